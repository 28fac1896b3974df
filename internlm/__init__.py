"""Drop-in import surface for code written against InternEvo's ``internlm`` package.

``import internlm`` / ``from internlm.core.context import global_context as gpc`` / ``from internlm.train import
initialize_model`` ... resolve to the corresponding ``internevo_b200`` modules (SURVEY Appendix B lists the public
surface).  Nothing is re-implemented here: a meta-path finder aliases module paths.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_PREFIX = __name__ + "."
_MAP = {  # longest prefix wins
    "internlm.model.losses.ce_loss": "internevo_b200.models.losses",
    "internlm.model.losses": "internevo_b200.models.losses",
    "internlm.model.moe.base_layer": "internevo_b200.models.moe",
    "internlm.model.moe.experts": "internevo_b200.models.moe",
    "internlm.model.moe.gshard_layer": "internevo_b200.models.moe",
    "internlm.model.moe.megablock.megablock_moe": "internevo_b200.models.moe",
    "internlm.model.moe.megablock.megablock_dmoe": "internevo_b200.models.moe",
    "internlm.model.moe.megablock.mlp": "internevo_b200.models.moe",
    "internlm.model.moe.megablock.utils": "internevo_b200.models.moe",
    "internlm.model.moe.megablock": "internevo_b200.models.moe",
    "internlm.model.moe.moe": "internevo_b200.models.moe",
    "internlm.model.moe.utils": "internevo_b200.models.moe",
    "internlm.accelerator.abstract_accelerator": "internevo_b200.accelerator",
    "internlm.accelerator.cuda_accelerator": "internevo_b200.accelerator",
    "internlm.data.tokenized.dataset": "internevo_b200.data.datasets",
    "internlm.solver.optimizer.base_optimizer": "internevo_b200.solver.optimizer.base_optimizer",
    "internlm.model.metrics": "internevo_b200.models.metrics",
    "internlm.model.moe": "internevo_b200.models.moe",
    "internlm.model.ops.linear": "internevo_b200.parallel.linear",
    "internlm.model.ops.norm": "internevo_b200.ops.norm",
    "internlm.model.modules.embedding": "internevo_b200.models.modules",
    "internlm.model.modules.mlp": "internevo_b200.models.modules",
    "internlm.model.modules.multi_head_attention": "internevo_b200.models.modules",
    "internlm.model.utils": "internevo_b200.parallel.functional",
    "internlm.model": "internevo_b200.models",
    "internlm.solver.optimizer.hybrid_zero_optim": "internevo_b200.solver.optimizer.hybrid_zero_optim",
    "internlm.solver.beta2_scheduler": "internevo_b200.solver.schedulers.beta2_scheduler",
    "internlm.solver.lr_scheduler": "internevo_b200.solver.schedulers.lr_scheduler",
    "internlm.data.tokenized.dummy_dataset": "internevo_b200.data.datasets",
    "internlm.data.tokenized.packed_dataset": "internevo_b200.data.datasets",
    "internlm.data.tokenized.single_dataset": "internevo_b200.data.datasets",
    "internlm.data.tokenized.batch_sampler": "internevo_b200.data.batch_sampler",
    "internlm.data.tokenized.collaters": "internevo_b200.data.collaters",
    "internlm.data.utils": "internevo_b200.data.datasets",
    "internlm.core.context.parallel_context": "internevo_b200.core.context.parallel_context",
    "internlm.core.context.process_group_initializer": "internevo_b200.core.context.process_groups",
    "internlm.core.scheduler.pipeline_scheduler": "internevo_b200.core.scheduler.pipeline_scheduler",
    "internlm.core.scheduler.no_pipeline_scheduler": "internevo_b200.core.scheduler.no_pipeline_scheduler",
    "internlm.core.scheduler.base_scheduler": "internevo_b200.core.scheduler.base_scheduler",
    "internlm.utils.utils": "internevo_b200.core.context.config",
    "internlm": "internevo_b200",
}


def _target(name: str):
    best = None
    for k in _MAP:
        if name == k or name.startswith(k + "."):
            if best is None or len(k) > len(best):
                best = k
    if best is None:
        return None
    return _MAP[best] + name[len(best):]


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        mod = importlib.import_module(self.target)
        return mod

    def exec_module(self, module):
        pass


class _NamespaceLoader(importlib.abc.Loader):
    """Empty package for a reference path that has no counterpart of its own but contains aliased modules
    (``internlm.model.ops`` -> only ``internlm.model.ops.linear`` / ``.norm`` map to something)."""

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        module.__path__ = []


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if not fullname.startswith(_PREFIX):
            return None
        tgt = _target(fullname)
        if tgt is not None:
            try:
                importlib.import_module(tgt)
                return importlib.util.spec_from_loader(fullname, _AliasLoader(tgt), is_package=True)
            except ImportError:
                pass
        if any(k.startswith(fullname + ".") for k in _MAP):      # a parent of an aliased module
            return importlib.util.spec_from_loader(fullname, _NamespaceLoader(), is_package=True)
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

from internevo_b200.initialize import (  # noqa: E402
    get_default_parser,
    initialize_distributed_env,
    initialize_trainer,
    launch_from_slurm,
    launch_from_torch,
)

__all__ = ["get_default_parser", "initialize_trainer", "launch_from_slurm", "launch_from_torch",
           "initialize_distributed_env"]
