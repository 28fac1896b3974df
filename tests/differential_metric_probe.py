"""``AccPerplex`` / ``LossWithTypeId`` on the same logits, labels and type ids, run against the reference and against this repository
(``internlm`` alias); see ``test_reference_differential_cpu.py``.

    python differential_metric_probe.py <root that provides `internlm`> <output json>
"""
import json, os, sys
root, dst = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
import torch, torch.distributed as dist
import internlm
import internlm.utils.common as common
from internlm.core.context import ParallelMode, global_context as gpc
from internlm.core.context.parallel_context import Config
cpu, orig = torch.device("cpu"), common.get_current_device
for mod in list(sys.modules.values()):
    if mod is not None and getattr(mod, "get_current_device", None) is orig:
        setattr(mod, "get_current_device", lambda: cpu)
dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:%d" % (20000 + os.getpid() % 20000))
gpc._config = Config(dict(model=dict(dtype=torch.float32, use_flash_attn=False, parallel_output=False), parallel=dict(sequence_parallel=False, tensor=dict(size=1, mode="mtp")), data=dict(micro_bsz=2, use_packed_dataset=True, seq_len=24), metric_dtype="fp32"))
try:
    for mode in ParallelMode:
        gpc._world_sizes[mode], gpc._local_ranks[mode], gpc._global_ranks[mode] = 1, 0, 0
        gpc._groups[mode], gpc._ranks_in_group[mode] = dist.group.WORLD, [0]
except Exception:
    pass
from internlm.accelerator import get_accelerator
from internlm.accelerator.abstract_accelerator import AcceleratorType
get_accelerator().get_accelerator_backend = lambda: AcceleratorType.OTHER      # no torch_scatter here: the plain scatter path
from internlm.model.metrics import AccPerplex, LossWithTypeId
torch.manual_seed(0)
types = ["cn", "code", "en"]
m = AccPerplex(device=cpu, tp_pg=dist.group.WORLD, dp_pg=dist.group.WORLD, dataset_types=types)
lt = LossWithTypeId(device=cpu, dp_pg=dist.group.WORLD, dataset_types=types)
for _ in range(3):
    logits = torch.randn(48, 32)
    labels = torch.randint(0, 32, (48,)); labels[::7] = -100
    tids = torch.randint(0, 3, (48,))
    m.set_current_type_ids(tids)
    m(logits.clone(), labels.clone())
    lt.update(logits.clone(), labels.clone(), tids)
res = m.get_metric()
res2 = lt.get_metric()
def plain(d): return {k: (float(v) if not isinstance(v, (int, float)) else v) for k, v in d.items()}
json.dump({"acc": plain(res), "loss": plain(res2)}, open(dst, "w"))
print("PROBE_OK", flush=True); os._exit(0)
