"""Global parallel context (``gpc``): config, rank bookkeeping, process groups, seeds and device binding.

Same public surface as the reference singleton (``internlm/core/context/parallel_context.py:130-673``), rebuilt around
the pure layout functions in ``process_groups.py``.  Differences that matter on a B200 node:

* backend is NCCL when a GPU is present, gloo otherwise (the reference asserts an accelerator and cannot run the
  ``configs/demo.py`` CPU plumbing config, reference ``parallel_context.py:627``);
* one process per GPU, device bound from ``LOCAL_RANK``; groups of size 1 never create a communicator;
* every multi-rank group may own a symmetric peer-memory heap (``internevo_b200/parallel/symm.py``) that the fused
  compute+collective kernels use instead of NCCL.
"""
from __future__ import annotations

import os
import random
import socket
from datetime import timedelta
from typing import Dict, List, Optional, Union

import numpy as np
import torch
import torch.distributed as dist

from internevo_b200.utils.logger import get_logger

from . import random as prng
from .config import Config, load_config
from .process_groups import ParallelMode, ParallelSizes, group_rank_lists, modes_to_build

logger = get_logger(__file__)

# parameter attribute names used to route gradient reduction / norm accounting (reference context/__init__.py)
IS_REPLICA_ZERO_PARALLEL = "is_replica_zero_parallel"
IS_TENSOR_ZERO_PARALLEL = "is_tensor_zero_parallel"
IS_TENSOR_DATA_PARALLEL = "is_tensor_data_parallel"
IS_TENSOR_EXPERT_DATA_PARALLEL = "is_tensor_expert_data_parallel"
IS_WEIGHT_ZERO_PARALLEL = "is_weight_zero_parallel"

LLM_NCCL_TIMEOUT = timedelta(seconds=int(os.getenv("NCCL_TIMEOUT", "1800")))


def _size_of(entry, default=1) -> int:
    if entry is None:
        return default
    if isinstance(entry, int):
        return entry
    return int(entry.get("size", default))


class ParallelContext:
    """Singleton; use the module-level ``global_context``."""

    def __init__(self):
        self._reset()

    def _reset(self):
        self._global_ranks: Dict[ParallelMode, int] = {}
        self._local_ranks: Dict[ParallelMode, int] = {}
        self._world_sizes: Dict[ParallelMode, int] = {}
        self._groups: Dict[ParallelMode, Optional[dist.ProcessGroup]] = {}
        self._cpu_groups: Dict[ParallelMode, Optional[dist.ProcessGroup]] = {}
        self._ranks_in_group: Dict[ParallelMode, List[int]] = {}
        self._config: Optional[Config] = None
        self.sizes: Optional[ParallelSizes] = None
        self.world_size = 1
        self.data_parallel_size = 1
        self.pipeline_parallel_size = 1
        self.tensor_parallel_size = 1
        self.weight_parallel_size = 1
        self.weight_data_parallel_size = 1
        self.sequence_parallel_size = 1
        self.zero1_parallel_size = -1
        self.nettest_parallel_size = 1
        self.expert_parallel_size = -1
        self.num_processes_on_current_node = -1
        self.virtual_pipeline_parallel_size = None
        self.virtual_pipeline_parallel_rank = None
        self._is_evaluating = False
        self.is_forward = True
        self.pipeline_bwd_group = None

    # ------------------------------------------------------------------ config
    @property
    def config(self) -> Config:
        return self._config

    def load_config(self, config: Union[dict, str, Config]):
        self._config = load_config(config)

    def set_config(self, config: Config):
        self._config = config

    # shorthands of the data section used all over training scripts (reference ``parallel_context.py:167-181``)
    @property
    def micro_bsz(self) -> int:
        return self._config.data.micro_bsz

    @property
    def micro_num(self) -> int:
        return self._config.data.micro_num

    @property
    def grad_accum_num(self) -> int:
        return self._config.data.get("gradient_accumulation", self._config.data.micro_num)

    @property
    def expert_parallel_group_names(self) -> list:
        """Names of the optimizer parameter groups that hold experts (one per expert-parallel size in use)."""
        n = self._config.model.get("num_experts", 1) if self._config is not None and "model" in self._config else 1
        return [f"moe_ep_size_{self.expert_parallel_size}"] if n > 1 else []

    # ------------------------------------------------------------------ queries
    @staticmethod
    def _check_mode(mode):
        assert isinstance(mode, ParallelMode), f"expected ParallelMode, got {type(mode)}"

    def get_global_rank(self) -> int:
        return self._global_ranks.get(ParallelMode.GLOBAL, 0)

    def get_local_rank(self, mode: ParallelMode) -> int:
        self._check_mode(mode)
        return self._local_ranks.get(mode, 0)

    def get_next_local_rank(self, mode):
        return (self.get_local_rank(mode) + 1) % self.get_world_size(mode)

    def get_prev_local_rank(self, mode):
        return (self.get_local_rank(mode) - 1) % self.get_world_size(mode)

    def get_next_global_rank(self, mode):
        ranks = self.get_ranks_in_group(mode)
        return ranks[(self.get_local_rank(mode) + 1) % len(ranks)]

    def get_prev_global_rank(self, mode):
        ranks = self.get_ranks_in_group(mode)
        return ranks[(self.get_local_rank(mode) - 1) % len(ranks)]

    def is_using_parallel_mode(self, mode) -> bool:
        return self.is_initialized(mode) and self.get_world_size(mode) > 1

    def is_first_rank(self, mode) -> bool:
        return self.get_local_rank(mode) == 0

    def is_last_rank(self, mode) -> bool:
        return self.get_local_rank(mode) == self.get_world_size(mode) - 1

    def is_rank_for_log(self) -> bool:
        """tp0 ∧ wp0 ∧ dp0 ∧ wdp0 ∧ last pipeline stage (reference ``parallel_context.py:284-293``)."""
        return (
            self.is_first_rank(ParallelMode.TENSOR)
            and self.is_first_rank(ParallelMode.WEIGHT)
            and self.is_first_rank(ParallelMode.DATA)
            and self.is_first_rank(ParallelMode.WEIGHT_DATA)
            and self.is_last_rank(ParallelMode.PIPELINE)
        )

    def is_last_rank_for_log(self) -> bool:
        return (
            self.is_last_rank(ParallelMode.TENSOR)
            and self.is_last_rank(ParallelMode.WEIGHT)
            and self.is_last_rank(ParallelMode.DATA)
            and self.is_last_rank(ParallelMode.WEIGHT_DATA)
            and self.is_last_rank(ParallelMode.PIPELINE)
        )

    def is_pipeline_first_stage(self, ignore_virtual=False) -> bool:
        if not ignore_virtual and self.virtual_pipeline_parallel_size is not None:
            if self.virtual_pipeline_parallel_rank != 0:
                return False
        return self.is_first_rank(ParallelMode.PIPELINE)

    def is_pipeline_last_stage(self, ignore_virtual=False) -> bool:
        if not ignore_virtual and self.virtual_pipeline_parallel_size is not None:
            if self.virtual_pipeline_parallel_rank != self.virtual_pipeline_parallel_size - 1:
                return False
        return self.is_last_rank(ParallelMode.PIPELINE)

    def is_no_pp_or_last_stage(self) -> bool:
        """Does this RANK produce the loss / logits?  Asked by code outside the schedulers (metrics, validation), after a schedule
        has reset the virtual (chunk) rank - so the chunk is ignored: with interleaving the last pipeline rank owns the last chunk."""
        return not self.is_initialized(ParallelMode.PIPELINE) or self.is_pipeline_last_stage(ignore_virtual=True)

    def get_world_size(self, mode) -> int:
        self._check_mode(mode)
        return self._world_sizes.get(mode, 1)

    def get_group(self, mode) -> Optional[dist.ProcessGroup]:
        self._check_mode(mode)
        return self._groups.get(mode, None)

    def get_cpu_group(self, mode):
        return self._cpu_groups.get(mode, None)

    def get_ranks_in_group(self, mode) -> List[int]:
        self._check_mode(mode)
        return self._ranks_in_group.get(mode, [self.get_global_rank()])

    def is_initialized(self, mode) -> bool:
        return mode in self._world_sizes

    def get_model_parallel_size(self):
        return self.tensor_parallel_size * self.pipeline_parallel_size

    @property
    def is_evaluating(self):
        return self._is_evaluating

    @is_evaluating.setter
    def is_evaluating(self, v):
        self._is_evaluating = v

    @property
    def is_distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized()

    # ------------------------------------------------------------------ init
    def _register(self, mode, local_rank, world_size, group, ranks, cpu_group=None):
        self._local_ranks[mode] = local_rank
        self._world_sizes[mode] = world_size
        self._groups[mode] = group
        self._cpu_groups[mode] = cpu_group
        self._ranks_in_group[mode] = ranks

    def init_global_dist(self, rank: int, world_size: int, backend: Optional[str], host: str, port: int,
                         use_cpu: bool = False):
        """``dist.init_process_group`` over tcp; registers the GLOBAL mode (reference ``:372-404``)."""
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if world_size > 1 or os.environ.get("INTERNEVO_FORCE_DIST", "0") == "1":
            if not dist.is_initialized():
                dist.init_process_group(
                    rank=rank, world_size=world_size, backend=backend, init_method=f"tcp://[{host}]:{port}"
                    if ":" in host else f"tcp://{host}:{port}", timeout=LLM_NCCL_TIMEOUT,
                )
            group = dist.group.WORLD
            # which ranks share this rank's node: the peer-memory (NVLink) back-ends only serve groups that stay inside one
            from internevo_b200.parallel import symm

            symm.exchange_node_ids()
        else:
            group = None  # single process: no communicator at all
        self._global_ranks[ParallelMode.GLOBAL] = rank
        self._register(ParallelMode.GLOBAL, rank, world_size, group, list(range(world_size)))
        self.world_size = world_size
        self.backend = backend

    def _sizes_from_config(self) -> ParallelSizes:
        pc = self._config.get("parallel", None)
        if pc is None:
            self._config._add_item("parallel", Config())
            pc = self._config.parallel
        if "zero1" not in pc:
            pc._add_item("zero1", dict(size=-1, fsdp=False))
        if "pipeline" not in pc:
            pc._add_item("pipeline", dict(size=1, interleaved_overlap=False))
        if "tensor" not in pc:
            pc._add_item("tensor", dict(size=1, mode="mtp"))
        if "weight" not in pc:
            pc._add_item("weight", dict(size=1, overlap=False, memory_pool=False))
        tmode = pc["tensor"].get("mode", "mtp") if isinstance(pc["tensor"], dict) else "mtp"
        model_cfg = self._config.get("model", {}) or {}
        zero1 = pc["zero1"]
        return ParallelSizes(
            world=self.world_size,
            pipeline=_size_of(pc["pipeline"]),
            tensor=_size_of(pc["tensor"]),
            weight=_size_of(pc["weight"]),
            zero1=_size_of(zero1, -1),
            num_experts=model_cfg.get("num_experts", 1),
            isp=tmode == "isp",
            fsdp=bool(zero1.get("fsdp", False)) if isinstance(zero1, dict) else False,
        )

    def init_parallel_groups(self):
        """Creates every group this configuration needs; all ranks call ``new_group`` in the same order."""
        rank = self.get_global_rank()
        s = self._sizes_from_config()
        self.sizes = s
        self.pipeline_parallel_size, self.tensor_parallel_size = s.pipeline, s.tensor
        self.weight_parallel_size, self.weight_data_parallel_size = s.weight, s.weight_data
        self.sequence_parallel_size, self.data_parallel_size = s.sequence, s.data
        self.zero1_parallel_size, self.expert_parallel_size = s.zero1, s.expert
        self.nettest_parallel_size = s.nettest
        pc = self._config.parallel
        tmode = pc["tensor"].get("mode", "mtp") if isinstance(pc["tensor"], dict) else "mtp"
        if "sequence_parallel" not in pc:
            pc._add_item("sequence_parallel", tmode != "mtp")
        if tmode == "mtp":
            pc["sequence_parallel"] = False
        self.check_sanity()

        gqa = bool(pc.get("gqa", False))
        for mode in modes_to_build(s, gqa):
            for ranks in group_rank_lists(mode, s):
                group = None
                if len(ranks) > 1 and self.is_distributed:
                    if len(ranks) == self.world_size and ranks == list(range(self.world_size)):
                        group = dist.group.WORLD
                    else:
                        group = dist.new_group(ranks, timeout=LLM_NCCL_TIMEOUT)
                if rank in ranks:
                    self._register(mode, ranks.index(rank), len(ranks), group, ranks)
                if mode is ParallelMode.PIPELINE and len(ranks) > 1 and self.is_distributed:
                    # second communicator for the backward (gradient) direction: with it, activations and gradients
                    # exchanged between the same pair of stages can never be mis-matched (matters for pp == 2, where
                    # the previous and the next stage are the same rank)
                    bwd = dist.new_group(ranks, timeout=LLM_NCCL_TIMEOUT)
                    if rank in ranks:
                        self.pipeline_bwd_group = bwd
        if ParallelMode.PIPELINE not in self._world_sizes and s.pipeline == 1:
            pass  # queries fall back to size 1 / rank 0

    def check_sanity(self):
        s = self.sizes
        assert self.world_size == s.data * s.pipeline * s.tensor, (
            f"world size {self.world_size} != dp {s.data} x pp {s.pipeline} x tp {s.tensor}"
        )
        assert self.world_size == s.weight_data * s.pipeline * s.weight, (
            f"world size {self.world_size} != wdp {s.weight_data} x pp {s.pipeline} x wp {s.weight}"
        )
        assert self.zero1_parallel_size > 0

    def detect_num_processes_on_current_node(self):
        hostname = socket.gethostname()
        if self.is_distributed and self.world_size > 1:
            names = [None] * self.world_size
            dist.all_gather_object(names, hostname)
            self.num_processes_on_current_node = sum(1 for n in names if n == hostname)
        else:
            self.num_processes_on_current_node = 1

    def set_device(self, device_ordinal: int = None):
        if not torch.cuda.is_available():
            return
        global_rank = self.get_global_rank()
        if device_ordinal is None:
            device_ordinal = global_rank % torch.cuda.device_count()
        torch.cuda.set_device(device_ordinal)
        logger.info(f"process rank {global_rank} is bound to cuda:{device_ordinal}")

    def set_seed(self, seed: int, dpseed_with_tpoffset: bool = False):
        """Seeds python / numpy / torch and registers the per-mode device RNG streams (reference ``:615-664``)."""
        pipeline_offset = self._local_ranks.get(ParallelMode.PIPELINE, 0)
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        prng.reset_seeds()
        dp_seed = seed + (pipeline_offset * 1024 if dpseed_with_tpoffset else 0)
        prng.add_seed(ParallelMode.DATA, dp_seed)
        prng.add_seed(ParallelMode.WEIGHT_DATA, dp_seed)
        prng.add_seed(ParallelMode.DUMMY, dp_seed)
        if self.is_initialized(ParallelMode.TENSOR):
            prng.add_seed(ParallelMode.TENSOR, seed + self.get_local_rank(ParallelMode.TENSOR) + pipeline_offset * 1024)
        if self.is_initialized(ParallelMode.WEIGHT):
            prng.add_seed(ParallelMode.WEIGHT, seed + self.get_local_rank(ParallelMode.WEIGHT) + pipeline_offset * 1024)
        prng.set_mode(ParallelMode.DUMMY)
        if self.is_using_parallel_mode(ParallelMode.TENSOR):
            prng.set_mode(ParallelMode.TENSOR)
        if self.is_using_parallel_mode(ParallelMode.WEIGHT):
            prng.set_mode(ParallelMode.WEIGHT)

    def set_virtual_pipeline_parallel_size(self, size):
        self.virtual_pipeline_parallel_size = size

    def set_virtual_pipeline_parallel_rank(self, rank):
        self.virtual_pipeline_parallel_rank = rank

    def destroy(self, graceful: bool = True):
        """Tear the groups down.  ``graceful=False`` is the error path: this rank is leaving because of an exception while its
        peers may sit in a collective that will never complete - no barrier, no collective teardown, just forget the groups
        (the launcher kills the remaining ranks when this process exits non-zero)."""
        # peer-memory back-ends cache symmetric buffers / flag epochs keyed by process group: drop them with the groups so a
        # later initialisation in the same process (another layout, another test) starts from a clean heap
        try:
            from internevo_b200.parallel import reset_caches

            reset_caches()
        except Exception:  # pragma: no cover - teardown must not raise
            pass
        if self.is_distributed and graceful:
            try:
                dist.barrier()
            except Exception:  # pragma: no cover
                pass
            dist.destroy_process_group()
        self._reset()


global_context = ParallelContext()
