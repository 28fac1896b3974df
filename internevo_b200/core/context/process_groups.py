"""Parallel modes and the rank layout of every process-group kind.

The reference builds one ``Initializer_*`` class per group kind, each looping over ``dist.new_group``
(``internlm/core/context/process_group_initializer.py:118-934``).  Here the layout is a pure function
(``group_rank_lists``) that can be unit-tested without any distributed runtime; ``ParallelContext`` walks the result once
to create the communicators.  Rank order (fastest → slowest): tensor (or weight), data, pipeline.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Dict, List


class ParallelMode(Enum):
    GLOBAL = "global"
    DATA = "data"
    MODEL = "model"  # enum member only (reference never instantiates it)
    PIPELINE = "pipe"
    TENSOR = "tensor"
    ZERO1 = "zero1"
    NETTEST = "nettest"
    ZERO3_DP = "zero3_dp"
    EXPERT = "expert"
    EXPERT_DATA = "expert_data"
    DUMMY = "dummy"
    WEIGHT = "weight"
    WEIGHT_DATA = "weight_data"
    SEQUENCE = "sequence"
    GQA = "gqa"


@dataclass
class ParallelSizes:
    world: int
    pipeline: int = 1
    tensor: int = 1
    weight: int = 1
    zero1: int = -1
    num_experts: int = 1
    isp: bool = False
    nettest: int = 32
    fsdp: bool = False
    gqa_q_heads: int = 32
    gqa_kv_heads: int = 8

    def __post_init__(self):
        assert self.world % self.pipeline == 0, "world size must be divisible by pipeline size"
        self.per_stage = self.world // self.pipeline
        assert self.per_stage % self.tensor == 0, "ranks per pipeline stage must be divisible by tensor size"
        assert self.per_stage % self.weight == 0, "ranks per pipeline stage must be divisible by weight size"
        self.sequence = self.tensor
        self.data = max(1, self.per_stage // self.tensor)
        self.weight_data = max(1, self.per_stage // self.weight)
        base = self.weight_data if self.isp else self.data
        if self.zero1 == -1:
            self.zero1 = base
        self.zero1 = max(1, self.zero1)
        assert self.zero1 <= base and base % self.zero1 == 0, (
            f"zero1 size {self.zero1} must divide the {'weight-' if self.isp else ''}data parallel size {base}"
        )
        assert self.data % self.num_experts == 0 or self.num_experts % self.data == 0, "can not place the experts evenly"
        self.expert = min(self.data, self.num_experts)
        self.expert_data = self.data // self.expert


def group_rank_lists(mode: ParallelMode, s: ParallelSizes) -> List[List[int]]:
    """All rank groups of ``mode`` (every rank appears in exactly one group, except NETTEST tails)."""
    pp, per = s.pipeline, s.per_stage
    out: List[List[int]] = []
    if mode is ParallelMode.GLOBAL:
        return [list(range(s.world))]
    if mode is ParallelMode.PIPELINE:  # reference process_group_initializer.py:167-168
        return [[i + j * per for j in range(pp)] for i in range(per)]
    if mode is ParallelMode.TENSOR:  # :227-229 contiguous ranks
        return [[i * s.tensor + j for j in range(s.tensor)] for i in range(s.world // s.tensor)]
    if mode is ParallelMode.WEIGHT:  # :680-681
        return [[i * s.weight + j for j in range(s.weight)] for i in range(s.world // s.weight)]
    if mode is ParallelMode.DATA:  # :747-752
        return [[i * per + j + k * s.sequence for k in range(s.data)] for i in range(pp) for j in range(s.sequence)]
    if mode is ParallelMode.WEIGHT_DATA:  # :831-836
        return [[i * per + j + k * s.weight for k in range(s.weight_data)] for i in range(pp) for j in range(s.weight)]
    if mode is ParallelMode.ZERO1:
        inner = s.weight if s.isp else s.tensor  # :305-311 / :394-400
        block = inner * s.zero1
        for i in range(pp):
            for j in range(per // block):
                for k in range(inner):
                    out.append([i * per + j * block + k + m * inner for m in range(s.zero1)])
        return out
    if mode is ParallelMode.ZERO3_DP:  # ranks with the same position inside their zero1(fsdp) group
        z = s.zero1
        n = s.data // z
        for i in range(pp):
            for j in range(s.tensor):
                for k in range(z):
                    out.append([i * per + j + (k + m * z) * s.tensor for m in range(n)])
        return out
    if mode is ParallelMode.NETTEST:  # :448-452
        n = (s.world + s.nettest - 1) // s.nettest
        return [[r for r in range(i * s.nettest, (i + 1) * s.nettest) if r < s.world] for i in range(n)]
    if mode in (ParallelMode.EXPERT, ParallelMode.EXPERT_DATA):  # :493-524
        ep_groups, edp_groups = [], []
        for dp_ranks in group_rank_lists(ParallelMode.DATA, s):
            part = [dp_ranks[i: i + s.expert] for i in range(0, s.data, s.expert)]
            ep_groups.extend(part)
            edp_groups.extend([list(t) for t in zip(*part)])
        return ep_groups if mode is ParallelMode.EXPERT else edp_groups
    if mode is ParallelMode.GQA:  # :857-934 ranks of one TP group that share a kv head
        q_per_kv = s.gqa_q_heads // s.gqa_kv_heads
        rep = max(1, s.tensor // s.gqa_kv_heads) if s.tensor > s.gqa_kv_heads else 1
        rep = max(rep, 1)
        for tp_ranks in group_rank_lists(ParallelMode.TENSOR, s):
            for a in range(0, len(tp_ranks), rep):
                out.append(tp_ranks[a: a + rep])
        del q_per_kv
        return out
    raise ValueError(f"no rank layout for {mode}")


def modes_to_build(s: ParallelSizes, gqa: bool = False) -> List[ParallelMode]:
    """Creation order (must be identical on every rank); mirrors reference ``parallel_context.py:546-574``."""
    modes = []
    if gqa:
        modes.append(ParallelMode.GQA)
    modes += [ParallelMode.WEIGHT, ParallelMode.WEIGHT_DATA, ParallelMode.TENSOR, ParallelMode.DATA, ParallelMode.ZERO1]
    if s.fsdp:
        modes.append(ParallelMode.ZERO3_DP)
    modes.append(ParallelMode.NETTEST)
    if s.pipeline > 1:
        modes.append(ParallelMode.PIPELINE)
    if s.num_experts > 1:
        modes += [ParallelMode.EXPERT, ParallelMode.EXPERT_DATA]
    return modes


def layout_for_rank(rank: int, s: ParallelSizes, gqa: bool = False) -> Dict[ParallelMode, List[int]]:
    res = {}
    for mode in modes_to_build(s, gqa):
        for ranks in group_rank_lists(mode, s):
            if rank in ranks:
                res[mode] = ranks
    return res


# ---------------------------------------------------------------------------------------------------------------------
# ``Initializer_*`` facade: user code written against the reference instantiates one class per group kind and calls
# ``init_dist_group()`` → ``(local_rank, group_world_size, process_group, cpu_group, ranks_in_group, mode)`` (two such tuples
# for the expert kinds).  Here every class is the SAME few lines around ``group_rank_lists``; only the mode and, for the ISP /
# expert / GQA flavours, a detail of the sizes differ.
# ---------------------------------------------------------------------------------------------------------------------
class ProcessGroupInitializer:
    MODE: ParallelMode = ParallelMode.GLOBAL
    ISP = False

    def __init__(self, rank: int, world_size: int, weight_parallel_size: int, weight_data_parallel_size: int,
                 sequence_parallel_size: int, data_parallel_size: int, pipeline_parallel_size: int,
                 tensor_parallel_size: int, zero1_parallel_size: int, nettest_parallel_size: int,
                 expert_parallel_size: int):
        assert sequence_parallel_size == tensor_parallel_size, "the sequence group is the tensor group"
        self.rank, self.world_size = rank, world_size
        self.weight_parallel_size, self.weight_data_parallel_size = weight_parallel_size, weight_data_parallel_size
        self.sequence_parallel_size, self.data_parallel_size = sequence_parallel_size, data_parallel_size
        self.pipeline_parallel_size, self.tensor_parallel_size = pipeline_parallel_size, tensor_parallel_size
        self.zero1_parallel_size, self.nettest_parallel_size = zero1_parallel_size, nettest_parallel_size
        self.expert_parallel_size = expert_parallel_size

    def sizes(self) -> ParallelSizes:
        return ParallelSizes(world=self.world_size, pipeline=self.pipeline_parallel_size, tensor=self.tensor_parallel_size,
                             weight=self.weight_parallel_size, zero1=self.zero1_parallel_size,
                             num_experts=max(1, self.expert_parallel_size), isp=self.ISP,
                             nettest=self.nettest_parallel_size)

    def _build(self, mode: ParallelMode, use_cpu: bool):
        """Create every group of ``mode`` (all ranks must do so, in the same order) and return this rank's tuple."""
        import torch.distributed as dist

        mine = None
        for ranks in group_rank_lists(mode, self.sizes()):
            group = dist.new_group(ranks) if dist.is_initialized() and len(ranks) > 1 else None
            cpu = None
            if use_cpu and dist.is_initialized() and len(ranks) > 1:
                cpu = dist.new_group(ranks, backend="gloo") if dist.get_backend() != "gloo" else group
            if self.rank in ranks:
                mine = (ranks.index(self.rank), len(ranks), group, cpu, list(ranks), mode)
        assert mine is not None, f"rank {self.rank} is in no {mode} group"
        return mine

    def init_dist_group(self, use_cpu: bool = False):
        return self._build(self.MODE, use_cpu)


def _initializer(name: str, mode: ParallelMode, isp: bool = False, doc: str = ""):
    cls = type(name, (ProcessGroupInitializer,), {"MODE": mode, "ISP": isp, "__doc__": doc})
    return cls


Initializer_Pipeline = _initializer("Initializer_Pipeline", ParallelMode.PIPELINE, doc="ranks with the same position in every stage")
Initializer_Tensor = _initializer("Initializer_Tensor", ParallelMode.TENSOR, doc="contiguous ranks")
Initializer_Weight = _initializer("Initializer_Weight", ParallelMode.WEIGHT, isp=True, doc="contiguous ranks (ISP weight shards)")
Initializer_Data = _initializer("Initializer_Data", ParallelMode.DATA, doc="same tensor position, stride = tensor size")
Initializer_Weight_Data = _initializer("Initializer_Weight_Data", ParallelMode.WEIGHT_DATA, isp=True,
                                       doc="same weight position, stride = weight size")
Initializer_Zero1 = _initializer("Initializer_Zero1", ParallelMode.ZERO1, doc="ZeRO sub-groups inside the data group")
Initializer_Zero1_ISP = _initializer("Initializer_Zero1_ISP", ParallelMode.ZERO1, isp=True,
                                     doc="ZeRO sub-groups inside the weight-data group")
Initializer_Nettest = _initializer("Initializer_Nettest", ParallelMode.NETTEST, doc="blocks of `nettest` ranks")
Initializer_Zero3_dp = _initializer("Initializer_Zero3_dp", ParallelMode.ZERO3_DP, doc="same position inside the FSDP group")
Initializer_GQA = _initializer("Initializer_GQA", ParallelMode.GQA, doc="tensor ranks that share a kv head")


class Initializer_Expert_Data(ProcessGroupInitializer):
    """Expert and expert-data groups together (the reference returns both tuples from one call)."""

    def init_dist_group(self, use_cpu: bool = False):
        return [self._build(ParallelMode.EXPERT, use_cpu), self._build(ParallelMode.EXPERT_DATA, use_cpu)]
