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
