"""Differentiable collectives used by the tensor/sequence/weight-parallel layers (NCCL path).

These are the *baseline* implementations of the N1–N10/N20 collective sites of the reference
(``internlm/model/utils.py:25-217``); the fused peer-memory kernels in ``parallel/fused.py`` replace the hot ones and
are validated against these.
"""
from __future__ import annotations


import torch
import torch.distributed as dist


def _ws(group) -> int:
    return 1 if group is None else dist.get_world_size(group)


def all_reduce_raw(x: torch.Tensor, group, async_op: bool = False, op=dist.ReduceOp.SUM):
    if _ws(group) <= 1:
        return x, None
    x = x.contiguous()
    h = dist.all_reduce(x, op=op, group=group, async_op=async_op)
    return x, h


def all_gather_raw(x: torch.Tensor, group, async_op: bool = False, gather_dim: int = 0):
    """Gather along ``gather_dim`` (implemented on dim 0 + permute so NCCL writes one contiguous buffer)."""
    ws = _ws(group)
    if ws <= 1:
        return x, None
    x = x.contiguous()
    out = torch.empty(ws * x.shape[0], *x.shape[1:], dtype=x.dtype, device=x.device)
    h = dist.all_gather_into_tensor(out, x, group=group, async_op=async_op)
    if gather_dim != 0:
        assert not async_op, "async gather only on dim 0"
        out = torch.cat(out.chunk(ws, dim=0), dim=gather_dim)
    return out, h


def reduce_scatter_raw(x: torch.Tensor, group, async_op: bool = False, op=dist.ReduceOp.SUM):
    """Reduce-scatter along dim 0."""
    ws = _ws(group)
    if ws <= 1:
        return x, None
    assert x.shape[0] % ws == 0
    x = x.contiguous()
    out = torch.empty(x.shape[0] // ws, *x.shape[1:], dtype=x.dtype, device=x.device)
    h = dist.reduce_scatter_tensor(out, x, op=op, group=group, async_op=async_op)
    return out, h


class _CopyToGroup(torch.autograd.Function):
    """identity fwd / all-reduce bwd (input of a column-parallel region)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, g):
        g, _ = all_reduce_raw(g.contiguous(), ctx.group)
        return g, None


class _ReduceFromGroup(torch.autograd.Function):
    """all-reduce fwd / identity bwd (output of a row-parallel region)."""

    @staticmethod
    def forward(ctx, x, group):
        x, _ = all_reduce_raw(x, group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


class _GatherForwardSplitBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, dim):
        ctx.group, ctx.dim = group, dim
        ws = _ws(group)
        if ws <= 1:
            return x
        parts = [torch.empty_like(x) for _ in range(ws)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=dim)

    @staticmethod
    def backward(ctx, g):
        ws = _ws(ctx.group)
        if ws <= 1:
            return g, None, None
        return g.chunk(ws, dim=ctx.dim)[dist.get_rank(ctx.group)].contiguous(), None, None


class _SplitForwardGatherBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, dim):
        ctx.group, ctx.dim = group, dim
        ws = _ws(group)
        if ws <= 1:
            return x
        return x.chunk(ws, dim=dim)[dist.get_rank(group)].contiguous()

    @staticmethod
    def backward(ctx, g):
        ws = _ws(ctx.group)
        if ws <= 1:
            return g, None, None
        parts = [torch.empty_like(g) for _ in range(ws)]
        dist.all_gather(parts, g.contiguous(), group=ctx.group)
        return torch.cat(parts, dim=ctx.dim), None, None


class _AllGatherSeq(torch.autograd.Function):
    """all-gather along dim 0 fwd / reduce-scatter bwd (sequence-parallel entry)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        out, _ = all_gather_raw(x, group)
        return out

    @staticmethod
    def backward(ctx, g):
        out, _ = reduce_scatter_raw(g, ctx.group)
        return out, None


class _ReduceScatterSeq(torch.autograd.Function):
    """reduce-scatter along dim 0 fwd / all-gather bwd (sequence-parallel exit)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        out, _ = reduce_scatter_raw(x, group)
        return out

    @staticmethod
    def backward(ctx, g):
        out, _ = all_gather_raw(g, ctx.group)
        return out, None


def copy_to_group(x, group):
    return _CopyToGroup.apply(x, group) if _ws(group) > 1 else x


def reduce_from_group(x, group):
    return _ReduceFromGroup.apply(x, group) if _ws(group) > 1 else x


def gather_forward_split_backward(x, group, dim: int = -1):
    return _GatherForwardSplitBackward.apply(x, group, dim) if _ws(group) > 1 else x


def split_forward_gather_backward(x, group, dim: int = 0):
    return _SplitForwardGatherBackward.apply(x, group, dim) if _ws(group) > 1 else x


def all_gather_seq(x, group):
    return _AllGatherSeq.apply(x, group) if _ws(group) > 1 else x


def reduce_scatter_seq(x, group):
    return _ReduceScatterSeq.apply(x, group) if _ws(group) > 1 else x


class _SeqAllToAll(torch.autograd.Function):
    """Ulysses transpose: scatter ``scatter_dim`` / gather ``gather_dim`` over the sequence-parallel group
    (reference ``internlm/model/modules/multi_head_attention.py:27-53``), one ``all_to_all_single`` instead of the
    reference's list API."""

    @staticmethod
    def forward(ctx, x, group, scatter_dim, gather_dim):
        ctx.group, ctx.scatter_dim, ctx.gather_dim = group, scatter_dim, gather_dim
        ws = _ws(group)
        if ws <= 1:
            return x
        parts = [t.contiguous() for t in x.chunk(ws, dim=scatter_dim)]
        inp = torch.stack(parts, 0)
        out = torch.empty_like(inp)
        dist.all_to_all_single(out, inp, group=group)
        return torch.cat(list(out.unbind(0)), dim=gather_dim)

    @staticmethod
    def backward(ctx, g):
        return _SeqAllToAll.apply(g, ctx.group, ctx.gather_dim, ctx.scatter_dim), None, None, None


def seq_all_to_all(x, group, scatter_dim: int, gather_dim: int):
    return _SeqAllToAll.apply(x, group, scatter_dim, gather_dim) if _ws(group) > 1 else x


def try_import_RMSNorm():
    """The RMSNorm class to use (reference ``model/utils.py:662-675`` picks apex or a torch fall-back; here it is always the
    fused residual-add + RMSNorm kernel wrapper, which degrades to plain PyTorch on CPU)."""
    from internevo_b200.ops.norm import RMSNorm

    return RMSNorm


def Silu(w1_o, w2_o):
    """``silu(w1_o) * w2_o`` — one fused kernel on the GPU (reference ``model/utils.py:684-688`` jit-scripts it)."""
    from internevo_b200.ops import silu_mul

    return silu_mul(w1_o, w2_o)


def is_moe_param(param) -> bool:
    return getattr(param, "is_expert", False)


# The functional dense API (``fused_dense_func`` & co., reference ``model/utils.py:31-69,220-660``) lives in ``parallel/dense.py``
# which needs ``parallel/linear.py`` which imports this module: resolved on first access.
_DENSE_NAMES = ("ReduceScatterFunc", "AllReduceFunc", "reduce_scatter", "all_reduce", "linear_bias_wgrad_torch",
                "FusedDenseFunc", "MegatronFusedDenseFunc", "ISPFusedDenseFunc", "fused_dense_func",
                "megatron_fused_dense_func", "isp_fused_dense_func")


def __getattr__(name):
    if name in _DENSE_NAMES:
        from . import dense

        return getattr(dense, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
