"""MegaBlocks-style functional API over the grouped tcgen05 GEMM.

The reference's dropless MoE multiplies a token matrix with a *block-diagonal* weight structure through the external
``stk`` / ``megablocks`` kernels: ``sdd`` (dense x dense -> sparse), ``dsd`` (sparse x dense -> dense), with tensor- or
weight-parallel autograd functions around them (``internlm/model/moe/megablock/utils.py:16-370``) and two expert MLP modules
built from those (``megablock/mlp.py:15-73``).  For a dropless MoE the sparse matrix is block diagonal - expert ``g`` owns
the token rows ``offsets[g] : offsets[g + 1]`` and the feature columns ``g * ffn : (g + 1) * ffn`` - so it is stored here as
the dense ``[rows, ffn]`` panel of its non-zero blocks plus the row offsets (``Topology``), and both products are ONE launch of
the grouped GEMM (``ops/grouped.py``: rows grouped along M, one weight tensor map per expert read from device memory; wgrad
grouped along K).  Nothing in here needs a host-side read of the routing result.

Parallel modes (``parallel_mode`` of ``sdd_nt`` / ``dsd_nn``):

* ``"tensor"``: every rank holds ``ffn / tp`` columns of every expert; ``dsd`` all-reduces its output, ``sdd`` all-reduces dX.
* ``"weight"``: every rank holds ``1 / wp`` of the rows of the stacked ``[E * ffn, hidden]`` weight; it is all-gathered for the
  product and its gradient reduce-scattered in fp32, pre-scaled by ``1 / wp``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.ops import grouped as _g
from internevo_b200.parallel.functional import Silu


def _ws(group) -> int:
    return 1 if group is None else dist.get_world_size(group)


class Topology:
    """Block-diagonal structure of one dropless MoE activation: ``offsets`` int32 ``[E + 1]`` (multiples of 128, on the
    device of the activations), ``ffn`` features per expert on this rank.  Stands where the reference passes a ``stk.Matrix``
    used only for its sparsity pattern (``megablock_dmoe.py:72-160`` builds it from the padded bins)."""

    blocking = _g.ALIGN

    def __init__(self, offsets: torch.Tensor, ffn: int, rows: Optional[int] = None) -> None:
        self.offsets = offsets.to(torch.int32)
        self.ffn = int(ffn)
        self.rows = rows

    @classmethod
    def from_counts(cls, tokens_per_expert: torch.Tensor, ffn: int, rows: Optional[int] = None) -> "Topology":
        return cls(_g.aligned_offsets(tokens_per_expert), ffn, rows)

    @property
    def num_experts(self) -> int:
        return self.offsets.numel() - 1

    def size(self):
        return (self.rows, self.num_experts * self.ffn)

    shape = property(size)


class BlockDiagonal:
    """The non-zero blocks of a ``[rows, E * ffn]`` block-diagonal matrix as one dense ``[rows, ffn]`` panel."""

    def __init__(self, data: torch.Tensor, topo: Topology) -> None:
        self.data, self.topo = data, topo

    def size(self):
        return (self.data.shape[0], self.topo.num_experts * self.topo.ffn)

    def to_dense(self) -> torch.Tensor:
        """The full matrix (tests / debugging; one host read of the offsets)."""
        E, f = self.topo.num_experts, self.topo.ffn
        out = self.data.new_zeros(self.data.shape[0], E * f)
        b = self.topo.offsets.tolist()
        for g in range(E):
            out[b[g]: b[g + 1], g * f: (g + 1) * f] = self.data[b[g]: b[g + 1]]
        return out


def _experts(w: torch.Tensor, E: int) -> List[torch.Tensor]:
    """Stacked ``[E * n, k]`` weight as ``E`` views ``[n, k]`` (each one contiguous: one tensor map per expert)."""
    assert w.dim() == 2 and w.shape[0] % E == 0, (tuple(w.shape), E)
    return list(w.view(E, w.shape[0] // E, w.shape[1]).unbind(0))


def _stacked_wgrad(dy: torch.Tensor, x: torch.Tensor, offsets: torch.Tensor, like: Sequence[torch.Tensor]) -> torch.Tensor:
    """``cat_g(dy_g^T @ x_g)`` as one ``[E * n, k]`` tensor."""
    parts = _g.grouped_wgrad(dy, x, offsets, [w.detach() for w in like])      # detached views carry no gradient arena
    return torch.cat([p.reshape(like[0].shape) for p in parts], 0)


def _gather_weights(w: torch.Tensor, group, async_op: bool = False):
    ws = _ws(group)
    if ws <= 1:
        return w, None
    full = torch.empty(w.shape[0] * ws, *w.shape[1:], dtype=w.dtype, device=w.device)
    h = dist.all_gather_into_tensor(full, w.contiguous(), group=group, async_op=async_op)
    return full, (h if async_op else None)


def _scaled_reduce_scatter(full_dw: torch.Tensor, group, async_op: bool = False):
    """fp32 reduce-scatter of the gathered-weight gradient, pre-scaled by ``1 / world`` (an average over the group)."""
    ws = _ws(group)
    if ws <= 1:
        return full_dw, None
    full_dw = full_dw.float() / ws
    dw = torch.empty(full_dw.shape[0] // ws, *full_dw.shape[1:], dtype=torch.float32, device=full_dw.device)
    h = dist.reduce_scatter_tensor(dw, full_dw.contiguous(), group=group, async_op=async_op)
    return dw, (h if async_op else None)


def _check(x: torch.Tensor, w: torch.Tensor):
    if not x.is_contiguous() or not w.is_contiguous():
        raise ValueError("Expected contiguous 'x' and 'w'.")


# ---------------------------------------------------------------------------------------------------------------------
# sdd: [rows, k] x blockdiag([n, k]^T per expert) -> [rows, n] panel
# ---------------------------------------------------------------------------------------------------------------------
class TensorParallelSddNt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, topo: Topology, group=None):
        _check(x, w)
        ctx.group, ctx.topo = group, topo
        ctx.save_for_backward(x, w)
        return _g.grouped_matmul(x, _experts(w, topo.num_experts), topo.offsets)

    @staticmethod
    def backward(ctx, grad):
        x, w = ctx.saved_tensors
        topo, grad = ctx.topo, grad.contiguous()
        ws_list = _experts(w, topo.num_experts)
        dx = handle = None
        if ctx.needs_input_grad[0]:
            dx = _g.grouped_matmul(grad, ws_list, topo.offsets, b_mn=True)
            if _ws(ctx.group) > 1:
                handle = dist.all_reduce(dx, group=ctx.group, async_op=True)      # overlaps the weight gradient
        dw = _stacked_wgrad(grad, x, topo.offsets, ws_list).to(w.dtype) if ctx.needs_input_grad[1] else None
        if handle is not None:
            handle.wait()
        return dx, dw, None, None


class WeightParallelSddNt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, topo: Topology, group=None):
        _check(x, w)
        ctx.group, ctx.topo = group, topo
        ctx.save_for_backward(x, w)
        full, _ = _gather_weights(w, group)
        return _g.grouped_matmul(x, _experts(full, topo.num_experts), topo.offsets)

    @staticmethod
    def backward(ctx, grad):
        x, w = ctx.saved_tensors
        topo, grad = ctx.topo, grad.contiguous()
        full, h_w = _gather_weights(w, ctx.group, async_op=True)                   # overlaps the weight-gradient GEMM
        E = topo.num_experts
        # (the views of the buffer being gathered only lend their shape / dtype to the weight-gradient GEMM)
        full_dw = _stacked_wgrad(grad, x, topo.offsets, _experts(full, E)) if ctx.needs_input_grad[1] else None
        if h_w is not None:
            h_w.wait()
        dw, h_dw = (None, None) if full_dw is None else _scaled_reduce_scatter(full_dw, ctx.group, async_op=True)
        dx = _g.grouped_matmul(grad, _experts(full, E), topo.offsets, b_mn=True) if ctx.needs_input_grad[0] else None
        if h_dw is not None:
            h_dw.wait()
        return dx, (None if dw is None else dw.to(w.dtype)), None, None


# ---------------------------------------------------------------------------------------------------------------------
# dsd: [rows, n] panel x blockdiag([n, k] per expert) -> [rows, k]
# ---------------------------------------------------------------------------------------------------------------------
class TensorParallelDsdNn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, topo: Topology, w, group=None):
        _check(data, w)
        ctx.group, ctx.topo = group, topo
        ctx.save_for_backward(data, w)
        out = _g.grouped_matmul(data, _experts(w, topo.num_experts), topo.offsets, b_mn=True)
        if _ws(group) > 1:
            dist.all_reduce(out, group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        data, w = ctx.saved_tensors
        topo, grad = ctx.topo, grad.contiguous()
        ws_list = _experts(w, topo.num_experts)
        dw = _stacked_wgrad(data, grad, topo.offsets, ws_list).to(w.dtype) if ctx.needs_input_grad[2] else None
        dx = _g.grouped_matmul(grad, ws_list, topo.offsets) if ctx.needs_input_grad[0] else None
        return dx, None, dw, None


class WeightParallelDsdNn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, topo: Topology, w, group=None):
        _check(data, w)
        ctx.group, ctx.topo = group, topo
        ctx.save_for_backward(data, w)
        full, _ = _gather_weights(w, group)
        return _g.grouped_matmul(data, _experts(full, topo.num_experts), topo.offsets, b_mn=True)

    @staticmethod
    def backward(ctx, grad):
        data, w = ctx.saved_tensors
        topo, grad = ctx.topo, grad.contiguous()
        E = topo.num_experts
        full, h_w = _gather_weights(w, ctx.group, async_op=True)
        full_dw = _stacked_wgrad(data, grad, topo.offsets, _experts(full, E)) if ctx.needs_input_grad[2] else None
        if h_w is not None:
            h_w.wait()
        dw, h_dw = (None, None) if full_dw is None else _scaled_reduce_scatter(full_dw, ctx.group, async_op=True)
        dx = _g.grouped_matmul(grad, _experts(full, E), topo.offsets) if ctx.needs_input_grad[0] else None
        if h_dw is not None:
            h_dw.wait()
        return dx, None, (None if dw is None else dw.to(w.dtype)), None


def sdd_nt(a: torch.Tensor, b: torch.Tensor, topo: Topology, group, parallel_mode: str) -> BlockDiagonal:
    """``a [rows, k]`` times the transposed per-expert blocks of ``b [E * n, k]`` -> block-diagonal ``[rows, E * n]``."""
    impl = WeightParallelSddNt if parallel_mode == "weight" else TensorParallelSddNt
    return BlockDiagonal(impl.apply(a, b, topo, group), topo)


def dsd_nn(a: BlockDiagonal, b: torch.Tensor, group, parallel_mode: str) -> torch.Tensor:
    """Block-diagonal ``a`` times the per-expert blocks of ``b [E * n, k]`` -> dense ``[rows, k]``."""
    impl = WeightParallelDsdNn if parallel_mode == "weight" else TensorParallelDsdNn
    return impl.apply(a.data, a.topo, b, group)


def act_fn(x1: BlockDiagonal, x2: BlockDiagonal, topo: Topology) -> BlockDiagonal:
    """``silu(x1) * x2`` on the non-zero blocks."""
    return BlockDiagonal(Silu(x1.data, x2.data), topo)


# ---------------------------------------------------------------------------------------------------------------------
# batched form for capacity-padded activations [E, C, k]
# ---------------------------------------------------------------------------------------------------------------------
def _bmm(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    E, C, k = x.shape
    if x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and C % _g.ALIGN == 0 and x.is_contiguous() \
            and w.is_contiguous():
        off = torch.arange(E + 1, device=x.device, dtype=torch.int32) * C
        return _g.grouped_matmul(x.view(E * C, k), list(w.unbind(0)), off, b_mn=True).view(E, C, w.shape[-1])
    return torch.bmm(x, w)


class TensorParallelBmm(torch.autograd.Function):
    """``[E, C, k] x [E, k, n]`` with the gradient of ``x`` all-reduced over ``group`` (column-parallel expert weights)."""

    @staticmethod
    def forward(ctx, x, w, group=None):
        _check(x, w)
        ctx.group = group
        ctx.save_for_backward(x, w)
        return _bmm(x, w)

    @staticmethod
    def backward(ctx, grad):
        x, w = ctx.saved_tensors
        grad = grad.contiguous()
        dx = handle = None
        if ctx.needs_input_grad[0]:
            dx = torch.bmm(grad, w.transpose(-2, -1))
            if _ws(ctx.group) > 1:
                handle = dist.all_reduce(dx, group=ctx.group, async_op=True)
        dw = torch.bmm(x.transpose(-2, -1), grad).to(w.dtype) if ctx.needs_input_grad[1] else None
        if handle is not None:
            handle.wait()
        return dx, dw, None


def tensor_parallel_bmm(x, w, group=None):
    return TensorParallelBmm.apply(x, w, group)


def promote_scalar(x: torch.Tensor) -> torch.Tensor:
    return x.view(1) if x.dim() == 0 else x


def check_megablock_installed() -> bool:
    """The dropless path needs no external package here: the grouped GEMM is part of this framework's extension."""
    return True


def check_stk_installed() -> bool:
    return True


def _tensor_group():
    return gpc.get_group(ParallelMode.TENSOR) if gpc.is_initialized(ParallelMode.TENSOR) and \
        gpc.get_world_size(ParallelMode.TENSOR) > 1 else None


class MegaBlockFeedForward(nn.Module):
    """SwiGLU experts on capacity-padded activations ``[E_local, C, in]``: ``w1`` / ``w3`` ``[E, in, hidden]`` (hidden sharded
    over the tensor group), ``w2`` ``[E, hidden, in]``; the output is all-reduced over the tensor group (reference
    ``megablock/mlp.py:15-44``, whose ``w2`` / ``w3`` shapes are swapped relative to their use - noted there as a TODO)."""

    def __init__(self, in_features: int, hidden_features: int, num_local_experts: int, device=None, dtype=None):
        super().__init__()
        kw = dict(device=device, dtype=dtype)
        self.w1 = nn.Parameter(torch.empty(num_local_experts, in_features, hidden_features, **kw))
        self.w3 = nn.Parameter(torch.empty(num_local_experts, in_features, hidden_features, **kw))
        self.w2 = nn.Parameter(torch.empty(num_local_experts, hidden_features, in_features, **kw))
        for p in (self.w1, self.w2, self.w3):
            nn.init.normal_(p, std=0.02)

    def forward(self, x):
        group = _tensor_group()
        h = Silu(tensor_parallel_bmm(x, self.w1, group), tensor_parallel_bmm(x, self.w3, group))
        out = tensor_parallel_bmm(h, self.w2)
        if group is not None:
            dist.all_reduce(out, group=group)
        return out


class MegaBlockGroupedFeedForward(nn.Module):
    """SwiGLU experts on a dropless, group-aligned row buffer: the three weights are the experts' matrices stacked along dim
    0, ``[E_local * ffn, in]`` (``hidden_features = E_local * ffn`` of THIS rank; reference ``megablock/mlp.py:47-73``)."""

    def __init__(self, in_features: int, hidden_features: int, parallel_mode: str = "tensor", device=None, dtype=None):
        super().__init__()
        kw = dict(device=device, dtype=dtype)
        self.w1 = nn.Parameter(torch.empty(hidden_features, in_features, **kw))
        self.w2 = nn.Parameter(torch.empty(hidden_features, in_features, **kw))
        self.w3 = nn.Parameter(torch.empty(hidden_features, in_features, **kw))
        for p in (self.w1, self.w2, self.w3):
            nn.init.normal_(p, std=0.02)
        self.parallel_mode = parallel_mode

    def _group(self):
        if self.parallel_mode == "weight":
            return gpc.get_group(ParallelMode.WEIGHT) if gpc.is_initialized(ParallelMode.WEIGHT) and \
                gpc.get_world_size(ParallelMode.WEIGHT) > 1 else None
        return _tensor_group()

    def forward(self, x, topo: Topology):
        group = self._group()
        w1_o = sdd_nt(x, self.w1, topo, group, self.parallel_mode)
        w3_o = sdd_nt(x, self.w3, topo, group, self.parallel_mode)
        return dsd_nn(act_fn(w1_o, w3_o, topo), self.w2, group, self.parallel_mode)


__all__ = ["Topology", "BlockDiagonal", "TensorParallelSddNt", "WeightParallelSddNt", "TensorParallelDsdNn",
           "WeightParallelDsdNn", "sdd_nt", "dsd_nn", "act_fn", "TensorParallelBmm", "tensor_parallel_bmm", "promote_scalar",
           "check_megablock_installed", "check_stk_installed", "MegaBlockFeedForward", "MegaBlockGroupedFeedForward"]
