"""Grouped tcgen05 GEMM for MoE experts: all local experts in ONE launch, problem sizes read from device memory.

The packed activation buffer holds the rows of expert ``g`` at ``offsets[g] .. offsets[g + 1]``; offsets are multiples of
128 and the padding rows are zero.  ``offsets`` is an int32 CUDA tensor produced by the routing arithmetic, so nothing here
(not the grid size, not a tensor shape) depends on a host-side read of the routing result - the per-expert Python loop and
its ``.tolist()`` sync are gone.  Forward and dgrad are "mode 1" of ``csrc/gemm_sm100.cu`` (rows grouped along M, every
expert's weight through its own tensor map in global memory), wgrad is "mode 2" (the contraction runs over the group's
rows, the result goes straight into each expert weight's gradient arena slot).

Replaces the block-sparse ``sdd`` / ``dsd`` products of MegaBlocks (reference ``internlm/model/moe/megablock/mlp.py:40-160``,
``megablock/utils.py:34-301``) and the per-expert loop of ``moe/experts.py:40-68``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import _lib
from .gemm import GEMM_ACCUMULATE, GEMM_SWIGLU, _bump
from .swiglu import swiglu_interleaved_bwd, swiglu_interleaved_fwd

ALIGN = 128

_bmaps: Dict[Tuple, torch.Tensor] = {}
_dptrs: Dict[Tuple, torch.Tensor] = {}


def aligned_offsets(counts: torch.Tensor, align: int = ALIGN) -> torch.Tensor:
    """``counts [G]`` rows per group → int32 ``[G + 1]`` start rows with every group start rounded up to ``align``
    (device arithmetic only)."""
    padded = (counts.to(torch.int64) + align - 1) // align * align
    off = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    off[1:] = padded.cumsum(0)
    return off.to(torch.int32)


def rows_capacity(n_rows_max: int, groups: int, align: int = ALIGN) -> int:
    """Static size of a packed buffer that can hold ``n_rows_max`` rows spread over ``groups`` aligned groups."""
    return (n_rows_max + groups * (align - 1) + align - 1) // align * align


def padding_rows(counts: torch.Tensor, offsets: torch.Tensor, dummy_row: int, align: int = ALIGN) -> torch.Tensor:
    """Indices of the (at most ``align - 1`` per group) padding rows between a group's last row and the next group's
    start, as a FIXED-size ``[G * align]`` tensor: entries that are not padding point at ``dummy_row`` (a row the caller
    reserves), so zeroing them is one ``index_fill_`` without a host read."""
    G = counts.numel()
    start = offsets[:-1].to(torch.int64) + counts.to(torch.int64)             # first padding row of every group
    end = offsets[1:].to(torch.int64)
    idx = start.unsqueeze(1) + torch.arange(align, device=counts.device).unsqueeze(0)
    return torch.where(idx < end.unsqueeze(1), idx, torch.full_like(idx, dummy_row)).reshape(G * align)


def _native(a: torch.Tensor, weights: Sequence[torch.Tensor]) -> bool:
    return _lib.use_native(a, weights[0]) and a.dtype == torch.bfloat16 and weights[0].dtype == torch.bfloat16


def _maps_for(weights: Sequence[torch.Tensor], b_mn: bool) -> torch.Tensor:
    key = (tuple(w.data_ptr() for w in weights), tuple(weights[0].shape), weights[0].stride(0), b_mn)
    m = _bmaps.get(key)
    if m is None:
        m = torch.ops.b200.grouped_b_maps(list(weights), b_mn)
        _bmaps[key] = m
    return m


def _ptrs_for(outs: Sequence[torch.Tensor]) -> torch.Tensor:
    key = tuple(o.data_ptr() for o in outs)
    t = _dptrs.get(key)
    if t is None:
        t = torch.tensor(key, dtype=torch.int64, device=outs[0].device)
        _dptrs[key] = t
    return t


def _bounds(offsets: torch.Tensor) -> List[int]:
    return [int(v) for v in offsets.tolist()]


def grouped_matmul(a: torch.Tensor, weights: Sequence[torch.Tensor], offsets: torch.Tensor, b_mn: bool = False) -> torch.Tensor:
    """``out[rows_g] = a[rows_g] @ W_g^T`` (``b_mn``: ``@ W_g``) for every group; rows outside the groups are not written."""
    N = weights[0].shape[1] if b_mn else weights[0].shape[0]
    out = torch.empty(a.shape[0], N, dtype=a.dtype, device=a.device)
    if _native(a, weights):
        torch.ops.b200.grouped_gemm(a, _maps_for(weights, b_mn), offsets, out, N, b_mn, 0, None)
        _bump()
        return out
    out.zero_()
    b = _bounds(offsets)
    for g, w in enumerate(weights):
        if b[g + 1] > b[g]:
            seg = a[b[g]: b[g + 1]].float()
            out[b[g]: b[g + 1]] = (seg @ (w.float() if b_mn else w.float().t())).to(out.dtype)
    return out


def grouped_matmul_swiglu(a: torch.Tensor, weights: Sequence[torch.Tensor], offsets: torch.Tensor):
    """``gu = a @ W13_g^T`` with interleaved (gate, up) rows; returns ``(gu, silu(gate) * up)`` - activation applied in the
    GEMM epilogue out of TMEM."""
    N = weights[0].shape[0]
    if _native(a, weights):
        gu = torch.empty(a.shape[0], N, dtype=a.dtype, device=a.device)
        h = torch.empty(a.shape[0], N // 2, dtype=a.dtype, device=a.device)
        torch.ops.b200.grouped_gemm(a, _maps_for(weights, False), offsets, gu, N, False, GEMM_SWIGLU, h)
        _bump()
        return gu, h
    gu = grouped_matmul(a, weights, offsets)
    return gu, swiglu_interleaved_fwd(gu)


def grouped_wgrad(dy: torch.Tensor, x: torch.Tensor, offsets: torch.Tensor, weights: Sequence[torch.Tensor]):
    """``dW_g = dy[rows_g]^T @ x[rows_g]``.  With gradient arenas (``weight.grad_buf``) the GEMM epilogue accumulates
    straight into them and ``None`` is returned for every weight; otherwise a list of gradients."""
    bufs = [getattr(w, "grad_buf", None) for w in weights]
    if all(b is not None for b in bufs) and _native(dy, weights):
        fresh = not any(getattr(w, "grad_ready", False) for w in weights)
        torch.ops.b200.grouped_wgrad(dy, x, offsets, _ptrs_for(bufs), bufs[0].stride(0), 0 if fresh else GEMM_ACCUMULATE)
        _bump()
        for w in weights:
            w.grad_ready = True
        for w in weights:
            hook = getattr(w, "grad_hook", None)
            if hook is not None:
                hook(w)
        return [None] * len(weights)
    if _native(dy, weights):
        outs = [torch.empty_like(w) for w in weights]
        torch.ops.b200.grouped_wgrad(dy, x, offsets, _ptrs_for(outs), outs[0].stride(0), 0)
        _bump()
        _dptrs.pop(tuple(o.data_ptr() for o in outs), None)   # temporaries: do not pin their addresses in the cache
        return outs
    b = _bounds(offsets)
    return [(dy[b[g]: b[g + 1]].float().t() @ x[b[g]: b[g + 1]].float()).to(w.dtype) for g, w in enumerate(weights)]


class _GroupedSwiGLUMLPFn(torch.autograd.Function):
    """All local experts' ``w2(silu(w1 x) * w3 x)`` on a packed, group-aligned row buffer: two grouped GEMMs forward
    (FC1 with the SwiGLU epilogue, FC2), five kernels backward (FC2 dgrad + wgrad, dSwiGLU, FC1 dgrad + wgrad)."""

    @staticmethod
    def forward(ctx, x, offsets, n_experts, *weights):
        w13, w2 = weights[:n_experts], weights[n_experts:]
        gu, h = grouped_matmul_swiglu(x, w13, offsets)
        y = grouped_matmul(h, w2, offsets)
        ctx.save_for_backward(x, offsets, gu, h)
        ctx.w13, ctx.w2 = w13, w2
        return y

    @staticmethod
    def backward(ctx, dy):
        x, offsets, gu, h = ctx.saved_tensors
        w13, w2 = ctx.w13, ctx.w2
        dy = dy.contiguous()
        dh = grouped_matmul(dy, w2, offsets, b_mn=True)
        dw2 = grouped_wgrad(dy, h, offsets, w2)
        dgu = swiglu_interleaved_bwd(dh, gu)
        dx = grouped_matmul(dgu, w13, offsets, b_mn=True) if ctx.needs_input_grad[0] else None
        dw13 = grouped_wgrad(dgu, x, offsets, w13)
        return (dx, None, None, *dw13, *dw2)


def grouped_swiglu_mlp(x: torch.Tensor, offsets: torch.Tensor, w13: Sequence[torch.Tensor], w2: Sequence[torch.Tensor]):
    """``x`` packed rows ``[R, hidden]`` (padding rows zero), ``offsets`` int32 ``[E + 1]`` → ``[R, hidden]``."""
    return _GroupedSwiGLUMLPFn.apply(x, offsets, len(w13), *w13, *w2)
