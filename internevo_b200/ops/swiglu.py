"""SwiGLU on interleaved (gate, up) columns. Forward normally lives in the GEMM epilogue (``matmul_swiglu``); this file
has the standalone kernels (used for backward and for non-fused weights).  Reference: ``internlm/model/utils.py:684-688``."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .gemm import _bump


def swiglu_interleaved_fwd(gu: torch.Tensor) -> torch.Tensor:
    if _lib.use_native(gu) and gu.dtype == torch.bfloat16:
        gu = gu.contiguous()
        h = torch.empty(*gu.shape[:-1], gu.shape[-1] // 2, device=gu.device, dtype=gu.dtype)
        torch.ops.b200.swiglu_fwd(gu, h)
        _bump()
        return h
    g, u = gu[..., 0::2].float(), gu[..., 1::2].float()
    return (F.silu(g) * u).to(gu.dtype)


def swiglu_interleaved_bwd(dh: torch.Tensor, gu: torch.Tensor) -> torch.Tensor:
    if _lib.use_native(gu, dh) and gu.dtype == torch.bfloat16:
        dgu = torch.empty_like(gu)
        torch.ops.b200.swiglu_bwd(dh.contiguous(), gu.contiguous(), dgu)
        _bump()
        return dgu
    g, u, d = gu[..., 0::2].float(), gu[..., 1::2].float(), dh.float()
    sg = torch.sigmoid(g)
    dg = d * u * sg * (1 + g * (1 - sg))
    du = d * g * sg
    return torch.stack([dg, du], -1).flatten(-2).to(gu.dtype)


class _SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu):
        ctx.save_for_backward(gu)
        return swiglu_interleaved_fwd(gu)

    @staticmethod
    def backward(ctx, dh):
        (gu,) = ctx.saved_tensors
        return swiglu_interleaved_bwd(dh, gu)


def swiglu_interleaved(gu: torch.Tensor) -> torch.Tensor:
    return _SwiGLUFn.apply(gu)


def silu_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Non-interleaved ``silu(a) * b`` (separate w1 / w3 outputs, reference ``Silu``)."""
    return F.silu(a) * b
