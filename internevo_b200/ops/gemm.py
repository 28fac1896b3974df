"""tcgen05 GEMM front-end and the autograd linear built on it.

``matmul(a, b, a_mn, b_mn)`` computes ``A @ B^T`` where the logical ``A`` is ``[M, K]`` and ``B`` is ``[N, K]``;
``a_mn`` / ``b_mn`` say that the tensor passed in is stored transposed (``[K, M]`` / ``[K, N]``, "MN-major").  One
kernel therefore covers the forward (``x @ W^T``), dgrad (``dy @ W``) and wgrad (``dy^T @ x``) GEMMs that the reference
sends to cuBLAS / cublasLt (reference ``internlm/model/utils.py:228-346``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

GEMM_OUT_F32 = 1
GEMM_ACCUMULATE = 2
GEMM_SWIGLU = 4
GEMM_SKIP_D = 8
GEMM_GELU = 16

_launches = 0  # counted for bench.py's "gpu_launches"


def launch_count() -> int:
    return _launches


def _bump(n: int = 1) -> None:
    global _launches
    _launches += n


def _ref_matmul(a, b, a_mn, b_mn, bias, out, accumulate, out_dtype):
    A = a.t() if a_mn else a
    Bt = b if b_mn else b.t()
    r = A.float() @ Bt.float()
    if bias is not None:
        r = r + bias.float()
    if out is not None:
        if accumulate:
            r = r + out.float()
        out.copy_(r.to(out.dtype))
        return out
    return r.to(out_dtype or a.dtype)


def matmul(
    a: torch.Tensor,
    b: torch.Tensor,
    a_mn: bool = False,
    b_mn: bool = False,
    bias: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    accumulate: bool = False,
    out_dtype: Optional[torch.dtype] = None,
    force_bn: int = 0,
    max_ctas: int = 0,
) -> torch.Tensor:
    """``A @ B^T (+ bias)`` with fp32 accumulation; bf16 inputs; bf16 or fp32 output (optionally accumulated in place)."""
    assert a.dim() == 2 and b.dim() == 2
    if not (_lib.use_native(a, b) and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16):
        return _ref_matmul(a, b, a_mn, b_mn, bias, out, accumulate, out_dtype)
    M = a.shape[1] if a_mn else a.shape[0]
    N = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype or torch.bfloat16)
        accumulate = False
    flags = (GEMM_OUT_F32 if out.dtype == torch.float32 else 0) | (GEMM_ACCUMULATE if accumulate else 0)
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    torch.ops.b200.gemm(a, b, out, a_mn, b_mn, bias, flags, None, force_bn, max_ctas)
    _bump()
    return out


def matmul_swiglu(x: torch.Tensor, w_gu: torch.Tensor, store_gu: bool = True, force_bn: int = 0):
    """``gu = x @ w_gu^T`` with interleaved (gate, up) rows in ``w_gu``; returns ``(gu, silu(gate) * up)``.

    The activation is applied in the GEMM epilogue straight out of TMEM, so the separate SwiGLU pass over ``[T, 2F]``
    of the reference (``internlm/model/utils.py:684-688``) disappears.
    """
    M, N = x.shape[0], w_gu.shape[0]
    if not (_lib.use_native(x, w_gu) and x.dtype == torch.bfloat16):
        gu = (x.float() @ w_gu.float().t()).to(x.dtype)
        g, u = gu[:, 0::2].float(), gu[:, 1::2].float()
        return gu, (torch.nn.functional.silu(g) * u).to(x.dtype)
    # with GEMM_SKIP_D the kernel never writes D; the (lazily backed) allocation only satisfies the shape checks
    d = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    h = torch.empty(M, N // 2, device=x.device, dtype=torch.bfloat16)
    flags = GEMM_SWIGLU | (0 if store_gu else GEMM_SKIP_D)
    torch.ops.b200.gemm(x, w_gu, d, False, False, None, flags, h, force_bn, 0)
    _bump()
    return (d if store_gu else None), h


def matmul_gelu(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None):
    """``pre = x @ w^T + bias`` and ``gelu_tanh(pre)`` from ONE GEMM (activation applied in the epilogue out of TMEM);
    returns ``(pre, act)`` — replaces the cuBLASLt GELU epilogue of the reference's fused-dense extension
    (``csrc/fused_dense_lib`` ``linear_act_forward``)."""
    if not (_lib.use_native(x, w) and x.dtype == torch.bfloat16):
        pre = torch.nn.functional.linear(x.float(), w.float(), None if bias is None else bias.float()).to(x.dtype)
        return pre, torch.nn.functional.gelu(pre.float(), approximate="tanh").to(x.dtype)
    M, N = x.shape[0], w.shape[0]
    pre = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    act = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    torch.ops.b200.gemm(x, w, pre, False, False, bias, GEMM_GELU, act, 0, 0)
    _bump()
    return pre, act


class _LinearGeluFn(torch.autograd.Function):
    """``gelu(x W^T + b)`` with the pre-activation kept for backward (``bias_act_linear_dgrad_bgrad`` of the reference)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        pre, act = matmul_gelu(x2, weight, bias)
        ctx.save_for_backward(x2, weight, pre)
        ctx.has_bias, ctx.x_shape = bias is not None, x.shape
        return act.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dact):
        x2, weight, pre = ctx.saved_tensors
        d2 = dact.reshape(-1, dact.shape[-1]).contiguous()
        if _lib.use_native(d2, pre):
            dpre = torch.empty_like(pre)
            torch.ops.b200.gelu_bwd(d2, pre, dpre)
            _bump()
        else:
            p = pre.float().requires_grad_(True)
            with torch.enable_grad():
                a = torch.nn.functional.gelu(p, approximate="tanh")
            dpre = torch.autograd.grad(a, p, d2.float())[0].to(pre.dtype)
        dx = matmul(dpre, weight, b_mn=True).reshape(ctx.x_shape) if ctx.needs_input_grad[0] else None
        dw = wgrad(dpre, x2, weight) if ctx.needs_input_grad[1] else None
        db = dpre.float().sum(0).to(dpre.dtype) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear_gelu(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _LinearGeluFn.apply(x, weight, bias)


class _LinearFn(torch.autograd.Function):
    """y = x @ W^T (+ b).  Backward runs the dgrad / wgrad tcgen05 GEMMs; the wgrad accumulates straight into the
    parameter's persistent gradient buffer (``weight.grad_buf``) when the optimizer provides one."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.x_shape = x.shape
        x2 = x.reshape(-1, x.shape[-1])
        y = matmul(x2, weight, bias=bias)
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = matmul(dy2, weight, b_mn=True).reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = wgrad(dy2, x2, weight)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.float().sum(0).to(dy.dtype)
        return dx, dw, db


def wgrad(dy2: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor):
    """dW = dy^T @ x.  If the parameter carries a persistent ``grad_buf`` the result is accumulated there in the GEMM
    epilogue and ``None`` is returned to autograd (no separate ``grad += dW`` pass, no flatten/unflatten copies)."""
    buf = getattr(weight, "grad_buf", None)
    if buf is not None:
        fresh = not getattr(weight, "grad_ready", False)
        matmul(dy2, x2, a_mn=True, b_mn=True, out=buf, accumulate=not fresh)
        weight.grad_ready = True
        hook = getattr(weight, "grad_hook", None)
        if hook is not None:
            hook(weight)
        return None
    return matmul(dy2, x2, a_mn=True, b_mn=True, out_dtype=weight.dtype)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    if x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16:
        return _LinearFn.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)
