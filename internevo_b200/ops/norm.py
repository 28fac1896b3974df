"""Fused residual-add + RMSNorm and fused dropout + residual-add + LayerNorm (forward and backward are one kernel each).

Replaces apex ``MixedFusedRMSNorm`` / ``RMSNormTorch`` (reference ``internlm/model/ops/norm.py:10-46``,
``internlm/model/utils.py:662-675``) and the separate ``dropout(h) + residual`` elementwise pass of the block
(reference ``internlm/model/modeling_internlm2.py:697-707``).  ``LayerNorm`` replaces flash-attn's ``dropout_add_layer_norm``
(``csrc/layernorm.cu``; reference ``norm_type="layernorm"`` blocks, ``internlm/model/modeling_internlm.py:215-248``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from .gemm import _bump


def rmsnorm_ref(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (y * weight.float()).to(weight.dtype if weight.dtype != torch.float32 or x.dtype == torch.float32 else x.dtype)


class _AddRMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, eps):
        # returns (y, new_residual); new_residual = x + residual (or x itself when residual is None)
        H = x.shape[-1]
        xc = x.contiguous()
        rows = xc.numel() // H
        y = torch.empty_like(xc)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        if residual is not None:
            res_out = torch.empty_like(xc)
            torch.ops.b200.rmsnorm_fwd(xc, residual.contiguous(), weight, y, res_out, rstd, eps)
        else:
            res_out = xc
            torch.ops.b200.rmsnorm_fwd(xc, None, weight, y, None, rstd, eps)
        _bump()
        ctx.save_for_backward(res_out, weight, rstd)
        ctx.has_res = residual is not None
        return y, res_out

    @staticmethod
    def backward(ctx, dy, dres_out):
        res, weight, rstd = ctx.saved_tensors
        H = res.shape[-1]
        rows = res.numel() // H
        dy = dy.contiguous()
        dx = torch.empty_like(res)
        nblk = torch.ops.b200.rmsnorm_bwd_blocks(rows)
        partial = torch.empty(nblk * H, device=res.device, dtype=torch.float32)
        buf = getattr(weight, "grad_buf", None)
        if buf is not None:
            fresh = not getattr(weight, "grad_ready", False)
            torch.ops.b200.rmsnorm_bwd(dy, res, weight, rstd, dres_out.contiguous() if dres_out is not None else None,
                                       dx, partial, buf, not fresh)
            weight.grad_ready = True
            hook = getattr(weight, "grad_hook", None)
            if hook is not None:
                hook(weight)
            dw = None
        else:
            dw = torch.empty_like(weight)
            torch.ops.b200.rmsnorm_bwd(dy, res, weight, rstd, dres_out.contiguous() if dres_out is not None else None,
                                       dx, partial, dw, False)
        _bump(2)
        # d(new_residual)/dx = d(new_residual)/d(residual) = identity: both inputs receive the same gradient
        return dx, (dx if ctx.has_res else None), dw, None


def add_rmsnorm(
    x: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor, eps: float
) -> Tuple[torch.Tensor, torch.Tensor]:
    """``new_res = x + residual``; ``y = RMSNorm(new_res) * weight``.  Returns ``(y, new_res)``."""
    H = x.shape[-1]
    if residual is not None and residual.dtype != x.dtype:
        # ``model.residual_in_fp32``: the residual stream stays fp32 across blocks (reference ``modeling_internlm2.py:697-707``);
        # the normalised activations go back to the compute dtype
        new_res = x.to(residual.dtype) + residual
        return rmsnorm_ref(new_res, weight, eps).to(x.dtype), new_res
    if (
        _lib.use_native(x, weight)
        and x.dtype == torch.bfloat16
        and weight.dtype == torch.bfloat16
        and H % 8 == 0
        and H <= 8192
    ):
        return _AddRMSNormFn.apply(x, residual, weight, eps)
    new_res = x if residual is None else x + residual
    return rmsnorm_ref(new_res, weight, eps).to(x.dtype), new_res


class RMSNorm(nn.Module):
    """RMSNorm with learnable scale.  ``forward(x)`` or ``forward(x, residual) -> (y, new_residual)``."""

    def __init__(self, hidden_size: int, eps: float = 1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))

    def reset_parameters(self):
        nn.init.ones_(self.weight)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None):
        if residual is None:
            return add_rmsnorm(x, None, self.weight, self.eps)[0]
        return add_rmsnorm(x, residual, self.weight, self.eps)

    def extra_repr(self):
        return f"{tuple(self.weight.shape)}, eps={self.eps}"


# ---------------------------------------------------------------------------------------------------------------------
# dropout + residual-add + LayerNorm
# ---------------------------------------------------------------------------------------------------------------------
def layernorm_ref(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = (xf - mean).pow(2).mean(-1, keepdim=True)
    y = (xf - mean) * torch.rsqrt(var + eps) * weight.float()
    if bias is not None:
        y = y + bias.float()
    return y.to(x.dtype)


class _DropAddLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, keep, drop_scale):
        H = x.shape[-1]
        xc = x.contiguous()
        rows = xc.numel() // H
        y = torch.empty_like(xc)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        fused_res = residual is not None or keep is not None
        res_out = torch.empty_like(xc) if fused_res else xc
        torch.ops.b200.layernorm_fwd(xc, residual.contiguous() if residual is not None else None, keep, drop_scale, weight,
                                     bias, y, res_out if fused_res else None, mean, rstd, eps)
        _bump()
        ctx.save_for_backward(res_out, weight, mean, rstd, keep)
        ctx.has_res, ctx.has_bias, ctx.drop_scale = residual is not None, bias is not None, drop_scale
        return y, res_out

    @staticmethod
    def backward(ctx, dy, dres_out):
        res, weight, mean, rstd, keep = ctx.saved_tensors
        H = res.shape[-1]
        rows = res.numel() // H
        dr = torch.empty_like(res)
        nblk = torch.ops.b200.layernorm_bwd_blocks(rows)
        partial = torch.empty(nblk * 2 * H, device=res.device, dtype=torch.float32)
        dwdb = torch.empty(2 * H, device=res.device, dtype=torch.float32)
        torch.ops.b200.layernorm_bwd(dy.contiguous(), res, weight, mean, rstd,
                                     dres_out.contiguous() if dres_out is not None else None, dr, partial, dwdb)
        _bump(2)
        dx = dr if keep is None else dr * keep.view_as(dr).to(dr.dtype) * ctx.drop_scale
        return (dx, dr if ctx.has_res else None, dwdb[:H].to(weight.dtype),
                dwdb[H:].to(weight.dtype) if ctx.has_bias else None, None, None, None)


def dropout_add_layernorm(x: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor,
                          bias: Optional[torch.Tensor], eps: float, dropout_p: float = 0.0, training: bool = False
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``new_res = dropout(x) + residual``; ``y = LayerNorm(new_res) * weight + bias``.  Returns ``(y, new_res)``."""
    H = x.shape[-1]
    p = dropout_p if training else 0.0
    if (_lib.use_native(x, weight) and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and (bias is None or bias.dtype == torch.bfloat16) and H % 8 == 0 and H <= 8192):
        keep = None
        if p > 0.0:
            keep = (torch.rand(x.shape, device=x.device) >= p).to(torch.uint8)
        return _DropAddLayerNormFn.apply(x, residual, weight, bias, eps, keep, 1.0 / (1.0 - p) if p > 0.0 else 1.0)
    if p > 0.0:
        x = torch.nn.functional.dropout(x, p, True)
    new_res = x if residual is None else x + residual
    return layernorm_ref(new_res, weight, bias, eps), new_res


class LayerNorm(nn.Module):
    """LayerNorm with the block-prologue calling convention of ``RMSNorm``: ``forward(x)`` or
    ``forward(x, residual) -> (y, new_residual)``; ``dropout_p`` is the dropout applied to ``x`` before the add."""

    def __init__(self, hidden_size: int, eps: float = 1e-5, bias: bool = True, dropout_p: float = 0.0, device=None,
                 dtype=None):
        super().__init__()
        self.eps, self.dropout_p = eps, dropout_p
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(hidden_size, device=device, dtype=dtype)) if bias else None

    def reset_parameters(self):
        nn.init.ones_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None):
        y, new_res = dropout_add_layernorm(x, residual, self.weight, self.bias, self.eps, self.dropout_p, self.training)
        return y if residual is None else (y, new_res)

    def extra_repr(self):
        return f"{tuple(self.weight.shape)}, eps={self.eps}"


def manual_rms_norm(my_input, normalized_shape, weight, eps):
    """Unfused RMSNorm over the trailing ``normalized_shape`` dims with fp32 statistics (reference ``ops/norm.py``)."""
    dims = tuple(range(-len(normalized_shape), 0))
    var = my_input.float().pow(2).mean(dims, keepdim=True)
    y = my_input * torch.rsqrt(var + eps)
    if weight is None:
        return y
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.to(weight.dtype)
    return weight * y


class RMSNormTorch(nn.Module):
    """Pure-PyTorch RMSNorm: CPU runs and numerical oracle for :class:`RMSNorm`."""

    def __init__(self, normalized_shape, eps: float = 1e-5, device=None, dtype=None):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = torch.Size(normalized_shape)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(*normalized_shape, device=device, dtype=dtype))

    def reset_parameters(self):
        nn.init.ones_(self.weight)

    def forward(self, x):
        return manual_rms_norm(x, self.normalized_shape, self.weight, self.eps)

    def extra_repr(self):
        return f"{tuple(self.normalized_shape)}, eps={self.eps}"
