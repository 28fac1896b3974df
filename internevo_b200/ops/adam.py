"""Fused AdamW / grad-norm kernels on flat shards (device-resident scalars: no host sync inside the optimizer step).

Replaces ``torch.optim.AdamW(fused=True)`` + the cast / unscale / copy-back passes of the reference
(``internlm/solver/optimizer/hybrid_zero_optim.py:740-797``) and apex ``multi_tensor_l2norm``
(``internlm/solver/optimizer/utils.py:177-204``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .gemm import _bump


# The kernels read bf16 or fp32 gradients; fp16 models (``model.dtype`` defaults to ``torch.float16`` as in the reference,
# ``initialize/launch.py``) take the plain PyTorch arithmetic below instead of tripping the kernel's dtype check.
_NATIVE_GRAD = (torch.bfloat16, torch.float32)


def sumsq_(g: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``out[0] += sum(g^2)`` (fp32)."""
    if g.numel() == 0:
        return out
    if _lib.use_native(g, out) and g.dtype in _NATIVE_GRAD:
        torch.ops.b200.sumsq(g.contiguous().view(-1), out)
        _bump()
    else:
        out += g.float().pow(2).sum()
    return out


def clip_scalars_(sumsq: torch.Tensor, scalars: torch.Tensor, loss_scale: float, clip: float) -> torch.Tensor:
    """scalars[0] = 1/(loss_scale*max(1, norm/clip)), scalars[1] = overflow flag, scalars[2] = unscaled norm (-1 if bad)."""
    if _lib.use_native(sumsq, scalars):
        torch.ops.b200.clip_scalars(sumsq, scalars, loss_scale, clip)
        _bump()
    else:
        ss = sumsq.reshape(-1)[0]
        bad = bool(torch.isnan(ss) or torch.isinf(ss))
        norm = float(ss.sqrt()) / loss_scale if not bad else -1.0
        mult = 1.0 / loss_scale
        if clip > 0 and not bad and norm / clip > 1:
            mult /= norm / clip
        scalars[0] = 0.0 if bad else mult
        scalars[1] = 1.0 if bad else 0.0
        scalars[2] = norm
    return scalars


def adamw_(p: torch.Tensor, m: torch.Tensor, v: torch.Tensor, g: torch.Tensor, p_lp: Optional[torch.Tensor], lr: float,
           beta1: float, beta2: float, eps: float, weight_decay: float, step: int,
           scalars: Optional[torch.Tensor] = None) -> None:
    """One AdamW step on flat fp32 ``p/m/v`` from (bf16|fp32) ``g``; writes the low-precision copy ``p_lp`` if given."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    if p.numel() == 0:
        return
    if _lib.use_native(p, g) and g.dtype in _NATIVE_GRAD and (p_lp is None or p_lp.dtype in _NATIVE_GRAD):
        torch.ops.b200.adamw(p, m, v, g, p_lp, lr, beta1, beta2, eps, weight_decay, bc1, bc2, scalars)
        _bump()
        return
    mult = 1.0
    if scalars is not None:
        if float(scalars[1]) != 0.0:
            return
        mult = float(scalars[0])
    gg = g.float() * mult
    m.mul_(beta1).add_(gg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
    p.mul_(1 - lr * weight_decay)
    p.addcdiv_(m / bc1, (v / bc2).sqrt() + eps, value=-lr)
    if p_lp is not None:
        p_lp.copy_(p)
