"""Rotary position embedding applied in place on the packed qkv tensor (q and k heads in ONE launch).

Replaces flash-attn's ``rotary_emb.apply_rotary`` wrappers (reference ``internlm/model/modules/embedding.py:89-257``).
cos/sin come from fp32 tables indexed by the packed-sequence position ids (``indexes``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .gemm import _bump


class RotaryTables:
    """Lazily grown fp32 cos/sin tables ``[max_pos, dim/2]`` (reference ``RotaryEmbedding._update_cos_sin_cache``,
    ``internlm/model/modules/embedding.py:263-376``); supports linear scaling and dynamic-NTK base rescaling."""

    def __init__(self, dim: int, base: float = 10000.0, scale_base: float = 0, device=None,
                 scaling_factor: float = 1.0, ntk_max_position: int = 0):
        self.dim, self.base, self.device = dim, float(base), device
        self.scaling_factor = scaling_factor  # LinearRotaryEmbedding: t / factor
        self.ntk_max_position = ntk_max_position  # DynamicNTK: rescale base beyond this length
        self.cos: Optional[torch.Tensor] = None
        self.sin: Optional[torch.Tensor] = None
        self.max_pos = 0

    def _inv_freq(self, seqlen: int, device) -> torch.Tensor:
        base = self.base
        if self.ntk_max_position and seqlen > self.ntk_max_position:
            base = self.base * (
                (self.scaling_factor * seqlen / self.ntk_max_position) - (self.scaling_factor - 1)
            ) ** (self.dim / (self.dim - 2))
        return 1.0 / (base ** (torch.arange(0, self.dim, 2, device=device, dtype=torch.float32) / self.dim))

    def get(self, max_pos: int, device) -> tuple:
        if self.cos is None or max_pos > self.max_pos or self.cos.device != torch.device(device):
            n = max(max_pos, 2 * self.max_pos, 4096)
            t = torch.arange(n, device=device, dtype=torch.float32)
            if not self.ntk_max_position and self.scaling_factor != 1.0:
                t = t / self.scaling_factor
            freqs = torch.outer(t, self._inv_freq(n, device))
            self.cos, self.sin, self.max_pos = freqs.cos().contiguous(), freqs.sin().contiguous(), n
        return self.cos, self.sin


def _rope_ref(x, pos, cos, sin, group, rot, conj, interleaved):
    # x: [T, heads, D]
    T, Hh, D = x.shape
    c = cos[pos.long()] if pos is not None else cos[:T]
    s = sin[pos.long()] if pos is not None else sin[:T]
    if conj:
        s = -s
    xf = x.float()
    mask = (torch.arange(Hh, device=x.device) % group) < rot
    if interleaved:
        x1, x2 = xf[..., 0::2], xf[..., 1::2]
        o = torch.stack([x1 * c[:, None] - x2 * s[:, None], x1 * s[:, None] + x2 * c[:, None]], -1).flatten(-2)
    else:
        x1, x2 = xf[..., : D // 2], xf[..., D // 2:]
        o = torch.cat([x1 * c[:, None] - x2 * s[:, None], x1 * s[:, None] + x2 * c[:, None]], -1)
    out = torch.where(mask[None, :, None], o, xf).to(x.dtype)
    x.copy_(out)
    return x


def rope_(x: torch.Tensor, pos: Optional[torch.Tensor], cos: torch.Tensor, sin: torch.Tensor, group: int = 1,
          rot_per_group: int = 1, conj: bool = False, interleaved: bool = False) -> torch.Tensor:
    """In-place rotation of ``x`` viewed as ``[T, heads, D]``; only heads with ``head % group < rot_per_group`` rotate."""
    if _lib.use_native(x) and x.dtype == torch.bfloat16:
        if pos is not None and pos.dtype != torch.int32:
            pos = pos.int()
        torch.ops.b200.rope(x, pos, cos, sin, group, rot_per_group, conj, interleaved)
        _bump()
        return x
    return _rope_ref(x, pos, cos, sin, group, rot_per_group, conj, interleaved)


class _RopeFn(torch.autograd.Function):
    """In place on ``x`` (``[T, heads * D]`` or ``[T, heads, D]``).  Call it on the tensor the projection GEMM returned,
    not on a view of it: an in-place autograd op on a view makes autograd insert CopySlices (two full copies of the
    gradient per call in backward)."""

    @staticmethod
    def forward(ctx, x, pos, cos, sin, group, rot, interleaved, head_dim):
        ctx.save_for_backward(pos, cos, sin)
        ctx.cfg = (group, rot, interleaved, head_dim)
        ctx.mark_dirty(x)
        rope_(x.view(x.shape[0], -1, head_dim), pos, cos, sin, group, rot, False, interleaved)
        return x

    @staticmethod
    def backward(ctx, dx):
        pos, cos, sin = ctx.saved_tensors
        group, rot, interleaved, head_dim = ctx.cfg
        dx = dx.contiguous()  # the incoming gradient is produced by our attention backward: safe to rotate in place
        rope_(dx.view(dx.shape[0], -1, head_dim), pos, cos, sin, group, rot, True, interleaved)
        return dx, None, None, None, None, None, None, None


def apply_rotary_packed(x: torch.Tensor, pos: Optional[torch.Tensor], cos: torch.Tensor, sin: torch.Tensor,
                        group: int = 1, rot_per_group: int = 1, interleaved: bool = False,
                        head_dim: Optional[int] = None) -> torch.Tensor:
    """Autograd-aware in-place RoPE on ``[T, heads, D]`` or, with ``head_dim``, on the flat ``[T, heads * D]`` projection
    output (the InternLM2 packed wqkv viewed as heads)."""
    return _RopeFn.apply(x, pos, cos, sin, group, rot_per_group, interleaved, head_dim or x.shape[-1])
