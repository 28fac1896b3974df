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
            # the tables outlive the call that builds them: a cache filled while validation runs under inference_mode() would be
            # made of inference tensors, which a later TRAINING step cannot save for backward - build them as normal tensors
            with torch.inference_mode(False), torch.no_grad():
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


# ---------------------------------------------------------------------------------------------------------------------
# flash-attn style autograd functions kept for code written against the reference (``internlm/model/modules/embedding.py:
# 89-170`` ``ApplyRotaryEmb``: out of place on ``[b, s, heads, d]``; ``:172-283`` ``ApplyRotaryEmbQKV_``: in place on q and k of
# a packed ``[total, 3, heads, d]`` or padded ``[b, s, 3, heads, d]`` qkv).  Full-width rotations of bf16 CUDA tensors go
# through ``rope_`` (one launch for q and k together); partial ``rotary_dim`` and separate key tables use the fp32 formula.
# ---------------------------------------------------------------------------------------------------------------------
def _rotate_prefix_(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, interleaved: bool, conj: bool) -> None:
    """Rotate the first ``2 * cos.shape[-1]`` features of ``x [..., rows, heads, d]`` in place; ``cos`` / ``sin`` are
    ``[rows, rd / 2]`` (broadcast over the leading dims and the heads)."""
    rd = cos.shape[-1] * 2
    c, s = cos.float().unsqueeze(-2), sin.float().unsqueeze(-2)
    if conj:
        s = -s
    ro = x[..., :rd]
    x1, x2 = (ro[..., 0::2], ro[..., 1::2]) if interleaved else (ro[..., : rd // 2], ro[..., rd // 2:])
    a, b = x1.float(), x2.float()
    o1, o2 = (a * c - b * s).to(x.dtype), (a * s + b * c).to(x.dtype)
    x1.copy_(o1)
    x2.copy_(o2)


def _native_full_width(x: torch.Tensor, cos: torch.Tensor) -> bool:
    return _lib.use_native(x) and x.dtype == torch.bfloat16 and x.is_contiguous() and cos.shape[-1] * 2 == x.shape[-1] \
        and cos.dtype == torch.float32


class ApplyRotaryEmb(torch.autograd.Function):
    """``out = rope(x)`` for ``x [batch, seqlen, heads, d]`` with tables ``[>= seqlen, rotary_dim / 2]`` (positions 0..s-1)."""

    @staticmethod
    def forward(ctx, x, cos, sin, interleaved=False):
        b, s, _, d = x.shape
        assert cos.shape == sin.shape and cos.shape[0] >= s and cos.shape[1] * 2 <= d
        ctx.save_for_backward(cos, sin)
        ctx.interleaved = interleaved
        out = x.clone(memory_format=torch.contiguous_format)
        ApplyRotaryEmb._run(out, cos, sin, interleaved, False)
        return out

    @staticmethod
    def _run(t, cos, sin, interleaved, conj):
        b, s, h, d = t.shape
        if _native_full_width(t, cos):
            pos = torch.arange(s, device=t.device, dtype=torch.int32).repeat(b)
            rope_(t.view(b * s, h, d), pos, cos.contiguous(), sin.contiguous(), 1, 1, conj, interleaved)
        else:
            _rotate_prefix_(t, cos[:s], sin[:s], interleaved, conj)

    @staticmethod
    def backward(ctx, do):
        cos, sin = ctx.saved_tensors
        dx = do.clone(memory_format=torch.contiguous_format)
        ApplyRotaryEmb._run(dx, cos, sin, ctx.interleaved, True)
        return dx, None, None, None


class ApplyRotaryEmbQKV_(torch.autograd.Function):
    """In-place RoPE of q and k inside ``qkv``: packed ``[total, 3, heads, d]`` with per-token tables ``[total, rd / 2]``, or
    padded ``[b, s, 3, heads, d]`` with tables indexed by position."""

    @staticmethod
    def forward(ctx, qkv, cos, sin, cos_k=None, sin_k=None, interleaved=False):
        packed = qkv.dim() == 4
        assert qkv.shape[1 if packed else 2] == 3
        ctx.save_for_backward(cos, sin, cos_k, sin_k)
        ctx.interleaved = interleaved
        ApplyRotaryEmbQKV_._run(qkv, cos, sin, cos_k, sin_k, interleaved, False)
        ctx.mark_dirty(qkv)
        return qkv

    @staticmethod
    def _run(qkv, cos, sin, cos_k, sin_k, interleaved, conj):
        packed = qkv.dim() == 4
        same = cos_k is None and sin_k is None
        if packed:
            total, _, h, d = qkv.shape
            if same and _native_full_width(qkv, cos) and cos.shape[0] >= total:
                # heads 0..h-1 are q, h..2h-1 are k, 2h..3h-1 are v: one launch rotates the first two thirds of every row
                rope_(qkv.view(total, 3 * h, d), None, cos.contiguous(), sin.contiguous(), 3 * h, 2 * h, conj, interleaved)
                return
            _rotate_prefix_(qkv[:, 0], cos[:total], sin[:total], interleaved, conj)
            _rotate_prefix_(qkv[:, 1], (cos if cos_k is None else cos_k)[:total], (sin if sin_k is None else sin_k)[:total],
                            interleaved, conj)
        else:
            b, s, _, h, d = qkv.shape
            if same and _native_full_width(qkv, cos):
                pos = torch.arange(s, device=qkv.device, dtype=torch.int32).repeat(b)
                rope_(qkv.view(b * s, 3 * h, d), pos, cos.contiguous(), sin.contiguous(), 3 * h, 2 * h, conj, interleaved)
                return
            _rotate_prefix_(qkv[:, :, 0], cos[:s], sin[:s], interleaved, conj)
            _rotate_prefix_(qkv[:, :, 1], (cos if cos_k is None else cos_k)[:s], (sin if sin_k is None else sin_k)[:s],
                            interleaved, conj)

    @staticmethod
    def backward(ctx, dqkv):
        cos, sin, cos_k, sin_k = ctx.saved_tensors
        dqkv = dqkv.contiguous()
        ApplyRotaryEmbQKV_._run(dqkv, cos, sin, cos_k, sin_k, ctx.interleaved, True)
        return dqkv, None, None, None, None, None


apply_rotary_emb = ApplyRotaryEmb.apply
apply_rotary_emb_qkv_ = ApplyRotaryEmbQKV_.apply
