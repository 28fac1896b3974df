"""Variable-length causal attention with GQA.

``flash_attention_varlen`` dispatches to the hand-written sm_100a kernel (``csrc/attention_sm100.cu``: TMA-fed
``tcgen05`` QK^T / PV with the score tile in TMEM and an online-softmax warpgroup); it replaces flash-attn's
``flash_attn_varlen_kvpacked_func`` (reference ``internlm/model/modeling_internlm2.py:446-468``), whose pinned build
refuses sm_100 altogether.  ``attention_ref`` is the fp32 PyTorch oracle and the CPU path.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import _lib
from .gemm import _bump

_IMPL = os.environ.get("INTERNEVO_ATTN_IMPL", "b200")  # b200 | flash_attn | sdpa


def set_attention_impl(name: str):
    global _IMPL
    assert name in ("b200", "flash_attn", "sdpa")
    _IMPL = name


def get_attention_impl() -> str:
    return _IMPL


def attention_ref(q, k, v, cu_seqlens, causal=True, scale=None):
    """fp32 reference. q ``[T, H, D]``, k/v ``[T, Hkv, D]``, ``cu_seqlens`` int32 ``[n+1]``."""
    T, H, D = q.shape
    Hkv = k.shape[1]
    scale = scale or 1.0 / math.sqrt(D)
    out = torch.empty(T, H, v.shape[-1], dtype=torch.float32, device=q.device)
    cu = cu_seqlens.tolist()
    rep = H // Hkv
    outs = []
    for a, b in zip(cu[:-1], cu[1:]):
        qs = q[a:b].float().transpose(0, 1)  # [H, S, D]
        ks = k[a:b].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        vs = v[a:b].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        s = qs @ ks.transpose(1, 2) * scale
        if causal:
            S = b - a
            mask = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
            s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        outs.append((p @ vs).transpose(0, 1))
    del out
    return torch.cat(outs, 0) if outs else q.new_zeros(0, H, D, dtype=torch.float32)


class _B200AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, scale, causal):
        T, H, D = q.shape
        out = torch.empty(T, H, D, device=q.device, dtype=q.dtype)
        lse = torch.empty(H, T, device=q.device, dtype=torch.float32)
        torch.ops.b200.attn_fwd(q, k, v, out, lse, cu_seqlens, max_seqlen, scale, causal)
        _bump()
        ctx.save_for_backward(q, k, v, out, lse, cu_seqlens)
        ctx.cfg = (max_seqlen, scale, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, cu = ctx.saved_tensors
        max_seqlen, scale, causal = ctx.cfg
        T, H, D = q.shape
        dout = dout.contiguous()
        # gradients are written in the same (possibly packed/strided) layout as the inputs when those are views of one
        # qkv tensor, so the RoPE backward and the wqkv wgrad consume them without a gather copy
        dq = torch.empty_like(q) if q.is_contiguous() else torch.empty(q.shape, device=q.device, dtype=q.dtype)
        dk = torch.empty(k.shape, device=k.device, dtype=k.dtype)
        dv = torch.empty(v.shape, device=v.device, dtype=v.dtype)
        delta = torch.empty(2, H, T, device=q.device, dtype=torch.float32)  # [0] rowsum(dO*O), [1] lse*log2e
        dq_acc = torch.zeros(T, H, D, device=q.device, dtype=torch.float32)
        torch.ops.b200.attn_bwd(dout, q, k, v, out, lse, dq, dk, dv, delta, dq_acc, cu, max_seqlen, scale, causal)
        _bump(3)
        return dq, dk, dv, None, None, None, None


class _B200AttnPackedFn(torch.autograd.Function):
    """Attention straight on the packed InternLM2 ``wqkv`` output ``[T, kv_heads, q_per_kv + 2, D]`` (post-RoPE).

    q / k / v are strided views of ONE buffer for the TMA descriptors, and the backward kernels write dq / dk / dv into
    ONE ``dqkv`` buffer of the same layout — no split / concat / zero-fill + add chains in autograd (those cost ~36 ms
    per 7B step with a q,k,v-separate op, see profiles/step_profile_r1_v1_7B_1gpu.txt)."""

    @staticmethod
    def forward(ctx, qkv, cu_seqlens, max_seqlen, scale, causal):
        T, G, gs, D = qkv.shape
        H = G * (gs - 2)
        out = torch.empty(T, H, D, device=qkv.device, dtype=qkv.dtype)
        lse = torch.empty(H, T, device=qkv.device, dtype=torch.float32)
        torch.ops.b200.attn_fwd(qkv[:, :, : gs - 2], qkv[:, :, gs - 2], qkv[:, :, gs - 1], out, lse, cu_seqlens,
                                max_seqlen, scale, causal)
        _bump()
        ctx.save_for_backward(qkv, out, lse, cu_seqlens)
        ctx.cfg = (max_seqlen, scale, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, cu = ctx.saved_tensors
        max_seqlen, scale, causal = ctx.cfg
        T, G, gs, D = qkv.shape
        H = G * (gs - 2)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(2, H, T, device=qkv.device, dtype=torch.float32)  # [0] rowsum(dO*O), [1] lse*log2e
        dq_acc = torch.zeros(T, H, D, device=qkv.device, dtype=torch.float32)
        torch.ops.b200.attn_bwd(dout.contiguous(), qkv[:, :, : gs - 2], qkv[:, :, gs - 2], qkv[:, :, gs - 1], out, lse,
                                dqkv[:, :, : gs - 2], dqkv[:, :, gs - 2], dqkv[:, :, gs - 1], delta, dq_acc, cu,
                                max_seqlen, scale, causal)
        _bump(3)
        return dqkv, None, None, None, None


def flash_attention_packed(qkv: torch.Tensor, cu_seqlens, max_seqlen: int, causal: bool = True,
                           scale: Optional[float] = None, impl: Optional[str] = None) -> torch.Tensor:
    """``qkv`` ``[T, kv_heads, q_per_kv + 2, D]`` → ``[T, H, D]``; native kernel on the packed buffer when possible."""
    T, G, gs, D = qkv.shape
    scale = scale or 1.0 / math.sqrt(D)
    impl = impl or _IMPL
    if qkv.is_cuda and impl == "b200" and qkv.dtype == torch.bfloat16 and D == 128 and _b200_available():
        if cu_seqlens.dtype != torch.int32:
            cu_seqlens = cu_seqlens.int()
        return _B200AttnPackedFn.apply(qkv, cu_seqlens.contiguous(), int(max_seqlen), float(scale), bool(causal))
    q = qkv[:, :, : gs - 2].reshape(T, G * (gs - 2), D)
    return flash_attention_varlen(q, qkv[:, :, gs - 2], qkv[:, :, gs - 1], cu_seqlens, max_seqlen, causal, scale, impl)


_b200_attn_ok: Optional[bool] = None


def _b200_available() -> bool:
    """The native attention kernel is mandatory on GPU unless explicitly overridden."""
    return _lib.available()


_warned_dropout = False
_warned_shape = set()


def _note_library_attention(q) -> None:
    """The tcgen05 attention kernels cover bf16 with head_dim 128 (every shipped 7B / 20B config).  Other head dims or
    dtypes run a LIBRARY kernel (flash-attn, else SDPA): said once per (dtype, head_dim) in the log, and refused outright
    when ``B200_STRICT_NATIVE=1`` so a benchmark cannot silently measure library code."""
    key = (str(q.dtype), int(q.shape[-1]))
    if key in _warned_shape:
        return
    _warned_shape.add(key)
    import logging
    import os

    msg = (f"attention with dtype {key[0]} / head_dim {key[1]} is outside the native tcgen05 kernels (bf16, head_dim 128): "
           f"using the flash-attn / SDPA library kernel")
    if os.environ.get("B200_STRICT_NATIVE", "0") == "1":
        raise RuntimeError(msg + " (B200_STRICT_NATIVE=1)")
    logging.getLogger(__name__).warning(msg)


def _attention_with_dropout(q, k, v, cu_seqlens, max_seqlen, causal, scale, dropout_p):
    """Attention-probability dropout (``model.attn_drop_rate`` > 0 while training; reference
    ``multi_head_attention.py:27-53`` passes it to flash-attn).  The tcgen05 kernel has no in-kernel Philox dropout, so this
    knob runs the flash-attn library kernel - said once in the log - or, without it / on CPU, an explicit softmax → dropout
    path per sequence.  It is never silently ignored."""
    global _warned_dropout
    if q.is_cuda:
        try:
            from flash_attn import flash_attn_varlen_func

            if not _warned_dropout:
                _warned_dropout = True
                import logging

                logging.getLogger(__name__).warning(
                    "attn_drop_rate=%.3f: attention dropout runs on the flash-attn library kernel, not the tcgen05 kernel",
                    dropout_p)
            return flash_attn_varlen_func(q, k, v, cu_seqlens.int(), cu_seqlens.int(), int(max_seqlen), int(max_seqlen),
                                          dropout_p=dropout_p, softmax_scale=scale, causal=causal)
        except ImportError:
            pass
    H, Hkv = q.shape[1], k.shape[1]
    outs = []
    cu = cu_seqlens.tolist()
    for a, b in zip(cu[:-1], cu[1:]):
        qs = q[a:b].transpose(0, 1).float()
        ks = k[a:b].transpose(0, 1).float().repeat_interleave(H // Hkv, 0)
        vs = v[a:b].transpose(0, 1).float().repeat_interleave(H // Hkv, 0)
        s_ = torch.einsum("hqd,hkd->hqk", qs, ks) * scale
        if causal:
            n = b - a
            s_ = s_.masked_fill(torch.ones(n, n, dtype=torch.bool, device=q.device).triu(1), float("-inf"))
        p = torch.nn.functional.dropout(torch.softmax(s_, -1), dropout_p, True)
        outs.append(torch.einsum("hqk,hkd->hqd", p, vs).transpose(0, 1).to(q.dtype))
    return torch.cat(outs, 0)


def flash_attention_varlen(q, k, v, cu_seqlens, max_seqlen: int, causal: bool = True, scale: Optional[float] = None,
                           impl: Optional[str] = None, dropout_p: float = 0.0) -> torch.Tensor:
    """q ``[T, H, D]`` bf16, k/v ``[T, Hkv, D]``; returns ``[T, H, D]``."""
    scale = scale or 1.0 / math.sqrt(q.shape[-1])
    impl = impl or _IMPL
    if dropout_p > 0.0:
        return _attention_with_dropout(q, k, v, cu_seqlens, max_seqlen, causal, scale, dropout_p)
    if not q.is_cuda:
        return attention_ref(q, k, v, cu_seqlens, causal, scale).to(q.dtype)
    if cu_seqlens.dtype != torch.int32:
        cu_seqlens = cu_seqlens.int()
    if impl == "b200" and q.dtype == torch.bfloat16 and q.shape[-1] == 128 and _b200_available():
        return _B200AttnFn.apply(q, k, v, cu_seqlens.contiguous(), int(max_seqlen), float(scale), bool(causal))
    if impl == "b200" and q.is_cuda:
        _note_library_attention(q)
    if impl in ("b200", "flash_attn"):
        try:
            from flash_attn import flash_attn_varlen_func

            return flash_attn_varlen_func(q, k, v, cu_seqlens, cu_seqlens, int(max_seqlen), int(max_seqlen),
                                          softmax_scale=scale, causal=causal)
        except ImportError:
            pass
    # torch SDPA per sequence (library fallback)
    H, Hkv = q.shape[1], k.shape[1]
    outs = []
    cu = cu_seqlens.tolist()
    for a, b in zip(cu[:-1], cu[1:]):
        qs, ks, vs = (t[a:b].transpose(0, 1)[None] for t in (q, k, v))
        o = torch.nn.functional.scaled_dot_product_attention(qs, ks, vs, is_causal=causal, scale=scale,
                                                             enable_gqa=H != Hkv)
        outs.append(o[0].transpose(0, 1))
    return torch.cat(outs, 0)


_decode_ws = {}


def decode_attention(q: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, seqlen: int,
                     scale: Optional[float] = None, seqlen_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One generation step: ``q [B, H, D]`` attends to the first ``seqlen`` cache positions (``[B, Smax, Hkv, D]``).
    Native split-KV kernel for bf16 / D = 128, library SDPA otherwise.

    ``seqlen_dev`` (cuda int32 ``[1]``): the valid length is ``seqlen_dev[0] + seqlen`` read ON THE DEVICE — the form used
    inside a captured CUDA graph, where one launch must serve every decode step (the split count is then sized for the
    whole cache; splits beyond the current length produce empty partials)."""
    B, H, D = q.shape
    Hkv = kcache.shape[2]
    scale = scale or 1.0 / math.sqrt(D)
    if q.is_cuda and q.dtype == torch.bfloat16 and D == 128 and _b200_available():
        ctas = B * Hkv
        span = kcache.shape[1] if seqlen_dev is not None else seqlen
        nsplit = max(1, min(32, (2 * 148 + ctas - 1) // ctas, (span + 63) // 64))
        key = (q.device, B, H, nsplit)
        ws = _decode_ws.get(key)
        if ws is None:
            ws = (torch.empty(B * H * nsplit * (D + 4), device=q.device, dtype=torch.float32),
                  torch.zeros(B * Hkv, device=q.device, dtype=torch.int32))
            _decode_ws[key] = ws
        out = torch.empty_like(q)
        torch.ops.b200.attn_decode(q.contiguous(), kcache, vcache, out, ws[0], ws[1], int(seqlen), nsplit, float(scale),
                                   seqlen_dev)
        _bump()
        return out
    assert seqlen_dev is None, "device-side sequence length needs the native decode kernel"
    kk, vv = kcache[:, :seqlen].transpose(1, 2), vcache[:, :seqlen].transpose(1, 2)
    o = torch.nn.functional.scaled_dot_product_attention(q[:, :, None], kk, vv, scale=scale, enable_gqa=H != Hkv)
    return o[:, :, 0]
