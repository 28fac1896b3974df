"""Sequence-parallel flash attention with in-kernel peer K / V (``csrc/attention_sm100.cu``, ``SP`` instantiations).

Every rank of the sequence group holds ``T / sp`` consecutive token rows of the packed stream (``cu_seqlens`` stay global).
The reference gets full sequences per head with three all-to-alls in front of flash-attn and one behind it
(Ulysses, ``internlm/model/modules/multi_head_attention.py:27-135``), which also caps ``sp`` at the number of kv heads.
Here nothing is exchanged ahead of the kernel:

* forward: K and V (after RoPE) are written into a symmetric slab; the kernel's TMA producer walks the key tiles of a
  sequence ACROSS rank boundaries - one tensor map per peer, the tile's owner is ``global_row / T_local`` - inside a single
  online-softmax pass, so there are no ring steps and no LSE merge, and any ``sp <= 8`` works with any head count;
* backward: a rank owns its key / value tiles; the query tiles they meet (its own and the later ranks' rows) are read from
  their owners - Q / dO through per-peer tensor maps, lse / delta through peer pointers - and dQ is reduce-added into the
  owner's fp32 accumulator by a TMA reduce over NVLink.

Slabs are ping-pong pairs; one device barrier per forward (K / V in place) and two per backward (Q / dO / stats in place,
all dQ contributions landed) order the phases - the same recycling argument as ``fused.TPFusedBackend``.

With a contiguous split causal attention is unbalanced across ranks for ONE long sequence (the last rank meets ``sp`` times
the keys of the first); packed batches of many shorter sequences - the SFT regime this is for - are balanced.  The layer picks
this path or the all-to-all path by ``B200_SP_ATTN`` (see ``models/modules.py``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from internevo_b200.ops.gemm import _bump
from internevo_b200.utils.logger import get_logger

from . import symm

logger = get_logger(__file__)


class SPAttentionBackend:
    def __init__(self, group: dist.ProcessGroup, t_local: int, H: int, Hkv: int, D: int):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.t_local, self.H, self.Hkv, self.D = t_local, H, Hkv, D
        self.flags = symm.flags_for(group)
        bf, f32 = torch.bfloat16, torch.float32
        mk = lambda n, dt: [symm.SymmBuffer(n, dt, group, zero=True) for _ in range(2)]  # noqa: E731
        self.kv = mk(2 * t_local * Hkv * D, bf)          # [2, T, Hkv, D]
        self.q = mk(t_local * H * D, bf)
        self.dout = mk(t_local * H * D, bf)
        self.stats = mk(2 * H * t_local, f32)            # [2, H, T]: rowsum(dO * O), lse * log2(e)
        self.dq_acc = mk(t_local * H * D, f32)
        self._fwd_i = 0
        self._bwd_i = 0

    def ptrs(self, buf: symm.SymmBuffer, elem_offset: int = 0):
        return [p + elem_offset * buf.elem for p in buf.base_ptrs]


_backends: Dict[Tuple, SPAttentionBackend] = {}


def backend_for(group, t_local: int, H: int, Hkv: int, D: int) -> Optional[SPAttentionBackend]:
    if group is None or dist.get_world_size(group) not in (2, 4, 8) or not symm.peer_addressable(group):
        return None
    key = (id(group), t_local, H, Hkv, D)
    if key not in _backends:
        try:
            _backends[key] = SPAttentionBackend(group, t_local, H, Hkv, D)
        except Exception as e:  # pragma: no cover - depends on driver / topology (no P2P between the group's GPUs)
            logger.warning(f"peer-K/V attention unavailable ({e}); using the all-to-all (Ulysses) form")
            _backends[key] = None
    return _backends[key]


def reset() -> None:
    _backends.clear()


class _SPAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, scale, causal, be: SPAttentionBackend):
        T, H, D = q.shape[0], be.H, be.D
        slab = be.kv[be._fwd_i]
        be._fwd_i ^= 1
        kv = slab.tensor.view(2, T, be.Hkv, D)
        kv[0].copy_(k)
        kv[1].copy_(v)
        be.flags.barrier()                      # every rank's K / V are in place
        out = torch.empty(T, H, D, device=q.device, dtype=q.dtype)
        lse = torch.empty(H, T, device=q.device, dtype=torch.float32)
        torch.ops.b200.attn_fwd_sp(q, kv[0], kv[1], out, lse, cu_seqlens, int(max_seqlen), float(scale), bool(causal),
                                   be.rank, be.ptrs(slab, 0), be.ptrs(slab, T * be.Hkv * D))
        _bump()
        ctx.save_for_backward(q, k, v, out, lse, cu_seqlens)
        ctx.cfg = (int(max_seqlen), float(scale), bool(causal), be)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, cu = ctx.saved_tensors
        max_seqlen, scale, causal, be = ctx.cfg
        T, H, D = q.shape[0], be.H, be.D
        i = be._bwd_i
        be._bwd_i ^= 1
        qs = be.q[i].tensor.view(T, H, D)
        ds = be.dout[i].tensor.view(T, H, D)
        stats = be.stats[i].tensor.view(2, H, T)
        acc = be.dq_acc[i].tensor.view(T, H, D)
        qs.copy_(q.reshape(T, H, D))
        ds.copy_(dout)
        acc.zero_()
        dq = torch.empty(T, H, D, device=q.device, dtype=q.dtype)
        dk = torch.empty(T, be.Hkv, D, device=q.device, dtype=q.dtype)
        dv = torch.empty_like(dk)
        peers = (be.ptrs(be.q[i]), be.ptrs(be.dout[i]), be.ptrs(be.dq_acc[i]), be.ptrs(be.stats[i]))

        def run(phase):
            torch.ops.b200.attn_bwd_sp(ds, qs, k, v, out, lse, dq, dk, dv, stats, acc, cu, max_seqlen, scale, causal, phase,
                                       be.rank, *peers)
            _bump()

        run(1)                                   # rowsum(dO * O) and lse * log2e into the symmetric stats slab
        be.flags.barrier()                       # Q / dO / stats of every rank are in place, accumulators are zero
        run(2)
        be.flags.barrier()                       # every contribution to this rank's dQ accumulator has landed
        run(3)
        return dq.view_as(q), dk.view_as(k), dv.view_as(v), None, None, None, None, None


def sp_flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int,
                       group, causal: bool = True, scale: Optional[float] = None) -> Optional[torch.Tensor]:
    """``q [T_local, H, D]``, ``k / v [T_local, Hkv, D]`` (this rank's rows), GLOBAL ``cu_seqlens`` → ``[T_local, H, D]``;
    ``None`` when the shapes are not supported (the caller falls back to the all-to-all path)."""
    T, H, D = q.shape[0], q.shape[-2] if q.dim() == 3 else q.shape[1] * q.shape[2], q.shape[-1]
    if not (q.is_cuda and q.dtype == torch.bfloat16 and D == 128 and T % 128 == 0):
        return None
    if q.dim() == 4:
        q = q.reshape(T, -1, D)
    be = backend_for(group, T, q.shape[1], k.shape[1], D)
    if be is None:
        return None
    if cu_seqlens.dtype != torch.int32:
        cu_seqlens = cu_seqlens.int()
    return _SPAttnFn.apply(q, k, v, cu_seqlens.contiguous(), int(max_seqlen), scale or 1.0 / math.sqrt(D), causal, be)
