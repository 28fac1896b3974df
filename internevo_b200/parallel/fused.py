"""Fused compute+collective back-ends over the symmetric peer-memory heap.

* ``TPFusedBackend``  — tensor-parallel linears: ``ag_gemm`` (all-gather → GEMM, sequence-parallel column linear) and
  ``gemm_rs`` (GEMM → reduce-scatter / all-reduce, row linear).  One kernel launch each: the tcgen05 GEMM runs on most
  SMs while the remaining CTAs of the same grid pull / reduce tiles over NVLink with per-tile flags, so the transfer
  overlaps the math tile by tile (reference sites N1/N3/N4, ``internlm/model/utils.py:25-217``).
* ``ZeroFusedBackend`` — Hybrid-ZeRO step: reduce-scatter by peer loads fused with mean + bf16 cast + grad-norm partials,
  then AdamW fused with the parameter all-gather (bf16 results are stored straight into every peer's arena)
  (reference sites N11/N13, ``internlm/solver/optimizer/hybrid_zero_optim.py:455-523,809-837``).

NCCL implementations of the same operations stay available as fall-back and numerical oracle
(``parallel/linear.py``, ``HybridZeroOptimizer`` without ``fused_comm``).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from internevo_b200.ops.gemm import _bump
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.nvtx import nvtx_range

from . import symm

logger = get_logger(__file__)


class TPFusedBackend:
    """Symmetric scratch + launch logic for the fused TP linears of one process group (CTA-pair tcgen05 tile).

    ``ag_gemm``  the epilogue warps of every CTA first push this rank's activation shard into every peer's gathered buffer
                 (8-row pieces, one flag each) while the tensor cores start on the local rows; a tile of remote rows is
                 loaded as soon as its pieces have landed.
    ``gemm_rs``  the GEMM epilogue stores partial blocks owned by a peer straight into that peer's staging slot over
                 NVLink and reduces the blocks this rank owns (scheduled last) with what the peers pushed; the all-reduce
                 form writes the reduced rows into every rank's output and ends with an in-kernel completion handshake.
    Scratch buffers are ping-pong pairs keyed by shape.  No barrier launch brackets the kernels: every call makes each rank
    wait for data of EVERY peer, so when a rank starts call k all peers have started call k - 1 and therefore finished
    call k - 2 - the last user of the slot call k overwrites."""

    def __init__(self, group: dist.ProcessGroup, comm_ctas: int = 0):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.flags = symm.flags_for(group)
        self._gath: Dict[Tuple[int, int], list] = {}     # (M, K) -> ping-pong gathered-activation buffers
        self._stage: Dict[Tuple[int, int], list] = {}    # (M, N) -> ping-pong [world, M/world, N] staging
        self._out: Dict[Tuple[int, int], list] = {}      # (M, N) -> ping-pong all-reduce outputs
        self._done_counter = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))

    def supports(self, M: int, x: torch.Tensor, weight: torch.Tensor) -> bool:
        """bf16, whole 256-row tiles per rank, 16-byte aligned rows, flag space for every block."""
        return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and M % (256 * self.world) == 0
                and x.shape[1] % 8 == 0 and weight.shape[0] % 8 == 0 and weight.shape[1] % 8 == 0
                and x.stride(-1) == 1 and weight.stride(1) == 1
                and (M // 128) * ((max(weight.shape) + 255) // 256) <= symm.RS_FLAG_WORDS)

    def _pp(self, cache, key, numel):
        if key not in cache:
            cache[key] = [0] + [symm.SymmBuffer(numel, torch.bfloat16, self.group, zero=False) for _ in range(2)]
        ent = cache[key]
        ent[0] ^= 1
        return ent[1 + ent[0]]

    def _pp_counted(self, cache, key, numel, blocks):
        """Ping-pong gathered buffers of one shape plus their arrival counters: one uint32 per 128-row block on every rank,
        16 arrivals per call, never reset - the call's target count is returned (``csrc/gemm_sm100.cu::AG_PARTS``)."""
        if key not in cache:
            cache[key] = [0] + [symm.SymmBuffer(numel, torch.bfloat16, self.group, zero=False) for _ in range(2)] + \
                [symm.SymmBuffer(max(64, blocks), torch.int32, self.group, zero=True), 0]
            torch.cuda.synchronize()
            dist.barrier(self.group)             # every rank's counters are zero before anyone pushes
        ent = cache[key]
        ent[0] ^= 1
        ent[4] += 1
        if os.environ.get("B200_DEBUG_NOWAIT") == "1":     # timing experiments only: never wait for the peers' pieces
            return ent[1 + ent[0]], ent[3].table_ptr(0), 0
        return ent[1 + ent[0]], ent[3].table_ptr(0), (16 * ent[4]) & 0xFFFFFFFF

    # -- all-gather -> GEMM -----------------------------------------------------------------------------------------
    def ag_gemm(self, x: torch.Tensor, weight: torch.Tensor, group=None, keep_gathered: bool = False,
                b_mn: bool = False):
        """``all_gather(x, dim=0) @ weight^T`` (``b_mn``: ``@ weight``); returns ``(y, gathered_x)``.

        ``gathered_x`` lives in a recycled symmetric slot: it is valid until the second next call with the same shape;
        pass ``keep_gathered=True`` to get a private copy (e.g. to save it for backward)."""
        x = x.contiguous()
        m_local, K = x.shape
        N = weight.shape[1] if b_mn else weight.shape[0]
        M = m_local * self.world
        gb, counters, target = self._pp_counted(self._gath, (M, K), M * K, M // 128)
        gathered = gb.tensor.view(M, K)
        out = torch.empty(M, N, device=x.device, dtype=x.dtype)
        if symm.DEBUG:
            self.flags.barrier()
            symm.poison(gathered)
            self.flags.barrier()
        torch.ops.b200.ag_gemm(x, gb.table_ptr(0), counters, self.rank, self.world, target, weight, b_mn, gathered, out, 0,
                               None, 0)
        _bump()
        if symm.DEBUG:
            symm.assert_clean(gathered, "ag_gemm gathered activations")
            symm.assert_clean(out, "ag_gemm output")
        return out, (gathered.clone() if keep_gathered else gathered)

    # -- GEMM -> reduce-scatter / all-reduce ------------------------------------------------------------------------------
    def gemm_rs(self, x: torch.Tensor, weight: torch.Tensor, group=None, all_reduce: bool = False,
                b_mn: bool = False) -> torch.Tensor:
        """``reduce_scatter(x @ weight^T, dim=0)`` (or all-reduce; ``b_mn``: ``x @ weight``)."""
        M, K = x.shape
        N = weight.shape[1] if b_mn else weight.shape[0]
        stage = self._pp(self._stage, (M, N), M * N)
        if all_reduce:
            outb = self._pp(self._out, (M, N), M * N)
            out = outb.tensor.view(M, N)
            out_ptrs = outb.table_ptr(0)
        else:
            out = torch.empty(M // self.world, N, device=x.device, dtype=x.dtype)
            out_ptrs = 0
        if symm.DEBUG:
            self.flags.barrier()
            symm.poison(stage.tensor)
            if all_reduce:
                symm.poison(out)
            self.flags.barrier()
        torch.ops.b200.gemm_rs(x, weight, out, stage.table_ptr(0), out_ptrs, symm.rs_flag_table(self.flags), self.rank,
                               self.world, self.flags.next_epoch(), b_mn, 1 if all_reduce else 0,
                               symm.done_flag_table(self.flags), self._done_counter)
        _bump()
        if all_reduce:
            if symm.DEBUG:
                symm.assert_clean(out, "gemm_rs all-reduce output")
            return out.clone()  # detach from the ping-pong slot (it is recycled two calls later)
        if symm.DEBUG:
            symm.assert_clean(out, "gemm_rs reduce-scatter output")
        return out


_tp_backends: Dict[int, TPFusedBackend] = {}


def enable_tp(group: Optional[dist.ProcessGroup]) -> Optional[TPFusedBackend]:
    """Create (once) the fused backend for ``group`` and route the TP linears through it."""
    if group is None or dist.get_world_size(group) <= 1 or not symm.peer_addressable(group):
        return None
    # Default on: the collective runs inside the CTA-pair tcgen05 GEMM (B200_TP_FUSED=0 selects 2-CTA GEMM + NCCL, which is
    # also the numerical oracle of tests/test_fused_comm_gpu.py).
    if os.environ.get("B200_TP_FUSED", "1") == "0":
        logger.info("fused TP linears disabled (B200_TP_FUSED=0); using the 2-CTA tcgen05 GEMM + NCCL")
        return None
    key = id(group)
    if key not in _tp_backends:
        try:
            _tp_backends[key] = TPFusedBackend(group)
        except Exception as e:  # pragma: no cover - depends on driver / topology (no P2P between the group's GPUs)
            logger.warning(f"fused TP linears unavailable ({e}); using the 2-CTA tcgen05 GEMM + NCCL")
            return None
    from . import linear

    linear.set_fused_backend(_tp_backends[key])
    return _tp_backends[key]


class ISPFusedBackend:
    """Weight-parallel (ISP) linears with the weight all-gather / gradient reduce-scatter INSIDE the tcgen05 GEMM
    (reference sites: ``internlm/core/communication/isp.py:255-297,486-526``, ``internlm/model/utils.py:466-586``).

    ``gather_gemm(x, w_shard)``          y = x @ all_gather(w_shard)^T: every rank's epilogue warps push their weight shard
                                         into the peers' gathered buffers - destination by destination, in the order the
                                         consumers need the shards - while the tensor cores start on the n tiles of the own
                                         shard (read in place); a tile of a remote shard is loaded once its 8-row pieces
                                         have landed.
    ``gather_gemm(dy, w_shard, b_mn)``   dx = dy @ all_gather(w_shard): the gathered rows are the contraction; the k loop
                                         walks the shards own-first and waits at shard boundaries.
    ``wgrad_rs(dy, x, out)``             out (+)= (1 / W) reduce_scatter(dy^T @ x): partial blocks owned by a peer go straight
                                         into that peer's staging slot from the GEMM epilogue, own blocks are reduced and
                                         accumulated into the gradient arena shard.
    No ``all_gather_into_tensor`` / ``reduce_scatter_tensor`` is left on this path.  Gathered buffers and staging slots are
    ping-pong pairs keyed by shape (same recycling argument as ``TPFusedBackend``)."""

    def __init__(self, group: dist.ProcessGroup):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.flags = symm.flags_for(group)
        self._gath: Dict[Tuple[int, int], list] = {}
        self._stage: Dict[Tuple[int, int], list] = {}

    _pp = TPFusedBackend._pp
    _pp_counted = TPFusedBackend._pp_counted

    def supports(self, x: torch.Tensor, w_shard: torch.Tensor) -> bool:
        rows, kin = w_shard.shape
        n_total = rows * self.world
        return (x.is_cuda and x.dtype == torch.bfloat16 and w_shard.dtype == torch.bfloat16 and x.dim() == 2
                and rows % 256 == 0 and kin % 8 == 0 and x.stride(-1) == 1 and w_shard.stride(1) == 1
                and w_shard.stride(0) % 8 == 0
                and (n_total // 128) * ((kin + 255) // 256) <= symm.RS_FLAG_WORDS)

    def prefer(self, op: str) -> bool:
        """Whether the in-kernel form of ``op`` ("fwd" | "dgrad" | "wgrad") is the faster one at this group size - from
        ``profiles/fused_comm_check_n{2,8}_r2_*.json``: the forward wins at 2 and 8 ranks, the wgrad reduce-scatter at 2 (and is
        assumed to at 4), the dgrad only ties at 2 and loses at 8 where the per-rank tiles are few.  ``B200_ISP_FUSED_BWD=1``
        / ``0`` forces both backward forms on / off."""
        force = os.environ.get("B200_ISP_FUSED_BWD", "")
        if op == "fwd" or force == "1":
            return True
        if force == "0":
            return False
        return self.world <= (2 if op == "dgrad" else 4)

    def gather_gemm(self, x: torch.Tensor, w_shard: torch.Tensor, b_mn: bool = False) -> torch.Tensor:
        rows, kin = w_shard.shape
        n_total = rows * self.world
        gb, counters, target = self._pp_counted(self._gath, (n_total, kin), n_total * kin, n_total // 128)
        gathered = gb.tensor.view(n_total, kin)
        out = torch.empty(x.shape[0], kin if b_mn else n_total, device=x.device, dtype=x.dtype)
        if symm.DEBUG:
            self.flags.barrier()
            symm.poison(gathered)
            self.flags.barrier()
        torch.ops.b200.gather_weight_gemm(x, w_shard, gb.table_ptr(0), gathered, counters, self.rank, self.world, target,
                                          b_mn, out)
        _bump()
        if symm.DEBUG:
            symm.assert_clean(out, "isp gather_gemm output")
        return out

    def wgrad_rs(self, dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, accumulate: bool) -> None:
        """``out [N_total / W, K_in]`` (a gradient-arena view or a fresh tensor) ``(+)= mean over the group of dy^T @ x``."""
        n_total, kin = dy.shape[1], x.shape[1]
        stage = self._pp(self._stage, (n_total, kin), n_total * kin)
        if symm.DEBUG:
            self.flags.barrier()
            symm.poison(stage.tensor)
            self.flags.barrier()
        torch.ops.b200.wgrad_rs(dy, x, out, stage.table_ptr(0), symm.rs_flag_table(self.flags), self.rank, self.world,
                                self.flags.next_epoch(), 1.0 / self.world, accumulate)
        _bump()


_isp_backends: Dict[int, ISPFusedBackend] = {}


def isp_backend(group: Optional[dist.ProcessGroup]) -> Optional[ISPFusedBackend]:
    """The fused weight-parallel backend of ``group`` (created on first use); ``None`` when peer memory is unavailable, the
    group is trivial or ``B200_ISP_FUSED=0`` (NCCL prefetch path, also the numerical oracle)."""
    if group is None or dist.get_world_size(group) not in (2, 4, 8) or not symm.peer_addressable(group):
        return None
    if os.environ.get("B200_ISP_FUSED", "1") == "0":
        return None
    key = id(group)
    if key not in _isp_backends:
        try:
            _isp_backends[key] = ISPFusedBackend(group)
        except Exception as e:  # pragma: no cover - depends on driver / topology
            logger.warning(f"fused ISP linears unavailable ({e}); using NCCL")
            return None
    return _isp_backends[key]


class ZeroFusedBackend:
    """Peer-memory primitives of ``HybridZeroOptimizer`` for groups whose ZeRO group is the DP group: the optimizer decides
    WHEN a range is reduced / updated (during backward / ahead of the next forward), this class does it over NVLink.

    ``reduce_range``  side stream: device barrier (every rank's backward has written the range), then ``rs_reduce``: this rank's
                      sub-slice = mean over the peers' arenas (P2P loads), cast to bf16 in place, Σ g² accumulated on the way.
    ``update_range``  unscale + clip + AdamW on the fp32 master sub-slice; the bf16 result is stored straight into EVERY
                      peer's parameter arena (the all-gather is the store), then a device barrier so that the event the next
                      forward waits on covers all peers' stores.
    The flag words and the epoch counter are private to this backend: its barriers run on the optimizer's side stream and
    must never share a counter with the main-stream barriers of the TP / MoE back-ends."""

    def __init__(self, opt):
        from internevo_b200.core.context import global_context as gpc

        self.gpc = gpc
        self.groups = {}
        for g in opt.groups:
            if not g.params or g.zero_size <= 1 or g.dtype is not torch.bfloat16:
                continue
            same = g.zero_size == g.dp_size and gpc.get_ranks_in_group(g.dp_mode) == gpc.get_ranks_in_group(g.zero_mode)
            if not same or g.zero_size not in (2, 4, 8):
                continue
            group = gpc.get_group(g.zero_mode)
            if not symm.group_is_intra_node(group):
                continue        # the ZeRO group leaves the node: its reduce-scatter / all-gather stay on NCCL
            # move both arenas into symmetric memory (parameters and grad_buf views are re-pointed)
            pbuf = symm.SymmBuffer(g.total, torch.bfloat16, group, zero=False)
            gbuf = symm.SymmBuffer(g.total, torch.bfloat16, group, zero=True)
            pbuf.tensor.copy_(g.param_arena)
            g.param_arena, g.grad_arena = pbuf.tensor, gbuf.tensor
            for p in g.ordered:
                o = g.offsets[id(p)]
                p.data = g.param_arena[o: o + p.numel()].view(p.shape)
                p.grad_buf = g.grad_arena[o: o + p.numel()].view(p.shape)
            self.groups[g.gid] = (pbuf, gbuf, symm.SymmFlags(group, words=1024))
        # NVLS variant: reduce in the switch (multimem.ld_reduce) and broadcast by one multicast store (multimem.st).  Opt-in:
        # a reduce-SCATTER / all-GATHER moves (W-1)/W of the arena over every GPU's links either way (only an all-reduce
        # halves its traffic in the switch), and measured at 2 GPUs it is slower than the unicast kernels (0.212 vs 0.114 ms
        # per 64 Mi elements, profiles/fused_comm_check_n2_r1_v3.json), so peer loads / stores stay the default.
        self.use_mc = os.environ.get("B200_ZERO_NVLS", "0") == "1" and bool(self.groups) and all(
            pb.mc_ptr and gb.mc_ptr for pb, gb, _ in self.groups.values())
        if gpc.is_rank_for_log():
            logger.info(f"fused Hybrid-ZeRO over peer memory enabled for groups {[opt.groups[i].name for i in self.groups]}"
                        f" (NVLS multicast: {self.use_mc})")

    @staticmethod
    def try_create(opt):
        if not symm.symm_available():
            return None
        try:
            be = ZeroFusedBackend(opt)
        except Exception as e:  # pragma: no cover - depends on driver / topology
            logger.warning(f"fused ZeRO backend unavailable ({e}); using NCCL")
            return None
        return be if be.groups else None

    # ---- phase 0 -------------------------------------------------------------------------------------------------------
    def _launch(self, g, i, hyper, phase):
        pbuf, gbuf, flags = self.groups[g.gid]
        a, m, n = g.sub(i)
        if n == 0:
            return
        mst, ea, es = g.master[m: m + n], g.exp_avg[m: m + n], g.exp_avg_sq[m: m + n]
        if self.use_mc:
            torch.ops.b200.reduce_scatter_adam_mc(gbuf.mc_ptr, pbuf.mc_ptr, gbuf.tensor.data_ptr(), g.zero_size, a, n, mst, ea,
                                                  es, g.scalars, *hyper, phase)
        else:
            torch.ops.b200.reduce_scatter_adam(gbuf.table_ptr(0), pbuf.table_ptr(0), flags.table_ptr(0), g.zero_rank,
                                               g.zero_size, 0, a, n, mst, ea, es, g.scalars, *hyper, phase)
        _bump()

    def reduce_range(self, opt, g, i):
        _, _, flags = self.groups[g.gid]
        side, main = opt._side(), torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)                      # the kernels that wrote this range's gradients are ahead of this point
        side.wait_event(ev)
        with torch.cuda.stream(side), nvtx_range("zero.rs_reduce"):
            if not g.scalars_fresh:
                g.scalars.zero_()
                g.scalars_fresh = True
            flags.barrier()                  # ... on every rank
            self._launch(g, i, (0.0, 0.9, 0.95, 1e-8, 0.0, 1.0, 1.0, float(g.zero_size)), 0)

    def join(self, opt):
        torch.cuda.current_stream().wait_stream(opt._side())

    def local_sumsq(self, opt, g, count_replica: bool):
        """Σ grad² of the owned sub-slices: the reduce kernels left it in ``scalars[3]``; replica parameters are counted on one
        tensor rank only, so the others subtract their (tiny) replica sub-slices again."""
        g.sumsq.copy_(g.scalars[3:4])
        if not count_replica:
            from internevo_b200 import ops

            rep = torch.zeros(1, device=g.sumsq.device)
            for i in range(g.n_sharded_ranges, len(g.ranges)):
                a, m, n = g.sub(i)
                if n:
                    ops.sumsq_(g.grad_arena[a: a + n], rep)
            g.sumsq -= rep

    # ---- phase 1 -------------------------------------------------------------------------------------------------------
    def update_range(self, opt, g, i, hyper):
        lr, beta1, beta2, eps, wd = hyper
        _, _, flags = self.groups[g.gid]
        with nvtx_range("zero.adamw_bcast"):
            self._launch(g, i, (lr, beta1, beta2, eps, wd, 1.0 - beta1 ** g.step, 1.0 - beta2 ** g.step, 1.0), 1)
            flags.barrier()   # every peer's stores of this range have landed before anyone's forward reads it

    def after_update(self, opt, g):
        """Serial path: nothing left to do (every range ended with its own barrier)."""
