"""Hybrid-ZeRO (ZeRO-1.5) optimizer on persistent arenas, overlapped with backward and with the next forward.

Semantics preserved from the reference (``internlm/solver/optimizer/hybrid_zero_optim.py:55-950``): parameter groups with
their own reduction group (DATA / WEIGHT_DATA / EXPERT_DATA), optimizer state + fp32 master sharded over the ZeRO
sub-group while gradients are averaged over the whole data-parallel group, dynamic loss scale with skip-on-overflow,
per-group grad-norm + clipping with replica parameters counted once, updated parameters redistributed inside the ZeRO
group, gradient reduction overlapped with backward (``overlap_sync_grad``, reference ``:290-365,427-523``), parameter
redistribution overlapped with the next forward (``overlap_sync_param``, reference ``core/communication/utils.py:134-235``),
resumable ``state_dict``.

Redesign (what changes on a B200 node):

* every group owns ONE contiguous low-precision parameter arena and ONE gradient arena for the life of the job; model
  parameters are views into the first and the wgrad GEMM epilogues / norm-backward kernels accumulate straight into
  the second (``param.grad_buf``).  The reference's per-step flatten → all-reduce → unflatten → copy chain
  (``store.py:315-322``, ``hybrid_zero_optim.py:455-523,740-797``) disappears;
* the arena is cut into RANGES (whole parameters, about one transformer block each, ``reduce_bucket_size`` caps them) and
  every range is sharded by element over the ZeRO group: rank ``r`` owns sub-slice ``r`` of EVERY range
  (range-interleaved ownership).  The fp32 master and the moments are the concatenation of the owned sub-slices.  Every
  rank therefore has work as soon as ANY range has its final gradient, which is what makes the overlap below possible;
* gradient sync: the kernels that write a parameter's final gradient of the step (last micro-batch) call
  ``param.grad_hook``; when all parameters of the next range in the fixed launch order (arena order reversed = backward
  order) have reported, its reduce-scatter is issued on a side stream while the backward of the earlier blocks continues.
  The order is the same on every rank whatever the timing, so the collectives / device barriers always match up;
  ranges whose hooks never fire are launched from ``step``;
* unscale + clip + AdamW + bf16 cast-back are ONE kernel per range whose multiplier / skip flag live on the device; the
  update and the parameter all-gather of range ``c`` run on the side stream and the forward of the NEXT step waits, block
  by block, only for the ranges that hold that block's parameters;
* with a peer-memory heap (``parallel/symm.py``) the reduce-scatter is a peer-load kernel fused with mean + cast +
  grad-norm partials and the all-gather is the store of the AdamW kernel into every peer's arena
  (``parallel/fused.py::ZeroFusedBackend``) — NCCL is the fallback and the oracle.
"""
from __future__ import annotations

import bisect
import math
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from internevo_b200 import ops
from internevo_b200.core.context import (
    IS_REPLICA_ZERO_PARALLEL,
    IS_TENSOR_DATA_PARALLEL,
    IS_TENSOR_EXPERT_DATA_PARALLEL,
    ParallelMode,
)
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import get_current_device
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.megatron_timers import megatron_timer as timer
from internevo_b200.utils.parallel import is_using_isp, is_using_sequence_parallel

from .utils import DynamicGradScaler

logger = get_logger(__file__)
_ALIGN = 128  # elements; keeps every parameter view 256-byte aligned for TMA descriptors and 16-byte vector access


def _group_size(mode: ParallelMode) -> int:
    return gpc.get_world_size(mode) if gpc.is_initialized(mode) else 1


def _all_reduce_avg(t: torch.Tensor, mode: ParallelMode):
    group, n = gpc.get_group(mode), _group_size(mode)
    if group is None or n <= 1:
        return
    if t.is_cuda:
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo has no AVG
        dist.all_reduce(t, group=group)
        t.div_(n)


class _GroupState:
    """Arena bookkeeping of one parameter group (see the module docstring for the range-interleaved layout)."""

    def __init__(self, gid: int, cfg: dict, params: List[torch.nn.Parameter], dp_mode: ParallelMode,
                 zero_mode: ParallelMode, device, bucket_elems: int):
        self.gid, self.cfg, self.params = gid, cfg, params
        self.name = cfg.get("name", f"group{gid}")
        self.dp_mode, self.zero_mode = dp_mode, zero_mode
        self.zero_size = _group_size(zero_mode)
        self.zero_rank = gpc.get_local_rank(zero_mode) if gpc.is_initialized(zero_mode) else 0
        self.dp_size = _group_size(dp_mode)
        self.dtype = params[0].dtype if params else torch.float32
        W = self.zero_size
        align = math.lcm(_ALIGN, 8 * W)   # every range (whole parameters) splits into W sub-slices of a multiple of 8 elements
        # sharded params first, replica params (norm weights / gates: identical on every TP rank) last
        rep = [p for p in params if getattr(p, IS_REPLICA_ZERO_PARALLEL, False)]
        shd = [p for p in params if not getattr(p, IS_REPLICA_ZERO_PARALLEL, False)]
        self.ordered = shd + rep
        off = 0
        self.offsets: Dict[int, int] = {}
        for p in shd:
            self.offsets[id(p)] = off
            off += (p.numel() + align - 1) // align * align
        self.replica_start = off
        for p in rep:
            self.offsets[id(p)] = off
            off += (p.numel() + align - 1) // align * align
        self.replica_end = off     # [replica_start, replica_end): same SIZE on every tensor rank (offsets may differ)
        quantum = math.lcm(W * 1024, align)
        self.total = max(quantum, (off + quantum - 1) // quantum * quantum)
        self.shard = self.total // W                      # elements of master / moments on this rank
        # ---- ranges: cut at parameter starts once `target` elements have accumulated; never across replica_start
        target = max(align, min(int(bucket_elems), max(1 << 22, self.replica_start // 48)))
        cuts = [0]
        for p in shd[1:]:
            s = self.offsets[id(p)]
            if s - cuts[-1] >= target:
                cuts.append(s)
        if self.replica_start > cuts[-1] and self.replica_start < self.total:
            cuts.append(self.replica_start)
        cuts.append(self.total)
        self.ranges: List[Tuple[int, int]] = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        self.n_sharded_ranges = sum(1 for lo, _ in self.ranges if lo < self.replica_start)
        self.replica_moff = self.replica_start // W if self.replica_start < self.total else self.shard
        # launch order of the gradient reduction = expected completion order of the backward: sharded ranges from the end
        # of the arena (head, last block) to its start (first block, embedding), the replica tail (norm weights of ALL blocks)
        # last
        self.order = list(range(self.n_sharded_ranges - 1, -1, -1)) + list(range(self.n_sharded_ranges, len(self.ranges)))
        self.range_of: Dict[int, int] = {}
        counts = [0] * len(self.ranges)
        starts = [lo for lo, _ in self.ranges]
        for p in self.ordered:
            i = bisect.bisect_right(starts, self.offsets[id(p)]) - 1
            self.range_of[id(p)] = i
            counts[i] += 1
        self.range_params = counts
        self.pending = list(counts)      # parameters of each range still waiting for their final gradient this step
        self.next = 0                    # position in `order` of the next range to launch
        self.launched = [False] * len(self.ranges)
        self.handles: list = []          # in-flight NCCL work of this step's gradient reduction
        self.copy_back: List[int] = []   # ranges reduced by all-reduce whose owned sub-slice still goes to `gshard`
        self.scalars_fresh = False       # fused path: scalars zeroed on the side stream for this step
        self.param_arena = torch.zeros(self.total, dtype=self.dtype, device=device)
        self.grad_arena = torch.zeros(self.total, dtype=self.dtype, device=device)
        for p in self.ordered:
            o = self.offsets[id(p)]
            view = self.param_arena[o: o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad_buf = self.grad_arena[o: o + p.numel()].view(p.shape)
            p.grad_ready = False
        # fp32 master + moments of the owned sub-slices (compact: range i lives at [lo_i / W, hi_i / W))
        self.master = torch.empty(self.shard, dtype=torch.float32, device=device)
        self.pull_master()
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.gshard: Optional[torch.Tensor] = None   # compact reduced gradient of the owned sub-slices (NCCL / gloo path)
        self.step = 0
        self.scalars = torch.zeros(4, dtype=torch.float32, device=device)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self.events = None      # per-range CUDA events of the overlapped update (HybridZeroOptimizer._ensure_events)
        self.ready_ev = None

    # -- geometry ---------------------------------------------------------------------------------------------------
    def sub(self, i: int) -> Tuple[int, int, int]:
        """``(arena offset, compact offset, length)`` of this rank's sub-slice of range ``i``."""
        lo, hi = self.ranges[i]
        n = (hi - lo) // self.zero_size
        return lo + self.zero_rank * n, lo // self.zero_size, n

    def is_replica_range(self, i: int) -> bool:
        return i >= self.n_sharded_ranges

    def owner_of(self, p) -> int:
        """Rank of the ZeRO group whose sub-slice holds the first element of ``p`` (a parameter larger than a sub-slice
        continues on the following ranks)."""
        lo, hi = self.ranges[self.range_of[id(p)]]
        return (self.offsets[id(p)] - lo) // ((hi - lo) // self.zero_size)

    def owned_views(self, arena: torch.Tensor):
        for i in range(len(self.ranges)):
            a, m, n = self.sub(i)
            yield arena[a: a + n], m, n

    def pull_master(self):
        """fp32 master ← the owned sub-slices of the low-precision arena."""
        for v, m, n in self.owned_views(self.param_arena):
            self.master[m: m + n].copy_(v)

    def push_master(self):
        """Owned sub-slices of the low-precision arena ← fp32 master."""
        for v, m, n in self.owned_views(self.param_arena):
            v.copy_(self.master[m: m + n])

    def grad_shard(self) -> torch.Tensor:
        """Compact reduced gradient of the owned sub-slices; with a ZeRO group of one the arena itself (same offsets)."""
        if self.zero_size == 1:
            return self.grad_arena
        if self.gshard is None:
            self.gshard = torch.zeros(self.shard, dtype=self.dtype, device=self.grad_arena.device)
        return self.gshard

    def reset_step_state(self):
        self.pending = list(self.range_params)
        self.next = 0
        self.launched = [False] * len(self.ranges)
        self.scalars_fresh = False


class HybridZeroOptimizer:
    """See module docstring. ``optimizer`` supplies ``param_groups`` (any object with that attribute, e.g. a
    ``torch.optim.AdamW`` built by ``initialize_optimizer``)."""

    def __init__(self, optimizer, cpu_offload=False, grad_scal_cfg=None, zero_cfg=None, param_bcast_sync_handler=None,
                 isp_communicator=None, use_fused_comm: Optional[bool] = None):
        assert not cpu_offload, "cpu_offload is not supported (180 GB HBM per GPU: keep optimizer state resident)"
        grad_scal_cfg = grad_scal_cfg or {}
        zero_cfg = zero_cfg or {}
        fp16_cfg = grad_scal_cfg.get("fp16", {}) if hasattr(grad_scal_cfg, "get") else {}
        self.param_groups = optimizer.param_groups
        self.optim = optimizer
        self._dtype = self.param_groups[0]["params"][0].dtype
        self.use_isp = is_using_isp()
        self._isp_communicator = isp_communicator
        self._clip_grad_norm = zero_cfg.get("clip_grad_norm", 0.0)
        # overlap_sync_grad: reduce a range as soon as its gradients are final (during backward) instead of inside step();
        # overlap_sync_param: keep the parameter all-gather of the NCCL path asynchronous until the owning block runs;
        # reduce_bucket_size: upper bound (elements) of one range.  B200_ZERO_OVERLAP=0 forces the serial path.
        self._overlap_sync_grad = bool(zero_cfg.get("overlap_sync_grad", False)) and \
            os.environ.get("B200_ZERO_OVERLAP", "1") != "0"
        self._overlap_sync_param = bool(zero_cfg.get("overlap_sync_param", False))
        self._reduce_bucket_size = int(zero_cfg.get("reduce_bucket_size", 512 * 1024 * 1024))
        self.skip_grad_reduce = False
        self.device = get_current_device()
        if self._dtype is torch.float32:
            self.grad_scaler = DynamicGradScaler(initial_scale=1, growth_factor=2, backoff_factor=0.5,
                                                 growth_interval=10**12, min_scale=1, max_scale=1, hysteresis=2)
        else:
            self.grad_scaler = DynamicGradScaler(
                initial_scale=fp16_cfg.get("initial_scale", 2**16), min_scale=fp16_cfg.get("min_scale", 1),
                growth_interval=fp16_cfg.get("growth_interval", 1000),
                growth_factor=grad_scal_cfg.get("growth_factor", 2), backoff_factor=grad_scal_cfg.get("backoff_factor", 0.5),
                max_scale=grad_scal_cfg.get("max_scale", 2**24), hysteresis=grad_scal_cfg.get("hysteresis", 2))
        self.groups: List[_GroupState] = []
        for gid, pg in enumerate(self.param_groups):
            params = [p for p in pg["params"] if p.requires_grad]
            dp_mode, zero_mode = self._modes_for_group(pg, params)
            self.groups.append(_GroupState(gid, pg, params, dp_mode, zero_mode, self.device, self._reduce_bucket_size))
        self.rank_unique_id = (
            f"gpus-{gpc.get_world_size(ParallelMode.GLOBAL)}_wp-{gpc.get_local_rank(ParallelMode.WEIGHT)}_"
            f"tp-{gpc.get_local_rank(ParallelMode.TENSOR)}_dp-{gpc.get_local_rank(ParallelMode.DATA)}_"
            f"pp-{gpc.get_local_rank(ParallelMode.PIPELINE)}_zo-{gpc.get_local_rank(ParallelMode.ZERO1)}.pt"
        )
        self._fused = None
        if use_fused_comm is None:
            use_fused_comm = bool(gpc.config.get("fused_comm", False)) if gpc.config is not None else False
        if use_fused_comm and torch.cuda.is_available():
            from internevo_b200.parallel import fused

            self._fused = fused.ZeroFusedBackend.try_create(self)
        self.has_params = sum(len(g.params) for g in self.groups) > 0
        # parameter update of step s overlapped with the forward of step s + 1 (CUDA; see _update_overlapped)
        self._adam_overlap = os.environ.get("B200_ADAM_OVERLAP", "1") != "0" and torch.cuda.is_available() \
            and not self.use_isp
        self._side_stream = None
        self._owner_events: Dict[int, list] = {}     # id(module) -> events of the ranges holding its parameters
        self._owner_seen: Dict[int, int] = {}        # id(module) -> update generation it has already waited for
        self._update_gen = 0
        self._model_attached = False
        self._model, self._param_names = None, {}       # bind_model(): names for the checkpoint plan / the state converters
        self._pp_group_names = None   # union of parameter-group names over the pipeline group (agreed at the first step)
        self._group_of: Dict[int, _GroupState] = {}
        self.overlap_stats = {"hook_launches": 0, "step_launches": 0}   # ranges reduced during backward / inside step()
        self._install_grad_hooks()

    # ------------------------------------------------------------------------------------------------------------
    def _modes_for_group(self, pg, params):
        """Reduction / sharding groups per parameter class (reference ``train/utils.py:40-79`` + ``optimizer/utils.py``)."""
        name = pg.get("name", "default")
        if "optimizer_mode" in pg and name.startswith("moe"):
            return ParallelMode.EXPERT_DATA, ParallelMode.EXPERT_DATA
        if params and getattr(params[0], IS_TENSOR_EXPERT_DATA_PARALLEL, False):
            return ParallelMode.EXPERT_DATA, ParallelMode.EXPERT_DATA
        if self.use_isp:
            if name == "embed_head" or (params and getattr(params[0], IS_TENSOR_DATA_PARALLEL, False)):
                return ParallelMode.DATA, ParallelMode.DATA
            return ParallelMode.WEIGHT_DATA, ParallelMode.ZERO1
        return ParallelMode.DATA, ParallelMode.ZERO1

    @property
    def dtype(self):
        return self._dtype

    @property
    def loss_scale(self):
        return torch.tensor(self.grad_scaler.scale, dtype=torch.float32)

    @property
    def num_param_groups(self):
        return len(self.param_groups)

    # ------------------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        """Gradients are *overwritten* by the first micro-batch's wgrad epilogue, so nothing is memset here."""
        for g in self.groups:
            for p in g.params:
                p.grad = None
                p.grad_ready = False
            g.reset_step_state()

    def backward(self, loss, retain_graph=False):
        (loss * self.grad_scaler.scale).backward(retain_graph=retain_graph)

    def backward_by_grad(self, tensor, grad):
        torch.autograd.backward(tensors=tensor, grad_tensors=grad)

    def wait_param_sync(self):
        """Kept for the scheduler protocol: parameter redistribution is tracked per range by CUDA events consumed in the
        blocks' pre-forward hooks (or completed inside ``step`` when the overlap is off), so there is nothing to wait for."""

    # ---- gradient readiness ------------------------------------------------------------------------------------------------
    def _install_grad_hooks(self):
        """``param.grad_hook(param)`` is called by the kernels that accumulate straight into ``grad_buf`` (wgrad GEMM
        epilogue, norm backward); parameters whose gradient arrives through autograd (``.grad``: embedding, biases, MoE gates)
        get a post-accumulate hook that folds it into the arena right away and then reports the same way."""
        for g in self.groups:
            for p in g.params:
                self._group_of[id(p)] = g
                p.grad_hook = self._on_grad_ready
                if hasattr(p, "register_post_accumulate_grad_hook"):
                    p.register_post_accumulate_grad_hook(self._on_autograd_grad)

    @staticmethod
    def _fold_autograd_grad(p):
        if p.grad is None:
            return
        if p.grad_ready:
            p.grad_buf.add_(p.grad)
        else:
            p.grad_buf.copy_(p.grad)
        p.grad = None
        p.grad_ready = True

    def _on_autograd_grad(self, p):
        if p.grad is None:
            return
        self._fold_autograd_grad(p)
        self._on_grad_ready(p)

    def _on_grad_ready(self, p):
        """A parameter's gradient of the current micro-batch is in the arena.  On the LAST micro-batch of the step (the
        schedulers clear ``skip_grad_reduce`` for it, as in the reference) this is the final value: count it and launch every
        range of the fixed order that is now complete."""
        if self.skip_grad_reduce or not self._overlap_sync_grad:
            return
        g = self._group_of.get(id(p))
        if g is None or not self._can_overlap_reduce(g):
            return
        g.pending[g.range_of[id(p)]] -= 1
        self._advance(g, final=False)

    def _can_overlap_reduce(self, g: _GroupState) -> bool:
        # expert gradients are pre-reduced over the tensor group inside step(); a group without data parallelism has
        # nothing to reduce
        return g.dp_size > 1 and g.dp_mode is not ParallelMode.EXPERT_DATA and not self.use_isp

    def _range_is_late(self, g: _GroupState, i: int) -> bool:
        """Replica ranges under sequence parallelism need their tensor-group reduction first (done in step())."""
        return g.is_replica_range(i) and (is_using_sequence_parallel() or self.use_isp) and \
            _group_size(ParallelMode.WEIGHT if self.use_isp else ParallelMode.TENSOR) > 1

    def _advance(self, g: _GroupState, final: bool):
        while g.next < len(g.order):
            i = g.order[g.next]
            if not final and (g.pending[i] > 0 or self._range_is_late(g, i)):
                return
            self._launch_reduce(g, i)
            self.overlap_stats["step_launches" if final else "hook_launches"] += 1
            g.next += 1

    # ---- gradient reduction of one range -----------------------------------------------------------------------------------
    def _side(self):
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        return self._side_stream

    def _launch_reduce(self, g: _GroupState, i: int):
        """Average range ``i`` over the data-parallel group; afterwards this rank's sub-slice holds the reduced values (in the
        arena for the peer-memory path, in ``g.gshard`` otherwise)."""
        g.launched[i] = True
        if g.dp_size <= 1:
            return
        lo, hi = g.ranges[i]
        a, m, n = g.sub(i)
        dp_group = gpc.get_group(g.dp_mode)
        if self._fused is not None and g.gid in self._fused.groups:
            self._fused.reduce_range(self, g, i)
            return
        same = g.zero_size == g.dp_size and gpc.get_ranks_in_group(g.dp_mode) == gpc.get_ranks_in_group(g.zero_mode)
        seg = g.grad_arena[lo:hi]
        if seg.is_cuda and same:
            h = dist.reduce_scatter_tensor(g.grad_shard()[m: m + n], seg, op=dist.ReduceOp.AVG, group=dp_group,
                                           async_op=True)
            g.handles.append(h)
        elif seg.is_cuda:
            g.handles.append(dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=dp_group, async_op=True))
            g.copy_back.append(i)
        else:  # gloo: no AVG, no reduce_scatter_tensor
            dist.all_reduce(seg, group=dp_group)
            seg.div_(g.dp_size)
            g.copy_back.append(i)

    def _finish_reduce(self, g: _GroupState):
        """Join the reductions of this step into the compute stream; ``g.gshard`` (or the arena sub-slices on the
        peer-memory path) is final afterwards."""
        for h in g.handles:
            h.wait()
        g.handles = []
        if self._fused is not None and g.gid in self._fused.groups:
            self._fused.join(self)
            return
        if g.zero_size == 1:
            g.copy_back = []        # grad_shard() IS the arena
        elif g.dp_size <= 1:        # pragma: no cover - zero_size > 1 implies data parallelism
            g.copy_back = list(range(len(g.ranges)))
        if g.copy_back:
            gs = g.grad_shard()
            for i in g.copy_back:
                a, m, n = g.sub(i)
                gs[m: m + n].copy_(g.grad_arena[a: a + n])
            g.copy_back = []

    # ------------------------------------------------------------------------------------------------------------
    def _collect_grads(self, g: _GroupState):
        """Fold autograd-produced ``.grad`` tensors into the arena; zero slots of parameters that got no gradient."""
        for p in g.params:
            if p.grad is not None:
                self._fold_autograd_grad(p)
            elif not p.grad_ready:
                p.grad_buf.zero_()

    def _reduce_replica_grads(self, g: _GroupState):
        """Sequence parallel: norm-weight gradients are partial sums over each rank's sequence shard."""
        if not is_using_sequence_parallel() and not self.use_isp:
            return
        mode = ParallelMode.WEIGHT if self.use_isp else ParallelMode.TENSOR
        if g.dp_mode is ParallelMode.EXPERT_DATA and g.params and getattr(g.params[0], "expert_tp_sharded", False):
            return   # tensor-sharded experts: the parallel linears already produce every shard's complete gradient
        if g.dp_mode is ParallelMode.EXPERT_DATA:
            # experts are replicated over the tensor (sequence) group - they are neither tensor- nor weight-sharded here; under
            # sequence parallelism each of those ranks routed a different sequence shard through them, so their gradients are
            # partial: summed for msp / fsp (the loss is normalised over the full sequence), averaged for isp (every rank
            # normalises over its own shard)
            if _group_size(ParallelMode.TENSOR) > 1:
                if self.use_isp:
                    _all_reduce_avg(g.grad_arena, ParallelMode.TENSOR)
                else:
                    dist.all_reduce(g.grad_arena, group=gpc.get_group(ParallelMode.TENSOR))
            return
        if _group_size(mode) <= 1:
            return
        if g.replica_start >= g.replica_end:
            return
        # the slice ends at the last replica parameter: the tail padding and the START offset may differ between tensor ranks
        # (a row-parallel bias lives on tensor rank 0 only), the replica region itself has the same size everywhere
        rep = g.grad_arena[g.replica_start: g.replica_end]
        if self.use_isp:
            _all_reduce_avg(rep, mode)
        else:
            dist.all_reduce(rep, group=gpc.get_group(mode))

    def _group_sumsq(self, g: _GroupState) -> torch.Tensor:
        """Σ grad² of the owned sub-slices with replica parameters counted only on tp/wp rank 0, reduced over the ZeRO and
        tensor (weight) groups (reference ``compute_norm``, ``optimizer/utils.py:265-378``).  The sum over pipeline stages
        is done for all groups at once by ``_reduce_sumsq_over_pipeline``."""
        model_mode = ParallelMode.WEIGHT if self.use_isp else ParallelMode.TENSOR
        count_replica = gpc.get_local_rank(model_mode) == 0
        if self._fused is not None and g.gid in self._fused.groups:
            self._fused.local_sumsq(self, g, count_replica)     # the reduce kernels already accumulated the squares
        else:
            g.sumsq.zero_()
            owned = g.grad_shard()
            if g.replica_moff > 0:
                ops.sumsq_(owned[: g.replica_moff], g.sumsq)
            if count_replica and g.replica_moff < g.shard:
                ops.sumsq_(owned[g.replica_moff:], g.sumsq)
        if g.zero_size > 1:
            dist.all_reduce(g.sumsq, group=gpc.get_group(g.zero_mode))
        if g.dp_mode is ParallelMode.EXPERT_DATA:
            if _group_size(ParallelMode.EXPERT) > 1:
                dist.all_reduce(g.sumsq, group=gpc.get_group(ParallelMode.EXPERT))
        # replicated experts: the same gradient on every tensor / weight rank (tensor-sharded experts add their shards)
        replicated_experts = g.dp_mode is ParallelMode.EXPERT_DATA and not (
            g.params and getattr(g.params[0], "expert_tp_sharded", False))
        if _group_size(model_mode) > 1 and not (self.use_isp and g.dp_mode is ParallelMode.DATA) and not replicated_experts:
            dist.all_reduce(g.sumsq, group=gpc.get_group(model_mode))
        return g.sumsq

    def _reduce_sumsq_over_pipeline(self, active) -> None:
        """Add the per-group Σ grad² of all pipeline stages.  Stages own different parameter groups (under ISP the embedding
        group exists on the first stage only, MoE / fp32 groups only where such layers live), so the reduction runs over the
        UNION of group names - agreed once over the pipeline group - as ONE vector all-reduce: every stage issues the same
        collective whatever it owns (a per-group all-reduce would dead-lock as soon as two stages disagree)."""
        if _group_size(ParallelMode.PIPELINE) <= 1:
            return
        group = gpc.get_group(ParallelMode.PIPELINE)
        if self._pp_group_names is None:
            gathered = [None] * _group_size(ParallelMode.PIPELINE)
            dist.all_gather_object(gathered, [g.name for g in self.groups if g.params], group=group)
            self._pp_group_names = sorted({n for names in gathered for n in names})
        by_name = {g.name: g for g in active}
        dev = active[0].sumsq.device if active else self.device
        vec = torch.zeros(len(self._pp_group_names), dtype=torch.float32, device=dev)
        for i, n in enumerate(self._pp_group_names):
            if n in by_name:
                vec[i:i + 1] = by_name[n].sumsq
        dist.all_reduce(vec, group=group)
        # an overflow in ANY group of ANY stage must skip the step on every stage (a group may exist on one stage only):
        # poison all entries when one is not finite - a device-side select, no host sync
        vec = torch.where(torch.isfinite(vec).all(), vec, torch.full_like(vec, float("inf")))
        for i, n in enumerate(self._pp_group_names):
            if n in by_name:
                by_name[n].sumsq.copy_(vec[i:i + 1])

    # ---- parameter update + redistribution, range by range --------------------------------------------------------------
    def _hyper(self, g: _GroupState):
        cfg = g.cfg
        beta1, beta2 = cfg.get("betas", (0.9, 0.95))
        return cfg["lr"], beta1, beta2, cfg.get("eps", 1e-8), cfg.get("weight_decay", 0.0)

    def _update_range(self, g: _GroupState, i: int):
        """AdamW on this rank's sub-slice of range ``i`` and redistribution of the new low-precision values inside the ZeRO
        group.  Runs on the current stream (the side stream when overlapped)."""
        lr, beta1, beta2, eps, wd = self._hyper(g)
        lo, hi = g.ranges[i]
        a, m, n = g.sub(i)
        if self._fused is not None and g.gid in self._fused.groups:
            self._fused.update_range(self, g, i, (lr, beta1, beta2, eps, wd))
            return
        lowp = g.dtype is not torch.float32
        lp = g.param_arena[a: a + n] if lowp else None
        ops.adamw_(g.master[m: m + n], g.exp_avg[m: m + n], g.exp_avg_sq[m: m + n], g.grad_shard()[m: m + n], lp, lr,
                   beta1, beta2, eps, wd, g.step, g.scalars)
        if lp is None:
            g.param_arena[a: a + n].copy_(g.master[m: m + n])
        if g.zero_size > 1:
            group = gpc.get_group(g.zero_mode)
            seg, mine = g.param_arena[lo:hi], g.param_arena[a: a + n]
            try:
                dist.all_gather_into_tensor(seg, mine, group=group)
            except RuntimeError:  # very old gloo builds
                parts = list(seg.chunk(g.zero_size))
                dist.all_gather(parts, mine.clone(), group=group)

    # The update is 28 bytes per parameter of pure HBM streaming during which the tensor cores idle, and nothing but the NEXT
    # forward depends on it - block by block.  It therefore runs range by range (norm weights first, then arena order =
    # forward order) on a side stream; the pre-forward hook of every block makes the compute stream wait for the ranges that
    # hold that block's parameters only.  The backward of the next step cannot start before its forward, so the gradients a
    # range reads are never overwritten early.  Same arithmetic, same order of operations per element.
    def bind_model(self, model) -> None:
        """Remember the model: parameter names go into the checkpoint's plan, and the checkpoint converters translate optimizer
        state through the model's state-dict hooks (``checkpoint/optimizer_interchange.py``)."""
        # the object the checkpoint manager saves: the AMP wrapper is looked through, a list of pipeline chunks is kept
        inner = model.model if hasattr(model, "model") and not isinstance(model, torch.nn.ModuleList) else model
        self._model = inner
        self._param_names = {id(p): n for n, p in inner.named_parameters()}

    def param_name(self, p) -> str:
        name = self._param_names.get(id(p))
        assert name is not None, "HybridZeroOptimizer.bind_model(model) has not been called: parameter names are unknown"
        return name

    def attach_model(self, model) -> None:
        """Register the pre-forward hooks (called once by ``initialize_optimizer``).  A transformer block waits as a whole -
        its forward may read ``self.w13.weight`` or a gate weight without calling the sub-module that owns it - and every
        parameter outside the blocks (embedding, final norm, head) waits on the module that owns it directly."""
        if not self._adam_overlap or self._model_attached:
            return
        modules = list(model) if isinstance(model, (list, torch.nn.ModuleList)) else [model]
        waiters: Dict[int, torch.nn.Module] = {}      # id(param) -> module whose pre-forward hook guards it
        for m in modules:
            for sub in m.modules():
                if getattr(sub, "is_zero_wait_block", False) or type(sub).__name__ == "DecoderLayer":
                    for p in sub.parameters():
                        waiters.setdefault(id(p), sub)
            for sub in m.modules():
                for p in sub.parameters(recurse=False):
                    waiters.setdefault(id(p), sub)
        for g in self.groups:
            if not g.params:
                continue
            self._ensure_events(g)
            for p in g.params:
                sub = waiters.get(id(p))
                if sub is None:      # a parameter no module owns: the overlap cannot be made safe
                    self._adam_overlap = False
                    return
                ev = g.events[g.range_of[id(p)]]
                lst = self._owner_events.setdefault(id(sub), [])
                if ev not in lst:
                    lst.append(ev)
        for m in modules:
            for sub in m.modules():
                if id(sub) in self._owner_events:
                    sub.register_forward_pre_hook(self._pre_forward_wait)
        self._model_attached = True

    def _ensure_events(self, g: _GroupState):
        if g.events is None:
            g.events = [torch.cuda.Event() for _ in g.ranges]
            g.ready_ev = torch.cuda.Event()

    def _pre_forward_wait(self, module, inputs):
        if self._owner_seen.get(id(module), 0) != self._update_gen:
            self._owner_seen[id(module)] = self._update_gen
            stream = torch.cuda.current_stream()
            for ev in self._owner_events[id(module)]:
                stream.wait_event(ev)

    def _can_overlap_update(self, g: _GroupState) -> bool:
        return self._adam_overlap and self._model_attached and g.master.is_cuda and g.events is not None

    def _update_order(self, g: _GroupState):
        """Replica tail (norm weights: needed by the very first block) first, then arena order = forward order."""
        return list(range(g.n_sharded_ranges, len(g.ranges))) + list(range(g.n_sharded_ranges))

    def _update_group(self, g: _GroupState) -> bool:
        g.step += 1
        if not self._can_overlap_update(g):
            for i in self._update_order(g):
                self._update_range(g, i)
            if self._fused is not None and g.gid in self._fused.groups:
                self._fused.after_update(self, g)
            return False
        side, main = self._side(), torch.cuda.current_stream()
        g.ready_ev.record(main)                 # gradients, clip multiplier and overflow flag are final
        side.wait_event(g.ready_ev)
        with torch.cuda.stream(side):
            for i in self._update_order(g):
                self._update_range(g, i)
                g.events[i].record(side)
        return True

    def flush_param_update(self) -> None:
        """Make the current stream wait for every in-flight update range (checkpointing, state loading, end of a timed
        region): after this call parameters and optimizer state can be read or written in stream order as usual."""
        if self._side_stream is not None:
            torch.cuda.current_stream().wait_stream(self._side_stream)

    # ------------------------------------------------------------------------------------------------------------
    def step(self, closure=None):
        """→ ``(success, {group_name: grad_norm})``; a non-finite norm skips the update and backs off the loss scale."""
        assert closure is None
        timer("sync_grad").start()
        active = [g for g in self.groups if g.params]
        for g in active:
            self._collect_grads(g)
            self._reduce_replica_grads(g)
        for g in active:
            self._advance(g, final=True)       # whatever the hooks did not launch during backward, in the same order
        for g in active:
            self._finish_reduce(g)
        timer("sync_grad").stop()
        timer("step").start()
        scale = self.grad_scaler.scale
        for g in active:
            self._group_sumsq(g)
        self._reduce_sumsq_over_pipeline(active)
        for g in active:
            ops.clip_scalars_(g.sumsq, g.scalars, scale, self._clip_grad_norm)
        # overflow anywhere must skip every group: fold the flags (tiny device op), still no host sync
        if len(active) > 1:
            flag = torch.stack([g.scalars[1] for g in active]).max()
            for g in active:
                g.scalars[1] = flag
        overlapped = False
        for g in active:
            overlapped |= self._update_group(g)
        if overlapped:
            self._update_gen += 1
        timer("step").stop()
        # single read-back for logging / loss-scale bookkeeping (everything above is already queued)
        host = torch.stack([g.scalars for g in active]).cpu() if self.has_params else torch.zeros(1, 4)
        found_inf = bool((host[:, 1] != 0).any())
        norms = {g.name: float(host[i, 2]) for i, g in enumerate(active)}
        self.grad_scaler.update(found_inf)
        if found_inf:
            for g in active:
                g.step -= 1
            if gpc.is_rank_for_log():
                logger.warning("Overflow occurs, please check it.")
            self.zero_grad()
            return False, {k: -1.0 for k in norms}
        self.zero_grad()
        return True, norms

    # ------------------------------------------------------------------------------------------------------------
    def clip_grad_norm(self, model, max_norm):
        """No-op: clipping happens inside ``step`` (reference ``hybrid_zero_optim.py:855-857``)."""

    def _redistribute_params(self, g: _GroupState):
        """Full (blocking) parameter all-gather of every range: state loading / master reload."""
        if g.zero_size <= 1:
            return
        group = gpc.get_group(g.zero_mode)
        for i, (lo, hi) in enumerate(g.ranges):
            a, m, n = g.sub(i)
            seg, mine = g.param_arena[lo:hi], g.param_arena[a: a + n]
            try:
                dist.all_gather_into_tensor(seg, mine, group=group)
            except RuntimeError:  # very old gloo builds
                parts = list(seg.chunk(g.zero_size))
                dist.all_gather(parts, mine.clone(), group=group)

    # ---- the reference's introspection surface (``hybrid_zero_optim.py:238-244,369-388,809-837``) on the arena layout ----
    @property
    def zero_local_rank(self) -> List[int]:
        """This rank's position in the ZeRO group of every parameter group."""
        return [g.zero_rank for g in self.groups]

    @property
    def zero_world_size(self) -> List[int]:
        return [g.zero_size for g in self.groups]

    def _state_of(self, param) -> Optional[_GroupState]:
        for g in self.groups:
            if id(param) in g.offsets:
                return g
        return None

    def belongs_to_current_rank(self, param) -> bool:
        """Does this rank hold optimizer state for (part of) ``param``?  Ranges are split evenly over the ZeRO group, so a
        parameter may span several ranks: true for each of them."""
        g = self._state_of(param)
        if g is None:
            return False
        lo, hi = g.ranges[g.range_of[id(param)]]
        n = (hi - lo) // g.zero_size
        start = g.offsets[id(param)] - lo
        return start // n <= g.zero_rank <= (start + param.numel() - 1) // n

    def broadcast_params(self) -> None:
        """Bring every rank's low-precision parameters up to date with the owners' fp32 masters (blocking).  ``step`` does this
        range by range, overlapped; this is the explicit form (after editing the masters by hand, in tests)."""
        self.flush_param_update()
        for g in self.groups:
            if g.params:
                g.push_master()
                self._redistribute_params(g)

    def accumulate_left_grads_after_backward(self) -> None:
        """ISP with overlap: fold the weight-gradient reduce-scatters still in flight into the gradient arena."""
        if self._isp_communicator is not None and getattr(self._isp_communicator, "overlap", False):
            self._isp_communicator.flush_grads()

    def state_dict(self):
        self.flush_param_update()
        states = {"grad_scaler": self.grad_scaler.state_dict(), "zero_devide_optim_plan": {}, "groups": []}
        for g in self.groups:
            states["groups"].append({
                "name": g.name, "step": g.step, "total": g.total, "layout": "range-interleaved",
                "ranges": [list(r) for r in g.ranges], "zero_rank": g.zero_rank, "zero_size": g.zero_size,
                "flat_fp32_weights": g.master.detach().cpu(), "exp_avg": g.exp_avg.detach().cpu(),
                "exp_avg_sq": g.exp_avg_sq.detach().cpu(),
                "hyper": {k: v for k, v in g.cfg.items() if k != "params"},
            })
            states["zero_devide_optim_plan"][g.name] = {
                "zero_rank": g.zero_rank, "zero_size": g.zero_size, "ranges": [list(r) for r in g.ranges],
                "offsets": [(list(p.shape), g.offsets[id(p)]) for p in g.ordered],
                # parameter names make the file layout-free: it can be re-sharded / converted (checkpoint/optimizer_interchange.py)
                "names": [self._param_names.get(id(p)) for p in g.ordered],
            }
        return states

    def load_state_dict(self, states):
        assert "grad_scaler" in states, "Not found grad_scaler state!"
        self.flush_param_update()
        self.grad_scaler.load_state_dict(states["grad_scaler"])
        if gpc.config is not None and gpc.config.get("only_load_lr", False):
            # resume of the schedule only: learning rates (and the loss scale above), no weights / moments
            for g, st in zip(self.groups, states["groups"]):
                if "lr" in st.get("hyper", {}):
                    g.cfg["lr"] = st["hyper"]["lr"]
            return
        for g, st in zip(self.groups, states["groups"]):
            same = (st.get("layout") == "range-interleaved" and st["total"] == g.total
                    and [tuple(r) for r in st["ranges"]] == g.ranges and st["zero_rank"] == g.zero_rank
                    and st["zero_size"] == g.zero_size)
            assert same, (f"optimizer checkpoint layout mismatch for group {g.name}: parallel sizes and "
                          f"reduce_bucket_size must match the checkpoint")
            g.step = st["step"]
            g.master.copy_(st["flat_fp32_weights"])
            g.exp_avg.copy_(st["exp_avg"])
            g.exp_avg_sq.copy_(st["exp_avg_sq"])
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in st.get("hyper", {}):
                    g.cfg[k] = st["hyper"][k]
            g.push_master()
            self._redistribute_params(g)

    def reload_zero_fp32_buff(self):
        """After a model-only load: refresh the fp32 master from the (new) low-precision parameters."""
        self.flush_param_update()
        for g in self.groups:
            g.pull_master()


def reload_zero_fp32_buff(optimizer):
    """Module-level form of :meth:`HybridZeroOptimizer.reload_zero_fp32_buff` (reference ``hybrid_zero_optim.py:939-950``):
    refresh the fp32 master shards after weights were loaded into the model; a no-op for other optimizers."""
    if isinstance(optimizer, HybridZeroOptimizer):
        optimizer.reload_zero_fp32_buff()
