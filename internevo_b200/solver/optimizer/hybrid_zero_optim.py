"""Hybrid-ZeRO (ZeRO-1.5) optimizer on persistent arenas.

Semantics preserved from the reference (``internlm/solver/optimizer/hybrid_zero_optim.py:55-950``): parameter groups with
their own reduction group (DATA / WEIGHT_DATA / EXPERT_DATA), optimizer state + fp32 master sharded over the ZeRO
sub-group while gradients are averaged over the whole data-parallel group, dynamic loss scale with skip-on-overflow,
per-group grad-norm + clipping with replica parameters counted once, updated parameters redistributed inside the ZeRO
group, resumable ``state_dict``.

Redesign (what changes on a B200 node):

* every group owns ONE contiguous low-precision parameter arena and ONE gradient arena for the life of the job; model
  parameters are views into the first and the wgrad GEMM epilogues / norm-backward kernels accumulate straight into
  the second (``param.grad_buf``).  The reference's per-step flatten → all-reduce → unflatten → copy chain
  (``store.py:315-322``, ``hybrid_zero_optim.py:455-523,740-797``) disappears;
* the arena is sharded by *element* (not by whole parameter), so gradient sync is a reduce-scatter and parameter sync an
  all-gather when the ZeRO group is the DP group; with a smaller ZeRO group the gradient is all-reduced over DP (as the
  reference does) and only the owned slice is consumed;
* unscale + clip + AdamW + bf16 cast-back are ONE kernel over the owned shard whose multiplier / skip flag live on the
  device, so the step issues no host sync until the norm is read back for logging;
* with a peer-memory heap (``parallel/symm.py``) the reduce-scatter, the update and the parameter all-gather are fused in
  one NVLink kernel (``reduce_scatter_adam``) — NCCL is the fallback and the oracle.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

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
    """Arena bookkeeping of one parameter group."""

    def __init__(self, gid: int, cfg: dict, params: List[torch.nn.Parameter], dp_mode: ParallelMode,
                 zero_mode: ParallelMode, device):
        self.gid, self.cfg, self.params = gid, cfg, params
        self.name = cfg.get("name", f"group{gid}")
        self.dp_mode, self.zero_mode = dp_mode, zero_mode
        self.zero_size = _group_size(zero_mode)
        self.zero_rank = gpc.get_local_rank(zero_mode) if gpc.is_initialized(zero_mode) else 0
        self.dp_size = _group_size(dp_mode)
        self.dtype = params[0].dtype if params else torch.float32
        # sharded params first, replica params (norm weights / gates: identical on every TP rank) last
        rep = [p for p in params if getattr(p, IS_REPLICA_ZERO_PARALLEL, False)]
        shd = [p for p in params if not getattr(p, IS_REPLICA_ZERO_PARALLEL, False)]
        self.ordered = shd + rep
        off = 0
        self.offsets: Dict[int, int] = {}
        for p in shd:
            self.offsets[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.replica_start = off
        for p in rep:
            self.offsets[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.replica_end = off     # [replica_start, replica_end): same SIZE on every tensor rank (offsets may differ)
        quantum = self.zero_size * 1024
        self.total = max(quantum, (off + quantum - 1) // quantum * quantum)
        self.shard = self.total // self.zero_size
        self.lo, self.hi = self.zero_rank * self.shard, (self.zero_rank + 1) * self.shard
        self.param_arena = torch.zeros(self.total, dtype=self.dtype, device=device)
        self.grad_arena = torch.zeros(self.total, dtype=self.dtype, device=device)
        for p in self.ordered:
            o = self.offsets[id(p)]
            view = self.param_arena[o: o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad_buf = self.grad_arena[o: o + p.numel()].view(p.shape)
            p.grad_ready = False
        # fp32 master + moments of the owned slice
        self.master = self.param_arena[self.lo: self.hi].float()
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.step = 0
        self.scalars = torch.zeros(4, dtype=torch.float32, device=device)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self.param_sync_handle = None
        self.plan = None      # chunk plan of the overlapped update (HybridZeroOptimizer._chunk_plan)

    def owned_grad(self) -> torch.Tensor:
        return self.grad_arena[self.lo: self.hi]


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
        self._overlap_sync_grad = zero_cfg.get("overlap_sync_grad", False)
        self._overlap_sync_param = zero_cfg.get("overlap_sync_param", False)
        self._reduce_bucket_size = zero_cfg.get("reduce_bucket_size", 512 * 1024 * 1024)
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
            self.groups.append(_GroupState(gid, pg, params, dp_mode, zero_mode, self.device))
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
        # AdamW of step s overlapped with the forward of step s + 1 (unsharded groups, see _update_overlapped)
        self._adam_overlap = os.environ.get("B200_ADAM_OVERLAP", "1") != "0" and torch.cuda.is_available() \
            and not self.use_isp
        self._opt_stream = None
        self._owner_events: Dict[int, list] = {}     # id(module) -> events of the chunks holding its parameters
        self._owner_seen: Dict[int, int] = {}        # id(module) -> update generation it has already waited for
        self._update_gen = 0
        self._model_attached = False
        self._pp_group_names = None   # union of parameter-group names over the pipeline group (agreed at the first step)

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

    def backward(self, loss, retain_graph=False):
        (loss * self.grad_scaler.scale).backward(retain_graph=retain_graph)

    def backward_by_grad(self, tensor, grad):
        torch.autograd.backward(tensors=tensor, grad_tensors=grad)

    def wait_param_sync(self):
        for g in self.groups:
            if g.param_sync_handle is not None:
                g.param_sync_handle.wait()
                g.param_sync_handle = None

    # ------------------------------------------------------------------------------------------------------------
    def _collect_grads(self, g: _GroupState):
        """Fold autograd-produced ``.grad`` tensors into the arena; zero slots of parameters that got no gradient."""
        for p in g.params:
            if p.grad is not None:
                if p.grad_ready:
                    p.grad_buf.add_(p.grad)
                else:
                    p.grad_buf.copy_(p.grad)
                p.grad = None
                p.grad_ready = True
            elif not p.grad_ready:
                p.grad_buf.zero_()

    def _reduce_replica_grads(self, g: _GroupState):
        """Sequence parallel: norm-weight gradients are partial sums over each rank's sequence shard."""
        if not is_using_sequence_parallel() and not self.use_isp:
            return
        mode = ParallelMode.WEIGHT if self.use_isp else ParallelMode.TENSOR
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

    def _sync_grads(self, g: _GroupState):
        """Average over the data-parallel group; afterwards ``g.owned_grad()`` holds this rank's reduced slice."""
        dp_group, zero_group = gpc.get_group(g.dp_mode), gpc.get_group(g.zero_mode)
        if g.dp_size <= 1 or dp_group is None:
            return
        same = g.zero_size == g.dp_size and gpc.get_ranks_in_group(g.dp_mode) == gpc.get_ranks_in_group(g.zero_mode)
        if same and g.grad_arena.is_cuda:
            dist.reduce_scatter_tensor(g.owned_grad(), g.grad_arena, op=dist.ReduceOp.AVG, group=dp_group)
        else:
            _all_reduce_avg(g.grad_arena, g.dp_mode)
        del zero_group

    def _group_sumsq(self, g: _GroupState) -> torch.Tensor:
        """Σ grad² of the owned slice with replica parameters counted only on tp/wp rank 0, reduced over the ZeRO and
        tensor (weight) groups (reference ``compute_norm``, ``optimizer/utils.py:265-378``).  The sum over pipeline stages
        is done for all groups at once by ``_reduce_sumsq_over_pipeline``."""
        g.sumsq.zero_()
        owned = g.owned_grad()
        rep_lo = max(g.replica_start, g.lo) - g.lo
        model_mode = ParallelMode.WEIGHT if self.use_isp else ParallelMode.TENSOR
        count_replica = gpc.get_local_rank(model_mode) == 0
        if rep_lo >= g.shard:
            ops.sumsq_(owned, g.sumsq)
        else:
            if rep_lo > 0:
                ops.sumsq_(owned[:rep_lo], g.sumsq)
            if count_replica:
                ops.sumsq_(owned[rep_lo:], g.sumsq)
        if g.zero_size > 1:
            dist.all_reduce(g.sumsq, group=gpc.get_group(g.zero_mode))
        if g.dp_mode is ParallelMode.EXPERT_DATA:
            if _group_size(ParallelMode.EXPERT) > 1:
                dist.all_reduce(g.sumsq, group=gpc.get_group(ParallelMode.EXPERT))
        replicated_experts = g.dp_mode is ParallelMode.EXPERT_DATA   # same gradient on every tensor / weight rank
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

    def _update(self, g: _GroupState):
        cfg = g.cfg
        beta1, beta2 = cfg.get("betas", (0.9, 0.95))
        g.step += 1
        lp = g.param_arena[g.lo: g.hi] if g.dtype is not torch.float32 else None
        ops.adamw_(g.master, g.exp_avg, g.exp_avg_sq, g.owned_grad(), lp, cfg["lr"], beta1, beta2, cfg.get("eps", 1e-8),
                   cfg.get("weight_decay", 0.0), g.step, g.scalars)
        if lp is None:
            g.param_arena[g.lo: g.hi].copy_(g.master)

    # ---- AdamW overlapped with the next forward ------------------------------------------------------------------------
    # Without ZeRO sharding (zero group of size 1, e.g. the 1-GPU benchmark) the update is 28 bytes per parameter of pure
    # HBM streaming (40 ms for 7.7 B parameters) during which the tensor cores idle, and nothing but the NEXT forward
    # depends on it - layer by layer.  The update therefore runs chunk by chunk (norm weights first, then arena order =
    # forward order) on a side stream; every module's pre-forward hook makes the compute stream wait for the chunks that
    # hold that module's parameters only.  The backward of the next step cannot start before its forward, so the
    # gradients an update chunk reads are never overwritten early.  Same arithmetic, same order of operations per element.
    def attach_model(self, model) -> None:
        """Register the pre-forward hooks (called once by ``initialize_optimizer``)."""
        if not self._adam_overlap or self._model_attached:
            return
        owner_of = {}
        modules = model if isinstance(model, (list, torch.nn.ModuleList)) else [model]
        for m in modules:
            for sub in m.modules():
                for p in sub.parameters(recurse=False):
                    owner_of.setdefault(id(p), sub)
        for g in self.groups:
            if not g.params or g.zero_size != 1:
                continue
            plan = self._chunk_plan(g)
            for p in g.params:
                sub = owner_of.get(id(p))
                if sub is None:      # a parameter no module owns directly: the overlap cannot be made safe
                    self._adam_overlap = False
                    return
                ev = plan["events"][self._chunk_of(plan, g.offsets[id(p)])]
                lst = self._owner_events.setdefault(id(sub), [])
                if ev not in lst:
                    lst.append(ev)
        for m in modules:
            for sub in m.modules():
                if id(sub) in self._owner_events:
                    sub.register_forward_pre_hook(self._pre_forward_wait)
        self._model_attached = True

    def _pre_forward_wait(self, module, inputs):
        if self._owner_seen.get(id(module), 0) != self._update_gen:
            self._owner_seen[id(module)] = self._update_gen
            stream = torch.cuda.current_stream()
            for ev in self._owner_events[id(module)]:
                stream.wait_event(ev)

    def _chunk_plan(self, g: _GroupState):
        if g.plan is None:
            ranges = []
            if g.replica_start < g.total:
                ranges.append((g.replica_start, g.total))          # norm weights / gates: needed by the very first layer
            starts = sorted(g.offsets[id(p)] for p in g.ordered if g.offsets[id(p)] < g.replica_start)
            target = max(1 << 22, g.replica_start // 48)
            lo = 0
            for s in starts[1:]:
                if s - lo >= target:
                    ranges.append((lo, s))
                    lo = s
            if g.replica_start > lo:
                ranges.append((lo, g.replica_start))
            g.plan = {"ranges": ranges, "starts": [r[0] for r in ranges],
                      "events": [torch.cuda.Event() for _ in ranges], "ready": torch.cuda.Event()}
        return g.plan

    @staticmethod
    def _chunk_of(plan, offset: int) -> int:
        best = 0
        for i, (lo, hi) in enumerate(plan["ranges"]):
            if lo <= offset < hi:
                best = i
        return best

    def _can_overlap(self, g: _GroupState) -> bool:
        return (self._adam_overlap and self._model_attached and g.zero_size == 1 and g.master.is_cuda
                and g.plan is not None)

    def _update_overlapped(self, g: _GroupState):
        cfg = g.cfg
        beta1, beta2 = cfg.get("betas", (0.9, 0.95))
        g.step += 1
        plan = g.plan
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=g.master.device)
        main = torch.cuda.current_stream()
        plan["ready"].record(main)                 # gradients, clip multiplier and overflow flag are final
        self._opt_stream.wait_event(plan["ready"])
        lowp = g.dtype is not torch.float32
        with torch.cuda.stream(self._opt_stream):
            for (lo, hi), ev in zip(plan["ranges"], plan["events"]):
                lp = g.param_arena[lo:hi] if lowp else None
                ops.adamw_(g.master[lo:hi], g.exp_avg[lo:hi], g.exp_avg_sq[lo:hi], g.grad_arena[lo:hi], lp, cfg["lr"], beta1,
                           beta2, cfg.get("eps", 1e-8), cfg.get("weight_decay", 0.0), g.step, g.scalars)
                if lp is None:
                    g.param_arena[lo:hi].copy_(g.master[lo:hi])
                ev.record(self._opt_stream)

    def flush_param_update(self) -> None:
        """Make the current stream wait for every in-flight update chunk (checkpointing, state loading, end of a timed
        region): after this call parameters and optimizer state can be read or written in stream order as usual."""
        if self._opt_stream is not None:
            torch.cuda.current_stream().wait_stream(self._opt_stream)

    def _sync_params(self, g: _GroupState):
        group = gpc.get_group(g.zero_mode)
        if g.zero_size <= 1 or group is None:
            return
        shard = g.param_arena[g.lo: g.hi]
        try:
            h = dist.all_gather_into_tensor(g.param_arena, shard, group=group, async_op=self._overlap_sync_param)
        except RuntimeError:  # very old gloo builds
            parts = list(g.param_arena.chunk(g.zero_size))
            h = dist.all_gather(parts, shard.clone(), group=group, async_op=self._overlap_sync_param)
        g.param_sync_handle = h if self._overlap_sync_param else None

    # ------------------------------------------------------------------------------------------------------------
    def step(self, closure=None):
        """→ ``(success, {group_name: grad_norm})``; a non-finite norm skips the update and backs off the loss scale."""
        assert closure is None
        self.wait_param_sync()
        timer("sync_grad").start()
        if self._fused is not None:
            ok, norms = self._fused.step(self)
            timer("sync_grad").stop()
            return ok, norms
        for g in self.groups:
            if not g.params:
                continue
            self._collect_grads(g)
            self._reduce_replica_grads(g)
            self._sync_grads(g)
        timer("sync_grad").stop()
        timer("step").start()
        scale = self.grad_scaler.scale
        active = [g for g in self.groups if g.params]
        for g in active:
            self._group_sumsq(g)
        self._reduce_sumsq_over_pipeline(active)
        for g in active:
            ops.clip_scalars_(g.sumsq, g.scalars, scale, self._clip_grad_norm)
        # overflow anywhere must skip every group: fold the flags (tiny device op), still no host sync
        if len(self.groups) > 1:
            flag = torch.stack([g.scalars[1] for g in self.groups if g.params]).max()
            for g in self.groups:
                if g.params:
                    g.scalars[1] = flag
        overlapped = False
        for g in self.groups:
            if g.params:
                if self._can_overlap(g):
                    self._update_overlapped(g)
                    overlapped = True
                else:
                    self._update(g)
                self._sync_params(g)
        if overlapped:
            self._update_gen += 1
        timer("step").stop()
        # single read-back for logging / loss-scale bookkeeping (everything above is already queued)
        host = torch.stack([g.scalars for g in self.groups if g.params]).cpu() if self.has_params else torch.zeros(1, 4)
        found_inf = bool((host[:, 1] != 0).any())
        norms = {}
        i = 0
        for g in self.groups:
            if g.params:
                norms[g.name] = float(host[i, 2])
                i += 1
        self.grad_scaler.update(found_inf)
        if found_inf:
            for g in self.groups:
                if g.params:
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

    def state_dict(self):
        self.flush_param_update()
        states = {"grad_scaler": self.grad_scaler.state_dict(), "zero_devide_optim_plan": {}, "groups": []}
        for g in self.groups:
            states["groups"].append({
                "name": g.name, "step": g.step, "lo": g.lo, "hi": g.hi, "total": g.total,
                "flat_fp32_weights": g.master.detach().cpu(), "exp_avg": g.exp_avg.detach().cpu(),
                "exp_avg_sq": g.exp_avg_sq.detach().cpu(),
                "hyper": {k: v for k, v in g.cfg.items() if k != "params"},
            })
            states["zero_devide_optim_plan"][g.name] = {
                "zero_rank": g.zero_rank, "zero_size": g.zero_size,
                "offsets": [(list(p.shape), g.offsets[id(p)]) for p in g.ordered],
            }
        return states

    def load_state_dict(self, states):
        assert "grad_scaler" in states, "Not found grad_scaler state!"
        self.flush_param_update()
        self.grad_scaler.load_state_dict(states["grad_scaler"])
        for g, st in zip(self.groups, states["groups"]):
            assert st["total"] == g.total and st["lo"] == g.lo, (
                f"optimizer checkpoint layout mismatch for group {g.name}: the parallel sizes must match the checkpoint"
            )
            g.step = st["step"]
            g.master.copy_(st["flat_fp32_weights"])
            g.exp_avg.copy_(st["exp_avg"])
            g.exp_avg_sq.copy_(st["exp_avg_sq"])
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in st.get("hyper", {}):
                    g.cfg[k] = st["hyper"][k]
            g.param_arena[g.lo: g.hi].copy_(g.master)
            self._sync_params(g)
        self.wait_param_sync()

    def reload_zero_fp32_buff(self):
        """After a model-only load: refresh the fp32 master from the (new) low-precision parameters."""
        self.flush_param_update()
        for g in self.groups:
            g.master.copy_(g.param_arena[g.lo: g.hi])


def reload_zero_fp32_buff(optimizer):
    """Module-level form of :meth:`HybridZeroOptimizer.reload_zero_fp32_buff` (reference ``hybrid_zero_optim.py:939-950``):
    refresh the fp32 master shards after weights were loaded into the model; a no-op for other optimizers."""
    if isinstance(optimizer, HybridZeroOptimizer):
        optimizer.reload_zero_fp32_buff()
