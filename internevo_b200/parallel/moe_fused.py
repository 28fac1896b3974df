"""Expert-parallel dispatch / combine fused over NVLink peer memory (``csrc/moe_comm.cu``).

The NCCL formulation of an MoE layer (reference ``internlm/moe/sharded_moe.py:369-498`` and
``internlm/moe/megablock/megablock_moe.py:155-247``) is  sort/permute → ``all_to_all`` → regroup → experts → regroup →
``all_to_all`` → un-permute → weighted sum.  Here every routed (token, j) *slot* gets its final address ``(rank, row)`` in
the owner GPU's expert slab from the ``[world, E]`` matrix of per-expert counts, and two kernels do the rest:

* dispatch  = ``moe_scatter_rows``: token rows are stored straight into the owner's slab over NVLink,
* combine   = ``moe_gather_combine``: the k expert outputs of a token are pulled from their owners, weighted in fp32 and
  summed.

The backward passes are the same two kernels with the roles swapped (``d dispatch`` is a gather with unit weights,
``d combine`` is a scatter of ``w · dOut`` that also produces ``d w`` from the pulled expert outputs).

``slot_plan`` is pure tensor arithmetic (no communication, any device) so the addressing is unit-tested on CPU against
the all-to-all ordering (``tests/test_moe_cpu.py``).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from internevo_b200.ops.gemm import _bump
from internevo_b200.utils.logger import get_logger

from . import symm

logger = get_logger(__file__)


def slot_plan(expert_of_slot: torch.Tensor, count_matrix: torch.Tensor, rank: int, num_local_experts: int, align: int = 1):
    """Addresses of this rank's routed slots inside the owners' expert slabs.

    ``expert_of_slot``  int64 ``[n_slots]`` global expert of every (token, j) slot of THIS rank (slot = token * k + j)
    ``count_matrix``    int64 ``[world, E]``: rows routed by rank ``s`` to expert ``e``

    The slab of a GPU is ordered (local expert, source rank, arrival order) — exactly the order produced by a
    variable-split all-to-all followed by the regroup-by-expert permutation, so a local expert's rows are contiguous.
    With ``align > 1`` every local expert's slab STARTS at a multiple of ``align`` rows (the grouped GEMM's requirement);
    the rows in between are padding that nobody writes.

    Returns ``(slot_rank int32[n_slots], slot_row int32[n_slots], rows_per_local_expert int64[El], rows_per_rank
    int64[world])``; ``rows_per_rank`` includes the padding.
    """
    world, E = count_matrix.shape
    El = num_local_experts
    assert E == world * El
    cd = count_matrix.view(world, world, El)                      # [src, dst, el]
    lay = cd.permute(1, 2, 0).reshape(world, El * world)          # per dst: (el, src) order
    off = (lay.cumsum(1) - lay).view(world, El, world)            # exclusive prefix: slab offset of (dst, el, src)
    if align > 1:
        el_tot = lay.view(world, El, world).sum(2)                # [dst, el] rows of each local expert
        el_pad = (el_tot + align - 1) // align * align
        el_start = el_pad.cumsum(1) - el_pad                      # aligned start of every local expert's slab
        within = off - (el_tot.cumsum(1) - el_tot).unsqueeze(2)   # offset of (src) inside its expert's slab
        off = el_start.unsqueeze(2) + within
    my_base = off[:, :, rank].reshape(E)                          # [E]: where MY rows for expert e start on its owner
    # arrival order inside (me -> expert e): stable order of the slots routed to e
    order = torch.argsort(expert_of_slot, stable=True)
    e_sorted = expert_of_slot[order]
    my_counts = count_matrix[rank]
    start = my_counts.cumsum(0) - my_counts
    pos_sorted = torch.arange(e_sorted.numel(), device=e_sorted.device) - start[e_sorted]
    row_sorted = my_base[e_sorted] + pos_sorted
    slot_row = torch.empty_like(row_sorted)
    slot_row[order] = row_sorted
    slot_rank = torch.div(expert_of_slot, El, rounding_mode="floor")
    per_rank = lay.sum(1) if align <= 1 else el_pad.sum(1)
    return slot_rank.to(torch.int32), slot_row.to(torch.int32), cd[:, rank, :].sum(0), per_rank


class MoEFusedBackend:
    """Symmetric slabs + flags of one expert-parallel group.  ``max_rows`` bounds the rows a GPU can receive."""

    def __init__(self, group: dist.ProcessGroup, hidden: int, max_rows: int, num_experts: int):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.hidden, self.max_rows, self.num_experts = hidden, max_rows, num_experts
        self.xbuf = symm.SymmBuffer(max_rows * hidden, torch.bfloat16, group, zero=False)   # rows pushed to the owner
        self.ybuf = symm.SymmBuffer(max_rows * hidden, torch.bfloat16, group, zero=False)   # rows pulled by the source
        self.cnt = symm.SymmBuffer(self.world * num_experts, torch.int32, group, zero=True)
        self.flags = symm.flags_for(group)

    # ---- counts: [world, E] on every rank after ONE launch (push + rendezvous)
    def exchange_counts(self, counts: torch.Tensor) -> torch.Tensor:
        torch.ops.b200.symm_allgather_small(self.cnt.table_ptr(0), counts.to(torch.int32).contiguous(),
                                            self.flags.table_ptr(0), self.rank, self.world, self.flags.next_epoch())
        _bump()
        return self.cnt.tensor.view(self.world, self.num_experts).to(torch.int64)

    def x_rows(self, n: int) -> torch.Tensor:
        return self.xbuf.view(0, (n, self.hidden))

    def y_rows(self, n: int) -> torch.Tensor:
        return self.ybuf.view(0, (n, self.hidden))


_backends: Dict[Tuple[int, int, int, int], MoEFusedBackend] = {}


def backend_for(group, hidden: int, max_rows: int, num_experts: int) -> Optional[MoEFusedBackend]:
    """One backend per (group, hidden): the slabs are shared by all MoE layers of the model (they are used strictly one
    layer at a time, every use bracketed by device barriers)."""
    if dist.get_world_size(group) not in (2, 4, 8) or not symm.peer_addressable(group):
        return None
    key = (id(group), hidden, num_experts, 0)
    be = _backends.get(key)
    if be is None or be.max_rows < max_rows:
        try:
            be = MoEFusedBackend(group, hidden, max_rows, num_experts)
        except Exception as e:  # pragma: no cover - depends on driver / topology
            logger.warning(f"fused MoE dispatch unavailable ({e}); using NCCL all-to-all")
            return None
        _backends[key] = be
    return be


class _Plan:
    __slots__ = ("slot_rank", "slot_row", "n_recv", "k", "n_tokens", "zero_fill", "pad_idx")

    def __init__(self, slot_rank, slot_row, n_recv, k, n_tokens, zero_fill=False, pad_idx=None):
        self.slot_rank, self.slot_row, self.n_recv, self.k, self.n_tokens = slot_rank, slot_row, n_recv, k, n_tokens
        self.zero_fill = zero_fill   # capacity layout: rows nobody writes must read as zero
        self.pad_idx = pad_idx       # group-aligned layout: (fixed-size) indices of the padding rows, zeroed before every scatter


class _FusedDispatch(torch.autograd.Function):
    """x [T, H] → rows received by this GPU's experts [n_recv, H] (a private copy: the slab is reused by the next layer)."""

    @staticmethod
    def forward(ctx, x, be: MoEFusedBackend, plan: _Plan):
        ctx.be, ctx.plan = be, plan
        if plan.zero_fill:
            be.x_rows(plan.n_recv).zero_()
            be.flags.barrier()                  # every slab is cleared before the first row arrives
        elif symm.DEBUG:
            symm.poison(be.x_rows(plan.n_recv))
            be.flags.barrier()
        if plan.pad_idx is not None:            # nobody writes the alignment padding: it must read as zero (wgrad runs over it)
            be.x_rows(plan.n_recv).index_fill_(0, plan.pad_idx, 0)
        torch.ops.b200.moe_scatter_rows(x, plan.slot_rank, plan.slot_row, None, be.xbuf.table_ptr(0), 0, None, plan.k)
        _bump()
        be.flags.barrier()                      # every peer's rows have landed in my slab
        rows = be.x_rows(plan.n_recv).clone()
        if symm.DEBUG:
            symm.assert_clean(rows, "moe dispatch slab")
        be.flags.barrier()                      # nobody overwrites a slab that is still being read
        return rows

    @staticmethod
    def backward(ctx, g_rows):
        be, plan = ctx.be, ctx.plan
        be.y_rows(plan.n_recv).copy_(g_rows)
        be.flags.barrier()
        gx = torch.empty(plan.n_tokens, be.hidden, dtype=torch.bfloat16, device=g_rows.device)
        torch.ops.b200.moe_gather_combine(gx, None, plan.slot_rank, plan.slot_row, be.ybuf.table_ptr(0), plan.k)
        _bump()
        be.flags.barrier()
        return gx, None, None


class _FusedCombine(torch.autograd.Function):
    """expert outputs [n_recv, H] (on their owner) + gate weights [T * k] → combined [T, H] on the token's GPU."""

    @staticmethod
    def forward(ctx, out_rows, w, be: MoEFusedBackend, plan: _Plan):
        ctx.be, ctx.plan = be, plan
        ctx.save_for_backward(out_rows, w)
        be.y_rows(plan.n_recv).copy_(out_rows)
        be.flags.barrier()
        out = torch.empty(plan.n_tokens, be.hidden, dtype=torch.bfloat16, device=out_rows.device)
        if symm.DEBUG:
            symm.poison(out)
        torch.ops.b200.moe_gather_combine(out, w, plan.slot_rank, plan.slot_row, be.ybuf.table_ptr(0), plan.k)
        _bump()
        be.flags.barrier()
        if symm.DEBUG:
            symm.assert_clean(out, "moe combine output")
        return out

    @staticmethod
    def backward(ctx, g_out):
        be, plan = ctx.be, ctx.plan
        out_rows, w = ctx.saved_tensors
        g_out = g_out.contiguous()
        be.y_rows(plan.n_recv).copy_(out_rows)   # the owners publish their outputs again for d(gate weight)
        if plan.zero_fill:
            be.x_rows(plan.n_recv).zero_()       # capacity rows nobody routed to receive a zero gradient
        if plan.pad_idx is not None:
            be.x_rows(plan.n_recv).index_fill_(0, plan.pad_idx, 0)
        be.flags.barrier()
        dw = torch.empty_like(w)
        torch.ops.b200.moe_scatter_rows(g_out, plan.slot_rank, plan.slot_row, w, be.xbuf.table_ptr(0),
                                        be.ybuf.table_ptr(0), dw, plan.k)
        _bump()
        be.flags.barrier()
        g_rows = be.x_rows(plan.n_recv).clone()
        be.flags.barrier()
        return g_rows, dw, None, None


def fused_dispatch(x2: torch.Tensor, expert_of_slot: torch.Tensor, counts: torch.Tensor, be: MoEFusedBackend, k: int):
    """Returns ``(rows [R, H], offsets int32 [El + 1], plan)``: ``rows`` is this GPU's group-aligned expert slab (local
    expert ``e`` at ``offsets[e] .. offsets[e + 1]``, padding rows zero) - the layout the grouped GEMM consumes.  Everything
    is sized statically (``R`` = the slab's capacity) and addressed by device arithmetic on the exchanged count matrix:
    there is NO host read.  A slab smaller than the worst case (``B200_MOE_CAPACITY`` < 1) is checked by a device-side
    assert of the row count instead."""
    from internevo_b200.ops.grouped import ALIGN, aligned_offsets, padding_rows

    cm = be.exchange_counts(counts)
    El = be.num_experts // be.world
    slot_rank, slot_row, per_expert, per_rank = slot_plan(expert_of_slot, cm, be.rank, El, align=ALIGN)
    offsets = aligned_offsets(per_expert)
    R = be.max_rows
    if R < x2.shape[0] * k * be.world + El * ALIGN:      # not the worst case: fail loudly (asynchronously) on overflow
        torch._assert_async((per_rank.max() < R).all() if hasattr(torch, "_assert_async") else True)
    pad_idx = padding_rows(per_expert, offsets, dummy_row=R - 1)
    plan = _Plan(slot_rank, slot_row, R, k, x2.shape[0], pad_idx=pad_idx)
    rows = _FusedDispatch.apply(x2.contiguous(), be, plan)
    return rows, offsets, plan


def fused_combine(out_rows: torch.Tensor, w_slots: torch.Tensor, be: MoEFusedBackend, plan: _Plan) -> torch.Tensor:
    return _FusedCombine.apply(out_rows.contiguous(), w_slots.float().contiguous(), be, plan)


def fused_capacity_dispatch(x2: torch.Tensor, slot_rank: torch.Tensor, slot_row: torch.Tensor, be: MoEFusedBackend, k: int,
                            rows_per_rank: int):
    """Capacity (GShard) layout: the slab addresses are a pure function of (expert, capacity slot, source rank), so there is
    no count exchange and no host sync; ``slot_row < 0`` marks a dropped slot."""
    if rows_per_rank > be.max_rows:
        raise RuntimeError(f"fused MoE dispatch: slab of {be.max_rows} rows, {rows_per_rank} needed")
    plan = _Plan(slot_rank.contiguous(), slot_row.contiguous(), rows_per_rank, k, x2.shape[0], zero_fill=True)
    return _FusedDispatch.apply(x2.contiguous(), be, plan), plan


def fused_capacity_combine(out_rows: torch.Tensor, w_slots: torch.Tensor, be: MoEFusedBackend, plan: _Plan) -> torch.Tensor:
    return _FusedCombine.apply(out_rows.contiguous(), w_slots.float().contiguous(), be, plan)
