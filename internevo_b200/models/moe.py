"""Mixture of experts: GShard top-1 / top-2 gating with capacity, expert parallelism over the EXPERT group.

Semantics follow the reference (``internlm/model/moe/{moe,gshard_layer,experts}.py``): softmax gate in fp32, capacity
``ceil(k * S / E * capacity_factor)`` (``>= min_capacity``), tokens beyond capacity dropped in token order (or by random
priority with ``use_rts`` for top-1), second expert chosen after Gumbel noise, load-balancing loss
``l_aux = E * sum(mean(gates) * mean(mask1))``, top-2 weights renormalised, equal-split ``all_to_all`` of the ``[E, C, h]``
dispatch buffer, optional residual expert.

Implementation is index based: the reference materialises dense one-hot ``[S, E, C]`` dispatch / combine tensors and
contracts them with einsums (O(S·E·C) memory, ``gshard_layer.py:458,490``); here tokens are scattered into / gathered
from the ``[E, C, h]`` buffer by row index and every expert runs the fused SwiGLU tcgen05 GEMMs on its contiguous slab.
``MegaBlock`` is the same layer with the capacity raised to the busiest expert's load instead of dropping;
``MegaBlock-D`` is ``DroplessMOELayer``: sorted tokens, variable-split all-to-all, grouped GEMM, no padding at all.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from internevo_b200 import ops
from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.core.naive_amp import set_fp32_attr_to_module
from internevo_b200.utils.registry import MOE_INITIALIZER

from .modules import FeedForward

uniform_map = {}
gumbel_map = {}


def multiplicative_jitter(x, device, epsilon=1e-2):
    if epsilon == 0:
        return x
    u = uniform_map.get(device)
    if u is None:
        u = torch.distributions.uniform.Uniform(low=torch.tensor(1.0 - epsilon, device=device),
                                                high=torch.tensor(1.0 + epsilon, device=device)).rsample
        uniform_map[device] = u
    return x * u(x.shape)


def gumbel_rsample(shape, device):
    g = gumbel_map.get(device)
    if g is None:
        g = torch.distributions.gumbel.Gumbel(torch.tensor(0.0, device=device), torch.tensor(1.0, device=device)).rsample
        gumbel_map[device] = g
    return g(shape)


def _capacity(num_tokens: int, num_experts: int, capacity_factor: float, min_capacity: int, k: int) -> int:
    cap = math.ceil(k * num_tokens / num_experts * capacity_factor)
    return max(cap, min_capacity)


class BaseMoELayer(nn.Module):
    """Attribute contract of an MoE layer — ``gate`` (or ``wg``), ``experts`` (:class:`Experts`), ``ep_group``, ``ep_size``,
    ``num_local_experts``, and after each forward ``l_aux`` / ``exp_counts`` (reference ``moe/base_layer.py:17-41``);
    ``MoE`` and the loss / metric code only rely on these."""


class _AllToAll(torch.autograd.Function):
    """Equal-split ``all_to_all_single`` (dispatch / combine, reference ``moe/utils.py:21-40``)."""

    @staticmethod
    def forward(ctx, group, x):
        ctx.group = group
        x = x.contiguous()
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, _AllToAll.apply(ctx.group, g)


def _route(logits, k, cf, min_capacity, noisy_gate_policy, drop_tokens, use_rts):
    """Routing decision from fp32 gate logits ``[S, E]`` in index form →
    ``(l_aux, weights[S, k], experts[S, k], slots[S, k], keep[S, k], capacity, counts[E])``."""
    S, E = logits.shape
    gates = F.softmax(logits, dim=1)
    cap = _capacity(S, E, cf, min_capacity, k)
    # ---- expert choice
    noisy1 = logits + gumbel_rsample(logits.shape, logits.device) if (
        k == 1 and noisy_gate_policy == "RSample") else gates
    idx1 = torch.argmax(noisy1, dim=1)
    mask1 = F.one_hot(idx1, E)
    if k == 2:
        noisy2 = (logits + gumbel_rsample(logits.shape, logits.device)).masked_fill(mask1.bool(), float("-inf"))
        idx2 = torch.argmax(noisy2, dim=1)
        mask2 = F.one_hot(idx2, E)
    counts = mask1.sum(0).detach()
    # ---- load-balancing loss (first choice only, as GShard)
    me, ce = gates.mean(0), mask1.float().mean(0)
    l_aux = (me * ce).sum() * E
    if not drop_tokens:  # dropless: grow capacity to the busiest expert (agreed across the expert group)
        load = counts.max() if k == 1 else (mask1.sum(0) + mask2.sum(0)).max()
        if gpc.is_initialized(ParallelMode.EXPERT) and gpc.get_world_size(ParallelMode.EXPERT) > 1:
            dist.all_reduce(load, op=dist.ReduceOp.MAX, group=gpc.get_group(ParallelMode.EXPERT))
        cap = max(cap, int(load.item()))
    # ---- slot of every token inside its expert's buffer
    if k == 1 and use_rts and drop_tokens:
        # random token selection: keep the `cap` tokens with the highest random priority per expert
        pri = mask1 * torch.rand_like(mask1, dtype=torch.float32)
        top = torch.topk(pri, k=min(cap, S), dim=0).indices  # [cap, E]
        sel = torch.zeros_like(mask1).scatter_(0, top, 1) * mask1
        loc1 = (torch.cumsum(sel, 0) - 1)
        keep1 = sel.sum(1).bool()
        slot1 = (loc1 * sel).sum(1)
    else:
        loc1 = torch.cumsum(mask1, 0) - 1
        slot1 = (loc1 * mask1).sum(1)
        keep1 = slot1 < cap
    g1 = (gates * mask1).sum(1)
    if k == 1:
        weights = (g1 * keep1).unsqueeze(1)
        return l_aux, weights, idx1.unsqueeze(1), slot1.unsqueeze(1), keep1.unsqueeze(1), cap, counts
    loc2 = torch.cumsum(mask2, 0) - 1 + mask1.sum(0, keepdim=True)  # second choices queue behind all first choices
    slot2 = (loc2 * mask2).sum(1)
    keep2 = slot2 < cap
    g2 = (gates * mask2).sum(1)
    g1, g2 = g1 * keep1, g2 * keep2
    denom = (g1 + g2).clamp_min(torch.finfo(gates.dtype).eps)
    weights = torch.stack([g1 / denom, g2 / denom], 1)
    return (l_aux, weights, torch.stack([idx1, idx2], 1), torch.stack([slot1, slot2], 1),
            torch.stack([keep1, keep2], 1), cap, counts)


class TopKGate(nn.Module):
    """Gate ``wg: hidden -> num_experts`` kept in fp32 (reference ``gshard_layer.py:287-366``).

    ``forward(x[S, h])`` returns ``(l_aux, weights[S, k], experts[S, k], slots[S, k], keep[S, k], capacity, counts[E])``.
    """

    def __init__(self, model_dim: int, num_experts: int, k: int = 1, capacity_factor: float = 1.0,
                 eval_capacity_factor: float = 1.0, min_capacity: int = 8, noisy_gate_policy: Optional[str] = None,
                 drop_tokens: bool = True, use_rts: bool = True, device=None) -> None:
        super().__init__()
        assert k in (1, 2), "Only top-1 and top-2 gatings are supported."
        self.wg = nn.Linear(model_dim, num_experts, bias=False, device=device, dtype=torch.float32)
        self.k, self.num_experts = k, num_experts
        self.capacity_factor, self.eval_capacity_factor, self.min_capacity = capacity_factor, eval_capacity_factor, min_capacity
        self.noisy_gate_policy, self.drop_tokens, self.use_rts = noisy_gate_policy, drop_tokens, use_rts
        set_fp32_attr_to_module(self)

    def forward(self, x: torch.Tensor):
        xf = x.float()
        if self.noisy_gate_policy == "Jitter" and self.training:
            xf = multiplicative_jitter(xf, device=xf.device)
        logits = F.linear(xf, self.wg.weight.float())
        cf = self.capacity_factor if self.training else self.eval_capacity_factor
        return _route(logits, self.k, cf, self.min_capacity, self.noisy_gate_policy, self.drop_tokens, self.use_rts)


class Experts(nn.Module):
    """Local experts; parameters are tagged for expert-data-parallel reduction and per-expert checkpoint files
    (reference ``moe/experts.py:13-68``)."""

    def __init__(self, experts, num_local_experts=1, expert_group_name=None):
        super().__init__()
        self.wrapped_experts = nn.ModuleList(experts)
        self.num_local_experts = num_local_experts
        for expert in self.wrapped_experts:
            for p in expert.parameters():
                p.is_expert = True
                p.group_name = expert_group_name

    def _grouped_weights(self):
        """``(w13 list, w2 list)`` when every local expert is a plain (not tensor-/weight-parallel) SwiGLU ``FeedForward``:
        the grouped tcgen05 GEMM then runs all of them in one launch (``ops/grouped.py``)."""
        ok = all(isinstance(e, FeedForward) and getattr(e, "process_group", None) is None and e.tp_mode != "isp"
                 for e in self.wrapped_experts)
        if not ok:
            return None
        return [e.w13.weight for e in self.wrapped_experts], [e.w2.weight for e in self.wrapped_experts]

    def forward_packed(self, rows: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        """``rows [R, h]``: the rows of local expert ``e`` at ``offsets[e] .. offsets[e + 1]`` (int32 on the rows' device,
        multiples of 128, padding rows zero) → ``[R, h]``.  No host-side knowledge of the per-expert counts is needed."""
        ws = self._grouped_weights()
        if ws is not None and rows.dim() == 2:
            return ops.grouped_swiglu_mlp(rows, offsets, ws[0], ws[1])
        b = [int(v) for v in offsets.tolist()]
        out = torch.zeros_like(rows)
        for e, expert in enumerate(self.wrapped_experts):
            if b[e + 1] > b[e]:
                out[b[e]: b[e + 1]] = expert(rows[b[e]: b[e + 1]])
        return out

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        # inputs: [E_local, rows, h] (capacity layout: the same number of rows for every expert)
        El, R, h = inputs.shape
        if self._grouped_weights() is not None and inputs.is_cuda and inputs.dtype == torch.bfloat16:
            Ra = (R + 127) // 128 * 128           # group starts must be multiples of 128 rows (zero padding)
            x = inputs if Ra == R else F.pad(inputs, (0, 0, 0, Ra - R))
            off = torch.arange(El + 1, device=inputs.device, dtype=torch.int32) * Ra
            y = self.forward_packed(x.reshape(El * Ra, h), off).view(El, Ra, h)
            return y if Ra == R else y[:, :R]
        outs = [expert(chunk.squeeze(0)) for chunk, expert in zip(inputs.chunk(self.num_local_experts, dim=0),
                                                                   self.wrapped_experts)]
        return torch.stack(outs, 0)


class GShardMOELayer(BaseMoELayer):
    """dispatch → all-to-all → experts → all-to-all → combine (reference ``gshard_layer.py:369-498``)."""

    def __init__(self, hidden_size, gate: TopKGate, experts: Experts, ep_group, ep_size, num_local_experts: int) -> None:
        super().__init__()
        self.gate, self.experts = gate, experts
        self.ep_group, self.ep_size, self.num_local_experts = ep_group, ep_size, num_local_experts
        self.exp_counts = None
        self.l_aux = None

    def forward(self, x: torch.Tensor):
        shape = x.shape
        h = shape[-1]
        x2 = x.reshape(-1, h)
        S = x2.shape[0]
        l_aux, weights, experts, slots, keep, cap, counts = self.gate(x2)
        self.l_aux, self.exp_counts = l_aux, counts
        E = self.gate.num_experts
        k = experts.shape[1]
        flat = (experts * cap + slots).reshape(-1)            # row in the [E * cap, h] dispatch buffer
        keep_f = keep.reshape(-1)
        be = self._fused_backend(x2, E * cap)
        if be is not None:
            return self._forward_fused(x2, be, weights, experts, slots, keep_f, cap).reshape(shape)
        tok = torch.arange(S, device=x2.device).repeat_interleave(k)
        rows, toks = flat[keep_f], tok[keep_f]
        dispatched = x2.new_zeros(E * cap, h).index_copy(0, rows, x2[toks])
        dispatched = dispatched.view(E, cap, h)
        if self.ep_size > 1:
            dispatched = _AllToAll.apply(self.ep_group, dispatched)
        # [ep, E_local, cap, h] -> [E_local, ep * cap, h]
        d = dispatched.view(self.ep_size, self.num_local_experts, cap, h).transpose(0, 1).reshape(
            self.num_local_experts, self.ep_size * cap, h)
        out = self.experts(d)
        out = out.view(self.num_local_experts, self.ep_size, cap, h).transpose(0, 1).reshape(E, cap, h)
        if self.ep_size > 1:
            out = _AllToAll.apply(self.ep_group, out)
        out = out.reshape(E * cap, h)
        w = weights.reshape(-1)[keep_f].to(out.dtype)
        combined = x2.new_zeros(S, h).index_add(0, toks, out[rows] * w.unsqueeze(1))
        return combined.reshape(shape)


    # ---- peer-memory path: dispatch / combine are ONE kernel each over NVLink (parallel/moe_fused.py, csrc/moe_comm.cu)
    def _fused_backend(self, x2: torch.Tensor, rows_per_rank: int):
        if self.ep_size <= 1 or not x2.is_cuda or x2.dtype is not torch.bfloat16 or x2.shape[1] % 8 != 0:
            return None
        if os.environ.get("B200_MOE_FUSED", "1") == "0":
            return None
        from internevo_b200.parallel.moe_fused import backend_for

        return backend_for(self.ep_group, x2.shape[1], rows_per_rank, self.gate.num_experts)

    def _forward_fused(self, x2, be, weights, experts, slots, keep_f, cap):
        """The owner's slab is the ``[E_local, ep * cap, h]`` expert input itself: slot (token, j) routed to expert ``e`` with
        capacity slot ``c`` lands at row ``(e % E_local) * ep * cap + my_rank * cap + c`` of GPU ``e // E_local``; dropped
        slots carry row -1 (nothing sent, zero contribution).  Unused capacity rows are zero, as in the dense buffer."""
        from internevo_b200.parallel.moe_fused import fused_capacity_combine, fused_capacity_dispatch

        El, ep = self.num_local_experts, self.ep_size
        e_flat, c_flat = experts.reshape(-1), slots.reshape(-1)
        row = (e_flat % El) * (ep * cap) + be.rank * cap + c_flat
        slot_row = torch.where(keep_f, row, torch.full_like(row, -1)).to(torch.int32)
        slot_rank = torch.div(e_flat, El, rounding_mode="floor").to(torch.int32)
        k = experts.shape[1]
        d, plan = fused_capacity_dispatch(x2, slot_rank, slot_row, be, k, El * ep * cap)
        out = self.experts(d.view(El, ep * cap, -1)).reshape(El * ep * cap, -1)
        w = (weights.reshape(-1) * keep_f.to(weights.dtype))
        return fused_capacity_combine(out, w, be, plan)


def _make_experts(num_local, hidden_size, mlp_ratio, device, dtype, expert_tensor_parallel):
    """Local expert MLPs.  Default: every tensor rank holds the whole expert (replicated over the tensor group, grouped
    GEMM path).  ``moe.expert_tensor_parallel=True`` shards every expert's hidden dimension over the tensor group the way the
    dense MLP is sharded - column-parallel w1 / w3, row-parallel w2, the mode's collectives inside the linears - which is what
    the reference's experts do (``moe/gshard_layer.py:423-427`` builds them on the TENSOR group; ``megablock/mlp.py:37-40``)."""
    group, mode = None, "mtp"
    if expert_tensor_parallel and gpc.is_initialized(ParallelMode.TENSOR) and gpc.get_world_size(ParallelMode.TENSOR) > 1:
        mode = gpc.config.parallel["tensor"].get("mode", "mtp") if gpc.config is not None else "mtp"
        assert mode != "isp", "expert_tensor_parallel needs a tensor mode (mtp / msp / fsp); under isp experts stay replicated"
        group = gpc.get_group(ParallelMode.TENSOR)
    experts = [FeedForward(hidden_size, int(hidden_size * mlp_ratio), out_features=hidden_size, process_group=group,
                           bias=False, device=device, dtype=dtype, tp_mode=mode) for _ in range(num_local)]
    if group is not None:
        for e in experts:
            for p in e.parameters():
                p.expert_tp_sharded = True   # not a replica over the tensor group: no broadcast, no replica-grad reduction
    return experts


@MOE_INITIALIZER.register_module("GShard")
def _build_gshard(hidden_size, num_experts, ep_group, ep_size, mlp_ratio, device, dtype, top_k=1, capacity_factor=1.0,
                  eval_capacity_factor=1.0, min_capacity=4, noisy_gate_policy=None, drop_tokens=True, use_rts=True,
                  expert_tensor_parallel=False, **unused):
    assert noisy_gate_policy is None or noisy_gate_policy in ("None", "Jitter", "RSample")
    num_local = num_experts // ep_size
    experts = _make_experts(num_local, hidden_size, mlp_ratio, device, dtype, expert_tensor_parallel)
    gate = TopKGate(hidden_size, num_experts, top_k, capacity_factor, eval_capacity_factor, min_capacity,
                    noisy_gate_policy, drop_tokens, use_rts, device=device)
    cls = unused.pop("_layer_cls", None) or GShardMOELayer
    return cls(hidden_size, gate, Experts(experts, num_local, f"moe_ep_size_{ep_size}"), ep_group, ep_size, num_local)


class _AllToAllV(torch.autograd.Function):
    """Variable-split ``all_to_all_single`` over rows (reference ``moe/megablock/megablock_moe.py:155-247``)."""

    @staticmethod
    def forward(ctx, group, x, send_splits, recv_splits):
        ctx.group, ctx.send, ctx.recv = group, send_splits, recv_splits
        out = x.new_empty(sum(recv_splits), *x.shape[1:])
        dist.all_to_all_single(out, x.contiguous(), output_split_sizes=recv_splits, input_split_sizes=send_splits,
                               group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, _AllToAllV.apply(ctx.group, g, ctx.recv, ctx.send), None, None


class DroplessMOELayer(BaseMoELayer):
    """Dropless top-k MoE (the ``MegaBlock-D`` registry entry): no capacity, no padding.

    Tokens are sorted by destination expert, exchanged with ONE variable-split all-to-all (row counts first, then the
    rows), every local expert runs its fused SwiGLU GEMMs on its contiguous, exactly-sized slab (a grouped GEMM: the
    tcgen05 kernel takes any M, ragged tiles are zero-filled by TMA), and results travel back through the inverse
    permutation.  Replaces the block-sparse SDD/DSD kernels of MegaBlocks (reference ``moe/megablock/*``): on Blackwell
    the dense grouped form keeps the 128 x 256 tensor-core tile busy without a block topology."""

    def __init__(self, hidden_size, num_experts, ep_group, ep_size, experts: Experts, top_k: int = 1,
                 noisy_gate_policy=None, device=None):
        super().__init__()
        self.wg = nn.Linear(hidden_size, num_experts, bias=False, device=device, dtype=torch.float32)
        set_fp32_attr_to_module(self.wg)
        self.experts = experts
        self.num_experts, self.top_k = num_experts, top_k
        self.ep_group, self.ep_size = ep_group, ep_size
        self.num_local_experts = num_experts // ep_size
        self.noisy_gate_policy = noisy_gate_policy
        self.l_aux, self.exp_counts = None, None

    def _fused_backend(self, x2: torch.Tensor, n_slots: int):
        """Peer-memory dispatch / combine when the expert group spans 2 / 4 / 8 NVLink-connected GPUs (``B200_MOE_FUSED=0``
        forces the NCCL all-to-all path, which stays as fall-back and numerical oracle)."""
        if self.ep_size <= 1 or not x2.is_cuda or x2.dtype is not torch.bfloat16 or x2.shape[1] % 8 != 0:
            return None
        if os.environ.get("B200_MOE_FUSED", "1") == "0":
            return None
        from internevo_b200.parallel.moe_fused import backend_for

        # worst case: every slot of every rank lands on one GPU; B200_MOE_CAPACITY < 1 trades memory for an overflow check
        # + one 128-row alignment gap per local expert and a reserved dummy row (ops/grouped.py::padding_rows)
        max_rows = max(1, int(n_slots * self.ep_size * float(os.environ.get("B200_MOE_CAPACITY", "1.0")))) \
            + (self.num_local_experts + 1) * 128
        return backend_for(self.ep_group, x2.shape[1], max_rows, self.num_experts)

    def forward(self, x: torch.Tensor):
        shape = x.shape
        h = shape[-1]
        x2 = x.reshape(-1, h)
        S, E, k = x2.shape[0], self.num_experts, self.top_k
        xf = x2.float()
        if self.noisy_gate_policy == "Jitter" and self.training:
            xf = multiplicative_jitter(xf, device=xf.device)
        gates = F.softmax(F.linear(xf, self.wg.weight.float()), dim=1)
        w, idx = torch.topk(gates, k, dim=1)                      # [S, k]
        if k > 1:
            w = w / w.sum(1, keepdim=True).clamp_min(torch.finfo(w.dtype).eps)
        mask1 = F.one_hot(idx[:, 0], E)
        self.exp_counts = mask1.sum(0).detach()
        self.l_aux = (gates.mean(0) * mask1.float().mean(0)).sum() * E
        # ---- sort assignments by expert: rows of expert e are contiguous, experts of one ep rank are adjacent
        flat_e = idx.reshape(-1)
        order = torch.argsort(flat_e, stable=True)
        tok = torch.arange(S, device=x2.device).repeat_interleave(k)[order]
        counts = torch.bincount(flat_e, minlength=E)
        be = self._fused_backend(x2, S * k)
        if be is not None:
            # dispatch / combine as ONE kernel each over NVLink peer memory (parallel/moe_fused.py, csrc/moe_comm.cu)
            from internevo_b200.parallel.moe_fused import fused_combine, fused_dispatch

            # every slot gets its address in the owner's group-aligned slab from the exchanged count matrix; the experts
            # run as ONE grouped GEMM per projection over that slab - no host read of the counts anywhere
            rows, offsets, plan = fused_dispatch(x2, flat_e, counts, be, k)
            out = self.experts.forward_packed(rows, offsets)
            return fused_combine(out, w.reshape(-1), be, plan).reshape(shape)
        send = x2[tok]
        if self.ep_size > 1:
            recv_counts = torch.empty_like(counts)
            dist.all_to_all_single(recv_counts, counts, group=self.ep_group)   # [ep, E_local] rows coming from each rank
            send_splits = counts.view(self.ep_size, -1).sum(1).tolist()
            rc = recv_counts.view(self.ep_size, self.num_local_experts)
            recv_splits = rc.sum(1).tolist()
            recv = _AllToAllV.apply(self.ep_group, send, send_splits, recv_splits)
            # received rows are grouped (source rank, local expert): regroup by local expert
            src_expert = torch.repeat_interleave(
                torch.arange(self.num_local_experts, device=x2.device).repeat(self.ep_size), rc.reshape(-1))
            regroup = torch.argsort(src_expert, stable=True)
            rows = recv[regroup]
            per_expert = rc.sum(0)
        else:
            rows, per_expert, regroup = send, counts, None
        # grouped GEMM over group-aligned slabs: rows of local expert e move to offsets[e] + (position inside e); the padding
        # rows stay zero.  The buffer size depends on the row count only (known from the all-to-all splits), not on the routing.
        El = self.num_local_experts
        offsets = ops.aligned_offsets(per_expert)
        exact = per_expert.cumsum(0) - per_expert
        e_of_row = torch.repeat_interleave(torch.arange(El, device=x2.device), per_expert, output_size=rows.shape[0])
        dest = offsets[:-1].to(torch.int64)[e_of_row] + torch.arange(rows.shape[0], device=x2.device) - exact[e_of_row]
        cap_rows = (rows.shape[0] + El * 127 + 127) // 128 * 128
        packed = rows.new_zeros(cap_rows, h).index_copy(0, dest, rows)
        out = self.experts.forward_packed(packed, offsets)[dest]
        if self.ep_size > 1:
            back = torch.empty_like(out).index_copy(0, regroup, out)
            out = _AllToAllV.apply(self.ep_group, back, recv_splits, send_splits)
        wsel = w.reshape(-1)[order].to(out.dtype)
        combined = x2.new_zeros(S, h).index_add(0, tok, out * wsel.unsqueeze(1))
        return combined.reshape(shape)


@MOE_INITIALIZER.register_module("MegaBlock")
def _build_megablock(**kw):
    kw["drop_tokens"] = False  # capacity-padded variant: padded to the busiest expert instead of dropping
    kw["_layer_cls"] = MegaBlockMoE
    return _build_gshard(**kw)


@MOE_INITIALIZER.register_module("MegaBlock-D")
def _build_megablock_d(hidden_size, num_experts, ep_group, ep_size, mlp_ratio, device, dtype, top_k=1,
                       noisy_gate_policy=None, expert_tensor_parallel=False, **unused):
    num_local = num_experts // ep_size
    experts = _make_experts(num_local, hidden_size, mlp_ratio, device, dtype, expert_tensor_parallel)
    return MegaBlockdMoE(hidden_size, num_experts, ep_group, ep_size, Experts(experts, num_local, f"moe_ep_size_{ep_size}"),
                         top_k=top_k, noisy_gate_policy=noisy_gate_policy, device=device)


class MoE(nn.Module):
    """Wrapper selected by ``model.moe_type`` with kwargs from the top-level ``moe = dict(...)`` config; optional residual
    expert mixed in by a learned 2-way coefficient (reference ``moe/moe.py:13-99``).
    ``forward(x) -> (output, l_aux, exp_counts)``."""

    def __init__(self, hidden_size, num_experts=1, ep_group=None, ep_size=None, device=None, dtype=None, mlp_ratio=4.0,
                 moe_use_residual=False, moe_type="GShard", **moe_kwargs):
        super().__init__()
        ep_size = ep_size or gpc.get_world_size(ParallelMode.EXPERT)
        ep_group = ep_group or gpc.get_group(ParallelMode.EXPERT)
        assert num_experts % ep_size == 0, f"num_experts ({num_experts}) must be divisible by ep size ({ep_size})"
        self.ep_size, self.num_experts = ep_size, num_experts
        self.num_local_experts = num_experts // ep_size
        self.moe_layer = MOE_INITIALIZER.get_module(moe_type)(
            hidden_size=hidden_size, num_experts=num_experts, ep_group=ep_group, ep_size=ep_size, mlp_ratio=mlp_ratio,
            device=device, dtype=dtype, **moe_kwargs)
        self.use_residual = moe_use_residual
        if self.use_residual:
            self.residual_mlp = FeedForward(hidden_size, int(hidden_size * mlp_ratio), out_features=hidden_size,
                                            process_group=gpc.get_group(ParallelMode.TENSOR), bias=False, device=device,
                                            dtype=dtype)
            self.coefficient = nn.Linear(hidden_size, 2, device=device, dtype=dtype)

    def forward(self, hidden_states, used_token=None):
        output = self.moe_layer(hidden_states)
        if self.use_residual:
            output_mlp = self.residual_mlp(hidden_states)
            if isinstance(output_mlp, tuple):
                output_mlp = output_mlp[0]
            coef = F.softmax(self.coefficient(hidden_states), dim=-1)
            output = output * coef[..., 0:1] + output_mlp * coef[..., 1:]
        return output, self.moe_layer.l_aux, self.moe_layer.exp_counts


def is_moe_param(param: torch.Tensor) -> bool:
    return getattr(param, "is_expert", False)


def all_to_all(x, output_split_sizes=None, input_split_sizes=None, group=None, async_op=False):
    """Differentiable ``all_to_all_single`` with optional row splits (reference ``moe/utils.py:62-63``)."""
    assert not async_op, "the autograd all_to_all is synchronous; overlap comes from the fused peer-memory dispatch"
    if output_split_sizes is None and input_split_sizes is None:
        return _AllToAll.apply(group, x)
    return _AllToAllV.apply(group, x, list(input_split_sizes), list(output_split_sizes))


def _dense_routing(l_aux, weights, experts, slots, keep, cap, counts, num_experts):
    """Index form → the dense GShard tensors ``combine_weights [S, E, C]`` / ``dispatch_mask [S, E, C]``."""
    S, k = experts.shape
    combine = weights.new_zeros(S, num_experts, cap)
    tok = torch.arange(S, device=experts.device).unsqueeze(1).expand(S, k)
    combine.index_put_((tok[keep], experts[keep], slots[keep]), weights[keep], accumulate=True)
    return l_aux, combine, combine.bool(), counts


def top1gating(logits, capacity_factor, min_capacity, used_token=None, noisy_gate_policy=None, drop_tokens=True,
               use_rts=True):
    """Top-1 gating on logits in the dense form of the reference (``gshard_layer.py:138-218``) →
    ``(l_aux, combine_weights[S, E, C], dispatch_mask[S, E, C], exp_counts[E])``.  The layers use the index form
    (:func:`_route`); this is the O(S·E·C) view of the same decision for inspection and tests."""
    assert used_token is None, "token masks are applied by the caller in this framework"
    E = logits.shape[1]
    return _dense_routing(*_route(logits.float(), 1, capacity_factor, min_capacity, noisy_gate_policy, drop_tokens, use_rts), E)


def top2gating(logits, capacity_factor, min_capacity):
    """Top-2 gating (second expert after Gumbel noise, weights renormalised; reference ``:221-284``), dense form."""
    E = logits.shape[1]
    return _dense_routing(*_route(logits.float(), 2, capacity_factor, min_capacity, None, True, False), E)


# ---------------------------------------------------------------------------------------------------------------------
# Routing arithmetic in the MegaBlocks vocabulary (reference ``megablock_moe.py:73-99,253-275``, ``megablock_dmoe.py:85-182``):
# sort-based "indices and bins", capacity, load-balancing loss, binned gather / scatter around the experts.  The layers above
# do the same work inline (fused with their all-to-all); these are the stand-alone pieces for user code and tests.
# ---------------------------------------------------------------------------------------------------------------------
class _BinnedRouting:
    blocking = 128

    def _n_experts(self) -> int:
        return self.gate.num_experts if hasattr(self, "gate") else self.num_experts

    def _top_k(self) -> int:
        return self.gate.k if hasattr(self, "gate") else self.top_k

    def expert_capacity(self, tokens: int, top_k: int) -> int:
        """Rows one expert accepts when tokens are dropped: ``capacity_factor * top_k * tokens * ep / E``."""
        cf = getattr(getattr(self, "gate", None), "capacity_factor", 1.0)
        return int(cf * top_k * tokens * max(1, self.ep_size) / self._n_experts())

    def indices_and_bins(self, top_expert: torch.Tensor):
        """``(indices, bin_ids, bins, tokens_per_expert)``: the stable order that groups the (token, choice) slots by expert,
        the sorted expert ids, the inclusive running row count per expert and the row count per expert."""
        flat = top_expert.reshape(-1).to(torch.int64)
        bin_ids, indices = torch.sort(flat, stable=True)
        tokens_per_expert = torch.bincount(flat, minlength=self._n_experts())
        return indices.to(torch.int32), bin_ids.to(torch.int32), tokens_per_expert.cumsum(0).to(torch.int32), \
            tokens_per_expert.to(torch.int32)

    def indices_and_padded_bins(self, selected_experts: torch.Tensor):
        """As ``indices_and_bins`` plus ``padded_bins``: the running count with every expert rounded up to the 128-row block of
        the grouped GEMM (= ``ops.aligned_offsets(tokens_per_expert)[1:]``)."""
        indices, bin_ids, bins, tpe = self.indices_and_bins(selected_experts)
        padded = (tpe.to(torch.int64) + self.blocking - 1) // self.blocking * self.blocking
        return indices, bin_ids, bins, padded.cumsum(0).to(torch.int32), tpe

    def load_balancing_loss(self, tokens_per_expert: torch.Tensor, expert_scores: torch.Tensor) -> torch.Tensor:
        """``E / (tokens * k) * <tokens_per_expert, mean score per expert>`` (Switch-style auxiliary loss)."""
        assert expert_scores.dim() == 2 and tokens_per_expert.dim() == 1
        tokens, E = expert_scores.shape
        assert E == self._n_experts() == tokens_per_expert.numel()
        return (E / (tokens * self._top_k())) * torch.dot(tokens_per_expert.to(expert_scores.dtype), expert_scores.mean(0))

    def topology(self, x: torch.Tensor, padded_bins: torch.Tensor):
        """Block-diagonal structure of the expert activations for a row buffer ``x`` laid out by ``padded_bins``."""
        from .megablock import Topology

        assert x.shape[0] % self.blocking == 0
        off = torch.cat([padded_bins.new_zeros(1), padded_bins]).to(torch.int32)
        first = self.experts.wrapped_experts[0]
        ffn = first.w2.weight.shape[1] if hasattr(first, "w2") else next(first.parameters()).shape[0]
        return Topology(off, ffn, rows=x.shape[0])

    @staticmethod
    def sparse_transpose(size, row_indices: torch.Tensor, column_indices: torch.Tensor, blocking: int = 128):
        """Transpose metadata of a block-COO pattern ``(row_indices, column_indices)`` (in row-major block order):
        ``(column_indices_t, offsets_t, block_offsets_t)`` = the block rows listed column by column, the CSC offsets, and for
        every transposed block the position of its data in the original order."""
        order = torch.sort(column_indices.to(torch.int64), stable=True)[1]
        per_col = torch.bincount(column_indices.to(torch.int64), minlength=size[1] // blocking)
        offsets_t = torch.cat([per_col.new_zeros(1), per_col.cumsum(0)]).to(torch.int32)
        return row_indices.gather(0, order), offsets_t, order.to(torch.int32)

    def permute_and_compute(self, x, indices, expert_weights, bins, expert_capacity, top_k):
        """Binned gather -> experts on ``[E, capacity, h]`` -> weighted binned scatter: slot ``indices[j]`` (token
        ``indices[j] // top_k``) is the ``j - bins[e - 1]``-th row of its expert ``e`` and is dropped beyond ``expert_capacity``."""
        x = x.reshape(-1, x.shape[-1])
        E = bins.numel()
        idx = indices.to(torch.int64)
        j = torch.arange(idx.numel(), device=x.device)
        expert = torch.searchsorted(bins.to(torch.int64), j, right=True)
        start = torch.cat([bins.new_zeros(1), bins[:-1]]).to(torch.int64)
        slot = j - start[expert]
        keep = slot < expert_capacity
        rows = (expert * expert_capacity + slot)[keep]
        toks = torch.div(idx, top_k, rounding_mode="floor")[keep]
        buf = x.new_zeros(E * expert_capacity, x.shape[1]).index_copy(0, rows, x[toks])
        out = self.experts(buf.view(E, expert_capacity, -1)).reshape(E * expert_capacity, -1)
        w = expert_weights.reshape(-1)[idx][keep].to(out.dtype)
        return x.new_zeros(x.shape).index_add(0, toks, out[rows] * w.unsqueeze(1))


class MegaBlockMoE(_BinnedRouting, GShardMOELayer):
    """Capacity-padded MegaBlocks layer: the GShard data path with the capacity raised to the busiest expert's load
    (``_build_megablock``) plus the binned-routing helpers."""


class MegaBlockdMoE(_BinnedRouting, DroplessMOELayer):
    """Dropless MegaBlocks layer (``MegaBlock-D``)."""


# reference class names (``moe/utils.py``)
AllToAll = _AllToAll


def einsum(rule: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """The handful of contractions of the dense GShard formulation written as broadcasts / matmuls (they map to GEMMs instead
    of the generic einsum planner), anything else falls through to ``torch.einsum`` (reference ``gshard_layer.py:75-112``)."""
    if rule == "s,se->se":
        return a.reshape(-1, 1) * b
    if rule == "se,sc->sec":
        return a.unsqueeze(2) * b.unsqueeze(1)
    if rule == "se,se->s":
        return (a * b).sum(-1)
    if rule == "sec,sm->ecm":
        s_, e_, c_ = a.shape
        return (a.reshape(s_, e_ * c_).t() @ b).reshape(e_, c_, b.shape[1])
    if rule == "sec,ecm->sm":
        return a.reshape(a.shape[0], -1) @ b.reshape(-1, b.shape[-1])
    if rule == "ks,ksm->sm":
        return (a.unsqueeze(-1) * b).sum(0)
    return torch.einsum(rule, a, b)


# MegaBlocks-style functional API (sdd / dsd over the grouped GEMM, expert MLP modules): reference ``moe/megablock/utils.py``, ``mlp.py``
from .megablock import (  # noqa: E402,F401
    BlockDiagonal,
    MegaBlockFeedForward,
    MegaBlockGroupedFeedForward,
    TensorParallelBmm,
    TensorParallelDsdNn,
    TensorParallelSddNt,
    Topology,
    WeightParallelDsdNn,
    WeightParallelSddNt,
    act_fn,
    check_megablock_installed,
    check_stk_installed,
    dsd_nn,
    promote_scalar,
    sdd_nt,
    tensor_parallel_bmm,
)
