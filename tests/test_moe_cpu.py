"""MoE: gating semantics (capacity, renormalised top-2 weights, l_aux) and expert-parallel training on CPU/gloo."""
import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config


def test_top2_gate_semantics():
    from internevo_b200.core.context import Config, global_context as gpc
    from internevo_b200.models.moe import TopKGate

    gpc.set_config(Config(dict(parallel=dict(tensor=dict(size=1, mode="mtp")))))
    torch.manual_seed(0)
    gate = TopKGate(16, 4, k=2, capacity_factor=1.0, min_capacity=2)
    x = torch.randn(32, 16)
    l_aux, w, e, slot, keep, cap, counts = gate(x)
    assert cap == 16 and w.shape == (32, 2) and e.shape == (32, 2)
    assert (e[:, 0] != e[:, 1]).all(), "second expert must differ from the first"
    both = keep.all(1)
    assert torch.allclose(w[both].sum(1), torch.ones(int(both.sum())), atol=1e-5)
    assert (slot[keep] < cap).all() and counts.sum() == 32
    # no two kept tokens share a buffer row
    rows = (e * cap + slot)[keep]
    assert rows.unique().numel() == rows.numel()
    assert l_aux.item() > 0


def test_top1_capacity_drops():
    from internevo_b200.models.moe import TopKGate

    torch.manual_seed(0)
    gate = TopKGate(8, 2, k=1, capacity_factor=0.5, min_capacity=1, use_rts=False)
    l_aux, w, e, slot, keep, cap, counts = gate(torch.randn(16, 8))
    assert cap == 4
    for ex in range(2):
        assert int((keep[:, 0] & (e[:, 0] == ex)).sum()) <= cap
    assert (w[~keep] == 0).all()


def _train_moe(rank, world, kw):
    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=2, num_experts=4, **kw)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2, capacity_factor=2.0, eval_capacity_factor=2.0, min_capacity=4, noisy_gate_policy=None,
                      drop_tokens=True, use_rts=False)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    trainer, opt, model, _ = build_trainer(cfg)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses = []
    for step in range(4):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=rank)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok and len(out) == 4
        losses.append(float(out[2]))
    return losses, sorted(norms)


def test_moe_single_process_trains():
    losses, groups = run_distributed(_train_moe, 1, {})[0]
    assert losses[-1] < losses[0]
    assert any(g.startswith("moe_ep_size") for g in groups)


def test_moe_expert_parallel_dp2():
    res = run_distributed(_train_moe, 2, {})
    for losses, groups in res:
        assert losses[-1] < losses[0]
        assert "moe_ep_size_2" in groups


def test_dropless_layer_matches_dense_reference():
    """MegaBlock-D: sorted / grouped computation == per-token sum_k w_k * expert_k(x), and gradients flow to every part."""
    from internevo_b200.models.modules import FeedForward
    from internevo_b200.models.moe import DroplessMOELayer, Experts

    torch.manual_seed(0)
    h, E, S = 16, 4, 40
    experts = [FeedForward(h, 32, out_features=h, process_group=None, bias=False, dtype=torch.float32) for _ in range(E)]
    layer = DroplessMOELayer(h, E, None, 1, Experts(experts, E, "moe_ep_size_1"), top_k=2)
    x = torch.randn(S, h, requires_grad=True)
    y = layer(x)
    gates = torch.softmax(x.detach() @ layer.wg.weight.t(), 1)
    w, idx = torch.topk(gates, 2, 1)
    w = w / w.sum(1, keepdim=True)
    ref = torch.zeros_like(y)
    for t in range(S):
        for j in range(2):
            ref[t] += w[t, j] * experts[int(idx[t, j])](x.detach()[t:t + 1])[0]
    assert torch.allclose(y, ref, atol=1e-5), (y - ref).abs().max()
    (y.sum() + layer.l_aux).backward()
    assert x.grad is not None and layer.wg.weight.grad is not None
    assert all(p.grad is not None for e in experts for p in e.parameters())
    assert int(layer.exp_counts.sum()) == S


def _train_dropless(rank, world, kw):
    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=2, num_experts=4, moe_type="MegaBlock-D", **kw)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    trainer, opt, model, _ = build_trainer(cfg)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses = []
    for step in range(4):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=rank)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok
        losses.append(float(out[2]))
    return losses


def test_dropless_expert_parallel_dp2():
    for losses in run_distributed(_train_dropless, 2, {}):
        assert losses[-1] < losses[0]


def test_fused_dispatch_slot_plan_matches_all_to_all_order():
    """The peer-memory dispatch addresses every routed slot by (owner rank, slab row).  Simulate `world` ranks in one
    process: pushing rows by `slot_plan` must build, on every owner, exactly the slab that the variable-split all-to-all +
    regroup-by-expert path builds (rows of a local expert contiguous, ordered by source rank then arrival), and pulling them
    back by the same addresses with the gate weights must reproduce the combine."""
    from internevo_b200.parallel.moe_fused import slot_plan

    torch.manual_seed(0)
    world, El, k, S, H = 4, 2, 2, 37, 8
    E = world * El
    xs = [torch.randn(S, H) for _ in range(world)]
    experts = [torch.stack([torch.randperm(E)[:k] for _ in range(S)]).reshape(-1) for _ in range(world)]  # [S * k]
    weights = [torch.rand(S * k) for _ in range(world)]
    counts = torch.stack([torch.bincount(e, minlength=E) for e in experts])          # [world, E]
    plans = [slot_plan(experts[r], counts, r, El) for r in range(world)]
    # ---- push
    slabs = [torch.full((int(plans[0][3][d]), H), float("nan")) for d in range(world)]
    for r in range(world):
        slot_rank, slot_row, per_expert, per_rank = plans[r]
        assert per_rank.tolist() == [int(counts[:, d * El:(d + 1) * El].sum()) for d in range(world)]
        for s in range(S * k):
            slabs[int(slot_rank[s])][int(slot_row[s])] = xs[r][s // k]
    # ---- oracle: what all_to_all_v + regroup gives on owner d
    for d in range(world):
        assert not torch.isnan(slabs[d]).any()                      # every row written exactly where expected
        recv = []
        for el in range(El):
            for src in range(world):
                e = d * El + el
                sel = (experts[src] == e).nonzero().flatten()       # arrival order = slot order (stable sort)
                recv.append(xs[src][sel // k])
        oracle = torch.cat(recv, 0)
        assert torch.equal(slabs[d], oracle)
        assert plans[d][2].tolist() == [int(counts[:, d * El + el].sum()) for el in range(El)]
    # ---- pull back with weights ("expert" = multiply by (global expert id + 1))
    outs = []
    for d in range(world):
        per_expert = plans[d][2].tolist()
        chunks, start = [], 0
        for el, n in enumerate(per_expert):
            chunks.append(slabs[d][start:start + n] * (d * El + el + 1))
            start += n
        outs.append(torch.cat(chunks, 0))
    for r in range(world):
        slot_rank, slot_row = plans[r][0], plans[r][1]
        got = torch.zeros(S, H)
        for s in range(S * k):
            got[s // k] += weights[r][s] * outs[int(slot_rank[s])][int(slot_row[s])]
        want = torch.zeros(S, H)
        for s in range(S * k):
            want[s // k] += weights[r][s] * xs[r][s // k] * (int(experts[r][s]) + 1)
        assert torch.allclose(got, want, atol=1e-5)


def _train_moe_tp(rank, world, kw):
    """MoE under tensor parallelism (experts and gates are replicated over the tensor group, not sharded): returns the loss
    trajectory and whether every replicated parameter is still bit-identical on all tensor ranks after training."""
    import torch.distributed as dist

    from internevo_b200.core.context import ParallelMode
    from internevo_b200.core.context import global_context as gpc

    kw = dict(kw)
    moe_type = kw.pop("moe_type", "MegaBlock-D")
    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=2, num_experts=4, moe_type=moe_type, **kw)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    trainer, opt, model, _ = build_trainer(cfg)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses = []
    for step in range(4):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=0)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok
        losses.append(float(out[2]) if out[2] is not None else None)
    same = True
    group = gpc.get_group(ParallelMode.TENSOR)
    for name, p in model.named_parameters():
        if getattr(p, "is_expert", False) or name.endswith("wg.weight") or "norm" in name:
            both = [torch.empty_like(p.data) for _ in range(gpc.get_world_size(ParallelMode.TENSOR))]
            dist.all_gather(both, p.data.contiguous(), group=group)
            same &= all(torch.equal(both[0], b) for b in both[1:])
    return losses, same


@pytest.mark.parametrize("kw", [dict(tp=2), dict(tp=2, mode="msp"), dict(tp=2, mode="fsp", moe_type="GShard"),
                                dict(tp=2, pp=2), dict(tp=2, wp=2, mode="isp", moe_type="GShard")],
                         ids=["mtp", "msp_dropless", "fsp_gshard", "mtp_pp2", "isp_gshard"])
def test_moe_under_tensor_parallel_keeps_replicas_identical(kw):
    """Regressions: (1) the replica gradient all-reduce used a slice whose size differed between tensor ranks when a
    row-parallel bias lives on rank 0 only (v1 blocks); (2) the dropless gate was not tagged as a replica; (3) replicated
    experts were initialised from the per-rank tensor RNG stream and never reduced under sequence parallelism."""
    world = 4 if kw.get("pp", 1) > 1 else 2
    res = run_distributed(_train_moe_tp, world, kw, timeout=600)
    reported = [r for r in res if r[0][0] is not None]
    assert reported and all(r[1] for r in res), [r[1] for r in res]
    for losses, _ in reported:
        if kw.get("mode") != "isp":                     # isp: every sequence rank reports the mean over its own shard
            assert losses == reported[0][0]             # tensor ranks see the same loss
        assert losses[-1] < losses[0]


def _train_v1_tp(rank, world, mode):
    import torch.distributed as dist

    from internevo_b200.core.context import ParallelMode
    from internevo_b200.core.context import global_context as gpc

    cfg = tiny_config(model_type="INTERNLM", num_layers=2, micro_num=2, tp=2, mode=mode)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    trainer, opt, model, _ = build_trainer(cfg)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses = []
    for step in range(4):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=0)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok
        losses.append(float(out[2]))
    same = True
    for name, p in model.named_parameters():
        if "norm" in name:
            both = [torch.empty_like(p.data) for _ in range(2)]
            dist.all_gather(both, p.data.contiguous(), group=gpc.get_group(ParallelMode.TENSOR))
            same &= torch.equal(both[0], both[1])
    n_params = sum(1 for _ in model.parameters())
    return losses, same, n_params


@pytest.mark.parametrize("mode", ["msp", "fsp"])
def test_internlm_v1_bias_blocks_under_sequence_parallel(mode):
    """InternLM v1 blocks carry biases; the row-parallel bias exists on tensor rank 0 only, so the ranks' gradient arenas
    have different layouts - the replica (norm) gradient all-reduce must still line up."""
    res = run_distributed(_train_v1_tp, 2, mode, timeout=600)
    assert res[0][2] != res[1][2]                       # the premise: different parameter counts per tensor rank
    assert res[0][0] == res[1][0] and res[0][0][-1] < res[0][0][0]
    assert res[0][1] and res[1][1]


def test_aligned_slot_plan_keeps_expert_slabs_contiguous_and_128_aligned():
    """``slot_plan(..., align=128)``: every local expert's slab starts at a multiple of 128 rows, holds its rows in the same
    (source rank, arrival) order as the unaligned plan, and ``aligned_offsets`` / ``padding_rows`` describe exactly the gaps."""
    from internevo_b200.ops.grouped import aligned_offsets, padding_rows
    from internevo_b200.parallel.moe_fused import slot_plan

    torch.manual_seed(1)
    world, El, k, S = 2, 3, 2, 301
    E = world * El
    experts = [torch.stack([torch.randperm(E)[:k] for _ in range(S)]).reshape(-1) for _ in range(world)]
    counts = torch.stack([torch.bincount(e, minlength=E) for e in experts])
    for d in range(world):
        written = {}
        for r in range(world):
            rank0, row0, per0, _ = slot_plan(experts[r], counts, r, El)
            rank1, row1, per1, per_rank = slot_plan(experts[r], counts, r, El, align=128)
            assert torch.equal(rank0, rank1) and torch.equal(per0, per1)
            if r == d:
                per_expert, slab_rows = per1, int(per_rank[d])
            for s in (rank1 == d).nonzero().flatten().tolist():
                written[int(row1[s])] = (int(experts[r][s]) - d * El, int(row0[s]))
        off = aligned_offsets(per_expert)
        assert slab_rows == int(off[-1]) and all(int(o) % 128 == 0 for o in off)
        exact = per_expert.cumsum(0) - per_expert
        for row, (el, row_unaligned) in written.items():
            assert int(off[el]) <= row < int(off[el]) + int(per_expert[el])
            assert row - int(off[el]) == row_unaligned - int(exact[el])          # same order inside the expert
        assert len(written) == int(per_expert.sum())
        pads = set(padding_rows(per_expert, off, dummy_row=10**6).tolist()) - {10**6}
        assert pads == set(range(int(off[-1]))) - set(written)


def test_grouped_swiglu_mlp_matches_the_per_expert_loop():
    """``ops.grouped_swiglu_mlp`` (reference path on CPU, the grouped tcgen05 kernels on GPU - tests/test_kernels_gpu.py)
    against one ``FeedForward`` call per expert: outputs, input gradient and every weight gradient."""
    from internevo_b200 import ops
    from internevo_b200.core.context import Config, global_context as gpc
    from internevo_b200.models.modules import FeedForward

    prev_cfg = gpc.config
    gpc.set_config(Config(dict(parallel=dict(sequence_parallel=False, tensor=dict(size=1, mode="mtp")))))
    torch.manual_seed(2)
    h, F_, El = 32, 64, 3
    experts = [FeedForward(h, F_, out_features=h, process_group=None, bias=False, dtype=torch.float32, multiple_of=32)
               for _ in range(El)]
    counts = torch.tensor([70, 0, 200])
    off = ops.aligned_offsets(counts)
    assert off.tolist() == [0, 128, 128, 384]
    x = torch.zeros(int(off[-1]), h)
    for e in range(El):
        x[int(off[e]): int(off[e]) + int(counts[e])] = torch.randn(int(counts[e]), h)
    x.requires_grad_(True)
    y = ops.grouped_swiglu_mlp(x, off, [e.w13.weight for e in experts], [e.w2.weight for e in experts])
    gy = torch.randn_like(y)
    for e in range(El):
        gy[int(off[e]) + int(counts[e]): int(off[e + 1])] = 0      # padding rows carry no gradient
    y.backward(gy)
    got = (y.detach().clone(), x.grad.clone(), [e.w13.weight.grad.clone() for e in experts],
           [e.w2.weight.grad.clone() for e in experts])
    x.grad = None
    for e in experts:
        e.zero_grad()
    outs = torch.zeros_like(y)
    for e in range(El):
        a, b = int(off[e]), int(off[e]) + int(counts[e])
        if b > a:
            outs[a:b] = experts[e](x[a:b])
    outs.backward(gy)
    assert torch.allclose(got[0], outs.detach(), atol=1e-5) and torch.allclose(got[1], x.grad, atol=1e-5)
    for e in range(El):
        w13g = experts[e].w13.weight.grad if experts[e].w13.weight.grad is not None else torch.zeros_like(got[2][e])
        w2g = experts[e].w2.weight.grad if experts[e].w2.weight.grad is not None else torch.zeros_like(got[3][e])
        assert torch.allclose(got[2][e], w13g, atol=1e-4) and torch.allclose(got[3][e], w2g, atol=1e-4)
    gpc._config = prev_cfg      # this test runs in-process: leave the global context as it was found


def _sharded_expert_matches_dense(rank, world):
    """``moe.expert_tensor_parallel``: one expert sharded over tp = 2 (column-parallel w13, row-parallel w2) against the dense
    MLP built from the all-gathered shards - output and input gradient (mtp: replicated tokens)."""
    import torch.distributed as dist

    from internevo_b200.core.context import ParallelMode
    from internevo_b200.core.context import global_context as gpc
    from internevo_b200.initialize import initialize_distributed_env
    from internevo_b200.models.moe import _make_experts
    from internevo_b200.ops.swiglu import swiglu_interleaved

    initialize_distributed_env(config=tiny_config(tp=2), launcher="torch", seed=7)
    group = gpc.get_group(ParallelMode.TENSOR)
    (expert,) = _make_experts(1, 64, 2.0, torch.device("cpu"), torch.float32, True)
    assert expert.process_group is group and all(getattr(p, "expert_tp_sharded", False) for p in expert.parameters())
    w13, w2 = expert.w13.weight.detach(), expert.w2.weight.detach()
    parts13 = [torch.empty_like(w13) for _ in range(2)]
    parts2 = [torch.empty_like(w2) for _ in range(2)]
    dist.all_gather(parts13, w13.contiguous(), group=group)
    dist.all_gather(parts2, w2.contiguous(), group=group)
    assert not torch.equal(parts13[0], parts13[1])                    # shards, not replicas
    full13, full2 = torch.cat(parts13, 0), torch.cat(parts2, 1)
    torch.manual_seed(3)                                              # the same tokens on both tensor ranks (mtp)
    x = torch.randn(24, 64, requires_grad=True)
    y = expert(x)
    y.sum().backward()
    xd = x.detach().clone().requires_grad_(True)
    yd = swiglu_interleaved(xd @ full13.t()) @ full2.t()
    yd.sum().backward()
    return bool(torch.allclose(y, yd, atol=1e-5)), bool(torch.allclose(x.grad, xd.grad, atol=1e-5))


def test_tensor_sharded_expert_equals_dense_mlp():
    for ok_y, ok_dx in run_distributed(_sharded_expert_matches_dense, 2):
        assert ok_y and ok_dx


def _train_moe_sharded_experts(rank, world, mode):
    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=2, num_experts=4, moe_type="GShard", tp=2, mode=mode)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2, expert_tensor_parallel=True)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    trainer, opt, model, _ = build_trainer(cfg)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses, norms = [], None
    for step in range(4):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=0)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        losses.append(float(out[2]))
    sharded = [p for p in model.parameters() if getattr(p, "expert_tp_sharded", False)]
    return losses, sorted(norms.items()), len(sharded)


@pytest.mark.parametrize("mode", ["mtp", "msp"])
def test_moe_trains_with_tensor_sharded_experts(mode):
    res = run_distributed(_train_moe_sharded_experts, 2, mode, timeout=600)
    (l0, n0, k0), (l1, n1, k1) = res
    assert k0 == k1 and k0 > 0
    assert l0 == l1 and l0[-1] < l0[0]          # tensor ranks agree on the loss, and it goes down
    assert n0 == n1                              # ... and on every group's gradient norm (expert shards summed over TP)


def test_binned_routing_helpers_of_the_megablock_layers():
    """``indices_and_bins`` / ``indices_and_padded_bins`` / ``load_balancing_loss`` / ``permute_and_compute`` / ``topology`` /
    ``sparse_transpose`` (reference ``megablock_moe.py:73-99,253-275``, ``megablock_dmoe.py:85-182``) against direct formulas."""
    from internevo_b200.models.moe import MOE_INITIALIZER, MegaBlockdMoE, MegaBlockMoE

    torch.manual_seed(0)
    h, E, k, S = 16, 4, 2, 12
    kw = dict(hidden_size=h, num_experts=E, ep_group=None, ep_size=1, mlp_ratio=2, device="cpu", dtype=torch.float32, top_k=k)
    mb = MOE_INITIALIZER.get_module("MegaBlock")(**kw)
    md = MOE_INITIALIZER.get_module("MegaBlock-D")(**kw)
    assert isinstance(mb, MegaBlockMoE) and isinstance(md, MegaBlockdMoE)
    scores = torch.softmax(torch.randn(S, E), -1)
    w, top = torch.topk(scores, k, dim=-1)
    indices, bin_ids, bins, padded_bins, tpe = md.indices_and_padded_bins(top)
    flat = top.reshape(-1)
    assert torch.equal(tpe.long(), torch.bincount(flat, minlength=E)) and int(bins[-1]) == S * k
    assert torch.equal(flat[indices.long()], bin_ids.long()) and bool((bin_ids[1:] >= bin_ids[:-1]).all())
    assert all(int(v) % 128 == 0 for v in padded_bins) and int(padded_bins[-1]) == 128 * int((tpe > 0).sum())
    i2, b2, bins2, tpe2 = mb.indices_and_bins(top)
    assert torch.equal(i2, indices) and torch.equal(bins2, bins) and torch.equal(tpe2, tpe) and torch.equal(b2, bin_ids)
    lb = md.load_balancing_loss(tpe, scores)
    want = E / (S * k) * sum(float(tpe[e]) * float(scores[:, e].mean()) for e in range(E))
    assert abs(float(lb) - want) < 1e-5
    assert mb.expert_capacity(S, k) == int(1.0 * k * S / E)
    # binned gather -> experts -> weighted scatter == every token through its chosen experts
    x = torch.randn(S, h)
    cap = int(tpe.max())
    out = mb.permute_and_compute(x, indices, w, bins, cap, k)
    ref = torch.zeros(S, h)
    for t in range(S):
        for j in range(k):
            ref[t] += w[t, j] * mb.experts.wrapped_experts[int(top[t, j])](x[t: t + 1])[0]
    assert torch.allclose(out, ref, atol=1e-5)
    dropped = mb.permute_and_compute(x, indices, w, bins, 1, k)      # capacity 1: only the first slot of every expert survives
    assert float((dropped - ref).detach().abs().max()) > 1e-4
    topo = md.topology(torch.zeros(int(padded_bins[-1]), h), padded_bins)
    assert topo.num_experts == E and topo.ffn == md.experts.wrapped_experts[0].w2.weight.shape[1]
    assert topo.offsets.tolist() == [0] + padded_bins.tolist()
    # block COO (2 x 3 blocks): rows [0, 0, 1], cols [0, 2, 1]
    ct, off_t, blk = md.sparse_transpose((256, 384), torch.tensor([0, 0, 1]), torch.tensor([0, 2, 1]))
    assert ct.tolist() == [0, 1, 0] and off_t.tolist() == [0, 1, 2, 3] and blk.tolist() == [0, 2, 1]


def _ep_equivalence(rank, world, state_file):
    """Dropless MoE from IDENTICAL weights: one process with 2 micro-batches / 2 expert-parallel ranks with one each."""
    import re

    from internevo_b200.core.context import global_context as gpc  # noqa: F401

    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=2 // world, num_experts=4, moe_type="MegaBlock-D")
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    cfg["hybrid_zero_optimizer"]["clip_grad_norm"] = 100.0        # no clipping: Adam's update does not see the gradient scale
    trainer, opt, model, _ = build_trainer(cfg)
    inner = model.model
    if world == 1:
        torch.save(inner.state_dict(), state_file)
    else:
        full, n_local, sd = torch.load(state_file), 4 // world, {}
        for k in inner.state_dict().keys():
            m = re.match(r"^(.*wrapped_experts\.)(\d+)(\..*)$", k)
            sd[k] = full[f"{m.group(1)}{int(m.group(2)) + rank * n_local}{m.group(3)}"] if m else full[k]
        inner.load_state_dict(sd, strict=True)
        opt.reload_zero_fp32_buff()
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses, norms = [], []
    for step in range(3):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=step)
        per = 2 // world
        data = {k: v[rank * per:(rank + 1) * per] for k, v in data.items()}
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels[rank * per:(rank + 1) * per]), forward_only=False, return_loss=True,
                                       return_output_label=False)
        ok, n = trainer.step()
        assert ok
        loss = out[2].detach().clone().reshape(1).float()
        if world > 1:
            torch.distributed.all_reduce(loss)
            loss /= world
        losses.append(float(loss))
        norms.append({("moe" if k.startswith("moe") else k): float(v) for k, v in n.items()})
    return losses, norms


def test_expert_parallel_follows_the_single_process_trajectory(tmp_path):
    """Expert parallel over 2 ranks == one process holding all experts: same loss trajectory, same dense gradient norm.  The
    EXPERT gradient norm is ``ep`` times the single-process one - reference semantics: expert gradients collect the tokens of
    every expert-parallel rank through the all-to-all and are averaged over the expert-DATA group only
    (``hybrid_zero_optim.py:166-167``, ``solver/optimizer/utils.py:120``), while each rank's loss is already a per-rank mean.
    Adam's update is invariant to that factor; the per-group clipping threshold is not (documented in ``doc/migration.md``)."""
    state = str(tmp_path / "moe_state.pt")
    single = run_distributed(_ep_equivalence, 1, state)[0]
    ep2 = run_distributed(_ep_equivalence, 2, state)
    for losses, norms in ep2:
        for a, b in zip(losses, single[0]):
            assert abs(a - b) < 2e-5 * max(1.0, abs(b)), (losses, single[0])
        for got, want in zip(norms, single[1]):
            assert abs(got["default"] - want["default"]) < 1e-4 * max(1.0, want["default"]), (got, want)
            assert abs(got["moe"] - 2 * want["moe"]) < 1e-4 * max(1.0, want["moe"]), (got, want)
