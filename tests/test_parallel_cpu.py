"""Hybrid-parallel equivalence on CPU/gloo: every layout (DP+ZeRO, TP in mtp/msp/fsp, 1F1B and interleaved pipeline)
must reproduce the single-process loss trajectory from identical weights and data — the CPU analogue of the
reference's ``tests/test_training/test_loss.py`` layout matrix and ``tests/test_core/test_pipeline.py``."""
import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config

STEPS = 3
MICRO_TOTAL = 4  # micro-batches per step (split over dp ranks)


def _golden_state(cfg, seed=7):
    """Full (unsharded) InternLM2 weights generated from the config's shapes."""
    m = cfg["model"]
    h, H, Hkv, V, L = m["hidden_size"], m["num_attention_heads"], m["num_kv_attention_heads"], m["vocab_size"], m["num_layers"]
    d = h // H
    F = 256 * ((int(h * m["mlp_ratio"]) + 255) // 256)
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g) * 0.05  # noqa: E731
    sd = {"tok_embeddings.weight": r(V, h), "norm.weight": 1 + r(h), "output.weight": r(V, h)}
    for i in range(L):
        p = f"layers.{i}."
        sd[p + "attention.wqkv.weight"] = r((H + 2 * Hkv) * d, h)
        sd[p + "attention.wo.weight"] = r(h, h)
        sd[p + "attention_norm.weight"] = 1 + r(h)
        sd[p + "ffn_norm.weight"] = 1 + r(h)
        sd[p + "feed_forward.w1.weight"] = r(F, h)
        sd[p + "feed_forward.w3.weight"] = r(F, h)
        sd[p + "feed_forward.w2.weight"] = r(h, F)
    return sd


def _load_golden(model, opt, cfg):
    from internevo_b200.core.context import ParallelMode, global_context as gpc
    from internevo_b200.models.sharding import pipeline_slice, shard_state_dict, shard_state_dict_isp
    from internevo_b200.solver.pipeline_utils import partition_uniform

    full = _golden_state(cfg)
    tp, tpr = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.TENSOR)
    pp, ppr = gpc.get_world_size(ParallelMode.PIPELINE), gpc.get_local_rank(ParallelMode.PIPELINE)
    L, chunks = cfg["model"]["num_layers"], cfg["model"].get("num_chunks", 1)
    parts = partition_uniform(L, pp, chunks)[ppr]
    inner = model.model
    mods = list(inner) if isinstance(inner, torch.nn.ModuleList) else [inner]
    for mod, (s, e) in zip(mods, parts):
        sd = pipeline_slice(full, s, e, first=s == 0, last=e == L)
        if cfg["parallel"]["tensor"]["mode"] == "isp":
            sd = shard_state_dict_isp(sd, gpc.get_local_rank(ParallelMode.WEIGHT), gpc.get_world_size(ParallelMode.WEIGHT))
        else:
            sd = shard_state_dict(sd, tpr, tp)
        missing, unexpected = mod.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
    opt.reload_zero_fp32_buff()


def _train(rank, world, kw):
    from internevo_b200.core.context import ParallelMode, global_context as gpc

    kw = dict(kw)
    overlap = kw.pop("overlap", None)
    cfg = tiny_config(**kw)
    if overlap is not None:   # gradient reduction launched range by range from the backward hooks
        cfg["hybrid_zero_optimizer"].update(overlap_sync_grad=True, reduce_bucket_size=int(overlap))
    trainer, opt, model, _ = build_trainer(cfg)
    _load_golden(model, opt, cfg)
    dp, dpr = gpc.get_world_size(ParallelMode.DATA), gpc.get_local_rank(ParallelMode.DATA)
    per = MICRO_TOTAL // dp
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses = []
    for step in range(STEPS):
        data, labels = synthetic_batch(MICRO_TOTAL, T, cfg["model"]["vocab_size"], seed=0)
        data = {k: v[dpr * per:(dpr + 1) * per] for k, v in data.items()}
        labels = labels[dpr * per:(dpr + 1) * per]
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        loss = out[2]
        if loss is not None:
            loss = loss.detach().clone().reshape(1).float()
            if dp > 1:
                torch.distributed.all_reduce(loss, group=gpc.get_group(ParallelMode.DATA))
                loss /= dp
            if cfg["parallel"]["tensor"]["mode"] == "isp" and cfg["parallel"]["tensor"]["size"] > 1:
                # every sequence shard reports the mean over its own tokens
                torch.distributed.all_reduce(loss, group=gpc.get_group(ParallelMode.TENSOR))
                loss /= cfg["parallel"]["tensor"]["size"]
            losses.append(float(loss))
        else:
            losses.append(None)
    if overlap is not None:
        g = max(opt.groups, key=lambda g: g.total)
        assert len(g.ranges) >= 4, g.ranges
        if gpc.get_world_size(ParallelMode.PIPELINE) == 1:   # every sharded range goes out while the backward is running
            assert opt.overlap_stats["hook_launches"] >= STEPS * g.n_sharded_ranges, (opt.overlap_stats, g.ranges)
    return losses, norms


@pytest.fixture(scope="module")
def baseline():
    return run_distributed(_train, 1, dict(micro_num=MICRO_TOTAL))[0]


def _check(res, baseline, tol=2e-4):
    ref_losses, ref_norms = baseline
    got = [r for r in res if r[0][0] is not None]
    assert got, "no rank reported a loss"
    for losses, norms in got:
        for a, b in zip(losses, ref_losses):
            assert abs(a - b) < tol * max(1.0, abs(b)), (losses, ref_losses)
        total = sum(v * v for v in norms.values()) ** 0.5        # ISP reports the embedding in its own group
        ref_total = sum(v * v for v in ref_norms.values()) ** 0.5
        assert abs(total - ref_total) < 1e-3 * max(1.0, ref_total), (norms, ref_norms)


def test_single_process_trains(baseline):
    losses, _ = baseline
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("zero1", [-1, 1])
def test_dp2_zero(baseline, zero1):
    _check(run_distributed(_train, 2, dict(micro_num=MICRO_TOTAL // 2, zero1=zero1)), baseline)


@pytest.mark.parametrize("mode", ["mtp", "msp", "fsp"])
def test_tp2(baseline, mode):
    _check(run_distributed(_train, 2, dict(tp=2, mode=mode, micro_num=MICRO_TOTAL)), baseline)


def test_pp2_1f1b(baseline):
    _check(run_distributed(_train, 2, dict(pp=2, micro_num=MICRO_TOTAL)), baseline)


def test_pp2_interleaved(baseline):
    _check(run_distributed(_train, 2, dict(pp=2, micro_num=MICRO_TOTAL, num_chunks=2)), baseline)


@pytest.mark.parametrize("zero1", [-1, 1])
def test_dp2_zero_overlapped_with_backward(baseline, zero1):
    """``overlap_sync_grad``: ranges of 8 Ki elements reduced from the grad hooks while the backward is still running must
    give the serial trajectory (range-interleaved ownership, fixed launch order)."""
    _check(run_distributed(_train, 2, dict(micro_num=MICRO_TOTAL // 2, zero1=zero1, overlap=8192)), baseline)


def test_tp2_dp2_zero2_overlapped(baseline):
    _check(run_distributed(_train, 4, dict(tp=2, mode="msp", micro_num=MICRO_TOTAL // 2, zero1=2, overlap=4096)), baseline)


def test_pp2_dp2_overlapped(baseline):
    _check(run_distributed(_train, 4, dict(pp=2, micro_num=MICRO_TOTAL // 2, overlap=4096)), baseline)


def test_tp2_dp2_zero2(baseline):
    _check(run_distributed(_train, 4, dict(tp=2, mode="fsp", micro_num=MICRO_TOTAL // 2, zero1=2)), baseline)


def test_activation_checkpoint_matches(baseline):
    _check(run_distributed(_train, 1, dict(micro_num=MICRO_TOTAL, checkpoint=True)), baseline, tol=1e-5)


def test_isp_tp2_wp2(baseline):
    """Intern sequence parallel: activations sequence-sharded over TENSOR, weights sharded over WEIGHT (all-gather on
    use, reduce-scatter of wgrads), Ulysses all-to-all around attention."""
    _check(run_distributed(_train, 2, dict(tp=2, mode="isp", wp=2, micro_num=MICRO_TOTAL)), baseline, tol=5e-4)


def test_pp4_1f1b(baseline):
    _check(run_distributed(_train, 4, dict(pp=4, micro_num=MICRO_TOTAL)), baseline)


def test_pp4_interleaved_8layers():
    """4 stages x 2 virtual chunks, 8 micro-batches: the steady-state 1F1B phase crosses chunk boundaries."""
    import test_parallel_cpu as me

    old = me.MICRO_TOTAL
    me.MICRO_TOTAL = 8
    try:
        kw = dict(micro_num=8, num_layers=8)
        base = run_distributed(_train8, 1, kw)[0]
        _check(run_distributed(_train8, 4, dict(pp=4, num_chunks=2, **kw)), base)
    finally:
        me.MICRO_TOTAL = old


def _train8(rank, world, kw):
    import test_parallel_cpu as me

    me.MICRO_TOTAL = 8
    return _train(rank, world, kw)


def _train_swapped(rank, world, kw):
    """The SAME token stream cut into ``micro_num`` micro-batches of ``micro_bsz`` packed sequences each."""
    from internevo_b200.core.context import global_context as gpc  # noqa: F401  (initialised by build_trainer)

    micro_num, micro_bsz, seq_len, total_rows = kw["micro_num"], kw["micro_bsz"], 32, 4
    cfg = tiny_config(micro_num=micro_num, micro_bsz=micro_bsz, seq_len=seq_len)
    trainer, opt, model, _ = build_trainer(cfg)
    _load_golden(model, opt, cfg)
    losses, norms = [], None
    for step in range(STEPS):
        # four sequences of seq_len tokens; a micro-batch packs micro_bsz of them into one row
        data, labels = synthetic_batch(total_rows, seq_len, cfg["model"]["vocab_size"], seed=step, segments=1)
        rows = total_rows // micro_bsz
        ids = data["input_ids"].reshape(rows, micro_bsz * seq_len)
        cu = torch.arange(0, micro_bsz * seq_len + 1, seq_len, dtype=torch.int32).repeat(rows, 1)
        idx = torch.arange(seq_len).repeat(rows, micro_bsz)
        assert rows == micro_num
        trainer.zero_grad()
        out = trainer.execute_schedule(({"input_ids": ids, "cu_seqlens": cu, "indexes": idx},
                                        labels.reshape(rows, micro_bsz * seq_len)),
                                       forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        losses.append(float(out[2]))
    return losses, norms


def test_swapping_micro_num_and_micro_bsz_keeps_loss_and_grad_norm():
    """Counterpart of the reference's ``tests/test_training/test_swap_nb_loss_and_gradnorm.py``: 4 micro-batches of one sequence,
    2 of two and 1 of four see the same tokens with the same per-token weights, so loss and gradient norm agree step by step."""
    runs = {mn: run_distributed(_train_swapped, 1, dict(micro_num=mn, micro_bsz=4 // mn))[0] for mn in (4, 2, 1)}
    ref_losses, ref_norms = runs[4]
    for mn in (2, 1):
        losses, norms = runs[mn]
        for a, b in zip(losses, ref_losses):
            assert abs(a - b) < 2e-5 * max(1.0, abs(b)), (mn, losses, ref_losses)
        for k, v in ref_norms.items():
            assert abs(norms[k] - v) < 1e-4 * max(1.0, v), (mn, norms, ref_norms)
