"""Checkpoint manager: file layout, completion marker, snapshot alternation, stop-file protocol and bit-exact auto-resume
(reference strategy: tests/test_utils/test_model_checkpoint.py — save, 'crash', resume, compare)."""
import os

import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config


def _step(trainer, cfg, dpr, seed):
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=seed * 10 + dpr)
    trainer.zero_grad()
    out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
    ok, norms = trainer.step()
    assert ok
    return float(out[2])


def _run(rank, world, folder, phase):
    from internevo_b200.checkpoint import CheckpointManager
    from internevo_b200.core.context import ParallelMode, global_context as gpc
    from internevo_b200.core.trainer import TrainState

    cfg = tiny_config(zero1=world, num_layers=2, micro_num=2)
    cfg["ckpt"] = dict(enable_save_ckpt=True, save_ckpt_folder=f"local:{folder}", checkpoint_every=2, oss_snapshot_freq=3,
                       auto_resume=(phase == "resume"), async_upload=False, stop_file_path=os.path.join(folder, "stop"))
    trainer, opt, model, _ = build_trainer(cfg)
    dpr = gpc.get_local_rank(ParallelMode.DATA)
    ts = TrainState(gpc.config, None)
    mm = CheckpointManager(ckpt_config=gpc.config.ckpt, model=model, optimizer=opt, lr_scheduler=trainer.engine._lr_scheduler,
                           model_config=gpc.config.model)
    mm.try_resume_training(ts)
    losses = {}
    if phase == "first":
        assert ts.step_count == 0
        for step in range(1, 5):  # same bookkeeping as train.py: 0-based batch_count, step_count = completed steps
            ts.batch_count = step - 1
            losses[step] = _step(trainer, cfg, dpr, step)
            ts.step_count += 1
            stop = mm.try_save_checkpoint(ts)
            assert not stop
        mm.wait_async_upload_finish()
        if rank == 0:
            for s in (2, 4):
                names = set(os.listdir(os.path.join(folder, str(s))))
                assert {"model_tp0_pp0.pt", "context.pt", "schedulder.pt", "model_config.pt", f"{s}.step"} <= names, names
                assert {f"optimizer_tp0_pp0_zo{z}.pt" for z in range(world)} <= names, names
            assert "3.step" in os.listdir(os.path.join(folder, "snapshot", "0"))   # step 3 snapshot
        # continue two more steps without saving: the reference trajectory for the resumed run
        for step in (5, 6):
            losses[step] = _step(trainer, cfg, dpr, step)
    else:
        assert ts.step_count == 4 and ts.batch_count == 4, (ts.step_count, ts.batch_count)  # newest complete ckpt: step 4
        for step in (5, 6):
            losses[step] = _step(trainer, cfg, dpr, step)
        # stop file: "save at step 6 and quit"
        if rank == 0:
            open(os.path.join(folder, "stop"), "w").write("6")
        torch.distributed.barrier()
        ts.step_count, ts.batch_count = 6, 5
        mm.checkpoint_every = 1000
        mm.oss_snapshot_freq = 1000
        assert mm.try_save_checkpoint(ts) is True
        if rank == 0:
            assert os.path.exists(os.path.join(folder, "6", "6.step"))
            assert open(os.path.join(folder, "stop")).read().strip() == "0"
        # a second request in the same run is honoured too: "-8" = save at step 8 and continue; nothing happens at step 7
        torch.distributed.barrier()
        if rank == 0:
            open(os.path.join(folder, "stop"), "w").write("-8")
        torch.distributed.barrier()
        ts.step_count, ts.batch_count = 7, 6
        assert mm.try_save_checkpoint(ts) is False
        torch.distributed.barrier()
        assert not os.path.exists(os.path.join(folder, "7"))
        ts.step_count, ts.batch_count = 8, 7
        assert mm.try_save_checkpoint(ts) is False
        mm.wait_async_upload_finish()
        if rank == 0:
            assert os.path.exists(os.path.join(folder, "8", "8.step"))
            assert open(os.path.join(folder, "stop")).read().strip() == "0"
    return losses


def test_save_resume_is_exact(tmp_path):
    first = run_distributed(_run, 2, str(tmp_path), "first")
    resumed = run_distributed(_run, 2, str(tmp_path), "resume")
    for r in range(2):
        for step in (5, 6):
            assert abs(first[r][step] - resumed[r][step]) < 1e-6, (first[r], resumed[r])


def _run_layout(rank, world, folder, phase, kw, moe):
    """Train 2 steps, save, train 2 more (reference trajectory) / resume from the checkpoint and train the same 2 steps."""
    from internevo_b200.checkpoint import CheckpointManager
    from internevo_b200.core.context import ParallelMode, global_context as gpc
    from internevo_b200.core.trainer import TrainState

    kw = dict(kw)
    bucket = kw.pop("overlap_bucket", None)
    ckpt_extra = kw.pop("ckpt_extra", {})
    top_level = kw.pop("top_level", {})
    cfg = tiny_config(num_layers=2, micro_num=2, **kw)
    cfg.update(top_level)
    if bucket:   # Hybrid-ZeRO with many small ranges reduced from the grad hooks: optimizer shards are range-interleaved
        cfg["hybrid_zero_optimizer"].update(overlap_sync_grad=True, reduce_bucket_size=bucket)
    if moe:
        cfg["model"].pop("no_bias", None)
        cfg["model"].pop("num_kv_attention_heads", None)
        cfg["moe"] = dict(top_k=2)
        cfg["loss"]["moe_loss_coeff"] = 0.1
    cfg["ckpt"] = dict(enable_save_ckpt=True, save_ckpt_folder=f"local:{folder}", checkpoint_every=2, oss_snapshot_freq=0,
                       auto_resume=(phase == "resume"), async_upload=False)
    if phase == "first":
        cfg["ckpt"].update(ckpt_extra)
    trainer, opt, model, _ = build_trainer(cfg)
    dpr = gpc.get_local_rank(ParallelMode.DATA)
    ts = TrainState(gpc.config, None)
    mm = CheckpointManager(ckpt_config=gpc.config.ckpt, model=model, optimizer=opt, lr_scheduler=trainer.engine._lr_scheduler,
                           model_config=gpc.config.model)
    mm.try_resume_training(ts)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]

    def step(seed):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=seed * 10 + dpr)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        return (float(out[2]) if out[2] is not None else None, sorted(norms.items()))

    res = {}
    if phase == "first":
        for s in (1, 2):
            ts.batch_count = s - 1
            step(s)
            ts.step_count += 1
            mm.try_save_checkpoint(ts)
        mm.wait_async_upload_finish()
    else:
        assert ts.step_count == 2
    for s in (3, 4):
        res[s] = step(s)
    return res


import pytest  # noqa: E402


@pytest.mark.parametrize("name,world,kw,moe", [
    ("tp2_pp2", 4, dict(tp=2, pp=2), False),
    ("isp_sp2_wp2", 2, dict(tp=2, wp=2, mode="isp"), False),
    # weight group wider than the sequence group: the head's vocabulary rows are re-sharded between file and memory layout
    ("isp_sp2_wp4", 4, dict(tp=2, wp=4, mode="isp"), False),
    ("moe_ep2", 2, dict(model_type="INTERNLM_MoE", num_experts=4, moe_type="MegaBlock-D"), True),  # dropless: no gate noise
    # range-interleaved optimizer shards (many ranges, reduction overlapped with backward) must resume bit-exactly, too
    ("dp2_zero2_overlap", 2, dict(zero1=2, overlap_bucket=4096), False),
    # MoE x pipeline: every stage numbers its blocks from 0 - the per-expert files must carry GLOBAL layer ids
    ("moe_pp2", 2, dict(pp=2, model_type="INTERNLM_MoE", num_experts=4, moe_type="MegaBlock-D"), True),
])
def test_resume_is_exact_for_model_parallel_and_moe_layouts(tmp_path, name, world, kw, moe):
    first = run_distributed(_run_layout, world, str(tmp_path), "first", kw, moe)
    resumed = run_distributed(_run_layout, world, str(tmp_path), "resume", kw, moe)
    for r in range(world):
        for s in (3, 4):
            (l0, n0), (l1, n1) = first[r][s], resumed[r][s]
            assert (l0 is None) == (l1 is None)
            if l0 is not None:
                assert abs(l0 - l1) < 1e-6, (name, first[r], resumed[r])
            for (k0, v0), (k1, v1) in zip(n0, n1):
                assert k0 == k1 and abs(v0 - v1) < 1e-5 * max(1.0, abs(v0)), (name, n0, n1)
    if name.startswith("isp"):
        # files carry the reference's ISP layout: embedding split along hidden, head rows split over the TENSOR group
        cfgm = tiny_config(num_layers=2, **{k: v for k, v in kw.items() if k != "overlap_bucket"})["model"]
        V, h, tp, wp = cfgm["vocab_size"], cfgm["hidden_size"], kw["tp"], kw["wp"]
        for w in range(wp):
            t = w % tp
            st = torch.load(os.path.join(str(tmp_path), "2", f"model_tp{t}_wp{w}_pp0.pt"), weights_only=False)
            emb = next(v for k, v in st.items() if k.endswith("tok_embeddings.weight"))
            head = next(v for k, v in st.items() if k.endswith("output.weight"))
            wqkv = next(v for k, v in st.items() if k.endswith("layers.0.attention.wqkv.weight"))
            assert tuple(emb.shape) == (V, h // tp) and tuple(head.shape) == (V // tp, h), (emb.shape, head.shape)
            assert wqkv.shape[0] * wp == (cfgm["num_attention_heads"] + 2 * cfgm["num_kv_attention_heads"]) * (
                h // cfgm["num_attention_heads"])
        # the tensor ranks' embedding slices are different columns of ONE matrix, the head slices different rows
        e0 = torch.load(os.path.join(str(tmp_path), "2", "model_tp0_wp0_pp0.pt"), weights_only=False)
        e1 = torch.load(os.path.join(str(tmp_path), "2", "model_tp1_wp1_pp0.pt"), weights_only=False)
        k_emb = next(k for k in e0 if k.endswith("tok_embeddings.weight"))
        assert not torch.equal(e0[k_emb], e1[k_emb])
    if moe:
        import glob
        import re as _re

        files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(str(tmp_path), "2", "model_moe_layer*")))
        layers = {int(_re.match(r"model_moe_layer(\d+)_", f).group(1)) for f in files}
        assert layers == {0, 1} and len(files) == 2 * 4, files    # 2 global layers x 4 global experts, no collisions
        st = torch.load(os.path.join(str(tmp_path), "2", "model_moe_layer1_expert3_tp0.pt"), weights_only=False)
        assert st and all(".wrapped_experts.3." in k for k in st), list(st)[:3]   # keys carry the global expert id


def _run_partial(rank, world, folder, phase):
    """Load ``content=("model", "scheduler")`` from a finished run into a FRESH optimizer: the fp32 master must follow the
    loaded weights (or the first step writes the init weights' update back) and the learning rate comes from the checkpoint."""
    from internevo_b200.checkpoint import CheckpointManager
    from internevo_b200.core.context import ParallelMode, global_context as gpc
    from internevo_b200.core.trainer import TrainState

    cfg = tiny_config(zero1=world, num_layers=2, micro_num=2)
    cfg["ckpt"] = dict(enable_save_ckpt=(phase == "first"), save_ckpt_folder=f"local:{folder}", checkpoint_every=2,
                       oss_snapshot_freq=0, auto_resume=False, async_upload=False)
    if phase != "first":
        cfg["ckpt"]["load_ckpt_info"] = dict(path=f"local:{folder}/2", content=("model", "scheduler"), ckpt_type="internevo")
    trainer, opt, model, _ = build_trainer(cfg, seed=1024 if phase == "first" else 77)    # different init on reload
    dpr = gpc.get_local_rank(ParallelMode.DATA)
    ts = TrainState(gpc.config, None)
    mm = CheckpointManager(ckpt_config=gpc.config.ckpt, model=model, optimizer=opt, lr_scheduler=trainer.engine._lr_scheduler,
                           model_config=gpc.config.model)
    mm.try_resume_training(ts)
    if phase == "first":
        for s in (1, 2):
            ts.batch_count = s - 1
            _step(trainer, cfg, dpr, s)
            ts.step_count += 1
            mm.try_save_checkpoint(ts)
        mm.wait_async_upload_finish()
        g = next(g for g in opt.groups if g.params)
        return {"w": g.param_arena.clone(), "lr": g.cfg["lr"]}
    g = next(g for g in opt.groups if g.params)
    loaded = g.param_arena.clone()
    # master == loaded weights on the owned sub-slices
    for v, m, n in g.owned_views(g.param_arena):
        assert torch.equal(g.master[m: m + n], v.float())
    lr = g.cfg["lr"]
    _step(trainer, cfg, dpr, 3)
    moved = float((g.param_arena - loaded).abs().max())
    return {"w": loaded, "lr": lr, "moved": moved, "step": ts.step_count}


def test_partial_load_keeps_loaded_weights_and_learning_rate(tmp_path):
    first = run_distributed(_run_partial, 2, str(tmp_path), "first")
    again = run_distributed(_run_partial, 2, str(tmp_path), "reload")
    for r in range(2):
        assert torch.equal(first[r]["w"], again[r]["w"])                     # the checkpoint's weights, not the new init
        assert abs(first[r]["lr"] - again[r]["lr"]) < 1e-12, (first[r]["lr"], again[r]["lr"])
        assert again[r]["step"] == 2
        # one AdamW step from the loaded weights moves them by ~lr, not by the distance to another initialisation
        assert 0 < again[r]["moved"] < 0.02, again[r]["moved"]
