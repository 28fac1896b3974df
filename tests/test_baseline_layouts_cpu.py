"""The parallel layouts BASELINE.json names, at their REAL rank counts (8 gloo ranks, tiny model): each must reproduce the
single-process trajectory from identical weights and data.

    configs/7B_internlm2.py   tensor 2 (mtp) x data 4, Hybrid-ZeRO over the 4 data ranks (and the ZeRO-1.5 sub-group form)
    configs/7B_isp_sft.py     ISP: sequence parallel 8 x weight parallel 8
    configs/20B_internlm2.py  tensor 4 x pipeline 2 (1F1B)
    configs/7B_MoE4_sft.py    4 experts: expert parallel 4 x expert-data parallel 2

The 2- / 4-rank forms of the same mechanisms live in ``test_parallel_cpu.py`` / ``test_parallel_combos_cpu.py``; the reference
checks these layouts with its 8-GPU loss-curve marks (``tests/test_training/test_loss.py:257-400``)."""
import pytest

import test_parallel_cpu as T
from common import run_distributed
from test_parallel_combos_cpu import _check_union

WORLD = 8
# model shapes wide enough for the group sizes (8 query heads for sp = 8; 4 kv heads for tp = 4)
SHAPES = {
    "h4kv2": dict(),
    "h8kv8": dict(heads=8, kv_heads=8),
    "h8kv4": dict(heads=8, kv_heads=4),
}
LAYOUTS = {
    "7B_internlm2_tp2_dp4_zero4": ("h4kv2", dict(tp=2, micro_num=1, zero1=-1)),
    "7B_internlm2_tp2_dp4_zero2_subgroups": ("h4kv2", dict(tp=2, micro_num=1, zero1=2)),
    "7B_internlm2_tp2_dp4_zero4_overlapped": ("h4kv2", dict(tp=2, micro_num=1, zero1=-1, overlap=4096)),
    "7B_isp_sft_sp8_wp8": ("h8kv8", dict(tp=8, wp=8, mode="isp", micro_num=4)),
    "20B_internlm2_tp4_pp2": ("h8kv4", dict(tp=4, pp=2, micro_num=4)),
    "20B_internlm2_tp4_pp2_msp_interleaved": ("h8kv4", dict(tp=4, pp=2, mode="msp", micro_num=4, num_chunks=2)),
    # the reference's three-way marks (``16GPU_4DP2TP2PP_{MTP,MSP,FSP}``) with data parallel 2: DP x TP x PP + Hybrid-ZeRO
    "dp2_tp2_pp2_mtp": ("h4kv2", dict(tp=2, pp=2, mode="mtp", micro_num=2)),
    "dp2_tp2_pp2_msp": ("h4kv2", dict(tp=2, pp=2, mode="msp", micro_num=2)),
    "dp2_tp2_pp2_fsp_interleaved": ("h4kv2", dict(tp=2, pp=2, mode="fsp", micro_num=2, num_chunks=2)),
    "dp2_isp_sp2_wp4_pp2": ("h4kv2", dict(tp=2, wp=4, pp=2, mode="isp", micro_num=2)),
}
_baselines = {}


def _baseline(shape, isp=False):
    """Single-process trajectory.  ISP layouts are compared with a single-process run in ISP mode: ISP keeps the embedding and the
    head in a parameter group of their own and gradient clipping is per GROUP (reference ``hybrid_zero_optim.py:863-876``), so from
    the second update on its trajectory differs (deterministically, ~2e-4) from the one-group trajectory of the other modes."""
    if (shape, isp) not in _baselines:
        kw = dict(micro_num=T.MICRO_TOTAL, **SHAPES[shape])
        if isp:
            kw["mode"] = "isp"
        _baselines[shape, isp] = run_distributed(T._train, 1, kw)[0]
    return _baselines[shape, isp]


@pytest.mark.parametrize("name", list(LAYOUTS))
def test_baseline_layout_at_8_ranks_matches_single_process(name):
    shape, kw = LAYOUTS[name]
    isp = kw.get("mode") == "isp"
    res = run_distributed(T._train, WORLD, dict(kw, **SHAPES[shape]), timeout=900)
    # ISP included: every sequence shard contributes its share of the mean over ALL valid tokens of the micro-batch (as the
    # reference, which gathers the sequence in front of the head), so 8 shards with unequal numbers of ignored labels follow the
    # single-process trajectory like every other layout
    _check_union(res, _baseline(shape, isp), 2e-4)


def _moe_ep4_edp2(rank, world):
    """7B_MoE4 layout: the 4 experts live on 4 expert-parallel ranks, replicated twice (expert-data parallel 2); dense
    parameters are data parallel over all 8 ranks."""
    import torch

    from common import build_trainer, synthetic_batch, tiny_config
    from internevo_b200.core.context import ParallelMode, global_context as gpc

    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=1, num_experts=4)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2, capacity_factor=2.0, eval_capacity_factor=2.0, min_capacity=4, noisy_gate_policy=None,
                      drop_tokens=True, use_rts=False)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    trainer, opt, model, _ = build_trainer(cfg)
    assert gpc.get_world_size(ParallelMode.EXPERT) == 4 and gpc.get_world_size(ParallelMode.EXPERT_DATA) == 2
    assert gpc.get_world_size(ParallelMode.DATA) == 8
    T_ = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    losses = []
    for _ in range(4):
        data, labels = synthetic_batch(1, T_, cfg["model"]["vocab_size"], seed=rank)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        losses.append(float(out[2]))
    # replicas stay identical: dense parameters over all ranks, each expert over its expert-data pair
    dense, expert = [], []
    for n, p in model.named_parameters():
        (expert if ".experts." in n or "wrapped_experts" in n else dense).append(p.detach().double().sum())
    d = torch.stack(dense).sum().reshape(1)
    e = torch.stack(expert).sum().reshape(1) if expert else torch.zeros(1, dtype=torch.float64)
    dl = [torch.zeros_like(d) for _ in range(world)]
    torch.distributed.all_gather(dl, d)
    el = [torch.zeros_like(e) for _ in range(2)]
    torch.distributed.all_gather(el, e, group=gpc.get_group(ParallelMode.EXPERT_DATA))
    return losses, sorted(norms), [float(x) for x in dl], [float(x) for x in el], len(expert)


def test_moe4_expert_parallel_4_expert_data_2():
    res = run_distributed(_moe_ep4_edp2, WORLD, timeout=900)
    for losses, groups, dense, expert, n_expert in res:
        assert losses[-1] < losses[0], losses
        assert "moe_ep_size_4" in groups, groups
        assert n_expert > 0
        assert max(dense) - min(dense) < 1e-9 * max(1.0, abs(dense[0])), dense
        assert abs(expert[0] - expert[1]) < 1e-9 * max(1.0, abs(expert[0])), expert
