"""Every shipped config (the BASELINE workloads: 7B InternLM2, 7B ISP SFT, 7B MoE4, 20B TP4xPP2, LLaMA-2, SFT, demo) must load,
describe a layout that tiles 8 GPUs, and name a registered model family — checked without building any model."""
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = sorted(p for p in glob.glob(os.path.join(ROOT, "configs", "*.py")) if not p.endswith("__init__.py"))


def _size(v):
    return v if isinstance(v, int) else v.get("size", 1)


@pytest.mark.parametrize("path", CONFIGS, ids=[os.path.basename(p) for p in CONFIGS])
def test_config_loads_and_tiles_8_gpus(path):
    import internevo_b200.models  # noqa: F401  (registers the model builders)
    from internevo_b200.core.context import Config
    from internevo_b200.core.context.process_groups import group_rank_lists
    from internevo_b200.utils.registry import MODEL_INITIALIZER

    cfg = Config.from_file(path)
    for key in ("model", "data", "parallel", "adam", "lr_scheduler", "hybrid_zero_optimizer", "grad_scaler", "loss"):
        assert key in cfg, f"{os.path.basename(path)}: missing section {key}"
    assert MODEL_INITIALIZER.has(cfg.get("model_type", "INTERNLM")), cfg.get("model_type")
    m, d, par = cfg.model, cfg.data, cfg.parallel
    assert m["hidden_size"] % m["num_attention_heads"] == 0
    assert m["num_attention_heads"] % m.get("num_kv_attention_heads", m["num_attention_heads"]) == 0
    assert d["seq_len"] > 0 and d["micro_bsz"] > 0 and d["micro_num"] > 0
    tp, pp, wp = _size(par["tensor"]), _size(par["pipeline"]), _size(par.get("weight", 1))
    world = 8 if os.path.basename(path) != "demo.py" else 2
    assert world % (tp * pp) == 0, (tp, pp)
    mode = par["tensor"].get("mode", "mtp") if not isinstance(par["tensor"], int) else "mtp"
    if mode == "isp":
        assert world % wp == 0 and m["num_attention_heads"] % tp == 0           # Ulysses: heads split over the sequence group
    else:
        assert wp == 1 and m["num_attention_heads"] % tp == 0 and m.get("num_kv_attention_heads", tp) % tp == 0
    assert m["num_layers"] % (pp * m.get("num_chunks", 1)) == 0 or pp == 1
    if "num_experts" in m:
        assert world % 1 == 0 and m["num_experts"] >= 1
    # the rank layout of every mode the config needs covers each rank exactly once
    from internevo_b200.core.context import ParallelMode, ParallelSizes, layout_for_rank, modes_to_build

    z = _size(par["zero1"])
    sizes = ParallelSizes(world=world, pipeline=pp, tensor=tp, weight=wp, zero1=z if z > 0 else -1,
                          num_experts=m.get("num_experts", 1), isp=mode == "isp")
    for pm in modes_to_build(sizes):
        if pm in (ParallelMode.NETTEST,):
            continue
        groups = group_rank_lists(pm, sizes)
        flat = sorted(r for g in groups for r in g)
        assert flat == list(range(world)), (pm, groups)
    assert layout_for_rank(0, sizes)[ParallelMode.TENSOR] is not None
