"""Checkpoint helpers (reference ``internlm/checkpoint/utils.py``)."""
from __future__ import annotations



from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)


def get_shard_state_dict(shard_model):
    return shard_model.state_dict()


def load_shard_state_dict(shard_model, shard_state, **kwargs):
    return shard_model.load_state_dict(shard_state, **kwargs)


def get_model_topology(model):
    """``{module_name: {"dim": 0}}`` for vocabulary-parallel embeddings (reference ``utils.py:52-70``)."""
    from internevo_b200.models.modules import VocabParallelEmbedding

    topos = {}
    for name, module in model.named_modules():
        if isinstance(module, VocabParallelEmbedding):
            topos[name] = {"dim": 0}
    return topos


def process_load_info(load_info):
    load_content_str = ""
    load_ckpt_folder = load_info["path"]
    load_content = load_info["content"]
    if gpc.is_rank_for_log():
        logger.info(f"Try load_ckpt_folder: {load_ckpt_folder}")
    return load_content_str, load_ckpt_folder, load_content


def try_get_tp_pp_from_fns(fns):
    max_tp = max_pp = 0
    for fn in fns:
        if fn.startswith("model_tp") and not fn.endswith(".md5"):
            segs = fn.replace(".pt", "").split("_")
            max_tp = max(max_tp, int(segs[1][2:]))
            max_pp = max(max_pp, int(segs[-1][2:]))
    return max_tp + 1, max_pp + 1


def get_non_moe_state_dict(full_state_dict):
    """Drop expert tensors (they go to per-expert files), keep the gates (reference ``checkpoint/utils.py:29-37``)."""
    for key in list(full_state_dict.keys()):
        if "expert" in key and "moe_layer.gate" not in key and ".wg." not in key:
            full_state_dict.pop(key)
    return full_state_dict
