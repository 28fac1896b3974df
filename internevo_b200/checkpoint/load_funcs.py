"""External weight loaders (reference ``internlm/checkpoint/load_funcs.py:16-192``): Meta ``llama`` consolidated shards
and HuggingFace ``hf_llama`` checkpoints into the LLaMA-2 family; sharding is applied on the fly."""
from __future__ import annotations

import os

import torch

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.models.sharding import shard_state_dict
from internevo_b200.solver.pipeline_utils import partition_uniform
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.storage_manager import get_fns, llm_load

logger = get_logger(__file__)


def _stage_layers():
    pp, ppr = gpc.get_world_size(ParallelMode.PIPELINE), gpc.get_local_rank(ParallelMode.PIPELINE)
    L = gpc.config.model.num_layers
    return partition_uniform(L, pp, gpc.config.model.get("num_chunks", 1))[ppr], L


def _finish(ckpt_mm, full, spec_prefix="layers"):
    from internevo_b200.models.sharding import pipeline_slice

    tp, tpr = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.TENSOR)
    parts, L = _stage_layers()
    model = ckpt_mm.model
    mods = list(model) if isinstance(model, torch.nn.ModuleList) else [model]
    for mod, (s, e) in zip(mods, parts):
        sd = shard_state_dict(pipeline_slice(full, s, e, first=s == 0, last=e == L), tpr, tp,
                              embed_split_hidden=gpc.config.model.get("embed_split_hidden", True))
        missing, unexpected = mod.load_state_dict(sd, strict=False)
        if gpc.is_rank_for_log():
            logger.info(f"external load: missing={list(missing)} unexpected={list(unexpected)}")
    if ckpt_mm.optimizer is not None and hasattr(ckpt_mm.optimizer, "reload_zero_fp32_buff"):
        ckpt_mm.optimizer.reload_zero_fp32_buff()


def load_llama_pretrained_weights(ckpt_mm, load_info, train_state=None):
    """Meta format: ``consolidated.{mp}.pth`` shards with ``layers.{i}.attention.{wq,wk,wv,wo}``, ``feed_forward.{w1,w2,w3}``."""
    folder = load_info["path"]
    fns = sorted(f for f in get_fns(folder) if f.endswith(".pth") or f.endswith(".pt"))
    shards = [llm_load(os.path.join(folder, f), map_location="cpu") for f in fns]
    full = {}
    col = ("wq.weight", "wk.weight", "wv.weight", "w1.weight", "w3.weight", "output.weight")
    row = ("wo.weight", "w2.weight", "tok_embeddings.weight")
    for k in shards[0]:
        if k.endswith("rope.freqs"):
            continue
        if any(k.endswith(c) for c in col):
            full[k] = torch.cat([s[k] for s in shards], 0)
        elif any(k.endswith(c) for c in row):
            full[k] = torch.cat([s[k] for s in shards], 1)
        else:
            full[k] = shards[0][k]
    _finish(ckpt_mm, full)
    return "model (llama), "


def load_hf_llama_pretrained_weights(ckpt_mm, load_info, train_state=None):
    """HF format: ``model.layers.{i}.self_attn.{q,k,v,o}_proj``, ``mlp.{gate,up,down}_proj`` (q/k row order depends on ``model.adapt_hf``)."""
    folder = load_info["path"]
    fns = sorted(f for f in get_fns(folder) if f.endswith(".bin") and f.startswith("pytorch_model"))
    hf = {}
    for f in fns:
        hf.update(llm_load(os.path.join(folder, f), map_location="cpu"))
    H = gpc.config.model.num_attention_heads
    Hkv = gpc.config.model.get("num_kv_attention_heads", H)
    h = gpc.config.model.hidden_size
    d = h // H

    hf_rope = bool(gpc.config.model.get("adapt_hf", False))

    def unpermute(w, nh):
        """HF stores q/k rows of a head as [2, d/2] (rotate-half RoPE).  With ``adapt_hf=True`` the model rotates the same
        way and the rows are taken as they are (the reference requires this setting, ``load_funcs.py:74``); with
        interleaved RoPE (``adapt_hf=False``) they go back to Meta's [d/2, 2] order."""
        if hf_rope:
            return w
        return w.view(nh, 2, d // 2, w.shape[-1]).transpose(1, 2).reshape(nh * d, w.shape[-1])

    full = {"tok_embeddings.weight": hf["model.embed_tokens.weight"], "norm.weight": hf["model.norm.weight"],
            "output.weight": hf["lm_head.weight"]}
    for i in range(gpc.config.model.num_layers):
        p, q = f"model.layers.{i}.", f"layers.{i}."
        full[q + "attention.wq.weight"] = unpermute(hf[p + "self_attn.q_proj.weight"], H)
        full[q + "attention.wk.weight"] = unpermute(hf[p + "self_attn.k_proj.weight"], Hkv)
        full[q + "attention.wv.weight"] = hf[p + "self_attn.v_proj.weight"]
        full[q + "attention.wo.weight"] = hf[p + "self_attn.o_proj.weight"]
        full[q + "feed_forward.w1.weight"] = hf[p + "mlp.gate_proj.weight"]
        full[q + "feed_forward.w3.weight"] = hf[p + "mlp.up_proj.weight"]
        full[q + "feed_forward.w2.weight"] = hf[p + "mlp.down_proj.weight"]
        full[q + "attention_norm.weight"] = hf[p + "input_layernorm.weight"]
        full[q + "ffn_norm.weight"] = hf[p + "post_attention_layernorm.weight"]
    _finish(ckpt_mm, full)
    return "model (hf_llama), "


LOAD_FUNC_DICT = {"llama": load_llama_pretrained_weights, "hf_llama": load_hf_llama_pretrained_weights}
