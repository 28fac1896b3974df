"""Model registry entries — names identical to the reference so configs port unchanged:
``INTERNLM`` (``modeling_internlm.py:463``), ``INTERNLM2_PUBLIC`` (``modeling_internlm2.py:1055``), ``LLAMA2``
(``modeling_llama.py:1023``), ``INTERNLM_MoE`` (``modeling_moe.py:487``)."""
from __future__ import annotations

import torch

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.registry import MODEL_INITIALIZER

from .decoder import INTERNLM2_SPEC, INTERNLM_SPEC, LLAMA_SPEC, build_generic_model_1d

_COMMON = dict(
    checkpoint=0.0, dtype=torch.float, embed_split_hidden=False, hidden_size=2048, vocab_size=50304, embed_grad_scale=1,
    parallel_output=True, num_attention_heads=32, max_position_embeddings=2048, mlp_ratio=4.0, residual_in_fp32=False,
    use_dynamic_ntk_rope=False, norm_type="rmsnorm", drop_rate=0, attn_drop_rate=0, apply_post_layer_norm=False,
    layer_norm_epsilon=1e-5, is_reward=False, dropout_selective_checkpoint=True, use_scaled_init=True, use_swiglu=True,
    use_flash_attn=True, rope_base=10000,
)
_V2_EXTRA = dict(
    num_kv_attention_heads=None, no_bias=True, embedding_init_std=0.02, attn_wqkv_init_std=0.02,
    attn_other_init_std=0.02, ffn_uplayer_init_std=0.02, ffn_other_init_std=0.02, out_head_init_std=0.02,
    init_type="normal", norm_head=False,
    # adapt_hf=True: rotate-half RoPE (the HF convention, no row permutation on conversion); False: interleaved pairs.
    # Defaults are the reference's: InternLM2 True (modeling_internlm2.py:1071), LLaMA-2 False (modeling_llama.py:1039).
    adapt_hf=True,
)


def _cfg(defaults, kwargs):
    cfg = dict(defaults)
    unknown = set(kwargs) - set(cfg) - {"num_layers", "num_chunks", "device", "deepnorm", "max_position_embeddings"}
    for k in unknown:  # tolerated-but-ignored reference knobs (e.g. sequence_parallel, deepnorm)
        kwargs.pop(k)
    cfg.update(kwargs)
    if isinstance(cfg.get("dtype"), str):
        cfg["dtype"] = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16,
                        "torch.float32": torch.float32, "torch.tf32": torch.float32}[cfg["dtype"]]
    return cfg


@MODEL_INITIALIZER.register_module("INTERNLM")
def build_model_with_cfg(num_chunks=1, num_layers=48, **kwargs):
    """InternLM (v1): MHA, fused ``Wqkv`` with bias."""
    cfg = _cfg(_COMMON, kwargs)
    cfg["num_kv_attention_heads"] = cfg["num_attention_heads"]
    return build_generic_model_1d(INTERNLM_SPEC, num_layers=num_layers, num_chunks=num_chunks, **cfg)


@MODEL_INITIALIZER.register_module("INTERNLM2_PUBLIC")
def build_model_with_cfg_internlm2(num_chunks=1, num_layers=48, **kwargs):
    """InternLM2: GQA, interleaved ``wqkv``."""
    cfg = _cfg({**_COMMON, **_V2_EXTRA}, kwargs)
    cfg["num_kv_attention_heads"] = cfg["num_kv_attention_heads"] or cfg["num_attention_heads"]
    return build_generic_model_1d(INTERNLM2_SPEC, num_layers=num_layers, num_chunks=num_chunks, **cfg)


@MODEL_INITIALIZER.register_module("LLAMA2")
def build_model_with_cfg_llama(num_chunks=1, num_layers=48, **kwargs):
    """LLaMA-2: separate ``wq / wk / wv``."""
    cfg = _cfg({**_COMMON, **_V2_EXTRA, "adapt_hf": False}, kwargs)
    cfg["num_kv_attention_heads"] = cfg["num_kv_attention_heads"] or cfg["num_attention_heads"]
    return build_generic_model_1d(LLAMA_SPEC, num_layers=num_layers, num_chunks=num_chunks, **cfg)


@MODEL_INITIALIZER.register_module("INTERNLM_MoE")
def build_model_with_moe_cfg(num_chunks=1, num_layers=48, num_experts=1, moe_use_residual=False, moe_type="GShard",
                             **kwargs):
    """InternLM-v1 block whose MLP is a mixture of experts; forward returns ``(hidden, moe_losses)``."""
    cfg = _cfg(_COMMON, kwargs)
    cfg["num_kv_attention_heads"] = cfg["num_attention_heads"]
    moe_kwargs = dict(gpc.config.get("moe", {}) or {}) if gpc.config is not None else {}
    cfg["moe_cfg"] = dict(num_experts=num_experts, moe_use_residual=moe_use_residual, moe_type=moe_type, **moe_kwargs)
    return build_generic_model_1d(INTERNLM_SPEC, num_layers=num_layers, num_chunks=num_chunks, **cfg)
