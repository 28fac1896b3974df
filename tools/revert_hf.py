"""HuggingFace InternLM2 / LLaMA folder → training checkpoint shards ``model_tp{t}_pp0.pt`` (+ ``model_config.pt``) that
``ckpt.load_ckpt_info=dict(path=..., content=("model",), ckpt_type="internevo")`` can load with tensor-parallel size
``--tp_size`` (reference: ``transformers/revert_internlm{,2}.py``).

    python tools/revert_hf.py --src hf_model --tgt llm_ckpts/from_hf --tp_size 2 [--interleaved_rope]
"""
import argparse
import json
import os

import torch
from ckpt_io import save_sharded


def interleave_rows(w: torch.Tensor, head_dim: int) -> torch.Tensor:
    """inverse of convert2hf.deinterleave_rows"""
    shp = w.shape
    w = w.reshape(-1, 2, head_dim // 2, *shp[1:])
    return torch.stack([w[:, 0], w[:, 1]], dim=2).reshape(shp)


def load_hf_tensors(src):
    tensors = {}
    for fn in sorted(os.listdir(src)):
        if fn.endswith(".safetensors"):
            from safetensors.torch import load_file

            tensors.update(load_file(os.path.join(src, fn)))
        elif fn.startswith("pytorch_model") and fn.endswith(".bin"):
            tensors.update(torch.load(os.path.join(src, fn), map_location="cpu", weights_only=False))
    return tensors


def from_hf(hf, cfg, interleaved_rope: bool):
    H, Hkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads", cfg["num_attention_heads"])
    d = cfg["hidden_size"] // H
    full = {}
    # InternLM2 by its config, or - a config.json without ``model_type`` - by its fused ``wqkv`` projections
    if cfg.get("model_type") == "internlm2" or any(".attention.wqkv." in k for k in hf):
        for k, v in hf.items():
            k = k[6:] if k.startswith("model.") else k
            if k.endswith("attention.wqkv.weight") and interleaved_rope:
                gs = H // Hkv + 2
                g = v.reshape(Hkv, gs, d, -1).clone()
                qk = g[:, : gs - 1].reshape(-1, v.shape[-1])
                g[:, : gs - 1] = interleave_rows(qk, d).reshape(Hkv, gs - 1, d, -1)
                v = g.reshape(v.shape)
            full[k] = v
    else:
        ren = {"self_attn.q_proj": "attention.wq", "self_attn.k_proj": "attention.wk", "self_attn.v_proj": "attention.wv",
               "self_attn.o_proj": "attention.wo", "mlp.gate_proj": "feed_forward.w1", "mlp.up_proj": "feed_forward.w3",
               "mlp.down_proj": "feed_forward.w2", "input_layernorm": "attention_norm",
               "post_attention_layernorm": "ffn_norm"}
        for k, v in hf.items():
            if k == "model.embed_tokens.weight":
                full["tok_embeddings.weight"] = v
            elif k == "model.norm.weight":
                full["norm.weight"] = v
            elif k == "lm_head.weight":
                full["output.weight"] = v
            elif k.startswith("model.layers."):
                k = k[6:]
                for a, b in ren.items():
                    if f".{a}." in k:
                        if a in ("self_attn.q_proj", "self_attn.k_proj") and interleaved_rope:
                            v = interleave_rows(v, d)
                        k = k.replace(a, b)
                        break
                full[k] = v
    return full


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--src", required=True)
    p.add_argument("--tgt", required=True)
    p.add_argument("--tp_size", type=int, default=1)
    p.add_argument("--embed_split", action="store_true", help="embed_split_hidden of the training config")
    p.add_argument("--interleaved_rope", action="store_true")
    a = p.parse_args()
    cfg = json.load(open(os.path.join(a.src, "config.json")))
    full = from_hf(load_hf_tensors(a.src), cfg, a.interleaved_rope)
    save_sharded(full, a.tgt, a.tp_size, a.embed_split)
    model_config = dict(
        hidden_size=cfg["hidden_size"], num_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
        num_kv_attention_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]), vocab_size=cfg["vocab_size"],
        mlp_ratio=cfg["intermediate_size"] / cfg["hidden_size"], layer_norm_epsilon=cfg.get("rms_norm_eps", 1e-5),
        rope_base=cfg.get("rope_theta", 10000), embed_split_hidden=a.embed_split, no_bias=not cfg.get("bias", False),
        norm_type="rmsnorm", dtype="torch.bfloat16", adapt_hf=not a.interleaved_rope)
    torch.save(model_config, os.path.join(a.tgt, "model_config.pt"))
    print(f"wrote {a.tp_size} shard(s) to {a.tgt}")


if __name__ == "__main__":
    main()
