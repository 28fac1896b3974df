"""Training checkpoint (``model_tp*_pp*.pt`` + ``model_config.pt``) → HuggingFace folder (``config.json`` + safetensors
shards + index), for the InternLM2 / LLaMA-2 / InternLM(v1) families.

    python tools/convert2hf.py --src llm_ckpts/1000 --tgt hf_out --family internlm2 [--tokenizer tokenizer.model]

Same job as the reference ``transformers/convert2hf_internlm{,2}.py``; here the tensors are written directly (no remote
model code needs to be importable).  ``wqkv`` keeps the grouped ``(kv_head, q_per_kv + 2, head_dim)`` layout that the HF
InternLM2 implementation uses; with interleaved RoPE (``adapt_hf=False`` at training time) the q/k rows of every head are
re-ordered from (even, odd) pairs to the half-split order HF's ``rotate_half`` expects.
"""
import argparse
import json
import os
import shutil

import torch
from ckpt_io import load_full_state, load_model_config


def deinterleave_rows(w: torch.Tensor, head_dim: int) -> torch.Tensor:
    """rows [..., d] of every head: (x0, x1, x2, ...) → (x0, x2, ..., x1, x3, ...)"""
    shp = w.shape
    w = w.reshape(-1, head_dim, *shp[1:])
    return torch.cat([w[:, 0::2], w[:, 1::2]], dim=1).reshape(shp)


def to_hf(full, cfg, family: str, interleaved_rope: bool):
    H = cfg["num_attention_heads"]
    Hkv = cfg.get("num_kv_attention_heads") or H
    d = cfg["hidden_size"] // H
    out = {}
    if family == "internlm2":
        for k, v in full.items():
            if k.endswith("attention.wqkv.weight") and interleaved_rope:
                gs = H // Hkv + 2
                g = v.reshape(Hkv, gs, d, -1).clone()
                qk = g[:, : gs - 1].reshape(-1, v.shape[-1])
                g[:, : gs - 1] = deinterleave_rows(qk, d).reshape(Hkv, gs - 1, d, -1)
                v = g.reshape(v.shape)
            out[k if k.startswith("output.") else "model." + k] = v
        arch, mtype = "InternLM2ForCausalLM", "internlm2"
    elif family == "llama":
        ren = {"attention.wq": "self_attn.q_proj", "attention.wk": "self_attn.k_proj", "attention.wv": "self_attn.v_proj",
               "attention.wo": "self_attn.o_proj", "feed_forward.w1": "mlp.gate_proj", "feed_forward.w3": "mlp.up_proj",
               "feed_forward.w2": "mlp.down_proj", "attention_norm": "input_layernorm", "ffn_norm": "post_attention_layernorm"}
        for k, v in full.items():
            if k == "tok_embeddings.weight":
                out["model.embed_tokens.weight"] = v
            elif k == "norm.weight":
                out["model.norm.weight"] = v
            elif k == "output.weight":
                out["lm_head.weight"] = v
            else:
                for a, b in ren.items():
                    if f".{a}." in k:
                        if a in ("attention.wq", "attention.wk") and interleaved_rope:
                            v = deinterleave_rows(v, d)
                        k = k.replace(a, b)
                        break
                out["model." + k] = v
        arch, mtype = "LlamaForCausalLM", "llama"
    else:  # internlm v1: fused Wqkv [3, H, d] → q/k/v projections with bias
        for k, v in full.items():
            if ".mixer.Wqkv." in k:
                q, kk, vv = v.reshape(3, -1, *v.shape[1:]).unbind(0)
                if interleaved_rope:
                    q, kk = deinterleave_rows(q, d), deinterleave_rows(kk, d)
                base = "model." + k.replace("blocks.", "layers.").split(".mixer.")[0] + ".self_attn."
                suf = k.rsplit(".", 1)[1]
                out[base + "q_proj." + suf], out[base + "k_proj." + suf], out[base + "v_proj." + suf] = q, kk, vv
                continue
            nk = (k.replace("blocks.", "layers.").replace(".mixer.out_proj.", ".self_attn.o_proj.")
                  .replace(".mlp.w1.", ".mlp.gate_proj.").replace(".mlp.w3.", ".mlp.up_proj.")
                  .replace(".mlp.w2.", ".mlp.down_proj.").replace(".norm1.", ".input_layernorm.")
                  .replace(".norm2.", ".post_attention_layernorm."))
            if nk == "embedding.weight":
                nk = "embed_tokens.weight"
            if nk == "head.weight":
                out["lm_head.weight"] = v
            else:
                out["model." + nk] = v
        arch, mtype = "InternLMForCausalLM", "internlm"
    mlp = cfg.get("mlp_ratio", 8 / 3)
    inter = int(cfg["hidden_size"] * mlp)
    inter = 256 * ((inter + 255) // 256)
    hf_cfg = {
        "architectures": [arch], "model_type": mtype, "hidden_size": cfg["hidden_size"], "num_hidden_layers": cfg["num_layers"],
        "num_attention_heads": H, "num_key_value_heads": Hkv, "intermediate_size": inter, "vocab_size": cfg["vocab_size"],
        "rms_norm_eps": cfg.get("layer_norm_epsilon", 1e-5), "hidden_act": "silu", "bias": family == "internlm",
        "rope_theta": cfg.get("rope_base", 10000), "tie_word_embeddings": False, "torch_dtype": "bfloat16",
    }
    return out, hf_cfg


def save_hf(tensors, hf_cfg, tgt, dtype, max_shard_bytes):
    from safetensors.torch import save_file

    os.makedirs(tgt, exist_ok=True)
    shards, cur, size = [], {}, 0
    for k in sorted(tensors):
        t = tensors[k].to(dtype).contiguous()
        n = t.numel() * t.element_size()
        if cur and size + n > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = t
        size += n
    shards.append(cur)
    index = {"metadata": {"total_size": 0}, "weight_map": {}}
    for i, sh in enumerate(shards):
        fn = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(tgt, fn), metadata={"format": "pt"})
        for k, t in sh.items():
            index["weight_map"][k] = fn
            index["metadata"]["total_size"] += t.numel() * t.element_size()
    json.dump(index, open(os.path.join(tgt, "model.safetensors.index.json"), "w"), indent=1)
    json.dump(hf_cfg, open(os.path.join(tgt, "config.json"), "w"), indent=1)


def install_remote_code(tgt: str, family: str):
    """Make the folder self-contained for ``from_pretrained(tgt, trust_remote_code=True)``: copy the HF model / tokenizer
    code of ``huggingface/{family}_model`` next to the weights (flattened: the v1 files import the shared helpers from the
    InternLM2 files) and register it in ``config.json`` / ``tokenizer_config.json`` (reference
    ``transformers/convert2hf_internlm2.py:246-262``)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "huggingface")
    files = ["internlm2_model/configuration_internlm2.py", "internlm2_model/modeling_internlm2.py",
             "internlm2_model/tokenization_internlm2.py"]
    if family == "internlm":
        files += ["internlm_model/configuration_internlm.py", "internlm_model/modeling_internlm.py",
                  "internlm_model/tokenization_internlm.py"]
    for f in files:
        src = open(os.path.join(root, f)).read()
        src = src.replace("from ..internlm2_model.", "from .")     # one flat folder
        with open(os.path.join(tgt, os.path.basename(f)), "w") as out:
            out.write(src)
    tag = "internlm2" if family == "internlm2" else "internlm"
    cls = "InternLM2" if family == "internlm2" else "InternLM"
    cfg_path = os.path.join(tgt, "config.json")
    cfg = json.load(open(cfg_path))
    cfg["auto_map"] = {"AutoConfig": f"configuration_{tag}.{cls}Config", "AutoModel": f"modeling_{tag}.{cls}ForCausalLM",
                       "AutoModelForCausalLM": f"modeling_{tag}.{cls}ForCausalLM"}
    json.dump(cfg, open(cfg_path, "w"), indent=1)
    json.dump({"auto_map": {"AutoTokenizer": [f"tokenization_{tag}.{cls}Tokenizer", None]}, "tokenizer_class": f"{cls}Tokenizer",
               "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>", "pad_token": "</s>", "add_bos_token": True,
               "add_eos_token": False, "clean_up_tokenization_spaces": False},
              open(os.path.join(tgt, "tokenizer_config.json"), "w"), indent=1)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--src", required=True)
    p.add_argument("--tgt", required=True)
    p.add_argument("--family", choices=["internlm2", "llama", "internlm"], default="internlm2")
    p.add_argument("--dtype", default="bfloat16")
    p.add_argument("--max_shard", default="10GB")
    p.add_argument("--max_pos", type=int, default=4096)
    p.add_argument("--tokenizer", default=None)
    p.add_argument("--interleaved_rope", action="store_true", help="the model was trained with adapt_hf=False")
    a = p.parse_args()
    cfg = load_model_config(a.src)
    full = load_full_state(a.src, cfg.get("embed_split_hidden", True))
    tensors, hf_cfg = to_hf(full, cfg, a.family, a.interleaved_rope or cfg.get("adapt_hf") is False)
    hf_cfg["max_position_embeddings"] = a.max_pos
    unit = {"GB": 1 << 30, "MB": 1 << 20}
    max_bytes = int(float(a.max_shard[:-2]) * unit[a.max_shard[-2:].upper()])
    save_hf(tensors, hf_cfg, a.tgt, getattr(torch, a.dtype), max_bytes)
    if a.tokenizer:
        shutil.copy(a.tokenizer, os.path.join(a.tgt, "tokenizer.model"))
    if a.family in ("internlm2", "internlm"):
        install_remote_code(a.tgt, a.family)
    print(f"wrote {len(tensors)} tensors to {a.tgt}")


if __name__ == "__main__":
    main()
