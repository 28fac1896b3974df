"""External weight loaders end to end (`ckpt.load_ckpt_info.ckpt_type = "hf_llama" | "llama"`): a random HF
`LlamaForCausalLM` is saved in HF format (and re-packed in Meta's consolidated format), loaded into the framework's LLAMA2
family through `internevo_b200.checkpoint.load_funcs`, and both models must produce the same logits - for both RoPE
conventions (`adapt_hf=True`: rotate-half as HF; `adapt_hf=False`: interleaved pairs as Meta, rows un-permuted on load)."""
import os

import pytest
import torch

from common import run_distributed, tiny_config


def _hf_llama(tmp, h=64, H=4, Hkv=2, L=2, V=96, F=256):
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=h, intermediate_size=F, num_hidden_layers=L, num_attention_heads=H, num_key_value_heads=Hkv,
                      vocab_size=V, max_position_embeddings=64, rms_norm_eps=1e-5, rope_theta=10000.0, tie_word_embeddings=False,
                      attention_bias=False)
    m = LlamaForCausalLM(cfg).float().eval()
    os.makedirs(tmp, exist_ok=True)
    torch.save(m.state_dict(), os.path.join(tmp, "pytorch_model.bin"))
    return m


def _to_meta(sd, H, Hkv, d):
    """HF llama -> Meta consolidated naming (q / k rows back to interleaved pairs)."""
    def unperm(w, nh):
        return w.view(nh, 2, d // 2, w.shape[-1]).transpose(1, 2).reshape(nh * d, w.shape[-1])

    out = {"tok_embeddings.weight": sd["model.embed_tokens.weight"], "norm.weight": sd["model.norm.weight"],
           "output.weight": sd["lm_head.weight"]}
    L = max(int(k.split(".")[2]) for k in sd if k.startswith("model.layers.")) + 1
    for i in range(L):
        p, q = f"model.layers.{i}.", f"layers.{i}."
        out[q + "attention.wq.weight"] = unperm(sd[p + "self_attn.q_proj.weight"], H)
        out[q + "attention.wk.weight"] = unperm(sd[p + "self_attn.k_proj.weight"], Hkv)
        out[q + "attention.wv.weight"] = sd[p + "self_attn.v_proj.weight"]
        out[q + "attention.wo.weight"] = sd[p + "self_attn.o_proj.weight"]
        out[q + "feed_forward.w1.weight"] = sd[p + "mlp.gate_proj.weight"]
        out[q + "feed_forward.w3.weight"] = sd[p + "mlp.up_proj.weight"]
        out[q + "feed_forward.w2.weight"] = sd[p + "mlp.down_proj.weight"]
        out[q + "attention_norm.weight"] = sd[p + "input_layernorm.weight"]
        out[q + "ffn_norm.weight"] = sd[p + "post_attention_layernorm.weight"]
    return out


def _load_and_forward(rank, world, folder, ckpt_type, adapt_hf, ids):
    from internevo_b200.checkpoint.load_funcs import LOAD_FUNC_DICT
    from internevo_b200.initialize import initialize_distributed_env
    from internevo_b200.train import initialize_model

    cfg = tiny_config(model_type="LLAMA2", num_layers=2, hidden=64, heads=4, kv_heads=2, vocab=96, seq_len=16, micro_bsz=1,
                      adapt_hf=adapt_hf)
    cfg["model"]["mlp_ratio"] = 4.0   # 256-aligned intermediate size = 256
    cfg["model"]["embed_split_hidden"] = False
    initialize_distributed_env(config=cfg, launcher="torch", seed=3)
    model = initialize_model()

    class _MM:   # what the CheckpointManager hands to a loader
        optimizer = None

    mm = _MM()
    mm.model = model.model
    LOAD_FUNC_DICT[ckpt_type](mm, dict(path=folder))
    model.eval()
    T = ids.shape[1]
    with torch.no_grad():
        out = model(input_ids=ids, cu_seqlens=torch.tensor([0, T], dtype=torch.int32), indexes=torch.arange(T)[None])
    out = out[0] if isinstance(out, (tuple, list)) else out
    return out.reshape(T, -1).float()


@pytest.mark.parametrize("ckpt_type,adapt_hf", [("hf_llama", True), ("hf_llama", False), ("llama", False)])
def test_llama_loaders_reproduce_hf_logits(tmp_path, ckpt_type, adapt_hf):
    hf = _hf_llama(str(tmp_path / "hf"))
    folder = str(tmp_path / "hf")
    if ckpt_type == "llama":
        folder = str(tmp_path / "meta")
        os.makedirs(folder)
        torch.save(_to_meta(hf.state_dict(), 4, 2, 16), os.path.join(folder, "consolidated.00.pth"))
    ids = torch.tensor([[1, 5, 9, 13, 40, 41, 7, 3, 90, 2]])
    with torch.no_grad():
        want = hf(input_ids=ids).logits[0]
    got = run_distributed(_load_and_forward, 1, folder, ckpt_type, adapt_hf, ids)[0]
    assert torch.allclose(got, want, atol=3e-4, rtol=1e-4), float((got - want).abs().max())
