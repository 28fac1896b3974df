"""Data-preparation and checkpoint-conversion tools (reference: tools/tokenizer.py, transformers/convert2hf_*.py,
revert_*.py have no tests of their own; the round trips below are the contract)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def sp_model(tmp_path_factory):
    import sentencepiece as spm

    d = tmp_path_factory.mktemp("sp")
    corpus = d / "corpus.txt"
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    rng = np.random.RandomState(0)
    corpus.write_text("\n".join(" ".join(rng.choice(words, 8)) for _ in range(200)))
    spm.SentencePieceTrainer.Train(input=str(corpus), model_prefix=str(d / "tok"), vocab_size=64, bos_id=1, eos_id=2,
                                   unk_id=0, pad_id=-1, model_type="bpe", minloglevel=2)
    return str(d / "tok.model"), str(corpus)


def test_tokenizer_bin_meta_roundtrip(sp_model, tmp_path):
    import tokenizer as tk

    from internevo_b200.data.datasets import JsonlDataset

    model, corpus = sp_model
    out = tmp_path / "train" / "en" / "c.bin"
    n = tk.text2bin(corpus, str(out), model)
    assert n == 200
    meta = np.load(str(out) + ".meta")
    assert meta.shape == (200, 2) and meta[0, 0] == 0
    ds = JsonlDataset(str(out), min_length=0)
    sp = tk.load_sp(model)
    first = ds[0]["tokens"]
    assert first[0] == sp.bos_id() and first[-1] == sp.eos_id() and len(first) == meta[0, 1]
    assert sp.decode(list(map(int, first[1:-1]))) == open(corpus).readline().strip()


def test_alpaca_masks_prompt(sp_model, tmp_path):
    model, _ = sp_model
    data = [{"instruction": "alpha beta", "input": "", "output": "gamma delta"},
            {"instruction": "eta", "input": "theta", "output": "iota"}] * 10
    src = tmp_path / "alpaca.json"
    src.write_text(json.dumps(data))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "alpaca_tokenizer.py"), str(src), str(tmp_path / "o"),
                        model, "--eoh_id", "60", "--eoa_id", "61", "--nl_id", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = open(tmp_path / "o" / "train" / "en" / "dataset.bin").readline()
    toks = json.loads(line)["tokens"]
    assert toks[0] == 1 and toks[-1] == 2 and any(t < 0 for t in toks) and toks[-3] == 61
    assert os.path.exists(tmp_path / "o" / "valid" / "en" / "dataset.bin.meta")


def _fake_internlm2(L=2, h=32, H=4, Hkv=2, F=64, V=48):
    g = torch.Generator().manual_seed(0)
    d = h // H
    sd = {"tok_embeddings.weight": torch.randn(V, h, generator=g), "norm.weight": torch.randn(h, generator=g),
          "output.weight": torch.randn(V, h, generator=g)}
    for i in range(L):
        p = f"layers.{i}."
        sd[p + "attention.wqkv.weight"] = torch.randn((H + 2 * Hkv) * d, h, generator=g)
        sd[p + "attention.wo.weight"] = torch.randn(h, h, generator=g)
        sd[p + "feed_forward.w1.weight"] = torch.randn(F, h, generator=g)
        sd[p + "feed_forward.w3.weight"] = torch.randn(F, h, generator=g)
        sd[p + "feed_forward.w2.weight"] = torch.randn(h, F, generator=g)
        sd[p + "attention_norm.weight"] = torch.randn(h, generator=g)
        sd[p + "ffn_norm.weight"] = torch.randn(h, generator=g)
    cfg = dict(hidden_size=h, num_layers=L, num_attention_heads=H, num_kv_attention_heads=Hkv, vocab_size=V, mlp_ratio=F / h)
    return sd, cfg


@pytest.mark.parametrize("interleaved", [False, True])
def test_hf_roundtrip_internlm2(tmp_path, interleaved):
    import ckpt_io
    import convert2hf
    import revert_hf

    full, cfg = _fake_internlm2()
    src = tmp_path / "ckpt"
    ckpt_io.save_sharded(full, str(src), tp_size=2, embed_split_hidden=True)
    torch.save(cfg, src / "model_config.pt")
    merged = ckpt_io.load_full_state(str(src), True)
    assert all(torch.equal(merged[k], full[k]) for k in full)

    hf, hf_cfg = convert2hf.to_hf(merged, cfg, "internlm2", interleaved)
    convert2hf.save_hf(hf, hf_cfg, str(tmp_path / "hf"), torch.float32, 1 << 12)
    assert len([f for f in os.listdir(tmp_path / "hf") if f.endswith(".safetensors")]) > 1  # sharding honoured
    back = revert_hf.from_hf(revert_hf.load_hf_tensors(str(tmp_path / "hf")), hf_cfg, interleaved)
    assert set(back) == set(full) and all(torch.equal(back[k], full[k]) for k in full)
    if interleaved:  # the permutation is a real one
        assert not torch.equal(hf["model.layers.0.attention.wqkv.weight"], full["layers.0.attention.wqkv.weight"])


def test_hf_roundtrip_llama_names():
    import convert2hf
    import revert_hf

    full, cfg = _fake_internlm2()
    d = cfg["hidden_size"] // cfg["num_attention_heads"]
    ll = {}
    for k, v in full.items():
        if k.endswith("wqkv.weight"):
            p = k[: -len("wqkv.weight")]
            ll[p + "wq.weight"], ll[p + "wk.weight"], ll[p + "wv.weight"] = v[: 4 * d], v[4 * d: 6 * d], v[6 * d:]
        else:
            ll[k] = v
    hf, hf_cfg = convert2hf.to_hf(ll, cfg, "llama", True)
    assert "model.layers.1.self_attn.q_proj.weight" in hf and "lm_head.weight" in hf
    back = revert_hf.from_hf(hf, hf_cfg, True)
    assert set(back) == set(ll) and all(torch.equal(back[k], ll[k]) for k in ll)
