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


# ---------------------------------------------------------------------------------------------------------------------
# generation interface, PAL runtime, MOSS SFT data
# ---------------------------------------------------------------------------------------------------------------------
class _CharTok:
    """Minimal tokenizer double: one id per character (ids 3..), bos 1, eos 2."""
    eos_token_id, bos_token_id = 2, 1

    def __call__(self, texts, return_tensors=None, **kw):
        return {"input_ids": torch.tensor([[1] + [3 + (ord(c) % 40) for c in texts[0]]])}

    def encode(self, text, add_special_tokens=True):
        return ([1] if add_special_tokens else []) + [3 + (ord(c) % 40) for c in text]

    def decode(self, ids, skip_special_tokens=True):
        return "".join(chr(97 + (i - 3) % 26) for i in ids if i > 2)


def test_generate_interactive_streams_greedy_tokens():
    sys.path.insert(0, ROOT)
    import interface
    from huggingface.internlm2_model import InternLM2Config, InternLM2ForCausalLM

    torch.manual_seed(0)
    m = InternLM2ForCausalLM(InternLM2Config(vocab_size=50, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                             num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64)).eval()
    tok = _CharTok()
    cfg = interface.GenerationConfig(max_new_tokens=6, do_sample=False)
    chunks = list(interface.generate_interactive(m, tok, "hello", cfg))
    assert 1 <= len(chunks) <= 6 and all(chunks[i + 1].startswith(chunks[i]) for i in range(len(chunks) - 1))
    ids = tok(["hello"])["input_ids"]
    with torch.no_grad():
        ref = m.generate(ids, max_new_tokens=6, do_sample=False, eos_token_id=2)
    assert chunks[-1] == tok.decode(ref[0, ids.shape[1]:].tolist())
    # sampling utilities: top-k = 1 is greedy, repetition penalty demotes seen tokens
    lg = torch.tensor([1.0, 3.0, 2.0])
    assert interface.sample_next(lg, [], interface.GenerationConfig(top_k=1, top_p=1.0)) == 1
    assert interface.sample_next(lg, [1], interface.GenerationConfig(do_sample=False, repetition_penalty=4.0)) == 2
    kept = interface._filter_logits(torch.tensor([0.0, 5.0, 4.9, -3.0]), 0, 0.6)
    assert torch.isinf(kept[0]) and torch.isinf(kept[3]) and not torch.isinf(kept[1])


def test_pal_runtime_extracts_runs_and_times_out():
    import pal_inference as pal

    gen = "Sure.\n```python\ndef solution():\n    a = 23 - 5 * 3\n    return a\n```\nDone"
    code = pal.extract_code(gen)
    assert code[0].startswith("def solution")
    assert pal.run_with_timeout(code, time_out=20) == ("ok", 8)
    assert pal.run_with_timeout(["def solution():", "    while True:", "        pass"], time_out=1.0)[0] == "timeout"
    assert pal.run_with_timeout(["def solution():", "    return 1 / 0"], time_out=20)[0] == "err"
    assert pal.gold_answer("blah blah #### 1,234") == 1234.0 and pal.is_correct(8.0004, 8.0) and not pal.is_correct("x", 8.0)
    iface = pal.PALInterface(lambda prompt: gen, time_out=20)
    value, g, status = iface.run("Olivia has $23 ...")
    assert value == 8 and status == "ok" and "How about this question" in pal.PROMPT_HEAD


def test_moss_sft_processing_masks_instruction_and_cuts_on_turn_boundary():
    import moss_002_sft as ms

    tok = _CharTok()
    sample = {"prefix": "sys:", "num_turns": 3, "plain_text": "<|Human|>: hi<eoh> <|MOSS|>: yo<eoa><|Human|>: a<eoh> <|MOSS|>: b<eoa>"
              "<|Human|>: c<eoh> <|MOSS|>: dddddddddddddddddddddddddddddddddddddddddddddddd<eoa>"}
    full = ms.process(sample, tok, 10_000)
    cut = ms.process(sample, tok, len(full["input_ids"]) - 5)
    assert full["no_loss_spans"] == [(0, 5)] and len(cut["input_ids"]) < len(full["input_ids"])
    assert ms.process(sample, tok, 8) == {"input_ids": [], "no_loss_spans": []}
    ds = ms.SFTDataset([full, cut])
    data, label = ds[0]
    assert torch.all(label[:5] == -100) and torch.equal(label[5:], data[5:])
    batch = ms.collate_fn([ds[0], ds[1]], tok)
    assert batch["input_ids"].shape == batch["labels"].shape == batch["attention_mask"].shape
    assert int(batch["attention_mask"][1].sum()) == len(cut["input_ids"]) and batch["labels"][1, -1] == -100


def test_openai_api_chat_completions_and_streaming(sp_model):
    """`/v1/models`, `/v1/chat/completions` (plain and SSE streaming) over a tiny random model: OpenAI response schema,
    `max_tokens` counts generated tokens, streamed deltas concatenate to a complete message."""
    import sentencepiece as spm
    from fastapi.testclient import TestClient

    import openai_api
    from load_internlm_model import initialize_internlm_model

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29661")
    cfg = dict(num_layers=2, hidden_size=64, num_attention_heads=4, num_kv_attention_heads=2, vocab_size=64, mlp_ratio=2.0,
               embed_split_hidden=False, no_bias=True, norm_type="rmsnorm", layer_norm_epsilon=1e-5, use_flash_attn=True,
               max_position_embeddings=256)
    torch.manual_seed(0)
    sp = spm.SentencePieceProcessor()
    sp.Load(sp_model[0])
    openai_api.STATE.update(model=initialize_internlm_model("INTERNLM2_PUBLIC", None, cfg, param_dtype=torch.float32),
                            tokenizer=sp, name="tiny")
    client = TestClient(openai_api.app)
    assert client.get("/v1/models").json()["data"][0]["id"] == "tiny"
    body = {"model": "tiny", "messages": [{"role": "system", "content": "be brief"}, {"role": "user", "content": "alpha beta"}],
            "max_tokens": 6, "temperature": 0.0}
    r = client.post("/v1/chat/completions", json=body).json()
    assert r["object"] == "chat.completion" and r["choices"][0]["message"]["role"] == "assistant"
    assert r["choices"][0]["finish_reason"] == "stop" and isinstance(r["choices"][0]["message"]["content"], str)
    prompt = openai_api.build_prompt([openai_api.Message(**m) for m in body["messages"]])
    assert prompt.startswith("<|System|>:be brief\n<|User|>:alpha beta<eoh>\n") and prompt.endswith("<|Bot|>:")
    with client.stream("POST", "/v1/chat/completions", json=dict(body, stream=True)) as resp:
        lines = [ln for ln in resp.iter_lines() if ln.startswith("data: ")]
    assert lines[-1] == "data: [DONE]" and 1 <= len(lines) - 1 <= 6          # at most max_tokens chunks
    deltas = [json.loads(ln[6:])["choices"][0]["delta"]["content"] for ln in lines[:-1]]
    assert "".join(deltas) == r["choices"][0]["message"]["content"]           # greedy: streaming == non-streaming


def test_ckpt_io_merges_weight_parallel_checkpoints(tmp_path):
    """``model_tp{t}_wp{w}_pp{p}.pt`` (isp) folders merge into the same full state dict they were cut from: linears by weight rank,
    embedding by the tensor ranks' hidden slices, head by their vocabulary rows."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ckpt_io

    torch.manual_seed(0)
    V, h, F = 32, 16, 24
    full = {"tok_embeddings.weight": torch.randn(V, h), "norm.weight": torch.randn(h), "output.weight": torch.randn(V, h)}
    for i in range(2):
        p = f"layers.{i}."
        full[p + "attention.wqkv.weight"] = torch.randn(24, h)
        full[p + "attention.wo.weight"] = torch.randn(h, h)
        full[p + "attention_norm.weight"] = torch.randn(h)
        full[p + "ffn_norm.weight"] = torch.randn(h)
        for n, s in (("w1", (F, h)), ("w3", (F, h)), ("w2", (h, F))):
            full[p + f"feed_forward.{n}.weight"] = torch.randn(*s)
    for tp, wp in ((2, 2), (2, 4), (1, 4)):
        d = str(tmp_path / f"tp{tp}wp{wp}")
        ckpt_io.save_sharded_isp(full, d, tp, wp)
        assert len(ckpt_io.find_isp_shards(d)) == wp
        one = torch.load(os.path.join(d, f"model_tp{(wp - 1) % tp}_wp{wp - 1}_pp0.pt"), weights_only=False)
        assert tuple(one["tok_embeddings.weight"].shape) == (V, h // tp) and tuple(one["output.weight"].shape) == (V // tp, h)
        assert tuple(one["layers.0.feed_forward.w2.weight"].shape) == (h // wp, F)
        back = ckpt_io.load_full_state(d)
        assert set(back) == set(full) and all(torch.equal(back[k], full[k]) for k in full), (tp, wp)


def test_ci_flow_driver_runs_the_user_journey_on_two_cpu_ranks(tmp_path):
    """``ci_scripts/flow.py all``: shards -> train.py (case merged over configs/demo.py, checkpoint file set checked) -> HF
    conversion -> Auto* load -> resumed second launch; plus the pieces on their own (case merge, expected file sets)."""
    import subprocess

    import numpy as np
    import sentencepiece as spm

    sys.path.insert(0, os.path.join(ROOT, "ci_scripts"))
    import flow

    assert flow.deep_merge({"a": {"b": 1, "c": 2}, "d": 3}, {"a": {"b": 5}, "e": 6}) == {"a": {"b": 5, "c": 2}, "d": 3, "e": 6}
    cfg = flow.write_case_config("tp2_dp4", str(tmp_path / "tp2.py"))
    assert cfg["parallel"]["tensor"]["size"] == 2 and cfg["lr_scheduler"]["total_steps"] == cfg["data"]["total_steps"] == 10
    want = flow.expected_checkpoint_files(cfg, 8, 10)
    assert {"model_tp0_pp0.pt", "model_tp1_pp0.pt", "optimizer_tp1_pp0_zo3.pt", "10.step"} <= want and len(want) == 5 + 2 + 8
    from internevo_b200.core.context import Config

    assert Config.from_file(str(tmp_path / "tp2.py")).model["num_layers"] == 8          # the generated file is a loadable config

    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    rng = np.random.RandomState(0)
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(rng.choice(words, 12)) for _ in range(400)))
    spm.SentencePieceTrainer.Train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=64, bos_id=1, eos_id=2,
                                   unk_id=0, pad_id=-1, model_type="bpe", minloglevel=2)
    from common import find_free_port

    r = subprocess.run([sys.executable, "ci_scripts/flow.py", "all", "--corpus", str(corpus), "--tokenizer",
                        str(tmp_path / "tok.model"), "--work", str(tmp_path / "work"), "--case", "cpu_smoke", "--cpu", "--nproc", "2",
                        "--port", str(find_free_port())], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "flow ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert sorted(d for d in os.listdir(tmp_path / "work" / "llm_ckpts") if d.isdigit()) == ["12", "6"]


@pytest.mark.parametrize("src_layout,tgt_layout", [((1, 1), (2, 2)), ((2, 2), (4, 1)), ((2, 1), (1, 4))])
def test_reshard_ckpt_between_tensor_and_pipeline_layouts(tmp_path, src_layout, tgt_layout):
    """``tools/reshard_ckpt.py``: files of one tp x pp layout → another; merging the result gives the same full model, every
    stage numbers its layers from 0 and only the first / last stage carry embedding / head."""
    import ckpt_io
    import reshard_ckpt

    full, cfg = _fake_internlm2(L=5)        # 5 layers over 2 / 4 stages: uneven stages, later ones larger
    src, tgt = str(tmp_path / "src"), str(tmp_path / "tgt")
    ckpt_io.save_sharded(full, src, tp_size=src_layout[0], embed_split_hidden=True, pp_size=src_layout[1])
    torch.save(dict(cfg, embed_split_hidden=True), os.path.join(src, "model_config.pt"))
    assert reshard_ckpt.main(["--src", src, "--tgt", tgt, "--tp", str(tgt_layout[0]), "--pp", str(tgt_layout[1])]) == 0
    assert ckpt_io.find_shards(tgt) == tgt_layout and os.path.exists(os.path.join(tgt, "model_config.pt"))
    back = ckpt_io.load_full_state(tgt, True)
    assert set(back) == set(full) and all(torch.equal(back[k], full[k]) for k in full)
    last = torch.load(os.path.join(tgt, f"model_tp0_pp{tgt_layout[1] - 1}.pt"), weights_only=False)
    first = torch.load(os.path.join(tgt, "model_tp0_pp0.pt"), weights_only=False)
    assert "tok_embeddings.weight" in first and "output.weight" in last and "layers.0.attention.wqkv.weight" in last
    if tgt_layout[1] > 1:
        assert "output.weight" not in first and "tok_embeddings.weight" not in last
        assert first["layers.0.attention.wqkv.weight"].shape[0] * tgt_layout[0] == full["layers.0.attention.wqkv.weight"].shape[0]
