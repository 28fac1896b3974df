"""HF `transformers` model code shipped with the converters (reference `transformers/internlm{,2}_model`): a checkpoint
written by `tools/convert2hf.py` must load into `huggingface/internlm2_model` and give the SAME logits as the framework's
own model; KV-cache generation must equal cache-free greedy decoding."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)


def _tiny_model(family="INTERNLM2_PUBLIC", **extra):
    from load_internlm_model import initialize_internlm_model

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29641")
    cfg = dict(num_layers=2, hidden_size=64, num_attention_heads=4, num_kv_attention_heads=2, vocab_size=96, mlp_ratio=2.0,
               embed_split_hidden=False, no_bias=True, norm_type="rmsnorm", layer_norm_epsilon=1e-5, use_flash_attn=True,
               max_position_embeddings=64)
    cfg.update(extra)
    torch.manual_seed(0)
    return initialize_internlm_model(family, None, cfg, param_dtype=torch.float32), cfg


def _framework_logits(model, ids):
    T = ids.shape[1]
    out = model(input_ids=ids, cu_seqlens=torch.tensor([0, T], dtype=torch.int32), indexes=torch.arange(T)[None])
    out = out[0] if isinstance(out, (tuple, list)) else out
    return out.reshape(T, -1)


@pytest.mark.parametrize("interleaved", [False, True])
def test_converted_checkpoint_gives_same_logits_in_hf_model(tmp_path, interleaved):
    import convert2hf
    from huggingface.internlm2_model import InternLM2Config, InternLM2ForCausalLM

    model, cfg = _tiny_model(adapt_hf=not interleaved)
    inner = model.model if hasattr(model, "model") else model
    assert inner.layer_list[0].attention.interleaved_rope == interleaved
    full = {k: v.detach().clone() for k, v in inner.state_dict().items()}
    hf_sd, hf_cfg = convert2hf.to_hf(full, cfg, "internlm2", interleaved)
    convert2hf.save_hf(hf_sd, hf_cfg, str(tmp_path / "hf"), torch.float32, 1 << 30)

    hf_cfg = {k: v for k, v in hf_cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    hf = InternLM2ForCausalLM(InternLM2Config(max_position_embeddings=64, **hf_cfg)).float().eval()
    from safetensors.torch import load_file

    sd = {}
    for f in os.listdir(tmp_path / "hf"):
        if f.endswith(".safetensors"):
            sd.update(load_file(str(tmp_path / "hf" / f)))
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m for m in missing), (missing, unexpected)

    ids = torch.tensor([[1, 5, 9, 13, 40, 41, 7, 3, 90, 2, 17]])
    with torch.no_grad():
        want = _framework_logits(model, ids)
        got = hf(input_ids=ids).logits[0]
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), float((got - want).abs().max())

    # KV-cache generation == cache-free greedy decoding
    with torch.no_grad():
        gen = hf.generate(ids[:, :4], max_new_tokens=6, do_sample=False, use_cache=True)
        seq = ids[:, :4]
        for _ in range(6):
            seq = torch.cat([seq, hf(input_ids=seq, use_cache=False).logits[:, -1].argmax(-1, keepdim=True)], 1)
    assert torch.equal(gen, seq), (gen, seq)


def test_hf_padding_mask_and_loss():
    from huggingface.internlm2_model import InternLM2Config, InternLM2ForCausalLM, InternLM2ForSequenceClassification

    torch.manual_seed(0)
    cfg = InternLM2Config(vocab_size=50, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                          num_key_value_heads=2, max_position_embeddings=32, num_labels=1)
    m = InternLM2ForCausalLM(cfg).eval()
    a = torch.tensor([[3, 4, 5, 6]])
    padded = torch.tensor([[3, 4, 5, 6, 0, 0]])
    mask = torch.tensor([[1, 1, 1, 1, 0, 0]])
    with torch.no_grad():
        la = m(input_ids=a).logits
        lp = m(input_ids=padded, attention_mask=mask).logits
    assert torch.allclose(la, lp[:, :4], atol=1e-5)           # right padding does not leak into real tokens
    out = m(input_ids=a, labels=a)
    assert out.loss is not None and out.loss.requires_grad
    rm = InternLM2ForSequenceClassification(cfg).eval()
    with torch.no_grad():
        s1 = rm(input_ids=a).logits
        s2 = rm(input_ids=padded, attention_mask=mask).logits
    assert s1.shape == (1, 1) and torch.allclose(s1, s2, atol=1e-5)   # score is taken at the last real token


def test_internlm_v1_converted_checkpoint_same_logits():
    import convert2hf
    from huggingface.internlm_model import InternLMConfig, InternLMForCausalLM

    model, cfg = _tiny_model("INTERNLM", num_kv_attention_heads=4)
    inner = model.model if hasattr(model, "model") else model
    full = {k: v.detach().clone() for k, v in inner.state_dict().items()}
    hf_sd, hf_cfg = convert2hf.to_hf(full, cfg, "internlm", False)
    hf_cfg = {k: v for k, v in hf_cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
    hf = InternLMForCausalLM(InternLMConfig(max_position_embeddings=64, **hf_cfg)).float().eval()
    missing, unexpected = hf.load_state_dict({k: v.float() for k, v in hf_sd.items()}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    ids = torch.tensor([[1, 5, 9, 13, 40, 41, 7, 3]])
    with torch.no_grad():
        want = _framework_logits(model, ids)
        got = hf(input_ids=ids).logits[0]
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), float((got - want).abs().max())


def test_hf_sentencepiece_tokenizer_roundtrip(tmp_path):
    import sentencepiece as spm

    corpus = tmp_path / "c.txt"
    corpus.write_text("\n".join(["hello world this is a tiny corpus", "the quick brown fox jumps over the lazy dog",
                                 "internlm tokenizer test sentence number three"] * 30))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=64, bos_id=1, eos_id=2,
                                   unk_id=0, pad_id=-1, model_type="bpe", minloglevel=2)
    from huggingface.internlm2_model.tokenization_internlm2 import InternLM2Tokenizer
    from huggingface.internlm_model.tokenization_internlm import InternLMTokenizer

    tok = InternLM2Tokenizer(str(tmp_path / "tok.model"))
    enc = tok("hello world")["input_ids"]
    assert enc[0] == tok.bos_token_id == 1 and tok.eos_token_id == 2
    assert tok.decode(enc, skip_special_tokens=True).strip() == "hello world"
    tok_eos = InternLMTokenizer(str(tmp_path / "tok.model"), add_eos_token=True)
    assert tok_eos("the fox")["input_ids"][-1] == 2
    out = tok.save_vocabulary(str(tmp_path))
    assert os.path.isfile(out[0]) and tok.vocab_size == 64 and len(tok.get_vocab()) >= 64


def test_convert2hf_installs_remote_code_loadable_with_auto_classes(tmp_path):
    """The converted folder is self-contained: AutoConfig / AutoModelForCausalLM load it with trust_remote_code."""
    import convert2hf
    from transformers import AutoConfig, AutoModelForCausalLM

    model, cfg = _tiny_model()
    inner = model.model if hasattr(model, "model") else model
    full = {k: v.detach().clone() for k, v in inner.state_dict().items()}
    hf_sd, hf_cfg = convert2hf.to_hf(full, cfg, "internlm2", False)
    hf_cfg["max_position_embeddings"] = 64
    tgt = str(tmp_path / "hf")
    convert2hf.save_hf(hf_sd, hf_cfg, tgt, torch.float32, 1 << 30)
    convert2hf.install_remote_code(tgt, "internlm2")
    assert {"modeling_internlm2.py", "configuration_internlm2.py", "tokenization_internlm2.py", "tokenizer_config.json"} <= set(
        os.listdir(tgt))
    conf = AutoConfig.from_pretrained(tgt, trust_remote_code=True)
    assert type(conf).__name__ == "InternLM2Config" and conf.num_key_value_heads == 2
    hf = AutoModelForCausalLM.from_pretrained(tgt, trust_remote_code=True, torch_dtype=torch.float32).eval()
    ids = torch.tensor([[1, 5, 9, 13, 40, 41, 7]])
    with torch.no_grad():
        assert torch.allclose(hf(input_ids=ids).logits[0], _framework_logits(model, ids), atol=2e-4, rtol=1e-4)


def test_stream_chat_yields_growing_responses_and_rope_scaling_classes():
    """``stream_chat`` (generator over a worker-thread ``generate``) ends on the same text as ``chat`` with greedy decoding;
    the named rope-scaling classes are the linear / dynamic-NTK configurations of the one rotary module."""
    import torch
    from huggingface.internlm2_model.configuration_internlm2 import InternLM2Config
    from huggingface.internlm2_model.modeling_internlm2 import (InternLM2DynamicNTKScalingRotaryEmbedding,
                                                                InternLM2ForCausalLM,
                                                                InternLM2LinearScalingRotaryEmbedding,
                                                                InternLM2RotaryEmbedding)

    class Tok:   # the slice of the tokenizer API the chat helpers use: ids are code points folded into the vocabulary
        bos_token, eos_token_id = "", 2

        def __call__(self, texts, return_tensors="pt"):
            ids = [[3 + (ord(c) % 90) for c in texts[0]][-24:]]
            return {"input_ids": torch.tensor(ids), "attention_mask": torch.ones(1, len(ids[0]), dtype=torch.long)}

        def convert_tokens_to_ids(self, t):
            return 1

        def decode(self, ids, skip_special_tokens=True):
            return "".join(chr(97 + (int(i) % 26)) for i in ids)

    torch.manual_seed(0)
    cfg = InternLM2Config(vocab_size=96, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                          num_key_value_heads=2, max_position_embeddings=128, pad_token_id=0)
    model = InternLM2ForCausalLM(cfg).float().eval()
    final, hist = model.chat(Tok(), "hi", max_new_tokens=8, do_sample=False)
    seen = list(model.stream_chat(Tok(), "hi", max_new_tokens=8, do_sample=False))
    assert seen[0][0] == "" and seen[-1][0] == final and seen[-1][1] == hist
    lens = [len(r) for r, _ in seen]
    assert lens == sorted(lens) and len(seen) >= 3
    pos = torch.arange(0, 300)[None]
    base = InternLM2RotaryEmbedding(16, 128)(pos, torch.float32)[0]
    lin = InternLM2LinearScalingRotaryEmbedding(16, 128, scaling_factor=2.0)(pos, torch.float32)[0]
    ntk = InternLM2DynamicNTKScalingRotaryEmbedding(16, 128, scaling_factor=2.0)(pos, torch.float32)[0]
    assert torch.allclose(lin[0, 200], base[0, 100], atol=1e-5)          # positions halved
    assert not torch.allclose(ntk[0, 200], base[0, 200], atol=1e-3)      # base stretched beyond max_position_embeddings
