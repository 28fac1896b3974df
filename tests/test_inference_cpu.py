"""Generation API (reference tests: tests/test_model/…; apis/inference.py has none): KV-cache decode must reproduce the
tokens of a cache-free full forward, beam search must return `num_return_sequences` finished hypotheses."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _tiny_model():
    from load_internlm_model import initialize_internlm_model

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29631")
    cfg = dict(num_layers=2, hidden_size=64, num_attention_heads=4, num_kv_attention_heads=2, vocab_size=96, mlp_ratio=2.0,
               embed_split_hidden=False, no_bias=True, norm_type="rmsnorm", layer_norm_epsilon=1e-5, use_flash_attn=True,
               max_position_embeddings=64)
    torch.manual_seed(0)
    return initialize_internlm_model("INTERNLM2_PUBLIC", None, cfg, param_dtype=torch.float32)


def test_generate_matches_full_forward_and_beam():
    from internevo_b200.apis.inference import SequenceGenerator

    model = _tiny_model()
    prompt = torch.tensor([[1, 5, 9, 13], [1, 7, 11, 15]])
    gen = SequenceGenerator(model, eos_token_id=None, pad_token_id=0, bos_token_id=1)
    out = gen.generate(prompt, max_length=12, do_sample=False)
    assert out.shape == (2, 1, 12)
    # reference: recompute the whole sequence at every step without any cache (packed training forward)
    seq = prompt.clone()
    for _ in range(8):
        rows = []
        for b in range(seq.shape[0]):
            T = seq.shape[1]
            logits = model(input_ids=seq[b:b + 1], cu_seqlens=torch.tensor([0, T], dtype=torch.int32),
                           indexes=torch.arange(T)[None])
            logits = logits[0] if isinstance(logits, (tuple, list)) else logits
            rows.append(logits.reshape(T, -1)[-1].argmax())
        seq = torch.cat([seq, torch.stack(rows)[:, None]], 1)
    assert torch.equal(out[:, 0], seq), (out[:, 0], seq)

    beams = gen.generate(prompt[:1], max_length=10, num_beams=3, num_return_sequences=2, do_sample=False)
    assert beams.shape[0] == 1 and beams.shape[1] == 2 and beams.shape[2] <= 10
    assert torch.equal(beams[0, 0, :4], prompt[0])

    stream = list(gen.streaming_generate(prompt[:1], max_length=8, do_sample=False))
    assert len(stream) == 4 and torch.equal(stream[-1][0, 0], out[0, 0, :8])


def _generate_with_layout(rank, world, kw):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_parallel_cpu as T
    from common import build_trainer, tiny_config

    from internevo_b200.apis.inference import SequenceGenerator

    cfg = tiny_config(**kw)
    trainer, opt, model, _ = build_trainer(cfg)
    T._load_golden(model, opt, cfg)            # identical full weights in every layout
    model.eval()
    gen = SequenceGenerator(model, eos_token_id=None, pad_token_id=0, bos_token_id=1)
    prompt = torch.tensor([[1, 5, 9, 13], [1, 7, 11, 15]])
    greedy = gen.generate(prompt, max_length=12, do_sample=False)[:, 0].tolist()
    beams = gen.generate(prompt[:1], max_length=10, num_beams=3, num_return_sequences=2, do_sample=False)[0].tolist()
    return greedy, beams


def test_generation_under_tensor_parallel_matches_single_rank():
    """tp = 2 (mtp): vocabulary-parallel logits are gathered, KV caches hold the local heads - same tokens as tp = 1, for
    greedy and beam search; sequence-parallel tensor modes refuse generation with a clear error."""
    import pytest

    from common import run_distributed

    single = run_distributed(_generate_with_layout, 1, dict())[0]
    for res in run_distributed(_generate_with_layout, 2, dict(tp=2)):
        assert res == single
    with pytest.raises(AssertionError, match="needs parallel.tensor.mode='mtp'"):
        run_distributed(_generate_with_layout, 2, dict(tp=2, mode="msp"))
