"""Loss-curve regression on one B200 (the reference's ``tests/test_training/test_loss.py`` idea without its stored
baselines): the bf16 run through the hand-written kernels must track an fp32 plain-PyTorch run of the same model, same
weights and same data step by step, and so must the gradient norm."""
import os

import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config
from test_parallel_cpu import _load_golden

pytestmark = pytest.mark.gpu
STEPS = 6


def _train(rank, world, dtype):
    on_gpu = dtype != "torch.float32"
    assert torch.cuda.is_available() == on_gpu
    cfg = tiny_config(dtype=dtype, num_layers=2, hidden=512, heads=4, kv_heads=2, seq_len=256, micro_bsz=1, vocab=1024,
                      micro_num=2)
    trainer, opt, model, _ = build_trainer(cfg)
    _load_golden(model, opt, cfg)
    if on_gpu:
        from internevo_b200 import ops

        n0 = ops.launch_count()
    T = cfg["data"]["seq_len"]
    out_l = []
    for _ in range(STEPS):
        data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=0)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        out_l.append((float(out[2]), float(list(norms.values())[0])))
    if on_gpu:
        assert ops.launch_count() - n0 > 50 * STEPS, "the native kernels were not on the path"
    return out_l


def test_bf16_native_tracks_fp32_reference():
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    try:
        ref = run_distributed(_train, 1, "torch.float32")[0]
    finally:
        if saved is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES")
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved
    got = run_distributed(_train, 1, "torch.bfloat16")[0]
    assert ref[-1][0] < ref[0][0]
    for (l0, n0), (l1, n1) in zip(ref, got):
        assert abs(l0 - l1) < 0.03 * abs(l0) + 0.02, (ref, got)
        assert abs(n0 - n1) < 0.1 * n0 + 0.02, (ref, got)
