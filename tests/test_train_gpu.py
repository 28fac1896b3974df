"""End-to-end on one B200: a small InternLM2 trains through the public API with the hand-written kernels on the path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_smoke_entry():
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "smoke ok" in r.stdout


def test_bench_debug_config_runs():
    r = subprocess.run([sys.executable, "bench.py", "--layers", "2", "--steps", "2", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json

    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["gpu_launches"] > 0 and res["value"] > 0 and res["e2e"]["value"] > 0


def _train_overlap(rank, world, overlap):
    """6 optimizer steps of a small bf16 model; returns losses, grad norms and a parameter checksum."""
    import torch

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import build_trainer, synthetic_batch, tiny_config

    os.environ["B200_ADAM_OVERLAP"] = "1" if overlap else "0"
    cfg = tiny_config(dtype="torch.bfloat16", num_layers=4, hidden=512, heads=4, kv_heads=2, seq_len=512, micro_bsz=1,
                      vocab=1024, micro_num=2)
    trainer, opt, model, _ = build_trainer(cfg)
    assert opt._model_attached == bool(overlap)
    out = []
    for step in range(6):
        data, labels = synthetic_batch(2, 512, 1024, seed=step % 2)
        trainer.zero_grad()
        res = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        out.append((float(res[2]), float(list(norms.values())[0])))
    opt.flush_param_update()
    torch.cuda.synchronize()
    checksum = float(sum(p.detach().float().abs().sum() for p in model.parameters()))
    return out, checksum


def test_adam_forward_overlap_matches_the_serial_update():
    """The update of step s runs on a side stream under the forward of step s + 1 (per-module event waits): same kernels on
    the same data -> the same training trajectory (up to the run-to-run noise of the atomic grad-norm reduction, ~1e-7)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import run_distributed

    serial = run_distributed(_train_overlap, 1, False)[0]
    overlapped = run_distributed(_train_overlap, 1, True)[0]
    for (l0, n0), (l1, n1) in zip(serial[0], overlapped[0]):
        assert abs(l0 - l1) < 2e-3 * abs(l0) and abs(n0 - n1) < 5e-3 * abs(n0), (serial[0], overlapped[0])
    assert abs(serial[1] - overlapped[1]) < 1e-4 * abs(serial[1])
    assert serial[0][-1][0] < serial[0][0][0]
