"""Peer-memory kernels on >= 2 GPUs of one node: device barrier, GEMM-epilogue reduce-scatter / all-reduce, TMA-pushed
all-gather + GEMM, fused Hybrid-ZeRO reduce-scatter + AdamW + parameter push — all checked against NCCL + the plain GEMM
inside ``tools/fused_comm_check.py``; and a tensor-parallel training run with the fused linears must follow the NCCL run."""
import json
import os
import subprocess
import sys

import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs with NVLink")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_comm_kernels_match_nccl(tmp_path):
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", "tools/fused_comm_check.py"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.load(open(os.path.join(ROOT, "gpurun_out", f"fused_comm_check_n{n}.json")))
    assert res["all_ok"], res


def _train(rank, world, fused, mode):
    cfg = tiny_config(tp=2, mode=mode, dtype="torch.bfloat16", num_layers=2, hidden=512, heads=4, kv_heads=2, seq_len=512,
                      micro_bsz=1, vocab=1024, micro_num=2)
    cfg["fused_comm"] = fused
    os.environ["B200_TP_FUSED"] = "1" if fused else "0"     # default on; "0" = 2-CTA GEMM + NCCL (the oracle)
    trainer, opt, model, _ = build_trainer(cfg)
    from internevo_b200.parallel import linear

    assert (linear._fused_backend is not None) == bool(fused)
    out_l = []
    for _ in range(4):
        data, labels = synthetic_batch(2, 512, 1024, seed=0)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        out_l.append((float(out[2]), float(list(norms.values())[0])))
    return out_l


@pytest.mark.parametrize("mode", ["mtp", "msp", "fsp"])
def test_tp2_fused_linears_track_nccl(mode):
    ref = run_distributed(_train, 2, False, mode)[0]
    got = run_distributed(_train, 2, True, mode)[0]
    for (l0, n0), (l1, n1) in zip(ref, got):
        assert abs(l0 - l1) < 0.02 * abs(l0) + 0.01, (ref, got)
        assert abs(n0 - n1) < 0.08 * n0 + 0.02, (ref, got)


def _train_isp(rank, world, fused):
    """Weight-parallel (ISP) run with sequence-parallel attention: sp = wp = 2."""
    cfg = tiny_config(tp=2, wp=2, mode="isp", dtype="torch.bfloat16", num_layers=2, hidden=512, heads=4, kv_heads=2,
                      seq_len=1024, micro_bsz=1, vocab=1024, micro_num=2)
    cfg["fused_comm"] = True
    os.environ["B200_ISP_FUSED"] = "1" if fused else "0"
    trainer, opt, model, _ = build_trainer(cfg)
    from internevo_b200 import ops

    n0 = ops.launch_count()
    out_l = []
    for _ in range(4):
        data, labels = synthetic_batch(2, 1024, 1024, seed=0)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        out_l.append((float(out[2]), sorted(norms.items())))
    from internevo_b200.parallel import fused as fmod

    used = bool(fmod._isp_backends) and any(be._gath for be in fmod._isp_backends.values())
    assert used == bool(fused), "the fused ISP path was (not) taken"
    return out_l


def test_isp_fused_linears_track_nccl():
    """ISP with the weight all-gather / gradient reduce-scatter inside the GEMM kernels must follow the NCCL-prefetch run."""
    ref = run_distributed(_train_isp, 2, False)[0]
    got = run_distributed(_train_isp, 2, True)[0]
    for (l0, n0), (l1, n1) in zip(ref, got):
        assert abs(l0 - l1) < 0.02 * abs(l0) + 0.01, (ref, got)
        for (k0, v0), (k1, v1) in zip(n0, n1):
            assert k0 == k1 and abs(v0 - v1) < 0.08 * v0 + 0.02, (ref, got)


def test_sequence_parallel_attention_with_peer_kv_matches_single_rank_kernel():
    """Forward and backward of the sequence-parallel attention kernels (K / V read from the peers' symmetric buffers inside
    the kernel, dQ reduce-added into its owner over NVLink) against the single-rank kernel on the full sequence: one long
    sequence, ragged packed sequences crossing rank boundaries at non-aligned rows, many short sequences."""
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29534", "tools/sp_attn_check.py"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.load(open(os.path.join(ROOT, "gpurun_out", f"sp_attn_check_n{n}.json")))
    assert res["all_ok"], res
