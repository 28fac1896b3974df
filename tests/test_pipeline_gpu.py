"""Pipeline parallelism on GPUs (NCCL p2p + the native kernels): 1F1B and interleaved 1F1B over 2 stages must follow the
single-GPU run of the same model and data (the CPU suite checks the same schedules on gloo in fp32)."""
import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _train(rank, world, kw):
    cfg = tiny_config(dtype="torch.bfloat16", num_layers=4, hidden=512, heads=4, kv_heads=2, seq_len=512, micro_bsz=1,
                      vocab=1024, micro_num=4, **kw)
    trainer, opt, model, _ = build_trainer(cfg)
    out = []
    for step in range(4):
        data, labels = synthetic_batch(4, 512, 1024, seed=step)
        trainer.zero_grad()
        res = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        loss = res[2]
        out.append((float(loss) if loss is not None else None, float(sum(v * v for v in norms.values()) ** 0.5)))
    return out


@pytest.mark.parametrize("chunks", [1, 2])
def test_pp2_training_tracks_single_gpu(chunks):
    ref = run_distributed(_train, 1, {})[0]
    got = run_distributed(_train, 2, dict(pp=2, num_chunks=chunks))
    last = [r for r in got if r[0][0] is not None]          # the last stage reports the loss
    assert last, got
    for (l0, n0), (l1, n1) in zip(ref, last[0]):
        assert abs(l0 - l1) < 0.02 * abs(l0) + 0.02, (ref, last[0])
