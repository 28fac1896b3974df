"""parallel.zero1.fsdp=True (ZeRO-3 through torch FSDP + FSDPadaptOptimizer) must follow the same loss / grad-norm
trajectory as Hybrid-ZeRO data parallelism from the same seed-initialised weights."""
import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config

STEPS = 3


def _train(rank, world, fsdp):
    from internevo_b200.core.context import ParallelMode, global_context as gpc

    cfg = tiny_config(zero1=world, fsdp=fsdp, micro_num=2, num_layers=2, dtype="torch.bfloat16", hidden=256, heads=2,
                      kv_heads=2, seq_len=128)
    trainer, opt, model, _ = build_trainer(cfg)
    if fsdp:
        from torch.distributed.fsdp import FullyShardedDataParallel as FSDP

        from internevo_b200.solver.optimizer import FSDPadaptOptimizer

        assert isinstance(model, FSDP) and isinstance(opt, FSDPadaptOptimizer)
    dpr = gpc.get_local_rank(ParallelMode.DATA)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    out_l = []
    for _ in range(STEPS):
        data, labels = synthetic_batch(4, T, cfg["model"]["vocab_size"], seed=0)
        data = {k: v[dpr * 2:(dpr + 1) * 2] for k, v in data.items()}
        labels = labels[dpr * 2:(dpr + 1) * 2]
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        loss = out[2].detach().clone().reshape(1).float()
        torch.distributed.all_reduce(loss, group=gpc.get_group(ParallelMode.DATA))
        out_l.append((float(loss) / world, float(list(norms.values())[0])))
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    return out_l


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="torch FSDP needs accelerators; 2 GPUs")
def test_fsdp_matches_zero_dp2():
    ref = run_distributed(_train, 2, False)[0]
    got = run_distributed(_train, 2, True)[0]
    for (l0, n0), (l1, n1) in zip(ref, got):
        assert abs(l0 - l1) < 3e-2 * max(1.0, abs(l0)), (ref, got)
        assert abs(n0 - n1) < 5e-2 * max(1.0, n0), (ref, got)
    assert got[-1][0] < got[0][0]
