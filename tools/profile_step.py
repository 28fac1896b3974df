"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities only) for the bench configuration.

    python tools/profile_step.py [--layers N] > gpurun_out/step_profile.txt

Used to decide what to optimise next; numbers taken under the profiler are never reported as benchmark values.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    a = bench.parse()
    world, rank = bench.setup_env(a)
    cfg, tp, dp = bench.build_config(a, world)
    if a.attn:
        os.environ["INTERNEVO_ATTN_IMPL"] = a.attn
    import internevo_b200 as fw
    from internevo_b200.initialize import initialize_distributed_env
    from internevo_b200.models.losses import FlashGPTLMLoss
    from internevo_b200.train import get_scheduler_hooks, initialize_model, initialize_optimizer

    initialize_distributed_env(config=cfg, launcher="torch", seed=1024)
    model = initialize_model()
    crit = FlashGPTLMLoss(parallel_output=True, label_smoothing=0)
    opt, b2, lrs = initialize_optimizer(model)
    trainer, _, _, _ = fw.initialize_trainer(model=model, optimizer=opt, criterion=crit, lr_scheduler=lrs,
                                             beta2_scheduler=b2, scheduler_hooks=get_scheduler_hooks(None, opt, None))
    trainer.train()
    batch = bench.make_batches(1, a.micro_num, a.seq_len, bench.MODEL_7B["vocab_size"], pin=True)[0]

    def step():
        trainer.zero_grad()
        trainer.execute_schedule(batch, forward_only=False, return_loss=True, return_output_label=False)
        trainer.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))


if __name__ == "__main__":
    main()
