"""Training-step time of an InternLM-MoE stack with the expert all-to-all fused over peer memory vs the NCCL formulation.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/moe_bench.py [--moe_type GShard|MegaBlock-D] [--layers 4]

Same model, data and kernels in both arms; only `B200_MOE_FUSED` differs (a fresh process group per arm would cost a second
model build, so the arms run back to back in one process and the switch is read per forward).  Expert parallel size = world
size, 7B-class layer shapes (hidden 4096, 32 heads, seq 4096), `--layers` MoE blocks, 2 experts per GPU, top-2 routing.
Device-timed with CUDA events, max over ranks.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--moe_type", default="GShard")
    p.add_argument("--layers", type=int, default=4)
    p.add_argument("--steps", type=int, default=6)
    p.add_argument("--warmup", type=int, default=3)
    a = p.parse_args()
    import torch
    import torch.distributed as dist

    from common import build_trainer, synthetic_batch, tiny_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=a.layers, micro_num=2, num_experts=2 * world, moe_type=a.moe_type,
                      dtype="torch.bfloat16", hidden=4096, heads=32, seq_len=4096, micro_bsz=1, vocab=32000)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["model"]["mlp_ratio"] = 3.5
    cfg["moe"] = dict(top_k=2)
    cfg["loss"]["moe_loss_coeff"] = 0.01
    trainer, opt, model, _ = build_trainer(cfg)
    rank = dist.get_rank()
    data, labels = synthetic_batch(2, 4096, 32000, seed=rank)

    def step():
        trainer.zero_grad()
        trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok

    res = {"world": world, "moe_type": a.moe_type, "layers": a.layers, "experts": 2 * world, "top_k": 2, "hidden": 4096,
           "tokens_per_step_per_gpu": 2 * 4096}
    for name, flag in (("nccl_all_to_all", "0"), ("fused_peer_memory", "1"), ("nccl_all_to_all_again", "0")):
        os.environ["B200_MOE_FUSED"] = flag
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.steps):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / a.steps], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[name + "_ms_per_step"] = round(float(ms), 2)
    if rank == 0:
        print(json.dumps(res), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"moe_bench_{a.moe_type}_n{world}.json"), "w"), indent=1)
    from internevo_b200.core.context import global_context as gpc

    gpc.destroy()


if __name__ == "__main__":
    main()
