"""Expert-parallel dispatch / combine over peer memory (``csrc/moe_comm.cu``) vs the NCCL all-to-all path of the same
layer: outputs, input gradients, gate gradients and expert weight gradients must agree (2 GPUs, bf16)."""
import os

import pytest
import torch

from common import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs with NVLink")]


def _layer_run(rank, world, fused, kind="dropless"):
    import torch.distributed as dist

    from internevo_b200 import ops
    from internevo_b200.models.modules import FeedForward
    from internevo_b200.models.moe import DroplessMOELayer, Experts, GShardMOELayer, TopKGate

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    os.environ["B200_MOE_FUSED"] = "1" if fused else "0"
    h, E, S, k = 512, 4, 1000, 2
    El = E // world
    torch.manual_seed(7)
    all_experts = [FeedForward(h, 1024, out_features=h, process_group=None, bias=False, device="cuda", dtype=torch.bfloat16)
                   for _ in range(E)]
    mine = all_experts[rank * El:(rank + 1) * El]
    if kind == "dropless":
        layer = DroplessMOELayer(h, E, dist.group.WORLD, world, Experts(mine, El, f"moe_ep_size_{world}"), top_k=k,
                                 device="cuda")
    else:  # GShard top-2 with capacity 1.0: tokens beyond capacity are dropped (slot row -1 in the fused path)
        gate = TopKGate(h, E, k, 1.0, 1.0, 4, None, True, True, device="cuda")
        layer = GShardMOELayer(h, gate, Experts(mine, El, f"moe_ep_size_{world}"), dist.group.WORLD, world, El)
    torch.manual_seed(100 + rank)
    x = (torch.randn(S, h, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    n0 = ops.launch_count()
    y = layer(x)
    gy = torch.randn(S, h, device="cuda").to(torch.bfloat16)
    (y.float() * gy.float()).sum().backward()
    torch.cuda.synchronize()
    launches = ops.launch_count() - n0
    wg = layer.wg if kind == "dropless" else layer.gate.wg
    res = dict(y=y.detach().float().cpu(), gx=x.grad.float().cpu(), gwg=wg.weight.grad.float().cpu(),
               gexp=[p.grad.float().cpu() if p.grad is not None else getattr(p, "grad_buf", torch.zeros(1)).float().cpu()
                     for e in mine for p in e.parameters()],
               launches=launches)
    dist.barrier()
    dist.destroy_process_group()
    return res


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("kind", ["dropless", "gshard"])
def test_fused_moe_dispatch_combine_matches_nccl(kind):
    ref = run_distributed(_layer_run, 2, False, kind)
    got = run_distributed(_layer_run, 2, True, kind)
    for r, g in zip(ref, got):
        assert _rel(g["y"], r["y"]) < 2e-2, _rel(g["y"], r["y"])
        assert _rel(g["gx"], r["gx"]) < 2e-2, _rel(g["gx"], r["gx"])
        assert _rel(g["gwg"], r["gwg"]) < 3e-2, _rel(g["gwg"], r["gwg"])
        for a, b in zip(g["gexp"], r["gexp"]):
            assert _rel(a, b) < 3e-2, _rel(a, b)
        assert g["launches"] > r["launches"]       # the peer-memory kernels really ran


def _train_moe(rank, world, fused, moe_type):
    """A tiny InternLM-MoE model trained through the public API (expert parallel over all ranks, bf16, native kernels)."""
    from common import build_trainer, synthetic_batch, tiny_config

    os.environ["B200_MOE_FUSED"] = "1" if fused else "0"
    cfg = tiny_config(model_type="INTERNLM_MoE", num_layers=2, micro_num=2, num_experts=2 * world, moe_type=moe_type,
                      dtype="torch.bfloat16", hidden=512, heads=4, seq_len=512, micro_bsz=1, vocab=1024)
    cfg["model"].pop("no_bias", None)
    cfg["model"].pop("num_kv_attention_heads", None)
    cfg["moe"] = dict(top_k=2)
    cfg["loss"]["moe_loss_coeff"] = 0.1
    trainer, opt, model, _ = build_trainer(cfg)
    losses = []
    for step in range(5):
        data, labels = synthetic_batch(2, 512, 1024, seed=rank)
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, norms = trainer.step()
        assert ok
        losses.append((float(out[2]), float(list(norms.values())[0])))
    return losses


@pytest.mark.parametrize("moe_type", ["MegaBlock-D", "GShard"])
def test_moe_training_with_fused_dispatch_tracks_nccl(moe_type):
    world = int(os.environ.get("B200_TEST_WORLD", "2"))
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ref = run_distributed(_train_moe, world, False, moe_type)
    got = run_distributed(_train_moe, world, True, moe_type)
    for r, g in zip(ref, got):
        assert g[-1][0] < g[0][0], g                                   # it learns
        for (l0, n0), (l1, n1) in zip(r, g):
            assert abs(l0 - l1) < 0.03 * abs(l0) + 0.02, (r, g)        # and follows the NCCL all-to-all run
            assert abs(n0 - n1) < 0.15 * n0 + 0.05, (r, g)
