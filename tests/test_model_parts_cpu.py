"""Model-part and fault tests mirroring the reference's `tests/test_model/*`, `test_fused_precision`, `test_norm_weight` and
`test_timeout`, on CPU / gloo with 2 ranks: hidden-split embedding, gather / split autograd pairs, reward head,
fp32-tagged modules under NaiveAMPModel + fp32 optimizer group, replica (norm) weights staying identical across the tensor
group after training, and a hung collective being turned into an error by the process-group timeout."""
import time

import pytest
import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config


# ------------------------------------------------------------------------------------------ embedding / gather / head
def _parts(rank, world):
    import torch.distributed as dist

    from internevo_b200.core.context import ParallelMode
    from internevo_b200.core.context import global_context as gpc
    from internevo_b200.initialize import initialize_distributed_env
    from internevo_b200.models.modules import Embedding1D
    from internevo_b200.parallel.functional import gather_forward_split_backward, split_forward_gather_backward
    from internevo_b200.parallel.linear import RewardModelLinear

    initialize_distributed_env(config=tiny_config(tp=2), launcher="torch", seed=7)
    group = gpc.get_group(ParallelMode.TENSOR)
    out = {}
    # gather_forward_split_backward: forward concatenates the shards, backward hands every rank its own slice
    x = (torch.arange(6, dtype=torch.float32).view(2, 3) + 10 * rank).requires_grad_(True)
    y = gather_forward_split_backward(x, group, dim=-1)
    w = torch.arange(12, dtype=torch.float32).view(2, 6)
    (y * w).sum().backward()
    out["gather"] = (y.detach().clone(), x.grad.clone())
    # split_forward_gather_backward is its transpose
    z = torch.arange(8, dtype=torch.float32).view(4, 2).requires_grad_(True)
    s = split_forward_gather_backward(z, group, dim=0)
    (s * (rank + 1)).sum().backward()
    out["split"] = (s.detach().clone(), z.grad.clone())
    # Embedding1D: hidden dimension sharded, output = full hidden
    torch.manual_seed(3)
    full = torch.randn(16, 8)
    emb = Embedding1D(16, 8)
    with torch.no_grad():
        emb.weight.copy_(full[:, rank * 4:(rank + 1) * 4])
    ids = torch.tensor([1, 5, 5, 9])
    e = emb(ids)
    g = torch.randn(4, 8, generator=torch.Generator().manual_seed(5))
    (e * g).sum().backward()
    ref_w = torch.zeros(16, 8).index_add_(0, ids, g)
    out["embed"] = (torch.allclose(e, full[ids]), torch.allclose(emb.weight.grad, ref_w[:, rank * 4:(rank + 1) * 4]))
    # reward head: replicated weights (rank 0's init is broadcast), scalar output
    head = RewardModelLinear(8, 1, process_group=group, bias=True)
    ws = [torch.empty_like(head.weight) for _ in range(world)]
    dist.all_gather(ws, head.weight.data, group=group)
    out["reward"] = (bool(torch.equal(ws[0], ws[1])), tuple(head(torch.randn(3, 8)).shape))
    return out


def test_embedding_gather_split_and_reward_head_tp2():
    res = run_distributed(_parts, 2)
    for rank, r in enumerate(res):
        y, gx = r["gather"]
        want_y = torch.cat([torch.arange(6.).view(2, 3), torch.arange(6.).view(2, 3) + 10], -1)
        assert torch.equal(y, want_y)
        assert torch.equal(gx, torch.arange(12.).view(2, 6)[:, rank * 3:(rank + 1) * 3])
        s, gz = r["split"]
        assert torch.equal(s, torch.arange(8.).view(4, 2)[rank * 2:(rank + 1) * 2])
        assert torch.equal(gz, torch.tensor([1., 1., 2., 2.]).repeat_interleave(2).view(4, 2))
        assert r["embed"] == (True, True)
        assert r["reward"] == (True, (3, 1))


# ------------------------------------------------------------------------------------------ fp32 modules / fp32 group
def _fp32_parts(rank, world):
    cfg = tiny_config(dtype="torch.bfloat16", num_layers=2)
    cfg["use_fp32_norm"] = True
    trainer, opt, model, _ = build_trainer(cfg)
    inner = model.model
    from internevo_b200.ops.norm import RMSNorm

    norm_dtypes = {p.dtype for m in inner.modules() if isinstance(m, RMSNorm) for p in m.parameters()}
    other = {p.dtype for n, p in inner.named_parameters() if "norm" not in n}
    groups = {g.name: sorted({p.dtype for p in g.params}, key=str) for g in opt.groups if g.params}
    data, labels = synthetic_batch(2, 64, 128, seed=0)
    losses = []
    for _ in range(3):
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok
        losses.append(float(out[2]))
    return norm_dtypes, other, groups, losses


def test_fp32_tagged_norms_stay_fp32_and_get_their_own_optimizer_group():
    norm_dtypes, other, groups, losses = run_distributed(_fp32_parts, 1)[0]
    assert norm_dtypes == {torch.float32} and other == {torch.bfloat16}
    assert groups.get("fp32") == [torch.float32] and groups.get("default") == [torch.bfloat16]
    assert losses[-1] < losses[0]


# ------------------------------------------------------------------------------------------ replica weights across TP
def _norm_weights_after_training(rank, world, mode):
    import torch.distributed as dist

    from internevo_b200.core.context import ParallelMode
    from internevo_b200.core.context import global_context as gpc

    trainer, opt, model, _ = build_trainer(tiny_config(tp=2, mode=mode, num_layers=2))
    for step in range(4):
        data, labels = synthetic_batch(2, 64, 128, seed=step)
        trainer.zero_grad()
        trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok
    flat = torch.cat([p.detach().float().reshape(-1) for n, p in model.model.named_parameters() if "norm" in n])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat, group=gpc.get_group(ParallelMode.TENSOR))
    return bool(torch.equal(both[0], both[1])), float((flat - 1).abs().max())


@pytest.mark.parametrize("mode", ["msp", "fsp", "isp"])
def test_norm_weights_identical_across_tensor_ranks_after_training(mode):
    for same, moved in run_distributed(_norm_weights_after_training, 2, mode):
        assert same          # replica parameters never drift apart (reference tests/test_training/test_norm_weight.py)
        assert moved > 0     # ... and they did train


# ------------------------------------------------------------------------------------------ hung collective → error
def _hang(rank, world):
    import datetime

    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=4))
    dist.barrier()
    if rank == 1:
        time.sleep(12)           # the "hung" rank never joins the collective in time
        return "slept"
    t0 = time.time()
    try:
        dist.all_reduce(torch.ones(4))
    except Exception as e:  # noqa: BLE001 - gloo raises RuntimeError / DistBackendError depending on the build
        return ("raised", time.time() - t0, type(e).__name__)
    return ("no error", time.time() - t0, "")


def test_hung_collective_raises_within_the_group_timeout():
    """Reference `tests/test_utils/test_timeout.py`: rank 1 sleeps, rank 0's all-reduce must fail by timeout, not hang."""
    res = run_distributed(_hang, 2, timeout=60)
    status, elapsed, _ = res[0]
    assert status == "raised" and elapsed < 11, res


def _knobs(rank, world, extra):
    """One training step with a model-config knob switched on; returns the loss and what the knob must have changed."""
    cfg = tiny_config(num_layers=2, dtype="torch.bfloat16", **extra)
    trainer, opt, model, _ = build_trainer(cfg)
    seen = {}

    def grab(mod, args, kwargs=None):
        if len(args) > 1 and torch.is_tensor(args[1]):
            seen["residual_dtype"] = str(args[1].dtype)

    layers = [m for m in model.modules() if type(m).__name__ == "DecoderLayer"]
    layers[1].register_forward_pre_hook(grab)
    data, labels = synthetic_batch(2, 64, cfg["model"]["vocab_size"], seed=3)
    losses = []
    for _ in range(2):
        trainer.zero_grad()
        out = trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok
        losses.append(float(out[2]))
    return losses, seen


def test_residual_in_fp32_keeps_an_fp32_residual_stream():
    (base, seen0), = run_distributed(_knobs, 1, dict())
    (fp32, seen1), = run_distributed(_knobs, 1, dict(residual_in_fp32=True))
    assert seen0["residual_dtype"] == "torch.bfloat16" and seen1["residual_dtype"] == "torch.float32"
    assert all(l == l for l in fp32) and abs(fp32[0] - base[0]) < 0.05


def test_attention_dropout_is_applied_in_training():
    (base, _), = run_distributed(_knobs, 1, dict())
    (drop, _), = run_distributed(_knobs, 1, dict(attn_drop_rate=0.5))
    assert all(l == l for l in drop)
    assert abs(drop[0] - base[0]) > 1e-4, "attn_drop_rate had no effect on the training loss"


def test_output_tf32_keeps_the_head_in_fp32():
    """``output_tf32 = True``: the LM head's weights stay fp32, it is fed fp32 activations and returns fp32 logits while the
    rest of the model runs in the low-precision dtype (reference ``core/naive_amp.py:203-208``)."""
    from internevo_b200.core.context import Config, global_context as gpc
    from internevo_b200.core.naive_amp import NaiveAMPModel
    from internevo_b200.parallel.linear import ScaleColumnParallelLinear

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(8, 8)
            self.output = ScaleColumnParallelLinear(8, 16, process_group=None, bias=False)

        def forward(self, x):
            return self.output(self.body(x))

    gpc.set_config(Config(dict(output_tf32=True, parallel=dict(tensor=dict(size=1, mode="mtp")))))
    amp = NaiveAMPModel(Tiny(), output_to_fp32=False, dtype=torch.bfloat16)
    assert amp.model.body.weight.dtype == torch.bfloat16 and amp.model.output.weight.dtype == torch.float32
    y = amp(torch.randn(4, 8))
    assert y.dtype == torch.float32 and y.shape == (4, 16)
    y.sum().backward()
    assert amp.model.output.weight.grad.dtype == torch.float32 and amp.model.body.weight.grad is not None
    gpc.set_config(Config(dict(output_tf32=False, parallel=dict(tensor=dict(size=1, mode="mtp")))))
    amp = NaiveAMPModel(Tiny(), output_to_fp32=False, dtype=torch.bfloat16)
    assert amp.model.output.weight.dtype == torch.bfloat16
