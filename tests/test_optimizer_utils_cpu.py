"""List-based optimizer utilities kept next to the arena optimizer: ``compute_norm`` / ``reduce_grads`` (reference
``solver/optimizer/utils.py:225-378``) must agree with the norm the arena optimizer computes inside its reduce path, and the
bucket / parameter / gradient stores (``solver/optimizer/store.py``) describe a live optimizer."""
import math

import torch

from common import build_trainer, run_distributed, synthetic_batch, tiny_config


def _norm_worker(rank, world, mode):
    from internevo_b200.core.context import ParallelMode, global_context as gpc
    from internevo_b200.solver.optimizer.store import stores_of
    from internevo_b200.solver.optimizer.utils import compute_norm, reduce_grads

    cfg = tiny_config(tp=2, mode=mode, wp=2 if mode == "isp" else 1)
    trainer, opt, model, _ = build_trainer(cfg)
    T = cfg["data"]["seq_len"] * cfg["data"]["micro_bsz"]
    data, labels = synthetic_batch(2, T, cfg["model"]["vocab_size"], seed=3)
    trainer.zero_grad()
    trainer.execute_schedule((data, labels), forward_only=False, return_loss=True, return_output_label=False)
    listed = {}
    for g in opt.groups:
        if not g.params:
            continue
        grads = [p.grad_buf.detach().clone() for p in g.params]
        listed[g.name] = math.sqrt(compute_norm(grads, g.params, norm_type=2, zero_mode=g.zero_mode))
        model_mode = ParallelMode.WEIGHT if mode == "isp" else ParallelMode.TENSOR
        kept = reduce_grads(grads, g.params, model_mode)
        n_rep = sum(1 for p in g.params if getattr(p, "is_replica_zero_parallel", False))
        assert len(kept) == len(grads) - (n_rep if gpc.get_local_rank(model_mode) != 0 else 0)
    if mode == "isp":       # the communicator's view of the chunk: blocks, weight-parallel linears per block, the head outside
        from internevo_b200.core.communication.isp import ISPOverlapState
        from internevo_b200.parallel.linear import ISPLinear

        comm = ISPLinear.communicator()
        st = comm.overlap_state
        assert isinstance(st, ISPOverlapState) and st.num_blocks == cfg["model"]["num_layers"] and st.ckpt_block_num == 0
        per_block = {len(v) for v in st.index_to_isp_modules.values()}
        assert len(per_block) == 1 and per_block.pop() >= 3, st.index_to_isp_modules        # wqkv, wo, w13 (/ w1, w3), w2
        assert all(st.block_of(m) == i for i, mods in st.index_to_isp_modules.items() for m in mods)
        assert len(st.isp_modules) == sum(len(v) for v in st.index_to_isp_modules.values()) + len(st.isp_outs)
        assert st.in_flight() == 0                     # nothing left gathered after forward + backward
        comm.switch_current_model_chunk(0)
    pstore, gstore, buckets = stores_of(opt)
    g0 = next(g for g in opt.groups if g.params)
    assert pstore.get_param_rank(g0.params[0]) == 0 and len(buckets) == sum(1 for g in opt.groups if g.params)
    assert pstore.get_flat_fp16_param_by_rank_group(pstore.local_rank, g0.gid) is g0.param_arena
    assert gstore.get_averaged_gradients_by_group(g0.gid)[0].numel() == g0.shard
    ok, norms = trainer.step()
    assert ok
    return listed, {k: float(v) for k, v in norms.items()}


def _check_norm(mode):
    for listed, fused in run_distributed(_norm_worker, 2, mode):
        assert set(listed) == set(fused), (listed, fused)
        for name in listed:
            assert abs(listed[name] - fused[name]) <= 1e-4 * max(1.0, fused[name]), (mode, name, listed, fused)


def test_list_based_norm_equals_the_arena_norm_tensor_parallel():
    _check_norm("mtp")


def test_list_based_norm_equals_the_arena_norm_sequence_and_weight_parallel():
    _check_norm("fsp")
    _check_norm("isp")


def test_compute_norm_sentinels_and_inf_norm_single_process():
    from internevo_b200.solver.optimizer.utils import compute_norm, get_grad_accumulate_object

    p = [torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(2))]
    g = [torch.tensor([3.0, 0.0, 4.0]), torch.tensor([0.0, 12.0])]
    assert abs(compute_norm(g, p, 2) - 169.0) < 1e-6
    assert compute_norm(g, p, math.inf) == 12.0
    assert compute_norm([torch.tensor([float("inf")])], p[:1], 2) == -1
    assert compute_norm([torch.tensor([float("nan")])], p[:1], 2) == -2
    acc = get_grad_accumulate_object(p[0])
    fired = []
    acc.register_hook(lambda *a: fired.append(1))
    (p[0] * 2).sum().backward()
    assert fired == [1] and torch.equal(p[0].grad, torch.full((3,), 2.0))


def test_tensor_bucket_and_stores():
    from internevo_b200.solver.optimizer.store import BucketStore, GradientStore, ParameterStore, TensorBucket

    b = TensorBucket(size=10)
    t1, t2 = torch.arange(4.0), torch.arange(5.0)
    b.add_to_bucket(t1)
    b.add_to_bucket(t2)
    assert b.current_size == 9 and not b.is_full_or_oversized() and b.will_exceed_max_size(2)
    try:
        b.add_to_bucket(torch.zeros(2))
        raise AssertionError("must refuse")
    except RuntimeError:
        pass
    b.add_to_bucket(torch.zeros(2), allow_oversize=True)
    assert b.is_full_or_oversized()
    flat = b.flatten()
    flat.mul_(2)                                    # stands for the collective's result
    b.set_unflatten_and_copy_flag(True)
    b.unflatten_and_copy()
    assert torch.equal(t1, torch.arange(4.0) * 2) and torch.equal(t2, torch.arange(5.0) * 2)
    b.empty()
    assert b.is_empty() and b.current_size == 0 and b.get_flat_tensor() is None

    from internevo_b200.core.context import ParallelMode

    bs = BucketStore(0, ParallelMode.DATA)
    bs.add_grad(t1, reduce_rank=None)
    bs.add_param(t1, reduce_rank=None)
    bs.add_num_elements_in_bucket(4)
    assert bs.num_elements_in_bucket() == 4 and bs.num_params_in_bucket() == 1 and bs.get_grad()[0] is t1
    bs.reset_by_rank(None)
    assert bs.num_elements_in_bucket() == 0 and bs.get_param() == []
    gs = GradientStore(ParallelMode.DATA)
    gs.add_average_gradient_by_group(1, t2)
    assert gs.get_averaged_gradients_by_group(1)[0] is t2
    gs.reset_average_gradients_by_group(1)
    assert gs.get_averaged_gradients_by_group(1) == []
    ps = ParameterStore(ParallelMode.ZERO1)
    w = torch.nn.Parameter(torch.zeros(2))
    ps.set_param_to_rank(w, 0)
    ps.set_param_reduction_state(w, True)
    assert ps.belongs_to_current_rank(w) and ps.is_param_reduced(w) and list(ps.get_param_reduction_states().values()) == [True]
    w.grad = torch.ones(2)
    ps.add_previous_reduced_param(w)
    ps.clear_grads_of_previous_reduced_params()
    assert w.grad is None
