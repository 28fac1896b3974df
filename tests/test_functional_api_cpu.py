"""Function-level API kept for code written against the reference: ``fused_dense_func`` & co. (``internlm/model/utils.py``),
the MegaBlocks-style ``sdd`` / ``dsd`` products and expert MLPs (``internlm/model/moe/megablock``), the flash-attn style rotary
functions, ``filling_ops_queue``.  Every check is against a plain single-process PyTorch computation."""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from common import run_distributed


# ----------------------------------------------------------------------------------------------------- dense functions
def _dense_worker(rank, world):
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from internevo_b200.parallel.functional import (AllReduceFunc, ReduceScatterFunc, fused_dense_func,
                                                    linear_bias_wgrad_torch, megatron_fused_dense_func)

    group = dist.group.WORLD
    torch.manual_seed(0)
    T, h, n = 8, 6, 10
    x_full = torch.randn(T, h)
    w_full = torch.randn(n, h)
    b_full = torch.randn(n)
    g_full = torch.randn(T, n)
    out = {}
    for name, fn in (("fsp", fused_dense_func), ("msp", megatron_fused_dense_func)):
        # sequence parallel column linear: x sharded along tokens, weight along out features
        x = x_full.chunk(world, 0)[rank].clone().requires_grad_()
        w = w_full.chunk(world, 0)[rank].clone().requires_grad_()
        b = b_full.chunk(world, 0)[rank].clone().requires_grad_()
        y = fn(x, w, b, process_group=group, sequence_parallel=True)
        assert y.shape == (T, n // world)
        (y * g_full.chunk(world, 1)[rank]).sum().backward()
        out[name] = (y.detach(), x.grad, w.grad, b.grad)
        # 3-D input gathered along dim 1 (the reference's packed [1, T / tp, h] layout)
        y3 = fn(x.detach()[None], w.detach(), None, process_group=group, sequence_parallel=True, gather_dim=1)
        assert y3.shape == (1, T, n // world)
        assert torch.allclose(y3[0], F.linear(x_full, w.detach()), atol=1e-5)
    # no sequence parallel: dgrad is all-reduced
    x = x_full.clone().requires_grad_()
    w = w_full.chunk(world, 0)[rank].clone().requires_grad_()
    y, res = fused_dense_func(x, w, None, return_residual=True, process_group=group, sequence_parallel=False)
    assert res is x
    (y * g_full.chunk(world, 1)[rank]).sum().backward()
    out["mtp"] = (y.detach(), x.grad, w.grad)
    # reduce-scatter / all-reduce functions
    z = (torch.arange(T * 2, dtype=torch.float32).reshape(T, 2) * (rank + 1)).requires_grad_()
    rs = ReduceScatterFunc.apply(z, group, 0)
    rs.sum().backward()
    ar = AllReduceFunc.apply(z.detach(), group)
    dw, db = linear_bias_wgrad_torch(x_full, g_full, True)
    out["misc"] = (rs.detach(), z.grad, ar, dw, db)
    dist.destroy_process_group()
    return out


def test_fused_dense_functions_match_the_unsharded_linear():
    world = 2
    res = run_distributed(_dense_worker, world)
    torch.manual_seed(0)
    T, h, n = 8, 6, 10
    x = torch.randn(T, h, requires_grad=True)
    w = torch.randn(n, h, requires_grad=True)
    b = torch.randn(n, requires_grad=True)
    g = torch.randn(T, n)
    y = F.linear(x, w, b)
    (y * g).sum().backward()
    for mode in ("fsp", "msp"):
        ys, dxs, dws, dbs = zip(*[r[mode] for r in res])
        assert torch.allclose(torch.cat(ys, 1), y.detach(), atol=1e-5)
        assert torch.allclose(torch.cat(dxs, 0), x.grad, atol=1e-5)          # reduce-scattered dgrad
        assert torch.allclose(torch.cat(dws, 0), w.grad, atol=1e-5)
        assert torch.allclose(torch.cat(dbs, 0), b.grad, atol=1e-5)
    y0 = F.linear(x, w)
    x.grad = None
    (y0 * g).sum().backward()
    ys, dxs, dws = zip(*[r["mtp"] for r in res])
    assert torch.allclose(torch.cat(ys, 1), y0.detach(), atol=1e-5)
    for dx in dxs:
        assert torch.allclose(dx, x.grad, atol=1e-5)                          # all-reduced dgrad, same on every rank
    z = torch.arange(T * 2, dtype=torch.float32).reshape(T, 2)
    for r, (rs, dz, ar, dw, db) in enumerate(x["misc"] for x in res):
        assert torch.equal(rs, (z * 3).chunk(world, 0)[r]) and torch.equal(dz, torch.ones_like(z))
        assert torch.equal(ar, z * 3)
        assert torch.allclose(dw, g.t() @ x.detach(), atol=1e-5) and torch.allclose(db, g.sum(0), atol=1e-5)


# ----------------------------------------------------------------------------------------------------- megablock ops
def _moe_inputs(E=3, f=8, h=16, counts=(5, 0, 130)):
    from internevo_b200.models.megablock import Topology

    torch.manual_seed(1)
    topo = Topology.from_counts(torch.tensor(counts), f)
    R = int(topo.offsets[-1])
    topo.rows = R
    b = topo.offsets.tolist()
    x = torch.randn(R, h)
    keep = torch.zeros(R, dtype=torch.bool)
    for g in range(E):
        keep[b[g]: b[g] + counts[g]] = True
    x = x * keep[:, None]                                     # padding rows are zero by contract
    w = [torch.randn(E * f, h) * 0.3 for _ in range(3)]
    gy = torch.randn(R, h)
    return topo, b, x, w, gy


def _dense_experts(x, w1, w2, w3, b, E, f):
    ref = torch.zeros(x.shape[0], x.shape[1])
    for g in range(E):
        xs = x[b[g]: b[g + 1]]
        a, c, d = (t[g * f: (g + 1) * f] for t in (w1, w2, w3))
        ref[b[g]: b[g + 1]] = (F.silu(xs @ a.t()) * (xs @ d.t())) @ c
    return ref


def test_block_diagonal_products_and_grouped_mlp_single_process():
    from internevo_b200.models.megablock import MegaBlockFeedForward, MegaBlockGroupedFeedForward, act_fn, dsd_nn, sdd_nt

    E, f, h = 3, 8, 16
    topo, b, x, (w1, w2, w3), gy = _moe_inputs(E, f, h)
    s = sdd_nt(x, w1, topo, None, "tensor")
    dense = s.to_dense()
    full = x @ w1.t()                                         # only the diagonal blocks survive
    for g in range(E):
        assert torch.allclose(dense[b[g]: b[g + 1], g * f: (g + 1) * f], full[b[g]: b[g + 1], g * f: (g + 1) * f], atol=1e-5)
    assert float(dense.abs().sum()) > 0 and torch.count_nonzero(dense[b[0]: b[1], f:]) == 0
    y = dsd_nn(act_fn(s, sdd_nt(x, w3, topo, None, "tensor"), topo), w2, None, "tensor")
    assert torch.allclose(y, _dense_experts(x, w1, w2, w3, b, E, f), atol=1e-4)

    m = MegaBlockGroupedFeedForward(h, E * f, dtype=torch.float32)
    xr = x.clone().requires_grad_()
    out = m(xr, topo)
    ref = _dense_experts(xr, m.w1, m.w2, m.w3, b, E, f)
    assert torch.allclose(out, ref, atol=1e-5)
    got = torch.autograd.grad((out * gy).sum(), [xr, m.w1, m.w2, m.w3])
    want = torch.autograd.grad((ref * gy).sum(), [xr, m.w1, m.w2, m.w3])
    for a, c in zip(got, want):
        assert torch.allclose(a, c, atol=1e-4)

    ff = MegaBlockFeedForward(h, f, E, dtype=torch.float32)
    xb = torch.randn(E, 4, h, requires_grad=True)
    o = ff(xb)
    r = torch.stack([(F.silu(xb[g] @ ff.w1[g]) * (xb[g] @ ff.w3[g])) @ ff.w2[g] for g in range(E)])
    assert torch.allclose(o, r, atol=1e-5)
    ga = torch.autograd.grad(o.sum(), [xb, ff.w1, ff.w2, ff.w3])
    gb = torch.autograd.grad(r.sum(), [xb, ff.w1, ff.w2, ff.w3])
    for a, c in zip(ga, gb):
        assert torch.allclose(a, c, atol=1e-5)


def _megablock_worker(rank, world, mode):
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from internevo_b200.models.megablock import Topology, act_fn, dsd_nn, sdd_nt

    group = dist.group.WORLD
    E, f, h = 3, 8, 16
    topo, b, x, ws, gy = _moe_inputs(E, f, h)
    if mode == "tensor":      # every rank holds f / world features of every expert
        fl = f // world
        topo = Topology(topo.offsets, fl, topo.rows)
        shard = [w.view(E, f, h)[:, rank * fl: (rank + 1) * fl].reshape(E * fl, h).contiguous().requires_grad_() for w in ws]
    else:                     # weight parallel: 1 / world of the rows of the stacked weight
        shard = [w.chunk(world, 0)[rank].contiguous().requires_grad_() for w in ws]
    w1, w2, w3 = shard
    xr = x.clone().requires_grad_()
    y = dsd_nn(act_fn(sdd_nt(xr, w1, topo, group, mode), sdd_nt(xr, w3, topo, group, mode), topo), w2, group, mode)
    (y * gy).sum().backward()
    dist.destroy_process_group()
    return y.detach(), xr.grad, w1.grad, w2.grad, w3.grad


def _check_megablock(mode):
    world, E, f, h = 2, 3, 8, 16
    res = run_distributed(_megablock_worker, world, mode)
    topo, b, x, ws, gy = _moe_inputs(E, f, h)
    xr = x.clone().requires_grad_()
    w1, w2, w3 = [w.clone().requires_grad_() for w in ws]
    ref = _dense_experts(xr, w1, w2, w3, b, E, f)
    (ref * gy).sum().backward()
    for r, (y, dx, d1, d2, d3) in enumerate(res):
        assert torch.allclose(y, ref.detach(), atol=1e-4), mode
        if mode == "tensor":
            fl = f // world
            assert torch.allclose(dx, xr.grad, atol=1e-4)                    # all-reduced over the feature shards
            for d, w in ((d1, w1), (d2, w2), (d3, w3)):
                want = w.grad.view(E, f, h)[:, r * fl: (r + 1) * fl].reshape(E * fl, h)
                assert torch.allclose(d, want, atol=1e-4)
        else:
            # every rank computed the SAME full product on the same tokens: dX is complete per rank, the weight gradient
            # is the group AVERAGE (pre-scaled reduce-scatter) = the single-process gradient
            assert torch.allclose(dx, xr.grad, atol=1e-4)
            for d, w in ((d1, w1), (d2, w2), (d3, w3)):
                assert torch.allclose(d, w.grad.chunk(world, 0)[r], atol=1e-4)


def test_tensor_parallel_block_diagonal_mlp_two_ranks():
    _check_megablock("tensor")


def test_weight_parallel_block_diagonal_mlp_two_ranks():
    _check_megablock("weight")


# ----------------------------------------------------------------------------------------------------- rotary / p2p helpers
def test_flash_attn_style_rotary_functions():
    from internevo_b200.models.modules import ApplyRotaryEmb, ApplyRotaryEmbQKV_

    torch.manual_seed(0)
    b, s, h, d = 2, 5, 3, 8
    t = torch.arange(16).float()[:, None] * torch.rand(d // 2)[None]
    cos, sin = t.cos(), t.sin()

    def rot(z, c, sn):
        z1, z2 = z[..., : d // 2], z[..., d // 2:]
        return torch.cat([z1 * c - z2 * sn, z1 * sn + z2 * c], -1)

    x = torch.randn(b, s, h, d, requires_grad=True)
    y = ApplyRotaryEmb.apply(x, cos, sin, False)
    ref = rot(x, cos[:s][None, :, None], sin[:s][None, :, None])
    g = torch.randn_like(y)
    assert torch.allclose(y, ref, atol=1e-6)
    assert torch.allclose(torch.autograd.grad((y * g).sum(), x)[0], torch.autograd.grad((ref * g).sum(), x)[0], atol=1e-6)
    # partial rotary dim, interleaved pairs: the tail passes through
    y = ApplyRotaryEmb.apply(x, cos[:, :2], sin[:, :2], True)
    a, c = x[..., 0:4:2], x[..., 1:4:2]
    cc, ss = cos[:s, :2][None, :, None], sin[:s, :2][None, :, None]
    want = torch.cat([torch.stack([a * cc - c * ss, a * ss + c * cc], -1).flatten(-2), x[..., 4:]], -1)
    assert torch.allclose(y, want, atol=1e-6)
    # packed qkv with separate key tables, in place
    total = 7
    qkv = torch.randn(total, 3, h, d)
    ck, sk = cos[3: 3 + total], sin[3: 3 + total]
    buf = qkv.clone().requires_grad_()
    work = buf * 1.0
    out = ApplyRotaryEmbQKV_.apply(work, cos[:total], sin[:total], ck, sk, False)
    assert out.data_ptr() == work.data_ptr()
    want = torch.stack([rot(qkv[:, 0], cos[:total, None], sin[:total, None]), rot(qkv[:, 1], ck[:, None], sk[:, None]),
                        qkv[:, 2]], 1)
    assert torch.allclose(out, want, atol=1e-6)
    go = torch.randn_like(out)
    (out * go).sum().backward()
    # the backward of a rotation is the inverse rotation of the incoming gradient
    wantg = torch.stack([rot(go[:, 0], cos[:total, None], -sin[:total, None]), rot(go[:, 1], ck[:, None], -sk[:, None]),
                         go[:, 2]], 1)
    assert torch.allclose(buf.grad, wantg, atol=1e-6)


def test_filling_ops_queue_builds_one_op_per_tensor():
    from internevo_b200.core.communication.p2p import filling_ops_queue

    dist.init_process_group("gloo", rank=0, world_size=1, store=dist.HashStore())
    try:
        q = []
        a, b = torch.zeros(2), torch.zeros(3)
        filling_ops_queue(a, dist.isend, 0, q)
        filling_ops_queue([a, b], dist.irecv, 0, q)
        assert len(q) == 3 and all(isinstance(o, dist.P2POp) for o in q)
        assert q[0].op is dist.isend and q[2].op is dist.irecv and q[2].tensor is b
    finally:
        dist.destroy_process_group()
