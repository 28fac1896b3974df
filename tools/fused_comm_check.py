"""Multi-GPU check of the peer-memory kernels (run with torchrun on N GPUs of one node):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/fused_comm_check.py

Validates against NCCL + our own GEMM: device barrier, GEMM→reduce-scatter, GEMM→all-reduce, all-gather→GEMM and the fused
Hybrid-ZeRO step; prints device-timed (max over ranks) latencies and achieved fraction of the NVLink / GEMM roofline.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from internevo_b200 import ops
from internevo_b200.parallel import fused, symm


# roofline denominators: sustained cuBLAS bf16 throughput from MEASURED_PEAKS.json when present (fallback: the profiling
# recipe's 1.4 PFLOP/s) and the recipe's measured 770 GB/s per direction peer copy
PEAK_FLOPS, LINK_BPS = 1.4e15, 770e9
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as _f:
        PEAK_FLOPS = json.load(_f)["bf16_tflops_sustained"] * 1e12
except Exception:
    pass


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    group = dist.group.WORLD
    ok = True
    res = {"world": world}
    flags = symm.flags_for(group)
    for _ in range(3):
        flags.barrier()
    torch.cuda.synchronize()
    res["barrier_us"] = round(timed(flags.barrier, iters=50) * 1e3, 1)

    def stage(msg):
        torch.cuda.synchronize()
        if rank == 0:
            print(f"[stage] {msg} ok={ok}", file=sys.stderr, flush=True)

    stage(f"barrier {res['barrier_us']} us")

    be = fused.TPFusedBackend(group)
    torch.manual_seed(1234 + rank)
    T, h, F = 4096, 4096, 14336
    # ---- GEMM -> RS / AR (row-parallel: K sharded)
    for name, (M, N, K) in {"wo": (T, h, h // world), "w2": (T, h, F // world)}.items():
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16) * 0.1
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.1
        full = ops.matmul(x, w).float()
        ref_rs = torch.empty(M // world, N, device="cuda")
        dist.reduce_scatter_tensor(ref_rs, full)
        out = be.gemm_rs(x, w, all_reduce=False)
        r = rel(out, ref_rs)
        out_b = be.gemm_rs(x, w.t().contiguous(), all_reduce=False, b_mn=True)
        ok &= rel(out_b, ref_rs) < 2e-2
        ok &= r < 2e-2
        stage(f"gemm_rs {name} rel={r}")
        ref_ar = full.clone()
        dist.all_reduce(ref_ar)
        out_ar = be.gemm_rs(x, w, all_reduce=True)
        r2 = rel(out_ar, ref_ar)
        ok &= r2 < 2e-2
        stage(f"gemm_ar {name} rel={r2}")

        def nccl_rs():
            y = ops.matmul(x, w)
            o = torch.empty(M // world, N, device="cuda", dtype=torch.bfloat16)
            dist.reduce_scatter_tensor(o, y)

        def nccl_ar():
            y = ops.matmul(x, w)
            dist.all_reduce(y)

        t_f, t_n = timed(lambda: be.gemm_rs(x, w)), timed(nccl_rs)
        t_fa, t_na = timed(lambda: be.gemm_rs(x, w, all_reduce=True)), timed(nccl_ar)
        t_g = timed(lambda: ops.matmul(x, w))
        flop_ms = 2.0 * M * N * K / PEAK_FLOPS * 1e3
        rs_link = M * N * 2 * (world - 1) / world / LINK_BPS * 1e3       # partial blocks sent to their owners
        res[f"gemm_rs_{name}"] = {"rel_err": r, "fused_ms": round(t_f, 4), "gemm+nccl_ms": round(t_n, 4),
                                  "gemm_only_ms": round(t_g, 4), "roofline_ms": round(max(flop_ms, rs_link), 4),
                                  "frac_of_roofline": round(max(flop_ms, rs_link) / t_f, 3)}
        res[f"gemm_ar_{name}"] = {"rel_err": r2, "fused_ms": round(t_fa, 4), "gemm+nccl_ms": round(t_na, 4),
                                  "roofline_ms": round(max(flop_ms, 2 * rs_link), 4),   # partials out + reduced rows out
                                  "frac_of_roofline": round(max(flop_ms, 2 * rs_link) / t_fa, 3)}
    # ---- AG -> GEMM (column-parallel with sequence parallel)
    for name, (N, K) in {"wqkv": (6144 // world, h), "w13": (2 * F // world, h)}.items():
        xs = torch.randn(T // world, K, device="cuda", dtype=torch.bfloat16) * 0.1
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.1
        xg = torch.empty(T, K, device="cuda", dtype=torch.bfloat16)
        dist.all_gather_into_tensor(xg, xs)
        ref = ops.matmul(xg, w)
        out, gathered = be.ag_gemm(xs, w)
        r = rel(out, ref)
        # dgrad form: all_gather(x) @ w_t with w_t given [K, N] (MN-major B)
        wt = w.t().contiguous()
        out2, _ = be.ag_gemm(xs, wt, b_mn=True)
        ok &= rel(out2, ref) < 2e-2
        ok &= r < 2e-2 and torch.equal(gathered, xg)
        stage(f"ag_gemm {name} rel={r} gathered_equal={torch.equal(gathered, xg)}")

        def nccl_ag():
            g = torch.empty(T, K, device="cuda", dtype=torch.bfloat16)
            dist.all_gather_into_tensor(g, xs)
            ops.matmul(g, w)

        t_f = timed(lambda: be.ag_gemm(xs, w))
        flop_ms = 2.0 * T * N * K / PEAK_FLOPS * 1e3
        link_ms = (T // world) * K * 2 * (world - 1) / LINK_BPS * 1e3     # bytes every GPU must send
        res[f"ag_gemm_{name}"] = {"rel_err": r, "fused_ms": round(t_f, 4),
                                  "nccl+gemm_ms": round(timed(nccl_ag), 4),
                                  "gemm_only_ms": round(timed(lambda: ops.matmul(xg, w)), 4),
                                  "roofline_ms": round(max(flop_ms, link_ms), 4),
                                  "frac_of_roofline": round(max(flop_ms, link_ms) / t_f, 3)}
    # ---- weight-parallel (ISP) forms: weight gather inside the GEMM (forward and dgrad), wgrad -> reduce-scatter (AVG)
    ib = fused.ISPFusedBackend(group)
    Tl = 4096
    for name, (Nt, K) in {"wqkv": (6144, h), "w2": (h, F)}.items():
        xs = torch.randn(Tl, K, device="cuda", dtype=torch.bfloat16) * 0.1          # every rank its own tokens
        wsh = torch.randn(Nt // world, K, device="cuda", dtype=torch.bfloat16) * 0.1
        wfull = torch.empty(Nt, K, device="cuda", dtype=torch.bfloat16)
        dist.all_gather_into_tensor(wfull, wsh)
        ref = ops.matmul(xs, wfull)
        y = ib.gather_gemm(xs, wsh)
        r_f = rel(y, ref)
        dy = torch.randn(Tl, Nt, device="cuda", dtype=torch.bfloat16) * 0.1
        ref_dx = ops.matmul(dy, wfull, b_mn=True)
        dx = ib.gather_gemm(dy, wsh, b_mn=True)
        r_d = rel(dx, ref_dx)
        dw_full = ops.matmul(dy, xs, a_mn=True, b_mn=True, out_dtype=torch.float32)
        ref_dw = torch.empty(Nt // world, K, device="cuda")
        dist.reduce_scatter_tensor(ref_dw, dw_full, op=dist.ReduceOp.AVG)
        dw = torch.empty(Nt // world, K, device="cuda", dtype=torch.bfloat16)
        ib.wgrad_rs(dy, xs, dw, accumulate=False)
        r_w = rel(dw, ref_dw)
        ib.wgrad_rs(dy, xs, dw, accumulate=True)
        r_w2 = rel(dw, 2 * ref_dw)
        ok &= r_f < 2e-2 and r_d < 2e-2 and r_w < 2e-2 and r_w2 < 3e-2
        stage(f"isp {name} fwd rel={r_f} dgrad rel={r_d} wgrad rel={r_w} accumulate rel={r_w2}")

        def nccl_fwd():
            g = torch.empty(Nt, K, device="cuda", dtype=torch.bfloat16)
            dist.all_gather_into_tensor(g, wsh)
            ops.matmul(xs, g)

        def nccl_dgrad():
            g = torch.empty(Nt, K, device="cuda", dtype=torch.bfloat16)
            dist.all_gather_into_tensor(g, wsh)
            ops.matmul(dy, g, b_mn=True)

        def nccl_wgrad():
            full = ops.matmul(dy, xs, a_mn=True, b_mn=True)
            o = torch.empty(Nt // world, K, device="cuda", dtype=torch.bfloat16)
            dist.reduce_scatter_tensor(o, full, op=dist.ReduceOp.AVG)

        t = {"fwd": (timed(lambda: ib.gather_gemm(xs, wsh)), timed(nccl_fwd)),
             "dgrad": (timed(lambda: ib.gather_gemm(dy, wsh, b_mn=True)), timed(nccl_dgrad)),
             "wgrad_rs": (timed(lambda: ib.wgrad_rs(dy, xs, dw, accumulate=False)), timed(nccl_wgrad))}
        flop_ms = 2.0 * Tl * Nt * K / PEAK_FLOPS * 1e3
        link_ms = (Nt // world) * K * 2 * (world - 1) / LINK_BPS * 1e3
        res[f"isp_{name}"] = {"fwd_rel_err": r_f, "dgrad_rel_err": r_d, "wgrad_rel_err": r_w,
                              **{f"{k}_fused_ms": round(v[0], 4) for k, v in t.items()},
                              **{f"{k}_nccl+gemm_ms": round(v[1], 4) for k, v in t.items()},
                              "roofline_ms": round(max(flop_ms, link_ms), 4),
                              "fwd_frac_of_roofline": round(max(flop_ms, link_ms) / t["fwd"][0], 3)}
    # ---- fused ZeRO kernels: reduce-scatter (mean) + sumsq, AdamW + parameter push
    n = 64 * 1024 * 1024
    gbuf = symm.SymmBuffer(n, torch.bfloat16, group)
    pbuf = symm.SymmBuffer(n, torch.bfloat16, group)
    gbuf.tensor.copy_(torch.randn(n, device="cuda") * 0.01)
    ref = gbuf.tensor.float().clone()
    dist.all_reduce(ref)
    ref /= world
    shard = n // world
    lo = rank * shard
    scal = torch.zeros(4, device="cuda")
    p32 = torch.randn(shard, device="cuda")
    m32, v32 = torch.zeros_like(p32), torch.zeros_like(p32)
    flags.barrier()
    torch.ops.b200.reduce_scatter_adam(gbuf.table_ptr(0), pbuf.table_ptr(0), flags.table_ptr(0), rank, world, 0, lo, shard,
                                       p32, m32, v32, scal, 0.0, 0.9, 0.95, 1e-8, 0.0, 1.0, 1.0, float(world), 0)
    torch.cuda.synchronize()
    r = rel(gbuf.tensor[lo:lo + shard], ref[lo:lo + shard])
    ss_ref = gbuf.tensor[lo:lo + shard].float().pow(2).sum()
    ok &= r < 1e-2 and abs(scal[3].item() - ss_ref.item()) / ss_ref.item() < 1e-3
    stage(f"zero rs rel={r} sumsq {scal[3].item()} vs {ss_ref.item()}")
    scal[0] = 1.0
    pref = p32.clone()
    ops.adamw_(pref, torch.zeros_like(pref), torch.zeros_like(pref), gbuf.tensor[lo:lo + shard], None, 1e-3, 0.9, 0.95,
               1e-8, 0.1, 1, None)
    torch.ops.b200.reduce_scatter_adam(gbuf.table_ptr(0), pbuf.table_ptr(0), flags.table_ptr(0), rank, world, 0, lo, shard,
                                       p32, m32, v32, scal, 1e-3, 0.9, 0.95, 1e-8, 0.1, 1 - 0.9, 1 - 0.95, 1.0, 1)
    flags.barrier()
    torch.cuda.synchronize()
    allp = [torch.empty(shard, device="cuda") for _ in range(world)]
    dist.all_gather(allp, pref)
    r_p = rel(pbuf.tensor, torch.cat(allp))
    ok &= r_p < 1e-2 and rel(p32, pref) < 1e-5
    stage(f"zero adam+bcast rel={r_p} p32 rel={rel(p32, pref)}")

    def fused_rs():
        torch.ops.b200.reduce_scatter_adam(gbuf.table_ptr(0), pbuf.table_ptr(0), flags.table_ptr(0), rank, world, 0, lo,
                                           shard, p32, m32, v32, scal, 0.0, 0.9, 0.95, 1e-8, 0.0, 1.0, 1.0, float(world), 0)

    def fused_adam():
        torch.ops.b200.reduce_scatter_adam(gbuf.table_ptr(0), pbuf.table_ptr(0), flags.table_ptr(0), rank, world, 0, lo,
                                           shard, p32, m32, v32, scal, 1e-3, 0.9, 0.95, 1e-8, 0.1, 0.1, 0.05, 1.0, 1)

    tmp = torch.empty(shard, device="cuda", dtype=torch.bfloat16)
    t_rs, t_ad = timed(fused_rs), timed(fused_adam)
    t_nrs = timed(lambda: dist.reduce_scatter_tensor(tmp, gbuf.tensor, op=dist.ReduceOp.AVG))
    t_nag = timed(lambda: dist.all_gather_into_tensor(pbuf.tensor, tmp))
    res["zero"] = {"rs_rel_err": r, "param_rel_err": r_p, "fused_rs_ms": round(t_rs, 3), "nccl_rs_ms": round(t_nrs, 3),
                   "fused_adam_allgather_ms": round(t_ad, 3), "nccl_allgather_only_ms": round(t_nag, 3),
                   "rs_GBps_per_gpu": round(shard * 2 * (world - 1) / t_rs / 1e6, 1),
                   "elements": n}
    # ---- the same two phases through the NVSwitch multicast mapping (multimem.ld_reduce / multimem.st)
    if gbuf.mc_ptr and pbuf.mc_ptr:
        gbuf.tensor.copy_(torch.randn(n, device="cuda") * 0.01)
        ref = gbuf.tensor.float().clone()
        dist.all_reduce(ref)
        ref /= world
        scal.zero_()
        p32 = torch.randn(shard, device="cuda")
        m32.zero_(); v32.zero_()
        flags.barrier()

        def mc(phase, *hyper):
            torch.ops.b200.reduce_scatter_adam_mc(gbuf.mc_ptr, pbuf.mc_ptr, gbuf.tensor.data_ptr(), world, lo, shard, p32,
                                                  m32, v32, scal, *hyper, phase)

        mc(0, 0.0, 0.9, 0.95, 1e-8, 0.0, 1.0, 1.0, float(world))
        torch.cuda.synchronize()
        r = rel(gbuf.tensor[lo:lo + shard], ref[lo:lo + shard])
        ss_ref = gbuf.tensor[lo:lo + shard].float().pow(2).sum()
        ok &= r < 1e-2 and abs(scal[3].item() - ss_ref.item()) / ss_ref.item() < 1e-3
        stage(f"zero NVLS rs rel={r}")
        scal[0] = 1.0
        pref = p32.clone()
        ops.adamw_(pref, torch.zeros_like(pref), torch.zeros_like(pref), gbuf.tensor[lo:lo + shard], None, 1e-3, 0.9,
                   0.95, 1e-8, 0.1, 1, None)
        mc(1, 1e-3, 0.9, 0.95, 1e-8, 0.1, 1 - 0.9, 1 - 0.95, 1.0)
        flags.barrier()
        torch.cuda.synchronize()
        allp = [torch.empty(shard, device="cuda") for _ in range(world)]
        dist.all_gather(allp, pref)
        r_p = rel(pbuf.tensor, torch.cat(allp))
        ok &= r_p < 1e-2 and rel(p32, pref) < 1e-5
        stage(f"zero NVLS adam+bcast rel={r_p}")
        t_rs = timed(lambda: mc(0, 0.0, 0.9, 0.95, 1e-8, 0.0, 1.0, 1.0, float(world)))
        t_ad = timed(lambda: mc(1, 1e-3, 0.9, 0.95, 1e-8, 0.1, 0.1, 0.05, 1.0))
        res["zero_nvls"] = {"rs_rel_err": r, "param_rel_err": r_p, "fused_rs_ms": round(t_rs, 3),
                            "fused_adam_allgather_ms": round(t_ad, 3)}
    else:
        res["zero_nvls"] = {"unavailable": "no multicast mapping (handle.multicast_ptr == 0)"}
    # ---- MoE dispatch / combine over peer memory vs permute + NCCL all-to-all(v) + un-permute
    from internevo_b200.parallel.moe_fused import MoEFusedBackend, slot_plan

    S, H, k, El = 4096, 4096, 2, 1
    E = world * El
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    xt = (torch.randn(S, H, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    eos = torch.stack([torch.randperm(E, device="cuda", generator=gen)[:k] for _ in range(S)]).reshape(-1) if E >= k else \
        torch.zeros(S * k, dtype=torch.int64, device="cuda")
    wts = torch.rand(S * k, device="cuda", generator=gen)
    counts = torch.bincount(eos, minlength=E)
    mbe = MoEFusedBackend(group, H, S * k * world, E)
    cm = mbe.exchange_counts(counts)
    slot_rank, slot_row, per_expert, per_rank = slot_plan(eos, cm, rank, El)
    n_recv = int(per_expert.sum())
    order = torch.argsort(eos, stable=True)
    tok = (torch.arange(S * k, device="cuda") // k)[order]
    send_splits = counts.view(world, El).sum(1).tolist()
    recv_splits = cm.view(world, world, El)[:, rank, :].sum(1).tolist()
    src_expert = torch.repeat_interleave(torch.arange(El, device="cuda").repeat(world), cm.view(world, world, El)[:, rank, :].reshape(-1))
    regroup = torch.argsort(src_expert, stable=True)

    def nccl_dispatch():
        send = xt[tok]
        recv = send.new_empty(n_recv, H)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_splits, input_split_sizes=send_splits)
        return recv[regroup]

    def fused_dispatch_k():
        torch.ops.b200.moe_scatter_rows(xt, slot_rank, slot_row, None, mbe.xbuf.table_ptr(0), 0, None, k)
        mbe.flags.barrier()

    fused_dispatch_k()
    torch.cuda.synchronize()
    r_d = rel(mbe.x_rows(n_recv), nccl_dispatch())
    ok &= r_d == 0.0
    stage(f"moe dispatch rel={r_d} rows={n_recv}")
    yrows = (mbe.x_rows(n_recv).float() * 1.5).to(torch.bfloat16)

    def nccl_combine():
        back = torch.empty_like(yrows).index_copy(0, regroup, yrows)
        outb = back.new_empty(S * k, H)
        dist.all_to_all_single(outb, back, output_split_sizes=send_splits, input_split_sizes=recv_splits)
        return torch.zeros(S, H, device="cuda", dtype=torch.bfloat16).index_add(0, tok, outb * wts[order].to(outb.dtype).unsqueeze(1))

    comb = torch.empty(S, H, device="cuda", dtype=torch.bfloat16)

    def fused_combine_k():
        mbe.y_rows(n_recv).copy_(yrows)
        mbe.flags.barrier()
        torch.ops.b200.moe_gather_combine(comb, wts, slot_rank, slot_row, mbe.ybuf.table_ptr(0), k)
        mbe.flags.barrier()

    fused_combine_k()
    torch.cuda.synchronize()
    r_c = rel(comb, nccl_combine())
    ok &= r_c < 1e-2
    stage(f"moe combine rel={r_c}")
    t_fd, t_nd = timed(fused_dispatch_k), timed(nccl_dispatch)
    t_fc, t_nc = timed(fused_combine_k), timed(nccl_combine)
    remote = S * k * H * 2 * (world - 1) / world
    res["moe"] = {"tokens": S, "k": k, "hidden": H, "dispatch_rel_err": r_d, "combine_rel_err": r_c,
                  "fused_dispatch_ms": round(t_fd, 4), "nccl_permute_a2a_ms": round(t_nd, 4),
                  "fused_combine_ms": round(t_fc, 4), "nccl_a2a_unpermute_ms": round(t_nc, 4),
                  "dispatch_remote_GBps": round(remote / t_fd / 1e6, 1), "combine_remote_GBps": round(remote / t_fc / 1e6, 1)}
    okt = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        res["all_ok"] = bool(okt.item())
        print(json.dumps(res, indent=1))
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open(f"gpurun_out/fused_comm_check_n{world}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
