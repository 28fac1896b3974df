"""GPU numerics + speed check for the hand-written kernels (run under gpurun; each section in its own process).

    python tools/kernel_check.py gemm        # correctness sweep of the tcgen05 GEMM, all operand layouts / epilogues
    python tools/kernel_check.py gemm_perf   # TFLOP/s vs torch.matmul (cuBLAS) on the model's shapes
    python tools/kernel_check.py elementwise # RMSNorm / RoPE / SwiGLU / CE / AdamW / sumsq vs fp32 PyTorch
    python tools/kernel_check.py attn        # flash attention fwd/bwd vs fp32 reference
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import internevo_b200.ops as ops
from internevo_b200.ops import _lib

dev = "cuda"


def err_report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    diff = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-6)
    rel = (diff.max() / denom).item()
    ok = rel < tol and not torch.isnan(got).any().item()
    print(f"  [{'OK ' if ok else 'BAD'}] {name}: max_abs={diff.max().item():.4e} rel_to_max={rel:.3e} (tol {tol})", flush=True)
    if not ok and got.dim() == 2:
        # localise the damage: error per 32x32 block, coarse map
        M, N = got.shape
        bm, bn = max(1, M // 8), max(1, N // 8)
        grid = diff[: bm * 8, : bn * 8].reshape(8, bm, 8, bn).amax(dim=(1, 3))
        print("    coarse error map (8x8 blocks):")
        for r in grid.tolist():
            print("     ", " ".join(f"{v:9.2e}" for v in r))
        bad = (diff > tol * denom).nonzero()
        print("    first bad idx:", bad[:5].tolist(), "got", [got[i, j].item() for i, j in bad[:5].tolist()],
              "ref", [ref[i, j].item() for i, j in bad[:5].tolist()])
    return ok


def check_gemm():
    torch.manual_seed(0)
    allok = True
    shapes = [(128, 256, 64), (128, 128, 64), (128, 256, 256), (256, 512, 512), (384, 768, 1024), (1000, 520, 264),
              (4096, 4096, 4096), (333, 46272 // 8, 512)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        ref = a.float() @ b.float().t()
        for bn in (256, 128, 512, 384):  # 512 / 384 = 256x256 / 256x192 tile on a CTA pair (cta_group::2)
            out = ops.matmul(a, b, force_bn=bn)
            allok &= err_report(f"NT  {M}x{N}x{K} bn{bn}", out, ref, 1e-2)
        if M % 8 == 0:
            bt = b.t().contiguous()  # [K, N]
            at = a.t().contiguous()  # [K, M]
            for bn in (0, 512):
                out = ops.matmul(a, bt, b_mn=True, force_bn=bn)
                allok &= err_report(f"NN  {M}x{N}x{K} (B MN-major) bn{bn}", out, ref, 1e-2)
                out = ops.matmul(at, bt, a_mn=True, b_mn=True, force_bn=bn)
                allok &= err_report(f"TN  {M}x{N}x{K} (A,B MN-major) bn{bn}", out, ref, 1e-2)
                out = ops.matmul(at, b, a_mn=True, force_bn=bn)
                allok &= err_report(f"TT  {M}x{N}x{K} (A MN-major) bn{bn}", out, ref, 1e-2)
    # ragged K with both operands MN-major (wgrad over an MoE expert's exactly-sized token slab): TMA zero-fills the tail
    for (M, N, K) in [(512, 1024, 1003), (1024, 512, 37), (256, 256, 1)]:
        at = torch.randn(K, M, device=dev, dtype=torch.bfloat16)
        bt = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        ref = at.float().t() @ bt.float()
        for bn in (0, 128, 256, 512):
            out = ops.matmul(at, bt, a_mn=True, b_mn=True, force_bn=bn)
            allok &= err_report(f"TN  {M}x{N}x{K} ragged K bn{bn}", out, ref, 1e-2)
    # epilogues
    M, N, K = 512, 1024, 512
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev, dtype=torch.bfloat16)
    ref = a.float() @ b.float().t()
    for bn in (128, 256, 512, 384):
        allok &= err_report(f"bias bn{bn}", ops.matmul(a, b, bias=bias, force_bn=bn), ref + bias.float(), 1e-2)
        allok &= err_report(f"fp32 out bn{bn}", ops.matmul(a, b, out_dtype=torch.float32, force_bn=bn), ref, 1e-3)
        acc = torch.randn(M, N, device=dev, dtype=torch.float32)
        acc0 = acc.clone()
        ops.matmul(a, b, out=acc, accumulate=True, force_bn=bn)
        allok &= err_report(f"fp32 accumulate bn{bn}", acc, ref + acc0, 1e-3)
        accb = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        accb0 = accb.clone()
        ops.matmul(a, b, out=accb, accumulate=True, force_bn=bn)
        allok &= err_report(f"bf16 accumulate bn{bn}", accb, ref + accb0.float(), 1e-2)
        gu, h = ops.matmul_swiglu(a, b, force_bn=bn)
        allok &= err_report(f"swiglu gu bn{bn}", gu, ref, 1e-2)
        g, u = gu[:, 0::2].float(), gu[:, 1::2].float()
        allok &= err_report(f"swiglu h bn{bn}", h, torch.nn.functional.silu(g) * u, 1e-2)
    # GELU epilogue (+ bias) and its backward kernel
    pre, act = ops.matmul_gelu(a, b, bias)
    pre_ref = ref + bias.float()
    allok &= err_report("gelu pre", pre, pre_ref, 1e-2)
    allok &= err_report("gelu act", act, torch.nn.functional.gelu(pre.float(), approximate="tanh"), 1e-2)
    xg = torch.randn(256, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    wg = torch.randn(384, 512, device=dev, dtype=torch.bfloat16, requires_grad=True) * 0.05
    wg = wg.detach().requires_grad_(True)
    bg = torch.randn(384, device=dev, dtype=torch.bfloat16, requires_grad=True)
    yg = ops.linear_gelu(xg, wg, bg)
    dyg = torch.randn_like(yg)
    yg.backward(dyg)
    xf, wf, bf = (t.detach().float().requires_grad_(True) for t in (xg, wg, bg))
    yf = torch.nn.functional.gelu(xf @ wf.t() + bf, approximate="tanh")
    yf.backward(dyg.float())
    allok &= err_report("linear_gelu fwd", yg, yf, 1e-2)
    allok &= err_report("linear_gelu dx", xg.grad, xf.grad, 2e-2)
    allok &= err_report("linear_gelu dw", wg.grad, wf.grad, 2e-2)
    allok &= err_report("linear_gelu db", bg.grad, bf.grad, 2e-2)
    # ragged M for the pair tile (M % 256 == 128) and wgrad-style accumulate into a strided view
    a3 = torch.randn(384 + 128, 320, device=dev, dtype=torch.bfloat16)[:384 + 128 - 0]
    a3 = torch.randn(640, 320, device=dev, dtype=torch.bfloat16)
    b3 = torch.randn(768, 320, device=dev, dtype=torch.bfloat16)
    allok &= err_report("2cta M=640", ops.matmul(a3, b3, force_bn=512), a3.float() @ b3.float().t(), 1e-2)
    # autograd linear
    x = torch.randn(256, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(384, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = ops.linear(x, w)
    dy = torch.randn_like(y)
    y.backward(dy)
    allok &= err_report("linear fwd", y, x.float() @ w.float().t(), 1e-2)
    allok &= err_report("linear dx", x.grad, dy.float() @ w.float(), 1e-2)
    allok &= err_report("linear dw", w.grad, dy.float().t() @ x.float(), 1e-2)
    print("GEMM_ALL_OK" if allok else "GEMM_HAS_FAILURES", flush=True)
    return allok


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2]


def gemm_perf():
    res = []
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
    T, h, F, V = 4096, 4096, 14336, 92544
    cases = [
        ("8192^3 NT", 8192, 8192, 8192, False, False),
        ("wqkv fwd  NT", T, 6144, h, False, False),
        ("wo   fwd  NT", T, h, h, False, False),
        ("w13  fwd  NT", T, 2 * F, h, False, False),
        ("w2   fwd  NT", T, h, F, False, False),
        ("head fwd  NT", T, V, h, False, False),
        ("w13 dgrad NN", T, h, 2 * F, False, True),
        ("w2  dgrad NN", T, F, h, False, True),
        ("w13 wgrad TN", 2 * F, h, T, True, True),
        ("w2  wgrad TN", h, F, T, True, True),
        ("head wgrad TN", V, h, T, True, True),
    ]
    for name, M, N, K, a_mn, b_mn in cases:
        a = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=torch.bfloat16)
        b = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        A = a.t() if a_mn else a
        Bt = b if b_mn else b.t()
        row = {"case": name, "M": M, "N": N, "K": K}
        # variants are measured round-robin (3 rounds) so that clock / power drift hits all of them alike
        variants = {"ours_default": (0, 1), "ours_2cta": (512, 1), "ours_2cta_nosplit": (512, 0), "ours_bn256": (256, 1)}
        if not b_mn:
            variants["ours_2cta_bn192"] = (384, 1)
        acc = {k: [] for k in list(variants) + ["cublas"]}
        for _ in range(3):
            for key, (bn, split) in variants.items():
                torch.ops.b200.set_gemm_tail_split(split)
                try:
                    acc[key].append(timeit(lambda: ops.matmul(a, b, a_mn=a_mn, b_mn=b_mn, out=out, force_bn=bn), flush=flush))
                except Exception as e:  # noqa
                    acc[key].append(float("nan"))
            acc["cublas"].append(timeit(lambda: torch.matmul(A, Bt, out=out), flush=flush))
        torch.ops.b200.set_gemm_tail_split(1)
        for key, ms in acc.items():
            row[f"{key}_tflops"] = round(2 * M * N * K / (sum(ms) / len(ms)) / 1e9, 1)
        print(json.dumps(row), flush=True)
        res.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gemm_perf.json", "w"), indent=1)


def check_elementwise():
    torch.manual_seed(0)
    ok = True
    T, H = 1000, 4096
    x = torch.randn(T, H, device=dev, dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(T, H, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = (torch.rand(H, device=dev) + 0.5).to(torch.bfloat16).requires_grad_(True)
    y, nr = ops.add_rmsnorm(x, r, w, 1e-5)
    nr_ref = (x.float() + r.float()).to(torch.bfloat16).float()
    y_ref = nr_ref * torch.rsqrt(nr_ref.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
    ok &= err_report("rmsnorm y", y, y_ref, 1e-2)
    ok &= err_report("rmsnorm res", nr, nr_ref, 1e-2)
    dy, dnr = torch.randn_like(y), torch.randn_like(nr)
    torch.autograd.backward([y, nr], [dy, dnr])
    xf = x.detach().float().requires_grad_(True)
    rf = r.detach().float().requires_grad_(True)
    wf = w.detach().float().requires_grad_(True)
    nrf = xf + rf
    yf = nrf * torch.rsqrt(nrf.pow(2).mean(-1, keepdim=True) + 1e-5) * wf
    torch.autograd.backward([yf, nrf], [dy.float(), dnr.float()])
    ok &= err_report("rmsnorm dx", x.grad, xf.grad, 2e-2)
    ok &= err_report("rmsnorm dres", r.grad, rf.grad, 2e-2)
    ok &= err_report("rmsnorm dw", w.grad, wf.grad, 2e-2)

    # RoPE on packed qkv [T, Hkv*(q_per_kv+2), D]
    Hkv, qpk, D = 8, 4, 128
    qkv = torch.randn(T, Hkv * (qpk + 2), D, device=dev, dtype=torch.bfloat16)
    pos = torch.randint(0, 4096, (T,), device=dev, dtype=torch.int32)
    tabs = ops.RotaryTables(D, 1e6)
    cos, sin = tabs.get(4096, dev)
    from internevo_b200.ops.rope import _rope_ref
    ref = _rope_ref(qkv.clone(), pos, cos, sin, qpk + 2, qpk + 1, False, False)
    got = ops.rope_(qkv.clone(), pos, cos, sin, qpk + 2, qpk + 1, False, False)
    ok &= err_report("rope packed", got.flatten(1), ref.flatten(1), 1e-2)
    back = ops.rope_(got.clone(), pos, cos, sin, qpk + 2, qpk + 1, True, False)
    ok &= err_report("rope conj roundtrip", back.flatten(1), qkv.flatten(1), 2e-2)
    ref = _rope_ref(qkv.clone(), pos, cos, sin, 1, 1, False, True)
    got = ops.rope_(qkv.clone(), pos, cos, sin, 1, 1, False, True)
    ok &= err_report("rope interleaved", got.flatten(1), ref.flatten(1), 1e-2)

    # SwiGLU
    gu = torch.randn(T, 2 * 1024, device=dev, dtype=torch.bfloat16, requires_grad=True)
    hh = ops.swiglu_interleaved(gu)
    dh = torch.randn_like(hh)
    hh.backward(dh)
    gf = gu.detach().float().requires_grad_(True)
    hf = torch.nn.functional.silu(gf[:, 0::2]) * gf[:, 1::2]
    hf.backward(dh.float())
    ok &= err_report("swiglu fwd", hh, hf, 1e-2)
    ok &= err_report("swiglu bwd", gu.grad, gf.grad, 1e-2)

    # cross entropy
    V = 92544
    logits = (torch.randn(512, V, device=dev) * 2).to(torch.bfloat16).requires_grad_(True)
    labels = torch.randint(0, V, (512,), device=dev)
    labels[::7] = -100
    for sm in (0.0, 0.1):
        lg = logits.detach().clone().requires_grad_(True)
        loss = ops.cross_entropy(lg, labels, label_smoothing=sm, inplace_backward=False)
        loss.sum().backward()
        lf = logits.detach().float().requires_grad_(True)
        lref = torch.nn.functional.cross_entropy(lf, labels, reduction="none", label_smoothing=sm, ignore_index=-100)
        lref.sum().backward()
        ok &= err_report(f"ce loss sm={sm}", loss[None], lref[None], 1e-3)
        ok &= err_report(f"ce grad sm={sm}", lg.grad, lf.grad, 2e-2)

    # adam + sumsq
    n = 1_000_003
    p = torch.randn(n, device=dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    g = torch.randn(n, device=dev).to(torch.bfloat16)
    plp = torch.empty(n, device=dev, dtype=torch.bfloat16)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    ss = torch.zeros(1, device=dev)
    ops.sumsq_(g, ss)
    ok &= err_report("sumsq", ss[None], g.float().pow(2).sum()[None, None], 1e-4)
    scal = torch.zeros(4, device=dev)
    ops.clip_scalars_(ss, scal, 1.0, 1.0)
    norm = g.float().norm().item()
    print("   scalars", scal.tolist(), "expected mult", 1.0 / norm, "norm", norm)
    for step in (1, 2, 3):
        ops.adamw_(p, m, v, g, plp, 1e-3, 0.9, 0.95, 1e-8, 0.1, step, scal)
        pr.grad = g.float() * (1.0 / norm)
        opt.step()
    ok &= err_report("adamw p", p[None], pr.detach()[None], 1e-5)
    ok &= err_report("adamw p_lp", plp[None], pr.detach()[None], 1e-2)
    print("ELEMENTWISE_ALL_OK" if ok else "ELEMENTWISE_HAS_FAILURES", flush=True)

    # bandwidth numbers
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
    T = 16384
    x = torch.randn(T, H, device=dev, dtype=torch.bfloat16)
    r = torch.randn(T, H, device=dev, dtype=torch.bfloat16)
    wq = torch.ones(H, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.add_rmsnorm(x, r, wq, 1e-5), flush=flush)
    print(f"  rmsnorm fwd {T}x{H}: {ms:.3f} ms  {4 * T * H * 2 / ms / 1e6:.0f} GB/s")
    lg = torch.randn(4096, V, device=dev, dtype=torch.bfloat16)
    lb = torch.randint(0, V, (4096,), device=dev)
    st = torch.empty(4, 4096, device=dev)
    ms = timeit(lambda: torch.ops.b200.ce_fwd(lg, lb, 0, st[0], st[1], st[2], st[3]), flush=flush)
    print(f"  ce fwd 4096x{V}: {ms:.3f} ms  {4096 * V * 2 / ms / 1e6:.0f} GB/s")
    n = 256 * 1024 * 1024
    p = torch.zeros(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    g = torch.zeros(n, device=dev, dtype=torch.bfloat16); plp = torch.empty(n, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.adamw_(p, m, v, g, plp, 1e-3, 0.9, 0.95, 1e-8, 0.1, 1, None))
    print(f"  adamw {n}: {ms:.3f} ms  {n * 28 / ms / 1e6:.0f} GB/s")
    return ok


def check_attn():
    from internevo_b200.ops.attention import attention_ref, flash_attention_varlen

    torch.manual_seed(0)
    ok = True
    for (seqs, H, Hkv, D) in [([128], 2, 2, 128), ([256, 384], 4, 2, 128), ([1000, 24, 513], 8, 2, 128), ([4096], 8, 2, 128)]:
        T = sum(seqs)
        cu = torch.tensor([0] + list(torch.tensor(seqs).cumsum(0)), device=dev, dtype=torch.int32)
        q = torch.randn(T, H, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(T, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(T, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
        out = flash_attention_varlen(q, k, v, cu, max(seqs), causal=True, impl="b200")
        dout = torch.randn_like(out)
        out.backward(dout)
        qf, kf, vf = [t.detach().float().requires_grad_(True) for t in (q, k, v)]
        ref = attention_ref(qf, kf, vf, cu, causal=True)
        ref.backward(dout.float())
        ok &= err_report(f"attn fwd {seqs} H{H}/{Hkv}", out.flatten(1), ref.flatten(1), 2e-2)
        ok &= err_report("attn dq", q.grad.flatten(1), qf.grad.flatten(1), 3e-2)
        ok &= err_report("attn dk", k.grad.flatten(1), kf.grad.flatten(1), 3e-2)
        ok &= err_report("attn dv", v.grad.flatten(1), vf.grad.flatten(1), 3e-2)
        # packed qkv path (one buffer in, one gradient buffer out)
        from internevo_b200.ops.attention import flash_attention_packed
        qpk = H // Hkv
        qkv = torch.cat([q.detach().view(T, Hkv, qpk, D), k.detach()[:, :, None], v.detach()[:, :, None]], 2).contiguous().requires_grad_(True)
        outp = flash_attention_packed(qkv, cu, max(seqs), causal=True, impl="b200")
        outp.backward(dout)
        ok &= err_report("attn packed fwd", outp.flatten(1), ref.flatten(1), 2e-2)
        ok &= err_report("attn packed dq", qkv.grad[:, :, :qpk].reshape(T, -1), qf.grad.flatten(1), 3e-2)
        ok &= err_report("attn packed dk", qkv.grad[:, :, qpk].reshape(T, -1), kf.grad.flatten(1), 3e-2)
        ok &= err_report("attn packed dv", qkv.grad[:, :, qpk + 1].reshape(T, -1), vf.grad.flatten(1), 3e-2)
    print("ATTN_ALL_OK" if ok else "ATTN_HAS_FAILURES", flush=True)
    # speed: 7B shape, T=4096 one sequence and 16k
    for S in (4096, 16384):
        H, Hkv, D = 32, 8, 128
        cu = torch.tensor([0, S], device=dev, dtype=torch.int32)
        q = torch.randn(S, H, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(S, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(S, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
        ms = timeit(lambda: flash_attention_varlen(q, k, v, cu, S, causal=True, impl="b200"))
        fl = 4 * S * S * H * D / 2
        print(f"  attn fwd S={S}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s (causal flops)")
        out = flash_attention_varlen(q, k, v, cu, S, causal=True, impl="b200")
        dout = torch.randn_like(out)
        ms = timeit(lambda: out.backward(dout, retain_graph=True))
        print(f"  attn bwd S={S}: {ms:.3f} ms {2.5 * fl / ms / 1e9:.0f} TFLOP/s")
        try:
            from flash_attn import flash_attn_varlen_func
            ms = timeit(lambda: flash_attn_varlen_func(q, k, v, cu, cu, S, S, causal=True))
            print(f"  flash_attn lib fwd S={S}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s")
            o2 = flash_attn_varlen_func(q, k, v, cu, cu, S, S, causal=True)
            ms = timeit(lambda: o2.backward(dout, retain_graph=True))
            print(f"  flash_attn lib bwd S={S}: {ms:.3f} ms {2.5 * fl / ms / 1e9:.0f} TFLOP/s")
        except Exception as e:  # noqa
            print("  flash_attn lib unavailable:", repr(e)[:200])
    return ok


def check_attn_fwd():
    """forward-only check of the native attention kernel (plain and packed-qkv grouped layouts) + speed."""
    from internevo_b200.ops.attention import attention_ref

    torch.manual_seed(0)
    ok = True
    for (seqs, H, Hkv) in [([128], 2, 2), ([256], 2, 1), ([384, 256], 4, 2), ([1000, 24, 513], 8, 2), ([4096], 8, 2)]:
        D = 128
        T = sum(seqs)
        cu = torch.tensor([0] + torch.tensor(seqs).cumsum(0).tolist(), device=dev, dtype=torch.int32)
        qpk = H // Hkv
        qkv = torch.randn(T, Hkv, qpk + 2, D, device=dev, dtype=torch.bfloat16)
        q4, k, v = qkv[:, :, :qpk], qkv[:, :, -2], qkv[:, :, -1]
        out = torch.empty(T, H, D, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(H, T, device=dev, dtype=torch.float32)
        scale = D ** -0.5
        torch.ops.b200.attn_fwd(q4, k, v, out, lse, cu, max(seqs), scale, True)
        torch.cuda.synchronize()
        q3 = q4.reshape(T, H, D)
        ref = attention_ref(q3, k, v, cu, True, scale)
        ok &= err_report(f"attn fwd packed {seqs} H{H}/{Hkv}", out.flatten(1), ref.flatten(1), 2e-2)
        out2 = torch.empty_like(out)
        torch.ops.b200.attn_fwd(q3.contiguous(), k.contiguous(), v.contiguous(), out2, lse, cu, max(seqs), scale, True)
        ok &= err_report("attn fwd contiguous", out2.flatten(1), ref.flatten(1), 2e-2)
        # lse check
        cul = cu.tolist()
        a, b = cul[0], cul[1]
        s = (q3[a:b].float().transpose(0, 1) @ k[a:b].float().transpose(0, 1).repeat_interleave(qpk, 0).transpose(1, 2)) * scale
        s = s.masked_fill(~torch.ones(b - a, b - a, dtype=torch.bool, device=dev).tril(), float("-inf"))
        ok &= err_report("attn lse", lse[:, a:b], torch.logsumexp(s, -1), 1e-2)
    print("ATTN_FWD_ALL_OK" if ok else "ATTN_FWD_HAS_FAILURES", flush=True)
    for S in (4096, 16384):
        H, Hkv, D = 32, 8, 128
        cu = torch.tensor([0, S], device=dev, dtype=torch.int32)
        q = torch.randn(S, H, D, device=dev, dtype=torch.bfloat16)
        k = torch.randn(S, Hkv, D, device=dev, dtype=torch.bfloat16)
        v = torch.randn(S, Hkv, D, device=dev, dtype=torch.bfloat16)
        out = torch.empty_like(q)
        lse = torch.empty(H, S, device=dev, dtype=torch.float32)
        ms = timeit(lambda: torch.ops.b200.attn_fwd(q, k, v, out, lse, cu, S, D ** -0.5, True))
        fl = 4 * S * S * H * D / 2
        print(f"  b200 attn fwd S={S}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s (causal flops)")
        try:
            from flash_attn import flash_attn_varlen_func
            ms = timeit(lambda: flash_attn_varlen_func(q, k, v, cu, cu, S, S, causal=True))
            print(f"  flash_attn lib fwd S={S}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s")
        except Exception as e:  # noqa
            print("  flash_attn lib unavailable:", repr(e)[:200])
    return ok


def check_small():
    """Every hand-written kernel once on a tiny problem: the workload of ``ci_scripts/sanitize_kernels.sh``
    (``compute-sanitizer --tool memcheck | racecheck | synccheck | initcheck`` slows kernels 10-100x)."""
    from internevo_b200.ops.attention import attention_ref, flash_attention_varlen

    torch.manual_seed(0)
    ok = True
    # GEMM: all layouts, ragged M / K, every tile variant, epilogues
    for (M, N, K) in [(200, 264, 136), (256, 512, 64)]:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        ref = a.float() @ b.float().t()
        for bn in (128, 256, 512):
            ok &= err_report(f"NT {M}x{N}x{K} bn{bn}", ops.matmul(a, b, force_bn=bn), ref, 1e-2)
            if M % 8 == 0:
                ok &= err_report(f"TN bn{bn}", ops.matmul(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True,
                                                         force_bn=bn), ref, 1e-2)
    a = torch.randn(256, 128, device=dev, dtype=torch.bfloat16)
    b = torch.randn(512, 128, device=dev, dtype=torch.bfloat16)
    gu, h = ops.matmul_swiglu(a, b)
    g, u = gu[:, 0::2].float(), gu[:, 1::2].float()
    ok &= err_report("swiglu epilogue", h, torch.nn.functional.silu(g) * u, 2e-2)
    # norms
    x = torch.randn(67, 1024, device=dev, dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(67, 1024, device=dev, dtype=torch.bfloat16)
    w = torch.ones(1024, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y, nr = ops.add_rmsnorm(x, r, w, 1e-5)
    (y.float().sum() + nr.float().sum()).backward()
    ln = ops.LayerNorm(1024, device=dev, dtype=torch.bfloat16)
    y2, _ = ln(x, r)
    y2.float().sum().backward()
    ok &= err_report("layernorm", y2, torch.nn.functional.layer_norm((x + r).float(), (1024,)), 2e-2)
    # rope, swiglu, CE, adam, sumsq
    qkv = torch.randn(67, 6, 128, device=dev, dtype=torch.bfloat16)
    pos = torch.randint(0, 512, (67,), device=dev, dtype=torch.int32)
    cos, sin = ops.RotaryTables(128, 1e6).get(512, dev)
    ops.rope_(qkv, pos, cos, sin, 3, 2, False, False)
    gu = torch.randn(67, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ops.swiglu_interleaved(gu).float().sum().backward()
    lg = torch.randn(67, 1000, device=dev, dtype=torch.bfloat16, requires_grad=True)
    lb = torch.randint(0, 1000, (67,), device=dev)
    loss = ops.cross_entropy(lg, lb, inplace_backward=False)
    loss.sum().backward()
    ok &= err_report("ce", loss[None], torch.nn.functional.cross_entropy(lg.detach().float(), lb, reduction="none")[None], 1e-3)
    n = 10_007
    p, m, v = torch.randn(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    gr = torch.randn(n, device=dev).to(torch.bfloat16)
    ss = torch.zeros(1, device=dev)
    ops.sumsq_(gr, ss)
    scal = torch.zeros(4, device=dev)
    ops.clip_scalars_(ss, scal, 1.0, 1.0)
    ops.adamw_(p, m, v, gr, torch.empty(n, device=dev, dtype=torch.bfloat16), 1e-3, 0.9, 0.95, 1e-8, 0.1, 1, scal)
    # attention fwd + bwd, two packed sequences, GQA
    seqs, H, Hkv, D = [200, 312], 4, 2, 128
    T = sum(seqs)
    cu = torch.tensor([0, 200, 512], device=dev, dtype=torch.int32)
    q = torch.randn(T, H, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(T, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    vv = torch.randn(T, Hkv, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    out = flash_attention_varlen(q, k, vv, cu, max(seqs), causal=True, impl="b200")
    out.backward(torch.randn_like(out))
    ok &= err_report("attn", out.flatten(1), attention_ref(q.detach().float(), k.detach().float(), vv.detach().float(), cu,
                                                          causal=True).flatten(1), 2e-2)
    torch.cuda.synchronize()
    print("SMALL_ALL_OK" if ok else "SMALL_HAS_FAILURES", flush=True)
    return ok


if __name__ == "__main__":
    what = sys.argv[1]
    assert torch.cuda.is_available(), "needs a GPU"
    assert _lib.available()
    print("device", torch.cuda.get_device_name(0), "section", what, flush=True)
    t0 = time.time()
    fn = {"gemm": check_gemm, "gemm_perf": gemm_perf, "elementwise": check_elementwise, "attn": check_attn,
          "attn_fwd": check_attn_fwd, "small": check_small}[what]
    r = fn()
    print(f"section {what} done in {time.time() - t0:.1f}s", flush=True)
    sys.exit(0 if r in (None, True) else 1)
