"""Per-kernel roofline of one training step of the bench configuration (InternLM2-7B, seq 4096, micro_num 4, 1 GPU).

    python tools/roofline.py [--layers N] --out gpurun_out/roofline_r2      # writes <out>.md, <out>.json, <out>_profile.txt

One step is captured with torch.profiler (CUDA activities only); the kernels are grouped into families, each family's device time
is set against the work it does per step - FLOPs for the tensor-core kernels, minimal HBM bytes for the streaming ones, both
derived from the model shape - and the achieved rate against the DRIVER-measured ceilings in ``MEASURED_PEAKS.json``
(``bf16_tflops_sustained`` for FLOPs, ``hbm_gbs`` copy bandwidth for bytes).  Times taken under the profiler are used for the
fractions only; the benchmark number comes from ``bench.py``.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

# family name -> substrings of the kernel name (first match wins, in this order)
FAMILIES = [
    ("gemm (tcgen05, all linears + head)", ("gemm_bf16_kernel",)),
    ("attention backward", ("attn_bwd_kernel",)),
    ("attention forward", ("attn_fwd_kernel",)),
    ("attention bwd pre/post (delta, dq convert)", ("attn_bwd_delta", "attn_bwd_dq_convert")),
    ("AdamW (+ clip / cast)", ("adamw",)),
    ("SwiGLU backward", ("swiglu_bwd",)),
    ("SwiGLU forward (stand-alone)", ("swiglu_fwd",)),
    ("RMSNorm backward", ("rmsnorm_bwd",)),
    ("RMSNorm forward", ("rmsnorm_fwd",)),
    ("RoPE (q + k in place, fwd + bwd)", ("rope_kernel",)),
    ("cross-entropy forward", ("ce_fwd",)),
    ("cross-entropy backward", ("ce_bwd",)),
    ("grad-norm / column sums", ("sumsq", "colsum", "clip_scalars")),
]


def model_work(m, micro_num, seq_len, micro_bsz=1, segments=1):
    """Analytic work of ONE optimizer step on one GPU (no tensor / pipeline sharding): FLOPs and minimal HBM bytes per family."""
    h, L, H = m["hidden_size"], m["num_layers"], m["num_attention_heads"]
    Hkv, V = m.get("num_kv_attention_heads", H), m["vocab_size"]
    D = h // H
    F = 256 * ((int(h * m["mlp_ratio"]) + 255) // 256)
    T = seq_len * micro_bsz                       # tokens per micro-batch
    tok = T * micro_num
    qkv = (H + 2 * Hkv) * D
    per_tok_layer = 2 * h * qkv + 2 * h * h + 2 * h * 2 * F + 2 * F * h
    gemm_fwd = tok * (L * per_tok_layer + 2 * h * V)
    S = seq_len // segments
    n_seq = micro_bsz * segments * micro_num
    attn_fwd = L * n_seq * 2 * (2 * S * S * D * H) / 2          # QK^T and PV, causal half
    n_params = L * (qkv * h + h * h + 3 * F * h + 2 * h) + 2 * V * h + h
    e = 2                                                        # bf16
    return {
        "gemm (tcgen05, all linears + head)": ("flops", 3.0 * gemm_fwd),          # forward + dgrad + wgrad
        "attention forward": ("flops", attn_fwd),
        "attention backward": ("flops", 2.5 * attn_fwd),                          # 5 GEMMs vs 2
        # grad (bf16) read, fp32 master + 2 moments read / written, bf16 parameter written
        "AdamW (+ clip / cast)": ("bytes", n_params * (2 + 3 * 8 + 2)),
        # read gate|up [T, 2F] and dh [T, F], write d(gate|up) [T, 2F]
        "SwiGLU backward": ("bytes", L * tok * F * e * 5),
        # fused residual-add + norm: read hidden + residual, write normed + new residual; (2L + 1) calls per micro-batch
        "RMSNorm forward": ("bytes", (2 * L + 1) * tok * h * e * 4),
        # read dy, saved input, incoming residual gradient; write dx (also the outgoing residual gradient)
        "RMSNorm backward": ("bytes", (2 * L + 1) * tok * h * e * 4),
        "RoPE (q + k in place, fwd + bwd)": ("bytes", 2 * L * tok * (H + Hkv) * D * e * 2),
        "cross-entropy forward": ("bytes", tok * V * e),
        "cross-entropy backward": ("bytes", tok * V * e * 2),
    }, {"tokens_per_step": tok, "params": n_params, "ffn": F, "head_dim": D}


def family_of(name: str):
    for fam, keys in FAMILIES:
        if any(k in name for k in keys):
            return fam
    if "nccl" in name.lower():
        return "NCCL"
    if name.startswith("Memcpy") or name.startswith("Memset"):
        return "memcpy / memset"
    return "other (ATen: embedding, fills, index ops)"


def build_table(events, work, peaks):
    """``events``: iterable of (kernel name, device time in us, calls).  Returns (rows, total ms)."""
    agg = {}
    for name, us, calls in events:
        fam = family_of(name)
        t, c = agg.get(fam, (0.0, 0))
        agg[fam] = (t + us, c + calls)
    total_us = sum(t for t, _ in agg.values()) or 1.0
    rows = []
    for fam, (us, calls) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        row = {"family": fam, "calls": calls, "ms": us / 1e3, "share": us / total_us}
        kind, amount = work.get(fam, (None, None))
        if kind == "flops":
            row.update(work=f"{amount / 1e12:.1f} TFLOP", achieved=amount / (us * 1e-6) / 1e12, unit="TFLOP/s",
                       peak=peaks["bf16_tflops_sustained"], peak_name="bf16_tflops_sustained")
        elif kind == "bytes":
            row.update(work=f"{amount / 1e9:.1f} GB", achieved=amount / (us * 1e-6) / 1e9, unit="GB/s", peak=peaks["hbm_gbs"],
                       peak_name="hbm_gbs")
        if "achieved" in row:
            row["fraction"] = row["achieved"] / row["peak"]
        rows.append(row)
    return rows, total_us / 1e3


def to_markdown(rows, total_ms, meta, peaks):
    out = [f"# Per-kernel roofline of one training step ({meta['model']}, {meta['tokens_per_step']} tokens / step, 1 GPU)", "",
           f"Ceilings from `MEASURED_PEAKS.json` (driver-measured): cuBLAS bf16 sustained {peaks['bf16_tflops_sustained']} TFLOP/s "
           f"(burst {peaks['bf16_tflops']}), copy bandwidth {peaks['hbm_gbs']} GB/s.  Kernel time of the step under the profiler: "
           f"{total_ms:.1f} ms.  FLOPs: causal attention counted once (not the Megatron formula); bytes: the minimum each kernel "
           "has to move.", "",
           "| kernel family | launches | ms / step | share | work / step | achieved | ceiling | fraction |",
           "|---|---|---|---|---|---|---|---|"]
    for r in rows:
        if "achieved" in r:
            out.append(f"| {r['family']} | {r['calls']} | {r['ms']:.2f} | {100 * r['share']:.1f} % | {r['work']} | "
                       f"{r['achieved']:.0f} {r['unit']} | {r['peak']:.0f} ({r['peak_name']}) | **{100 * r['fraction']:.0f} %** |")
        else:
            out.append(f"| {r['family']} | {r['calls']} | {r['ms']:.2f} | {100 * r['share']:.1f} % | | | | |")
    return "\n".join(out) + "\n"


def main():
    argv = sys.argv[1:]
    out = "gpurun_out/roofline"
    if "--out" in argv:
        i = argv.index("--out")
        out = argv[i + 1]
        del argv[i: i + 2]
    sys.argv = [sys.argv[0]] + argv
    a = bench.parse()
    world, _ = bench.setup_env(a)
    cfg, _, _ = bench.build_config(a, world)
    import torch

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
    mdl = cfg["model"]
    T = a.seq_len * a.micro_bsz
    batches = bench.make_batches(4, a.micro_num, T, mdl["vocab_size"], pin=True, segments=a.segments)
    dev = [({k: v.cuda() for k, v in d.items()}, l.cuda()) for d, l in batches]

    def step(i):
        d, l = dev[i % len(dev)]
        trainer.zero_grad()
        trainer.execute_schedule((dict(d), l), forward_only=False, return_loss=True, return_output_label=False)
        ok, _ = trainer.step()
        assert ok, "optimizer step skipped"

    for i in range(3):
        step(i)
    getattr(opt, "flush_param_update", lambda: None)()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step(3)
        getattr(opt, "flush_param_update", lambda: None)()
        torch.cuda.synchronize()
    avg = prof.key_averages()
    events = []
    for e in avg:
        us = getattr(e, "self_device_time_total", None)
        if us is None:
            us = getattr(e, "self_cuda_time_total", 0.0)
        if us and us > 0:
            events.append((e.key, float(us), int(e.count)))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    work, meta = model_work(mdl, a.micro_num, a.seq_len, a.micro_bsz, a.segments)
    meta["model"] = f"h{mdl['hidden_size']} L{mdl['num_layers']} H{mdl['num_attention_heads']}/kv" \
                    f"{mdl.get('num_kv_attention_heads')} V{mdl['vocab_size']}, seq {a.seq_len} x micro_num {a.micro_num}"
    rows, total_ms = build_table(events, work, peaks)
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out + ".md", "w") as f:
        f.write(to_markdown(rows, total_ms, meta, peaks))
    with open(out + ".json", "w") as f:
        json.dump({"meta": meta, "peaks": peaks, "kernel_ms_total": total_ms, "rows": rows,
                   "clocks": bench.ClockSampler.snapshot() if hasattr(bench.ClockSampler, "snapshot") else None}, f, indent=1)
    try:
        table = avg.table(sort_by="self_device_time_total", row_limit=45, max_name_column_width=90)
    except Exception:   # older attribute name
        table = avg.table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
    with open(out + "_profile.txt", "w") as f:
        f.write(table)
    print(to_markdown(rows, total_ms, meta, peaks))


if __name__ == "__main__":
    main()
