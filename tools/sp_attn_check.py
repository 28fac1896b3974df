"""Sequence-parallel attention with in-kernel peer K / V against the single-GPU kernel on the full sequence (2..8 ranks):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sp_attn_check.py

Every rank builds the SAME full q / k / v / dO, runs the single-rank tcgen05 attention on it (the oracle; itself checked
against fp32 PyTorch in tools/kernel_check.py) and the sequence-parallel kernels on its slice of the rows; outputs, dQ, dK, dV
must agree.  Also times the peer kernel against the Ulysses form (3 all-to-alls + local attention + 1 all-to-all)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from internevo_b200.ops.attention import flash_attention_varlen
from internevo_b200.parallel.functional import seq_all_to_all
from internevo_b200.parallel.sp_attention import sp_flash_attention


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    group = dist.group.WORLD
    H, Hkv, D = 16, 4, 128
    res, ok = {"world": world}, True
    cases = {
        "one_sequence": [0, 4096 * world],
        "packed_ragged": None,     # filled below: sequence boundaries that are NOT multiples of 128 and cross rank boundaries
        "many_short": None,
    }
    T = 4096 * world
    g = torch.Generator().manual_seed(5)
    cuts = sorted(set(torch.randint(1, T - 1, (5,), generator=g).tolist()))
    cases["packed_ragged"] = [0] + cuts + [T]
    cases["many_short"] = list(range(0, T + 1, 1024))
    for name, cu_l in cases.items():
        torch.manual_seed(11)
        q = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16) * 0.5
        k = torch.randn(T, Hkv, D, device="cuda", dtype=torch.bfloat16) * 0.5
        v = torch.randn(T, Hkv, D, device="cuda", dtype=torch.bfloat16) * 0.5
        do = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16) * 0.1
        cu = torch.tensor(cu_l, device="cuda", dtype=torch.int32)
        maxlen = max(b - a for a, b in zip(cu_l[:-1], cu_l[1:]))
        qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
        ref = flash_attention_varlen(qf, kf, vf, cu, maxlen, causal=True)
        ref.backward(do)
        Tl = T // world
        sl = slice(rank * Tl, (rank + 1) * Tl)
        ql, kl, vl = (t[sl].clone().requires_grad_(True) for t in (q, k, v))
        out = sp_flash_attention(ql, kl, vl, cu, maxlen, group, causal=True)
        assert out is not None, "sp attention refused a supported shape"
        out.backward(do[sl])
        r = {"out": rel(out, ref[sl]), "dq": rel(ql.grad, qf.grad[sl]), "dk": rel(kl.grad, kf.grad[sl]),
             "dv": rel(vl.grad, vf.grad[sl])}
        worst = torch.tensor([max(r.values())], device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        ok &= worst.item() < 3e-2
        res[name] = {k_: round(v_, 5) for k_, v_ in r.items()}
        res[name]["worst_over_ranks"] = round(worst.item(), 5)
        if rank == 0:
            print(f"[sp_attn] {name} {res[name]} ok={ok}", file=sys.stderr, flush=True)
    # ---- timing: peer kernel vs Ulysses (heads scattered over the group) at 4096 local tokens
    for name in ("one_sequence", "many_short"):
        cu_l = cases[name]
        cu = torch.tensor(cu_l, device="cuda", dtype=torch.int32)
        maxlen = max(b - a for a, b in zip(cu_l[:-1], cu_l[1:]))
        Hb, Hkvb = 32, 8
        Tl = T // world
        ql = torch.randn(Tl, Hb, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        kl = torch.randn(Tl, Hkvb, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        vl = torch.randn(Tl, Hkvb, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        dol = torch.randn(Tl, Hb, D, device="cuda", dtype=torch.bfloat16)

        def peer():
            o = sp_flash_attention(ql, kl, vl, cu, maxlen, group, causal=True)
            o.backward(dol)

        def ulysses():
            qa = seq_all_to_all(ql.contiguous(), group, scatter_dim=1, gather_dim=0)
            ka = seq_all_to_all(kl.contiguous(), group, scatter_dim=1, gather_dim=0)
            va = seq_all_to_all(vl.contiguous(), group, scatter_dim=1, gather_dim=0)
            o = flash_attention_varlen(qa, ka, va, cu, maxlen, causal=True)
            o = seq_all_to_all(o, group, scatter_dim=0, gather_dim=1)
            o.backward(dol)

        res[f"time_{name}"] = {"peer_fwd_bwd_ms": round(timed(peer), 3), "ulysses_fwd_bwd_ms": round(timed(ulysses), 3),
                               "tokens_local": Tl, "heads": Hb}
    res["all_ok"] = bool(ok)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open(f"gpurun_out/sp_attn_check_n{world}.json", "w"), indent=1)
        print(json.dumps(res))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
