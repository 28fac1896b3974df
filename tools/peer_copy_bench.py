"""SM-driven peer copy over NVLink: GB/s of the all-gather push loop (``csrc/gemm_sm100.cu::ag_push_pieces``) in isolation
as a function of the loads a thread keeps in flight, the piece size and the number of CTAs.  2 ranks:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_copy_bench.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from internevo_b200.parallel import symm


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    group = dist.group.WORLD
    n = 64 << 20
    src = torch.randn(n // 2, device="cuda", dtype=torch.bfloat16)
    buf = symm.SymmBuffer(n // 2, torch.bfloat16, group, zero=True)
    peer = buf.base_ptrs[(rank + 1) % dist.get_world_size()]
    res = {}
    for target, name in ((peer, "peer"), (buf.base_ptrs[rank], "local")):
        for unroll in (4, 8, 16):
            for piece in (16384, 65536):
                for ctas in (37, 74, 148, 296):
                    def fn():
                        torch.ops.b200.peer_copy_bench(src, target, n, piece, unroll, ctas)
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize(); dist.barrier()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    for _ in range(10):
                        fn()
                    e.record(); torch.cuda.synchronize()
                    ms = s.elapsed_time(e) / 10
                    res[f"{name}_u{unroll}_p{piece}_c{ctas}"] = round(n / ms / 1e6, 1)   # GB/s (both ranks push at once)
    ok = torch.equal(buf.tensor.view(torch.int16), src.view(torch.int16)) if True else True
    dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/peer_copy_bench.json", "w"), indent=1)
        best = sorted(res.items(), key=lambda kv: -kv[1])[:8]
        print(json.dumps({"best": best, "u4_p65536_c148_peer": res.get("peer_u4_p65536_c148")}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
