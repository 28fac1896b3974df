import os, sys, json
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist
from internevo_b200 import ops
from internevo_b200.parallel import fused
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
g = dist.group.WORLD
def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return round(t.item(), 4)
ib = fused.ISPFusedBackend(g)
Tl, Nt, K = 4096, 6144, 4096
xs = torch.randn(Tl, K, device="cuda", dtype=torch.bfloat16) * 0.1
wsh = torch.randn(Nt // world, K, device="cuda", dtype=torch.bfloat16) * 0.1
wfull = torch.randn(Nt, K, device="cuda", dtype=torch.bfloat16) * 0.1
dy = torch.randn(Tl, Nt, device="cuda", dtype=torch.bfloat16) * 0.1
res = {}
res["plain_dgrad"] = timed(lambda: ops.matmul(dy, wfull, b_mn=True))
res["plain_fwd"] = timed(lambda: ops.matmul(xs, wfull))
res["fused_fwd"] = timed(lambda: ib.gather_gemm(xs, wsh))
res["fused_dgrad"] = timed(lambda: ib.gather_gemm(dy, wsh, b_mn=True))
os.environ["B200_DEBUG_NOWAIT"] = "1"
res["fused_fwd_nowait"] = timed(lambda: ib.gather_gemm(xs, wsh))
res["fused_dgrad_nowait"] = timed(lambda: ib.gather_gemm(dy, wsh, b_mn=True))
if rank == 0: print("PROBE", json.dumps(res))
dist.destroy_process_group()
