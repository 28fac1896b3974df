"""Launch the tcgen05 GEMM a few times on one shape (for `ncu --set full -k regex:gemm_bf16 -c 1 ...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from internevo_b200 import ops

M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 28672, 4096)))
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.matmul(a, b, out=out)
torch.cuda.synchronize()
