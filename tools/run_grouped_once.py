"""Tiny driver for ncu captures of the grouped (MoE expert) GEMM: 4 experts, ragged row counts, 7B-MoE4 expert shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from internevo_b200 import ops
from internevo_b200.ops import grouped

h, F2 = 4096, 2 * 5632
counts = torch.tensor([5000, 3100, 4300, 3984], device="cuda")
off = ops.aligned_offsets(counts)
R = int(off[-1])
x = torch.randn(R, h, device="cuda", dtype=torch.bfloat16)
w13 = [torch.randn(F2, h, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(4)]
w2 = [torch.randn(h, F2 // 2, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(4)]
for _ in range(3):
    gu, hh = grouped.grouped_matmul_swiglu(x, w13, off)
    y = grouped.grouped_matmul(hh, w2, off)
    dw = grouped.grouped_wgrad(y, hh, off, w2)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    gu, hh = grouped.grouped_matmul_swiglu(x, w13, off)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"grouped FC1+SwiGLU: {ms:.3f} ms, {2.0 * int(counts.sum()) * F2 * h / ms / 1e9:.0f} TFLOPS (useful rows)")
