"""Tiny driver for ncu captures of the attention kernels: a few fwd+bwd calls at the 7B shape (S=4096, 32/8 heads)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from internevo_b200.ops.attention import flash_attention_packed

S, G, qpk, D = int(os.environ.get("SEQ", 4096)), 8, 4, 128
qkv = torch.randn(S, G, qpk + 2, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
cu = torch.tensor([0, S], device="cuda", dtype=torch.int32)
for _ in range(3):
    out = flash_attention_packed(qkv, cu, S, impl="b200")
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("done")
