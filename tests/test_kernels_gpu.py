"""GPU numerics of every hand-written sm_100a kernel against plain PyTorch fp32 references of the same op."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def _native_loaded():
    from internevo_b200.ops import _lib

    assert _lib.available(), "native extension must be loaded on a GPU box"


def test_native_extension_loaded():
    _native_loaded()
    for op in ("gemm", "rmsnorm_fwd", "rope", "ce_fwd", "adamw", "attn_fwd"):
        assert hasattr(torch.ops.b200, op)


def test_gemm_all_layouts_and_epilogues():
    _native_loaded()
    import kernel_check

    assert kernel_check.check_gemm()


def test_elementwise_kernels():
    _native_loaded()
    import kernel_check

    assert kernel_check.check_elementwise()


def test_attention_fwd_bwd():
    _native_loaded()
    import kernel_check
    from internevo_b200.ops.attention import get_attention_impl

    if get_attention_impl() != "b200":
        pytest.skip("native attention not selected")
    assert kernel_check.check_attn()


def test_linear_autograd_accumulates_into_grad_buf():
    _native_loaded()
    from internevo_b200 import ops

    x = torch.randn(256, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = torch.nn.Parameter(torch.randn(384, 512, device="cuda", dtype=torch.bfloat16))
    w.grad_buf = torch.zeros_like(w)
    w.grad_ready = False
    for _ in range(2):
        y = ops.linear(x, w)
        y.backward(torch.ones_like(y))
    ref = 2 * (torch.ones(256, 384, device="cuda").t() @ x.detach().float())
    assert w.grad is None and w.grad_ready
    assert (w.grad_buf.float() - ref).abs().max() / ref.abs().max() < 2e-2
