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


@pytest.mark.parametrize("H", [1024, 4096, 5120])
@pytest.mark.parametrize("with_res,p", [(True, 0.0), (False, 0.0), (True, 0.1)])
def test_dropout_add_layernorm_matches_fp32_reference(H, with_res, p):
    """csrc/layernorm.cu vs plain fp32 PyTorch: outputs, new residual, dx, dres, dweight, dbias."""
    _native_loaded()
    from internevo_b200 import ops
    from internevo_b200.ops.norm import _DropAddLayerNormFn

    torch.manual_seed(0)
    rows = 777
    x = (torch.randn(rows, H, device="cuda") * 2 + 0.5).to(torch.bfloat16).requires_grad_(True)
    res = torch.randn(rows, H, device="cuda").to(torch.bfloat16).requires_grad_(True) if with_res else None
    w = torch.nn.Parameter((torch.rand(H, device="cuda") + 0.5).to(torch.bfloat16))
    b = torch.nn.Parameter((torch.randn(H, device="cuda") * 0.1).to(torch.bfloat16))
    keep = (torch.rand(rows, H, device="cuda") >= p).to(torch.uint8) if p > 0 else None
    scale = 1.0 / (1.0 - p) if p > 0 else 1.0
    n0 = ops.launch_count()
    y, new_res = _DropAddLayerNormFn.apply(x, res, w, b, 1e-5, keep, scale)
    gy, gr = torch.randn_like(y), torch.randn_like(y)
    ((y.float() * gy.float()).sum() + (new_res.float() * gr.float()).sum()).backward()
    assert ops.launch_count() - n0 == 3
    got = [y, new_res, x.grad, res.grad if with_res else None, w.grad, b.grad]
    # fp32 oracle
    xf = x.detach().float().requires_grad_(True)
    rf = res.detach().float().requires_grad_(True) if with_res else None
    wf, bf = w.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    d = xf * keep.float() * scale if keep is not None else xf
    nr = d + rf if with_res else d
    nr_b = nr + (nr.to(torch.bfloat16).float() - nr).detach()  # the kernel normalises the bf16-rounded residual
    yr = torch.nn.functional.layer_norm(nr_b, (H,), wf, bf, 1e-5)
    ((yr * gy.float()).sum() + (nr * gr.float()).sum()).backward()
    want = [yr, nr, xf.grad, rf.grad if with_res else None, wf.grad, bf.grad]
    tol = [2e-2, 1e-2, 2e-2, 2e-2, 2e-2, 2e-2]
    errs = {}
    for name, g, r, t in zip(["y", "new_res", "dx", "dres", "dw", "db"], got, want, tol):
        if r is None:
            continue
        errs[name] = (float((g.detach().float() - r.detach()).norm() / (r.detach().norm() + 1e-9)), t)
    assert all(e < t for e, t in errs.values()), (H, with_res, p, errs)


def test_layernorm_module_in_decoder_block():
    """norm_type='layernorm' models run the native kernel (module API: forward(x) and forward(x, residual))."""
    _native_loaded()
    from internevo_b200 import ops

    ln = ops.LayerNorm(512, eps=1e-5, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(64, 512, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(64, 512, device="cuda", dtype=torch.bfloat16)
    n0 = ops.launch_count()
    y1 = ln(x)
    y2, nr = ln(x, r)
    assert ops.launch_count() - n0 == 2
    assert torch.allclose(y1.float(), torch.nn.functional.layer_norm(x.float(), (512,)), atol=3e-2)
    assert torch.allclose(nr.float(), (x + r).float(), atol=1e-6)
    assert torch.allclose(y2.float(), torch.nn.functional.layer_norm((x + r).float(), (512,)), atol=3e-2)


def test_grouped_gemm_matches_fp32_reference_per_group():
    """The grouped tcgen05 GEMM (MoE experts in one launch, row ranges read from device memory): forward K-major, dgrad
    MN-major, SwiGLU epilogue, and wgrad over the groups' rows (fresh and accumulating, with an EMPTY group) against fp32
    matmuls per group."""
    _native_loaded()
    from internevo_b200 import ops
    from internevo_b200.ops import grouped

    torch.manual_seed(0)
    h, F2, El = 1024, 1536, 4
    counts = torch.tensor([700, 0, 129, 2050], device="cuda")
    off = ops.aligned_offsets(counts)
    R = int(off[-1]) + 256      # spare rows past the last group must stay untouched
    b = off.tolist()
    x = torch.zeros(R, h, device="cuda", dtype=torch.bfloat16)
    for e in range(El):
        x[b[e]: b[e] + int(counts[e])] = torch.randn(int(counts[e]), h, device="cuda") * 0.5
    w13 = [torch.randn(F2, h, device="cuda", dtype=torch.bfloat16) * 0.05 for _ in range(El)]
    w2 = [torch.randn(h, F2 // 2, device="cuda", dtype=torch.bfloat16) * 0.05 for _ in range(El)]

    def rel(a, ref):
        return ((a.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()

    # forward (K-major B) + SwiGLU epilogue
    gu, hh = grouped.grouped_matmul_swiglu(x, w13, off)
    for e in range(El):
        if b[e + 1] > b[e]:
            ref = x[b[e]: b[e + 1]].float() @ w13[e].float().t()
            assert rel(gu[b[e]: b[e + 1]], ref) < 1e-2
            g_, u_ = gu[b[e]: b[e + 1], 0::2].float(), gu[b[e]: b[e + 1], 1::2].float()
            assert rel(hh[b[e]: b[e + 1]], torch.nn.functional.silu(g_) * u_) < 2e-2
    y = grouped.grouped_matmul(hh, w2, off)
    # dgrad (MN-major B)
    dy = torch.randn(R, h, device="cuda", dtype=torch.bfloat16) * 0.1
    for e in range(El):
        dy[b[e] + int(counts[e]): b[e + 1]] = 0
    dh = grouped.grouped_matmul(dy, w2, off, b_mn=True)
    for e in range(El):
        if b[e + 1] > b[e]:
            assert rel(y[b[e]: b[e + 1]], hh[b[e]: b[e + 1]].float() @ w2[e].float().t()) < 1e-2
            assert rel(dh[b[e]: b[e + 1]], dy[b[e]: b[e + 1]].float() @ w2[e].float()) < 1e-2
    # wgrad: fresh into gradient arenas, then accumulate; the empty group must come out exactly zero
    for w in w2:
        w.grad_buf = torch.full_like(w, 7.0)
        w.grad_ready = False
    assert grouped.grouped_wgrad(dy, hh, off, w2) == [None] * El
    refs = [dy[b[e]: b[e + 1]].float().t() @ hh[b[e]: b[e + 1]].float() for e in range(El)]
    for e in range(El):
        if b[e + 1] > b[e]:
            assert rel(w2[e].grad_buf, refs[e]) < 1e-2
        else:
            assert float(w2[e].grad_buf.abs().max()) == 0.0
    grouped.grouped_wgrad(dy, hh, off, w2)
    for e in range(El):
        if b[e + 1] > b[e]:
            assert rel(w2[e].grad_buf, 2 * refs[e]) < 2e-2
    # whole MLP through autograd against the per-expert oracle
    xr = x.clone().requires_grad_(True)
    out = ops.grouped_swiglu_mlp(xr, off, w13, w2)
    out.backward(dy)
    xo = x.clone().requires_grad_(True)
    oracle = torch.zeros_like(out)
    for e in range(El):
        if b[e + 1] > b[e]:
            seg = xo[b[e]: b[e + 1]].float()
            gu_ = (seg @ w13[e].float().t()).to(torch.bfloat16).float()
            h_ = (torch.nn.functional.silu(gu_[:, 0::2]) * gu_[:, 1::2]).to(torch.bfloat16).float()
            oracle[b[e]: b[e + 1]] = (h_ @ w2[e].float().t()).to(torch.bfloat16)
    oracle.backward(dy)
    live = torch.zeros(R, dtype=torch.bool, device="cuda")
    for e in range(El):
        live[b[e]: b[e] + int(counts[e])] = True
    assert rel(out[live], oracle[live].float()) < 2e-2
    assert rel(xr.grad[live], xo.grad[live].float()) < 5e-2
