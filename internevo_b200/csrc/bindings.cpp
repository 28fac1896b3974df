// torch op registrations for the sm_100a kernels (namespace torch.ops.b200). Thin: shape checks + raw launches on the
// current stream; allocation and autograd live in python (internevo_b200/ops).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <cstring>
#include <vector>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "attention_sm100.h"
#include "comm_kernels.h"
#include "elementwise.h"
#include "gemm_sm100.h"

using at::Tensor;
using c10::optional;

namespace {

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
inline const void* optptr(const optional<Tensor>& t) { return t.has_value() ? t->data_ptr() : nullptr; }

#define CHECK_RC(rc, what) TORCH_CHECK((rc) == 0, what, " failed with code ", (rc))
#define CHECK_BF16(t) TORCH_CHECK((t).is_cuda() && (t).scalar_type() == at::kBFloat16, #t " must be a CUDA bf16 tensor")

// D[M,N] (+)= A * B^T.  a: [M,K] (a_mn=false) or [K,M] (a_mn=true); b: [N,K] (b_mn=false) or [K,N] (b_mn=true)
void gemm(const Tensor& a, const Tensor& b, Tensor& out, bool a_mn, bool b_mn, const optional<Tensor>& bias,
          int64_t flags, const optional<Tensor>& h, int64_t force_bn, int64_t max_ctas) {
    CHECK_BF16(a);
    CHECK_BF16(b);
    TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && out.dim() == 2, "gemm expects 2-D tensors");
    TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1 && out.stride(1) == 1, "innermost dim must be contiguous");
    c10::cuda::CUDAGuard guard(a.device());
    b200::GemmDesc g;
    g.M = a_mn ? a.size(1) : a.size(0);
    g.K = a_mn ? a.size(0) : a.size(1);
    g.N = b_mn ? b.size(1) : b.size(0);
    const int64_t kb = b_mn ? b.size(0) : b.size(1);
    TORCH_CHECK(kb == g.K, "gemm: K mismatch ", kb, " vs ", g.K);
    TORCH_CHECK(out.size(0) == g.M && out.size(1) == g.N, "gemm: bad output shape");
    // K is the contiguous dimension of a K-major operand (16-byte rows for TMA); with both operands MN-major (wgrad:
    // dy^T @ x over a ragged token count) K is the outer TMA dimension and the last k-block is zero-filled by TMA.
    TORCH_CHECK(g.N % 8 == 0 && (g.K % 8 == 0 || (a_mn && b_mn)), "gemm: N (and K for K-major operands) must be multiples of 8");
    TORCH_CHECK(a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0, "gemm: row strides must be multiples of 8 elements");
    TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(b.data_ptr()) & 15) == 0,
                "gemm: operands must be 16-byte aligned");
    if (a_mn) TORCH_CHECK(g.M % 8 == 0, "gemm: M must be a multiple of 8 for MN-major A");
    const bool f32 = flags & b200::GEMM_OUT_F32;
    TORCH_CHECK(out.scalar_type() == (f32 ? at::kFloat : at::kBFloat16), "gemm: output dtype/flag mismatch");
    TORCH_CHECK(out.stride(0) % (f32 ? 4 : 8) == 0, "gemm: output row stride alignment");
    g.A = a.data_ptr(); g.lda = a.stride(0); g.a_mn_major = a_mn;
    g.B = b.data_ptr(); g.ldb = b.stride(0); g.b_mn_major = b_mn;
    g.D = out.data_ptr(); g.ldd = out.stride(0);
    g.bias = optptr(bias);
    if (bias.has_value()) { CHECK_BF16(*bias); TORCH_CHECK(bias->numel() == g.N, "gemm: bias size"); }
    if (flags & b200::GEMM_SWIGLU) {
        TORCH_CHECK(h.has_value() && h->scalar_type() == at::kBFloat16 && h->size(0) == g.M && h->size(1) == g.N / 2 &&
                        h->stride(1) == 1 && h->stride(0) % 8 == 0 && g.N % 16 == 0,
                    "gemm: bad swiglu output");
        g.H = h->data_ptr(); g.ldh = h->stride(0);
    }
    if (flags & b200::GEMM_GELU) {
        TORCH_CHECK(h.has_value() && h->scalar_type() == at::kBFloat16 && h->size(0) == g.M && h->size(1) == g.N &&
                        h->stride(1) == 1 && h->stride(0) % 8 == 0 && !(flags & b200::GEMM_OUT_F32),
                    "gemm: bad gelu output");
        g.H = h->data_ptr(); g.ldh = h->stride(0);
    }
    g.flags = static_cast<int>(flags);
    g.force_bn = static_cast<int>(force_bn);
    g.max_ctas = static_cast<int>(max_ctas);
    CHECK_RC(b200::gemm_bf16(g, cur_stream()), "b200::gemm");
}

void rmsnorm_fwd(const Tensor& x, const optional<Tensor>& res_in, const Tensor& w, Tensor& y,
                 const optional<Tensor>& res_out, const optional<Tensor>& rstd, double eps) {
    CHECK_BF16(x); CHECK_BF16(w); CHECK_BF16(y);
    TORCH_CHECK(x.is_contiguous() && y.is_contiguous() && w.is_contiguous(), "rmsnorm: contiguous tensors required");
    c10::cuda::CUDAGuard guard(x.device());
    const int H = x.size(-1);
    const int rows = x.numel() / H;
    CHECK_RC(b200::rmsnorm_fwd(x.data_ptr(), optptr(res_in), w.data_ptr(), y.data_ptr(),
                               res_out.has_value() ? res_out->data_ptr() : nullptr,
                               rstd.has_value() ? rstd->data_ptr<float>() : nullptr, rows, H, (float)eps, cur_stream()),
             "b200::rmsnorm_fwd");
}

int64_t rmsnorm_bwd_blocks(int64_t rows) { return b200::rmsnorm_bwd_blocks((int)rows); }

void rmsnorm_bwd(const Tensor& dy, const Tensor& res, const Tensor& w, const Tensor& rstd, const optional<Tensor>& dres,
                 Tensor& dx, Tensor& dw_partial, Tensor& dw, bool accumulate) {
    CHECK_BF16(dy); CHECK_BF16(res); CHECK_BF16(w); CHECK_BF16(dx);
    TORCH_CHECK(dy.is_contiguous() && res.is_contiguous() && dx.is_contiguous(), "rmsnorm_bwd: contiguous required");
    c10::cuda::CUDAGuard guard(dy.device());
    const int H = dy.size(-1);
    const int rows = dy.numel() / H;
    TORCH_CHECK(dw_partial.numel() >= (int64_t)b200::rmsnorm_bwd_blocks(rows) * H, "rmsnorm_bwd: partial buffer too small");
    const bool f32 = dw.scalar_type() == at::kFloat;
    CHECK_RC(b200::rmsnorm_bwd(dy.data_ptr(), res.data_ptr(), w.data_ptr(), rstd.data_ptr<float>(), optptr(dres),
                               dx.data_ptr(), dw_partial.data_ptr<float>(), f32 ? dw.data_ptr<float>() : nullptr,
                               f32 ? nullptr : dw.data_ptr(), accumulate, rows, H, cur_stream()),
             "b200::rmsnorm_bwd");
}

void layernorm_fwd(const Tensor& x, const optional<Tensor>& res_in, const optional<Tensor>& keep, double drop_scale,
                   const Tensor& w, const optional<Tensor>& b, Tensor& y, const optional<Tensor>& res_out, Tensor& mean,
                   Tensor& rstd, double eps) {
    CHECK_BF16(x); CHECK_BF16(w); CHECK_BF16(y);
    TORCH_CHECK(x.is_contiguous() && y.is_contiguous() && w.is_contiguous(), "layernorm: contiguous tensors required");
    if (keep.has_value())
        TORCH_CHECK(keep->scalar_type() == at::kByte && keep->is_contiguous() && keep->numel() == x.numel(),
                    "layernorm: keep mask must be uint8 of x's shape");
    c10::cuda::CUDAGuard guard(x.device());
    const int H = x.size(-1);
    const int rows = x.numel() / H;
    CHECK_RC(b200::layernorm_fwd(x.data_ptr(), optptr(res_in), keep.has_value() ? keep->data_ptr<uint8_t>() : nullptr,
                                 (float)drop_scale, w.data_ptr(), optptr(b), y.data_ptr(),
                                 res_out.has_value() ? res_out->data_ptr() : nullptr, mean.data_ptr<float>(),
                                 rstd.data_ptr<float>(), rows, H, (float)eps, cur_stream()),
             "b200::layernorm_fwd");
}

int64_t layernorm_bwd_blocks(int64_t rows) { return b200::layernorm_bwd_blocks((int)rows); }

void layernorm_bwd(const Tensor& dy, const Tensor& res, const Tensor& w, const Tensor& mean, const Tensor& rstd,
                   const optional<Tensor>& dres, Tensor& dx, Tensor& partial, Tensor& dwdb) {
    CHECK_BF16(dy); CHECK_BF16(res); CHECK_BF16(w); CHECK_BF16(dx);
    TORCH_CHECK(dy.is_contiguous() && res.is_contiguous() && dx.is_contiguous(), "layernorm_bwd: contiguous required");
    c10::cuda::CUDAGuard guard(dy.device());
    const int H = dy.size(-1);
    const int rows = dy.numel() / H;
    TORCH_CHECK(partial.numel() >= (int64_t)b200::layernorm_bwd_blocks(rows) * 2 * H && dwdb.numel() == 2 * H &&
                    partial.scalar_type() == at::kFloat && dwdb.scalar_type() == at::kFloat,
                "layernorm_bwd: bad partial / dwdb buffers");
    CHECK_RC(b200::layernorm_bwd(dy.data_ptr(), res.data_ptr(), w.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(),
                                 optptr(dres), dx.data_ptr(), partial.data_ptr<float>(), dwdb.data_ptr<float>(), rows, H,
                                 cur_stream()),
             "b200::layernorm_bwd");
}

// x: [T, heads, D] view with contiguous (heads, D) and token stride x.stride(0)
void rope(Tensor& x, const optional<Tensor>& pos, const Tensor& cos_t, const Tensor& sin_t, int64_t group,
          int64_t rot_per_group, bool conj, bool interleaved) {
    CHECK_BF16(x);
    TORCH_CHECK(x.dim() == 3 && x.stride(2) == 1 && x.stride(1) == x.size(2), "rope: x must be [T, heads, D] with dense heads");
    TORCH_CHECK(cos_t.scalar_type() == at::kFloat && cos_t.is_contiguous() && sin_t.is_contiguous(), "rope: tables fp32");
    if (pos.has_value()) TORCH_CHECK(pos->scalar_type() == at::kInt && pos->is_contiguous(), "rope: pos must be int32");
    c10::cuda::CUDAGuard guard(x.device());
    CHECK_RC(b200::rope_inplace(x.data_ptr(), pos.has_value() ? pos->data_ptr<int>() : nullptr, cos_t.data_ptr<float>(),
                                sin_t.data_ptr<float>(), x.size(0), x.size(1), x.size(2), x.stride(0), group,
                                rot_per_group, conj, interleaved, cur_stream()),
             "b200::rope");
}

void swiglu_fwd(const Tensor& gu, Tensor& h) {
    CHECK_BF16(gu); CHECK_BF16(h);
    TORCH_CHECK(gu.is_contiguous() && h.is_contiguous() && gu.numel() == 2 * h.numel(), "swiglu_fwd: shapes");
    c10::cuda::CUDAGuard guard(gu.device());
    const int64_t F = h.size(-1);
    CHECK_RC(b200::swiglu_fwd(gu.data_ptr(), h.data_ptr(), h.numel() / F, F, cur_stream()), "b200::swiglu_fwd");
}
// one decode step: q [B, H, D] against the first `seqlen` positions of the caches [B, Smax, Hkv, D]
void attn_decode(const Tensor& q, const Tensor& kcache, const Tensor& vcache, Tensor& out, Tensor& work, Tensor& tickets,
                 int64_t seqlen, int64_t nsplit, double scale, const optional<Tensor>& seqlen_dev) {
    CHECK_BF16(q); CHECK_BF16(kcache); CHECK_BF16(vcache); CHECK_BF16(out);
    TORCH_CHECK(q.dim() == 3 && q.is_contiguous() && out.is_contiguous() && kcache.dim() == 4, "attn_decode: shapes");
    const int64_t B = q.size(0), H = q.size(1), D = q.size(2), Hkv = kcache.size(2);
    TORCH_CHECK(kcache.stride(3) == 1 && kcache.stride(2) == D && vcache.strides() == kcache.strides(), "attn_decode: cache layout");
    TORCH_CHECK(work.scalar_type() == at::kFloat && work.numel() >= B * H * nsplit * (D + 4), "attn_decode: work size");
    TORCH_CHECK(tickets.scalar_type() == at::kInt && tickets.numel() >= B * Hkv, "attn_decode: tickets");
    if (seqlen_dev.has_value())
        TORCH_CHECK(seqlen_dev->is_cuda() && seqlen_dev->scalar_type() == at::kInt && seqlen_dev->numel() >= 1,
                    "attn_decode: seqlen_dev must be a cuda int32 tensor");
    c10::cuda::CUDAGuard guard(q.device());
    CHECK_RC(b200::attn_decode(q.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), out.data_ptr(), work.data_ptr<float>(),
                               reinterpret_cast<unsigned int*>(tickets.data_ptr<int>()), B, H, Hkv, D, seqlen,
                               seqlen_dev.has_value() ? seqlen_dev->data_ptr<int>() : nullptr, nsplit,
                               kcache.stride(0), kcache.stride(1), static_cast<float>(scale), cur_stream()),
             "b200::attn_decode");
}
void gelu_bwd(const Tensor& dh, const Tensor& pre, Tensor& dpre) {
    CHECK_BF16(dh); CHECK_BF16(pre); CHECK_BF16(dpre);
    TORCH_CHECK(dh.is_contiguous() && pre.is_contiguous() && dpre.is_contiguous() && dh.numel() == pre.numel(),
                "gelu_bwd: contiguous tensors of equal size");
    c10::cuda::CUDAGuard guard(dh.device());
    CHECK_RC(b200::gelu_bwd(dh.data_ptr(), pre.data_ptr(), dpre.data_ptr(), dh.numel(), cur_stream()), "b200::gelu_bwd");
}
void swiglu_bwd(const Tensor& dh, const Tensor& gu, Tensor& dgu) {
    CHECK_BF16(dh); CHECK_BF16(gu); CHECK_BF16(dgu);
    TORCH_CHECK(dh.is_contiguous() && gu.is_contiguous() && dgu.is_contiguous(), "swiglu_bwd: contiguous");
    c10::cuda::CUDAGuard guard(gu.device());
    const int64_t F = dh.size(-1);
    CHECK_RC(b200::swiglu_bwd(dh.data_ptr(), gu.data_ptr(), dgu.data_ptr(), dh.numel() / F, F, cur_stream()),
             "b200::swiglu_bwd");
}

void ce_fwd(const Tensor& logits, const Tensor& labels, int64_t vocab_start, Tensor& out_max, Tensor& out_sum,
            Tensor& out_sumx, Tensor& out_tgt) {
    CHECK_BF16(logits);
    TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1 && labels.scalar_type() == at::kLong && labels.is_contiguous(),
                "ce_fwd: logits [rows, V] bf16, labels int64");
    c10::cuda::CUDAGuard guard(logits.device());
    CHECK_RC(b200::ce_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr<int64_t>(), logits.size(0), logits.size(1),
                          vocab_start, out_max.data_ptr<float>(), out_sum.data_ptr<float>(), out_sumx.data_ptr<float>(),
                          out_tgt.data_ptr<float>(), cur_stream()),
             "b200::ce_fwd");
}
void ce_bwd(Tensor& logits, const Tensor& labels, const Tensor& lse, const Tensor& gscale, int64_t vocab_start,
            double smoothing, int64_t total_classes, int64_t ignore_index) {
    CHECK_BF16(logits);
    c10::cuda::CUDAGuard guard(logits.device());
    CHECK_RC(b200::ce_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr<int64_t>(), lse.data_ptr<float>(),
                          gscale.data_ptr<float>(), logits.size(0), logits.size(1), vocab_start, (float)smoothing,
                          total_classes, ignore_index, cur_stream()),
             "b200::ce_bwd");
}

void adamw(Tensor& p, Tensor& m, Tensor& v, const Tensor& g, const optional<Tensor>& p_lp, double lr, double beta1,
           double beta2, double eps, double wd, double bc1, double bc2, const optional<Tensor>& scalars) {
    TORCH_CHECK(p.scalar_type() == at::kFloat && m.scalar_type() == at::kFloat && v.scalar_type() == at::kFloat,
                "adamw: fp32 state required");
    TORCH_CHECK(p.is_contiguous() && m.is_contiguous() && v.is_contiguous() && g.is_contiguous(), "adamw: contiguous");
    TORCH_CHECK(g.numel() == p.numel(), "adamw: grad size");
    const bool gbf = g.scalar_type() == at::kBFloat16;
    TORCH_CHECK(gbf || g.scalar_type() == at::kFloat, "adamw: grad must be bf16 or fp32");
    c10::cuda::CUDAGuard guard(p.device());
    CHECK_RC(b200::adamw_step(p.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), g.data_ptr(), gbf,
                              p_lp.has_value() ? p_lp->data_ptr() : nullptr, p.numel(), (float)lr, (float)beta1,
                              (float)beta2, (float)eps, (float)wd, (float)bc1, (float)bc2,
                              scalars.has_value() ? scalars->data_ptr<float>() : nullptr, cur_stream()),
             "b200::adamw");
}

void sumsq(const Tensor& g, Tensor& out) {
    TORCH_CHECK(g.is_contiguous() && out.scalar_type() == at::kFloat, "sumsq: contiguous input, fp32 output");
    const bool gbf = g.scalar_type() == at::kBFloat16;
    TORCH_CHECK(gbf || g.scalar_type() == at::kFloat, "sumsq: bf16 or fp32");
    c10::cuda::CUDAGuard guard(g.device());
    CHECK_RC(b200::sumsq(g.data_ptr(), gbf, g.numel(), out.data_ptr<float>(), cur_stream()), "b200::sumsq");
}
void clip_scalars(const Tensor& sumsq_in, Tensor& scalars, double loss_scale, double clip) {
    c10::cuda::CUDAGuard guard(scalars.device());
    CHECK_RC(b200::clip_scalars(sumsq_in.data_ptr<float>(), scalars.data_ptr<float>(), (float)loss_scale, (float)clip,
                                cur_stream()),
             "b200::clip_scalars");
}

// ---------------------------------------------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------------------------------------------
static void fill_qkv(b200::AttnDesc& d, const Tensor& q, const Tensor& k, const Tensor& v) {
    // q: [T, H, D] or the grouped view [T, Hkv, q_per_kv, D] of a packed qkv buffer; k, v: [T, Hkv, D]
    TORCH_CHECK((q.dim() == 3 || q.dim() == 4) && k.dim() == 3 && v.dim() == 3, "attn: q [T,H,D] | [T,Hkv,qpk,D], k/v [T,Hkv,D]");
    TORCH_CHECK(q.stride(-1) == 1 && k.stride(2) == 1 && v.stride(2) == 1, "attn: last dim must be contiguous");
    d.q = q.data_ptr(); d.k = k.data_ptr(); d.v = v.data_ptr();
    d.T = q.size(0); d.Hkv = k.size(1); d.D = q.size(-1);
    d.q_stride_t = q.stride(0);
    if (q.dim() == 4) {
        d.H = q.size(1) * q.size(2); d.q_stride_g = q.stride(1); d.q_stride_h = q.stride(2);
    } else {
        d.H = q.size(1); d.q_stride_h = q.stride(1); d.q_stride_g = q.stride(1) * (d.H / d.Hkv);
    }
    d.k_stride_t = k.stride(0); d.k_stride_h = k.stride(1);
    d.v_stride_t = v.stride(0); d.v_stride_h = v.stride(1);
    TORCH_CHECK(d.q_stride_t % 8 == 0 && d.k_stride_t % 8 == 0 && d.v_stride_t % 8 == 0, "attn: token strides must be multiples of 8");
    // TMA descriptors address whole token rows starting at the tensor's data pointer: that pointer must be 16B aligned
    TORCH_CHECK((reinterpret_cast<uintptr_t>(d.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.k) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d.v) & 15) == 0, "attn: q/k/v must be 16-byte aligned");
}

void attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, Tensor& out, Tensor& lse, const Tensor& cu_seqlens,
              int64_t max_seqlen, double scale, bool causal) {
    CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_BF16(out);
    TORCH_CHECK(out.is_contiguous(), "attn: out must be contiguous [T, H, D]");
    TORCH_CHECK(cu_seqlens.scalar_type() == at::kInt && cu_seqlens.is_contiguous(), "attn: cu_seqlens int32");
    c10::cuda::CUDAGuard guard(q.device());
    b200::AttnDesc d;
    fill_qkv(d, q, k, v);
    d.o = out.data_ptr(); d.lse = lse.data_ptr<float>();
    d.cu_seqlens = cu_seqlens.data_ptr<int>(); d.num_seqs = cu_seqlens.numel() - 1; d.max_seqlen = max_seqlen;
    d.scale = (float)scale; d.causal = causal;
    CHECK_RC(b200::attn_fwd(d, cur_stream()), "b200::attn_fwd");
}

void attn_bwd(const Tensor& dout, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& out, const Tensor& lse,
              Tensor& dq, Tensor& dk, Tensor& dv, Tensor& delta, Tensor& dq_acc, const Tensor& cu_seqlens,
              int64_t max_seqlen, double scale, bool causal) {
    CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_BF16(dout);
    c10::cuda::CUDAGuard guard(q.device());
    b200::AttnBwdDesc d;
    fill_qkv(d.f, q, k, v);
    d.f.o = const_cast<void*>(out.data_ptr());
    d.f.lse = const_cast<float*>(lse.data_ptr<float>());
    d.f.cu_seqlens = cu_seqlens.data_ptr<int>(); d.f.num_seqs = cu_seqlens.numel() - 1; d.f.max_seqlen = max_seqlen;
    d.f.scale = (float)scale; d.f.causal = causal;
    TORCH_CHECK(dout.is_contiguous() && out.is_contiguous(), "attn_bwd: dout/out contiguous");
    TORCH_CHECK(dq.stride(-1) == 1 && dk.stride(2) == 1 && dv.stride(2) == 1, "attn_bwd: grads last dim contiguous");
    d.dout = dout.data_ptr();
    d.dq = dq.data_ptr(); d.dq_stride_t = dq.stride(0);
    if (dq.dim() == 4) { d.dq_stride_g = dq.stride(1); d.dq_stride_h = dq.stride(2); }
    else { d.dq_stride_h = dq.stride(1); d.dq_stride_g = dq.stride(1) * (d.f.H / d.f.Hkv); }
    d.dk = dk.data_ptr(); d.dk_stride_t = dk.stride(0); d.dk_stride_h = dk.stride(1);
    d.dv = dv.data_ptr(); d.dv_stride_t = dv.stride(0); d.dv_stride_h = dv.stride(1);
    d.delta = delta.data_ptr<float>();
    d.dq_acc = dq_acc.data_ptr<float>();
    CHECK_RC(b200::attn_bwd(d, cur_stream()), "b200::attn_bwd");
}

// ---- sequence-parallel attention: K / V (forward) and Q / dO / stats / dQ accumulator (backward) of every rank are read
// (or reduce-added) in place through peer pointers of symmetric buffers; q / out / lse hold this rank's token rows, cu_seqlens
// are global.  `k`, `v` (forward) and `q`, `dout`, `delta`, `dq_acc` (backward) must BE this rank's slices of those buffers.
static std::vector<const void*> as_ptrs(at::IntArrayRef v) {
    std::vector<const void*> out;
    for (int64_t p : v) out.push_back(reinterpret_cast<const void*>(p));
    return out;
}

void attn_fwd_sp(const Tensor& q, const Tensor& k, const Tensor& v, Tensor& out, Tensor& lse, const Tensor& cu_seqlens,
                 int64_t max_seqlen, double scale, bool causal, int64_t sp_rank, at::IntArrayRef k_peers,
                 at::IntArrayRef v_peers) {
    CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_BF16(out);
    TORCH_CHECK(out.is_contiguous(), "attn: out must be contiguous [T, H, D]");
    TORCH_CHECK(cu_seqlens.scalar_type() == at::kInt && cu_seqlens.is_contiguous(), "attn: cu_seqlens int32");
    TORCH_CHECK(k_peers.size() == v_peers.size() && k_peers.size() >= 2 && k_peers.size() <= 8, "attn_fwd_sp: 2..8 peers");
    c10::cuda::CUDAGuard guard(q.device());
    b200::AttnDesc d;
    fill_qkv(d, q, k, v);
    d.o = out.data_ptr(); d.lse = lse.data_ptr<float>();
    d.cu_seqlens = cu_seqlens.data_ptr<int>(); d.num_seqs = cu_seqlens.numel() - 1; d.max_seqlen = max_seqlen;
    d.scale = (float)scale; d.causal = causal;
    auto kp = as_ptrs(k_peers), vp = as_ptrs(v_peers);
    d.sp_rank = (int)sp_rank; d.sp_world = (int)k_peers.size(); d.k_peers = kp.data(); d.v_peers = vp.data();
    CHECK_RC(b200::attn_fwd(d, cur_stream()), "b200::attn_fwd_sp");
}

void attn_bwd_sp(const Tensor& dout, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& out, const Tensor& lse,
                 Tensor& dq, Tensor& dk, Tensor& dv, Tensor& delta, Tensor& dq_acc, const Tensor& cu_seqlens,
                 int64_t max_seqlen, double scale, bool causal, int64_t phase, int64_t sp_rank, at::IntArrayRef q_peers,
                 at::IntArrayRef dout_peers, at::IntArrayRef dq_acc_peers, at::IntArrayRef delta_peers) {
    CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_BF16(dout);
    c10::cuda::CUDAGuard guard(q.device());
    b200::AttnBwdDesc d;
    fill_qkv(d.f, q, k, v);
    d.f.o = const_cast<void*>(out.data_ptr());
    d.f.lse = const_cast<float*>(lse.data_ptr<float>());
    d.f.cu_seqlens = cu_seqlens.data_ptr<int>(); d.f.num_seqs = cu_seqlens.numel() - 1; d.f.max_seqlen = max_seqlen;
    d.f.scale = (float)scale; d.f.causal = causal;
    TORCH_CHECK(dout.is_contiguous() && out.is_contiguous(), "attn_bwd: dout/out contiguous");
    TORCH_CHECK(dq.stride(-1) == 1 && dk.stride(2) == 1 && dv.stride(2) == 1, "attn_bwd: grads last dim contiguous");
    d.dout = dout.data_ptr();
    d.dq = dq.data_ptr(); d.dq_stride_t = dq.stride(0);
    if (dq.dim() == 4) { d.dq_stride_g = dq.stride(1); d.dq_stride_h = dq.stride(2); }
    else { d.dq_stride_h = dq.stride(1); d.dq_stride_g = dq.stride(1) * (d.f.H / d.f.Hkv); }
    d.dk = dk.data_ptr(); d.dk_stride_t = dk.stride(0); d.dk_stride_h = dk.stride(1);
    d.dv = dv.data_ptr(); d.dv_stride_t = dv.stride(0); d.dv_stride_h = dv.stride(1);
    d.delta = delta.data_ptr<float>();
    d.dq_acc = dq_acc.data_ptr<float>();
    const size_t W = q_peers.size();
    TORCH_CHECK(W >= 2 && W <= 8 && dout_peers.size() == W && dq_acc_peers.size() == W && delta_peers.size() == W,
                "attn_bwd_sp: 2..8 peers");
    auto qp = as_ptrs(q_peers), dop = as_ptrs(dout_peers), dlp = as_ptrs(delta_peers);
    std::vector<void*> dqp;
    for (int64_t p : dq_acc_peers) dqp.push_back(reinterpret_cast<void*>(p));
    d.f.sp_rank = (int)sp_rank; d.f.sp_world = (int)W;
    d.phase = (int)phase; d.q_peers = qp.data(); d.dout_peers = dop.data(); d.dq_acc_peers = dqp.data(); d.delta_peers = dlp.data();
    CHECK_RC(b200::attn_bwd(d, cur_stream()), "b200::attn_bwd_sp");
}

// ---------------------------------------------------------------------------------------------------------------
// peer-memory communication kernels (pointers come from the symmetric heap, see internevo_b200/parallel/symm.py)
// ---------------------------------------------------------------------------------------------------------------
void symm_barrier(int64_t flags_ptrs, int64_t rank, int64_t world, int64_t epoch) {
    CHECK_RC(b200::symm_barrier(reinterpret_cast<uint32_t* const*>(flags_ptrs), rank, world, (uint32_t)epoch, cur_stream()),
             "b200::symm_barrier");
}

void peer_copy_bench(const Tensor& src, int64_t dst_ptr, int64_t bytes, int64_t piece_bytes, int64_t unroll, int64_t ctas) {
    c10::cuda::CUDAGuard guard(src.device());
    CHECK_RC(b200::peer_copy_bench(src.data_ptr(), reinterpret_cast<void*>(dst_ptr), bytes, piece_bytes, (int)unroll, (int)ctas,
                                   cur_stream()), "b200::peer_copy_bench");
}

void reduce_scatter_adam(int64_t grad_ptrs, int64_t param_ptrs, int64_t flags_ptrs, int64_t rank, int64_t world,
                         int64_t epoch, int64_t shard_off, int64_t shard_n, Tensor& p, Tensor& m, Tensor& v,
                         const Tensor& scalars, double lr, double beta1, double beta2, double eps, double wd, double bc1,
                         double bc2, double grad_div, int64_t phase) {
    c10::cuda::CUDAGuard guard(p.device());
    b200::RsAdamDesc d;
    d.grad_ptrs = reinterpret_cast<void* const*>(grad_ptrs);
    d.param_ptrs = reinterpret_cast<void* const*>(param_ptrs);
    d.flags_ptrs = reinterpret_cast<uint32_t* const*>(flags_ptrs);
    d.rank = rank; d.world = world; d.epoch = (uint32_t)epoch;
    d.shard_off = shard_off; d.shard_n = shard_n;
    d.p = p.data_ptr<float>(); d.m = m.data_ptr<float>(); d.v = v.data_ptr<float>();
    d.scalars = scalars.data_ptr<float>();
    d.lr = lr; d.beta1 = beta1; d.beta2 = beta2; d.eps = eps; d.wd = wd; d.bc1 = bc1; d.bc2 = bc2; d.grad_div = grad_div;
    d.phase = phase;
    CHECK_RC(b200::reduce_scatter_adam(d, cur_stream()), "b200::reduce_scatter_adam");
}

// NVLS form of the two ZeRO phases: `grad_mc` / `param_mc` are the NVSwitch multicast addresses of the gradient / parameter
// arenas, `grad_local` this rank's own mapping of the gradient arena.
void reduce_scatter_adam_mc(int64_t grad_mc, int64_t param_mc, int64_t grad_local, int64_t world, int64_t shard_off,
                            int64_t shard_n, Tensor& p, Tensor& m, Tensor& v, const Tensor& scalars, double lr, double beta1,
                            double beta2, double eps, double wd, double bc1, double bc2, double grad_div, int64_t phase) {
    c10::cuda::CUDAGuard guard(p.device());
    TORCH_CHECK(grad_mc != 0 && param_mc != 0 && grad_local != 0, "reduce_scatter_adam_mc: null multicast mapping");
    b200::RsAdamDesc d;
    d.grad_mc = reinterpret_cast<const void*>(grad_mc);
    d.param_mc = reinterpret_cast<void*>(param_mc);
    d.grad_local = reinterpret_cast<void*>(grad_local);
    d.world = world;
    d.shard_off = shard_off; d.shard_n = shard_n;
    d.p = p.data_ptr<float>(); d.m = m.data_ptr<float>(); d.v = v.data_ptr<float>();
    d.scalars = scalars.data_ptr<float>();
    d.lr = lr; d.beta1 = beta1; d.beta2 = beta2; d.eps = eps; d.wd = wd; d.bc1 = bc1; d.bc2 = bc2; d.grad_div = grad_div;
    d.phase = phase;
    CHECK_RC(b200::reduce_scatter_adam(d, cur_stream()), "b200::reduce_scatter_adam_mc");
}

// reduce_scatter(a @ b^T) over rows (mode 0: `out` is this rank's [M/world, N]) or all-reduce (mode 1: `out` is the
// [M, N] symmetric output, out_ptrs its per-rank pointer table).  stage_ptrs: per-rank symmetric staging
// [world, M/world, N] that the peers' epilogues push their partial tiles into.
void gemm_rs(const Tensor& a, const Tensor& b, Tensor& out, int64_t stage_ptrs, int64_t out_ptrs, int64_t flags_ptrs,
             int64_t rank, int64_t world, int64_t epoch, bool b_mn, int64_t mode, int64_t done_ptrs,
             const optional<Tensor>& done_counter) {
    CHECK_BF16(a); CHECK_BF16(b); CHECK_BF16(out);
    c10::cuda::CUDAGuard guard(a.device());
    b200::GemmCommDesc d;
    d.g.M = a.size(0); d.g.K = a.size(1); d.g.N = b_mn ? b.size(1) : b.size(0);
    TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == d.g.K, "gemm_rs: K mismatch");
    TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1 && out.stride(1) == 1, "gemm_rs: unit inner strides");
    TORCH_CHECK(out.size(1) == d.g.N && out.size(0) == (mode == 1 ? d.g.M : d.g.M / world), "gemm_rs: out shape");
    d.g.A = a.data_ptr(); d.g.lda = a.stride(0); d.g.a_mn_major = 0;
    d.g.B = b.data_ptr(); d.g.ldb = b.stride(0); d.g.b_mn_major = b_mn;
    d.g.D = nullptr; d.g.ldd = d.g.N;
    d.peer_ptrs = reinterpret_cast<void* const*>(stage_ptrs);
    d.out_ptrs = reinterpret_cast<void* const*>(out_ptrs);
    d.flags_ptrs = reinterpret_cast<uint32_t* const*>(flags_ptrs);
    d.rank = rank; d.world = world; d.epoch = (uint32_t)epoch; d.mode = mode;
    d.out_local = out.data_ptr(); d.ld_out = out.stride(0);
    d.done_ptrs = reinterpret_cast<uint32_t* const*>(done_ptrs);
    if (done_counter.has_value()) {
        TORCH_CHECK(done_counter->is_cuda() && done_counter->scalar_type() == at::kInt, "gemm_rs: done_counter must be cuda int32");
        d.done_counter = reinterpret_cast<uint32_t*>(done_counter->data_ptr<int>());
    }
    CHECK_RC(b200::gemm_reduce_scatter(d, cur_stream()), "b200::gemm_rs");
}

// out = all_gather(x shards) @ b^T.  `gathered` is this rank's symmetric [M, K] buffer (gathered_ptrs its per-rank
// pointer table): copy CTAs of the same launch push x_local into every rank's buffer while the GEMM CTAs start on the
// local rows (read in place from x_local).
void ag_gemm(const Tensor& x_local, int64_t gathered_ptrs, int64_t flags_ptrs, int64_t rank, int64_t world,
             int64_t epoch, const Tensor& b, bool b_mn, Tensor& gathered, Tensor& out, int64_t flags,
             const optional<Tensor>& h, int64_t comm_ctas) {
    CHECK_BF16(b); CHECK_BF16(x_local); CHECK_BF16(gathered); CHECK_BF16(out);
    c10::cuda::CUDAGuard guard(b.device());
    b200::GemmCommDesc d;
    d.m_local = x_local.size(0);
    d.g.M = d.m_local * world; d.g.K = x_local.size(1); d.g.N = b_mn ? b.size(1) : b.size(0);
    TORCH_CHECK(x_local.is_contiguous() && gathered.is_contiguous(), "ag_gemm: contiguous shard / gathered buffer");
    TORCH_CHECK(gathered.size(0) == d.g.M && gathered.size(1) == d.g.K, "ag_gemm: gathered shape");
    TORCH_CHECK(out.size(0) == d.g.M && out.size(1) == d.g.N, "ag_gemm: out shape");
    d.g.A = gathered.data_ptr(); d.g.lda = d.g.K; d.g.a_mn_major = 0;
    d.g.B = b.data_ptr(); d.g.ldb = b.stride(0); d.g.b_mn_major = b_mn;
    d.g.D = out.data_ptr(); d.g.ldd = out.stride(0);
    d.g.flags = flags;
    if (h.has_value()) { d.g.H = h->data_ptr(); d.g.ldh = h->stride(0); }
    d.peer_ptrs = reinterpret_cast<void* const*>(gathered_ptrs);
    d.flags_ptrs = reinterpret_cast<uint32_t* const*>(flags_ptrs);
    d.rank = rank; d.world = world; d.epoch = (uint32_t)epoch;
    d.out_local = gathered.data_ptr(); d.ld_out = d.g.K;
    d.x_local = x_local.data_ptr();
    d.comm_ctas = comm_ctas;
    CHECK_RC(b200::allgather_gemm(d, cur_stream()), "b200::ag_gemm");
}

}  // namespace

// ---- MoE dispatch / combine over peer memory
void symm_allgather_small(int64_t buf_ptrs, const Tensor& src, int64_t flags_ptrs, int64_t rank, int64_t world, int64_t epoch) {
    c10::cuda::CUDAGuard guard(src.device());
    TORCH_CHECK(src.is_cuda() && src.is_contiguous() && src.element_size() == 4, "symm_allgather_small: 32-bit contiguous src");
    CHECK_RC(b200::symm_allgather_small(reinterpret_cast<uint32_t* const*>(buf_ptrs),
                                        reinterpret_cast<const uint32_t*>(src.data_ptr()), (int)src.numel(),
                                        reinterpret_cast<uint32_t* const*>(flags_ptrs), rank, world, (uint32_t)epoch,
                                        cur_stream()),
             "b200::symm_allgather_small");
}

static void check_slots(const Tensor& slot_rank, const Tensor& slot_row, const c10::optional<Tensor>& scale, const char* who) {
    TORCH_CHECK(slot_rank.scalar_type() == at::kInt && slot_row.scalar_type() == at::kInt && slot_rank.is_contiguous() &&
                    slot_row.is_contiguous() && slot_rank.numel() == slot_row.numel(),
                who, ": slot_rank / slot_row must be contiguous int32 of equal length");
    if (scale.has_value())
        TORCH_CHECK(scale->scalar_type() == at::kFloat && scale->is_contiguous() && scale->numel() == slot_row.numel(), who,
                    ": scale must be contiguous fp32 [n_slots]");
}

void moe_scatter_rows(const Tensor& x, const Tensor& slot_rank, const Tensor& slot_row, const c10::optional<Tensor>& scale,
                      int64_t x_ptrs, int64_t y_ptrs, c10::optional<Tensor> dw, int64_t k) {
    c10::cuda::CUDAGuard guard(x.device());
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.stride(1) == 1, "moe_scatter_rows: x bf16 [T, H]");
    check_slots(slot_rank, slot_row, scale, "moe_scatter_rows");
    TORCH_CHECK(slot_row.numel() == x.size(0) * k, "moe_scatter_rows: n_slots != T * k");
    TORCH_CHECK(x_ptrs != 0, "moe_scatter_rows: null destination table");
    b200::MoeCommDesc d;
    d.x = x.data_ptr(); d.ldx = x.stride(0);
    d.slot_rank = slot_rank.data_ptr<int>(); d.slot_row = slot_row.data_ptr<int>();
    d.scale = scale.has_value() ? scale->data_ptr<float>() : nullptr;
    d.x_ptrs = reinterpret_cast<void* const*>(x_ptrs);
    d.y_ptrs = reinterpret_cast<void* const*>(y_ptrs);
    if (dw.has_value()) {
        TORCH_CHECK(dw->scalar_type() == at::kFloat && dw->is_contiguous() && dw->numel() == slot_row.numel() && y_ptrs != 0,
                    "moe_scatter_rows: dw fp32 [n_slots] needs y_ptrs");
        d.dw = dw->data_ptr<float>();
    }
    d.n_slots = (int)slot_row.numel(); d.k = (int)k; d.H = (int)x.size(1);
    CHECK_RC(b200::moe_scatter_rows(d, cur_stream()), "b200::moe_scatter_rows");
}

void moe_gather_combine(Tensor& out, const c10::optional<Tensor>& w, const Tensor& slot_rank, const Tensor& slot_row,
                        int64_t y_ptrs, int64_t k) {
    c10::cuda::CUDAGuard guard(out.device());
    TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kBFloat16 && out.dim() == 2 && out.stride(1) == 1,
                "moe_gather_combine: out bf16 [T, H]");
    check_slots(slot_rank, slot_row, w, "moe_gather_combine");
    TORCH_CHECK(slot_row.numel() == out.size(0) * k && y_ptrs != 0, "moe_gather_combine: n_slots != T * k or null table");
    b200::MoeCommDesc d;
    d.out = out.data_ptr(); d.ldo = out.stride(0);
    d.slot_rank = slot_rank.data_ptr<int>(); d.slot_row = slot_row.data_ptr<int>();
    d.scale = w.has_value() ? w->data_ptr<float>() : nullptr;
    d.y_ptrs = reinterpret_cast<void* const*>(y_ptrs);
    d.n_slots = (int)slot_row.numel(); d.k = (int)k; d.H = (int)out.size(1);
    CHECK_RC(b200::moe_gather_combine(d, cur_stream()), "b200::moe_gather_combine");
}

// out = x @ all_gather(w_shard)^T (b_mn = false: w_shard [N / W, K]) or x @ all_gather(w_shard) (b_mn = true: x is
// [M, N_total], w_shard [N_total / W, K_in], the gathered rows are the contraction).  `gathered` is this rank's symmetric
// [N_total, K_in] buffer: every rank's epilogue warps push their shard into it while the tensor cores start on the own shard.
void gather_weight_gemm(const Tensor& x, const Tensor& w_shard, int64_t gathered_ptrs, Tensor& gathered, int64_t flags_ptrs,
                        int64_t rank, int64_t world, int64_t epoch, bool b_mn, Tensor& out) {
    CHECK_BF16(x); CHECK_BF16(w_shard); CHECK_BF16(gathered); CHECK_BF16(out);
    TORCH_CHECK(x.dim() == 2 && w_shard.dim() == 2 && x.stride(1) == 1 && w_shard.stride(1) == 1 && out.stride(1) == 1 &&
                    gathered.is_contiguous(), "gather_weight_gemm: 2-D operands with unit inner stride");
    c10::cuda::CUDAGuard guard(x.device());
    b200::GemmCommDesc d;
    const int64_t rows = w_shard.size(0), kin = w_shard.size(1);
    TORCH_CHECK(gathered.size(0) == rows * world && gathered.size(1) == kin, "gather_weight_gemm: gathered buffer shape");
    d.g.M = x.size(0);
    if (!b_mn) { d.g.K = kin; d.g.N = rows * world; TORCH_CHECK(x.size(1) == kin, "gather_weight_gemm: K mismatch"); }
    else       { d.g.K = rows * world; d.g.N = kin; TORCH_CHECK(x.size(1) == rows * world, "gather_weight_gemm: K mismatch"); }
    TORCH_CHECK(out.size(0) == d.g.M && out.size(1) == d.g.N, "gather_weight_gemm: out shape");
    d.g.A = x.data_ptr(); d.g.lda = x.stride(0); d.g.a_mn_major = 0;
    d.g.B = w_shard.data_ptr(); d.g.ldb = w_shard.stride(0); d.g.b_mn_major = b_mn;
    d.g.D = out.data_ptr(); d.g.ldd = out.stride(0);
    d.peer_ptrs = reinterpret_cast<void* const*>(gathered_ptrs);
    d.flags_ptrs = reinterpret_cast<uint32_t* const*>(flags_ptrs);
    d.rank = rank; d.world = world; d.epoch = (uint32_t)epoch;
    d.out_local = gathered.data_ptr(); d.m_local = rows; d.x_local = w_shard.data_ptr();
    CHECK_RC(b200::gather_weight_gemm(d, cur_stream()), "b200::gather_weight_gemm");
}

// out[N_total / W, K_in] (+)= scale * reduce_scatter(dy^T @ x) over the rows: the weight-parallel wgrad.  dy [T, N_total],
// x [T, K_in]; stage_ptrs: per-rank symmetric staging [world, N_total / W, K_in].
void wgrad_rs(const Tensor& dy, const Tensor& x, Tensor& out, int64_t stage_ptrs, int64_t flags_ptrs, int64_t rank,
              int64_t world, int64_t epoch, double scale, bool accumulate) {
    CHECK_BF16(dy); CHECK_BF16(x); CHECK_BF16(out);
    TORCH_CHECK(dy.dim() == 2 && x.dim() == 2 && dy.size(0) == x.size(0) && dy.stride(1) == 1 && x.stride(1) == 1 &&
                    out.stride(1) == 1, "wgrad_rs: shapes");
    c10::cuda::CUDAGuard guard(dy.device());
    b200::GemmCommDesc d;
    d.g.M = dy.size(1); d.g.K = dy.size(0); d.g.N = x.size(1);
    TORCH_CHECK(out.size(0) * world == d.g.M && out.size(1) == d.g.N, "wgrad_rs: out shape");
    d.g.A = dy.data_ptr(); d.g.lda = dy.stride(0); d.g.a_mn_major = 1;
    d.g.B = x.data_ptr(); d.g.ldb = x.stride(0); d.g.b_mn_major = 1;
    d.g.D = nullptr; d.g.ldd = d.g.N;
    d.g.flags = accumulate ? b200::GEMM_ACCUMULATE : 0;
    d.peer_ptrs = reinterpret_cast<void* const*>(stage_ptrs);
    d.flags_ptrs = reinterpret_cast<uint32_t* const*>(flags_ptrs);
    d.rank = rank; d.world = world; d.epoch = (uint32_t)epoch; d.mode = 0;
    d.out_local = out.data_ptr(); d.ld_out = out.stride(0); d.out_scale = static_cast<float>(scale);
    CHECK_RC(b200::gemm_reduce_scatter(d, cur_stream()), "b200::wgrad_rs");
}

// ---- grouped GEMM (MoE experts) -----------------------------------------------------------------------------------------
// 128-byte tensor maps of every expert's weight, built on the host and uploaded once (the weights live in the optimizer's
// arena: their addresses never change), returned as a uint8 cuda tensor [G * 128].
Tensor grouped_b_maps(at::TensorList weights, bool b_mn) {
    TORCH_CHECK(!weights.empty(), "grouped_b_maps: no weights");
    const int64_t G = static_cast<int64_t>(weights.size());
    auto host = at::empty({G * 128}, at::TensorOptions().dtype(at::kByte).pinned_memory(true));
    for (int64_t g = 0; g < G; ++g) {
        const Tensor& w = weights[g];
        CHECK_BF16(w);
        TORCH_CHECK(w.dim() == 2 && w.stride(1) == 1 && w.stride(0) % 8 == 0, "grouped_b_maps: 2-D weights, unit inner stride");
        const int N = b_mn ? w.size(1) : w.size(0), K = b_mn ? w.size(0) : w.size(1);
        CUtensorMap m;
        CHECK_RC(b200::grouped_b_map(&m, w.data_ptr(), N, K, w.stride(0), b_mn), "b200::grouped_b_map");
        std::memcpy(static_cast<uint8_t*>(host.data_ptr()) + g * 128, &m, sizeof(CUtensorMap));
    }
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    return host.to(weights[0].device(), /*non_blocking=*/false);
}

// rows of group g: offsets[g] .. offsets[g + 1] (int32, device).  out[rows, N] = a[rows, K] @ W_g^T (b_mn: @ W_g)
void grouped_gemm(const Tensor& a, const Tensor& bmaps, const Tensor& offsets, Tensor& out, int64_t N, bool b_mn, int64_t flags,
                  const optional<Tensor>& h) {
    CHECK_BF16(a); CHECK_BF16(out);
    TORCH_CHECK(offsets.is_cuda() && offsets.scalar_type() == at::kInt && offsets.is_contiguous(), "grouped_gemm: int32 offsets");
    TORCH_CHECK(bmaps.is_cuda() && bmaps.scalar_type() == at::kByte && bmaps.numel() == (offsets.numel() - 1) * 128,
                "grouped_gemm: one tensor map per group");
    TORCH_CHECK((reinterpret_cast<uintptr_t>(bmaps.data_ptr()) & 63) == 0, "grouped_gemm: tensor maps must be 64-byte aligned");
    TORCH_CHECK(a.dim() == 2 && out.dim() == 2 && a.stride(1) == 1 && out.stride(1) == 1 && out.size(0) == a.size(0) &&
                    out.size(1) == N, "grouped_gemm: shapes");
    c10::cuda::CUDAGuard guard(a.device());
    b200::GroupedGemmDesc g;
    g.mode = 1; g.num_groups = static_cast<int>(offsets.numel() - 1); g.grp_off = offsets.data_ptr<int>();
    g.rows_cap = a.size(0); g.N = static_cast<int>(N); g.K = static_cast<int>(a.size(1));
    g.A = a.data_ptr(); g.lda = a.stride(0); g.b_mn_major = b_mn;
    g.b_maps = reinterpret_cast<const CUtensorMap*>(bmaps.data_ptr());
    g.D = out.data_ptr(); g.ldd = out.stride(0); g.flags = static_cast<int>(flags);
    if (flags & b200::GEMM_SWIGLU) {
        TORCH_CHECK(h.has_value() && h->scalar_type() == at::kBFloat16 && h->size(0) == a.size(0) && h->size(1) == N / 2 &&
                        h->stride(1) == 1, "grouped_gemm: bad swiglu output");
        g.H = h->data_ptr(); g.ldh = h->stride(0);
    }
    CHECK_RC(b200::gemm_bf16_grouped(g, cur_stream()), "b200::grouped_gemm");
}

// dW_g [M, N] (+)= dy[rows_g, M]^T @ x[rows_g, N]; d_ptrs: int64 cuda tensor with one output pointer per group
void grouped_wgrad(const Tensor& dy, const Tensor& x, const Tensor& offsets, const Tensor& d_ptrs, int64_t ldd, int64_t flags) {
    CHECK_BF16(dy); CHECK_BF16(x);
    TORCH_CHECK(offsets.is_cuda() && offsets.scalar_type() == at::kInt && offsets.is_contiguous(), "grouped_wgrad: int32 offsets");
    TORCH_CHECK(d_ptrs.is_cuda() && d_ptrs.scalar_type() == at::kLong && d_ptrs.numel() == offsets.numel() - 1,
                "grouped_wgrad: one output pointer per group");
    TORCH_CHECK(dy.dim() == 2 && x.dim() == 2 && dy.size(0) == x.size(0) && dy.stride(1) == 1 && x.stride(1) == 1,
                "grouped_wgrad: shapes");
    c10::cuda::CUDAGuard guard(dy.device());
    b200::GroupedGemmDesc g;
    g.mode = 2; g.num_groups = static_cast<int>(offsets.numel() - 1); g.grp_off = offsets.data_ptr<int>();
    g.rows_cap = dy.size(0); g.M = static_cast<int>(dy.size(1)); g.N = static_cast<int>(x.size(1));
    g.A = dy.data_ptr(); g.lda = dy.stride(0); g.B = x.data_ptr(); g.ldb = x.stride(0);
    g.d_ptrs = reinterpret_cast<void* const*>(d_ptrs.data_ptr<int64_t>()); g.ldd = ldd;
    g.flags = static_cast<int>(flags);
    CHECK_RC(b200::gemm_bf16_grouped(g, cur_stream()), "b200::grouped_wgrad");
}

TORCH_LIBRARY(b200, m) {
    m.def("gemm(Tensor a, Tensor b, Tensor(a!) out, bool a_mn, bool b_mn, Tensor? bias, int flags, Tensor? h, int force_bn, int max_ctas) -> ()", &gemm);
    m.def("set_gemm_tail_split(int on) -> ()", [](int64_t on) { b200::set_gemm_tail_split(static_cast<int>(on)); });
    m.def("set_gemm_group_m(int g) -> ()", [](int64_t g) { b200::set_gemm_group_m(static_cast<int>(g)); });
    m.def("rmsnorm_fwd(Tensor x, Tensor? res_in, Tensor w, Tensor(a!) y, Tensor? res_out, Tensor? rstd, float eps) -> ()", &rmsnorm_fwd);
    m.def("rmsnorm_bwd_blocks(int rows) -> int", &rmsnorm_bwd_blocks);
    m.def("layernorm_fwd(Tensor x, Tensor? res_in, Tensor? keep, float drop_scale, Tensor w, Tensor? b, Tensor(a!) y, Tensor? res_out, Tensor(b!) mean, Tensor(c!) rstd, float eps) -> ()", &layernorm_fwd);
    m.def("layernorm_bwd_blocks(int rows) -> int", &layernorm_bwd_blocks);
    m.def("layernorm_bwd(Tensor dy, Tensor res, Tensor w, Tensor mean, Tensor rstd, Tensor? dres, Tensor(a!) dx, Tensor(b!) partial, Tensor(c!) dwdb) -> ()", &layernorm_bwd);
    m.def("rmsnorm_bwd(Tensor dy, Tensor res, Tensor w, Tensor rstd, Tensor? dres, Tensor(a!) dx, Tensor(b!) dw_partial, Tensor(c!) dw, bool accumulate) -> ()", &rmsnorm_bwd);
    m.def("rope(Tensor(a!) x, Tensor? pos, Tensor cos_t, Tensor sin_t, int group, int rot_per_group, bool conj, bool interleaved) -> ()", &rope);
    m.def("swiglu_fwd(Tensor gu, Tensor(a!) h) -> ()", &swiglu_fwd);
    m.def("swiglu_bwd(Tensor dh, Tensor gu, Tensor(a!) dgu) -> ()", &swiglu_bwd);
    m.def("gelu_bwd(Tensor dh, Tensor pre, Tensor(a!) dpre) -> ()", &gelu_bwd);
    m.def("attn_decode(Tensor q, Tensor kcache, Tensor vcache, Tensor(a!) out, Tensor(b!) work, Tensor(c!) tickets, int seqlen, int nsplit, float scale, Tensor? seqlen_dev=None) -> ()", &attn_decode);
    m.def("ce_fwd(Tensor logits, Tensor labels, int vocab_start, Tensor(a!) out_max, Tensor(b!) out_sum, Tensor(c!) out_sumx, Tensor(d!) out_tgt) -> ()", &ce_fwd);
    m.def("ce_bwd(Tensor(a!) logits, Tensor labels, Tensor lse, Tensor gscale, int vocab_start, float smoothing, int total_classes, int ignore_index) -> ()", &ce_bwd);
    m.def("adamw(Tensor(a!) p, Tensor(b!) m, Tensor(c!) v, Tensor g, Tensor? p_lp, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2, Tensor? scalars) -> ()", &adamw);
    m.def("sumsq(Tensor g, Tensor(a!) out) -> ()", &sumsq);
    m.def("clip_scalars(Tensor sumsq_in, Tensor(a!) scalars, float loss_scale, float clip) -> ()", &clip_scalars);
    m.def("attn_fwd(Tensor q, Tensor k, Tensor v, Tensor(a!) out, Tensor(b!) lse, Tensor cu_seqlens, int max_seqlen, float scale, bool causal) -> ()", &attn_fwd);
    m.def("attn_bwd(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor lse, Tensor(a!) dq, Tensor(b!) dk, Tensor(c!) dv, Tensor(d!) delta, Tensor(e!) dq_acc, Tensor cu_seqlens, int max_seqlen, float scale, bool causal) -> ()", &attn_bwd);
    m.def("peer_copy_bench(Tensor src, int dst_ptr, int bytes, int piece_bytes, int unroll, int ctas) -> ()", &peer_copy_bench);
    m.def("attn_fwd_sp(Tensor q, Tensor k, Tensor v, Tensor(a!) out, Tensor(b!) lse, Tensor cu_seqlens, int max_seqlen, float scale, bool causal, int sp_rank, int[] k_peers, int[] v_peers) -> ()", &attn_fwd_sp);
    m.def("attn_bwd_sp(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor lse, Tensor(a!) dq, Tensor(b!) dk, Tensor(c!) dv, Tensor(d!) delta, Tensor(e!) dq_acc, Tensor cu_seqlens, int max_seqlen, float scale, bool causal, int phase, int sp_rank, int[] q_peers, int[] dout_peers, int[] dq_acc_peers, int[] delta_peers) -> ()", &attn_bwd_sp);
    m.def("symm_barrier(int flags_ptrs, int rank, int world, int epoch) -> ()", &symm_barrier);
    m.def("reduce_scatter_adam(int grad_ptrs, int param_ptrs, int flags_ptrs, int rank, int world, int epoch, int shard_off, int shard_n, Tensor(a!) p, Tensor(b!) m, Tensor(c!) v, Tensor scalars, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2, float grad_div, int phase) -> ()", &reduce_scatter_adam);
    m.def("reduce_scatter_adam_mc(int grad_mc, int param_mc, int grad_local, int world, int shard_off, int shard_n, Tensor(a!) p, Tensor(b!) m, Tensor(c!) v, Tensor scalars, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2, float grad_div, int phase) -> ()", &reduce_scatter_adam_mc);
    m.def("symm_allgather_small(int buf_ptrs, Tensor src, int flags_ptrs, int rank, int world, int epoch) -> ()", &symm_allgather_small);
    m.def("moe_scatter_rows(Tensor x, Tensor slot_rank, Tensor slot_row, Tensor? scale, int x_ptrs, int y_ptrs, Tensor(a!)? dw, int k) -> ()", &moe_scatter_rows);
    m.def("moe_gather_combine(Tensor(a!) out, Tensor? w, Tensor slot_rank, Tensor slot_row, int y_ptrs, int k) -> ()", &moe_gather_combine);
    m.def("gather_weight_gemm(Tensor x, Tensor w_shard, int gathered_ptrs, Tensor(a!) gathered, int flags_ptrs, int rank, int world, int epoch, bool b_mn, Tensor(b!) out) -> ()", &gather_weight_gemm);
    m.def("wgrad_rs(Tensor dy, Tensor x, Tensor(a!) out, int stage_ptrs, int flags_ptrs, int rank, int world, int epoch, float scale, bool accumulate) -> ()", &wgrad_rs);
    m.def("grouped_b_maps(Tensor[] weights, bool b_mn) -> Tensor", &grouped_b_maps);
    m.def("grouped_gemm(Tensor a, Tensor bmaps, Tensor offsets, Tensor(a!) out, int N, bool b_mn, int flags, Tensor(b!)? h) -> ()", &grouped_gemm);
    m.def("grouped_wgrad(Tensor dy, Tensor x, Tensor offsets, Tensor d_ptrs, int ldd, int flags) -> ()", &grouped_wgrad);
    m.def("gemm_rs(Tensor a, Tensor b, Tensor(a!) out, int stage_ptrs, int out_ptrs, int flags_ptrs, int rank, int world, int epoch, bool b_mn, int mode, int done_ptrs, Tensor? done_counter) -> ()", &gemm_rs);
    m.def("ag_gemm(Tensor x_local, int gathered_ptrs, int flags_ptrs, int rank, int world, int epoch, Tensor b, bool b_mn, Tensor(a!) gathered, Tensor(b!) out, int flags, Tensor? h, int comm_ctas) -> ()", &ag_gemm);
}
