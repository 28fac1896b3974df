// Bandwidth-bound sm_100a kernels: fused residual-add + RMSNorm (fwd/bwd), in-place RoPE on the packed InternLM2 qkv
// layout, SwiGLU (fwd/bwd, interleaved gate/up), vocab-parallel cross-entropy (online LSE fwd, in-place bwd),
// fused AdamW (unscale + clip + update + bf16 cast-back) and flat L2-norm.  All use 128-bit accesses and fp32 math.
//
// Reference kernels replaced: apex cuApplyRMSNorm / cuComputeGradInput (third_party/apex/csrc/layer_norm_cuda_kernel.cu),
// rotary_emb.apply_rotary (third_party/flash-attention/csrc/rotary/rotary_cuda.cu), the jit-scripted Silu
// (internlm/model/utils.py:684-688), xentropy_cuda_lib (third_party/flash-attention/csrc/xentropy/xentropy_kernel.cu),
// ATen fused AdamW + the flatten/cast/unscale glue (internlm/solver/optimizer/hybrid_zero_optim.py:740-797) and
// apex multi_tensor_l2norm.
#include "elementwise.h"

#include <cuda_bf16.h>
#include <math.h>

#include "launch.h"
#include "sm100_ptx.cuh"

namespace b200 {

// ----------------------------------------------------------------------------------------------------------------
// helpers
// ----------------------------------------------------------------------------------------------------------------
B200_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
B200_DEVICE float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int THREADS>
B200_DEVICE float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < THREADS / 32) ? red[l] : 0.f;
    return warp_sum(t);
}
template <int THREADS>
B200_DEVICE float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < THREADS / 32) ? red[l] : -INFINITY;
    return warp_max(t);
}
B200_DEVICE void unpack8(const uint4& u, float (&f)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
B200_DEVICE uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}

// ----------------------------------------------------------------------------------------------------------------
// RMSNorm: one CTA (256 threads) per row, row cached in registers (H <= 256*8*MAXV)
// ----------------------------------------------------------------------------------------------------------------
static constexpr int RN_THREADS = 256;

// res_out = x (+ res_in);  y = res_out * rstd * w        (RN_MAXV * 2048 >= H)
template <int RN_MAXV>
__global__ void __launch_bounds__(RN_THREADS, RN_MAXV <= 2 ? 3 : 2) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ res_in,
                                                                 const __nv_bfloat16* __restrict__ w,
                                                                 __nv_bfloat16* __restrict__ y,
                                                                 __nv_bfloat16* __restrict__ res_out,
                                                                 float* __restrict__ rstd_out, int rows, int H,
                                                                 float eps) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red[32];
    const int nvec = H / 8;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + (int64_t)row * H);
        const uint4* rr = res_in ? reinterpret_cast<const uint4*>(res_in + (int64_t)row * H) : nullptr;
        float v[RN_MAXV][8];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {
            const int idx = threadIdx.x + i * RN_THREADS;
            if (idx < nvec) {
                unpack8(xr[idx], v[i]);
                if (rr) {
                    float r[8];
                    unpack8(rr[idx], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i][j] += r[j];
                }
                if (res_out) {
                    uint4 o = pack8(v[i]);
                    reinterpret_cast<uint4*>(res_out + (int64_t)row * H)[idx] = o;
                    unpack8(o, v[i]);  // normalise the bf16-rounded residual, as the unfused composition would
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
            }
        }
        ss = block_sum<RN_THREADS>(ss, red);
        const float rstd = rsqrtf(ss / H + eps);
        if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {
            const int idx = threadIdx.x + i * RN_THREADS;
            if (idx < nvec) {
                float wv[8], o[8];
                unpack8(reinterpret_cast<const uint4*>(w)[idx], wv);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rstd * wv[j];
                reinterpret_cast<uint4*>(y + (int64_t)row * H)[idx] = pack8(o);
            }
        }
    }
}

// dx = rstd * (dy*w - xhat * mean(dy*w*xhat)) (+ dres);  dw_partial[block] += dy * xhat
template <int RN_MAXV>
__global__ void __launch_bounds__(RN_THREADS, RN_MAXV == 1 ? 4 : (RN_MAXV == 2 ? 2 : 1)) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                 const __nv_bfloat16* __restrict__ res,
                                                                 const __nv_bfloat16* __restrict__ w,
                                                                 const float* __restrict__ rstd_in,
                                                                 const __nv_bfloat16* __restrict__ dres,
                                                                 __nv_bfloat16* __restrict__ dx,
                                                                 float* __restrict__ dw_partial, int rows, int H) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red[32];
    const int nvec = H / 8;
    float dwacc[RN_MAXV][8];
    float wv[RN_MAXV][8];
#pragma unroll
    for (int i = 0; i < RN_MAXV; ++i) {
        const int idx = threadIdx.x + i * RN_THREADS;
#pragma unroll
        for (int j = 0; j < 8; ++j) dwacc[i][j] = 0.f;
        if (idx < nvec) unpack8(reinterpret_cast<const uint4*>(w)[idx], wv[i]);
    }
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const float rstd = rstd_in[row];
        float g[RN_MAXV][8], xh[RN_MAXV][8];
        uint4 ldy[RN_MAXV], lres[RN_MAXV], ldr[RN_MAXV];
#pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {  // all loads of this row in flight before any use
            const int idx = threadIdx.x + i * RN_THREADS;
            if (idx < nvec) {
                ldy[i] = ld_nc_v4(reinterpret_cast<const uint4*>(dy + (int64_t)row * H) + idx);
                lres[i] = ld_nc_v4(reinterpret_cast<const uint4*>(res + (int64_t)row * H) + idx);
                if (dres) ldr[i] = ld_nc_v4(reinterpret_cast<const uint4*>(dres + (int64_t)row * H) + idx);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {
            const int idx = threadIdx.x + i * RN_THREADS;
            if (idx < nvec) {
                unpack8(ldy[i], g[i]);
                unpack8(lres[i], xh[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] *= rstd;
                    dwacc[i][j] += g[i][j] * xh[i][j];
                    g[i][j] *= wv[i][j];
                    dot += g[i][j] * xh[i][j];
                }
            }
        }
        dot = block_sum<RN_THREADS>(dot, red) / H;
#pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {
            const int idx = threadIdx.x + i * RN_THREADS;
            if (idx < nvec) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - xh[i][j] * dot);
                if (dres) {
                    float r[8];
                    unpack8(ldr[i], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += r[j];
                }
                reinterpret_cast<uint4*>(dx + (int64_t)row * H)[idx] = pack8(o);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < RN_MAXV; ++i) {
        const int idx = threadIdx.x + i * RN_THREADS;
        if (idx < nvec) {
            float* p = dw_partial + (int64_t)blockIdx.x * H + idx * 8;
            reinterpret_cast<float4*>(p)[0] = make_float4(dwacc[i][0], dwacc[i][1], dwacc[i][2], dwacc[i][3]);
            reinterpret_cast<float4*>(p)[1] = make_float4(dwacc[i][4], dwacc[i][5], dwacc[i][6], dwacc[i][7]);
        }
    }
}

// dw[c] (+)= sum_b partial[b, c]; block = 32 columns x 8 row-groups (coalesced 128-byte rows, H/32 blocks)
__global__ void colsum_kernel(const float* __restrict__ partial, float* __restrict__ out_f32,
                              __nv_bfloat16* __restrict__ out_bf16, int nblocks, int H, int accumulate) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < H)
        for (int b = threadIdx.y; b < nblocks; b += 8) s += partial[(int64_t)b * H + c];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < H) {
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
        if (out_f32) out_f32[c] = accumulate ? out_f32[c] + s : s;
        if (out_bf16) out_bf16[c] = __float2bfloat16_rn(accumulate ? __bfloat162float(out_bf16[c]) + s : s);
    }
}

int rmsnorm_fwd(const void* x, const void* res_in, const void* w, void* y, void* res_out, float* rstd, int rows, int H,
                float eps, cudaStream_t s) {
    if (H % 8 != 0 || H > RN_THREADS * 8 * 4) return -1;
    const int grid = rows < 148 * 6 ? rows : 148 * 6;
#define RN_FWD(MV)                                                                                                   \
    launch_pdl(rmsnorm_fwd_kernel<MV>, dim3(grid), dim3(RN_THREADS), 0, s, 1, (const __nv_bfloat16*)x, (const __nv_bfloat16*)res_in,        \
                                                       (const __nv_bfloat16*)w, (__nv_bfloat16*)y,                   \
                                                       (__nv_bfloat16*)res_out, rstd, rows, H, eps)
    if (H <= 2048) RN_FWD(1);
    else if (H <= 4096) RN_FWD(2);
    else RN_FWD(4);
#undef RN_FWD
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int rmsnorm_bwd_blocks(int rows) { return rows < 148 * 2 ? rows : 148 * 2; }

int rmsnorm_bwd(const void* dy, const void* res, const void* w, const float* rstd, const void* dres, void* dx,
                float* dw_partial, float* dw_f32, void* dw_bf16, int accumulate, int rows, int H, cudaStream_t s) {
    if (H % 8 != 0 || H > RN_THREADS * 8 * 4) return -1;
    const int grid = rmsnorm_bwd_blocks(rows);
#define RN_BWD(MV)                                                                                                   \
    launch_pdl(rmsnorm_bwd_kernel<MV>, dim3(grid), dim3(RN_THREADS), 0, s, 1, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)res,          \
                                                       (const __nv_bfloat16*)w, rstd, (const __nv_bfloat16*)dres,    \
                                                       (__nv_bfloat16*)dx, dw_partial, rows, H)
    if (H <= 2048) RN_BWD(1);
    else if (H <= 4096) RN_BWD(2);
    else RN_BWD(4);
#undef RN_BWD
    launch_pdl(colsum_kernel, dim3((H + 31) / 32), dim3(dim3(32, 8)), 0, s, 1, dw_partial, dw_f32, (__nv_bfloat16*)dw_bf16, grid, H, accumulate);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// RoPE, in place, on x viewed as [T, nheads_total, D] where only heads with (head % group) < rot_per_group rotate
// (InternLM2 packed wqkv "(h gs d)": group = q_per_kv + 2, rot_per_group = q_per_kv + 1 -> q and k rotate, v not).
// Non-interleaved (GPT-NeoX halves) or interleaved pairs.  cos/sin tables are [max_pos, D/2] fp32.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rope_kernel(__nv_bfloat16* __restrict__ x, const int* __restrict__ pos,
                                                   const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                   int T, int heads, int D, int64_t stride_t, int group,
                                                   int rot_per_group, int rot_heads, float sign, int interleaved) {
    griddep_launch_dependents();
    griddep_wait();
    // one thread per 8 rotation pairs (16 elements: 2 x 16 B of x, 2 x 32 B of the tables); only rotating heads are
    // enumerated, so v never costs an instruction: item -> (token, rotating head index, chunk)
    const int chunks = D / 16;
    const int64_t items = (int64_t)T * rot_heads * chunks;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int ch = it % chunks;
        const int64_t th = it / chunks;
        const int rh = th % rot_heads;
        const int t = th / rot_heads;
        const int h = (rh / rot_per_group) * group + rh % rot_per_group;
        const int p = pos ? pos[t] : t;
        __nv_bfloat16* xp = x + (int64_t)t * stride_t + (int64_t)h * D;
        const float* c = cos_t + (int64_t)p * (D / 2) + ch * 8;
        const float* s = sin_t + (int64_t)p * (D / 2) + ch * 8;
        const float4 c0 = *reinterpret_cast<const float4*>(c), c1 = *reinterpret_cast<const float4*>(c + 4);
        float4 s0 = *reinterpret_cast<const float4*>(s), s1 = *reinterpret_cast<const float4*>(s + 4);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float ss[8] = {s0.x * sign, s0.y * sign, s0.z * sign, s0.w * sign,
                             s1.x * sign, s1.y * sign, s1.z * sign, s1.w * sign};
        float a[8], b[8];
        if (!interleaved) {
            __nv_bfloat16* pa = xp + ch * 8;
            __nv_bfloat16* pb = xp + D / 2 + ch * 8;
            unpack8(*reinterpret_cast<const uint4*>(pa), a);
            unpack8(*reinterpret_cast<const uint4*>(pb), b);
            float oa[8], ob[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                oa[j] = a[j] * cc[j] - b[j] * ss[j];
                ob[j] = a[j] * ss[j] + b[j] * cc[j];
            }
            *reinterpret_cast<uint4*>(pa) = pack8(oa);
            *reinterpret_cast<uint4*>(pb) = pack8(ob);
        } else {
            __nv_bfloat16* pa = xp + ch * 16;
            unpack8(*reinterpret_cast<const uint4*>(pa), a);       // pairs 0..3
            unpack8(*reinterpret_cast<const uint4*>(pa + 8), b);   // pairs 4..7
            float oa[8], ob[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                oa[2 * j] = a[2 * j] * cc[j] - a[2 * j + 1] * ss[j];
                oa[2 * j + 1] = a[2 * j] * ss[j] + a[2 * j + 1] * cc[j];
                ob[2 * j] = b[2 * j] * cc[4 + j] - b[2 * j + 1] * ss[4 + j];
                ob[2 * j + 1] = b[2 * j] * ss[4 + j] + b[2 * j + 1] * cc[4 + j];
            }
            *reinterpret_cast<uint4*>(pa) = pack8(oa);
            *reinterpret_cast<uint4*>(pa + 8) = pack8(ob);
        }
    }
}

int rope_inplace(void* x, const int* pos, const float* cos_t, const float* sin_t, int T, int heads, int D,
                 int64_t stride_t, int group, int rot_per_group, int conj, int interleaved, cudaStream_t s) {
    if (D % 16 != 0 || heads % group != 0) return -1;
    const int rot_heads = heads / group * rot_per_group;
    const int64_t items = (int64_t)T * rot_heads * (D / 16);
    const int threads = 256;
    int64_t blocks = (items + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks == 0) return 0;
    launch_pdl(rope_kernel, dim3((unsigned)blocks), dim3(threads), 0, s, 1, (__nv_bfloat16*)x, pos, cos_t, sin_t, T, heads, D, stride_t, group,
                                                     rot_per_group, rot_heads, conj ? -1.f : 1.f, interleaved);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// SwiGLU on interleaved (gate, up) columns: gu [rows, 2F] -> h [rows, F];  bwd: dgu from dh, gu
// ----------------------------------------------------------------------------------------------------------------
__global__ void swiglu_fwd_kernel(const uint4* __restrict__ gu, uint2* __restrict__ h, int64_t nvec) {
    griddep_launch_dependents();
    griddep_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float f[8];
        unpack8(gu[i], f);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = f[2 * j] / (1.f + __expf(-f[2 * j])) * f[2 * j + 1];
        uint2 r;
        r.x = pack_bf16(o[0], o[1]); r.y = pack_bf16(o[2], o[3]);
        h[i] = r;
    }
}
__global__ void swiglu_bwd_kernel(const uint2* __restrict__ dh, const uint4* __restrict__ gu, uint4* __restrict__ dgu,
                                  int64_t nvec) {
    griddep_launch_dependents();
    griddep_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float f[8], o[8];
        unpack8(gu[i], f);
        uint2 d = dh[i];
        float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y);
        const float dv[4] = {d0.x, d0.y, d1.x, d1.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = f[2 * j], u = f[2 * j + 1];
            const float sg = 1.f / (1.f + __expf(-g));
            o[2 * j] = dv[j] * u * sg * (1.f + g * (1.f - sg));
            o[2 * j + 1] = dv[j] * g * sg;
        }
        dgu[i] = pack8(o);
    }
}
// dpre = dh * d/dx gelu_tanh(pre)
__global__ void gelu_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ pre, uint4* __restrict__ dpre,
                                int64_t nvec) {
    griddep_launch_dependents();
    griddep_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float x[8], d[8], o[8];
        unpack8(pre[i], x);
        unpack8(dh[i], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float u = 0.7978845608028654f * (x[j] + 0.044715f * x[j] * x[j] * x[j]);
            const float t = tanhf(u);
            const float du = 0.7978845608028654f * (1.f + 3.f * 0.044715f * x[j] * x[j]);
            o[j] = d[j] * (0.5f * (1.f + t) + 0.5f * x[j] * (1.f - t * t) * du);
        }
        dpre[i] = pack8(o);
    }
}
int gelu_bwd(const void* dh, const void* pre, void* dpre, int64_t n, cudaStream_t s) {
    if (n % 8 != 0) return -1;
    const int64_t nvec = n / 8;
    const int blocks = (int)((nvec + 255) / 256 < 148 * 16 ? (nvec + 255) / 256 : 148 * 16);
    if (blocks == 0) return 0;
    launch_pdl(gelu_bwd_kernel, dim3(blocks), dim3(256), 0, s, 1, (const uint4*)dh, (const uint4*)pre, (uint4*)dpre, nvec);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int swiglu_fwd(const void* gu, void* h, int64_t rows, int64_t F, cudaStream_t s) {
    if (F % 4 != 0) return -1;
    const int64_t nvec = rows * F / 4;
    const int blocks = (int)((nvec + 255) / 256 < 148 * 16 ? (nvec + 255) / 256 : 148 * 16);
    launch_pdl(swiglu_fwd_kernel, dim3(blocks), dim3(256), 0, s, 1, (const uint4*)gu, (uint2*)h, nvec);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int swiglu_bwd(const void* dh, const void* gu, void* dgu, int64_t rows, int64_t F, cudaStream_t s) {
    if (F % 4 != 0) return -1;
    const int64_t nvec = rows * F / 4;
    const int blocks = (int)((nvec + 255) / 256 < 148 * 16 ? (nvec + 255) / 256 : 148 * 16);
    launch_pdl(swiglu_bwd_kernel, dim3(blocks), dim3(256), 0, s, 1, (const uint2*)dh, (const uint4*)gu, (uint4*)dgu, nvec);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// Cross entropy over (a vocab shard of) bf16 logits [rows, V], row stride ld.
//   fwd: per row online (max, sum exp(x - max), sum x, logit[target]) in ONE pass
//   bwd: logits <- gscale[row] * (exp(x - lse) * (1 - 0) - (1-eps)*onehot - eps/Vtotal)   in place
// ----------------------------------------------------------------------------------------------------------------
static constexpr int CE_THREADS = 512;

__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld,
                                                            const int64_t* __restrict__ labels, int V, int vocab_start,
                                                            float* __restrict__ out_max, float* __restrict__ out_sum,
                                                            float* __restrict__ out_sumx, float* __restrict__ out_tgt) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red[32];
    const int row = blockIdx.x;
    const __nv_bfloat16* x = logits + (int64_t)row * ld;
    const int nvec = V / 8;
    float m = -INFINITY, s = 0.f, sx = 0.f;
    for (int i = threadIdx.x; i < nvec; i += CE_THREADS) {
        float f[8];
        unpack8(ld_nc_v4(x + i * 8), f);
        float lm = f[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) lm = fmaxf(lm, f[j]);
        const float nm = fmaxf(m, lm);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc += __expf(f[j] - nm); sx += f[j]; }
        s = s * __expf(m - nm) + acc;
        m = nm;
    }
    for (int i = nvec * 8 + threadIdx.x; i < V; i += CE_THREADS) {  // tail
        const float f = __bfloat162float(x[i]);
        const float nm = fmaxf(m, f);
        s = s * __expf(m - nm) + __expf(f - nm);
        m = nm;
        sx += f;
    }
    const float gm = block_max<CE_THREADS>(m, red);
    s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
    s = block_sum<CE_THREADS>(s, red);
    sx = block_sum<CE_THREADS>(sx, red);
    if (threadIdx.x == 0) {
        out_max[row] = gm;
        out_sum[row] = s;
        out_sumx[row] = sx;
        const int64_t t = labels[row] - vocab_start;
        out_tgt[row] = (t >= 0 && t < V) ? __bfloat162float(x[t]) : 0.f;
    }
}

__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(__nv_bfloat16* __restrict__ logits, int64_t ld,
                                                            const int64_t* __restrict__ labels,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ gscale, int V, int vocab_start,
                                                            float smoothing, int total_classes, int ignore_index) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x;
    __nv_bfloat16* x = logits + (int64_t)row * ld;
    const int64_t label = labels[row];
    const float g = (label == ignore_index) ? 0.f : gscale[row];
    const float l = lse[row];
    const int64_t t = label - vocab_start;
    const float sm = smoothing / total_classes;
    const int nvec = V / 8;
    for (int i = threadIdx.x; i < nvec; i += CE_THREADS) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + i * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float p = __expf(f[j] - l) - sm;
            if (i * 8 + j == t) p -= (1.f - smoothing);
            f[j] = g * p;
        }
        *reinterpret_cast<uint4*>(x + i * 8) = pack8(f);
    }
    for (int i = nvec * 8 + threadIdx.x; i < V; i += CE_THREADS) {
        float p = __expf(__bfloat162float(x[i]) - l) - sm;
        if (i == t) p -= (1.f - smoothing);
        x[i] = __float2bfloat16_rn(g * p);
    }
}

int ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int rows, int V, int vocab_start, float* out_max,
           float* out_sum, float* out_sumx, float* out_tgt, cudaStream_t s) {
    if (ld % 8 != 0) return -1;
    launch_pdl(ce_fwd_kernel, dim3(rows), dim3(CE_THREADS), 0, s, 1, (const __nv_bfloat16*)logits, ld, labels, V, vocab_start, out_max, out_sum,
                                              out_sumx, out_tgt);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int ce_bwd(void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* gscale, int rows, int V,
           int vocab_start, float smoothing, int total_classes, int ignore_index, cudaStream_t s) {
    if (ld % 8 != 0) return -1;
    launch_pdl(ce_bwd_kernel, dim3(rows), dim3(CE_THREADS), 0, s, 1, (__nv_bfloat16*)logits, ld, labels, lse, gscale, V, vocab_start, smoothing,
                                              total_classes, ignore_index);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// Fused AdamW on a flat shard: g (bf16 or fp32) -> unscale/clip -> m, v, p (fp32) -> bf16 copy of p.
// `scalars` lives on the device so the step needs no host sync: [0] = combined grad multiplier (1/(loss_scale*clip)),
// [1] = skip flag (non-zero => overflow, leave everything untouched).
// ----------------------------------------------------------------------------------------------------------------
template <typename G>
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                             const G* __restrict__ g, __nv_bfloat16* __restrict__ p_lp, int64_t n, float lr, float beta1,
                             float beta2, float eps, float wd, float bc1, float bc2, const float* __restrict__ scalars) {
    griddep_launch_dependents();
    griddep_wait();
    const float mult = scalars ? scalars[0] : 1.f;
    if (scalars && scalars[1] != 0.f) return;
    const int64_t n4 = n / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float gv[4];
        if constexpr (sizeof(G) == 2) {
            uint2 u = reinterpret_cast<const uint2*>(g)[i];
            float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
            gv[0] = a.x; gv[1] = a.y; gv[2] = b.x; gv[3] = b.y;
        } else {
            float4 u = reinterpret_cast<const float4*>(g)[i];
            gv[0] = u.x; gv[1] = u.y; gv[2] = u.z; gv[3] = u.w;
        }
        float* pp = reinterpret_cast<float*>(&pv);
        float* mp = reinterpret_cast<float*>(&mv);
        float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gg = gv[j] * mult;
            mp[j] = beta1 * mp[j] + (1.f - beta1) * gg;
            vp[j] = beta2 * vp[j] + (1.f - beta2) * gg * gg;
            const float mh = mp[j] / bc1;
            const float vh = vp[j] / bc2;
            pp[j] = pp[j] * (1.f - lr * wd) - lr * mh / (sqrtf(vh) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (p_lp) {
            uint2 o;
            o.x = pack_bf16(pp[0], pp[1]); o.y = pack_bf16(pp[2], pp[3]);
            reinterpret_cast<uint2*>(p_lp)[i] = o;
        }
    }
    // tail (n not a multiple of 4)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = n4 * 4 + threadIdx.x;
        float gg;
        if constexpr (sizeof(G) == 2) gg = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[i]) * mult;
        else gg = reinterpret_cast<const float*>(g)[i] * mult;
        const float mm = beta1 * m[i] + (1.f - beta1) * gg;
        const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
        m[i] = mm; v[i] = vv;
        const float np = p[i] * (1.f - lr * wd) - lr * (mm / bc1) / (sqrtf(vv / bc2) + eps);
        p[i] = np;
        if (p_lp) p_lp[i] = __float2bfloat16_rn(np);
    }
}

int adamw_step(float* p, float* m, float* v, const void* g, int g_is_bf16, void* p_lp, int64_t n, float lr, float beta1,
               float beta2, float eps, float wd, float bc1, float bc2, const float* scalars, cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t want = (n / 4 + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
    if (g_is_bf16)
        launch_pdl(adamw_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, s, 1, p, m, v, (const __nv_bfloat16*)g, (__nv_bfloat16*)p_lp, n, lr,
                                                           beta1, beta2, eps, wd, bc1, bc2, scalars);
    else
        launch_pdl(adamw_kernel<float>, dim3(blocks), dim3(256), 0, s, 1, p, m, v, (const float*)g, (__nv_bfloat16*)p_lp, n, lr, beta1, beta2,
                                                   eps, wd, bc1, bc2, scalars);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// sum of squares of a flat buffer (bf16 or fp32) accumulated into *out (atomicAdd, fp32). inf/nan propagate.
// ----------------------------------------------------------------------------------------------------------------
template <typename G>
__global__ void sumsq_kernel(const G* __restrict__ g, int64_t n, float* __restrict__ out) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ float red[32];
    float acc = 0.f;
    constexpr int VEC = 16 / sizeof(G);
    const int64_t nv = n / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(g) + i);
        if constexpr (sizeof(G) == 2) {
            float f[8];
            unpack8(u, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += f[j] * f[j];
        } else {
            const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += f[j] * f[j];
        }
    }
    if (blockIdx.x == 0) {
        for (int64_t i = nv * VEC + threadIdx.x; i < n; i += blockDim.x) {
            float f;
            if constexpr (sizeof(G) == 2) f = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[i]);
            else f = reinterpret_cast<const float*>(g)[i];
            acc += f * f;
        }
    }
    acc = block_sum<256>(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
int sumsq(const void* g, int is_bf16, int64_t n, float* out, cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t want = (n / 8 + 255) / 256;
    const int blocks = (int)(want < 148 * 4 ? (want > 0 ? want : 1) : 148 * 4);
    if (is_bf16) launch_pdl(sumsq_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, s, 1, (const __nv_bfloat16*)g, n, out);
    else launch_pdl(sumsq_kernel<float>, dim3(blocks), dim3(256), 0, s, 1, (const float*)g, n, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// scalars[0] = 1 / (loss_scale * max(1, norm/clip)) ; scalars[1] = overflow flag; scalars[2] = norm (unscaled)
__global__ void clip_scalars_kernel(const float* __restrict__ sumsq_in, float* __restrict__ scalars, float loss_scale,
                                    float clip) {
    griddep_launch_dependents();
    griddep_wait();
    const float ss = *sumsq_in;
    const bool bad = !(ss == ss) || isinf(ss);
    const float norm = sqrtf(ss) / loss_scale;
    float mult = 1.f / loss_scale;
    if (clip > 0.f) {
        const float c = norm / clip;
        if (c > 1.f) mult /= c;
    }
    scalars[0] = bad ? 0.f : mult;
    scalars[1] = bad ? 1.f : 0.f;
    scalars[2] = bad ? -1.f : norm;
}
int clip_scalars(const float* sumsq_in, float* scalars, float loss_scale, float clip, cudaStream_t s) {
    launch_pdl(clip_scalars_kernel, dim3(1), dim3(1), 0, s, 1, sumsq_in, scalars, loss_scale, clip);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// Decode attention (one new token per sequence against the KV cache), split over the keys: bandwidth bound, CUDA cores.
//   q [B, H, D] bf16, kcache / vcache [B, Smax, Hkv, D] bf16 (first `seqlen` positions valid), out [B, H, D] bf16.
//   grid (nsplit, Hkv, B); one warp per q head of the kv group; 8 keys in flight per warp (4 lanes x 32 dims per key).
//   Partials (fp32 acc, m, l) go to `work`; the last-arriving split of each (b, head group) merges them (ticket counter).
// Replaces the split-KV flash-decoding kernel / FasterTransformer decode MHA the reference links (flash_fwd_splitkv,
// ft_attention).
// ----------------------------------------------------------------------------------------------------------------
static constexpr int DEC_D = 128;

__global__ void __launch_bounds__(256) attn_decode_kernel(const __nv_bfloat16* __restrict__ q,
                                                          const __nv_bfloat16* __restrict__ kc,
                                                          const __nv_bfloat16* __restrict__ vc,
                                                          __nv_bfloat16* __restrict__ out, float* __restrict__ work,
                                                          unsigned int* __restrict__ tickets, int H, int Hkv, int seqlen_host,
                                                          const int* __restrict__ seqlen_dev, int64_t stride_b,
                                                          int64_t stride_s, float scale_log2) {
    griddep_launch_dependents();
    griddep_wait();
    // CUDA-graph decode: the number of valid cache positions lives on the device (`*seqlen_dev + seqlen_host`), so one
    // captured launch serves every step; splits beyond the current length simply contribute empty partials
    const int seqlen = seqlen_dev != nullptr ? *seqlen_dev + seqlen_host : seqlen_host;
    const int split = blockIdx.x, nsplit = gridDim.x, hk = blockIdx.y, b = blockIdx.z;
    const int qpk = H / Hkv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int per = (seqlen + nsplit - 1) / nsplit;
    const int k0 = split * per, k1 = min(seqlen, k0 + per);
    __shared__ int s_last;
    for (int j = warp; j < qpk; j += blockDim.x >> 5) {
        const int h = hk * qpk + j;
        float qf[32];
        {
            const uint4* qp = reinterpret_cast<const uint4*>(q + ((int64_t)b * H + h) * DEC_D + t * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f[8];
                unpack8(qp[i], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) qf[i * 8 + e] = f[e] * scale_log2;
            }
        }
        float m = -INFINITY, l = 0.f, acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        for (int base = k0; base < k1; base += 8) {  // uniform trip count: the shuffles below need the whole warp
            const int key = base + g;
            const bool valid = key < k1;
            const int64_t off = (int64_t)b * stride_b + (int64_t)(valid ? key : k0) * stride_s + (int64_t)hk * DEC_D + t * 32;
            const uint4* kp = reinterpret_cast<const uint4*>(kc + off);
            const uint4* vp = reinterpret_cast<const uint4*>(vc + off);
            uint4 kr[4], vr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { kr[i] = kp[i]; vr[i] = vp[i]; }
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f[8];
                unpack8(kr[i], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) s = fmaf(qf[i * 8 + e], f[e], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (valid) {
                const float mn = fmaxf(m, s);
                const float corr = exp2f(m - mn), p = exp2f(s - mn);
                l = l * corr + p;
                m = mn;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float f[8];
                    unpack8(vr[i], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[i * 8 + e] = fmaf(acc[i * 8 + e], corr, p * f[e]);
                }
            }
        }
        // merge the 8 key groups of the warp (lanes with equal t hold the same 32 dims)
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
            const float mn = fmaxf(m, m2);
            const float c1 = (m == -INFINITY) ? 0.f : exp2f(m - mn), c2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - mn);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = acc[i] * c1 + __shfl_xor_sync(0xffffffffu, acc[i], o) * c2;
            l = l * c1 + l2 * c2;
            m = mn;
        }
        if (g == 0) {
            float* w = work + (((int64_t)b * H + h) * nsplit + split) * (DEC_D + 4);
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(w + t * 32 + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            if (t == 0) { w[DEC_D] = m; w[DEC_D + 1] = l; }
        }
    }
    // the last split to finish for this (b, kv group) combines all partials
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&tickets[b * Hkv + hk], 1u);
        s_last = (prev == (unsigned)nsplit - 1);
        if (s_last) tickets[b * Hkv + hk] = 0;  // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < qpk * DEC_D; idx += blockDim.x) {
        const int j = idx / DEC_D, d = idx % DEC_D;
        const int h = hk * qpk + j;
        const float* w = work + ((int64_t)b * H + h) * nsplit * (DEC_D + 4);
        float M = -INFINITY;
        for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(w + s * (DEC_D + 4) + DEC_D));
        float num = 0.f, den = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float ms = __ldcg(w + s * (DEC_D + 4) + DEC_D);
            const float c = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
            num += c * __ldcg(w + s * (DEC_D + 4) + d);
            den += c * __ldcg(w + s * (DEC_D + 4) + DEC_D + 1);
        }
        out[((int64_t)b * H + h) * DEC_D + d] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    }
}

int attn_decode(const void* q, const void* kc, const void* vc, void* out, float* work, unsigned int* tickets, int B, int H,
                int Hkv, int D, int seqlen, const int* seqlen_dev, int nsplit, int64_t stride_b, int64_t stride_s, float scale,
                cudaStream_t s) {
    if (D != DEC_D || H % Hkv != 0 || (seqlen_dev == nullptr && seqlen <= 0)) return -1;
    const int qpk = H / Hkv;
    const int warps = qpk < 8 ? qpk : 8;
    dim3 grid(nsplit, Hkv, B);
    launch_pdl(attn_decode_kernel, dim3(grid), dim3(warps * 32), 0, s, 1, (const __nv_bfloat16*)q, (const __nv_bfloat16*)kc,
                                                  (const __nv_bfloat16*)vc, (__nv_bfloat16*)out, work, tickets, H, Hkv,
                                                  seqlen, seqlen_dev, stride_b, stride_s, scale * 1.4426950408889634f);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace b200
