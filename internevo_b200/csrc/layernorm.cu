// Fused dropout + residual-add + LayerNorm (forward and backward), bf16 in / out, fp32 statistics.
//
// Replaces flash-attn's `dropout_layer_norm` extension used by the reference for `norm_type="layernorm"` and the
// dropout-add-norm prologue of a block (third_party/flash-attention/csrc/layer_norm/ln_fwd_kernels.cuh,
// ln_bwd_kernels.cuh; call sites internlm/model/modeling_internlm.py:215-248).
//
//   fwd:  r = keep * x * scale + res_in          (keep: optional uint8 mask, scale = 1 / (1 - p))
//         y = (r - mean(r)) * rstd(r) * w + b    r, mean, rstd are written for the backward
//   bwd:  g = dy * w;  xh = (r - mean) * rstd
//         dr = rstd * (g - mean(g) - xh * mean(g * xh)) + dres        (gradient of BOTH x (through keep*scale) and res_in)
//         dw_partial[block] += dy * xh;  db_partial[block] += dy      (reduced over blocks by colsum2_kernel)
//
// One 256-thread CTA per row, the row lives in registers (H <= 8192), every global access is 128 bit.
#include <cuda_bf16.h>
#include <math.h>

#include "elementwise.h"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

constexpr int LN_THREADS = 256;

B200_DEVICE float ln_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// two sums at once (one barrier pair)
B200_DEVICE float2 ln_block_sum2(float a, float b, float2* red) {
    a = ln_warp_sum(a);
    b = ln_warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = make_float2(a, b);
    __syncthreads();
    float2 t = l < LN_THREADS / 32 ? red[l] : make_float2(0.f, 0.f);
    return make_float2(ln_warp_sum(t.x), ln_warp_sum(t.y));
}
B200_DEVICE void ln_unpack8(const uint4& u, float (&f)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
B200_DEVICE uint4 ln_pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

template <int MAXV>
__global__ void __launch_bounds__(LN_THREADS) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                    const __nv_bfloat16* __restrict__ res_in,
                                                                    const uint8_t* __restrict__ keep, float drop_scale,
                                                                    const __nv_bfloat16* __restrict__ w,
                                                                    const __nv_bfloat16* __restrict__ b,
                                                                    __nv_bfloat16* __restrict__ y,
                                                                    __nv_bfloat16* __restrict__ res_out,
                                                                    float* __restrict__ mean_out,
                                                                    float* __restrict__ rstd_out, int rows, int H,
                                                                    float eps) {
    __shared__ float2 red[32];
    const int nvec = H / 8;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t base = (int64_t)row * H;
        float v[MAXV][8];
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = threadIdx.x + i * LN_THREADS;
            if (idx < nvec) {
                ln_unpack8(ld_nc_v4(reinterpret_cast<const uint4*>(x + base) + idx), v[i]);
                if (keep) {
                    const uint2 m = *reinterpret_cast<const uint2*>(keep + base + idx * 8);
                    const uint8_t* mb = reinterpret_cast<const uint8_t*>(&m);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i][j] = mb[j] ? v[i][j] * drop_scale : 0.f;
                }
                if (res_in) {
                    float r[8];
                    ln_unpack8(ld_nc_v4(reinterpret_cast<const uint4*>(res_in + base) + idx), r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i][j] += r[j];
                }
                if (res_out) {
                    const uint4 o = ln_pack8(v[i]);
                    reinterpret_cast<uint4*>(res_out + base)[idx] = o;
                    ln_unpack8(o, v[i]);  // statistics of the bf16-rounded residual, as the unfused composition has
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) s1 += v[i][j];
            }
        }
        const float mean = ln_block_sum2(s1, 0.f, red).x / H;
        float s2 = 0.f;  // two-pass variance on the register copy: no cancellation
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = threadIdx.x + i * LN_THREADS;
            if (idx < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[i][j] -= mean;
                    s2 += v[i][j] * v[i][j];
                }
            }
        }
        const float rstd = rsqrtf(ln_block_sum2(s2, 0.f, red).x / H + eps);
        if (threadIdx.x == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = threadIdx.x + i * LN_THREADS;
            if (idx < nvec) {
                float wv[8], bv[8], o[8];
                ln_unpack8(reinterpret_cast<const uint4*>(w)[idx], wv);
                if (b) ln_unpack8(reinterpret_cast<const uint4*>(b)[idx], bv);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rstd * wv[j] + (b ? bv[j] : 0.f);
                reinterpret_cast<uint4*>(y + base)[idx] = ln_pack8(o);
            }
        }
    }
}

template <int MAXV>
__global__ void __launch_bounds__(LN_THREADS) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                    const __nv_bfloat16* __restrict__ res,
                                                                    const __nv_bfloat16* __restrict__ w,
                                                                    const float* __restrict__ mean_in,
                                                                    const float* __restrict__ rstd_in,
                                                                    const __nv_bfloat16* __restrict__ dres,
                                                                    __nv_bfloat16* __restrict__ dx,
                                                                    float* __restrict__ partial, int rows, int H) {
    __shared__ float2 red[32];
    const int nvec = H / 8;
    float dwacc[MAXV][8], dbacc[MAXV][8], wv[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = threadIdx.x + i * LN_THREADS;
#pragma unroll
        for (int j = 0; j < 8; ++j) dwacc[i][j] = dbacc[i][j] = 0.f;
        if (idx < nvec) ln_unpack8(reinterpret_cast<const uint4*>(w)[idx], wv[i]);
    }
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t base = (int64_t)row * H;
        const float mean = mean_in[row], rstd = rstd_in[row];
        float g[MAXV][8], xh[MAXV][8];
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = threadIdx.x + i * LN_THREADS;
            if (idx < nvec) {
                ln_unpack8(ld_nc_v4(reinterpret_cast<const uint4*>(dy + base) + idx), g[i]);
                ln_unpack8(ld_nc_v4(reinterpret_cast<const uint4*>(res + base) + idx), xh[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] = (xh[i][j] - mean) * rstd;
                    dwacc[i][j] += g[i][j] * xh[i][j];
                    dbacc[i][j] += g[i][j];
                    g[i][j] *= wv[i][j];
                    sg += g[i][j];
                    sgx += g[i][j] * xh[i][j];
                }
            }
        }
        const float2 tot = ln_block_sum2(sg, sgx, red);
        const float mg = tot.x / H, mgx = tot.y / H;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = threadIdx.x + i * LN_THREADS;
            if (idx < nvec) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - mg - xh[i][j] * mgx);
                if (dres) {
                    float r[8];
                    ln_unpack8(ld_nc_v4(reinterpret_cast<const uint4*>(dres + base) + idx), r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += r[j];
                }
                reinterpret_cast<uint4*>(dx + base)[idx] = ln_pack8(o);
            }
        }
    }
    // partial[block] = [dw (H) | db (H)]
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = threadIdx.x + i * LN_THREADS;
        if (idx < nvec) {
            float* p = partial + (int64_t)blockIdx.x * 2 * H + idx * 8;
            reinterpret_cast<float4*>(p)[0] = make_float4(dwacc[i][0], dwacc[i][1], dwacc[i][2], dwacc[i][3]);
            reinterpret_cast<float4*>(p)[1] = make_float4(dwacc[i][4], dwacc[i][5], dwacc[i][6], dwacc[i][7]);
            reinterpret_cast<float4*>(p + H)[0] = make_float4(dbacc[i][0], dbacc[i][1], dbacc[i][2], dbacc[i][3]);
            reinterpret_cast<float4*>(p + H)[1] = make_float4(dbacc[i][4], dbacc[i][5], dbacc[i][6], dbacc[i][7]);
        }
    }
}

// out[c] = sum_b partial[b, c] for c < W (W = 2H): 32 columns x 8 row groups per block
__global__ void colsum2_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblocks, int W) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < W)
        for (int b = threadIdx.y; b < nblocks; b += 8) s += partial[(int64_t)b * W + c];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < W) {
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
        out[c] = s;
    }
}

}  // namespace

int layernorm_bwd_blocks(int rows) { return rows < 148 * 2 ? rows : 148 * 2; }

int layernorm_fwd(const void* x, const void* res_in, const uint8_t* keep, float drop_scale, const void* w, const void* b,
                  void* y, void* res_out, float* mean, float* rstd, int rows, int H, float eps, cudaStream_t s) {
    if (H % 8 != 0 || H > LN_THREADS * 8 * 4 || rows <= 0) return -1;
    const int grid = rows < 148 * 6 ? rows : 148 * 6;
#define LN_FWD(MV)                                                                                                      \
    layernorm_fwd_kernel<MV><<<grid, LN_THREADS, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)res_in, keep,   \
                                                         drop_scale, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b,  \
                                                         (__nv_bfloat16*)y, (__nv_bfloat16*)res_out, mean, rstd, rows, H, eps)
    if (H <= 2048) LN_FWD(1);
    else if (H <= 4096) LN_FWD(2);
    else LN_FWD(4);
#undef LN_FWD
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int layernorm_bwd(const void* dy, const void* res, const void* w, const float* mean, const float* rstd, const void* dres,
                  void* dx, float* partial, float* dwdb, int rows, int H, cudaStream_t s) {
    if (H % 8 != 0 || H > LN_THREADS * 8 * 4 || rows <= 0) return -1;
    const int grid = layernorm_bwd_blocks(rows);
#define LN_BWD(MV)                                                                                                      \
    layernorm_bwd_kernel<MV><<<grid, LN_THREADS, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)res,           \
                                                         (const __nv_bfloat16*)w, mean, rstd, (const __nv_bfloat16*)dres, \
                                                         (__nv_bfloat16*)dx, partial, rows, H)
    if (H <= 2048) LN_BWD(1);
    else if (H <= 4096) LN_BWD(2);
    else LN_BWD(4);
#undef LN_BWD
    colsum2_kernel<<<(2 * H + 31) / 32, dim3(32, 8), 0, s>>>(partial, dwdb, grid, 2 * H);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace b200
