// Peer-memory (NVLink 5 / NVSwitch) kernels.  Every pointer table is a device array of `world` base pointers into the
// SAME symmetric allocation on each rank (torch symmetric memory does the cuMem export/import plumbing; all data
// movement and signalling below is ours).
//
//  * symm_barrier          device-side barrier over monotonically increasing flags (st.release.sys / ld.acquire.sys)
//  * rs_reduce             Hybrid-ZeRO phase 0: this rank's arena slice = mean over peers (P2P loads), cast to bf16,
//                          sum of squares of the reduced slice accumulated for the global grad norm
//  * adamw_bcast           Hybrid-ZeRO phase 1: unscale+clip+AdamW on the fp32 master slice, bf16 parameters pushed
//                          straight into EVERY peer's parameter arena (the all-gather is the store)
//  * gemm_reduce_scatter / allgather_gemm: the tcgen05 GEMM fused with its collective in the same launch
//                          (epilogue pushes / TMA copy CTAs, see gemm_sm100.cu: GemmCommArgs)
//
// Replaces: bucketed all_reduce(AVG) + flatten/unflatten + per-owner broadcast of the reference
// (internlm/solver/optimizer/hybrid_zero_optim.py:455-523,809-837) and the NCCL calls around the TP linears
// (internlm/model/utils.py:25-217).
#include "comm_kernels.h"

#include <cstdio>

#include "sm100_ptx.cuh"

namespace b200 {

// ----------------------------------------------------------------------------------------------------------------
// barrier: flags[p][r] (on rank p) = epoch written by rank r
// ----------------------------------------------------------------------------------------------------------------
__global__ void symm_barrier_kernel(uint32_t* const* flags_ptrs, int rank, int world, uint32_t epoch) {
    const int p = threadIdx.x;
    if (p < world) {
        __threadfence_system();
        st_release_sys(flags_ptrs[p] + rank, epoch);
        const uint32_t* mine = flags_ptrs[rank] + p;
        while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
        }
    }
}

int symm_barrier(uint32_t* const* flags_ptrs, int rank, int world, uint32_t epoch, cudaStream_t s) {
    symm_barrier_kernel<<<1, 32, 0, s>>>(flags_ptrs, rank, world, epoch);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// ZeRO phase 0: reduce-scatter by peer loads
// ----------------------------------------------------------------------------------------------------------------
static constexpr int MAX_WORLD = 8;

B200_DEVICE uint4 ld_peer_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
B200_DEVICE void st_peer_v2(void* p, uint2 v) {
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

template <int W>
__global__ void __launch_bounds__(256) rs_reduce_kernel(void* const* grad_ptrs, int rank, int64_t shard_off,
                                                        int64_t shard_n, float inv_div, float* scalars) {
    __shared__ float red[32];
    const __nv_bfloat16* src[W];
#pragma unroll
    for (int p = 0; p < W; ++p)  // start with the local copy, then walk the ring so link load is spread
        src[p] = reinterpret_cast<const __nv_bfloat16*>(grad_ptrs[(rank + p) % W]) + shard_off;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(grad_ptrs[rank]) + shard_off;
    const int64_t nvec = shard_n / 8;
    float ss = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v[W];
#pragma unroll
        for (int p = 0; p < W; ++p) v[p] = ld_peer_v4(src[p] + i * 8);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < W; ++p) {
            float2 a = unpack_bf16(v[p].x), b = unpack_bf16(v[p].y), c = unpack_bf16(v[p].z), d = unpack_bf16(v[p].w);
            acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
            acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
        }
        uint4 o;
        o.x = pack_bf16(acc[0] * inv_div, acc[1] * inv_div);
        o.y = pack_bf16(acc[2] * inv_div, acc[3] * inv_div);
        o.z = pack_bf16(acc[4] * inv_div, acc[5] * inv_div);
        o.w = pack_bf16(acc[6] * inv_div, acc[7] * inv_div);
        *reinterpret_cast<uint4*>(dst + i * 8) = o;
        float2 a = unpack_bf16(o.x), b = unpack_bf16(o.y), c = unpack_bf16(o.z), d = unpack_bf16(o.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    // block reduce -> one atomic per block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) atomicAdd(scalars + 3, t);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// ZeRO phase 1: AdamW on the owned slice, parameters pushed to every peer
// ----------------------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) adamw_bcast_kernel(float* __restrict__ p, float* __restrict__ m,
                                                          float* __restrict__ v, void* const* grad_ptrs,
                                                          void* const* param_ptrs, int rank, int64_t shard_off,
                                                          int64_t n, float lr, float beta1, float beta2, float eps,
                                                          float wd, float bc1, float bc2, const float* scalars) {
    const float mult = scalars[0];
    if (scalars[1] != 0.f) return;  // overflow: skip the step everywhere (flag is identical on all ranks)
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(grad_ptrs[rank]) + shard_off;
    __nv_bfloat16* dst[W];
#pragma unroll
    for (int q = 0; q < W; ++q) dst[q] = reinterpret_cast<__nv_bfloat16*>(param_ptrs[(rank + q) % W]) + shard_off;
    const int64_t n4 = n / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        const uint2 gu = reinterpret_cast<const uint2*>(g)[i];
        const float2 g0 = unpack_bf16(gu.x), g1 = unpack_bf16(gu.y);
        const float gv[4] = {g0.x, g0.y, g1.x, g1.y};
        float* pp = reinterpret_cast<float*>(&pv);
        float* mp = reinterpret_cast<float*>(&mv);
        float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gg = gv[j] * mult;
            mp[j] = beta1 * mp[j] + (1.f - beta1) * gg;
            vp[j] = beta2 * vp[j] + (1.f - beta2) * gg * gg;
            pp[j] = pp[j] * (1.f - lr * wd) - lr * (mp[j] / bc1) / (sqrtf(vp[j] / bc2) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
        uint2 o;
        o.x = pack_bf16(pp[0], pp[1]);
        o.y = pack_bf16(pp[2], pp[3]);
#pragma unroll
        for (int q = 0; q < W; ++q) st_peer_v2(dst[q] + i * 4, o);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// NVLS variants: the arenas are also mapped through an NVSwitch MULTICAST address.  `multimem.ld_reduce` makes the
// switch fetch the same 16 bytes from every GPU and add them (fp32 accumulation) on the way back: the owner reads its
// slice ONCE instead of once per peer.  `multimem.st` makes the switch replicate one store into every GPU's arena: the
// owner writes its updated parameters ONCE instead of once per peer.  Link traffic per GPU drops from (W-1)/W * arena to
// 1/W * arena in both phases.
// ----------------------------------------------------------------------------------------------------------------
B200_DEVICE uint4 multimem_ld_reduce_bf16x8(const void* mc) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
B200_DEVICE void multimem_st_v4(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

__global__ void __launch_bounds__(256) rs_reduce_mc_kernel(const void* grad_mc, void* grad_local, int64_t shard_off,
                                                           int64_t shard_n, float inv_div, float* scalars) {
    __shared__ float red[32];
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(grad_mc) + shard_off;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(grad_local) + shard_off;
    const int64_t nvec = shard_n / 8;
    float ss = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = multimem_ld_reduce_bf16x8(src + i * 8);
        float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
        uint4 o;
        o.x = pack_bf16(a.x * inv_div, a.y * inv_div);
        o.y = pack_bf16(b.x * inv_div, b.y * inv_div);
        o.z = pack_bf16(c.x * inv_div, c.y * inv_div);
        o.w = pack_bf16(d.x * inv_div, d.y * inv_div);
        *reinterpret_cast<uint4*>(dst + i * 8) = o;
        a = unpack_bf16(o.x); b = unpack_bf16(o.y); c = unpack_bf16(o.z); d = unpack_bf16(o.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) atomicAdd(scalars + 3, t);
    }
}

__global__ void __launch_bounds__(256) adamw_bcast_mc_kernel(float* __restrict__ p, float* __restrict__ m,
                                                             float* __restrict__ v, const void* grad_local,
                                                             void* param_mc, int64_t shard_off, int64_t n, float lr,
                                                             float beta1, float beta2, float eps, float wd, float bc1,
                                                             float bc2, const float* scalars) {
    const float mult = scalars[0];
    if (scalars[1] != 0.f) return;
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(grad_local) + shard_off;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(param_mc) + shard_off;
    const int64_t n8 = n / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        float pv[8], mv[8], vv[8], gv[8];
        *reinterpret_cast<float4*>(pv) = reinterpret_cast<float4*>(p)[2 * i];
        *reinterpret_cast<float4*>(pv + 4) = reinterpret_cast<float4*>(p)[2 * i + 1];
        *reinterpret_cast<float4*>(mv) = reinterpret_cast<float4*>(m)[2 * i];
        *reinterpret_cast<float4*>(mv + 4) = reinterpret_cast<float4*>(m)[2 * i + 1];
        *reinterpret_cast<float4*>(vv) = reinterpret_cast<float4*>(v)[2 * i];
        *reinterpret_cast<float4*>(vv + 4) = reinterpret_cast<float4*>(v)[2 * i + 1];
        const uint4 gu = reinterpret_cast<const uint4*>(g)[i];
        float2 g0 = unpack_bf16(gu.x), g1 = unpack_bf16(gu.y), g2 = unpack_bf16(gu.z), g3 = unpack_bf16(gu.w);
        gv[0] = g0.x; gv[1] = g0.y; gv[2] = g1.x; gv[3] = g1.y; gv[4] = g2.x; gv[5] = g2.y; gv[6] = g3.x; gv[7] = g3.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gg = gv[j] * mult;
            mv[j] = beta1 * mv[j] + (1.f - beta1) * gg;
            vv[j] = beta2 * vv[j] + (1.f - beta2) * gg * gg;
            pv[j] = pv[j] * (1.f - lr * wd) - lr * (mv[j] / bc1) / (sqrtf(vv[j] / bc2) + eps);
        }
        reinterpret_cast<float4*>(p)[2 * i] = *reinterpret_cast<float4*>(pv);
        reinterpret_cast<float4*>(p)[2 * i + 1] = *reinterpret_cast<float4*>(pv + 4);
        reinterpret_cast<float4*>(m)[2 * i] = *reinterpret_cast<float4*>(mv);
        reinterpret_cast<float4*>(m)[2 * i + 1] = *reinterpret_cast<float4*>(mv + 4);
        reinterpret_cast<float4*>(v)[2 * i] = *reinterpret_cast<float4*>(vv);
        reinterpret_cast<float4*>(v)[2 * i + 1] = *reinterpret_cast<float4*>(vv + 4);
        uint4 o;
        o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]);
        o.z = pack_bf16(pv[4], pv[5]); o.w = pack_bf16(pv[6], pv[7]);
        multimem_st_v4(dst + i * 8, o);
    }
}

template <int W>
static int launch_zero(const RsAdamDesc& d, cudaStream_t s) {
    const int64_t want = (d.shard_n / 8 + 255) / 256;
    const int blocks = (int)(want < 148 * 4 ? (want > 0 ? want : 1) : 148 * 4);
    if (d.phase == 0 && d.grad_mc != nullptr) {
        rs_reduce_mc_kernel<<<blocks, 256, 0, s>>>(d.grad_mc, d.grad_local, d.shard_off, d.shard_n, (float)(1.0 / d.grad_div),
                                                   d.scalars);
    } else if (d.phase == 1 && d.param_mc != nullptr) {
        adamw_bcast_mc_kernel<<<blocks, 256, 0, s>>>(d.p, d.m, d.v, d.grad_local, d.param_mc, d.shard_off, d.shard_n,
                                                     (float)d.lr, (float)d.beta1, (float)d.beta2, (float)d.eps, (float)d.wd,
                                                     (float)d.bc1, (float)d.bc2, d.scalars);
    } else if (d.phase == 0) {
        rs_reduce_kernel<W><<<blocks, 256, 0, s>>>(d.grad_ptrs, d.rank, d.shard_off, d.shard_n, (float)(1.0 / d.grad_div),
                                                   d.scalars);
    } else {
        adamw_bcast_kernel<W><<<blocks, 256, 0, s>>>(d.p, d.m, d.v, d.grad_ptrs, d.param_ptrs, d.rank, d.shard_off,
                                                     d.shard_n, (float)d.lr, (float)d.beta1, (float)d.beta2, (float)d.eps,
                                                     (float)d.wd, (float)d.bc1, (float)d.bc2, d.scalars);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int reduce_scatter_adam(const RsAdamDesc& d, cudaStream_t s) {
    if (d.shard_n % 8 != 0 || d.shard_off % 8 != 0) return -1;
    switch (d.world) {
        case 2: return launch_zero<2>(d, s);
        case 4: return launch_zero<4>(d, s);
        case 8: return launch_zero<8>(d, s);
        default: return -3;
    }
    static_assert(MAX_WORLD == 8, "");
}

// ----------------------------------------------------------------------------------------------------------------
// peer-copy micro-benchmark: the inner loop of the all-gather push (gemm_sm100.cu::ag_push_pieces) in isolation, with the
// number of 16-byte loads a thread keeps in flight as a template parameter.  Used by tools/peer_copy_bench.py to choose
// the unroll of the production loop; not on any training path.
// ----------------------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(128) peer_copy_kernel(const uint8_t* __restrict__ src, uint8_t* dst, int64_t bytes,
                                                        int64_t piece_bytes) {
    const int t = threadIdx.x;
    const int64_t pieces = bytes / piece_bytes;
    for (int64_t p = blockIdx.x; p < pieces; p += gridDim.x) {
        const uint8_t* s = src + p * piece_bytes;
        uint8_t* d = dst + p * piece_bytes;
        for (int64_t o = (int64_t)t * 16; o < piece_bytes; o += (int64_t)U * 2048) {
            uint4 v[U];
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (o + j * 2048 < piece_bytes) v[j] = ld_nc_v4(s + o + j * 2048);
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (o + j * 2048 < piece_bytes)
                    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(d + o + j * 2048), "r"(v[j].x), "r"(v[j].y),
                                 "r"(v[j].z), "r"(v[j].w)
                                 : "memory");
        }
    }
}

int peer_copy_bench(const void* src, void* dst, int64_t bytes, int64_t piece_bytes, int unroll, int ctas, cudaStream_t s) {
    const uint8_t* a = reinterpret_cast<const uint8_t*>(src);
    uint8_t* b = reinterpret_cast<uint8_t*>(dst);
    switch (unroll) {
        case 4: peer_copy_kernel<4><<<ctas, 128, 0, s>>>(a, b, bytes, piece_bytes); break;
        case 8: peer_copy_kernel<8><<<ctas, 128, 0, s>>>(a, b, bytes, piece_bytes); break;
        case 16: peer_copy_kernel<16><<<ctas, 128, 0, s>>>(a, b, bytes, piece_bytes); break;
        default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------------------------
// GEMM + collective: thin wrappers over the comm-aware GEMM launch (gemm_sm100.cu)
// ----------------------------------------------------------------------------------------------------------------
int gemm_reduce_scatter(const GemmCommDesc& d, cudaStream_t s) {
    GemmCommArgs c;
    c.mode = d.mode == 1 ? GEMM_COMM_ALL_REDUCE : GEMM_COMM_REDUCE_SCATTER;
    c.peer_ptrs = d.peer_ptrs; c.out_ptrs = d.out_ptrs; c.flags_ptrs = d.flags_ptrs; c.rank = d.rank; c.world = d.world;
    c.epoch = d.epoch; c.out_local = d.out_local; c.ld_out = d.ld_out; c.m_local = d.g.M / d.world;
    c.comm_ctas = d.comm_ctas; c.done_ptrs = d.done_ptrs; c.done_counter = d.done_counter; c.out_scale = d.out_scale;
    return gemm_bf16_comm(d.g, c, s);
}

int gather_weight_gemm(const GemmCommDesc& d, cudaStream_t s) {
    GemmCommArgs c;
    c.mode = GEMM_COMM_GATHER_B;
    c.peer_ptrs = d.peer_ptrs; c.flags_ptrs = d.flags_ptrs; c.rank = d.rank; c.world = d.world; c.epoch = d.epoch;
    c.out_local = d.out_local; c.m_local = d.m_local; c.x_local = d.x_local;
    return gemm_bf16_comm(d.g, c, s);
}

int allgather_gemm(const GemmCommDesc& d, cudaStream_t s) {
    GemmCommArgs c;
    c.mode = GEMM_COMM_ALL_GATHER;
    c.peer_ptrs = d.peer_ptrs; c.flags_ptrs = d.flags_ptrs; c.rank = d.rank; c.world = d.world; c.epoch = d.epoch;
    c.out_local = d.out_local; c.ld_out = d.ld_out; c.m_local = d.m_local; c.comm_ctas = d.comm_ctas;
    c.x_local = d.x_local;
    return gemm_bf16_comm(d.g, c, s);
}

}  // namespace b200
