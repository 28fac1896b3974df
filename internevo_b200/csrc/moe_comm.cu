// MoE expert-parallel dispatch / combine over NVLink peer memory (no NCCL all-to-all, no permute buffers).
//
// The reference moves tokens with  permute -> all_to_all(_single) -> un-permute  (internlm/moe/sharded_moe.py:369-498,
// internlm/moe/megablock/megablock_moe.py:155-247): three full passes over the routed activations plus the NCCL
// staging.  Here every routed (token, k) slot knows its final address on the expert's GPU (rank, row) from a [world, E]
// count matrix, so
//
//  * symm_allgather_small   all ranks publish their per-expert counts into every peer and rendezvous - one launch
//  * moe_scatter_rows       dispatch: each warp stores one token row STRAIGHT into the owner GPU's expert slab
//                           (permute + all-to-all fused; 512-byte coalesced NVLink writes).  With `scale` it is the
//                           backward of combine (rows w[t,j] * dOut[t]); with `y_ptrs` it also pulls the expert output row
//                           back and produces d(gate weight) = <dOut[t], y_row> in the same pass.
//  * moe_gather_combine     combine: each warp PULLS the k expert-output rows of one token from their owner GPUs,
//                           applies the gate weights in fp32 and writes the combined row (all-to-all + un-permute +
//                           weighted sum fused).  With w == nullptr it is the backward of dispatch.
//
// A slot with row < 0 is dropped (capacity overflow in the GShard variants): it sends nothing and contributes zero.
#include "comm_kernels.h"
#include "sm100_ptx.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) symm_allgather_small_kernel(uint32_t* const* buf_ptrs, const uint32_t* __restrict__ src,
                                                                   int nwords, uint32_t* const* flags_ptrs, int rank,
                                                                   int world, uint32_t epoch) {
    for (int idx = threadIdx.x; idx < world * nwords; idx += blockDim.x) {
        const int p = idx / nwords, i = idx - p * nwords;
        buf_ptrs[p][(int64_t)rank * nwords + i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int p = threadIdx.x;
        st_release_sys(flags_ptrs[p] + rank, epoch);
        const uint32_t* mine = flags_ptrs[rank] + p;
        while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
        }
    }
}

int symm_allgather_small(uint32_t* const* buf_ptrs, const uint32_t* src, int nwords, uint32_t* const* flags_ptrs, int rank,
                         int world, uint32_t epoch, cudaStream_t s) {
    if (world > 256 || nwords <= 0) return -1;
    symm_allgather_small_kernel<<<1, 256, 0, s>>>(buf_ptrs, src, nwords, flags_ptrs, rank, world, epoch);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

B200_DEVICE uint4 ld_sys_v4(const void* p) {  // peer memory: never through the non-coherent path
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
B200_DEVICE void st_v4(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// one warp per (token, j) slot
__global__ void __launch_bounds__(256) moe_scatter_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                               const int* __restrict__ dst_rank,
                                                               const int* __restrict__ dst_row,
                                                               const float* __restrict__ scale, void* const* x_ptrs,
                                                               void* const* y_ptrs, float* __restrict__ dw, int n_slots,
                                                               int k, int H) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nvec = H >> 3;
    for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < n_slots; s += gridDim.x * warps_per_block) {
        const int row = dst_row[s];
        if (row < 0) {
            if (dw != nullptr && lane == 0) dw[s] = 0.f;
            continue;
        }
        const int r = dst_rank[s];
        const int t = s / k;
        const uint4* src = reinterpret_cast<const uint4*>(x + (int64_t)t * ldx);
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(x_ptrs[r]) + (int64_t)row * H);
        const uint4* yrow = y_ptrs == nullptr
                                ? nullptr
                                : reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(y_ptrs[r]) + (int64_t)row * H);
        const float sc = scale == nullptr ? 1.f : scale[s];
        float dot = 0.f;
        for (int v0 = lane; v0 < nvec; v0 += 128) {
            uint4 a[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = v0 + u * 32;
                if (v < nvec) {
                    a[u] = ld_nc_v4(src + v);
                    if (yrow != nullptr) y[u] = ld_sys_v4(yrow + v);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = v0 + u * 32;
                if (v >= nvec) continue;
                if (yrow != nullptr || scale != nullptr) {
                    const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a[u]);
                    const uint32_t* yw = reinterpret_cast<const uint32_t*>(&y[u]);
                    uint32_t ow[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 av = unpack_bf16(aw[q]);
                        if (yrow != nullptr) {
                            const float2 yv = unpack_bf16(yw[q]);
                            dot += av.x * yv.x + av.y * yv.y;
                        }
                        ow[q] = pack_bf16(av.x * sc, av.y * sc);
                    }
                    a[u] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
                st_v4(dst + v, a[u]);
            }
        }
        if (dw != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (lane == 0) dw[s] = dot;
        }
    }
}

int moe_scatter_rows(const MoeCommDesc& d, cudaStream_t s) {
    if (d.H % 8 != 0 || d.k <= 0 || d.n_slots < 0) return -1;
    if (d.n_slots == 0) return 0;
    const int blocks = (d.n_slots + 7) / 8;
    moe_scatter_rows_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(d.x), d.ldx, d.slot_rank, d.slot_row, d.scale, d.x_ptrs, d.y_ptrs, d.dw,
        d.n_slots, d.k, d.H);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// one warp per token; K = slots per token (compile-time unrolled up to 8, generic loop above)
template <int K>
__global__ void __launch_bounds__(256) moe_gather_combine_kernel(__nv_bfloat16* __restrict__ out, int64_t ldo,
                                                                 const float* __restrict__ w,
                                                                 const int* __restrict__ src_rank,
                                                                 const int* __restrict__ src_row, void* const* y_ptrs,
                                                                 int n_tokens, int k_rt, int H) {
    const int k = K > 0 ? K : k_rt;
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nvec = H >> 3;
    for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < n_tokens; t += gridDim.x * warps_per_block) {
        for (int v0 = lane; v0 < nvec; v0 += 64) {
            float acc[2][8];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[u][q] = 0.f;
#pragma unroll
            for (int j = 0; j < (K > 0 ? K : 1); ++j) {
                for (int jj = j; jj < k; jj += (K > 0 ? k : 1)) {  // K > 0: exactly one pass per unrolled j
                    const int slot = t * k + jj;
                    const int row = src_row[slot];
                    if (row < 0) continue;
                    const float wt = w == nullptr ? 1.f : w[slot];
                    const uint4* y = reinterpret_cast<const uint4*>(
                        reinterpret_cast<const __nv_bfloat16*>(y_ptrs[src_rank[slot]]) + (int64_t)row * H);
                    uint4 yv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (v0 + u * 32 < nvec) yv[u] = ld_sys_v4(y + v0 + u * 32);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (v0 + u * 32 >= nvec) continue;
                        const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yv[u]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 f = unpack_bf16(yw[q]);
                            acc[u][2 * q] += wt * f.x;
                            acc[u][2 * q + 1] += wt * f.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = v0 + u * 32;
                if (v >= nvec) continue;
                uint4 o = make_uint4(pack_bf16(acc[u][0], acc[u][1]), pack_bf16(acc[u][2], acc[u][3]),
                                     pack_bf16(acc[u][4], acc[u][5]), pack_bf16(acc[u][6], acc[u][7]));
                st_v4(reinterpret_cast<uint4*>(out + (int64_t)t * ldo) + v, o);
            }
        }
    }
}

int moe_gather_combine(const MoeCommDesc& d, cudaStream_t s) {
    if (d.H % 8 != 0 || d.k <= 0 || d.n_slots % d.k != 0) return -1;
    const int n_tokens = d.n_slots / d.k;
    if (n_tokens == 0) return 0;
    int blocks = (n_tokens + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
#define B200_LAUNCH_COMBINE(KK)                                                                                         \
    moe_gather_combine_kernel<KK><<<blocks, 256, 0, s>>>(out, d.ldo, d.scale, d.slot_rank, d.slot_row, d.y_ptrs, n_tokens, \
                                                         d.k, d.H)
    switch (d.k) {
        case 1: B200_LAUNCH_COMBINE(1); break;
        case 2: B200_LAUNCH_COMBINE(2); break;
        case 4: B200_LAUNCH_COMBINE(4); break;
        default: B200_LAUNCH_COMBINE(0); break;
    }
#undef B200_LAUNCH_COMBINE
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace b200
