// Flash attention for sm_100a (head_dim 128, bf16, causal, variable-length packed sequences, GQA).
//
// Forward (attn_fwd_kernel): one CTA = one 256-row query block (two 128-row tiles A/B) of one (sequence, q-head).
//   warp 0      TMA producer: Q once, then K_j / V_j tiles through two 2-deep smem rings
//   warp 1      MMA issuer:   S_X = Q_X K_j^T  (SS, accumulator in TMEM), O_X += P_X V_j  (TS: P read from TMEM)
//   warps 4-7   softmax for tile A, warps 8-11 for tile B: tcgen05.ld S -> online softmax (lazy rescale) -> P (bf16)
//               written back over S in TMEM (tcgen05.st); O rescaled in TMEM only when the running max moved by > 2^8
//   The two tiles ping-pong: while the softmax warps of one tile run, the tensor core works on the other tile.
//   TMEM map (512 cols): [0,128) S_A/P_A, [128,256) O_A, [256,384) S_B/P_B, [384,512) O_B.
//
// Backward (attn_bwd_*): dO·O row sums, then one CTA per (kv tile, kv head-group member) looping over q tiles with five
// tcgen05 GEMMs per step (S = QK^T, dP = dO V^T, dV += P^T dO, dK += dS^T Q, dQ += dS K) — dQ accumulated in fp32
// global memory with red.add, converted by a small post-pass.
//
// Replaces flash-attn 2.x flash_fwd / flash_bwd (reference third_party/flash-attention/csrc/flash_attn/src), which are
// sm_80 mma.sync kernels and refuse sm_100 at runtime (flash_api.cpp:246-248).
#include "attention_sm100.h"

#include <cstdio>
#include <math.h>

#include "gemm_sm100.h"
#include "sm100_ptx.cuh"

namespace b200 {

static constexpr int D = 128;        // head dim
static constexpr int TM = 128;       // rows per q tile
static constexpr int TN = 128;       // kv rows per tile
static constexpr int KV_STAGES = 2;
static constexpr int FWD_THREADS = 384;
static constexpr int TILE_BYTES = TM * D * 2;  // 32 KB

struct FwdSmem {
    // offsets into the 1024-aligned dynamic smem block
    static constexpr int Q = 0;                               // 2 tiles x 32 KB
    static constexpr int K = Q + 2 * TILE_BYTES;              // KV_STAGES x 32 KB
    static constexpr int V = K + KV_STAGES * TILE_BYTES;      // KV_STAGES x 32 KB
    static constexpr int BARS = V + KV_STAGES * TILE_BYTES;   // mbarriers
    static constexpr int TOTAL = BARS + 256 + 1024;
};

struct AttnKernelArgs {
    __nv_bfloat16* o;
    float* lse;
    const int* cu_seqlens;
    int num_seqs, T, H, Hkv;
    int64_t q_stride_g, q_stride_h;  // element offsets of q head (g, j): g * q_stride_g + j * q_stride_h
    int64_t k_stride_h, v_stride_h;
    float scale_log2;  // softmax_scale * log2(e)
    float scale;
    int causal;
};

// maps blockIdx.x -> (sequence, 256-row q block); heavy (late) blocks of each sequence first for causal balance
B200_DEVICE bool find_qblock(const int* cu, int num_seqs, int idx, int block_rows, int& seq, int& blk, int& s0, int& len) {
    int acc = 0;
    for (int b = 0; b < num_seqs; ++b) {
        const int a = cu[b], e = cu[b + 1];
        const int nb = (e - a + block_rows - 1) / block_rows;
        if (idx < acc + nb) {
            seq = b; s0 = a; len = e - a;
            blk = nb - 1 - (idx - acc);
            return true;
        }
        acc += nb;
    }
    return false;
}

B200_DEVICE float fast_exp2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnKernelArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BARS);
    uint64_t* q_full = bars;                 // 1
    uint64_t* k_full = bars + 1;             // KV_STAGES
    uint64_t* k_empty = k_full + KV_STAGES;  // KV_STAGES
    uint64_t* v_full = k_empty + KV_STAGES;
    uint64_t* v_empty = v_full + KV_STAGES;
    uint64_t* s_full = v_empty + KV_STAGES;  // 2 (tile A, B)
    uint64_t* p_full = s_full + 2;           // 2
    uint64_t* o_final = p_full + 2;          // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int seq, blk, s0, len;
    if (!find_qblock(args.cu_seqlens, args.num_seqs, blockIdx.x, 2 * TM, seq, blk, s0, len)) return;
    const int h = blockIdx.y;
    const int qpk = args.H / args.Hkv;
    const int hk = h / qpk;
    const int q_row0 = blk * 2 * TM;  // local row of the block inside the sequence
    // number of kv tiles: causal -> up to the block's last row
    const int kv_limit = args.causal ? min(len, q_row0 + 2 * TM) : len;
    const int n_kv = (kv_limit + TN - 1) / TN;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_final[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- Q: two 128-row tiles, each two 64-column boxes
            const int qcol = (int)((int64_t)hk * args.q_stride_g + (int64_t)(h % qpk) * args.q_stride_h);
            mbar_expect_tx(q_full, 2 * TILE_BYTES);
            for (int t = 0; t < 2; ++t)
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(smem + FwdSmem::Q + t * TILE_BYTES + c * (TM * 128), &tmap_q, q_full, qcol + c * 64,
                                s0 + q_row0 + t * TM);
            const int kcol = (int)((int64_t)hk * args.k_stride_h), vcol = (int)((int64_t)hk * args.v_stride_h);
            int st = 0; uint32_t ph = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(smem + FwdSmem::K + st * TILE_BYTES + c * (TN * 128), &tmap_k, &k_full[st], kcol + c * 64,
                                s0 + j * TN);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(smem + FwdSmem::V + st * TILE_BYTES + c * (TN * 128), &tmap_v, &v_full[st], vcol + c * 64,
                                s0 + j * TN);
                if (++st == KV_STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = make_idesc_f16(TM, TN, 0, 0);  // A,B K-major
            const uint32_t idesc_pv = make_idesc_f16(TM, D, 0, 1);   // A from TMEM (K-major), B = V MN-major
            const uint32_t sq = smem_u32(smem + FwdSmem::Q);
            auto issue_s = [&](int tile, int st) {
                const uint32_t a0 = sq + tile * TILE_BYTES, b0 = smem_u32(smem + FwdSmem::K + st * TILE_BYTES);
                const uint32_t d = tmem + tile * 256;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k >> 2) * (TM * 128) + (k & 3) * 32;  // 64-col atom, then 32 B per k-step
                    umma_f16_ss<1>(d, make_smem_desc_sw128(a0 + off, 16, 1024), make_smem_desc_sw128(b0 + off, 16, 1024),
                                   idesc_qk, k != 0);
                }
            };
            auto issue_pv = [&](int tile, int st, bool acc) {
                const uint32_t b0 = smem_u32(smem + FwdSmem::V + st * TILE_BYTES);
                const uint32_t p = tmem + tile * 256, o = tmem + tile * 256 + 128;
#pragma unroll
                for (int k = 0; k < TN / 16; ++k) {
                    // V is [kv, d] (MN-major B): 64-col d blocks are TN*128 B apart (LBO), 16 kv rows per k-step
                    umma_f16_ts(o, p + k * 8, make_smem_desc_sw128(b0 + k * (16 * 128), TN * 128, 1024), idesc_pv,
                                acc || k != 0);
                }
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0); umma_commit<1>(&s_full[0]);
            issue_s(1, 0); umma_commit<1>(&s_full[1]);
            umma_commit<1>(&k_empty[0]);
            int st = 0; uint32_t ph = 0;
            for (int j = 0; j < n_kv; ++j) {
                const int nst = (st + 1 == KV_STAGES) ? 0 : st + 1;
                const uint32_t nph = (st + 1 == KV_STAGES) ? ph ^ 1 : ph;
                mbar_wait(&v_full[st], ph);
                for (int tile = 0; tile < 2; ++tile) {
                    mbar_wait(&p_full[tile], j & 1);
                    tc_fence_after();
                    issue_pv(tile, st, j > 0);
                    if (tile == 1) umma_commit<1>(&v_empty[st]);
                    if (j + 1 < n_kv) {
                        if (tile == 0) { mbar_wait(&k_full[nst], nph); tc_fence_after(); }
                        issue_s(tile, nst);
                        umma_commit<1>(&s_full[tile]);
                        if (tile == 1) umma_commit<1>(&k_empty[nst]);
                    } else {
                        umma_commit<1>(&o_final[tile]);
                    }
                }
                st = nst; ph = nph;
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax warpgroups =====================
        const int tile = (warp - 4) >> 2;          // 0: tile A, 1: tile B
        const int q = warp & 3;                    // TMEM lane quadrant
        const int row_in_tile = q * 32 + lane;
        const int row = q_row0 + tile * TM + row_in_tile;  // local q index in the sequence
        const uint32_t t_s = tmem + tile * 256 + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t t_o = t_s + 128;
        float m_used = -INFINITY, l = 0.f;
        const float sl2 = args.scale_log2;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[tile], j & 1);
            tc_fence_after();
            const int kv0 = j * TN;
            // the causal / length mask only matters on tiles that touch the diagonal or the sequence end
            const bool need_mask = (kv0 + TN > len) || (args.causal && kv0 + TN - 1 > q_row0 + tile * TM);
            const int kv_max = args.causal ? min(row, len - 1) : len - 1;  // last visible kv index for this row
            // ---- pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < TN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_s + c, r);
                tmem_ld_wait();
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (kv0 + c + i <= kv_max) mx = fmaxf(mx, __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                }
            }
            float m_new = fmaxf(m_used, mx);
            // ---- lazy rescale: only when some row of this warp moved its max by more than 2^8 (warp-uniform branch)
            const bool want = (j == 0) || ((m_new - m_used) * sl2 > 8.f);
            if (__any_sync(0xffffffffu, want)) {
                if (j > 0) {
                    const float alpha = (m_new == -INFINITY) ? 1.f : fast_exp2((m_used - m_new) * sl2);
                    l *= alpha;
#pragma unroll 1
                    for (int c = 0; c < D; c += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(t_o + c, r);
                        tmem_ld_wait();
                        uint32_t lo[16], hi[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            lo[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            hi[i] = __float_as_uint(__uint_as_float(r[16 + i]) * alpha);
                        }
                        tmem_st_32x32b_x16(t_o + c, lo);
                        tmem_st_32x32b_x16(t_o + c + 16, hi);
                    }
                }
                m_used = m_new;
            }
            const float mref = (m_used == -INFINITY) ? 0.f : m_used * sl2;
            // ---- pass 2: P = exp2(s * scale_log2 - m), packed bf16 written over S
            float rs = 0.f;
#pragma unroll 1
            for (int c = 0; c < TN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_s + c, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), sl2, -mref));
                    float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), sl2, -mref));
                    if (need_mask) {
                        if (kv0 + c + i > kv_max) p0 = 0.f;
                        if (kv0 + c + i + 1 > kv_max) p1 = 0.f;
                    }
                    rs += p0 + p1;
                    pk[i >> 1] = pack_bf16(p0, p1);
                }
                tmem_st_32x32b_x16(t_s + (c >> 1), pk);
            }
            l += rs;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[tile]);
        }
        // ---- epilogue: O / l -> global, LSE
        mbar_wait(&o_final[tile], 0);
        tc_fence_after();
        const bool valid = row < len;
        const float inv = (l > 0.f) ? 1.f / l : 0.f;
        const int64_t tok = (int64_t)s0 + row;
        __nv_bfloat16* op = args.o + (tok * args.H + h) * D;
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_o + c, r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 o4;
                    o4.x = pack_bf16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                    o4.y = pack_bf16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
                    o4.z = pack_bf16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
                    o4.w = pack_bf16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(op + c + i) = o4;
                }
            }
        }
        if (valid) args.lse[(int64_t)h * args.T + tok] = m_used * args.scale + __logf(l);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<1>(tmem, 512);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// host
// ----------------------------------------------------------------------------------------------------------------
static int upper_qblocks(int T, int num_seqs, int rows) { return (T + rows - 1) / rows + num_seqs; }

int attn_fwd(const AttnDesc& d, cudaStream_t stream) {
    if (d.D != D) return -10;
    if (d.T == 0) return 0;
    CUtensorMap tq, tk, tv;
    // each map spans whole token rows of the underlying (possibly packed qkv) buffer; heads are column offsets
    const uint64_t q_row = (uint64_t)d.q_stride_t, k_row = (uint64_t)d.k_stride_t, v_row = (uint64_t)d.v_stride_t;
    if (make_tmap_2d_bf16(&tq, d.q, q_row, d.T, q_row, 64, TM)) return -11;
    if (make_tmap_2d_bf16(&tk, d.k, k_row, d.T, k_row, 64, TN)) return -11;
    if (make_tmap_2d_bf16(&tv, d.v, v_row, d.T, v_row, 64, TN)) return -11;
    AttnKernelArgs a;
    a.o = (__nv_bfloat16*)d.o; a.lse = d.lse; a.cu_seqlens = d.cu_seqlens; a.num_seqs = d.num_seqs;
    a.T = d.T; a.H = d.H; a.Hkv = d.Hkv;
    a.q_stride_g = d.q_stride_g ? d.q_stride_g : d.q_stride_h * (d.H / d.Hkv);
    a.q_stride_h = d.q_stride_h; a.k_stride_h = d.k_stride_h; a.v_stride_h = d.v_stride_h;
    a.scale = d.scale; a.scale_log2 = d.scale * 1.4426950408889634f; a.causal = d.causal;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL) != cudaSuccess)
            return -12;
        attr = true;
    }
    dim3 grid(upper_qblocks(d.T, d.num_seqs, 2 * TM), d.H);
    attn_fwd_kernel<<<grid, FWD_THREADS, FwdSmem::TOTAL, stream>>>(tq, tk, tv, a);
    return cudaGetLastError() == cudaSuccess ? 0 : -13;
}

int attn_bwd(const AttnBwdDesc&, cudaStream_t) { return -100; }

}  // namespace b200
