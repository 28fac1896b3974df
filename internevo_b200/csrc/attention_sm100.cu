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
#include "launch.h"
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

// Sequence-parallel form (SP): this rank holds the query rows [rank * t_local, (rank + 1) * t_local) of the packed token
// stream (cu_seqlens stay GLOBAL) and the K / V rows of every rank are reachable through one tensor map per peer (the
// peers' symmetric buffers, mapped over NVLink): the K / V producer walks the key tiles of a sequence across rank
// boundaries inside ONE online-softmax pass - no all-to-all, no ring steps, no LSE merge, no limit on sp vs. kv heads.
// Key tiles are aligned to GLOBAL multiples of 128 rows (t_local % 128 == 0), so a tile never straddles two peers.
struct SpArgs {
    int rank, world, t_local;
};
struct PeerMaps {
    CUtensorMap k[8];
    CUtensorMap v[8];
};

// SP: blockIdx -> (sequence, 256-row q block of the part of the sequence that lies in this rank's window)
B200_DEVICE bool find_qblock_sp(const int* cu, int num_seqs, int idx, int block_rows, int w0, int w1, int& s0, int& len,
                                int& q_row0, int& q_end, bool reverse = true) {
    int acc = 0;
    for (int b = 0; b < num_seqs; ++b) {
        const int a = cu[b], e = cu[b + 1];
        const int fa = max(a, w0), fe = min(e, w1);
        if (fa >= fe) continue;
        const int nb = (fe - fa + block_rows - 1) / block_rows;
        if (idx < acc + nb) {
            s0 = a; len = e - a;
            const int blk = reverse ? nb - 1 - (idx - acc) : idx - acc;
            q_row0 = fa - a + blk * block_rows;   // position of the block's first row inside the sequence
            q_end = fe - a;                       // rows of the sequence beyond this one belong to the next rank
            return true;
        }
        acc += nb;
    }
    return false;
}

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

template <bool SP>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnKernelArgs args, const SpArgs sp,
                const __grid_constant__ PeerMaps peers) {
    griddep_launch_dependents();  // PDL (launch.h)
    griddep_wait();               // cu_seqlens is read right away: no prologue to overlap here
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
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
    int seq, blk, s0, len, q_row0, q_end;
    // grid = (heads, q blocks): CTAs are dispatched x-fastest, so the heaviest (latest) q blocks of ALL heads start first
    if constexpr (SP) {
        if (!find_qblock_sp(args.cu_seqlens, args.num_seqs, blockIdx.y, 2 * TM, sp.rank * sp.t_local,
                            (sp.rank + 1) * sp.t_local, s0, len, q_row0, q_end))
            return;
    } else {
        if (!find_qblock(args.cu_seqlens, args.num_seqs, blockIdx.y, 2 * TM, seq, blk, s0, len)) return;
        q_row0 = blk * 2 * TM;  // local row of the block inside the sequence
        q_end = len;
    }
    const int h = blockIdx.x;
    const int qpk = args.H / args.Hkv;
    const int hk = h / qpk;
    // number of kv tiles: causal -> up to the block's last row.  SP: tiles are aligned to global multiples of TN, the first
    // one may start before the sequence does (those keys are masked)
    const int kv_limit = args.causal ? min(len, q_row0 + 2 * TM) : len;
    const int jt0 = SP ? s0 / TN : 0;                                   // first global key tile (SP)
    const int n_kv = SP ? (s0 + kv_limit + TN - 1) / TN - jt0 : (kv_limit + TN - 1) / TN;
    const int q_tok0 = SP ? s0 + q_row0 - sp.rank * sp.t_local : s0 + q_row0;   // row of the block in the local q / o buffers

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
                                q_tok0 + t * TM);
            const int kcol = (int)((int64_t)hk * args.k_stride_h), vcol = (int)((int64_t)hk * args.v_stride_h);
            int st = 0; uint32_t ph = 0;
            for (int j = 0; j < n_kv; ++j) {
                const CUtensorMap *mk = &tmap_k, *mv = &tmap_v;
                int krow = s0 + j * TN;
                if constexpr (SP) {   // the tile lives on exactly one peer: its map, its local row
                    const int g = (jt0 + j) * TN, pr = g / sp.t_local;
                    mk = &peers.k[pr]; mv = &peers.v[pr];
                    krow = g - pr * sp.t_local;
                }
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(smem + FwdSmem::K + st * TILE_BYTES + c * (TN * 128), mk, &k_full[st], kcol + c * 64, krow);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], TILE_BYTES);
                for (int c = 0; c < 2; ++c)
                    tma_load_2d(smem + FwdSmem::V + st * TILE_BYTES + c * (TN * 128), mv, &v_full[st], vcol + c * 64, krow);
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
            // key positions relative to the sequence start; SP tiles are globally aligned, so the first one may begin at a
            // negative position (keys of the previous sequence / rank: masked by kv_min)
            const int kv0 = SP ? (jt0 + j) * TN - s0 : j * TN;
            // the causal / length mask only matters on tiles that touch the diagonal or the sequence ends
            const bool need_mask = (kv0 + TN > len) || (args.causal && kv0 + TN - 1 > q_row0 + tile * TM) || (SP && kv0 < 0);
            const int kv_max = args.causal ? min(row, len - 1) : len - 1;  // last visible kv index for this row
            const int kv_min = 0;
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
                        if (kv0 + c + i <= kv_max && (!SP || kv0 + c + i >= kv_min)) mx = fmaxf(mx, __uint_as_float(r[i]));
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
                        if (kv0 + c + i > kv_max || (SP && kv0 + c + i < kv_min)) p0 = 0.f;
                        if (kv0 + c + i + 1 > kv_max || (SP && kv0 + c + i + 1 < kv_min)) p1 = 0.f;
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
        const bool valid = row < q_end;
        const float inv = (l > 0.f) ? 1.f / l : 0.f;
        const int64_t tok = (int64_t)q_tok0 + (row - q_row0);   // row in the (local) output
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
        if (cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL) !=
                cudaSuccess ||
            cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL) !=
                cudaSuccess)
            return -12;
        attr = true;
    }
    dim3 grid(d.H, upper_qblocks(d.T, d.num_seqs, 2 * TM));
    if (d.sp_world > 1) {
        // sequence parallel: T is the LOCAL token count, K / V of peer p through its own map (same strides everywhere)
        if (d.sp_world > 8 || d.T % TN != 0 || d.k_peers == nullptr || d.v_peers == nullptr) return -14;
        PeerMaps pm;
        for (int p = 0; p < d.sp_world; ++p) {
            if (make_tmap_2d_bf16(&pm.k[p], d.k_peers[p], k_row, d.T, k_row, 64, TN)) return -11;
            if (make_tmap_2d_bf16(&pm.v[p], d.v_peers[p], v_row, d.T, v_row, 64, TN)) return -11;
        }
        for (int p = d.sp_world; p < 8; ++p) { pm.k[p] = pm.k[0]; pm.v[p] = pm.v[0]; }
        SpArgs spa{d.sp_rank, d.sp_world, d.T};
        launch_pdl(attn_fwd_kernel<true>, dim3(grid), dim3(FWD_THREADS), FwdSmem::TOTAL, stream, 1, tq, tk, tv, a, spa, pm);
    } else {
        PeerMaps pm;
        pm.k[0] = tk;   // unused by the single-rank instantiation
        launch_pdl(attn_fwd_kernel<false>, dim3(grid), dim3(FWD_THREADS), FwdSmem::TOTAL, stream, 1, tq, tk, tv, a, SpArgs{0, 1, d.T},
                   pm);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -13;
}


// ================================================================================================================
// backward
// ================================================================================================================
static constexpr int BQ = 64;                  // q rows per inner step
static constexpr int BWD_THREADS = 384;
static constexpr int QT_BYTES = BQ * D * 2;    // 16 KB
static constexpr int QSTAGES = 3;              // Q / dO ring depth: the slot of step i is released when dQ(i) retires,
                                               // so two stages left the TMA one MMA batch of lead -> S^T(i+1) issued late

struct BwdSmem {
    static constexpr int K = 0;                          // 32 KB
    static constexpr int V = K + TILE_BYTES;             // 32 KB
    static constexpr int Q = V + TILE_BYTES;             // 2 x 16 KB
    static constexpr int DO = Q + QSTAGES * QT_BYTES;    // QSTAGES x 16 KB
    static constexpr int DS = DO + QSTAGES * QT_BYTES;   // 2 x 16 KB  (dS^T, [128 kv rows x 64 q] bf16)
    static constexpr int DQ = DS + 2 * QT_BYTES;         // 32 KB fp32 staging [64 q x 128 d] for the TMA reduce-add
    static constexpr int LD = DQ + BQ * D * 4;           // 2 x (64 lse2 + 64 delta) floats
    static constexpr int BARS = LD + 2 * 2 * BQ * 4;
    static constexpr int TOTAL = BARS + 256 + 1024;
};

// sequence-parallel backward: the kv tile is local; the q tiles it meets (this rank's and the later ranks' rows of the same
// sequence) are read from their owners - Q / dO through per-peer tensor maps, lse / delta through per-peer pointers - and
// dQ is reduce-added into the owner's fp32 accumulator with a TMA reduce to peer memory.  q tiles are aligned to GLOBAL
// multiples of 64 rows, so a tile never straddles two peers (t_local % 64 == 0).
struct BwdPeers {
    CUtensorMap q[8];
    CUtensorMap dout[8];
    CUtensorMap dq[8];
};
struct BwdSpArgs {
    int rank, world, t_local;
    const float* lse2[8];
    const float* delta[8];
};

struct AttnBwdArgs {
    AttnKernelArgs f;
    const float* lse;
    const float* delta;
    float* dq_acc;            // [T, H, D] fp32
    __nv_bfloat16 *dk, *dv;   // strided outputs
    int64_t dk_stride_t, dk_stride_h, dv_stride_t, dv_stride_h;
};

// delta[h, t] = sum_d dO[t,h,d] * O[t,h,d];   lse2[h, t] = lse[h, t] * log2(e)   (stored behind delta: [2, H, T])
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                      const float* __restrict__ lse, float* __restrict__ delta, int T, int H) {
    griddep_launch_dependents();
    griddep_wait();
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per (t, h)
    const int lane = threadIdx.x & 31;
    if (gw >= (int64_t)T * H) return;
    const int t = gw / H, h = gw % H;
    const uint2 a = *reinterpret_cast<const uint2*>(dout + gw * D + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(out + gw * D + lane * 4);
    float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y);
    float s = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        delta[(int64_t)h * T + t] = s;
        delta[(int64_t)H * T + (int64_t)h * T + t] = lse[(int64_t)h * T + t] * 1.4426950408889634f;
    }
}

// dq (bf16, strided [T, (g, j), D]) = dq_acc (fp32 [T, H, D])
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int T, int H,
                                           int qpk, int64_t st, int64_t sg, int64_t sh, float scale) {
    griddep_launch_dependents();
    griddep_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 8 elements
    const int64_t n = (int64_t)T * H * (D / 8);
    if (i >= n) return;
    const int c = (i % (D / 8)) * 8;
    const int64_t th = i / (D / 8);
    const int h = th % H;
    const int64_t t = th / H;
    const float4 x = *reinterpret_cast<const float4*>(acc + th * D + c);
    const float4 y = *reinterpret_cast<const float4*>(acc + th * D + c + 4);
    uint4 o;
    o.x = pack_bf16(x.x * scale, x.y * scale); o.y = pack_bf16(x.z * scale, x.w * scale);
    o.z = pack_bf16(y.x * scale, y.y * scale); o.w = pack_bf16(y.z * scale, y.w * scale);
    *reinterpret_cast<uint4*>(dq + t * st + (int64_t)(h / qpk) * sg + (int64_t)(h % qpk) * sh + c) = o;
}

template <bool SP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, const AttnBwdArgs args, const BwdSpArgs sp,
                const __grid_constant__ BwdPeers peers) {
    griddep_launch_dependents();  // PDL (launch.h)
    griddep_wait();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BARS);
    uint64_t* kv_full = bars;            // 1
    uint64_t* qdo_full = bars + 17;      // QSTAGES
    uint64_t* qdo_empty = bars + 17 + QSTAGES;  // QSTAGES
    uint64_t* s_full = bars + 5;         // 2
    uint64_t* p_ready = bars + 7;        // 2 (count 4)
    uint64_t* dp_full = bars + 9;        // 1
    uint64_t* ds_ready = bars + 10;      // 1 (count 4)
    uint64_t* ds_free = bars + 11;       // 2
    uint64_t* dq_full = bars + 13;       // 1
    uint64_t* dq_free = bars + 14;       // 1 (count 4)
    uint64_t* dkv_done = bars + 15;      // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const AttnKernelArgs& f = args.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int seq, blk, s0, len, kv0, kv_end;
    // grid = (kv heads, kv tiles): x-fastest dispatch starts the heaviest kv tiles of all heads first (LPT order)
    if constexpr (SP) {
        // kv tiles of the sequence fragments inside this rank's window, aligned to the fragment start, early tiles first
        if (!find_qblock_sp(f.cu_seqlens, f.num_seqs, blockIdx.y, TN, sp.rank * sp.t_local, (sp.rank + 1) * sp.t_local, s0, len,
                            kv0, kv_end, false))
            return;
    } else {
        if (!find_qblock(f.cu_seqlens, f.num_seqs, blockIdx.y, TN, seq, blk, s0, len)) return;
        const int nkv = (len + TN - 1) / TN;
        blk = nkv - 1 - blk;                      // find_qblock reverses; bwd wants early (heavy) kv tiles first
        kv0 = blk * TN;
        kv_end = len;
    }
    const int hk = blockIdx.x;
    const int qpk = f.H / f.Hkv;
    // q tiles: non-SP seq-relative index mq (rows s0 + mq * BQ); SP GLOBAL index (rows mq * BQ of the packed stream)
    const int mq0 = SP ? (f.causal ? (s0 + kv0) / BQ : s0 / BQ) : (f.causal ? kv0 / BQ : 0);
    const int nq = SP ? (s0 + len + BQ - 1) / BQ - mq0 : (len + BQ - 1) / BQ - mq0;   // q tiles per head
    const int n_steps = nq * qpk;
    const int w0 = SP ? sp.rank * sp.t_local : 0;   // first global row held locally

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 8); mbar_init(&ds_free[i], 1);
        }
        for (int i = 0; i < QSTAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
        mbar_init(dp_full, 1); mbar_init(ds_ready, 8); mbar_init(dq_full, 1); mbar_init(dq_free, 8);
        mbar_init(dkv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns
    const uint32_t T_S0 = 0, T_DP = 128, T_DQ = 192, T_DV = 256, T_DK = 384;  // S^T buffers at 0 and 64

    if (warp == 0) {
        if (lane == 0) {
            const int kcol = (int)((int64_t)hk * f.k_stride_h), vcol = (int)((int64_t)hk * f.v_stride_h);
            mbar_expect_tx(kv_full, 2 * TILE_BYTES);
            for (int c = 0; c < 2; ++c) {
                tma_load_2d(smem + BwdSmem::K + c * (TN * 128), &tmap_k, kv_full, kcol + c * 64, s0 + kv0 - w0);
                tma_load_2d(smem + BwdSmem::V + c * (TN * 128), &tmap_v, kv_full, vcol + c * 64, s0 + kv0 - w0);
            }
            for (int i = 0; i < n_steps; ++i) {
                const int st = i % QSTAGES;
                const int j = i / nq, mq = mq0 + i % nq;
                const int h = hk * qpk + j;
                mbar_wait(&qdo_empty[st], ((i / QSTAGES) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[st], 2 * QT_BYTES);
                const int qcol = (int)((int64_t)hk * f.q_stride_g + (int64_t)j * f.q_stride_h);
                const CUtensorMap *mq_map = &tmap_q, *mdo_map = &tmap_do;
                int qrow = s0 + mq * BQ;
                if constexpr (SP) {   // the q tile lives on exactly one peer
                    const int pr = (mq * BQ) / sp.t_local;
                    mq_map = &peers.q[pr]; mdo_map = &peers.dout[pr];
                    qrow = mq * BQ - pr * sp.t_local;
                }
                for (int c = 0; c < 2; ++c) {
                    tma_load_2d(smem + BwdSmem::Q + st * QT_BYTES + c * (BQ * 128), mq_map, &qdo_full[st], qcol + c * 64, qrow);
                    tma_load_2d(smem + BwdSmem::DO + st * QT_BYTES + c * (BQ * 128), mdo_map, &qdo_full[st],
                                h * D + c * 64, qrow);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_s = make_idesc_f16(128, BQ, 0, 0);    // S^T / dP^T : A,B K-major
            const uint32_t id_kv = make_idesc_f16(128, D, 0, 1);    // dV / dK    : A TMEM, B MN-major
            const uint32_t id_dq = make_idesc_f16(128, BQ, 1, 1);   // dQ^T       : A,B MN-major
            const uint32_t sK = smem_u32(smem + BwdSmem::K), sV = smem_u32(smem + BwdSmem::V);
            auto issue_st = [&](uint32_t a_base, uint32_t b_base, uint32_t d_col) {
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t ao = (k >> 2) * (TN * 128) + (k & 3) * 32;
                    const uint32_t bo = (k >> 2) * (BQ * 128) + (k & 3) * 32;
                    umma_f16_ss<1>(tmem + d_col, make_smem_desc_sw128(a_base + ao, 16, 1024),
                                   make_smem_desc_sw128(b_base + bo, 16, 1024), id_s, k != 0);
                }
            };
            auto issue_acc = [&](uint32_t a_col, uint32_t b_base, uint32_t d_col, bool acc) {
#pragma unroll
                for (int k = 0; k < BQ / 16; ++k)
                    // packed bf16 rows live in the first 16 columns of each 32-column half (written in place by the
                    // warps that own that half, so they never clobber the other half's unread fp32 inputs)
                    umma_f16_ts(tmem + d_col, tmem + a_col + (k >> 1) * 32 + (k & 1) * 8,
                                make_smem_desc_sw128(b_base + k * (16 * 128), BQ * 128, 1024), id_kv, acc || k != 0);
            };
            mbar_wait(kv_full, 0);
            mbar_wait(&qdo_full[0], 0);
            tc_fence_after();
            issue_st(sK, smem_u32(smem + BwdSmem::Q), T_S0);
            umma_commit<1>(&s_full[0]);
            // dP^T(0)
            issue_st(sV, smem_u32(smem + BwdSmem::DO), T_DP);
            umma_commit<1>(dp_full);
            for (int i = 0; i < n_steps; ++i) {
                const int st = i & 1;
                const uint32_t ph2 = (i >> 1) & 1;
                const int qs = i % QSTAGES, nqs = (i + 1) % QSTAGES;
                const uint32_t sQ = smem_u32(smem + BwdSmem::Q + qs * QT_BYTES);
                const uint32_t sDO = smem_u32(smem + BwdSmem::DO + qs * QT_BYTES);
                // dV += P^T dO
                mbar_wait(&p_ready[st], ph2);
                tc_fence_after();
                issue_acc(T_S0 + st * 64, sDO, T_DV, i > 0);
                // S^T(i+1)
                const int ns = (i + 1) & 1;
                if (i + 1 < n_steps) {
                    mbar_wait(&qdo_full[nqs], ((i + 1) / QSTAGES) & 1);
                    tc_fence_after();
                    issue_st(sK, smem_u32(smem + BwdSmem::Q + nqs * QT_BYTES), T_S0 + ns * 64);
                    umma_commit<1>(&s_full[ns]);
                }
                // dK += dS^T Q
                mbar_wait(ds_ready, i & 1);
                tc_fence_after();
                issue_acc(T_DP, sQ, T_DK, i > 0);
                // dP^T(i+1) as early as possible (its TMEM region is free once dK(i) has been issued: in-order pipe)
                if (i + 1 < n_steps) {
                    issue_st(sV, smem_u32(smem + BwdSmem::DO + nqs * QT_BYTES), T_DP);
                    umma_commit<1>(dp_full);
                }
                // dQ^T(i) = K^T dS^T
                if (i > 0) { mbar_wait(dq_free, (i - 1) & 1); tc_fence_after(); }
                {
                    const uint32_t sDS = smem_u32(smem + BwdSmem::DS + st * QT_BYTES);
#pragma unroll
                    for (int k = 0; k < TN / 16; ++k)
                        umma_f16_ss<1>(tmem + T_DQ, make_smem_desc_sw128(sK + k * (16 * 128), TN * 128, 1024),
                                       make_smem_desc_sw128(sDS + k * (16 * 128), TN * 128, 1024), id_dq, k != 0);
                }
                umma_commit<1>(dq_full);
                umma_commit<1>(&ds_free[st]);
                umma_commit<1>(&qdo_empty[qs]);  // Q(i) / dO(i) fully consumed (dV, dK and the early dP^T(i+1) excluded)
            }
            umma_commit<1>(dkv_done);
        }
    } else if (warp >= 4) {
        // ===================== P / dS / dQ warps (8): lane = kv row (P, dS) or head-dim index (dQ^T) ==============
        // warps 4-7 own q columns [0,32) of the step, warps 8-11 columns [32,64): two warps per scheduler hide each
        // other's TMEM / MUFU latency (with one warpgroup the step was latency-bound at ~7 us, see profiles/).
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;     // which 32-column half of the 64 q columns
        const int r = q * 32 + lane;          // TMEM lane: kv row inside the tile (or d index for dQ^T)
        const int kv = kv0 + r;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const float sl2 = f.scale_log2;
        const int tid = threadIdx.x - 128;    // 0..255
        float* sld = reinterpret_cast<float*>(smem + BwdSmem::LD);  // [2][2][64]: buffer, {lse2, delta}, q
        float* stage = reinterpret_cast<float*>(smem + BwdSmem::DQ);
        const bool issuer = (tid == 0);
        // (head j, q tile mq) of a step are tracked incrementally: no integer division in the loop
        auto load_rowstats = [&](int step, int j, int mq) {
            if (tid < 2 * BQ && step < n_steps) {
                const int h = hk * qpk + j;
                if constexpr (SP) {
                    // rows outside the sequence are masked later; the tile lies inside one peer's buffer, so no clamp
                    const int pr = (mq * BQ) / sp.t_local;
                    const int qi = mq * BQ - pr * sp.t_local + (tid & (BQ - 1));
                    const float* src = (tid < BQ ? sp.lse2[pr] : sp.delta[pr]) + (int64_t)h * f.T + qi;
                    sld[(step & 1) * 2 * BQ + tid] = *reinterpret_cast<const volatile float*>(src);
                } else {
                    const int qi = min(mq * BQ + (tid & (BQ - 1)), len - 1);
                    const float* src = (tid < BQ ? args.lse : args.delta) + (int64_t)h * f.T + s0 + qi;
                    sld[(step & 1) * 2 * BQ + tid] = __ldg(src);
                }
            }
        };
        auto reduce_dq = [&](int step, int j, int mq) {
            // dQ^T(step): lane = d, this warp's 32 q columns -> fp32 staging rows -> one TMA reduce-add for the tile
            const int h = hk * qpk + j;
            mbar_wait(dq_full, step & 1);
            tc_fence_after();
            uint32_t x[32];
            tmem_ld_32x32b_x32(tmem + T_DQ + lane_off + half * 32, x);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_free);
            if (issuer) tma_store_wait_read<0>();  // previous reduce has finished reading the staging tile
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) stage[(half * 32 + c) * D + r] = __uint_as_float(x[c]);
            fence_proxy_async();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (issuer) {
                const CUtensorMap* mdq = &tmap_dq;
                int qrow = s0 + mq * BQ;
                if constexpr (SP) {   // the owner's accumulator: a TMA reduce-add into peer memory over NVLink
                    const int pr = (mq * BQ) / sp.t_local;
                    mdq = &peers.dq[pr];
                    qrow = mq * BQ - pr * sp.t_local;
                }
                asm volatile(
                    "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                    ::"l"(reinterpret_cast<uint64_t>(mdq)), "r"(smem_u32(stage)), "r"(h * D), "r"(qrow)
                    : "memory");
                tma_store_commit();
            }
        };
        load_rowstats(0, 0, mq0);
        int cj = 0, cm = mq0;          // step i
        int pj = 0, pm = mq0;          // step i - 1
        for (int i = 0; i < n_steps; ++i) {
            const int st = i & 1;
            const int mq = cm;
            const int q0 = (SP ? mq * BQ - s0 : mq * BQ) + half * 32;   // first q row (sequence-relative) handled by this warp
            int nj = cj, nm = cm + 1;  // step i + 1
            if (nm == mq0 + nq) { nm = mq0; ++nj; }
            asm volatile("bar.sync 2, 256;" ::: "memory");  // row stats of step i are in smem
            load_rowstats(i + 1, nj, nm);
            const float* lse2 = sld + st * 2 * BQ + half * 32;
            const float* delt = lse2 + BQ;
            const bool edge = (q0 + 32 > len) || (kv0 + TN > kv_end) || (f.causal && q0 < kv0 + TN) || (SP && q0 < 0);
            uint32_t pk[16];  // P^T row (this warp's 32 q columns), bf16 pairs
            float pf[32];     // the same values in fp32, reused by phase B (no unpack)
            // ---- phase A: P^T = exp2(S^T * scale_log2 - lse2)
            mbar_wait(&s_full[st], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t t_s = tmem + T_S0 + st * 64 + lane_off;
            {
                uint32_t x[32];
                tmem_ld_32x32b_x32(t_s + half * 32, x);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(lse2 + e);
                    float p0 = fast_exp2(fmaf(__uint_as_float(x[e]), sl2, -l4.x));
                    float p1 = fast_exp2(fmaf(__uint_as_float(x[e + 1]), sl2, -l4.y));
                    float p2 = fast_exp2(fmaf(__uint_as_float(x[e + 2]), sl2, -l4.z));
                    float p3 = fast_exp2(fmaf(__uint_as_float(x[e + 3]), sl2, -l4.w));
                    if (edge) {
                        const int qa = q0 + e;
                        const bool kvok = kv < kv_end;
                        // unsigned compare: a q row before the sequence start (SP: globally aligned tiles) is out, too
                        if (!(kvok && (unsigned)qa < (unsigned)len && (!f.causal || kv <= qa))) p0 = 0.f;
                        if (!(kvok && (unsigned)(qa + 1) < (unsigned)len && (!f.causal || kv <= qa + 1))) p1 = 0.f;
                        if (!(kvok && (unsigned)(qa + 2) < (unsigned)len && (!f.causal || kv <= qa + 2))) p2 = 0.f;
                        if (!(kvok && (unsigned)(qa + 3) < (unsigned)len && (!f.causal || kv <= qa + 3))) p3 = 0.f;
                    }
                    pk[e >> 1] = pack_bf16(p0, p1);
                    pk[(e >> 1) + 1] = pack_bf16(p2, p3);
                    pf[e] = p0; pf[e + 1] = p1; pf[e + 2] = p2; pf[e + 3] = p3;
                }
            }
            tmem_st_32x32b_x16(t_s + half * 32, pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[st]);
            // ---- phase B: dS^T = P^T * (dP^T - delta); the softmax scale is applied once to dK (epilogue) and dQ (convert)
            mbar_wait(dp_full, i & 1);
            if (i >= 2) mbar_wait(&ds_free[st], ((i >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t t_dp = tmem + T_DP + lane_off;
            uint8_t* ds_row = smem + BwdSmem::DS + st * QT_BYTES + r * 128;
            {
                uint32_t x[32];
                tmem_ld_32x32b_x32(t_dp + half * 32, x);
                tmem_ld_wait();
                uint32_t o[16];
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 d4 = *reinterpret_cast<const float4*>(delt + e);
                    o[e >> 1] = pack_bf16(pf[e] * (__uint_as_float(x[e]) - d4.x),
                                          pf[e + 1] * (__uint_as_float(x[e + 1]) - d4.y));
                    o[(e >> 1) + 1] = pack_bf16(pf[e + 2] * (__uint_as_float(x[e + 2]) - d4.z),
                                                pf[e + 3] * (__uint_as_float(x[e + 3]) - d4.w));
                }
                tmem_st_32x32b_x16(t_dp + half * 32, o);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {  // 4 x 16-byte chunks = this warp's 32 q values of row r
                    const int chunk = half * 4 + ch;
                    *reinterpret_cast<uint4*>(ds_row + ((chunk ^ (r & 7)) << 4)) =
                        make_uint4(o[ch * 4], o[ch * 4 + 1], o[ch * 4 + 2], o[ch * 4 + 3]);
                }
            }
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_ready);
            // ---- dQ^T of the previous step (long finished): transpose + bulk reduce-add
            if (i > 0) reduce_dq(i - 1, pj, pm);
            pj = cj; pm = cm; cj = nj; cm = nm;
        }
        if (n_steps > 0) reduce_dq(n_steps - 1, pj, pm);
        if (issuer) tma_store_wait<0>();
        // ---- epilogue: dV, dK rows of this kv tile (each warpgroup-half writes 64 of the 128 head-dim columns)
        mbar_wait(dkv_done, 0);
        tc_fence_after();
        const bool valid = kv < kv_end;
        const int64_t tok = (int64_t)s0 + kv - w0;
        for (int which = 0; which < 2; ++which) {
            const uint32_t t_src = tmem + (which == 0 ? T_DV : T_DK) + lane_off;
            const float mul = which == 0 ? 1.f : f.scale;
            __nv_bfloat16* dst = which == 0 ? args.dv + tok * args.dv_stride_t + (int64_t)hk * args.dv_stride_h
                                            : args.dk + tok * args.dk_stride_t + (int64_t)hk * args.dk_stride_h;
#pragma unroll 1
            for (int c = half * 64; c < half * 64 + 64; c += 32) {
                uint32_t x[32];
                tmem_ld_32x32b_x32(t_src + c, x);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        uint4 o4;
                        o4.x = pack_bf16(mul * __uint_as_float(x[e]), mul * __uint_as_float(x[e + 1]));
                        o4.y = pack_bf16(mul * __uint_as_float(x[e + 2]), mul * __uint_as_float(x[e + 3]));
                        o4.z = pack_bf16(mul * __uint_as_float(x[e + 4]), mul * __uint_as_float(x[e + 5]));
                        o4.w = pack_bf16(mul * __uint_as_float(x[e + 6]), mul * __uint_as_float(x[e + 7]));
                        *reinterpret_cast<uint4*>(dst + c + e) = o4;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<1>(tmem, 512);
    }
}

int attn_bwd(const AttnBwdDesc& b, cudaStream_t stream) {
    const AttnDesc& d = b.f;
    if (d.D != D) return -10;
    if (d.T == 0) return 0;
    const bool sp = d.sp_world > 1;
    // phases: 0 = everything (single rank); sequence parallel runs 1 (dO.O row sums + lse*log2e into the symmetric stats
    // buffer), 2 (main kernel: reads the peers' Q / dO / stats, reduce-adds dQ into the owners' accumulators) and 3 (dQ
    // fp32 -> bf16) with a device barrier of the sequence group between them (parallel/sp_attention.py)
    const int phase = b.phase;
    if (sp && (phase < 1 || phase > 3)) return -15;
    if (phase == 0 || phase == 1) {
        const int64_t warps = (int64_t)d.T * d.H;
        launch_pdl(attn_bwd_delta_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, stream, 1,
            (const __nv_bfloat16*)b.dout, (const __nv_bfloat16*)d.o, d.lse, b.delta, d.T, d.H);
    }
    if (phase == 0 || phase == 2) {
        CUtensorMap tq, tk, tv, tdo;
        if (make_tmap_2d_bf16(&tq, d.q, (uint64_t)d.q_stride_t, d.T, (uint64_t)d.q_stride_t, 64, BQ)) return -11;
        if (make_tmap_2d_bf16(&tk, d.k, (uint64_t)d.k_stride_t, d.T, (uint64_t)d.k_stride_t, 64, TN)) return -11;
        if (make_tmap_2d_bf16(&tv, d.v, (uint64_t)d.v_stride_t, d.T, (uint64_t)d.v_stride_t, 64, TN)) return -11;
        if (make_tmap_2d_bf16(&tdo, b.dout, (uint64_t)d.H * D, d.T, (uint64_t)d.H * D, 64, BQ)) return -11;
        CUtensorMap tdq;
        if (make_tmap_2d_f32_noswizzle(&tdq, b.dq_acc, (uint64_t)d.H * D, d.T, (uint64_t)d.H * D, D, BQ)) return -11;
        AttnBwdArgs a;
        a.f.o = (__nv_bfloat16*)d.o; a.f.lse = d.lse; a.f.cu_seqlens = d.cu_seqlens; a.f.num_seqs = d.num_seqs;
        a.f.T = d.T; a.f.H = d.H; a.f.Hkv = d.Hkv;
        a.f.q_stride_g = d.q_stride_g ? d.q_stride_g : d.q_stride_h * (d.H / d.Hkv);
        a.f.q_stride_h = d.q_stride_h; a.f.k_stride_h = d.k_stride_h; a.f.v_stride_h = d.v_stride_h;
        a.f.scale = d.scale; a.f.scale_log2 = d.scale * 1.4426950408889634f; a.f.causal = d.causal;
        a.lse = b.delta + (int64_t)d.H * d.T;  // lse * log2e, written by the delta pre-pass
        a.delta = b.delta; a.dq_acc = b.dq_acc;
        a.dk = (__nv_bfloat16*)b.dk; a.dv = (__nv_bfloat16*)b.dv;
        a.dk_stride_t = b.dk_stride_t; a.dk_stride_h = b.dk_stride_h; a.dv_stride_t = b.dv_stride_t; a.dv_stride_h = b.dv_stride_h;
        static bool attr = false;
        if (!attr) {
            if (cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL) !=
                    cudaSuccess ||
                cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL) !=
                    cudaSuccess)
                return -12;
            attr = true;
        }
        dim3 grid(d.Hkv, upper_qblocks(d.T, d.num_seqs, TN));
        BwdPeers pm;
        BwdSpArgs spa{d.sp_rank, d.sp_world, d.T, {}, {}};
        if (sp) {
            if (d.sp_world > 8 || d.T % TN != 0 || !b.q_peers || !b.dout_peers || !b.dq_acc_peers || !b.delta_peers) return -14;
            for (int p = 0; p < 8; ++p) {
                const int s_ = p < d.sp_world ? p : 0;
                if (make_tmap_2d_bf16(&pm.q[p], b.q_peers[s_], (uint64_t)d.q_stride_t, d.T, (uint64_t)d.q_stride_t, 64, BQ)) return -11;
                if (make_tmap_2d_bf16(&pm.dout[p], b.dout_peers[s_], (uint64_t)d.H * D, d.T, (uint64_t)d.H * D, 64, BQ)) return -11;
                if (make_tmap_2d_f32_noswizzle(&pm.dq[p], b.dq_acc_peers[s_], (uint64_t)d.H * D, d.T, (uint64_t)d.H * D, D, BQ))
                    return -11;
                spa.delta[p] = reinterpret_cast<const float*>(b.delta_peers[s_]);
                spa.lse2[p] = spa.delta[p] + (int64_t)d.H * d.T;
            }
            launch_pdl(attn_bwd_kernel<true>, dim3(grid), dim3(BWD_THREADS), BwdSmem::TOTAL, stream, 1, tq, tk, tv, tdo, tdq, a, spa,
                       pm);
        } else {
            pm.q[0] = tq;   // unused by the single-rank instantiation
            launch_pdl(attn_bwd_kernel<false>, dim3(grid), dim3(BWD_THREADS), BwdSmem::TOTAL, stream, 1, tq, tk, tv, tdo, tdq, a, spa,
                       pm);
        }
    }
    if (phase == 0 || phase == 3) {
        const int qpk = d.H / d.Hkv;
        const int64_t n = (int64_t)d.T * d.H * (D / 8);
        const int64_t sg = b.dq_stride_g ? b.dq_stride_g : b.dq_stride_h * qpk;
        launch_pdl(attn_bwd_dq_convert_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, 1,
            b.dq_acc, (__nv_bfloat16*)b.dq, d.T, d.H, qpk, b.dq_stride_t, sg, b.dq_stride_h, d.scale);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -13;
}

}  // namespace b200
