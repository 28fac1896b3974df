// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (fp32 accumulators in TMEM,
// double-buffered) -> tcgen05.ld epilogue with fused bias / accumulate / SwiGLU / fp32-out.
//
//   D[M,N] (+)= A[M,K] * B[N,K]^T        (both operands may be K-major or MN-major in memory, so the same kernel
//                                          serves forward  Y = X W^T   (A K-major, B K-major),
//                                                 dgrad    dX = dY W   (A K-major, B MN-major),
//                                                 wgrad    dW = dY^T X (A MN-major, B MN-major))
//
// This replaces the cuBLAS / cublasLt calls of the reference (F.linear + fused_dense_lib.linear_bias_wgrad,
// reference internlm/model/utils.py:228-586) and is the building block of the fused compute+collective kernels.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
#include "gemm_sm100.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "launch.h"
#include "sm100_ptx.cuh"

namespace b200 {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B atom row
static constexpr int UMMA_K = 16;
static constexpr int NUM_THREADS = 192;

// CG = CTAs cooperating on one tile (tcgen05 cta_group): with CG == 2 a cluster of two CTAs computes a 256 x BN tile,
// every CTA stages its own 128 rows of A and HALF of B, and the pair's tensor cores read both halves -> one third less
// shared-memory traffic per SM and room for a deeper ring.
template <int BN, int CG = 1>
struct GemmCfg {
    static constexpr int STAGES = (BN == 256 && CG == 1) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (BN / CG) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;  // double-buffered accumulator (allocations are powers of 2)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int SMEM_BYTES_COMM = SMEM_BYTES + 4 * 8192;  // + per-warp transpose staging of the RS push
};

struct GemmKernelArgs {
    int M, N, K;
    void* D;
    int64_t ldd;
    const __nv_bfloat16* bias;
    void* H;
    int64_t ldh;
    int flags;
    int a_mn, b_mn;
    int tiles_m, tiles_n;
    int group_m;      // raster group height in tile rows
    int tiles_full;   // work items [0, tiles_full) are whole BM x BN tiles
    int tail_split;   // the remaining tiles are split into this many column slices (1, 2 or 4) to fill the last wave
    // ---- grouped GEMM (MoE experts): `grp_num` problems in ONE launch, their sizes read from DEVICE memory ---------------
    // grp_off[g] .. grp_off[g + 1] are the rows of group g in the packed activation buffer (multiples of 128, padding rows
    // zero).  mode 1 (forward / dgrad): rows of A and D, B = weights of group g through its own tensor map b_maps[g];
    // mode 2 (wgrad): the contraction runs over the group's rows, D = d_ptrs[g].
    const int* grp_off;
    int grp_num, grp_mode;
    const CUtensorMap* b_maps;   // global memory, one map per group (mode 1)
    void* const* d_ptrs;         // global memory, one output per group (mode 2)
};

static constexpr int MAX_GROUPS = 64;


// Raster: groups of `GROUP_M` tile rows are swept column by column (8 rows: square-ish footprint of one wave of tiles;
// taller groups were not measurably better on the 7B shapes, see profiles/gemm_perf_r1_v4_group_ab.json).
B200_DEVICE void tile_coords(int tile, int tiles_m, int tiles_n, int& m, int& n, int GROUP_M) {
    const int per_group = GROUP_M * tiles_n;
    const int g = tile / per_group;
    const int first_m = g * GROUP_M;
    const int gsize = min(tiles_m - first_m, GROUP_M);
    const int r = tile - g * per_group;
    m = first_m + r % gsize;
    n = r / gsize;
}

// work item -> (m tile, n tile, column offset inside the tile, slice width)
template <int BN>
B200_DEVICE void item_coords(int item, const GemmKernelArgs& a, int& tm, int& tn, int& n_off, int& width) {
    int tile = item, slice = 0;
    width = BN;
    if (item >= a.tiles_full) {
        const int w = item - a.tiles_full;
        tile = a.tiles_full + w / a.tail_split;
        slice = w % a.tail_split;
        width = BN / a.tail_split;
    }
    tile_coords(tile, a.tiles_m, a.tiles_n, tm, tn, a.group_m);
    n_off = slice * width;
}

// A "block" below is a 128-row slab of the output (one CTA's share of a tile): with CG == 2 the pair's tile tm covers
// blocks 2 tm and 2 tm + 1.  Flags, ownership and staging are all per block, so the 1-CTA and the 2-CTA kernels share the
// protocol.
// An all-gather block (128 rows) is pushed as 16 pieces of 8 rows by different CTAs; every piece adds 1 to the block's
// arrival COUNTER on the destination.  Counters only grow: call number c of a buffer waits for 16 * c (`comm.epoch` carries
// that target), so nothing is ever reset and a fast peer that is already pushing call c + 1 cannot be mistaken for call c
// (its own call-c pieces were counted first).
static constexpr int AG_PARTS = 16;

struct CommKernelArgs {
    int mode;
    void* const* peer_ptrs;        // AG: every rank's gathered A [M, K]; RS/AR: every rank's staging [world, m_local, N]
    uint32_t* const* flags_ptrs;   // every rank's flag words for this operation
    void* const* out_ptrs;         // AR: every rank's output [M, N]
    uint32_t* const* done_ptrs;    // AR: every rank's `world` completion words (end-of-kernel handshake)
    uint32_t* done_counter;        // AR: local counter of finished CTAs (wraps to 0 by atomicInc)
    int rank, world;
    uint32_t epoch;
    int m_local;          // rows per rank
    void* out_local;      // RS: reduced rows [m_local, N]
    int64_t ld_out;
    const void* x_local;  // AG: this rank's shard [m_local, K] (contiguous); weight gather: this rank's weight shard
    float out_scale;      // RS: factor applied to the reduced rows (1 / W: the weight-parallel gradient is an average)
};

B200_DEVICE void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
B200_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
B200_DEVICE uint4 ld_cg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
B200_DEVICE void st_v4(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Rotation of the m-tile order: AG starts on the local shard (no waiting), then the shards in the order the peers push
// them; RS/AR starts with the rows of rank+1 (pushed to their owner first) and finishes on the rows this rank owns.
B200_DEVICE int comm_remap_m(int m, int tiles_m, const CommKernelArgs& c) {
    if (c.mode == GEMM_COMM_NONE || c.mode == GEMM_COMM_GATHER_B) return m;
    const int per = tiles_m / c.world;
    const int shift = (c.mode == GEMM_COMM_ALL_GATHER ? c.rank : c.rank + 1) * per;
    return (m + shift) % tiles_m;
}

// weight gather with a K-major B: the n tiles of this rank's own shard come first, then the shards in arrival order
B200_DEVICE int comm_remap_n(int n, int tiles_n, const CommKernelArgs& c, int b_mn) {
    if (c.mode != GEMM_COMM_GATHER_B || b_mn) return n;
    return (n + c.rank * (tiles_n / c.world)) % tiles_n;
}

// ---- all-gather push: the four epilogue warps of EVERY CTA, before their first tile is due --------------------------------
// The local shard [m_local rows of `row_bytes`] goes to every peer's gathered buffer in pieces of 8 rows: 128 threads move
// 16 bytes each per step with plain vector loads / stores - NVLink writes are posted, so nothing waits for a round trip -
// and the piece's flag is released on the destination once the CTA's stores are ordered before it.  With ~148 CTAs pushing,
// no single SM's store path limits the link; the accumulators are double-buffered, so the tensor cores run two tiles ahead
// while the epilogue warps are busy here.  Destinations are served one after the other, nearest consumer first: rank d
// consumes the shards in the order d, d + 1, ..., so shard s goes to s - 1, then s - 2, ... and every destination receives
// whole shards in exactly the order its GEMM asks for them (a ring schedule at full NVSwitch bandwidth).  `self_copy`
// appends the local copy (activations: wgrad reads the gathered buffer later; weights are read in place instead).
struct PushState {
    int q, p;   // destination index, piece index of the NEXT piece this CTA pushes
};
B200_DEVICE bool push_done(const PushState& st, const CommKernelArgs& c, bool self_copy) {
    return st.q >= c.world - 1 + (self_copy ? 1 : 0);
}
// one piece (8 rows) to one destination; all 128 epilogue threads take part (named barrier 1)
B200_DEVICE void push_one_piece(PushState& st, const CommKernelArgs& c, int64_t row_bytes, int t /*0..127*/) {
    const int blocks_local = c.m_local / BM;
    const int pieces = blocks_local * AG_PARTS;
    const int64_t piece_bytes = (int64_t)(BM / AG_PARTS) * row_bytes;
    if (st.p < pieces) {
        const int dest = (c.rank - 1 - st.q + 2 * c.world) % c.world;   // q == world - 1: this rank itself
        const uint8_t* src = reinterpret_cast<const uint8_t*>(c.x_local) + (int64_t)st.p * piece_bytes;
        uint8_t* dst = reinterpret_cast<uint8_t*>(c.peer_ptrs[dest]) + (int64_t)c.rank * c.m_local * row_bytes +
                       (int64_t)st.p * piece_bytes;
        // 16 x 16-byte loads in flight per thread: next to a GEMM that saturates L2 the copy needs the depth to keep the
        // link busy (tools/peer_copy_bench.py: 4 in flight reach 570 GB/s on an idle GPU but starve beside the GEMM)
        constexpr int U = 16;
        for (int64_t o = (int64_t)t * 16; o < piece_bytes; o += (int64_t)U * 128 * 16) {
            uint4 v[U];
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (o + j * 2048 < piece_bytes) v[j] = ld_nc_v4(src + o + j * 2048);
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (o + j * 2048 < piece_bytes) st_v4(dst + o + j * 2048, v[j]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // publish the piece on its destination: one arrival on the counter of its 128-row block (release: the CTA's
        // stores, ordered before this thread by the barrier, are visible before the count)
        if (t == 0) red_add_release_sys(c.flags_ptrs[dest] + c.rank * blocks_local + st.p / AG_PARTS, 1u);
    }
    st.p += gridDim.x;
    if (st.p >= pieces) { st.p = blockIdx.x; ++st.q; }
}

template <int CG>
B200_DEVICE void grp_coords(int item, const GemmKernelArgs& a, const int* s_off, const int* s_start, int& g, int& tm, int& tn) {
    if (a.grp_mode == 1) {
        g = 0;
        while (g + 1 < a.grp_num && item >= s_start[g + 1]) ++g;
        const int local = item - s_start[g];
        const int tiles_m_g = (s_off[g + 1] - s_off[g] + BM * CG - 1) / (BM * CG);
        tm = local % tiles_m_g;   // m fastest: consecutive units share the (large) weight tile through L2
        tn = local / tiles_m_g;
    } else {
        const int per = a.tiles_m * a.tiles_n;
        g = item / per;
        tile_coords(item - g * per, a.tiles_m, a.tiles_n, tm, tn, a.group_m);
    }
}

template <int BN, bool COMM, int CG = 1, bool GRP = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_bt, const GemmKernelArgs args, const CommKernelArgs comm) {
    static_assert(CG == 1 || BN == 256 || BN == 192, "2-CTA tiles: BN = 256, or 192 for a K-major B (wave quantisation)");
    static_assert(!(GRP && COMM), "grouped GEMM has no fused collective");
    __shared__ int s_off[GRP ? MAX_GROUPS + 1 : 1], s_start[GRP ? MAX_GROUPS + 1 : 1];
    using Cfg = GemmCfg<BN, CG>;
    griddep_launch_dependents();  // PDL (launch.h): the next kernel's CTAs may take over SMs this grid's tail has left
    const int cta_rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
    // persistent schedule over work units (a unit = one CTA, or one CTA pair)
    const int grid_ctas = static_cast<int>(gridDim.x) / CG;
    const int unit = static_cast<int>(blockIdx.x) / CG;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = args.tiles_m * args.tiles_n;
    int num_items = args.tiles_full + (num_tiles - args.tiles_full) * args.tail_split;
    const int num_kb = (args.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * CG);  // the leader collects the epilogue warps of the whole pair
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<CG>(tmem_base_smem, Cfg::TMEM_COLS);
    griddep_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
    if constexpr (GRP) {
        // group table: row offsets (written by the routing kernels earlier in the stream) and first work item of each group
        if (threadIdx.x == 64) {
            int acc_items = 0;
            for (int g = 0; g <= args.grp_num; ++g) {
                s_off[g] = args.grp_off[g];
                s_start[g] = acc_items;
                if (g < args.grp_num && args.grp_mode == 1)
                    acc_items += (args.grp_off[g + 1] - args.grp_off[g] + BM * CG - 1) / (BM * CG) * args.tiles_n;
            }
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync();  // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    if constexpr (GRP) num_items = args.grp_mode == 1 ? s_start[args.grp_num] : args.grp_num * num_tiles;

    if (warp == 0) {
        // ===================== TMA producer =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            uint64_t ready = 0;  // all-gather: blocks (of the first 64) whose pieces are known to have landed
            for (int tile = unit; tile < num_items; tile += grid_ctas) {
                int tm, tn, n_off, width, g = 0;
                if constexpr (GRP) { grp_coords<CG>(tile, args, s_off, s_start, g, tm, tn); n_off = 0; width = BN; }
                else item_coords<BN>(tile, args, tm, tn, n_off, width);
                bool own_rows = false;
                if constexpr (COMM) {
                    tm = comm_remap_m(tm, args.tiles_m, comm);
                    if (comm.mode == GEMM_COMM_ALL_GATHER) {
                        own_rows = tm / (args.tiles_m / comm.world) == comm.rank;
                        const int blk = tm * CG + cta_rank;
                        if (!own_rows && !(blk < 64 && ((ready >> blk) & 1))) {
                            // these rows are pushed into the local gathered buffer by their owner's epilogue warps: wait until
                            // all 16 pieces of the block have been counted
                            if (lane == 0) {
                                const uint32_t* f = comm.flags_ptrs[comm.rank] + blk;
                                while (static_cast<int32_t>(ld_acquire_sys(f) - comm.epoch) < 0) {
                                }
                            }
                            __syncwarp();
                            fence_proxy_async_all();
                            if (blk < 64) ready |= 1ull << blk;
                        }
                    }
                }
                int m0 = (tm * CG + cta_rank) * BM;
                int n0 = tn * BN + n_off + cta_rank * (width / CG);
                const int m0_own = m0 - comm.rank * comm.m_local;
                const CUtensorMap* map_b = &tmap_b;
                int kb_n = num_kb, k_base = 0;
                bool gb_k = false;   // weight gather along the contraction (dgrad): shards are walked inside the k loop
                if constexpr (COMM) {
                    if (comm.mode == GEMM_COMM_GATHER_B) {
                        if (!args.b_mn) {
                            // forward: this CTA's half of the n tile is ONE 128-row block of the gathered weight
                            tn = comm_remap_n(tn, args.tiles_n, comm, 0);
                            n0 = tn * BN + cta_rank * (BN / CG);
                            const int nblk = n0 / BM;
                            if (nblk / (comm.m_local / BM) == comm.rank) {      // own shard: read in place
                                map_b = &tmap_bt;
                                n0 -= comm.rank * comm.m_local;
                            } else if (!(nblk < 64 && ((ready >> nblk) & 1))) {
                                if (lane == 0) {
                                    const uint32_t* f = comm.flags_ptrs[comm.rank] + nblk;
                                    while (static_cast<int32_t>(ld_acquire_sys(f) - comm.epoch) < 0) {
                                    }
                                }
                                __syncwarp();
                                fence_proxy_async_all();
                                if (nblk < 64) ready |= 1ull << nblk;
                            }
                        } else {
                            gb_k = true;
                        }
                    }
                }
                if constexpr (GRP) {
                    if (args.grp_mode == 1) { m0 += s_off[g]; map_b = args.b_maps + g; }
                    else { k_base = s_off[g]; kb_n = (s_off[g + 1] - s_off[g]) / BK; }
                }
                // 2-CTA: every CTA loads its rows of A and its half of B, all bytes are counted on the leader's barrier
                auto ld = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
                    if constexpr (CG == 2) tma_load_2d_2sm(dst, map, bar, c0, c1);
                    else tma_load_2d(dst, map, bar, c0, c1);
                };
                const int nb64 = width / CG / 64;  // B arrives in 64-row (K-major) or 64-column (MN-major) boxes
                // every lane walks the k loop (only the weight-gather dgrad has work for lanes 1..31: polling the flags of
                // the next shard); lane 0 alone waits for free slots and issues the TMA loads
                // weight gather along k: the walk over the shards is tracked incrementally (the producer lane has ~500 cycles
                // per k block next to a CTA-pair MMA: no divisions, no local-memory state in this loop)
                int gk_seg = 0, gk_i = 0, gk_shard = 0, gk_kbs = 1;
                if constexpr (COMM) {
                    if (gb_k) { gk_kbs = kb_n / comm.world; gk_shard = comm.rank; }
                }
                for (int kb = 0; kb < kb_n; ++kb) {
                    int kk = kb, kb_off = 0;   // k block loaded for A / k offset of the B coordinate
                    const CUtensorMap* mb = map_b;
                    if constexpr (COMM) {
                        if (gb_k) {
                            kk = gk_shard * gk_kbs + gk_i;
                            if (gk_seg == 0) {
                                mb = &tmap_bt;                             // own shard, read in place
                                kb_off = -comm.rank * comm.m_local;
                            } else if ((gk_i & 1) == 0 && !((ready >> gk_shard) & 1)) {
                                // entering a new 128-row block (two k blocks) of a remote shard: the peers push the blocks of
                                // a shard in this very order, so the k loop streams behind the push instead of waiting for
                                // whole shards.  `ready` holds one bit per SHARD here: set once this CTA has walked it.
                                if (lane == 0) {
                                    const uint32_t* f = comm.flags_ptrs[comm.rank] + (kk >> 1);
                                    while (static_cast<int32_t>(ld_acquire_sys(f) - comm.epoch) < 0) {
                                    }
                                }
                                __syncwarp();
                                fence_proxy_async_all();
                            }
                            if (++gk_i == gk_kbs) {
                                if (gk_seg > 0) ready |= 1ull << gk_shard;
                                gk_i = 0;
                                ++gk_seg;
                                gk_shard = gk_shard + 1 == comm.world ? 0 : gk_shard + 1;
                            }
                        }
                    }
                    if (lane == 0) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES * CG + width * (BK * 2));
                        uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
                        uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
                        const int k0 = k_base + kk * BK;
                        if (COMM && own_rows) {  // the local shard is read in place (tmap_bt doubles as its map)
                            ld(sa, &tmap_bt, &full_bar[stage], k0, m0_own);
                        } else if (!args.a_mn) {
                            ld(sa, &tmap_a, &full_bar[stage], k0, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                ld(sa + j * (BK * 128), &tmap_a, &full_bar[stage], m0 + j * 64, k0);
                        }
                        if (width == BN) {
                            if (!args.b_mn) {
                                ld(sb, mb, &full_bar[stage], k0, n0);
                            } else {
#pragma unroll
                                for (int j = 0; j < BN / CG / 64; ++j)
                                    ld(sb + j * (BK * 128), mb, &full_bar[stage], n0 + j * 64, k0 + kb_off);
                            }
                        } else if (!args.b_mn) {  // tail slice: 64-row boxes of the K-major operand
                            for (int j = 0; j < nb64; ++j)
                                ld(sb + j * (64 * 128), &tmap_bt, &full_bar[stage], k0, n0 + j * 64);
                        } else {
                            for (int j = 0; j < nb64; ++j)
                                ld(sb + j * (BK * 128), &tmap_b, &full_bar[stage], n0 + j * 64, k0);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            __syncwarp();   // lanes 1..31 leave the loop early: reconverge before the warp-aligned teardown below
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && cta_rank == 0) {  // the leader CTA issues the MMAs of the pair
            const uint32_t idesc_full = make_idesc_f16(BM * CG, BN, args.a_mn, args.b_mn);
            // K-major: 8-row groups are 1024 B apart, advance 32 B per UMMA_K.
            // MN-major: 64-element column blocks are BK*128 B apart (LBO), 8-k groups 1024 B apart, advance 2048 B.
            const uint32_t a_lbo = args.a_mn ? BK * 128 : 16, b_lbo = args.b_mn ? BK * 128 : 16;
            const uint32_t a_kstep = args.a_mn ? (UMMA_K * 128) : (UMMA_K * 2);
            const uint32_t b_kstep = args.b_mn ? (UMMA_K * 128) : (UMMA_K * 2);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = unit; tile < num_items; tile += grid_ctas) {
                const uint32_t idesc = tile < args.tiles_full
                                           ? idesc_full
                                           : make_idesc_f16(BM * CG, BN / args.tail_split, args.a_mn, args.b_mn);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                int kb_n = num_kb;
                if constexpr (GRP) {
                    if (args.grp_mode == 2) {
                        const int g = tile / num_tiles;
                        kb_n = (s_off[g + 1] - s_off[g]) / BK;
                    }
                }
                for (int kb = 0; kb < kb_n; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
                    const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
                        const uint64_t db = make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
                        umma_f16_ss<CG>(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    umma_commit<CG>(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit<CG>(&tmem_full[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool out_f32 = args.flags & GEMM_OUT_F32;
        const bool accumulate = args.flags & GEMM_ACCUMULATE;
        const bool swiglu = args.flags & GEMM_SWIGLU;
        const bool no_store_d = args.flags & GEMM_SKIP_D;
        bool rs_mode = false;
        if constexpr (COMM) {
            rs_mode = comm.mode == GEMM_COMM_REDUCE_SCATTER || comm.mode == GEMM_COMM_ALL_REDUCE;
        }
        // all-gather forms: the local shard (activations, or the weight shard [N / W, K_in] under ISP) goes to every peer's
        // gathered buffer, piece by piece, IN BETWEEN the tile epilogues: whenever no accumulator is waiting the epilogue
        // warps push the next piece, so the tensor cores never stall behind the push and the push never waits for a tile
        PushState push{0, static_cast<int>(blockIdx.x)};
        bool pushing = false, push_self = false;
        int64_t push_row_bytes = 0;
        if constexpr (COMM) {
            pushing = comm.mode == GEMM_COMM_ALL_GATHER || comm.mode == GEMM_COMM_GATHER_B;
            push_self = comm.mode == GEMM_COMM_ALL_GATHER;
            push_row_bytes = (int64_t)((comm.mode == GEMM_COMM_GATHER_B && args.b_mn) ? args.N : args.K) * 2;
        }
        __shared__ int s_tile_ready[2];
        int poll_par = 0;
        for (int tile = unit; tile < num_items; tile += grid_ctas) {
            if constexpr (COMM) {
                while (pushing && !push_done(push, comm, push_self)) {
                    if (threadIdx.x == 64) s_tile_ready[poll_par] = mbar_try_wait(&tmem_full[acc], acc_phase) ? 1 : 0;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int rdy = s_tile_ready[poll_par];
                    poll_par ^= 1;
                    if (rdy) break;
                    push_one_piece(push, comm, push_row_bytes, static_cast<int>(threadIdx.x) - 64);
                }
            }
            int tm, tn, n_off, width, g = 0;
            if constexpr (GRP) { grp_coords<CG>(tile, args, s_off, s_start, g, tm, tn); n_off = 0; width = BN; }
            else item_coords<BN>(tile, args, tm, tn, n_off, width);
            if constexpr (COMM) {
                tm = comm_remap_m(tm, args.tiles_m, comm);
                tn = comm_remap_n(tn, args.tiles_n, comm, args.b_mn);
            }
            const int blk = tm * CG + cta_rank;   // 128-row block of this CTA
            int row = blk * BM + q * 32 + lane;
            int row_end = args.M;
            void* d_base = args.D;
            bool empty_k = false;
            if constexpr (GRP) {
                if (args.grp_mode == 1) { row += s_off[g]; row_end = s_off[g + 1]; }   // rows past the group belong to the next one
                else { d_base = args.d_ptrs[g]; empty_k = s_off[g + 1] == s_off[g]; }   // no rows: the accumulator was never written
            }
            const int n0 = tn * BN + n_off;
            const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
            if constexpr (COMM) {
                if (rs_mode) {
                    // ---- reduce-scatter / all-reduce -------------------------------------------------------------------
                    // A block owned by a peer: the bf16 partial goes straight into that peer's staging slot.  A block owned
                    // by this rank (scheduled last): add what the peers staged and write the final rows - to the local
                    // output (RS) or to every rank's output (AR).  Either way the 32 x 128 sub-tile of a warp is transposed
                    // through shared memory (16-byte units XOR-swizzled by row) so that 16 lanes store one contiguous
                    // 256-byte row segment: NVLink wants whole 128-byte requests.
                    const int per_blk = args.tiles_m * CG / comm.world;
                    const int owner = blk / per_blk;
                    const bool own = owner == comm.rank;
                    const int blk_local = blk - owner * per_blk;
                    const int row0_local = blk_local * BM + q * 32;   // first row of this warp inside the owner's shard
                    if (own) {
                        if (warp == 2 && lane < comm.world && lane != comm.rank) {
                            const uint32_t* f = comm.flags_ptrs[comm.rank] + ((int64_t)lane * per_blk + blk_local) * args.tiles_n + tn;
                            while (static_cast<int32_t>(ld_acquire_sys(f) - comm.epoch) < 0) {
                            }
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    }
                    mbar_wait(&tmem_full[acc], acc_phase);
                    tc_fence_after();
                    uint8_t* wst = smem + STAGES * Cfg::STAGE_BYTES + 256 + q * 8192;
                    const __nv_bfloat16* staged = reinterpret_cast<const __nv_bfloat16*>(comm.peer_ptrs[comm.rank]);
                    int ndst = 1;
                    if (own && comm.mode == GEMM_COMM_ALL_REDUCE) ndst = comm.world;
#pragma unroll 1
                    for (int half = 0; half < BN / 128; ++half) {
#pragma unroll 1
                        for (int c4 = 0; c4 < 4; ++c4) {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(taddr + half * 128 + c4 * 32, r);
                            tmem_ld_wait();
                            float v[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                            const int col = n0 + half * 128 + c4 * 32;
                            if (own && col < args.N) {
                                for (int pr = 1; pr < comm.world; ++pr) {
                                    const int sl = (comm.rank + pr) % comm.world;
                                    const __nv_bfloat16* src = staged + ((int64_t)sl * comm.m_local + row0_local + lane) * args.N + col;
#pragma unroll
                                    for (int i = 0; i < 32; i += 8) {
                                        if (col + i < args.N) {
                                            const uint4 x = ld_cg_v4(src + i);
                                            float2 a = unpack_bf16(x.x), b = unpack_bf16(x.y), cq = unpack_bf16(x.z),
                                                   dd = unpack_bf16(x.w);
                                            v[i] += a.x; v[i + 1] += a.y; v[i + 2] += b.x; v[i + 3] += b.y;
                                            v[i + 4] += cq.x; v[i + 5] += cq.y; v[i + 6] += dd.x; v[i + 7] += dd.y;
                                        }
                                    }
                                }
                            }
                            if (own && comm.mode == GEMM_COMM_REDUCE_SCATTER && col < args.N) {
                                // weight-parallel wgrad: average over the group and (from the second micro-batch on) add
                                // what the gradient arena already holds
                                if (comm.out_scale != 1.f) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) v[i] *= comm.out_scale;
                                }
                                if (accumulate) {
                                    const __nv_bfloat16* prev = reinterpret_cast<const __nv_bfloat16*>(comm.out_local) +
                                                                (int64_t)(row0_local + lane) * comm.ld_out + col;
#pragma unroll
                                    for (int i = 0; i < 32; i += 8) {
                                        if (col + i < args.N) {
                                            const uint4 x = *reinterpret_cast<const uint4*>(prev + i);
                                            float2 a = unpack_bf16(x.x), b = unpack_bf16(x.y), cq = unpack_bf16(x.z),
                                                   dd = unpack_bf16(x.w);
                                            v[i] += a.x; v[i + 1] += a.y; v[i + 2] += b.x; v[i + 3] += b.y;
                                            v[i + 4] += cq.x; v[i + 5] += cq.y; v[i + 6] += dd.x; v[i + 7] += dd.y;
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 o;
                                o.x = pack_bf16(v[8 * j], v[8 * j + 1]);
                                o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                                o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                                o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                                const int u = c4 * 4 + j;
                                *reinterpret_cast<uint4*>(wst + lane * 256 + ((u ^ (lane & 7)) << 4)) = o;
                            }
                        }
                        __syncwarp();
                        const int u = lane & 15;
                        const int col = n0 + half * 128 + u * 8;
                        for (int d = 0; d < ndst; ++d) {
                            __nv_bfloat16* dbase;
                            int64_t ldd;
                            if (!own) {            // the owner's staging slot of this rank
                                dbase = reinterpret_cast<__nv_bfloat16*>(comm.peer_ptrs[owner]) +
                                        ((int64_t)comm.rank * comm.m_local + row0_local) * args.N;
                                ldd = args.N;
                            } else if (comm.mode == GEMM_COMM_REDUCE_SCATTER) {
                                dbase = reinterpret_cast<__nv_bfloat16*>(comm.out_local) + (int64_t)row0_local * comm.ld_out;
                                ldd = comm.ld_out;
                            } else {               // every rank's output, peers first
                                dbase = reinterpret_cast<__nv_bfloat16*>(comm.out_ptrs[(comm.rank + 1 + d) % comm.world]) +
                                        (int64_t)(blk * BM + q * 32) * comm.ld_out;
                                ldd = comm.ld_out;
                            }
#pragma unroll 4
                            for (int rr = 0; rr < 32; rr += 2) {
                                const int rw = rr + (lane >> 4);
                                const uint4 o = *reinterpret_cast<const uint4*>(wst + rw * 256 + ((u ^ (rw & 7)) << 4));
                                if (col < args.N) st_v4(dbase + (int64_t)rw * ldd + col, o);
                            }
                        }
                        __syncwarp();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(&tmem_empty[acc], 0);
                        else mbar_arrive(&tmem_empty[acc]);
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    if (!own) {
                        // the partial block x n-tile has been stored into its owner's staging slot: publish it there
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (warp == 2 && lane == 0) {
                            fence_acq_rel_sys();
                            st_release_sys(comm.flags_ptrs[owner] + ((int64_t)comm.rank * per_blk + blk_local) * args.tiles_n + tn,
                                           comm.epoch);
                        }
                    }
                    continue;
                }
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < width; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + c, r);
                tmem_ld_wait();
                const int col = n0 + c;
                if (row < row_end && col < args.N) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = empty_k ? 0.f : __uint_as_float(r[i]);
                    const int ncols = min(32, args.N - col);  // multiple of 8 (host asserts N % 8 == 0)
                    if (args.bias != nullptr) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            if (i < ncols) {
                                uint4 b = *reinterpret_cast<const uint4*>(args.bias + col + i);
                                float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z),
                                       b3 = unpack_bf16(b.w);
                                v[i] += b0.x; v[i + 1] += b0.y; v[i + 2] += b1.x; v[i + 3] += b1.y;
                                v[i + 4] += b2.x; v[i + 5] += b2.y; v[i + 6] += b3.x; v[i + 7] += b3.y;
                            }
                        }
                    }
                    if (out_f32) {
                        float* d = reinterpret_cast<float*>(d_base) + static_cast<int64_t>(row) * args.ldd + col;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            if (i < ncols) {
                                float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                if (accumulate) {
                                    float4 p = *reinterpret_cast<float4*>(d + i);
                                    o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                                }
                                *reinterpret_cast<float4*>(d + i) = o;
                            }
                        }
                    } else {
                        __nv_bfloat16* d =
                            reinterpret_cast<__nv_bfloat16*>(d_base) + static_cast<int64_t>(row) * args.ldd + col;
                        if (!no_store_d) {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                if (i < ncols) {
                                    if (accumulate) {
                                        uint4 p = *reinterpret_cast<uint4*>(d + i);
                                        float2 p0 = unpack_bf16(p.x), p1 = unpack_bf16(p.y), p2 = unpack_bf16(p.z),
                                               p3 = unpack_bf16(p.w);
                                        v[i] += p0.x; v[i + 1] += p0.y; v[i + 2] += p1.x; v[i + 3] += p1.y;
                                        v[i + 4] += p2.x; v[i + 5] += p2.y; v[i + 6] += p3.x; v[i + 7] += p3.y;
                                    }
                                    uint4 o;
                                    o.x = pack_bf16(v[i], v[i + 1]);
                                    o.y = pack_bf16(v[i + 2], v[i + 3]);
                                    o.z = pack_bf16(v[i + 4], v[i + 5]);
                                    o.w = pack_bf16(v[i + 6], v[i + 7]);
                                    *reinterpret_cast<uint4*>(d + i) = o;
                                }
                            }
                        }
                        if (args.flags & GEMM_GELU) {
                            // H = gelu_tanh(D) on the bf16-rounded pre-activation (backward reads the stored D)
                            __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(args.H) +
                                               static_cast<int64_t>(row) * args.ldh + col;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                if (i < ncols) {
                                    float hv[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float x = __bfloat162float(__float2bfloat16_rn(v[i + j]));
                                        const float t = tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x));
                                        hv[j] = 0.5f * x * (1.f + t);
                                    }
                                    uint4 o;
                                    o.x = pack_bf16(hv[0], hv[1]);
                                    o.y = pack_bf16(hv[2], hv[3]);
                                    o.z = pack_bf16(hv[4], hv[5]);
                                    o.w = pack_bf16(hv[6], hv[7]);
                                    *reinterpret_cast<uint4*>(h + i) = o;
                                }
                            }
                        }
                        if (swiglu) {
                            // columns are interleaved (gate, up) pairs; h = silu(gate) * up, computed on the
                            // bf16-rounded values so that backward (which reads the stored gate/up) matches
                            __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(args.H) +
                                               static_cast<int64_t>(row) * args.ldh + (col >> 1);
#pragma unroll
                            for (int i = 0; i < 32; i += 16) {
                                if (i < ncols) {
                                    float hv[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        float g = __bfloat162float(__float2bfloat16_rn(v[i + 2 * j]));
                                        float u = __bfloat162float(__float2bfloat16_rn(v[i + 2 * j + 1]));
                                        hv[j] = g / (1.f + __expf(-g)) * u;
                                    }
                                    uint4 o;
                                    o.x = pack_bf16(hv[0], hv[1]);
                                    o.y = pack_bf16(hv[2], hv[3]);
                                    o.z = pack_bf16(hv[4], hv[5]);
                                    o.w = pack_bf16(hv[6], hv[7]);
                                    *reinterpret_cast<uint4*>(h + (i >> 1)) = o;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster(&tmem_empty[acc], 0);
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (COMM) {   // pieces left after the last tile (or a CTA without tiles)
            while (pushing && !push_done(push, comm, push_self))
                push_one_piece(push, comm, push_row_bytes, static_cast<int>(threadIdx.x) - 64);
        }
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync();  // no CTA of the pair may exit while the other still uses its smem / TMEM
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
    }
    if constexpr (COMM) {
        // all-reduce: the kernel must not complete before every peer has written its rows into this rank's output.  The last
        // CTA to finish (atomicInc wraps the counter back to 0 for the next call) tells every peer that this rank is done
        // and waits for the same word from each of them - the trailing device barrier, without a launch.
        if (comm.mode == GEMM_COMM_ALL_REDUCE && threadIdx.x == 0) {
            fence_acq_rel_sys();
            if (atomicInc(comm.done_counter, gridDim.x - 1) == gridDim.x - 1) {
                fence_acq_rel_sys();
                for (int pr = 1; pr < comm.world; ++pr)
                    st_release_sys(comm.done_ptrs[(comm.rank + pr) % comm.world] + comm.rank, comm.epoch);
                for (int pr = 1; pr < comm.world; ++pr) {
                    const uint32_t* f = comm.done_ptrs[comm.rank] + (comm.rank + pr) % comm.world;
                    while (static_cast<int32_t>(ld_acquire_sys(f) - comm.epoch) < 0) {
                    }
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
            fprintf(stderr, "[b200] cannot resolve cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(e));
            fn = nullptr;
        } else {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    return fn;
}

// 2D bf16 tensor map: `inner` contiguous elements, `outer` rows with `ld` elements stride, 128B swizzle
// Driver-API calls need a context bound to the calling thread; autograd worker threads may not have one yet
// (the runtime binds the primary context lazily), so make sure before encoding a tensor map.
static void ensure_context() {
    using CtxGetFn = CUresult (*)(CUcontext*);
    static CtxGetFn get_ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            get_ctx = reinterpret_cast<CtxGetFn>(p);
    });
    CUcontext ctx = nullptr;
    if (get_ctx == nullptr || get_ctx(&ctx) != CUDA_SUCCESS || ctx == nullptr) cudaFree(0);
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    ensure_context();
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[b200] cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u\n",
                (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems,
                box_inner, box_outer);
        return -2;
    }
    return 0;
}

// 2D fp32 tensor map without swizzle (TMA stores / reduce-adds of accumulator tiles)
int make_tmap_2d_f32_noswizzle(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                               uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    ensure_context();
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld_elems * 4};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[b200] cuTensorMapEncodeTiled(f32) failed (%d)\n", (int)r);
        return -2;
    }
    return 0;
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

// The 256 x 192 pair tile is OFF by default: measured (profiles/gemm_perf_r2_v1_bn192.json) it runs at ~70 % of the 256-wide
// tile's rate per tile (1090 vs 1553 TFLOPS on 8192^3), which eats the wave-quantisation gain it was meant to buy (wqkv
// forward: 1206 vs 1191 / 1240 without tail split).  B200_GEMM_BN192=1 turns the model-based choice on, force_bn = 384 pins it.
static int g_bn192 = [] {
    const char* e = std::getenv("B200_GEMM_BN192");
    return (e && e[0] == '1') ? 1 : 0;
}();
static int g_tail_split = 1;
static int g_group_m = 0;  // 0 = default (8)
void set_gemm_tail_split(int on) { g_tail_split = on; }
void set_gemm_group_m(int g) { g_group_m = g; }

template <int BN, int CG = 1>
static int launch(const GemmDesc& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, CG>;
    CUtensorMap ta, tb, tbt;
    int rc;
    // A: logical [M, K]
    if (!g.a_mn_major) rc = make_tmap_2d_bf16(&ta, g.A, g.K, g.M, g.lda, BK, BM);
    else               rc = make_tmap_2d_bf16(&ta, g.A, g.M, g.K, g.lda, 64, BK);
    if (rc) return rc;
    // B: logical [N, K]
    if (!g.b_mn_major) rc = make_tmap_2d_bf16(&tb, g.B, g.K, g.N, g.ldb, BK, BN / CG);
    else               rc = make_tmap_2d_bf16(&tb, g.B, g.N, g.K, g.ldb, 64, BK);
    if (rc) return rc;
    tbt = tb;  // tail slices of a K-major B come in 64-row boxes

    GemmKernelArgs a = {};
    a.M = g.M; a.N = g.N; a.K = g.K;
    a.D = g.D; a.ldd = g.ldd;
    a.bias = reinterpret_cast<const __nv_bfloat16*>(g.bias);
    a.H = g.H; a.ldh = g.ldh;
    a.flags = g.flags;
    a.a_mn = g.a_mn_major; a.b_mn = g.b_mn_major;
    a.tiles_m = (g.M + BM * CG - 1) / (BM * CG);
    a.tiles_n = (g.N + BN - 1) / BN;
    a.group_m = g_group_m > 0 ? g_group_m : 8;
    const int tiles = a.tiles_m * a.tiles_n;
    int sms = (g.max_ctas > 0 ? g.max_ctas : num_sms()) / CG;  // work units: CTAs or CTA pairs
    const int grid = tiles < sms ? tiles : sms;
    // Wave quantisation: split the tiles of the last (partial) wave into 2 or 4 column slices when that lets the
    // whole tail run as ONE wave of narrower tiles (e.g. 512 tiles on 148 SMs: 3 full waves + 68 tiles -> 136 half
    // tiles, 3.5 wave-times instead of 4).  The partial last n-tile (N % BN != 0) keeps the whole-tile path.
    a.tiles_full = tiles;
    a.tail_split = 1;
    if (g_tail_split && tiles > grid && g.N % BN == 0) {
        const int tail = tiles % grid;
        if (tail > 0) {
            int split = 1;  // every CTA of a pair must still stage whole 64-row boxes of B
            if (tail * 4 <= grid && BN / 4 / CG >= 64) split = 4;
            else if (tail * 2 <= grid && BN / 2 / CG >= 64) split = 2;
            if (split > 1) { a.tiles_full = tiles - tail; a.tail_split = split; }
        }
    }
    if (a.tail_split > 1 && !g.b_mn_major) {
        rc = make_tmap_2d_bf16(&tbt, g.B, g.K, g.N, g.ldb, BK, 64);
        if (rc) return rc;
    }

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, false, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            fprintf(stderr, "[b200] cudaFuncSetAttribute(smem=%d) failed: %s\n", Cfg::SMEM_BYTES, cudaGetErrorString(e));
            return -3;
        }
        attr_set = true;
    }
    // PDL launch (launch.h); CG = 2 adds the cluster dimension of the CTA pair
    cudaError_t e = launch_pdl(gemm_bf16_kernel<BN, false, CG>, dim3(grid * CG), dim3(NUM_THREADS), Cfg::SMEM_BYTES, stream, CG,
                               ta, tb, tbt, a, CommKernelArgs{});
    if (e != cudaSuccess) {
        fprintf(stderr, "[b200] gemm launch failed: %s\n", cudaGetErrorString(e));
        return -4;
    }
    return 0;
}

// GEMM fused with its tensor-parallel collective in ONE launch (see GemmCommArgs).
//   ALL_GATHER      A = all-gather of row shards: copy CTAs push the local shard into every peer's gathered buffer
//                   through the TMA unit; the GEMM starts on the local shard (read in place) and acquires per-chunk
//                   flags for the rest.
//   REDUCE_SCATTER  the epilogue stores partial tiles owned by a peer directly into that peer's staging slot (posted
//   / ALL_REDUCE    NVLink writes) and reduces the tiles this rank owns, which are scheduled last, with the slots the
//                   peers filled.  No extra CTAs, no second pass over the output.
int gemm_bf16_comm(const GemmDesc& g, const GemmCommArgs& c, cudaStream_t stream) {
    constexpr int BN = 256, CG = 2;   // the CTA-pair tile (cta_group::2), same as the plain GEMM's default
    using Cfg = GemmCfg<BN, CG>;
    const bool ag = c.mode == GEMM_COMM_ALL_GATHER, gb = c.mode == GEMM_COMM_GATHER_B;
    const bool rs = c.mode == GEMM_COMM_REDUCE_SCATTER || c.mode == GEMM_COMM_ALL_REDUCE;
    if (c.world < 2) return -20;
    if (g.a_mn_major && c.mode != GEMM_COMM_REDUCE_SCATTER) return -20;   // MN-major A: the wgrad -> reduce-scatter form only
    const int tiles_m = (g.M + BM * CG - 1) / (BM * CG), tiles_n = (g.N + BN - 1) / BN;
    if ((ag || rs) && g.M % (BM * CG * c.world) != 0) return -21;  // every rank owns whole 256-row tiles
    if (g.N % 8 != 0 || (g.K % 8 != 0 && !g.a_mn_major)) return -22;
    if (c.mode == GEMM_COMM_ALL_REDUCE && (c.done_ptrs == nullptr || c.done_counter == nullptr)) return -23;
    if (gb) {
        // weight shards of whole 128-row blocks; K-major B: whole n tiles per rank, MN-major B: whole k blocks per rank
        if (c.m_local % BM != 0) return -24;
        if (!g.b_mn_major && (g.N != c.m_local * c.world || g.N % (BN * c.world) != 0)) return -25;
        if (g.b_mn_major && (g.K != c.m_local * c.world)) return -26;
    }
    CUtensorMap ta, tb, tx;
    int rc;
    // A: the gathered buffer (AG) or the local activations (RS / AR / weight gather)
    if (!g.a_mn_major) rc = make_tmap_2d_bf16(&ta, ag ? c.out_local : g.A, g.K, g.M, ag ? (int64_t)g.K : g.lda, BK, BM);
    else               rc = make_tmap_2d_bf16(&ta, g.A, g.M, g.K, g.lda, 64, BK);
    if (rc) return rc;
    const void* bptr = gb ? c.out_local : g.B;     // weight gather: B is this rank's gathered buffer
    const int64_t ldb = gb ? (g.b_mn_major ? (int64_t)g.N : (int64_t)g.K) : g.ldb;
    if (!g.b_mn_major) rc = make_tmap_2d_bf16(&tb, bptr, g.K, g.N, ldb, BK, BN / CG);
    else               rc = make_tmap_2d_bf16(&tb, bptr, g.N, g.K, ldb, 64, BK);
    if (rc) return rc;
    tx = tb;
    if (ag) {  // this rank's shard, read in place
        rc = make_tmap_2d_bf16(&tx, c.x_local, g.K, c.m_local, g.K, BK, BM);
        if (rc) return rc;
    }
    if (gb) {  // this rank's weight shard [m_local, K_in], read in place
        if (!g.b_mn_major) rc = make_tmap_2d_bf16(&tx, c.x_local, g.K, c.m_local, g.ldb, BK, BN / CG);
        else               rc = make_tmap_2d_bf16(&tx, c.x_local, g.N, c.m_local, g.ldb, 64, BK);
        if (rc) return rc;
    }
    GemmKernelArgs a = {};
    a.tiles_full = tiles_m * tiles_n; a.tail_split = 1; a.group_m = 8;
    a.M = g.M; a.N = g.N; a.K = g.K;
    a.D = g.D; a.ldd = g.ldd;
    a.bias = nullptr; a.H = g.H; a.ldh = g.ldh; a.flags = g.flags;
    a.a_mn = g.a_mn_major; a.b_mn = g.b_mn_major;
    a.tiles_m = tiles_m; a.tiles_n = tiles_n;
    CommKernelArgs k;
    k.mode = c.mode; k.peer_ptrs = c.peer_ptrs; k.flags_ptrs = c.flags_ptrs; k.out_ptrs = c.out_ptrs;
    k.done_ptrs = c.done_ptrs; k.done_counter = c.done_counter;
    k.rank = c.rank; k.world = c.world; k.epoch = c.epoch; k.m_local = (int)c.m_local;
    k.out_local = c.out_local; k.ld_out = c.ld_out; k.x_local = c.x_local;
    k.out_scale = c.out_scale;
    int units = num_sms() / CG;
    // all-gather forms: every CTA also pushes a share of the local shard, so the whole machine is launched even for few tiles
    if (rs && units > tiles_m * tiles_n) units = tiles_m * tiles_n;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_bf16_kernel<BN, true, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES_COMM) != cudaSuccess)
            return -3;
        attr_set = true;
    }
    cudaError_t e = launch_pdl(gemm_bf16_kernel<BN, true, CG>, dim3(units * CG), dim3(NUM_THREADS), Cfg::SMEM_BYTES_COMM, stream,
                               CG, ta, tb, tx, a, k);
    if (e != cudaSuccess) {
        fprintf(stderr, "[b200] comm gemm launch failed: %s\n", cudaGetErrorString(e));
        return -4;
    }
    return 0;
}

int grouped_b_map(CUtensorMap* out, const void* ptr, int N, int K, int64_t ldb, int b_mn_major) {
    constexpr int BN = 256, CG = 2;
    if (!b_mn_major) return make_tmap_2d_bf16(out, ptr, K, N, ldb, BK, BN / CG);
    return make_tmap_2d_bf16(out, ptr, N, K, ldb, 64, BK);
}

int gemm_bf16_grouped(const GroupedGemmDesc& g, cudaStream_t stream) {
    constexpr int BN = 256, CG = 2;
    using Cfg = GemmCfg<BN, CG>;
    if (g.num_groups <= 0 || g.num_groups > MAX_GROUPS || g.grp_off == nullptr) return -30;
    if (g.N % 8 != 0) return -31;
    CUtensorMap ta, tb;
    int rc;
    GemmKernelArgs a = {};
    a.grp_off = g.grp_off; a.grp_num = g.num_groups; a.grp_mode = g.mode;
    a.b_maps = g.b_maps; a.d_ptrs = g.d_ptrs;
    a.N = g.N; a.flags = g.flags; a.bias = nullptr; a.H = g.H; a.ldh = g.ldh;
    a.tiles_n = (g.N + BN - 1) / BN;
    a.group_m = 8; a.tail_split = 1;
    if (g.mode == 1) {
        if (g.b_maps == nullptr || g.K % 8 != 0) return -32;
        rc = make_tmap_2d_bf16(&ta, g.A, g.K, g.rows_cap, g.lda, BK, BM);
        if (rc) return rc;
        tb = ta;   // unused: every group brings its own B map
        a.M = (int)g.rows_cap; a.K = g.K; a.D = g.D; a.ldd = g.ldd;
        a.a_mn = 0; a.b_mn = g.b_mn_major;
        a.tiles_m = 0; a.tiles_full = 0;
    } else {
        if (g.d_ptrs == nullptr || g.M % 8 != 0) return -33;
        rc = make_tmap_2d_bf16(&ta, g.A, g.M, g.rows_cap, g.lda, 64, BK);   // dY [rows, M], MN-major A
        if (rc) return rc;
        rc = make_tmap_2d_bf16(&tb, g.B, g.N, g.rows_cap, g.ldb, 64, BK);   // X  [rows, N], MN-major B
        if (rc) return rc;
        a.M = g.M; a.K = 0; a.D = nullptr; a.ldd = g.ldd;
        a.a_mn = 1; a.b_mn = 1;
        a.tiles_m = (g.M + BM * CG - 1) / (BM * CG);
        a.tiles_full = a.tiles_m * a.tiles_n;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_bf16_kernel<BN, false, CG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES) != cudaSuccess)
            return -3;
        attr_set = true;
    }
    int units = num_sms() / CG;
    if (g.mode == 2 && units > g.num_groups * a.tiles_full) units = g.num_groups * a.tiles_full;
    cudaError_t e = launch_pdl(gemm_bf16_kernel<BN, false, CG, true>, dim3(units * CG), dim3(NUM_THREADS), Cfg::SMEM_BYTES, stream,
                               CG, ta, tb, tb, a, CommKernelArgs{});
    if (e != cudaSuccess) {
        fprintf(stderr, "[b200] grouped gemm launch failed: %s\n", cudaGetErrorString(e));
        return -4;
    }
    return 0;
}

int gemm_bf16(const GemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    // pick the narrower tile when the 256-wide grid would leave SMs idle
    const int tiles256 = ((g.M + BM - 1) / BM) * ((g.N + 255) / 256);
    const bool use128 = g.force_bn == 128 || (g.force_bn == 0 && tiles256 < num_sms());
    // default for large problems: 256 x 256 tile on a CTA pair (cta_group::2); force_bn 512 / 384 / 256 / 128 pin a variant
    // (512 = pair x 256 columns, 384 = pair x 192 columns)
    if (g.force_bn == 384) {
        if (g.b_mn_major) return -5;   // MN-major B arrives in 64-column boxes: 96 columns per CTA do not tile
        return launch<192, 2>(g, stream);
    }
    if (g.force_bn == 0 && !use128 && g.max_ctas == 0 && !g.b_mn_major && g_bn192) {
        // Wave quantisation: the pair grid runs ceil(tiles / 74) rounds.  With M = 4096 (one micro-batch) a 256-wide n tile
        // leaves e.g. wqkv (N = 6144) at 384 tiles = 5.19 rounds -> 6; 192-wide tiles make it 512 tiles = 6.92 rounds -> 7
        // of 3/4 the work: 12 % fewer tensor cycles.  Chosen when the modelled cost is at least 3 % lower (the 256-wide
        // tile amortises the A operand and the epilogue better, and its tail wave can be split).
        const int units = num_sms() / 2;
        const int tm = (g.M + 2 * BM - 1) / (2 * BM);
        const int t256 = tm * ((g.N + 255) / 256), t192 = tm * ((g.N + 191) / 192);
        const double r256 = (t256 % units == 0 || t256 < units) ? (double)((t256 + units - 1) / units)
                                                                : t256 / units + 0.6;   // split tail: ~0.6 of a round
        const double c256 = r256 * 256.0, c192 = (double)((t192 + units - 1) / units) * 192.0;
        if (c192 < 0.97 * c256) return launch<192, 2>(g, stream);
    }
    if (g.force_bn == 512 || (g.force_bn == 0 && !use128 && g.max_ctas == 0)) return launch<256, 2>(g, stream);
    return use128 ? launch<128>(g, stream) : launch<256>(g, stream);
}

}  // namespace b200
