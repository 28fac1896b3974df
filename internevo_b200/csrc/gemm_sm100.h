// Host API of the sm_100a tcgen05 GEMM family (no torch dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum GemmFlags : int {
    GEMM_OUT_F32 = 1,      // D is fp32 (default bf16)
    GEMM_ACCUMULATE = 2,   // D += A*B^T (reads the previous D)
    GEMM_SWIGLU = 4,       // columns of D are interleaved (gate, up); also write H[:, j] = silu(gate_j) * up_j
    GEMM_SKIP_D = 8,       // with GEMM_SWIGLU: do not materialise D
    GEMM_GELU = 16,        // also write H[:, j] = gelu_tanh(D[:, j]) (bias included), H is bf16 [M, N]
};

struct GemmDesc {
    int M = 0, N = 0, K = 0;
    const void* A = nullptr;  // bf16; K-major: [M, K] with row stride lda; MN-major: [K, M] with row stride lda
    int64_t lda = 0;
    int a_mn_major = 0;
    const void* B = nullptr;  // bf16; K-major: [N, K] with row stride ldb; MN-major: [K, N] with row stride ldb
    int64_t ldb = 0;
    int b_mn_major = 0;
    void* D = nullptr;  // [M, N] row stride ldd, bf16 or fp32
    int64_t ldd = 0;
    const void* bias = nullptr;  // bf16 [N] or null
    void* H = nullptr;           // bf16 [M, N/2] (GEMM_SWIGLU) or [M, N] (GEMM_GELU)
    int64_t ldh = 0;
    int flags = 0;
    int force_bn = 0;   // 0 = auto, 128 or 256
    int max_ctas = 0;   // 0 = all SMs; otherwise cap the persistent grid (leave SMs for a concurrent comm kernel)
};

int gemm_bf16(const GemmDesc& g, cudaStream_t stream);
// Enable / disable splitting the tiles of the last partial wave into column slices (default on).
void set_gemm_tail_split(int on);
// Raster group height in tile rows (0 = automatic: whole M when it fits 32 tile rows, else 8).
void set_gemm_group_m(int g);

// ---- GEMM fused with its tensor-parallel collective (one launch, peer memory over NVLink) -------------------------
enum GemmCommMode : int {
    GEMM_COMM_NONE = 0,
    GEMM_COMM_ALL_GATHER = 1,      // A = all-gather of per-rank row shards, pushed by the epilogue warps of every CTA first
    GEMM_COMM_REDUCE_SCATTER = 2,  // epilogue pushes partial tiles to their owner; the owner's epilogue reduces
    GEMM_COMM_ALL_REDUCE = 3,      // as above, reduced rows are pushed into every peer's output
    GEMM_COMM_GATHER_B = 4,        // B = all-gather of per-rank weight shards [N / W, K] (weight / ISP parallelism): pushed
                                   // like the activations above; the GEMM walks the shards own-first (n tiles when B is
                                   // K-major, k blocks when B is MN-major) and reads its own shard in place
};

struct GemmCommArgs {
    int mode = GEMM_COMM_NONE;
    void* const* peer_ptrs = nullptr;      // AG: per-rank gathered A [M, K]; RS/AR: per-rank staging [world, m_local, N]
    uint32_t* const* flags_ptrs = nullptr; // per-rank flag arrays (uint32): AG 16 words per 128-row block, RS/AR blocks * tiles_n
    void* const* out_ptrs = nullptr;       // AR: per-rank final output [M, N]
    int rank = 0, world = 1;
    uint32_t epoch = 0;
    int64_t m_local = 0;                   // rows per rank (AG: contributed, RS: owned)
    void* out_local = nullptr;             // AG / weight gather: this rank's gathered buffer; RS: reduced rows [m_local, N]
    int64_t ld_out = 0;                    // RS / AR: row stride of the output
    const void* x_local = nullptr;         // AG: this rank's shard [m_local, K] (contiguous)
    int comm_ctas = 0;                     // unused (the all-gather push runs on the epilogue warps of every CTA)
    uint32_t* const* done_ptrs = nullptr;  // AR: per-rank `world` completion words (end-of-kernel handshake)
    uint32_t* done_counter = nullptr;      // AR: local counter of finished CTAs
    float out_scale = 1.f;                 // RS: scale of the reduced rows (GEMM_ACCUMULATE in g.flags adds the previous out_local)
};

int gemm_bf16_comm(const GemmDesc& g, const GemmCommArgs& c, cudaStream_t stream);

// ---- grouped GEMM: `num_groups` problems whose row ranges live in DEVICE memory (MoE experts, no host sync) --------------
// mode 1 (forward / dgrad):  D[rows_g, N] = A[rows_g, K] * B_g^T   for the rows grp_off[g] .. grp_off[g + 1] of A / D;
//                            B_g comes through its own tensor map (array in global memory, see grouped_b_maps).
// mode 2 (wgrad):            D_g[M, N] (+)= A[rows_g, M]^T * B[rows_g, N]   (contraction over the group's rows).
// Row offsets must be multiples of 128 and the padding rows of A zero (mode 2) / ignorable (mode 1: D rows are written).
struct GroupedGemmDesc {
    int mode = 1;
    int num_groups = 0;
    const int* grp_off = nullptr;            // device, [num_groups + 1]
    int64_t rows_cap = 0;                    // rows of the packed buffers (tensor-map extent)
    int M = 0, N = 0, K = 0;                 // mode 1: N, K (M unused); mode 2: M, N (K unused)
    const void* A = nullptr; int64_t lda = 0;
    const void* B = nullptr; int64_t ldb = 0;   // mode 2: the packed second operand
    int b_mn_major = 0;                      // mode 1: layout of every B_g
    const CUtensorMap* b_maps = nullptr;     // mode 1: device array of num_groups maps
    void* D = nullptr; int64_t ldd = 0;      // mode 1
    void* const* d_ptrs = nullptr;           // mode 2: device array of num_groups outputs (row stride ldd)
    void* H = nullptr; int64_t ldh = 0;      // mode 1 + GEMM_SWIGLU
    int flags = 0;
};
int gemm_bf16_grouped(const GroupedGemmDesc& g, cudaStream_t stream);
// host-side tensor map of one group's B operand for mode 1 (K-major [N, K] or MN-major [K, N]), 128 bytes into `out`
int grouped_b_map(CUtensorMap* out, const void* ptr, int N, int K, int64_t ldb, int b_mn_major);

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer);

int make_tmap_2d_f32_noswizzle(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                               uint32_t box_inner, uint32_t box_outer);

}  // namespace b200
