// Host API of the peer-memory (NVLink / NVSwitch) kernels: barrier over symmetric flags, fused
// reduce-scatter + AdamW + all-gather for Hybrid-ZeRO, GEMM->reduce-scatter and all-gather->GEMM for tensor parallel.
// Pointer tables (`*_ptrs`) are device arrays of `world` peer-mapped base pointers (see parallel/symm.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_sm100.h"

namespace b200 {

int symm_barrier(uint32_t* const* flags_ptrs, int rank, int world, uint32_t epoch, cudaStream_t s);
// micro-benchmark of the SM-driven peer copy loop (unroll = 16-byte loads in flight per thread: 4, 8 or 16)
int peer_copy_bench(const void* src, void* dst, int64_t bytes, int64_t piece_bytes, int unroll, int ctas, cudaStream_t s);

struct RsAdamDesc {
    void* const* grad_ptrs = nullptr;    // per-rank bf16 gradient arena (same layout on every rank)
    void* const* param_ptrs = nullptr;   // per-rank bf16 parameter arena
    uint32_t* const* flags_ptrs = nullptr;
    // optional NVSwitch multicast mappings of the same arenas (nullptr -> unicast peer loads / stores) + local pointers
    const void* grad_mc = nullptr;
    void* param_mc = nullptr;
    void* grad_local = nullptr;
    int rank = 0, world = 1;
    uint32_t epoch = 0;
    int64_t shard_off = 0, shard_n = 0;  // this rank's owned slice [shard_off, shard_off + shard_n) of the arena
    float *p = nullptr, *m = nullptr, *v = nullptr;  // fp32 master / moments of the owned slice
    float* scalars = nullptr;  // [0] grad multiplier, [1] overflow flag, [2] norm, [3] local sum of squares
    double lr = 0, beta1 = 0.9, beta2 = 0.95, eps = 1e-8, wd = 0, bc1 = 1, bc2 = 1, grad_div = 1;
    int phase = 0;  // 0: reduce owned slice (peer loads) -> fp32 grad in `m`-sized scratch + sumsq ; 1: adam + multicast params
};
int reduce_scatter_adam(const RsAdamDesc& d, cudaStream_t s);

struct GemmCommDesc {
    GemmDesc g;
    void* const* peer_ptrs = nullptr;   // AG: per-rank gathered buffers; RS/AR: per-rank staging [world, M/world, N]
    void* const* out_ptrs = nullptr;    // AR: per-rank output buffers
    uint32_t* const* flags_ptrs = nullptr;
    int rank = 0, world = 1;
    uint32_t epoch = 0;
    int mode = 0;          // gemm_rs: 0 = reduce-scatter rows, 1 = all-reduce
    int64_t m_local = 0;   // ag_gemm: rows contributed by each rank
    void* out_local = nullptr;
    int64_t ld_out = 0;
    const void* x_local = nullptr;      // ag_gemm: this rank's shard
    int comm_ctas = 0;
    uint32_t* const* done_ptrs = nullptr;   // all-reduce: per-rank completion words
    uint32_t* done_counter = nullptr;       // all-reduce: local finished-CTA counter
    float out_scale = 1.f;                  // reduce-scatter: scale of the reduced rows
};
int gemm_reduce_scatter(const GemmCommDesc& d, cudaStream_t s);
int allgather_gemm(const GemmCommDesc& d, cudaStream_t s);
// weight parallelism (ISP): B = all-gather of per-rank weight shards, consumed shard by shard by the same launch
int gather_weight_gemm(const GemmCommDesc& d, cudaStream_t s);

// ---- MoE expert-parallel dispatch / combine over peer memory (moe_comm.cu)
// all ranks push `nwords` words into slot `rank` of every peer's table and rendezvous (one launch)
int symm_allgather_small(uint32_t* const* buf_ptrs, const uint32_t* src, int nwords, uint32_t* const* flags_ptrs, int rank,
                         int world, uint32_t epoch, cudaStream_t s);

struct MoeCommDesc {
    const void* x = nullptr;     // scatter: local bf16 rows [n_slots / k, H] (row of slot s is s / k)
    int64_t ldx = 0;
    void* out = nullptr;         // combine: local bf16 output [n_slots / k, H]
    int64_t ldo = 0;
    const int* slot_rank = nullptr;  // [n_slots] owner GPU of the slot's expert row
    const int* slot_row = nullptr;   // [n_slots] row inside the owner's slab (< 0: dropped slot)
    const float* scale = nullptr;    // [n_slots] optional per-slot weight
    void* const* x_ptrs = nullptr;   // scatter destination: per-rank bf16 [rows, H] symmetric slabs
    void* const* y_ptrs = nullptr;   // combine source (and, for scatter, rows to dot with for d(gate weight))
    float* dw = nullptr;             // scatter: optional [n_slots] <x[t], y_row>
    int n_slots = 0, k = 1, H = 0;
};
int moe_scatter_rows(const MoeCommDesc& d, cudaStream_t s);
int moe_gather_combine(const MoeCommDesc& d, cudaStream_t s);

}  // namespace b200
