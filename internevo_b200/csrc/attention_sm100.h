// Host API of the sm_100a flash-attention kernels (varlen, causal, GQA; bf16 in / fp32 softmax).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

struct AttnDesc {
    const void* q = nullptr;  // [T, H, D]   strides in elements (q_stride_t, q_stride_h, 1)
    const void* k = nullptr;  // [T, Hkv, D]
    const void* v = nullptr;  // [T, Hkv, D]
    void* o = nullptr;        // [T, H, D] contiguous
    float* lse = nullptr;     // [H, T] natural-log LSE (scaled scores)
    int T = 0, H = 0, Hkv = 0, D = 0;
    int64_t q_stride_t = 0, q_stride_h = 0, k_stride_t = 0, k_stride_h = 0, v_stride_t = 0, v_stride_h = 0;
    int64_t q_stride_g = 0;  // stride between kv groups of q heads (0: q_stride_h * H/Hkv); head (g,j) = g*q_stride_g + j*q_stride_h
    const int* cu_seqlens = nullptr;  // [num_seqs + 1]
    int num_seqs = 0, max_seqlen = 0;
    float scale = 1.f;
    int causal = 1;
    // sequence parallel (sp_world > 1): q / o / lse hold this rank's T token rows of the packed stream, cu_seqlens are
    // GLOBAL, and K / V of rank p are read in place through k_peers[p] / v_peers[p] (host arrays of peer-mapped pointers
    // into symmetric buffers with the strides given above); T % 128 == 0
    int sp_rank = 0, sp_world = 1;
    const void* const* k_peers = nullptr;
    const void* const* v_peers = nullptr;
};

struct AttnBwdDesc {
    AttnDesc f;
    const void* dout = nullptr;  // [T, H, D] contiguous
    void* dq = nullptr;          // [T, H, D]
    void* dk = nullptr;          // [T, Hkv, D]
    void* dv = nullptr;          // [T, Hkv, D]
    int64_t dq_stride_g = 0;  // as q_stride_g
    int64_t dq_stride_t = 0, dq_stride_h = 0, dk_stride_t = 0, dk_stride_h = 0, dv_stride_t = 0, dv_stride_h = 0;
    float* delta = nullptr;   // [H, T] scratch: rowsum(dO * O)
    float* dq_acc = nullptr;  // [T, H, D] fp32 scratch (zero-initialised by the caller)
    // sequence parallel (f.sp_world > 1): per-peer pointers (host arrays) of the SYMMETRIC copies of q, dout, the fp32 dQ
    // accumulator and the [2, H, T] stats buffer (`delta` above is this rank's); `phase` selects 1 = stats, 2 = main kernel,
    // 3 = dQ conversion (0 = all three, single rank only)
    int phase = 0;
    const void* const* q_peers = nullptr;
    const void* const* dout_peers = nullptr;
    void* const* dq_acc_peers = nullptr;
    const void* const* delta_peers = nullptr;
};

int attn_fwd(const AttnDesc& d, cudaStream_t s);
int attn_bwd(const AttnBwdDesc& d, cudaStream_t s);

}  // namespace b200
