// Host API of the bandwidth-bound sm_100a kernels (no torch dependency). All return 0 on success.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

int rmsnorm_fwd(const void* x, const void* res_in, const void* w, void* y, void* res_out, float* rstd, int rows, int H,
                float eps, cudaStream_t s);
int rmsnorm_bwd_blocks(int rows);
int rmsnorm_bwd(const void* dy, const void* res, const void* w, const float* rstd, const void* dres, void* dx,
                float* dw_partial, float* dw_f32, void* dw_bf16, int accumulate, int rows, int H, cudaStream_t s);

// fused dropout + residual-add + LayerNorm (layernorm.cu); partial: [layernorm_bwd_blocks(rows), 2H] fp32, dwdb: [2H] fp32
int layernorm_bwd_blocks(int rows);
int layernorm_fwd(const void* x, const void* res_in, const uint8_t* keep, float drop_scale, const void* w, const void* b,
                  void* y, void* res_out, float* mean, float* rstd, int rows, int H, float eps, cudaStream_t s);
int layernorm_bwd(const void* dy, const void* res, const void* w, const float* mean, const float* rstd, const void* dres,
                  void* dx, float* partial, float* dwdb, int rows, int H, cudaStream_t s);

int rope_inplace(void* x, const int* pos, const float* cos_t, const float* sin_t, int T, int heads, int D,
                 int64_t stride_t, int group, int rot_per_group, int conj, int interleaved, cudaStream_t s);

int swiglu_fwd(const void* gu, void* h, int64_t rows, int64_t F, cudaStream_t s);
int attn_decode(const void* q, const void* kc, const void* vc, void* out, float* work, unsigned int* tickets, int B, int H,
                int Hkv, int D, int seqlen, const int* seqlen_dev, int nsplit, int64_t stride_b, int64_t stride_s, float scale,
                cudaStream_t s);
int gelu_bwd(const void* dh, const void* pre, void* dpre, int64_t n, cudaStream_t s);
int swiglu_bwd(const void* dh, const void* gu, void* dgu, int64_t rows, int64_t F, cudaStream_t s);

int ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int rows, int V, int vocab_start, float* out_max,
           float* out_sum, float* out_sumx, float* out_tgt, cudaStream_t s);
int ce_bwd(void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* gscale, int rows, int V,
           int vocab_start, float smoothing, int total_classes, int ignore_index, cudaStream_t s);

int adamw_step(float* p, float* m, float* v, const void* g, int g_is_bf16, void* p_lp, int64_t n, float lr, float beta1,
               float beta2, float eps, float wd, float bc1, float bc2, const float* scalars, cudaStream_t s);
int sumsq(const void* g, int is_bf16, int64_t n, float* out, cudaStream_t s);
int clip_scalars(const float* sumsq_in, float* scalars, float loss_scale, float clip, cudaStream_t s);

}  // namespace b200
