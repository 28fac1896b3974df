#include "comm_kernels.h"
namespace b200 {
int symm_barrier(uint32_t* const*, int, int, uint32_t, cudaStream_t) { return -100; }
int reduce_scatter_adam(const RsAdamDesc&, cudaStream_t) { return -100; }
int gemm_reduce_scatter(const GemmCommDesc&, cudaStream_t) { return -100; }
int allgather_gemm(const GemmCommDesc&, cudaStream_t) { return -100; }
}  // namespace b200
