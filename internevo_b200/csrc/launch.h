// Kernel launch with programmatic dependent launch (PDL, sm_90+).
//
// A training step is ~16,000 back-to-back launches of our own kernels on one stream.  With plain stream order the grid of
// kernel B is dispatched only after kernel A has drained and its memory has been flushed; the dispatch latency plus B's
// prologue (mbarrier init, TMEM allocation, descriptor prefetch) sit on the critical path of every boundary.
// With PDL
//   * every kernel calls `griddep_launch_dependents()` first: as soon as all CTAs of A are resident, the CTAs of B may be
//     scheduled onto SMs that A's tail wave has left,
//   * B runs its prologue there and then blocks in `griddep_wait()` until A has completed and its writes are visible.
// Correctness rule used throughout: a kernel touches global memory only after `griddep_wait()` (executed by every thread).
// Kernels launched without the attribute (torch's own, or B200_PDL=0) see both instructions as no-ops.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace b200 {

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("B200_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

// cluster_x > 1 launches thread-block clusters of that size along x (cta_group::2 GEMM)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace b200
