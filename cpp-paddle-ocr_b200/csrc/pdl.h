// Programmatic dependent launch (PDL) for the layer kernels.
//
// The three networks are chains of 57-77 short kernels; at the shapes this path runs, most of them are bound by the
// per-launch floor (grid drain -> next grid's launch -> its prologue), not by bandwidth.  Every kernel of the path
//   * calls pdl_trigger() first: once all CTAs of a grid have started, the NEXT kernel of the stream may be scheduled
//     onto free SM slots and run its prologue (barrier init, TMEM allocation, tensor-map prefetch, filter / bias
//     loads -- nothing a previous kernel produces) while this grid is still computing;
//   * calls pdl_wait() before it touches any activation (read OR write: the arena re-uses buffers, so an output may
//     alias something the previous kernel still reads): griddepcontrol.wait returns when the previous grid has
//     completed and its memory operations are visible.
// launch_k() launches with cudaLaunchAttributeProgrammaticStreamSerialization (also inside stream capture: the edges
// become programmatic dependencies of the CUDA graph).  Without the attribute (B200OCR_PDL=0) both instructions are
// no-ops and the stream is fully serialised as before.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace b200ocr {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef B200OCR_PDL_EARLY
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
// measured (profiles/r02_notes.md): with three workers per GPU an early trigger costs 3 % -- the dependents that sit in
// griddepcontrol.wait hold SM slots the other streams' runnable CTAs would have used.  Default: implicit trigger at exit.
__device__ __forceinline__ void pdl_trigger() {}
#endif

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("B200OCR_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// The same for a kernel that runs as thread-block clusters of `cluster_x` CTAs along x (grid.x is a multiple of it).
template <class... KArgs, class... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(cluster_x);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace b200ocr
