// Programmatic Dependent Launch (PDL) helpers.
//
// A local-SGD step is ~200 small kernels in a dependency chain; at ~1 ms per step the per-node
// launch latency and prologue (barrier init, TMEM alloc, descriptor prefetch, smem tables) is a
// large share of the time.  Every kernel of this library
//   1. calls griddep_launch_dependents() first thing  -> the NEXT kernel in the stream / captured
//      graph may become resident and run its prologue while this one is still executing, and
//   2. calls griddep_wait() before its first access to global memory -> it only consumes the
//      previous kernel's results once that grid has completed and flushed.
// Launches go through launch_pdl(), which sets cudaLaunchAttributeProgrammaticStreamSerialization
// (captured into CUDA graphs as programmatic dependency edges).  BATON_PDL=0 disables it.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace b200 {

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("BATON_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

}  // namespace b200
