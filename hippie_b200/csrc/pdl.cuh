// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may have its CTAs scheduled while the
// preceding kernel of the stream is still draining; it must call pdl_wait() before its first access to global memory
// (griddepcontrol.wait returns once every prerequisite grid has completed and its writes are visible).  pdl_trigger()
// at the top of a kernel lets ITS successor be scheduled as soon as all of its own CTAs have started.  Launch latency,
// barrier / TMEM set-up and tensor-map fetches of kernel N+1 thereby overlap the tail of kernel N -- the train step is a
// chain of ~450 short dependent kernels, so this is where the time goes.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace hp {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("HIPPIE_B200_PDL");
    return !v || atoi(v) != 0;
  }();
  return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace hp
