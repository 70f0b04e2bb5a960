// launch.cuh -- programmatic dependent launch (PDL) for the kernels of the sampler iteration: on for small batches only (measured).
//
// One ADPM2 iteration is ~500 dependent kernels of 10-200 us each.  A kernel launched through launch_k() can carry
// cudaLaunchAttributeProgrammaticStreamSerialization (captured into the CUDA graph as a programmatic dependency edge): its CTAs may
// become resident while the previous grid is still finishing, run their set-up (barrier init, TMEM allocation, descriptor prefetch),
// and block in pdl_wait() (griddepcontrol.wait) until the previous grid has completed and its writes are visible.  Rules that keep
// this exactly equivalent to stream order:
//   * every kernel launched through launch_k() executes pdl_wait() in all threads before its first global memory access;
//   * pdl_trigger() (griddepcontrol.launch_dependents) comes after pdl_wait(), so at most two grids are ever in flight and
//     completion stays transitive (grid k + 1 complete => grid k complete);
//   * without the launch attribute both instructions are no-ops.
// Measured on B200 (cfg2, B = 4096, tf32, 2-step bench; profiles/README.md): off 2473 samples/s; persistent tcgen05 kernels only
// (MDT_PDL=1) 2438; element-wise / normalisation kernels only (MDT_PDL=2) 2456; both (MDT_PDL=3) 2397; both without the explicit
// trigger 2455.  The persistent kernels fill an SM (220 KB of shared memory, all 512 TMEM columns), so the next grid's CTAs cannot
// co-reside and there is no set-up to overlap; what remains is the cost of the programmatic edges.  At B = 4 (cfg1: one or two CTAs
// per grid) it is the other way round: 259.4 vs 274.9 ms per 64-step sample() with bit 0, no change with bit 1.  Hence the automatic default.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <utility>

namespace mdt {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef MDT_PDL_NO_TRIGGER
__device__ __forceinline__ void pdl_trigger() {}
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

// MDT_PDL bit 0: the persistent tcgen05 kernels, bit 1: the element-wise / normalisation kernels.  Unset = automatic: the plan executor
// turns bit 0 on for calls whose every grid is smaller than the GPU (small batches: the next grid's CTAs find free SMs, and overlapping
// its launch + set-up with the running kernel is worth -5.6 % of the B = 4 latency), off otherwise (measured slower, see above).
inline int& pdl_env() {
  static int m = [] { const char* e = getenv("MDT_PDL"); return e ? atoi(e) : -1; }();
  return m;
}
inline int& pdl_auto() { static int a = 0; return a; }      // written by the plan executor before it launches a program
inline int pdl_mask() { const int m = pdl_env(); return m >= 0 ? m : pdl_auto(); }

template <int CLASS, typename... KArgs, typename... Args>
inline cudaError_t launch_k_class(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl_mask() & CLASS) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_k_class<1>(kernel, grid, block, smem, s, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_light(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_k_class<2>(kernel, grid, block, smem, s, std::forward<Args>(args)...);
}

}  // namespace mdt
