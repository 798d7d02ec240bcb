// Shared helpers for libvargp_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vargp_sm100.h"

namespace vargp {

extern int64_t g_launches;   // counted on the host at every kernel launch (bench.py: gpu_launches)

#define VARGP_ERR_ARG (-1)
#define VARGP_ERR_UNSUPPORTED (-2)
#define VARGP_ERR_NOT_INIT (-3)

extern int g_pdl;             // programmatic dependent launch between the library's kernels: 0 off, 1 every kernel,
                              // 2 only kernels with a small footprint (VARGP_PDL; see launch_k)

// Every kernel of the library starts with pdl_enter(): `launch_dependents` lets the NEXT kernel of the stream (or
// graph) be scheduled onto SMs as they drain instead of after this grid has fully retired, `wait` blocks until
// the PREVIOUS grid has completed and its writes are visible.  Both are no-ops for launches without the
// programmatic-serialization attribute.  The step is ~80 short dependent kernels, so the per-boundary launch
// latency is a measurable share of it.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}

// A dependent kernel launched with the programmatic-serialization attribute becomes RESIDENT while its predecessor still
// runs and parks in griddepcontrol.wait.  For light kernels that hides the launch latency; a tensor-core GEMM CTA parks
// with ~200 KB of shared memory and its TMEM allocation, i.e. it takes a whole SM away from the grid it is waiting for
// (measured at the Split-MNIST shape: 974 steps/s with the attribute on every launch, 1015 without it).  Mode 2 keeps
// the attribute for launches with at most 48 KB of dynamic shared memory.
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  unsigned n = 0;
  if (g_pdl == 1 || (g_pdl == 2 && smem <= 48 * 1024)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int launch_status() {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)cudaGetLastError();
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0 (and broadcast to all when `all` is set). scratch: >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

}  // namespace vargp
