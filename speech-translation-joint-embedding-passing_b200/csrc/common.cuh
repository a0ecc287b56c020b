// Shared helpers for libb200st (sm_100a).  Host-side error plumbing + device-side dtype/reduction utils.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200st.h"

namespace b200st {

int set_error(const char* fmt, ...);       // formats into the thread-local error slot, returns -1
void count_launch(int n = 1);

#define B200ST_LAUNCH_CHECK(name)                                                        \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) return b200st::set_error("%s: %s", name, cudaGetErrorString(e__)); \
    b200st::count_launch();                                                              \
  } while (0)

#define B200ST_CUDA(call)                                                                \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) return b200st::set_error("%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

#define B200ST_DISPATCH(dtype, T, ...)                                                   \
  do {                                                                                   \
    if ((dtype) == B200ST_F32) { using T = float; __VA_ARGS__; }                         \
    else if ((dtype) == B200ST_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }           \
    else return b200st::set_error("unsupported dtype %d", (int)(dtype));                 \
  } while (0)

// four consecutive elements (16-byte aligned for fp32, 8-byte for bf16) <-> fp32 registers
__device__ __forceinline__ void load4(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float* v) {
  const uint2 a = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(a.x << 16); v[1] = __uint_as_float(a.x & 0xffff0000u);
  v[2] = __uint_as_float(a.y << 16); v[3] = __uint_as_float(a.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float* v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// Hot-loop kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the grid may be scheduled
// while its predecessor in the stream is still draining, so launch latency and per-CTA prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlap the predecessor's tail.  `pdl_wait()` (griddepcontrol.wait) blocks until
// every prerequisite grid has completed and its memory is visible; it is placed before the first access to any
// buffer another kernel may have written or may still be reading, so the data semantics stay strictly serial.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// PDL + a thread-block cluster of `cz` CTAs along grid.z (cluster split-K kernels)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster_z(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                               cudaStream_t st, unsigned cz, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 1;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = cz;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 2^v on the SFU (MUFU.EX2, relative error 2^-22): for bf16-output kernels, exp(x - m) = ex2_approx(x * log2e - m * log2e) is one
// FFMA + one MUFU where expf() is ~10 instructions
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide reductions through a 32-float shared scratch; every thread gets the result.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  return warp_max(r);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b200st
