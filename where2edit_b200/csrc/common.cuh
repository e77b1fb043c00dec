// Shared helpers for libw2e (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/w2e.h"

namespace w2e {

char* last_error_buffer();  // thread-local, defined in api.cu
int set_error(int code, const char* fmt, ...);

#define W2E_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return ::w2e::set_error(W2E_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define W2E_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::w2e::set_error(W2E_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                              __FILE__, __LINE__);                                             \
  } while (0)

#define W2E_LAUNCH_OK() W2E_CUDA_OK(cudaGetLastError())

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count();        // of the CURRENT device (cached per device ordinal), api.cu
int current_device();  // cudaGetDevice; -1 if it fails

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember it per device ordinal
struct PerDeviceOnce {
  bool flag[64] = {};
  bool done() const { const int d = current_device(); return d >= 0 && d < 64 && flag[d]; }
  void mark() { const int d = current_device(); if (d >= 0 && d < 64) flag[d] = true; }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float lrelu_gain(float v, float slope, float gain) {
  return (v > 0.f ? v : v * slope) * gain;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` needs 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// fp64 variants for the long, cancellation-heavy reductions of the backward (sums over up to 2^20 pixels of signed
// gradient terms: condition number ~ sqrt(N); the kernels are memory-bound, the extra DADD per element is free)
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum; result valid in thread 0.  `red` needs 32 doubles of shared memory.
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum_d(v);
  __syncthreads();
  return v;
}

}  // namespace w2e
