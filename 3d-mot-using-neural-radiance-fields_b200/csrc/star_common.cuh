// Shared device helpers for the STaR B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/star_b200.h"

#define STAR_FULL_MASK 0xffffffffu
#define STAR_EPS_F32 1.1920928955078125e-07f  // torch.finfo(float32).eps  (utils/constants.py:3)
#define STAR_MAX_V 8                          // max dynamic objects handled by the compositing kernels

extern int g_star_last_cuda_error;

static inline int star_check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_star_last_cuda_error = (int)e;
    return STAR_E_CUDA;
  }
  return STAR_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(STAR_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(STAR_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(STAR_FULL_MASK, v, o));
  return v;
}
// inclusive prefix product across the warp
__device__ __forceinline__ float warp_scan_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(STAR_FULL_MASK, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
// inclusive SUFFIX sum across the warp (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_rscan_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_down_sync(STAR_FULL_MASK, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}
// inclusive prefix sum (double) across the warp
__device__ __forceinline__ double warp_scan_sum_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = __shfl_up_sync(STAR_FULL_MASK, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
// d softplus / dx with torch's threshold semantics
__device__ __forceinline__ float softplus_grad_f(float x) { return x > 20.f ? 1.f : sigmoid_f(x); }
