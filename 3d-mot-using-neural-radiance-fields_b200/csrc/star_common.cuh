// Shared device helpers for the STaR B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/star_b200.h"

#define STAR_FULL_MASK 0xffffffffu
#define STAR_EPS_F32 1.1920928955078125e-07f  // torch.finfo(float32).eps  (utils/constants.py:3)
#define STAR_MAX_V 8                          // max dynamic objects handled by the compositing kernels

extern int g_star_last_cuda_error;
// device pointer of the watchdog word of a tensor-core kernel family (c_api.cu); NULL if mapped host memory is unavailable
enum { STAR_WD_FWD = 0, STAR_WD_DX = 1, STAR_WD_DW = 2, STAR_WD_DX2 = 3, STAR_WD_MIP_FWD = 4, STAR_WD_MIP_DX = 5 };
int* star_watchdog_dev(int family);

extern int g_star_sync_launches;   // STAR_B200_SYNC_LAUNCHES=1 (diagnostics): every entry point waits for its kernels, so that
                                   // an asynchronous fault is reported by the call that caused it
static inline int star_check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && g_star_sync_launches) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    g_star_last_cuda_error = (int)e;
    return STAR_E_CUDA;
  }
  return STAR_OK;
}

// Where a kernel finds the sample positions: either a materialised pts [M,3] array (the tensor that crosses the
// reference API between sample_pts and render_*, models/rendering__.py:108-110) or, when pts == NULL, the ray and its
// depths -- p = rays_o[r] + rays_d[r] * z_vals[r, s] with the reference's two roundings (no FMA contraction), so both
// forms are bit-identical and the fused render path never writes or reads 12 B/sample of positions.
struct StarPtsSrc {
  const float* pts;
  const float* rays_o;
  const float* rays_d;
  const float* z_vals;     // [R, S]
};
__device__ __forceinline__ void star_load_pt(const StarPtsSrc& s, int64_t gi, int64_t r, float& px, float& py, float& pz) {
  if (s.pts != nullptr) {
    px = s.pts[gi * 3 + 0]; py = s.pts[gi * 3 + 1]; pz = s.pts[gi * 3 + 2];
  } else {
    const float z = s.z_vals[gi];
    px = __fadd_rn(s.rays_o[r * 3 + 0], __fmul_rn(s.rays_d[r * 3 + 0], z));
    py = __fadd_rn(s.rays_o[r * 3 + 1], __fmul_rn(s.rays_d[r * 3 + 1], z));
    pz = __fadd_rn(s.rays_o[r * 3 + 2], __fmul_rn(s.rays_d[r * 3 + 2], z));
  }
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
// Sums four values across the warp with 6 shuffles instead of 20: each butterfly step halves the number of values
// a lane carries.  Returns sum(a) in lanes 0-7, sum(b) in lanes 8-15, sum(c) in lanes 16-23, sum(d) in lanes 24-31.
__device__ __forceinline__ float warp_sum4(float a, float b, float c, float d, int lane) {
  const bool h16 = lane & 16;
  const float x = (h16 ? c : a) + __shfl_xor_sync(STAR_FULL_MASK, h16 ? a : c, 16);
  const float y = (h16 ? d : b) + __shfl_xor_sync(STAR_FULL_MASK, h16 ? b : d, 16);
  const bool h8 = lane & 8;
  float v = (h8 ? y : x) + __shfl_xor_sync(STAR_FULL_MASK, h8 ? x : y, 8);
  v += __shfl_xor_sync(STAR_FULL_MASK, v, 4);
  v += __shfl_xor_sync(STAR_FULL_MASK, v, 2);
  v += __shfl_xor_sync(STAR_FULL_MASK, v, 1);
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

// Transcendentals of the compositing kernels.  These kernels are graded against the HBM roofline, and with libm-grade
// expf/log1pf/IEEE division they are issue-bound at ~1/3 of it, so they use the SFU approximations (ex2.approx /
// lg2.approx / rcp.approx: relative error <= ~2^-21 for the arguments that occur) -- two orders of magnitude inside
// the 1e-4 absolute parity bar.  Nothing here feeds an index decision (sample_pdf has its own defined arithmetic).
// The .ftz forms: without them every ex2 / lg2 / rcp carries a denormal fix-up (FSETP + two or three FMUL: 40 % of the
// multi-field kernel's instructions, profiles/r2k_composite_multi.md).  Flushing only changes values below 2^-126: an
// exp() that underflows gives 0 instead of a denormal (alpha = 1 - exp is the same float either way), a density below
// 1e-38 counts as 0.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fexp(float x) { return ex2_ftz(x * 1.4426950408889634f); }
__device__ __forceinline__ float flog(float x) { return lg2_ftz(x) * 0.6931471805599453f; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * rcp_ftz(b); }
// log(1 - c) for c in [eps, 1-eps]; lg2.approx has an ABSOLUTE error of ~2^-22 near 1, so small c takes the series
__device__ __forceinline__ float flog1m(float c) { return c < 1e-3f ? -c * (1.f + 0.5f * c) : flog(1.f - c); }
// torch.nn.functional.softplus (beta=1, threshold=20).  Below -8 the series e - e^2/2 keeps the RELATIVE accuracy of
// tiny densities (they are multiplied by far_dist = 1e10 at the last sample, rendering__.py:321).
__device__ __forceinline__ float softplus_f(float x) {
  if (x > 20.f) return x;
  const float e = fexp(x);
  return x < -8.f ? e * (1.f - 0.5f * e) : flog(1.f + e);
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.f + fexp(-x)); }
// d softplus / dx with torch's threshold semantics
__device__ __forceinline__ float softplus_grad_f(float x) { return x > 20.f ? 1.f : sigmoid_f(x); }

// torch.max(a, b) element-wise propagates NaN (fmaxf drops it): disp = 1 / max(1e-10, depth / acc) of a ray whose weights
// are all exactly 0 is 0 / 0 = NaN in the reference (models/rendering__.py:353-357) and must be NaN here
__device__ __forceinline__ float max_nan_f(float a, float b) { return (b != b) ? b : fmaxf(a, b); }
