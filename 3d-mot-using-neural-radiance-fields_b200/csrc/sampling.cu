// Ray sampling kernels: stratified samples (a1), positional encoding (a3, stand-alone form),
// inverse-CDF sampling (a9) and the coarse->fine hierarchical step (a10).
// All arithmetic that feeds index decisions is written with explicit round-to-nearest intrinsics
// (__fadd_rn / __fmul_rn / __fdiv_rn) so that nvcc cannot contract it into FMAs: every reference
// op (one eager ATen kernel each) is rounded separately, and so is every op here.
#include "ray_device.cuh"

// ------------------------------------------------------------------------------------------ a1
// models/rendering__.py:87-110.  One thread per (ray, sample); pts written as 3 coalesced floats.
__global__ void sample_pts_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                  const float* __restrict__ t_vals, const float* __restrict__ t_rand,
                                  float near_, float far_, int R, int Nc, int lindisp,
                                  float* __restrict__ pts, float* __restrict__ z_vals) {
  const int64_t total = (int64_t)R * Nc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Nc), s = (int)(i % Nc);
    auto zval = [&](int k) -> float {
      const float t = t_vals[k];
      if (!lindisp) return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.f, t)), __fmul_rn(far_, t));
      const float a = __fmul_rn(__fdiv_rn(1.f, near_), __fsub_rn(1.f, t));
      const float b = __fmul_rn(__fdiv_rn(1.f, far_), t);
      return __fdiv_rn(1.f, __fadd_rn(a, b));
    };
    float z = zval(s);
    if (t_rand != nullptr) {  // :99-106 stratified jitter with injected noise
      const float lo = (s == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, zval(s - 1)));
      const float hi = (s == Nc - 1) ? z : __fmul_rn(0.5f, __fadd_rn(zval(s + 1), z));
      z = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[i]));
    }
    z_vals[i] = z;
    if (pts != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        pts[i * 3 + c] = __fadd_rn(rays_o[r * 3 + c], __fmul_rn(rays_d[r * 3 + c], z));
    }
  }
}

// Nc % 4 == 0: one thread per four consecutive samples of a ray (one quad per thread, the grid covers all quads), 32-bit
// index arithmetic, same per-element arithmetic as above.  z leaves as one 16-byte store per thread.  The 12 position
// floats of a quad are 48 contiguous bytes: stored straight from the thread, a warp's three store instructions each touch
// 48 sectors half-way (16 of every 48 bytes) and the kernel ran at 0.53 of the HBM peak; staged through shared memory every
// warp store covers 512 contiguous bytes: 0.82 (profiles/r2s_ab_sample_pts.txt).
#define SP_THREADS 256
__global__ void __launch_bounds__(SP_THREADS) sample_pts_x4_kernel(const float* __restrict__ rays_o,
                                     const float* __restrict__ rays_d,
                                     const float* __restrict__ t_vals, const float* __restrict__ t_rand,
                                     float near_, float far_, int R, int Nc, int lindisp,
                                     float* __restrict__ pts, float* __restrict__ z_vals) {
  __shared__ float4 stage[3 * SP_THREADS];
  const uint32_t Q = (uint32_t)Nc >> 2;
  const uint32_t total = (uint32_t)R * Q;
  auto zval = [&](int k) -> float {
    const float t = t_vals[k];
    if (!lindisp) return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.f, t)), __fmul_rn(far_, t));
    const float a = __fmul_rn(__fdiv_rn(1.f, near_), __fsub_rn(1.f, t));
    const float b = __fmul_rn(__fdiv_rn(1.f, far_), t);
    return __fdiv_rn(1.f, __fadd_rn(a, b));
  };
  const uint32_t q0 = blockIdx.x * SP_THREADS, q = q0 + threadIdx.x;
  if (q < total) {
    const uint32_t r = q / Q, s0 = (q - r * Q) * 4u;
    float z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = zval((int)s0 + j);
    if (t_rand != nullptr) {  // :99-106 stratified jitter with injected noise
      const float4 tr = *reinterpret_cast<const float4*>(t_rand + (size_t)q * 4);
      const float trv[4] = {tr.x, tr.y, tr.z, tr.w};
      const float zm = (s0 == 0) ? 0.f : zval((int)s0 - 1), zp = ((int)s0 + 4 >= Nc) ? 0.f : zval((int)s0 + 4);
      float zj[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = (int)s0 + j;
        const float prev = (j == 0) ? zm : z[j - 1], next = (j == 3) ? zp : z[j + 1];
        const float lo = (s == 0) ? z[j] : __fmul_rn(0.5f, __fadd_rn(z[j], prev));
        const float hi = (s == Nc - 1) ? z[j] : __fmul_rn(0.5f, __fadd_rn(next, z[j]));
        zj[j] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), trv[j]));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = zj[j];
    }
    __stcs(reinterpret_cast<float4*>(z_vals + (size_t)q * 4), make_float4(z[0], z[1], z[2], z[3]));
    if (pts != nullptr) {
      const float ox = rays_o[r * 3 + 0], oy = rays_o[r * 3 + 1], oz = rays_o[r * 3 + 2];
      const float dx = rays_d[r * 3 + 0], dy = rays_d[r * 3 + 1], dz = rays_d[r * 3 + 2];
      float p[12];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[3 * j + 0] = __fadd_rn(ox, __fmul_rn(dx, z[j]));
        p[3 * j + 1] = __fadd_rn(oy, __fmul_rn(dy, z[j]));
        p[3 * j + 2] = __fadd_rn(oz, __fmul_rn(dz, z[j]));
      }
      stage[3 * threadIdx.x + 0] = make_float4(p[0], p[1], p[2], p[3]);
      stage[3 * threadIdx.x + 1] = make_float4(p[4], p[5], p[6], p[7]);
      stage[3 * threadIdx.x + 2] = make_float4(p[8], p[9], p[10], p[11]);
    }
  }
  if (pts == nullptr) return;          // depths only: the MLP kernels form the positions themselves (StarPtsSrc)
  __syncthreads();
  const uint32_t nq = (total - q0 < (uint32_t)SP_THREADS) ? (total - q0) : (uint32_t)SP_THREADS;
  float4* po = reinterpret_cast<float4*>(pts + (size_t)q0 * 12);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t i = k * SP_THREADS + threadIdx.x;
    if (i < 3 * nq) __stcs(po + i, stage[i]);
  }
}

extern "C" int star_sample_pts(const float* rays_o, const float* rays_d, const float* t_vals,
                               const float* t_rand, float near_, float far_, int R, int Nc, int lindisp,
                               float* pts, float* z_vals, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!t_vals || !z_vals || (pts && (!rays_o || !rays_d))) return STAR_E_NULL;     // pts == NULL: depths only
  if (R < 0 || Nc < 1) return STAR_E_BAD_SHAPE;
  const int64_t total = (int64_t)R * Nc;
  const int threads = 256;
  if ((Nc & 3) == 0 && total / 4 < (int64_t)0x7fffffff &&
      (((uintptr_t)pts | (uintptr_t)z_vals | (uintptr_t)t_rand) & 15) == 0) {
    const int64_t nq = total / 4;
    sample_pts_x4_kernel<<<(unsigned)((nq + SP_THREADS - 1) / SP_THREADS), SP_THREADS, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, t_vals, t_rand, near_, far_, R, Nc, lindisp, pts, z_vals);
    return star_check_launch();
  }
  const int blocks = (int)((total + threads - 1) / threads < 148 * 16 ? (total + threads - 1) / threads : 148 * 16);
  sample_pts_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_vals, t_rand, near_, far_, R,
                                                                  Nc, lindisp, pts, z_vals);
  return star_check_launch();
}

// ------------------------------------------------------------------------------------------ a2
// models/rendering__.py:41-55 get_rays: pinhole rays of an H x W view (or of the pixel rows [row0, row0 + nrows)):
// dirs = [(i - cx) / fx, -(j - cy) / fy, -1]; rays_d = sum_k dirs[k] * c2w[c][k]; rays_o = c2w[:, 3].
// One rounding per reference op (no FMA contraction); optional fused viewdirs = rays_d / ||rays_d||.
__global__ void get_rays_kernel(int W, float fx, float fy, float cx, float cy, const float* __restrict__ c2w, int row0,
                                int64_t n, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                float* __restrict__ viewdirs) {
  __shared__ float m[12];
  if (threadIdx.x < 12) m[threadIdx.x] = c2w[threadIdx.x];
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int j = row0 + (int)(p / W), i = (int)(p % W);
    const float d0 = __fdiv_rn(__fsub_rn((float)i, cx), fx);
    const float d1 = -__fdiv_rn(__fsub_rn((float)j, cy), fy);
    const float d2 = -1.f;
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      r[c] = __fadd_rn(__fadd_rn(__fmul_rn(d0, m[c * 4 + 0]), __fmul_rn(d1, m[c * 4 + 1])), __fmul_rn(d2, m[c * 4 + 2]));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rays_d[p * 3 + c] = r[c];
      rays_o[p * 3 + c] = m[c * 4 + 3];
    }
    if (viewdirs != nullptr) {
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2])));
#pragma unroll
      for (int c = 0; c < 3; ++c) viewdirs[p * 3 + c] = __fdiv_rn(r[c], nrm);
    }
  }
}

extern "C" int star_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int row0, int nrows,
                             float* rays_o, float* rays_d, float* viewdirs, void* stream) {
  if (!c2w || !rays_o || !rays_d) return STAR_E_NULL;
  if (H < 1 || W < 1 || row0 < 0 || nrows < 0 || row0 + nrows > H) return STAR_E_BAD_SHAPE;
  if (nrows == 0) return STAR_OK;
  const int64_t n = (int64_t)nrows * W;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  get_rays_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(W, fx, fy, cx, cy, c2w, row0, n, rays_o, rays_d, viewdirs);
  return star_check_launch();
}

// ------------------------------------------------------------------------------------------ a3
// models/embedder.py:81-112 -- one thread per output element so that writes are coalesced.
__global__ void embed_kernel(const float* __restrict__ x, int64_t M, int L, const float* __restrict__ scale,
                             float* __restrict__ out) {
  const int D = 3 + 6 * L;
  const int64_t total = M * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / D;
    const int j = (int)(i % D);
    float v;
    if (j < 3) {
      v = x[m * 3 + j];
    } else {
      const int q = j - 3, k = q / 6, w = q % 6, c = w % 3;
      const float a = __fmul_rn(x[m * 3 + c], exp2f((float)k));
      v = (w < 3) ? sinf(a) : cosf(a);
    }
    if (scale != nullptr) v = __fmul_rn(v, scale[j]);
    out[i] = v;
  }
}

extern "C" int star_embed(const float* x, int M, int L, const float* scale, float* out, void* stream) {
  if (!x || !out) return STAR_E_NULL;
  if (M < 0 || L < 0 || L > 16) return STAR_E_BAD_SHAPE;
  if (M == 0) return STAR_OK;
  const int64_t total = (int64_t)M * (3 + 6 * L);
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(x, M, L, scale, out);
  return star_check_launch();
}

// ------------------------------------------------------------------------------------------ a9
// One warp per ray; cdf / bins of the ray live in shared memory (build_cdf_warp, invert_one: ray_device.cuh).

template <bool HAVE_CDF>
__global__ void sample_pdf_kernel(const float* __restrict__ bins, int64_t bins_stride,
                                  const float* __restrict__ weights, int64_t w_stride,
                                  const float* __restrict__ cdf_in, const float* __restrict__ u,
                                  const float* __restrict__ u_det, int R, int nb, int Ni,
                                  float* __restrict__ samples, int64_t* __restrict__ inds_o,
                                  int64_t* __restrict__ below_o, int64_t* __restrict__ above_o,
                                  float* __restrict__ cdf_o) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* cdf = smem + (size_t)warp * 2 * nb;
  float* sb = cdf + nb;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    if (HAVE_CDF) {
      for (int k = lane; k < nb; k += 32) cdf[k] = cdf_in[(int64_t)r * nb + k];
    } else {
      build_cdf_warp(weights + (int64_t)r * w_stride, nb - 1, cdf, lane);
    }
    for (int k = lane; k < nb; k += 32) sb[k] = bins[(int64_t)r * bins_stride + k];
    __syncwarp();
    if (cdf_o != nullptr)
      for (int k = lane; k < nb; k += 32) cdf_o[(int64_t)r * nb + k] = cdf[k];
    for (int j = lane; j < Ni; j += 32) {
      const float uu = (u != nullptr) ? u[(int64_t)r * Ni + j] : u_det[j];
      int i0, b0, a0;
      const float s = invert_one(cdf, sb, nb, uu, i0, b0, a0);
      const int64_t o = (int64_t)r * Ni + j;
      samples[o] = s;
      if (inds_o) inds_o[o] = i0;
      if (below_o) below_o[o] = b0;
      if (above_o) above_o[o] = a0;
    }
    __syncwarp();
  }
}

static int launch_cfg_warp_per_ray(int R, size_t smem_per_warp, int& blocks, int& threads, size_t& smem) {
  int wpb = 4;
  while (wpb > 1 && smem_per_warp * wpb > 200 * 1024) wpb >>= 1;
  if (smem_per_warp * wpb > 200 * 1024) return STAR_E_BAD_SHAPE;
  threads = wpb * 32;
  smem = smem_per_warp * wpb;
  int64_t b = ((int64_t)R + wpb - 1) / wpb;
#ifndef STAR_FUSED_GRID_CAP
#define STAR_FUSED_GRID_CAP 0x3fffffff     // grid over all rays (profiles/r2y_ab_grid_cap.txt)
#endif
  if (b > STAR_FUSED_GRID_CAP) b = STAR_FUSED_GRID_CAP;
  blocks = (int)b;
  return STAR_OK;
}

extern "C" int star_sample_pdf(const float* bins, int64_t bins_stride, const float* weights, int64_t w_stride,
                               const float* u, const float* u_det, int R, int nb, int Ni, float* samples,
                               int64_t* inds, int64_t* below, int64_t* above, float* cdf, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!bins || !weights || !samples || (!u && !u_det)) return STAR_E_NULL;
  if (R < 0 || nb < 2 || Ni < 1) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  int rc = launch_cfg_warp_per_ray(R, sizeof(float) * 2 * nb, blocks, threads, smem);
  if (rc) return rc;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sample_pdf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_pdf_kernel<false><<<blocks, threads, smem, (cudaStream_t)stream>>>(
      bins, bins_stride, weights, w_stride, nullptr, u, u_det, R, nb, Ni, samples, inds, below, above, cdf);
  return star_check_launch();
}

extern "C" int star_invert_cdf(const float* bins, const float* cdf, const float* u, int R, int nb, int Ni,
                               float* samples, int64_t* inds, int64_t* below, int64_t* above, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!bins || !cdf || !u || !samples) return STAR_E_NULL;
  if (R < 0 || nb < 1 || Ni < 1) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  int rc = launch_cfg_warp_per_ray(R, sizeof(float) * 2 * nb, blocks, threads, smem);
  if (rc) return rc;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sample_pdf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_pdf_kernel<true><<<blocks, threads, smem, (cudaStream_t)stream>>>(
      bins, nb, nullptr, 0, cdf, u, nullptr, R, nb, Ni, samples, inds, below, above, nullptr);
  return star_check_launch();
}

// ------------------------------------------------------------------------------------------ a10
// z_mid -> sample_pdf -> sort(cat(z_vals, z_samples)) -> z_std, pts.  One warp per ray (hier_ray, ray_device.cuh).
// GIVEN = true: the fine samples are supplied in z_samples (read, not written) and only the merge,
// z_std and pts are computed (star_merge_samples).
template <bool GIVEN>
__global__ void hierarchical_kernel(const float* __restrict__ z_vals, const float* __restrict__ weights,
                                    const float* __restrict__ u, const float* __restrict__ u_det,
                                    const float* __restrict__ rays_o, const float* __restrict__ rays_d, int R,
                                    int Nc, int Ni, int P, float* z_samples,
                                    float* __restrict__ z_all, float* __restrict__ z_std,
                                    float* __restrict__ pts_fine) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nb = Nc - 1, Nf = Nc + Ni;
  float* za = smem + (size_t)warp * hier_smem_floats(Nc, Ni, P);   // 16-byte aligned (region is a multiple of 4)
  float* zs = za + ((Nf + 3) & ~3);
  float* zc = zs + P;
  float* cdf = zc + Nc;
  float* sb = cdf + nb;
  int* gs = reinterpret_cast<int*>(sb + nb);   // per fine sample: below + 1 = first guess of its rank among zc
  const bool vec4 = ((Nf & 3) == 0) && (((uintptr_t)z_all | (uintptr_t)pts_fine) & 15) == 0;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    hier_ray<GIVEN, false>(z_vals + (int64_t)r * Nc, GIVEN ? nullptr : weights + (int64_t)r * Nc, u, u_det, rays_o, rays_d, r,
                           Nc, Ni, P, z_samples, z_all, z_std, pts_fine, za, zs, zc, cdf, sb, gs, vec4, lane);
  }
}

extern "C" int star_hierarchical(const float* z_vals, const float* weights, const float* u, const float* u_det,
                                 const float* rays_o, const float* rays_d, int R, int Nc, int Ni,
                                 float* z_samples, float* z_all, float* z_std, float* pts_fine, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!z_vals || !weights || !z_samples || !z_all || !z_std || (!u && !u_det)) return STAR_E_NULL;
  if (pts_fine && (!rays_o || !rays_d)) return STAR_E_NULL;
  if (R < 0 || Nc < 3 || Ni < 1 || Nc + Ni > 8192) return STAR_E_BAD_SHAPE;
  int P = 2;
  while (P < Ni) P <<= 1;
  int blocks, threads;
  size_t smem;
  int rc = launch_cfg_warp_per_ray(R, sizeof(float) * (size_t)hier_smem_floats(Nc, Ni, P), blocks, threads, smem);
  if (rc) return rc;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(hierarchical_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  hierarchical_kernel<false><<<blocks, threads, smem, (cudaStream_t)stream>>>(
      z_vals, weights, u, u_det, rays_o, rays_d, R, Nc, Ni, P, z_samples, z_all, z_std, pts_fine);
  return star_check_launch();
}

extern "C" int star_merge_samples(const float* z_vals, const float* z_samples, const float* rays_o,
                                  const float* rays_d, int R, int Nc, int Ni, float* z_all, float* z_std,
                                  float* pts_fine, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!z_vals || !z_samples || !z_all || !z_std) return STAR_E_NULL;
  if (pts_fine && (!rays_o || !rays_d)) return STAR_E_NULL;
  if (R < 0 || Nc < 1 || Ni < 1 || Nc + Ni > 8192) return STAR_E_BAD_SHAPE;
  int P = 2;
  while (P < Ni) P <<= 1;
  int blocks, threads;
  size_t smem;
  int rc = launch_cfg_warp_per_ray(R, sizeof(float) * (size_t)hier_smem_floats(Nc, Ni, P), blocks, threads, smem);
  if (rc) return rc;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(hierarchical_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  hierarchical_kernel<true><<<blocks, threads, smem, (cudaStream_t)stream>>>(
      z_vals, nullptr, nullptr, nullptr, rays_o, rays_d, R, Nc, Ni, P, const_cast<float*>(z_samples), z_all,
      z_std, pts_fine);
  return star_check_launch();
}
