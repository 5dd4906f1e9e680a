// Per-ray device code shared by the compositing kernels (composite.cu), the sampling kernels (sampling.cu) and the fused
// coarse-pass kernel (ray_fused.cu): one warp owns one ray.  The stand-alone kernels and the fused kernel instantiate the
// same functions, so their outputs agree bit for bit (tests/test_gpu_parity.py compares them).
#pragma once
#include "star_common.cuh"

struct RayGeom {
  float norm;  // ||rays_d||  (:325)
};

__device__ __forceinline__ float ray_norm(const float* __restrict__ rays_d, int r) {
  const float x = rays_d[r * 3 + 0], y = rays_d[r * 3 + 1], z = rays_d[r * 3 + 2];
  return sqrtf(x * x + y * y + z * z);
}

// dists (:318-325): z[s+1]-z[s], last = far_dist, times ||rays_d||
__device__ __forceinline__ float sample_dist(const float* __restrict__ zr, int s, int S, float far_dist, float norm) {
  const float d = (s == S - 1) ? far_dist : (zr[s + 1] - zr[s]);
  return d * norm;
}

// alpha = 1 - exp(-softplus(raw) * dist)   (:301-303)
__device__ __forceinline__ float alpha_of(float raw, float dist) { return 1.f - fexp(-softplus_f(raw) * dist); }
// d alpha / d raw
__device__ __forceinline__ float dalpha_draw(float raw, float dist) {
  return dist * fexp(-softplus_f(raw) * dist) * softplus_grad_f(raw);
}

// exclusive prefix product of m over the warp given the running carry; updates carry
__device__ __forceinline__ float excl_transmittance(float m, float& carry, int lane) {
  const float incl = warp_scan_prod(m, lane);
  float excl = __shfl_up_sync(STAR_FULL_MASK, incl, 1);
  if (lane == 0) excl = 1.f;
  const float T = carry * excl;
  carry *= __shfl_sync(STAR_FULL_MASK, incl, 31);
  return T;
}

// Single-field compositing of ray r, even S (models/rendering__.py:307-379): every lane owns two consecutive samples
// (8-byte loads / stores, one transmittance scan and one trip of the loop per 64 samples) and the five per-ray sums
// share 11 shuffles (warp_sum4).  w_sm / z_sm (optional, S floats each, shared memory of this warp) also receive the
// weights and the depths for a consumer in the same kernel (ray_fused.cu); weights_o may then be NULL.
__device__ __forceinline__ void composite_single_ray_x2(const float* __restrict__ raw_alpha, const float* __restrict__ raw_rgb,
                                                        const float* __restrict__ z_vals, const float* __restrict__ rays_d,
                                                        int r, int S, float far_dist, int white_bkgd,
                                                        float* __restrict__ rgb_o, float* __restrict__ disp_o,
                                                        float* __restrict__ acc_o, float* __restrict__ depth_o,
                                                        float* __restrict__ weights_o, float* __restrict__ dists_o,
                                                        float* w_sm, float* z_sm, int lane) {
    const float norm = ray_norm(rays_d, r);
    const int64_t row = (int64_t)r * S;
    const float* zr = z_vals + row;
    float carry = 1.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
    for (int base = 0; base < S; base += 64) {
      const int s0 = base + 2 * lane;
      const bool ok = s0 < S;   // S even: both samples of the pair are valid together
      float2 z = make_float2(0.f, 0.f), ra = make_float2(0.f, 0.f);
      float2 c01 = make_float2(0.f, 0.f), c23 = c01, c45 = c01;
      if (ok) {
        z = *reinterpret_cast<const float2*>(zr + s0);
        ra = *reinterpret_cast<const float2*>(raw_alpha + row + s0);
        const float2* cp = reinterpret_cast<const float2*>(raw_rgb + (row + s0) * 3);
        c01 = cp[0]; c23 = cp[1]; c45 = cp[2];
      }
      float znext = __shfl_down_sync(STAR_FULL_MASK, z.x, 1);
      if (lane == 31 && base + 64 < S) znext = zr[base + 64];
      float2 dist = make_float2(0.f, 0.f), alpha = make_float2(0.f, 0.f);
      if (ok) {
        dist.x = (z.y - z.x) * norm;
        dist.y = ((s0 + 1 == S - 1) ? far_dist : (znext - z.y)) * norm;
        alpha.x = alpha_of(ra.x, dist.x);
        alpha.y = alpha_of(ra.y, dist.y);
      }
      const float m0 = 1.f - alpha.x + 1e-10f, m1 = 1.f - alpha.y + 1e-10f;   // (alpha = 0 -> m = 1 when !ok)
      const float T0 = excl_transmittance(ok ? m0 * m1 : 1.f, carry, lane);
      if (ok) {
        const float w0 = alpha.x * T0, w1 = alpha.y * (T0 * m0);
        if (weights_o) *reinterpret_cast<float2*>(weights_o + row + s0) = make_float2(w0, w1);
        if (dists_o) *reinterpret_cast<float2*>(dists_o + row + s0) = dist;
        if (w_sm) { w_sm[s0] = w0; w_sm[s0 + 1] = w1; }
        if (z_sm) { z_sm[s0] = z.x; z_sm[s0 + 1] = z.y; }
        sr += w0 * sigmoid_f(c01.x) + w1 * sigmoid_f(c23.y);
        sg += w0 * sigmoid_f(c01.y) + w1 * sigmoid_f(c45.x);
        sb += w0 * sigmoid_f(c23.x) + w1 * sigmoid_f(c45.y);
        sd += w0 * z.x + w1 * z.y;
        sa += w0 + w1;
      }
    }
    sa = warp_sum(sa);
    const float v = warp_sum4(sr, sg, sb, sd, lane);   // lanes 0-7: r, 8-15: g, 16-23: b, 24-31: depth
    const float bg = white_bkgd ? (1.f - sa) : 0.f;                   // :360-361
    if ((lane & 7) == 0) {
      if (lane < 24) {
        rgb_o[r * 3 + (lane >> 3)] = v + bg;
      } else {
        const float wsum = (sa >= 0.f) ? sa : 1e-7f;                  // :353-354
        disp_o[r] = 1.f / max_nan_f(1e-10f, v / wsum);                    // :355-357
        depth_o[r] = v;
        acc_o[r] = sa;
      }
    }
}

// ------------------------------------------------------------------------------------------ a9
// One warp per ray.  cdf / bins of the ray live in shared memory.
// Defined arithmetic (DESIGN.md "sample_pdf arithmetic"):
//   w   = weights + 1e-5f                                   (rendering__.py:722)
//   s   = fp32( sum_k w_k accumulated in fp64 )             exactly rounded normaliser (:723)
//   pdf = w / s                                             IEEE fp32 division
//   cdf = fp32( prefix sums of pdf accumulated in fp64 )    == torch CPU cumsum (:733-734)
// The fp64 sums of <= 2^10 fp32 values spanning <= 2^17 in magnitude are exact, so the warp-parallel
// order gives the same bits as a sequential loop.
__device__ __forceinline__ void build_cdf_warp(const float* __restrict__ w_row, int nw, float* cdf, int lane) {
  double part = 0.0;
  for (int k = lane; k < nw; k += 32) part += (double)__fadd_rn(w_row[k], 1e-5f);
  const float s = (float)warp_sum_d(part);
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < nw; base += 32) {
    const int k = base + lane;
    double p = 0.0;
    if (k < nw) p = (double)__fdiv_rn(__fadd_rn(w_row[k], 1e-5f), s);
    const double incl = warp_scan_sum_d(p, lane) + carry;
    if (k < nw) cdf[k + 1] = (float)incl;
    carry = __shfl_sync(STAR_FULL_MASK, incl, 31);
  }
}

// searchsorted(cdf, u, right=True) (:745) + gather + lerp (:746-759)
__device__ __forceinline__ float invert_one(const float* cdf, const float* bins, int nb, float u, int& inds,
                                            int& below, int& above) {
  int lo = 0, hi = nb;  // first index with cdf[idx] > u
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  inds = lo;
  below = max(0, lo - 1);
  above = min(nb - 1, lo);
  const float c0 = cdf[below], c1 = cdf[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.f;
  const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  const float b0 = bins[below], b1 = bins[above];
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

// ------------------------------------------------------------------------------------------ a10
// z_mid -> sample_pdf -> sort(cat(z_vals, z_samples)) -> z_std, pts.  One warp per ray.
// smem per warp: see hier_smem_floats (P = next pow2 >= Ni; zs is padded with +inf up to P)
// The concatenation is sorted as a MERGE: the coarse samples are sorted by construction, the fine samples are
// sorted whenever u is (always in eval mode: inverse-CDF sampling is monotone); only otherwise (random u in
// training) are the Ni fine samples bitonic-sorted first.  Every element then finds its output slot with one
// binary search in the other list (coarse before fine on ties; torch.sort returns values only, :136,279).
// GIVEN = true: the fine samples are supplied in z_samples (read, not written) and only the merge,
// z_std and pts are computed (star_merge_samples).
__device__ __forceinline__ int count_less(const float* a, int n, float x) {      // # a[i] <  x, a sorted
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int count_less_equal(const float* a, int n, float x) {  // # a[i] <= x, a sorted
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}

// floats of shared memory per warp: zall[Nf, padded to 4] | zs[P] | zc[Nc] | cdf[nb] | bins[nb] | guess[Ni]
__host__ __device__ __forceinline__ int hier_smem_floats(int Nc, int Ni, int P) {
  return ((((Nc + Ni + 3) & ~3) + P + Nc + 2 * (Nc - 1) + Ni) + 3) & ~3;
}

// The coarse -> fine step of ray r: z_mid -> sample_pdf(weights[1:-1]) -> z_std -> sort(cat(z_vals, z_samples)) as a merge
// (+ optional fine positions).  zr = the ray's Nc coarse depths, w_row = its Nc coarse weights (global or shared memory);
// za / zs / zc / cdf / sb / gs = this warp's shared memory (hier_smem_floats).  ZC_READY: zc already holds the coarse
// depths (the fused kernel's compositing part put them there).
template <bool GIVEN, bool ZC_READY>
__device__ __forceinline__ void hier_ray(const float* zr, const float* w_row, const float* __restrict__ u,
                                         const float* __restrict__ u_det, const float* __restrict__ rays_o,
                                         const float* __restrict__ rays_d, int r, int Nc, int Ni, int P, float* z_samples,
                                         float* __restrict__ z_all, float* __restrict__ z_std,
                                         float* __restrict__ pts_fine, float* za, float* zs, float* zc, float* cdf,
                                         float* sb, int* gs, bool vec4, int lane) {
  const int nb = Nc - 1, Nf = Nc + Ni;
  if (!GIVEN) {
    build_cdf_warp(w_row + 1, Nc - 2, cdf, lane);  // weights[..., 1:-1]  (:131,274)
    for (int k = lane; k < nb; k += 32) sb[k] = __fmul_rn(0.5f, __fadd_rn(zr[k + 1], zr[k]));  // z_mid (:128)
  }
  if (!ZC_READY) for (int k = lane; k < Nc; k += 32) zc[k] = zr[k];
  for (int k = Ni + lane; k < P; k += 32) zs[k] = __int_as_float(0x7f800000);
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < Ni; j += 32) {
    float s;
    if (GIVEN) {
      s = z_samples[(int64_t)r * Ni + j];
    } else {
      const float uu = (u != nullptr) ? u[(int64_t)r * Ni + j] : u_det[j];
      int i0, b0, a0;
      s = invert_one(cdf, sb, nb, uu, i0, b0, a0);
      z_samples[(int64_t)r * Ni + j] = s;
      gs[j] = b0 + 1;
    }
    zs[j] = s;
    sum += s;
  }
  // z_std = std(z_samples, unbiased=False)   (:144,296)
  const float mean = warp_sum(sum) / (float)Ni;
  __syncwarp();
  float var = 0.f;
  bool sorted = true;
  for (int j = lane; j < Ni; j += 32) {
    const float d = zs[j] - mean;
    var += d * d;
    if (j + 1 < Ni && zs[j] > zs[j + 1]) sorted = false;
  }
  var = warp_sum(var) / (float)Ni;
  if (lane == 0) z_std[r] = sqrtf(var);
  const bool all_sorted = __all_sync(STAR_FULL_MASK, sorted);
  if (!all_sorted) {
    // bitonic sort of zs[0..P)
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        __syncwarp();
        for (int i = lane; i < P; i += 32) {
          const int l = i ^ j;
          if (l > i) {
            const float a = zs[i], b = zs[l];
            const bool up = ((i & k) == 0);
            if ((a > b) == up) { zs[i] = b; zs[l] = a; }
          }
        }
      }
    }
    __syncwarp();
  }
  // merge by rank
  for (int i = lane; i < Nc; i += 32) {
    const float v = zc[i];
    za[i + count_less(zs, Ni, v)] = v;
  }
  for (int j = lane; j < Ni; j += 32) {
    const float v = zs[j];
    int k;
    if (!GIVEN && all_sorted) {
      // a sample drawn from bin [z_mid[b], z_mid[b+1]] has b+1 or b+2 coarse samples at or below it: start from
      // the guess and walk (exact for any input, O(1) here)
      k = gs[j];
      while (k < Nc && zc[k] <= v) ++k;
      while (k > 0 && zc[k - 1] > v) --k;
    } else {
      k = count_less_equal(zc, Nc, v);
    }
    za[j + k] = v;
  }
  __syncwarp();
  const float ox = rays_o ? rays_o[r * 3 + 0] : 0.f, oy = rays_o ? rays_o[r * 3 + 1] : 0.f, oz = rays_o ? rays_o[r * 3 + 2] : 0.f;
  const float dx = rays_d ? rays_d[r * 3 + 0] : 0.f, dy = rays_d ? rays_d[r * 3 + 1] : 0.f, dz = rays_d ? rays_d[r * 3 + 2] : 0.f;
  if (vec4) {
    for (int q = lane; q < (Nf >> 2); q += 32) {
      const float4 z = *reinterpret_cast<const float4*>(za + 4 * q);
      __stcs(reinterpret_cast<float4*>(z_all + (int64_t)r * Nf) + q, z);
      if (pts_fine != nullptr) {
        float4* po = reinterpret_cast<float4*>(pts_fine + ((int64_t)r * Nf + 4 * q) * 3);
        __stcs(po + 0, make_float4(__fadd_rn(ox, __fmul_rn(dx, z.x)), __fadd_rn(oy, __fmul_rn(dy, z.x)),
                                   __fadd_rn(oz, __fmul_rn(dz, z.x)), __fadd_rn(ox, __fmul_rn(dx, z.y))));
        __stcs(po + 1, make_float4(__fadd_rn(oy, __fmul_rn(dy, z.y)), __fadd_rn(oz, __fmul_rn(dz, z.y)),
                                   __fadd_rn(ox, __fmul_rn(dx, z.z)), __fadd_rn(oy, __fmul_rn(dy, z.z))));
        __stcs(po + 2, make_float4(__fadd_rn(oz, __fmul_rn(dz, z.z)), __fadd_rn(ox, __fmul_rn(dx, z.w)),
                                   __fadd_rn(oy, __fmul_rn(dy, z.w)), __fadd_rn(oz, __fmul_rn(dz, z.w))));
      }
    }
  } else {
    for (int k = lane; k < Nf; k += 32) z_all[(int64_t)r * Nf + k] = za[k];
    if (pts_fine != nullptr) {
      for (int i = lane; i < Nf * 3; i += 32) {
        const int s = i / 3, c = i - s * 3;
        const float o = c == 0 ? ox : (c == 1 ? oy : oz), d = c == 0 ? dx : (c == 1 ? dy : dz);
        pts_fine[(int64_t)r * Nf * 3 + i] = __fadd_rn(o, __fmul_rn(d, za[s]));
      }
    }
  }
  __syncwarp();
}
