// Sampling and compositing kernels of the mip-NeRF variant:
//   uniform frustum edges       nerfstudio UniformSampler            (models/star_mipnerf.py:75-77,271)
//   importance frustum edges    nerfstudio PDFSampler                (star_mipnerf.py:78-81,286-288)
//   density-space compositing   models/rendering_starmip.py:32-175   (+ nerfstudio median DepthRenderer)
// One warp per ray; lanes stride over samples; prefix sums are warp-shuffle scans with a running carry.
// Index decisions (searchsorted) use explicitly rounded arithmetic, like sampling.cu.
#include "star_common.cuh"

// inclusive prefix sum across the warp
__device__ __forceinline__ float warp_scan_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(STAR_FULL_MASK, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

__device__ __forceinline__ float to_euclid(float x, float near_, float far_) {   // x * far + (1 - x) * near
  return __fadd_rn(__fmul_rn(x, far_), __fmul_rn(__fsub_rn(1.f, x), near_));
}

// ------------------------------------------------------------------------------------------ uniform edges
__global__ void mip_uniform_bins_kernel(const float* __restrict__ lin, const float* __restrict__ t_rand, float near_,
                                        float far_, int R, int nb, float* __restrict__ spacing,
                                        float* __restrict__ euclid) {
  const int64_t total = (int64_t)R * nb;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % nb);
    float b = lin[k];
    if (t_rand != nullptr) {   // stratified jitter of the edges (SpacedSampler, train_stratified)
      const float lo = (k == 0) ? b : __fdiv_rn(__fadd_rn(b, lin[k - 1]), 2.f);
      const float hi = (k == nb - 1) ? b : __fdiv_rn(__fadd_rn(lin[k + 1], b), 2.f);
      b = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[i]));
    }
    spacing[i] = b;
    euclid[i] = to_euclid(b, near_, far_);
  }
}

extern "C" int star_mip_uniform_bins(const float* lin, const float* t_rand, float near_, float far_, int R, int Nc,
                                     float* spacing, float* euclid, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!lin || !spacing || !euclid) return STAR_E_NULL;
  if (R < 0 || Nc < 1) return STAR_E_BAD_SHAPE;
  const int64_t total = (int64_t)R * (Nc + 1);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mip_uniform_bins_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(lin, t_rand, near_, far_, R, Nc + 1, spacing,
                                                                          euclid);
  return star_check_launch();
}

// ------------------------------------------------------------------------------------------ PDF sampler
// Defined arithmetic (same policy as sample_pdf): w = weights + 0.01; normaliser = exactly rounded fp32
// sum (fp64 accumulate) with nerfstudio's eps padding; pdf = w / sum (IEEE); cdf = min(1, fp64 prefix sums rounded
// per element), 0 prepended.  smem per warp: cdf[Nc+1] | bins[Nc+1].
__global__ void mip_pdf_sample_kernel(const float* __restrict__ spacing_bins, const float* __restrict__ weights,
                                      int64_t w_stride, const float* __restrict__ u_base,
                                      const float* __restrict__ u_rand, float near_, float far_, int R, int Nc, int Ni,
                                      float* __restrict__ spacing_out, float* __restrict__ euclid_out,
                                      int64_t* __restrict__ inds_o, float* __restrict__ cdf_o) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int ne = Nc + 1, nb = Ni + 1;
  float* cdf = smem + (size_t)warp * 2 * ne;
  float* sb = cdf + ne;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float* wr = weights + (int64_t)r * w_stride;
    double part = 0.0;
    for (int k = lane; k < Nc; k += 32) part += (double)__fadd_rn(wr[k], 0.01f);
    float wsum = (float)warp_sum_d(part);
    const float padding = fmaxf(__fsub_rn(1e-5f, wsum), 0.f);
    const float pad_each = __fdiv_rn(padding, (float)Nc);
    wsum = __fadd_rn(wsum, padding);
    double carry = 0.0;
    if (lane == 0) cdf[0] = 0.f;
    for (int base = 0; base < Nc; base += 32) {
      const int k = base + lane;
      double p = 0.0;
      if (k < Nc) p = (double)__fdiv_rn(__fadd_rn(__fadd_rn(wr[k], 0.01f), pad_each), wsum);
      const double incl = warp_scan_sum_d(p, lane) + carry;
      if (k < Nc) cdf[k + 1] = fminf(1.f, (float)incl);
      carry = __shfl_sync(STAR_FULL_MASK, incl, 31);
    }
    for (int k = lane; k < ne; k += 32) sb[k] = spacing_bins[(int64_t)r * ne + k];
    __syncwarp();
    if (cdf_o != nullptr)
      for (int k = lane; k < ne; k += 32) cdf_o[(int64_t)r * ne + k] = cdf[k];
    for (int j = lane; j < nb; j += 32) {
      float u = u_base[j];
      if (u_rand != nullptr) u = __fadd_rn(u, __fdiv_rn(u_rand[(int64_t)r * nb + j], (float)nb));
      int lo = 0, hi = ne;   // searchsorted(cdf, u, side="right"): first index with cdf[idx] > u
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      const int below = min(max(lo - 1, 0), ne - 1), above = min(max(lo, 0), ne - 1);
      const float c0 = cdf[below], c1 = cdf[above], b0 = sb[below], b1 = sb[above];
      float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
      t = (t != t) ? 0.f : fminf(fmaxf(t, 0.f), 1.f);   // clip(nan_to_num(., 0), 0, 1)
      const float b = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      const int64_t o = (int64_t)r * nb + j;
      spacing_out[o] = b;
      euclid_out[o] = to_euclid(b, near_, far_);
      if (inds_o != nullptr) inds_o[o] = lo;
    }
    __syncwarp();
  }
}

extern "C" int star_mip_pdf_sample(const float* spacing_bins, const float* weights, int64_t w_stride,
                                   const float* u_base, const float* u_rand, float near_, float far_, int R, int Nc,
                                   int Ni, float* spacing_out, float* euclid_out, int64_t* inds, float* cdf,
                                   void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!spacing_bins || !weights || !u_base || !spacing_out || !euclid_out) return STAR_E_NULL;
  if (R < 0 || Nc < 1 || Ni < 1 || Nc > 8192) return STAR_E_BAD_SHAPE;
  const int wpb = 4;
  const size_t smem = sizeof(float) * 2 * (size_t)(Nc + 1) * wpb;
  int64_t blocks = ((int64_t)R + wpb - 1) / wpb;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(mip_pdf_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mip_pdf_sample_kernel<<<(int)blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
      spacing_bins, weights, w_stride, u_base, u_rand, near_, far_, R, Nc, Ni, spacing_out, euclid_out, inds, cdf);
  return star_check_launch();
}

// ------------------------------------------------------------------------------------------ compositing
// sigma = softplus(raw) (DensityFieldHead), colour = sigmoid(raw) (RGBFieldHead); dd = delta * sigma;
// alpha = 1 - exp(-dd); T = exp(-exclusive_cumsum(dd)); weights = nan_to_num(alpha * T)   (rendering_starmip.py:32-63)
struct ScanState {
  float carry;   // sum of dd over the samples before this chunk
};

__device__ __forceinline__ float excl_scan_sum(float v, float& carry, int lane) {
  const float incl = warp_scan_sum(v, lane);
  const float excl = carry + (incl - v);
  carry += __shfl_sync(STAR_FULL_MASK, incl, 31);
  return excl;
}

__device__ __forceinline__ float nan_to_num_f(float x) {
  if (x != x) return 0.f;
  if (isinf(x)) return x > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return x;
}

// running median-depth search: first sample whose inclusive cumulative weight reaches 0.5 (searchsorted side="left")
struct MedianState {
  float cw;      // cumulative weight before this chunk
  int idx;       // -1 until found
};
__device__ __forceinline__ void median_step(float w, bool ok, int s, MedianState& st, int lane) {
  const float incl = warp_scan_sum(ok ? w : 0.f, lane) + st.cw;
  const unsigned hit = __ballot_sync(STAR_FULL_MASK, ok && incl >= 0.5f);
  if (st.idx < 0 && hit != 0u) st.idx = s - lane + (__ffs(hit) - 1);
  st.cw = __shfl_sync(STAR_FULL_MASK, incl, 31);
}
__device__ __forceinline__ float median_depth(const float* __restrict__ br, int idx, int S) {
  const int i = idx < 0 ? S - 1 : idx;   // clamp(searchsorted, 0, S - 1)
  return (br[i] + br[i + 1]) * 0.5f;
}

__global__ void mip_composite_single_fwd_kernel(const float* __restrict__ raw_sigma, const float* __restrict__ raw_rgb,
                                                const float* __restrict__ bins, int R, int S,
                                                float* __restrict__ rgb_o, float* __restrict__ acc_o,
                                                float* __restrict__ depth_o, float* __restrict__ weights_o) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    const float* br = bins + (int64_t)r * (S + 1);
    const int64_t row = (int64_t)r * S;
    float carry = 0.f, sr = 0.f, sg = 0.f, sbl = 0.f, sa = 0.f;
    MedianState med = {0.f, -1};
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float dd = 0.f;
      if (ok) dd = (br[s + 1] - br[s]) * softplus_f(raw_sigma[row + s]);
      const float T = fexp(-excl_scan_sum(dd, carry, lane));
      float w = 0.f;
      if (ok) {
        const float alpha = 1.f - fexp(-dd);
        const float ta = T * alpha;
        w = nan_to_num_f(alpha * T);
        weights_o[row + s] = w;
        const float* c = raw_rgb + (row + s) * 3;
        sr += ta * sigmoid_f(c[0]); sg += ta * sigmoid_f(c[1]); sbl += ta * sigmoid_f(c[2]);
        sa += w;
      }
      median_step(w, ok, s, med, lane);
    }
    sa = warp_sum(sa);
    const float v = warp_sum4(sr, sg, sbl, 0.f, lane);
    if ((lane & 7) == 0 && lane < 24) rgb_o[r * 3 + (lane >> 3)] = v;
    if (lane == 0) {
      acc_o[r] = sa;
      depth_o[r] = median_depth(br, med.idx, S);
    }
  }
}

// Gradients of (rgb, acc, weights); the median depth is piecewise constant in the densities.  smem per warp: wG[S]
__global__ void mip_composite_single_bwd_kernel(const float* __restrict__ raw_sigma, const float* __restrict__ raw_rgb,
                                                const float* __restrict__ bins, int R, int S,
                                                const float* __restrict__ g_rgb, const float* __restrict__ g_acc,
                                                const float* __restrict__ g_weights, float* __restrict__ d_raw_sigma,
                                                float* __restrict__ d_raw_rgb) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* s_T = smem + (size_t)warp * S;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float* br = bins + (int64_t)r * (S + 1);
    const int64_t row = (int64_t)r * S;
    float carry = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float dd = 0.f;
      if (ok) dd = (br[s + 1] - br[s]) * softplus_f(raw_sigma[row + s]);
      const float T = fexp(-excl_scan_sum(dd, carry, lane));
      if (ok) s_T[s] = T;
    }
    const float gr = g_rgb ? g_rgb[r * 3 + 0] : 0.f, gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f,
                gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
    const float gA = g_acc ? g_acc[r] : 0.f;
    __syncwarp();
    float suffix = 0.f;
    for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float v = 0.f, G = 0.f, alpha = 0.f, T = 0.f, delta = 0.f, raw = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (ok) {
        raw = raw_sigma[row + s];
        delta = br[s + 1] - br[s];
        alpha = 1.f - fexp(-delta * softplus_f(raw));
        T = s_T[s];
        const float* c = raw_rgb + (row + s) * 3;
        c0 = sigmoid_f(c[0]); c1 = sigmoid_f(c[1]); c2 = sigmoid_f(c[2]);
        G = gr * c0 + gg * c1 + gb * c2 + gA + (g_weights ? g_weights[row + s] : 0.f);
        v = alpha * T * G;
      }
      const float incl = warp_rscan_sum(v, lane);
      const float after = incl - v + suffix;
      suffix += __shfl_sync(STAR_FULL_MASK, incl, 0);
      if (ok) {
        const float d_dd = T * G * (1.f - alpha) - after;
        d_raw_sigma[row + s] = d_dd * delta * softplus_grad_f(raw);
        const float w = alpha * T;
        float* o = d_raw_rgb + (row + s) * 3;
        o[0] = w * gr * c0 * (1.f - c0);
        o[1] = w * gg * c1 * (1.f - c1);
        o[2] = w * gb * c2 * (1.f - c2);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------- multi field
#define MRAD(r, v, s) raw_sigma_d[((int64_t)(r) * V + (v)) * S + (s)]
#define MRCD(r, v, s, c) raw_rgb_d[(((int64_t)(r) * V + (v)) * S + (s)) * 3 + (c)]

__device__ __forceinline__ float mip_chunk_inv_count(int r, int R, int chunk) {
  const int c0 = (r / chunk) * chunk;
  return 1.f / (float)(min(R, c0 + chunk) - c0);
}
__device__ __forceinline__ float mip_bin_entropy_term(float a) {
  const float c = fminf(fmaxf(a, STAR_EPS_F32), 1.f - STAR_EPS_F32);
  return a * flog(c) + (1.f - a) * flog1m(c);
}

// regs[5] = alpha_entropy, dynamic_vs_static, ray_reg (mip form: NO max over samples), static_reg (== 0), dynamic_reg
template <int VT>
__global__ void __launch_bounds__(128)
mip_composite_multi_fwd_kernel(const float* __restrict__ raw_sigma_s, const float* __restrict__ raw_rgb_s,
                               const float* __restrict__ raw_sigma_d, const float* __restrict__ raw_rgb_d,
                               const float* __restrict__ bins, int R, int V, int S, int chunk, StarMipMultiOut out,
                               float* __restrict__ reg_partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float reg_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float* br = bins + (int64_t)r * (S + 1);
    const int64_t row = (int64_t)r * S;
    float cT = 0.f, cTs = 0.f, cTd[VT];
    float s_rgb[3] = {0, 0, 0}, s_rgbs[3] = {0, 0, 0}, s_rgbd[VT][3];
    MedianState med = {0.f, -1}, med_s = {0.f, -1}, med_d[VT];
    float s_acc = 0.f, ent = 0.f, dvs = 0.f, dyn = 0.f, rayreg = 0.f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
      cTd[v] = 0.f; med_d[v].cw = 0.f; med_d[v].idx = -1;
      s_rgbd[v][0] = s_rgbd[v][1] = s_rgbd[v][2] = 0.f;
    }
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float delta = 0.f, sig_s = 0.f, a_s = 0.f, sig_tot = 0.f;
      float cs[3] = {0, 0, 0}, mixd[3] = {0, 0, 0};
      float a_dv[VT], sig_dv[VT];
      float a_dsum = 0.f, ent_d = 0.f;
      if (ok) {
        delta = br[s + 1] - br[s];
        sig_s = softplus_f(raw_sigma_s[row + s]);
        a_s = 1.f - fexp(-delta * sig_s);
#pragma unroll
        for (int c = 0; c < 3; ++c) cs[c] = sigmoid_f(raw_rgb_s[(row + s) * 3 + c]);
        sig_tot = sig_s;
      }
#pragma unroll
      for (int v = 0; v < VT; ++v) {
        float sg = 0.f, a_d = 0.f, cd[3] = {0, 0, 0};
        if (ok) {
          sg = softplus_f(MRAD(r, v, s));
          a_d = 1.f - fexp(-delta * sg);
#pragma unroll
          for (int c = 0; c < 3; ++c) cd[c] = sigmoid_f(MRCD(r, v, s, c));
          sig_tot += sg;
        }
        a_dv[v] = a_d; sig_dv[v] = sg;
        const float Td = fexp(-excl_scan_sum(delta * sg, cTd[v], lane));
        float wd = 0.f;
        if (ok) {
          const float ta = Td * a_d;
          wd = nan_to_num_f(a_d * Td);
#pragma unroll
          for (int c = 0; c < 3; ++c) { s_rgbd[v][c] += ta * cd[c]; mixd[c] += a_d * cd[c]; }
          if (s == S - 1) out.dynamic_transmittance[(int64_t)r * V + v] = Td;   // rendering_starmip.py:167
          a_dsum += a_d;
          ent_d += mip_bin_entropy_term(a_d);
          dyn += sg;
        }
        median_step(wd, ok, s, med_d[v], lane);
      }
      const float a_t = ok ? 1.f - fexp(-delta * sig_tot) : 0.f;
      const float T = fexp(-excl_scan_sum(delta * sig_tot, cT, lane));
      const float Ts = fexp(-excl_scan_sum(delta * sig_s, cTs, lane));
      float w = 0.f, ws = 0.f;
      if (ok) {
        w = nan_to_num_f(a_t * T);
        ws = nan_to_num_f(a_s * Ts);
        out.weights[row + s] = w;
        s_acc += w;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s_rgb[c] += T * (a_s * cs[c] + mixd[c]);      // rendering_starmip.py:132-136
          s_rgbs[c] += Ts * a_s * cs[c];                 // :142
        }
        ent += mip_bin_entropy_term(a_s) + ent_d;
        const float tot = a_s + a_dsum;
        const float tc = fmaxf(tot, STAR_EPS_F32);
        float E;
        {
          const float p = fmaxf(fdiv(a_s, tc), STAR_EPS_F32);
          E = p * flog(p);
        }
        const float sc = fmaxf(sig_tot, STAR_EPS_F32);
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          const float p = fmaxf(fdiv(a_dv[v], tc), STAR_EPS_F32);
          E += p * flog(p);
          const float n = fdiv(sig_dv[v], sc);
          rayreg += n * n;       // compute_ray_reg on [R,V,S,1]: the max runs over the singleton (see oracle)
        }
        dvs += tot * E;
      }
      median_step(w, ok, s, med, lane);
      median_step(ws, ok, s, med_s, lane);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { s_rgb[c] = warp_sum(s_rgb[c]); s_rgbs[c] = warp_sum(s_rgbs[c]); }
    s_acc = warp_sum(s_acc); ent = warp_sum(ent); dvs = warp_sum(dvs); dyn = warp_sum(dyn); rayreg = warp_sum(rayreg);
#pragma unroll
    for (int v = 0; v < VT; ++v)
#pragma unroll
      for (int c = 0; c < 3; ++c) s_rgbd[v][c] = warp_sum(s_rgbd[v][c]);
    if (lane == 0) {
      out.acc[r] = s_acc;
      out.depth[r] = median_depth(br, med.idx, S);
      out.depth_static[r] = median_depth(br, med_s.idx, S);
#pragma unroll
      for (int c = 0; c < 3; ++c) { out.rgb[r * 3 + c] = s_rgb[c]; out.rgb_static[r * 3 + c] = s_rgbs[c]; }
#pragma unroll
      for (int v = 0; v < VT; ++v) {
#pragma unroll
        for (int c = 0; c < 3; ++c) out.rgb_dynamic[((int64_t)r * V + v) * 3 + c] = s_rgbd[v][c];
        out.depth_dynamic[(int64_t)r * V + v] = median_depth(br, med_d[v].idx, S);
      }
      const float inc = mip_chunk_inv_count(r, R, chunk);
      const float invS = 1.f / (float)S;
      reg_acc[0] += -ent * inc * invS / (float)(V + 1);
      reg_acc[1] += -dvs * inc * invS;
      reg_acc[2] += rayreg * inc / (float)V;
      reg_acc[4] += dyn * inc * invS / (float)V;
    }
  }
  __shared__ float s_part[5][4];
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 5; ++k) s_part[k][warp] = reg_acc[k];
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int w = 0; w < wpb; ++w) t += s_part[threadIdx.x][w];
    reg_partial[threadIdx.x * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void mip_reg_finalize_kernel(const float* __restrict__ reg_partial, int nblocks, float* __restrict__ regs) {
  const int k = threadIdx.x;
  if (k < 5) {
    float t = 0.f;
    for (int b = 0; b < nblocks; ++b) t += reg_partial[k * nblocks + b];
    regs[k] = t;
  }
}

// Differentiated outputs: rgb, acc, weights and the regularisers (the per-field products and the median depths are
// declared non-differentiable by the host wrapper, as in the vanilla path).  smem per warp: T[S]
template <int VT>
__global__ void __launch_bounds__(128)
mip_composite_multi_bwd_kernel(const float* __restrict__ raw_sigma_s, const float* __restrict__ raw_rgb_s,
                               const float* __restrict__ raw_sigma_d, const float* __restrict__ raw_rgb_d,
                               const float* __restrict__ bins, int R, int V, int S, int chunk,
                               const float* __restrict__ g_rgb, const float* __restrict__ g_acc,
                               const float* __restrict__ g_weights, const float* __restrict__ g_regs,
                               float* __restrict__ d_raw_sigma_s, float* __restrict__ d_raw_rgb_s,
                               float* __restrict__ d_raw_sigma_d, float* __restrict__ d_raw_rgb_d) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* s_T = smem + (size_t)warp * S;
  float greg[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (g_regs != nullptr)
#pragma unroll
    for (int k = 0; k < 5; ++k) greg[k] = g_regs[k];
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float* br = bins + (int64_t)r * (S + 1);
    const int64_t row = (int64_t)r * S;
    float cT = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float dd = 0.f;
      if (ok) {
        float sig = softplus_f(raw_sigma_s[row + s]);
#pragma unroll
        for (int v = 0; v < VT; ++v) sig += softplus_f(MRAD(r, v, s));
        dd = (br[s + 1] - br[s]) * sig;
      }
      const float T = fexp(-excl_scan_sum(dd, cT, lane));
      if (ok) s_T[s] = T;
    }
    const float gr = g_rgb ? g_rgb[r * 3 + 0] : 0.f, gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f,
                gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
    const float gA = g_acc ? g_acc[r] : 0.f;
    const float inc = mip_chunk_inv_count(r, R, chunk);
    const float invS = 1.f / (float)S;
    const float k_ent = -greg[0] * inc * invS / (float)(V + 1);
    const float k_dvs = -greg[1] * inc * invS;
    const float k_ray = greg[2] * inc / (float)V;
    const float k_dyn = greg[4] * inc * invS / (float)V;
    __syncwarp();
    float suffix = 0.f;
    for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float T = 0.f, delta = 0.f, raw_s = 0.f, sig_s = 0.f, a_s = 0.f, a_t = 0.f, sig_tot = 0.f, Gw = 0.f, val = 0.f;
      float cs[3] = {0, 0, 0};
      float a_dv[VT], sig_dv[VT];
      float a_dsum = 0.f;
      if (ok) {
        T = s_T[s];
        delta = br[s + 1] - br[s];
        raw_s = raw_sigma_s[row + s];
        sig_s = softplus_f(raw_s);
        a_s = 1.f - fexp(-delta * sig_s);
#pragma unroll
        for (int c = 0; c < 3; ++c) cs[c] = sigmoid_f(raw_rgb_s[(row + s) * 3 + c]);
        sig_tot = sig_s;
        float B0 = a_s * cs[0], B1 = a_s * cs[1], B2 = a_s * cs[2];
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          const float sg = softplus_f(MRAD(r, v, s));
          const float a_d = 1.f - fexp(-delta * sg);
          sig_dv[v] = sg; a_dv[v] = a_d;
          sig_tot += sg; a_dsum += a_d;
          B0 += a_d * sigmoid_f(MRCD(r, v, s, 0));
          B1 += a_d * sigmoid_f(MRCD(r, v, s, 1));
          B2 += a_d * sigmoid_f(MRCD(r, v, s, 2));
        }
        a_t = 1.f - fexp(-delta * sig_tot);
        Gw = gA + (g_weights ? g_weights[row + s] : 0.f);
        val = T * (gr * B0 + gg * B1 + gb * B2 + a_t * Gw);
      }
      const float incl = warp_rscan_sum(val, lane);
      const float after = incl - val + suffix;
      suffix += __shfl_sync(STAR_FULL_MASK, incl, 0);
      if (ok) {
        // gradient w.r.t. the total density (shared by every field)
        float d_sig_all = delta * (T * Gw * (1.f - a_t) - after);
        // regularisers
        const float tot = a_s + a_dsum;
        const float tc = fmaxf(tot, STAR_EPS_F32);
        const float tot_ok = (tot >= STAR_EPS_F32) ? 1.f : 0.f;
        const float sc = fmaxf(sig_tot, STAR_EPS_F32);
        float E = 0.f, Csum = 0.f, rsum = 0.f;
        {
          const float q = fdiv(a_s, tc), p = fmaxf(q, STAR_EPS_F32);
          E += p * flog(p);
          if (q >= STAR_EPS_F32) Csum += (flog(p) + 1.f) * a_s;
        }
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          const float q = fdiv(a_dv[v], tc), p = fmaxf(q, STAR_EPS_F32);
          E += p * flog(p);
          if (q >= STAR_EPS_F32) Csum += (flog(p) + 1.f) * a_dv[v];
          const float n = fdiv(sig_dv[v], sc);
          rsum += n * n;
        }
        const float reg_common = E - tot * tot_ok * fdiv(Csum, tc * tc);
        if (sig_tot >= STAR_EPS_F32) d_sig_all += -k_ray * 2.f * fdiv(rsum, sc);   // d/d sigma_tot of sum_v (sigma_v/sc)^2
        {
          float d_as = T * (gr * cs[0] + gg * cs[1] + gb * cs[2]);
          const float q = fdiv(a_s, tc), p = fmaxf(q, STAR_EPS_F32);
          d_as += k_dvs * (reg_common + (q >= STAR_EPS_F32 ? fdiv(tot * (flog(p) + 1.f), tc) : 0.f));
          const float c = fminf(fmaxf(a_s, STAR_EPS_F32), 1.f - STAR_EPS_F32);
          d_as += k_ent * (flog(c) - flog1m(c));
          const float d_sig = d_as * delta * (1.f - a_s) + d_sig_all;
          d_raw_sigma_s[row + s] = d_sig * softplus_grad_f(raw_s);
          float* o = d_raw_rgb_s + (row + s) * 3;
          const float k = T * a_s;
          o[0] = k * gr * cs[0] * (1.f - cs[0]);
          o[1] = k * gg * cs[1] * (1.f - cs[1]);
          o[2] = k * gb * cs[2] * (1.f - cs[2]);
        }
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          const float rd = MRAD(r, v, s);
          const float a_d = a_dv[v];
          const float c0 = sigmoid_f(MRCD(r, v, s, 0)), c1 = sigmoid_f(MRCD(r, v, s, 1)), c2 = sigmoid_f(MRCD(r, v, s, 2));
          float d_ad = T * (gr * c0 + gg * c1 + gb * c2);
          const float q = fdiv(a_d, tc), p = fmaxf(q, STAR_EPS_F32);
          d_ad += k_dvs * (reg_common + (q >= STAR_EPS_F32 ? fdiv(tot * (flog(p) + 1.f), tc) : 0.f));
          const float c = fminf(fmaxf(a_d, STAR_EPS_F32), 1.f - STAR_EPS_F32);
          d_ad += k_ent * (flog(c) - flog1m(c));
          const float d_sig = d_ad * delta * (1.f - a_d) + d_sig_all + k_dyn + k_ray * 2.f * fdiv(sig_dv[v], sc * sc);
          const int64_t o1 = ((int64_t)r * V + v) * S + s;
          d_raw_sigma_d[o1] = d_sig * softplus_grad_f(rd);
          const float k = T * a_d;
          d_raw_rgb_d[o1 * 3 + 0] = k * gr * c0 * (1.f - c0);
          d_raw_rgb_d[o1 * 3 + 1] = k * gg * c1 * (1.f - c1);
          d_raw_rgb_d[o1 * 3 + 2] = k * gb * c2 * (1.f - c2);
        }
      }
    }
    __syncwarp();
  }
}

// =========================================================================== host entry points
static void mip_warp_per_ray_cfg(int R, size_t smem_per_warp, int& blocks, int& threads, size_t& smem) {
  int wpb = 4;
  threads = wpb * 32;
  smem = smem_per_warp * wpb;
  int64_t b = ((int64_t)R + wpb - 1) / wpb;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  blocks = (int)b;
}

extern "C" int star_mip_composite_single_forward(const float* raw_sigma, const float* raw_rgb, const float* bins, int R,
                                                 int S, float* rgb, float* acc, float* depth, float* weights,
                                                 void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!raw_sigma || !raw_rgb || !bins || !rgb || !acc || !depth || !weights) return STAR_E_NULL;
  if (R < 0 || S < 1) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  mip_warp_per_ray_cfg(R, 0, blocks, threads, smem);
  mip_composite_single_fwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(raw_sigma, raw_rgb, bins, R, S, rgb, acc,
                                                                                 depth, weights);
  return star_check_launch();
}

extern "C" int star_mip_composite_single_backward(const float* raw_sigma, const float* raw_rgb, const float* bins, int R,
                                                  int S, const float* g_rgb, const float* g_acc, const float* g_weights,
                                                  float* d_raw_sigma, float* d_raw_rgb, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!raw_sigma || !raw_rgb || !bins || !d_raw_sigma || !d_raw_rgb) return STAR_E_NULL;
  if (R < 0 || S < 1 || S > 12288) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  mip_warp_per_ray_cfg(R, sizeof(float) * S, blocks, threads, smem);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(mip_composite_single_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mip_composite_single_bwd_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>(
      raw_sigma, raw_rgb, bins, R, S, g_rgb, g_acc, g_weights, d_raw_sigma, d_raw_rgb);
  return star_check_launch();
}

static int mip_multi_blocks(int R) {
  int64_t b = ((int64_t)R + 3) / 4;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" size_t star_mip_composite_multi_ws_bytes(int R) { return sizeof(float) * 5 * (size_t)mip_multi_blocks(R); }

extern "C" int star_mip_composite_multi_forward(const float* raw_sigma_s, const float* raw_rgb_s,
                                                const float* raw_sigma_d, const float* raw_rgb_d, const float* bins,
                                                int R, int V, int S, int chunk, const StarMipMultiOut* out,
                                                void* workspace, void* stream) {
  if (!raw_sigma_s || !raw_rgb_s || !raw_sigma_d || !raw_rgb_d || !bins || !out || !workspace) return STAR_E_NULL;
  if (!out->rgb || !out->acc || !out->depth || !out->weights || !out->rgb_static || !out->depth_static ||
      !out->rgb_dynamic || !out->depth_dynamic || !out->dynamic_transmittance || !out->regs)
    return STAR_E_NULL;
  if (R < 1 || S < 1 || V < 1 || V > STAR_MAX_V || chunk < 1) return STAR_E_BAD_SHAPE;
  const int blocks = mip_multi_blocks(R);
#define STAR_MIP_FWD(VT)                                                                                        \
  case VT:                                                                                                      \
    mip_composite_multi_fwd_kernel<VT><<<blocks, 128, 0, (cudaStream_t)stream>>>(                               \
        raw_sigma_s, raw_rgb_s, raw_sigma_d, raw_rgb_d, bins, R, V, S, chunk, *out, (float*)workspace);         \
    break;
  switch (V) {
    STAR_MIP_FWD(1) STAR_MIP_FWD(2) STAR_MIP_FWD(3) STAR_MIP_FWD(4)
    STAR_MIP_FWD(5) STAR_MIP_FWD(6) STAR_MIP_FWD(7) STAR_MIP_FWD(8)
  }
#undef STAR_MIP_FWD
  int rc = star_check_launch();
  if (rc) return rc;
  mip_reg_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float*)workspace, blocks, out->regs);
  return star_check_launch();
}

extern "C" int star_mip_composite_multi_backward(const float* raw_sigma_s, const float* raw_rgb_s,
                                                 const float* raw_sigma_d, const float* raw_rgb_d, const float* bins,
                                                 int R, int V, int S, int chunk, const float* g_rgb, const float* g_acc,
                                                 const float* g_weights, const float* g_regs, float* d_raw_sigma_s,
                                                 float* d_raw_rgb_s, float* d_raw_sigma_d, float* d_raw_rgb_d,
                                                 void* stream) {
  if (!raw_sigma_s || !raw_rgb_s || !raw_sigma_d || !raw_rgb_d || !bins || !d_raw_sigma_s || !d_raw_rgb_s ||
      !d_raw_sigma_d || !d_raw_rgb_d)
    return STAR_E_NULL;
  if (R < 1 || S < 1 || S > 12288 || V < 1 || V > STAR_MAX_V || chunk < 1) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  mip_warp_per_ray_cfg(R, sizeof(float) * S, blocks, threads, smem);
#define STAR_MIP_BWD(VT)                                                                                        \
  case VT:                                                                                                      \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(mip_composite_multi_bwd_kernel<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                           (int)smem);                                                                          \
    mip_composite_multi_bwd_kernel<VT><<<blocks, threads, smem, (cudaStream_t)stream>>>(                        \
        raw_sigma_s, raw_rgb_s, raw_sigma_d, raw_rgb_d, bins, R, V, S, chunk, g_rgb, g_acc, g_weights, g_regs,  \
        d_raw_sigma_s, d_raw_rgb_s, d_raw_sigma_d, d_raw_rgb_d);                                                \
    break;
  switch (V) {
    STAR_MIP_BWD(1) STAR_MIP_BWD(2) STAR_MIP_BWD(3) STAR_MIP_BWD(4)
    STAR_MIP_BWD(5) STAR_MIP_BWD(6) STAR_MIP_BWD(7) STAR_MIP_BWD(8)
  }
#undef STAR_MIP_BWD
  return star_check_launch();
}
