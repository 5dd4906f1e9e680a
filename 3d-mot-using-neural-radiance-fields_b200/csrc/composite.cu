// Alpha compositing kernels: single field (raw2outputs, models/rendering__.py:307-379) and
// static + V dynamic fields with the five STaR regularisers (raw2outputs_star :383-576, :612-715),
// forward and backward.  One warp per ray; lanes stride over samples in chunks of 32; the exclusive
// transmittance product is a warp-shuffle prefix product with a running carry, the backward suffix
// sums are warp-shuffle reverse scans.  HBM-bound: every input is read once (forward) or twice
// (backward: forward recompute + reverse sweep, the second read hits L1/L2).
#include "ray_device.cuh"


// =========================================================================== single field forward
__global__ void __launch_bounds__(128, 16) composite_single_fwd_kernel(const float* __restrict__ raw_alpha, const float* __restrict__ raw_rgb,
                                            const float* __restrict__ z_vals, const float* __restrict__ rays_d,
                                            int R, int S, float far_dist, int white_bkgd, float* __restrict__ rgb_o,
                                            float* __restrict__ disp_o, float* __restrict__ acc_o,
                                            float* __restrict__ depth_o, float* __restrict__ weights_o,
                                            float* __restrict__ dists_o) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    const float norm = ray_norm(rays_d, r);
    const float* zr = z_vals + (int64_t)r * S;
    const float* ar = raw_alpha + (int64_t)r * S;
    const float* cr = raw_rgb + (int64_t)r * S * 3;
    float carry = 1.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float alpha = 0.f, z = 0.f, dist = 0.f;
      if (ok) {
        z = zr[s];
        dist = sample_dist(zr, s, S, far_dist, norm);
        alpha = alpha_of(ar[s], dist);
      }
      const float m = ok ? (1.f - alpha + 1e-10f) : 1.f;
      const float T = excl_transmittance(m, carry, lane);
      if (ok) {
        const float w = alpha * T;
        weights_o[(int64_t)r * S + s] = w;
        if (dists_o) dists_o[(int64_t)r * S + s] = dist;
        sr += w * sigmoid_f(cr[s * 3 + 0]);
        sg += w * sigmoid_f(cr[s * 3 + 1]);
        sb += w * sigmoid_f(cr[s * 3 + 2]);
        sd += w * z;
        sa += w;
      }
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
    if (lane == 0) {
      const float wsum = (sa >= 0.f) ? sa : 1e-7f;                  // :353-354
      disp_o[r] = 1.f / max_nan_f(1e-10f, sd / wsum);                   // :355-357
      acc_o[r] = sa;
      depth_o[r] = sd;
      const float bg = white_bkgd ? (1.f - sa) : 0.f;               // :360-361
      rgb_o[r * 3 + 0] = sr + bg;
      rgb_o[r * 3 + 1] = sg + bg;
      rgb_o[r * 3 + 2] = sb + bg;
    }
  }
}

// Even S: every lane owns two consecutive samples (composite_single_ray_x2, ray_device.cuh).
__global__ void __launch_bounds__(128, 16) composite_single_fwd_x2_kernel(const float* __restrict__ raw_alpha, const float* __restrict__ raw_rgb,
                                               const float* __restrict__ z_vals, const float* __restrict__ rays_d,
                                               int R, int S, float far_dist, int white_bkgd,
                                               float* __restrict__ rgb_o, float* __restrict__ disp_o,
                                               float* __restrict__ acc_o, float* __restrict__ depth_o,
                                               float* __restrict__ weights_o, float* __restrict__ dists_o) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    composite_single_ray_x2(raw_alpha, raw_rgb, z_vals, rays_d, r, S, far_dist, white_bkgd, rgb_o, disp_o, acc_o, depth_o,
                            weights_o, dists_o, nullptr, nullptr, lane);
  }
}

// gradients of disp = 1/max(1e-10, depth/where(A>=0,A,eps)) folded into depth / acc gradients
__device__ __forceinline__ void fold_disp_grad(float g_disp, float depth, float A, float eps_guard, float& g_depth,
                                               float& g_A) {
  if (g_disp == 0.f) return;
  const bool a_ok = (A >= 0.f);
  const float Ap = a_ok ? A : eps_guard;
  const float q = depth / Ap;
  if (q > 1e-10f) {
    const float dq = -g_disp / (q * q);
    g_depth += dq / Ap;
    if (a_ok) g_A += -dq * depth / (Ap * Ap);
  }
}

// =========================================================================== single field backward
// smem per warp: alpha[S] | T[S]
__global__ void composite_single_bwd_kernel(const float* __restrict__ raw_alpha, const float* __restrict__ raw_rgb,
                                            const float* __restrict__ z_vals, const float* __restrict__ rays_d,
                                            int R, int S, float far_dist, int white_bkgd,
                                            const float* __restrict__ g_rgb, const float* __restrict__ g_disp,
                                            const float* __restrict__ g_acc, const float* __restrict__ g_depth,
                                            const float* __restrict__ g_weights, float* __restrict__ d_raw_alpha,
                                            float* __restrict__ d_raw_rgb) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* s_alpha = smem + (size_t)warp * 2 * S;
  float* s_T = s_alpha + S;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float norm = ray_norm(rays_d, r);
    const float* zr = z_vals + (int64_t)r * S;
    const float* ar = raw_alpha + (int64_t)r * S;
    const float* cr = raw_rgb + (int64_t)r * S * 3;
    float carry = 1.f, sd = 0.f, sa = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float alpha = 0.f;
      if (ok) alpha = alpha_of(ar[s], sample_dist(zr, s, S, far_dist, norm));
      const float m = ok ? (1.f - alpha + 1e-10f) : 1.f;
      const float T = excl_transmittance(m, carry, lane);
      if (ok) {
        s_alpha[s] = alpha;
        s_T[s] = T;
        sd += alpha * T * zr[s];
        sa += alpha * T;
      }
    }
    sd = warp_sum(sd);
    sa = warp_sum(sa);
    const float gr = g_rgb ? g_rgb[r * 3 + 0] : 0.f, gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f,
                gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
    float gD = g_depth ? g_depth[r] : 0.f;
    float gA = (g_acc ? g_acc[r] : 0.f) - (white_bkgd ? (gr + gg + gb) : 0.f);
    fold_disp_grad(g_disp ? g_disp[r] : 0.f, sd, sa, 1e-7f, gD, gA);
    __syncwarp();
    float suffix = 0.f;  // sum_{k > current chunk} w_k G_k
    for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float v = 0.f, G = 0.f, alpha = 0.f, T = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (ok) {
        alpha = s_alpha[s];
        T = s_T[s];
        c0 = sigmoid_f(cr[s * 3 + 0]); c1 = sigmoid_f(cr[s * 3 + 1]); c2 = sigmoid_f(cr[s * 3 + 2]);
        G = gr * c0 + gg * c1 + gb * c2 + gD * zr[s] + gA + (g_weights ? g_weights[(int64_t)r * S + s] : 0.f);
        v = alpha * T * G;
      }
      const float incl = warp_rscan_sum(v, lane);
      const float after = incl - v + suffix;
      suffix += __shfl_sync(STAR_FULL_MASK, incl, 0);
      if (ok) {
        const float m = 1.f - alpha + 1e-10f;
        const float d_alpha = T * G - fdiv(after, m);
        const float raw = ar[s];
        d_raw_alpha[(int64_t)r * S + s] = d_alpha * dalpha_draw(raw, sample_dist(zr, s, S, far_dist, norm));
        const float w = alpha * T;
        float* o = d_raw_rgb + ((int64_t)r * S + s) * 3;
        o[0] = w * gr * c0 * (1.f - c0);
        o[1] = w * gg * c1 * (1.f - c1);
        o[2] = w * gb * c2 * (1.f - c2);
      }
    }
    __syncwarp();
  }
}

// =========================================================================== multi field
// Per-sample quantities of the static + V dynamic fields.
struct RegRayAcc {
  float ent, dvs, rayreg, dyn;  // per-ray sums (see below)
  float Z, clogc, sig_s_sum;    // static_reg pieces
};

__device__ __forceinline__ float bin_entropy_term(float a) {  // alpha*log(clamp) + (1-alpha)*log1p(-clamp) (:620-628)
  const float c = fminf(fmaxf(a, STAR_EPS_F32), 1.f - STAR_EPS_F32);
  return a * flog(c) + (1.f - a) * flog1m(c);
}

// Layout helpers for the reference [R,V,S] / [R,V,S,3] dynamic tensors
#define RAD(r, v, s) raw_alpha_d[((int64_t)(r) * V + (v)) * S + (s)]
#define RCD(r, v, s, c) raw_rgb_d[(((int64_t)(r) * V + (v)) * S + (s)) * 3 + (c)]

__device__ __forceinline__ float chunk_inv_count(int r, int R, int chunk) {
  // number of rays in the ray chunk that holds ray r (star__.py:84-85)
  const int c0 = (r / chunk) * chunk;
  const int n = min(R, c0 + chunk) - c0;
  return 1.f / (float)n;
}

#ifndef STAR_MULTI_FWD_MINBLOCKS
#define STAR_MULTI_FWD_MINBLOCKS 4      // 128 registers: 16 warps per SM instead of 12 (V = 5: 36 bytes of spills)
#endif
template <int VT>
__global__ void __launch_bounds__(128, STAR_MULTI_FWD_MINBLOCKS) composite_multi_fwd_kernel(const float* __restrict__ raw_alpha_s, const float* __restrict__ raw_rgb_s,
                                           const float* __restrict__ raw_alpha_d, const float* __restrict__ raw_rgb_d,
                                           const float* __restrict__ z_vals, const float* __restrict__ rays_d, int R,
                                           int /*V == VT*/, int S, float far_dist, int white_bkgd, int chunk, StarMultiOut out,
                                           float* __restrict__ reg_partial) {
  constexpr int V = VT;     // one instantiation per object count: the per-object loops and the index arithmetic are static
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float reg_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float norm = ray_norm(rays_d, r);
    const float* zr = z_vals + (int64_t)r * S;
    float cT = 1.f, cTs = 1.f, cTall = 1.f;
    float cTd[VT];
    float s_rgb[3] = {0, 0, 0}, s_rgbs[3] = {0, 0, 0}, s_all[3] = {0, 0, 0};
    float s_depth = 0.f, s_acc = 0.f, s_depth_s = 0.f;
    float s_rgbd[VT][3], s_depth_d[VT], mx[VT];
#pragma unroll
    for (int v = 0; v < VT; ++v) {
      cTd[v] = 1.f; s_depth_d[v] = 0.f; mx[v] = -1.f;
      s_rgbd[v][0] = s_rgbd[v][1] = s_rgbd[v][2] = 0.f;
    }
    float ent = 0.f, dvs = 0.f, dyn = 0.f, Z = 0.f, clogc = 0.f, sig_s_sum = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float z = 0.f, dist = 0.f, ra_s = 0.f, a_s = 0.f, sig_s = 0.f, raw_sum_d = 0.f;
      float cs[3] = {0, 0, 0};
      // Every load of this trip first: the per-field warp scans below are convergence points the compiler does not move
      // loads across, and with the loads next to their uses a trip made V + 1 serialised DRAM round trips
      // (profiles/r2j: 0.13 of HBM peak at 12 warps per SM).
      float in_rad[VT], in_rcd[VT][3];
      if (ok) {
        z = zr[s];
        const float znext = (s == S - 1) ? 0.f : zr[s + 1];
        ra_s = raw_alpha_s[(int64_t)r * S + s];
#pragma unroll
        for (int c = 0; c < 3; ++c) cs[c] = raw_rgb_s[((int64_t)r * S + s) * 3 + c];
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          if (v < V) {
            in_rad[v] = RAD(r, v, s);
#pragma unroll
            for (int c = 0; c < 3; ++c) in_rcd[v][c] = RCD(r, v, s, c);
          }
        }
        dist = ((s == S - 1) ? far_dist : (znext - z)) * norm;    // sample_dist
        a_s = alpha_of(ra_s, dist);
        sig_s = softplus_f(ra_s);
#pragma unroll
        for (int c = 0; c < 3; ++c) cs[c] = sigmoid_f(cs[c]);
      }
      // dynamic fields
      float mixd[3] = {0, 0, 0};
      float a_dsum = 0.f, sig_dsum = 0.f, ent_d = 0.f;
      float a_dv[VT], sig_dv[VT];
#pragma unroll
      for (int v = 0; v < VT; ++v) {
        if (v < V) {
          float a_d = 0.f, sg = 0.f, cd[3] = {0, 0, 0};
          if (ok) {
            const float rd = in_rad[v];
            raw_sum_d += rd;
            a_d = alpha_of(rd, dist);
            sg = softplus_f(rd);
#pragma unroll
            for (int c = 0; c < 3; ++c) cd[c] = sigmoid_f(in_rcd[v][c]);
          }
          a_dv[v] = a_d;
          sig_dv[v] = sg;
          const float m = ok ? (1.f - a_d + 1e-10f) : 1.f;
          const float Td = excl_transmittance(m, cTd[v], lane);
          if (ok) {
            const float wd = Td * a_d;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              s_rgbd[v][c] += wd * cd[c];
              mixd[c] += a_d * cd[c];
            }
            s_depth_d[v] += wd * z;
            // dynamic_transmittance = T_d[..., -1]: exclusive product at the last sample (:568)
            if (s == S - 1) out.dynamic_transmittance[(int64_t)r * V + v] = Td;
            a_dsum += a_d;
            sig_dsum += sg;
            ent_d += bin_entropy_term(a_d);
            dyn += sg;
          }
        }
      }
      const float a_t = ok ? alpha_of(ra_s + raw_sum_d, dist) : 0.f;   // :416-418
      const float a_all = ok ? alpha_of(raw_sum_d, dist) : 0.f;        // :535-537
      const float T = excl_transmittance(ok ? (1.f - a_t + 1e-10f) : 1.f, cT, lane);
      const float Ts = excl_transmittance(ok ? (1.f - a_s + 1e-10f) : 1.f, cTs, lane);
      const float Tall = excl_transmittance(ok ? (1.f - a_all + 1e-10f) : 1.f, cTall, lane);
      if (ok) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s_rgb[c] += T * (a_s * cs[c] + mixd[c]);                    // :456-463
          s_rgbs[c] += Ts * a_s * cs[c];                               // :482-484
          s_all[c] += Tall * mixd[c];                                  // :552-555
        }
        const float w = T * a_t;                                       // :503
        out.weights[(int64_t)r * S + s] = w;
        s_depth += w * z;
        s_acc += w;
        s_depth_s += Ts * a_s * z;
        // regularisers
        ent += bin_entropy_term(a_s) + ent_d;
        const float tot = a_s + a_dsum;
        const float tc = fmaxf(tot, STAR_EPS_F32);
        float E;
        {
          const float p = fmaxf(fdiv(a_s, tc), STAR_EPS_F32);
          E = p * flog(p);
        }
        const float sc = fmaxf(sig_s + sig_dsum, STAR_EPS_F32);
#pragma unroll
        for (int v = 0; v < VT; ++v) {
          if (v < V) {
            const float p = fmaxf(fdiv(a_dv[v], tc), STAR_EPS_F32);
            E += p * flog(p);
            mx[v] = fmaxf(mx[v], fdiv(sig_dv[v], sc));
          }
        }
        dvs += tot * E;
        const float cc = fminf(fmaxf(a_s, STAR_EPS_F32), 1.f - STAR_EPS_F32);
        Z += cc;
        clogc += cc * flog(cc);
        sig_s_sum += sig_s;
      }
    }
    // per-ray reductions
#pragma unroll
    for (int c = 0; c < 3; ++c) { s_rgb[c] = warp_sum(s_rgb[c]); s_rgbs[c] = warp_sum(s_rgbs[c]); s_all[c] = warp_sum(s_all[c]); }
    s_depth = warp_sum(s_depth); s_acc = warp_sum(s_acc); s_depth_s = warp_sum(s_depth_s);
    ent = warp_sum(ent); dvs = warp_sum(dvs); dyn = warp_sum(dyn);
    Z = warp_sum(Z); clogc = warp_sum(clogc); sig_s_sum = warp_sum(sig_s_sum);
    float rayreg = 0.f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
      if (v < V) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s_rgbd[v][c] = warp_sum(s_rgbd[v][c]);
        s_depth_d[v] = warp_sum(s_depth_d[v]);
        const float M = warp_max(mx[v]);
        rayreg += M * M;
      }
    }
    if (lane == 0) {
      const float wsum = (s_acc >= 0.f) ? s_acc : STAR_EPS_F32;        // :509-511
      out.disp[r] = 1.f / max_nan_f(1e-10f, s_depth / wsum);
      out.acc[r] = s_acc;
      out.depth[r] = s_depth;
      const float bg = white_bkgd ? (1.f - s_acc) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        out.rgb[r * 3 + c] = s_rgb[c] + bg;
        out.rgb_static[r * 3 + c] = s_rgbs[c];
        if (out.rgb_dynamic_all) out.rgb_dynamic_all[r * 3 + c] = s_all[c];
      }
      out.depth_static[r] = s_depth_s;
#pragma unroll
      for (int v = 0; v < VT; ++v) {
        if (v < V) {
#pragma unroll
          for (int c = 0; c < 3; ++c) out.rgb_dynamic[((int64_t)r * V + v) * 3 + c] = s_rgbd[v][c];
          out.depth_dynamic[(int64_t)r * V + v] = s_depth_d[v];
        }
      }
      const float inc = chunk_inv_count(r, R, chunk);
      const float invS = 1.f / (float)S;
      reg_acc[0] += -ent * inc * invS / (float)(V + 1);                 // :620-629
      reg_acc[1] += -dvs * inc * invS;                                  // :645-651
      reg_acc[2] += rayreg * inc / (float)V;                            // :690-693
      const float pl = clogc / Z - flog(Z);                             // sum_s p log p, p = c/Z
      reg_acc[3] += (sig_s_sum < 0.1f ? 0.f : 1.f) * (-pl * invS) * inc;  // :707-709
      reg_acc[4] += dyn * inc * invS / (float)V;                        // :715
    }
  }
  // block-level partials -> workspace (deterministic two-stage reduction)
  __shared__ float s_part[5][32];
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

__global__ void reg_finalize_kernel(const float* __restrict__ reg_partial, int nblocks, float* __restrict__ regs) {
  const int k = threadIdx.x;
  if (k < 5) {
    float t = 0.f;
    for (int b = 0; b < nblocks; ++b) t += reg_partial[k * nblocks + b];
    regs[k] = t;
  }
}

// =========================================================================== multi field backward
// Appendix B of SURVEY.md.  Differentiated outputs: rgb, disp, acc, depth, weights and the five
// regularisers (what the reference training losses read, train_online__.py:158-273).  The per-field
// visualisation products (rgb_static, rgb_dynamic, depth_static, depth_dynamic,
// dynamic_transmittance, rgb_dynamic_all) are declared non-differentiable by the host wrapper.
// smem per warp: T[S] (total transmittance).
__global__ void composite_multi_bwd_kernel(const float* __restrict__ raw_alpha_s, const float* __restrict__ raw_rgb_s,
                                           const float* __restrict__ raw_alpha_d, const float* __restrict__ raw_rgb_d,
                                           const float* __restrict__ z_vals, const float* __restrict__ rays_d, int R,
                                           int V, int S, float far_dist, int white_bkgd, int chunk,
                                           const float* __restrict__ g_rgb, const float* __restrict__ g_disp,
                                           const float* __restrict__ g_acc, const float* __restrict__ g_depth,
                                           const float* __restrict__ g_weights, const float* __restrict__ g_regs,
                                           float* __restrict__ d_raw_alpha_s, float* __restrict__ d_raw_rgb_s,
                                           float* __restrict__ d_raw_alpha_d, float* __restrict__ d_raw_rgb_d) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* s_T = smem + (size_t)warp * S;
  float greg[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (g_regs != nullptr)
#pragma unroll
    for (int k = 0; k < 5; ++k) greg[k] = g_regs[k];
  const bool need_regs = (greg[0] != 0.f) || (greg[1] != 0.f) || (greg[2] != 0.f) || (greg[3] != 0.f) || (greg[4] != 0.f);

  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float norm = ray_norm(rays_d, r);
    const float* zr = z_vals + (int64_t)r * S;
    // ---------------- pass 1: forward recompute (T, depth, acc, per-ray regulariser statistics)
    float cT = 1.f, s_depth = 0.f, s_acc = 0.f, Z = 0.f, clogc = 0.f, sig_s_sum = 0.f;
    float mx[STAR_MAX_V];
    int amx[STAR_MAX_V];
#pragma unroll
    for (int v = 0; v < STAR_MAX_V; ++v) { mx[v] = -1.f; amx[v] = 0x7fffffff; }
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float a_t = 0.f;
      if (ok) {
        const float dist = sample_dist(zr, s, S, far_dist, norm);
        const float ra_s = raw_alpha_s[(int64_t)r * S + s];
        float raw_sum = ra_s;
        float sig_sum = softplus_f(ra_s);
        float sig_dv[STAR_MAX_V];
#pragma unroll
        for (int v = 0; v < STAR_MAX_V; ++v)
          if (v < V) {
            const float rd = RAD(r, v, s);
            raw_sum += rd;
            sig_dv[v] = softplus_f(rd);
            sig_sum += sig_dv[v];
          }
        a_t = alpha_of(raw_sum, dist);
        if (need_regs) {
          const float a_s = alpha_of(ra_s, dist);
          const float cc = fminf(fmaxf(a_s, STAR_EPS_F32), 1.f - STAR_EPS_F32);
          Z += cc;
          clogc += cc * flog(cc);
          sig_s_sum += softplus_f(ra_s);
          const float sc = fmaxf(sig_sum, STAR_EPS_F32);
#pragma unroll
          for (int v = 0; v < STAR_MAX_V; ++v)
            if (v < V) {
              const float n = fdiv(sig_dv[v], sc);
              if (n > mx[v]) { mx[v] = n; amx[v] = s; }
            }
        }
      }
      const float T = excl_transmittance(ok ? (1.f - a_t + 1e-10f) : 1.f, cT, lane);
      if (ok) {
        s_T[s] = T;
        s_depth += T * a_t * zr[s];
        s_acc += T * a_t;
      }
    }
    s_depth = warp_sum(s_depth);
    s_acc = warp_sum(s_acc);
    float PL = 0.f, mask = 0.f;
    if (need_regs) {
      Z = warp_sum(Z); clogc = warp_sum(clogc); sig_s_sum = warp_sum(sig_s_sum);
      PL = clogc / Z - flog(Z);
      mask = sig_s_sum < 0.1f ? 0.f : 1.f;
#pragma unroll
      for (int v = 0; v < STAR_MAX_V; ++v)
        if (v < V) {
          // warp arg-max, ties -> smallest sample index
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(STAR_FULL_MASK, mx[v], o);
            const int oa = __shfl_xor_sync(STAR_FULL_MASK, amx[v], o);
            if (om > mx[v] || (om == mx[v] && oa < amx[v])) { mx[v] = om; amx[v] = oa; }
          }
        }
    }
    const float gr = g_rgb ? g_rgb[r * 3 + 0] : 0.f, gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f,
                gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
    float gD = g_depth ? g_depth[r] : 0.f;
    float gA = (g_acc ? g_acc[r] : 0.f) - (white_bkgd ? (gr + gg + gb) : 0.f);
    fold_disp_grad(g_disp ? g_disp[r] : 0.f, s_depth, s_acc, STAR_EPS_F32, gD, gA);
    const float inc = chunk_inv_count(r, R, chunk);
    const float invS = 1.f / (float)S;
    const float k_ent = -greg[0] * inc * invS / (float)(V + 1);
    const float k_dvs = -greg[1] * inc * invS;
    const float k_ray = greg[2] * inc / (float)V;
    const float k_sta = -greg[3] * inc * mask * invS / Z;
    const float k_dyn = greg[4] * inc * invS / (float)V;
    __syncwarp();
    // ---------------- pass 2: reverse sweep
    float suffix = 0.f;
    for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float T = 0.f, dist = 0.f, ra_s = 0.f, a_s = 0.f, a_t = 0.f, raw_tot = 0.f, Gw = 0.f, val = 0.f;
      float cs[3] = {0, 0, 0};
      float a_dv[STAR_MAX_V], sig_dv[STAR_MAX_V];
      float a_dsum = 0.f, sig_dsum = 0.f;
      if (ok) {
        T = s_T[s];
        dist = sample_dist(zr, s, S, far_dist, norm);
        ra_s = raw_alpha_s[(int64_t)r * S + s];
        a_s = alpha_of(ra_s, dist);
#pragma unroll
        for (int c = 0; c < 3; ++c) cs[c] = sigmoid_f(raw_rgb_s[((int64_t)r * S + s) * 3 + c]);
        raw_tot = ra_s;
        float B0 = a_s * cs[0], B1 = a_s * cs[1], B2 = a_s * cs[2];
#pragma unroll
        for (int v = 0; v < STAR_MAX_V; ++v)
          if (v < V) {
            const float rd = RAD(r, v, s);
            raw_tot += rd;
            const float a_d = alpha_of(rd, dist);
            a_dv[v] = a_d;
            sig_dv[v] = softplus_f(rd);
            a_dsum += a_d;
            sig_dsum += sig_dv[v];
            B0 += a_d * sigmoid_f(RCD(r, v, s, 0));
            B1 += a_d * sigmoid_f(RCD(r, v, s, 1));
            B2 += a_d * sigmoid_f(RCD(r, v, s, 2));
          }
        a_t = alpha_of(raw_tot, dist);
        Gw = gD * zr[s] + gA + (g_weights ? g_weights[(int64_t)r * S + s] : 0.f);
        val = T * (gr * B0 + gg * B1 + gb * B2 + a_t * Gw);   // T_k * dL/dT_k
      }
      const float incl = warp_rscan_sum(val, lane);
      const float after = incl - val + suffix;
      suffix += __shfl_sync(STAR_FULL_MASK, incl, 0);
      if (ok) {
        const float m_t = 1.f - a_t + 1e-10f;
        const float d_at = T * Gw - fdiv(after, m_t);
        const float shared_raw = d_at * dalpha_draw(raw_tot, dist);  // flows to every field's raw (:416-418)
        // direct alpha / sigma gradients
        float d_as = T * (gr * cs[0] + gg * cs[1] + gb * cs[2]);
        float d_sig_s = 0.f, d_sig_all = 0.f;  // d_sig_all: gradient w.r.t. sigma_sum (added to every field)
        float reg_common = 0.f, tc = 1.f, tot = 0.f;
        if (need_regs) {
          const float sig_s = softplus_f(ra_s);
          tot = a_s + a_dsum;
          tc = fmaxf(tot, STAR_EPS_F32);
          const float tot_ok = (tot >= STAR_EPS_F32) ? 1.f : 0.f;
          // dynamic-vs-static (:634-651)
          float E = 0.f, Csum = 0.f;
          {
            const float q = fdiv(a_s, tc), p = fmaxf(q, STAR_EPS_F32);
            E += p * flog(p);
            if (q >= STAR_EPS_F32) Csum += (flog(p) + 1.f) * a_s;
          }
#pragma unroll
          for (int v = 0; v < STAR_MAX_V; ++v)
            if (v < V) {
              const float q = fdiv(a_dv[v], tc), p = fmaxf(q, STAR_EPS_F32);
              E += p * flog(p);
              if (q >= STAR_EPS_F32) Csum += (flog(p) + 1.f) * a_dv[v];
            }
          reg_common = E - tot * tot_ok * fdiv(Csum, tc * tc);
          {
            const float q = fdiv(a_s, tc), p = fmaxf(q, STAR_EPS_F32);
            d_as += k_dvs * (reg_common + (q >= STAR_EPS_F32 ? fdiv(tot * (flog(p) + 1.f), tc) : 0.f));
          }
          // alpha entropy (:612-631)
          {
            const float c = fminf(fmaxf(a_s, STAR_EPS_F32), 1.f - STAR_EPS_F32);
            d_as += k_ent * (flog(c) - flog1m(c));
            // static reg (:698-711)
            if (a_s >= STAR_EPS_F32 && a_s <= 1.f - STAR_EPS_F32) d_as += k_sta * (flog(fdiv(c, Z)) - PL);
          }
          // ray reg (:682-695): gradient lands on the arg-max sample of each object
          const float ssum = sig_s + sig_dsum;
          const float sc = fmaxf(ssum, STAR_EPS_F32);
#pragma unroll
          for (int v = 0; v < STAR_MAX_V; ++v)
            if (v < V && amx[v] == s) {
              const float dn = k_ray * 2.f * mx[v];
              if (ssum >= STAR_EPS_F32) d_sig_all += -dn * sig_dv[v] / (sc * sc);
            }
          d_sig_s = d_sig_all;
        }
        const float spg_s = softplus_grad_f(ra_s);
        d_raw_alpha_s[(int64_t)r * S + s] = d_as * dalpha_draw(ra_s, dist) + shared_raw + d_sig_s * spg_s;
        {
          float* o = d_raw_rgb_s + ((int64_t)r * S + s) * 3;
          const float k = T * a_s;
          o[0] = k * gr * cs[0] * (1.f - cs[0]);
          o[1] = k * gg * cs[1] * (1.f - cs[1]);
          o[2] = k * gb * cs[2] * (1.f - cs[2]);
        }
#pragma unroll
        for (int v = 0; v < STAR_MAX_V; ++v)
          if (v < V) {
            const float rd = RAD(r, v, s);
            const float a_d = a_dv[v];
            const float c0 = sigmoid_f(RCD(r, v, s, 0)), c1 = sigmoid_f(RCD(r, v, s, 1)), c2 = sigmoid_f(RCD(r, v, s, 2));
            float d_ad = T * (gr * c0 + gg * c1 + gb * c2);
            float d_sig = 0.f;
            if (need_regs) {
              const float q = fdiv(a_d, tc), p = fmaxf(q, STAR_EPS_F32);
              d_ad += k_dvs * (reg_common + (q >= STAR_EPS_F32 ? fdiv(tot * (flog(p) + 1.f), tc) : 0.f));
              const float c = fminf(fmaxf(a_d, STAR_EPS_F32), 1.f - STAR_EPS_F32);
              d_ad += k_ent * (flog(c) - flog1m(c));
              d_sig = d_sig_all + k_dyn;
              if (amx[v] == s) {
                const float sc = fmaxf(softplus_f(ra_s) + sig_dsum, STAR_EPS_F32);
                d_sig += k_ray * 2.f * mx[v] / sc;
              }
            }
            const int64_t o1 = ((int64_t)r * V + v) * S + s;
            d_raw_alpha_d[o1] = d_ad * dalpha_draw(rd, dist) + shared_raw + d_sig * softplus_grad_f(rd);
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
#ifndef STAR_RAY_GRID_CAP
// one warp per ray, the grid over all rays: measured 2-5 % faster than a 148 x 16 grid with a grid-stride loop for the
// one-trip kernels (profiles/r2y_ab_grid_cap.txt); the kernels keep their loops, so any cap stays correct
#define STAR_RAY_GRID_CAP 0x3fffffff
#endif
static void warp_per_ray_cfg(int R, size_t smem_per_warp, int& blocks, int& threads, size_t& smem) {
  int wpb = 4;
  while (wpb > 1 && smem_per_warp * wpb > 200 * 1024) wpb >>= 1;
  threads = wpb * 32;
  smem = smem_per_warp * wpb;
  int64_t b = ((int64_t)R + wpb - 1) / wpb;
  if (b > STAR_RAY_GRID_CAP) b = STAR_RAY_GRID_CAP;
  blocks = (int)b;
}

extern "C" int star_composite_single_forward(const float* raw_alpha, const float* raw_rgb, const float* z_vals,
                                             const float* rays_d, int R, int S, float far_dist, int white_bkgd,
                                             float* rgb, float* disp, float* acc, float* depth, float* weights,
                                             float* dists, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!raw_alpha || !raw_rgb || !z_vals || !rays_d || !rgb || !disp || !acc || !depth || !weights) return STAR_E_NULL;
  if (R < 0 || S < 1) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  warp_per_ray_cfg(R, 0, blocks, threads, smem);
  const bool aligned8 = (((uintptr_t)raw_alpha | (uintptr_t)raw_rgb | (uintptr_t)z_vals | (uintptr_t)weights |
                          (uintptr_t)dists) & 7) == 0;
  if ((S & 1) == 0 && aligned8)
    composite_single_fwd_x2_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        raw_alpha, raw_rgb, z_vals, rays_d, R, S, far_dist, white_bkgd, rgb, disp, acc, depth, weights, dists);
  else
    composite_single_fwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        raw_alpha, raw_rgb, z_vals, rays_d, R, S, far_dist, white_bkgd, rgb, disp, acc, depth, weights, dists);
  return star_check_launch();
}

extern "C" int star_composite_single_backward(const float* raw_alpha, const float* raw_rgb, const float* z_vals,
                                              const float* rays_d, int R, int S, float far_dist, int white_bkgd,
                                              const float* g_rgb, const float* g_disp, const float* g_acc,
                                              const float* g_depth, const float* g_weights, float* d_raw_alpha,
                                              float* d_raw_rgb, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!raw_alpha || !raw_rgb || !z_vals || !rays_d || !d_raw_alpha || !d_raw_rgb) return STAR_E_NULL;
  if (R < 0 || S < 1 || S > 16384) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  warp_per_ray_cfg(R, sizeof(float) * 2 * S, blocks, threads, smem);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(composite_single_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  composite_single_bwd_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>(
      raw_alpha, raw_rgb, z_vals, rays_d, R, S, far_dist, white_bkgd, g_rgb, g_disp, g_acc, g_depth, g_weights,
      d_raw_alpha, d_raw_rgb);
  return star_check_launch();
}

static int multi_fwd_blocks(int R) {
  int64_t b = ((int64_t)R + 3) / 4;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" size_t star_composite_multi_ws_bytes(int R) { return sizeof(float) * 5 * (size_t)multi_fwd_blocks(R); }

extern "C" int star_composite_multi_forward(const float* raw_alpha_s, const float* raw_rgb_s,
                                            const float* raw_alpha_d, const float* raw_rgb_d, const float* z_vals,
                                            const float* rays_d, int R, int V, int S, float far_dist, int white_bkgd,
                                            int chunk, const StarMultiOut* out, void* workspace, void* stream) {
  if (!raw_alpha_s || !raw_rgb_s || !raw_alpha_d || !raw_rgb_d || !z_vals || !rays_d || !out || !workspace)
    return STAR_E_NULL;
  if (!out->rgb || !out->disp || !out->acc || !out->depth || !out->weights || !out->rgb_static ||
      !out->depth_static || !out->rgb_dynamic || !out->depth_dynamic || !out->dynamic_transmittance || !out->regs)
    return STAR_E_NULL;
  if (R < 1 || S < 1 || V < 1 || V > STAR_MAX_V || chunk < 1) return STAR_E_BAD_SHAPE;
  const int blocks = multi_fwd_blocks(R);
#define STAR_MULTI_FWD(VT)                                                                                       \
  case VT:                                                                                                       \
    composite_multi_fwd_kernel<VT><<<blocks, 128, 0, (cudaStream_t)stream>>>(                                    \
        raw_alpha_s, raw_rgb_s, raw_alpha_d, raw_rgb_d, z_vals, rays_d, R, V, S, far_dist, white_bkgd, chunk,    \
        *out, (float*)workspace);                                                                                \
    break;
  switch (V) {   // per-object accumulators live in registers: one instantiation per object count
    STAR_MULTI_FWD(1) STAR_MULTI_FWD(2) STAR_MULTI_FWD(3) STAR_MULTI_FWD(4)
    STAR_MULTI_FWD(5) STAR_MULTI_FWD(6) STAR_MULTI_FWD(7) STAR_MULTI_FWD(8)
  }
#undef STAR_MULTI_FWD
  int rc = star_check_launch();
  if (rc) return rc;
  reg_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float*)workspace, blocks, out->regs);
  return star_check_launch();
}

extern "C" int star_composite_multi_backward(const float* raw_alpha_s, const float* raw_rgb_s,
                                             const float* raw_alpha_d, const float* raw_rgb_d, const float* z_vals,
                                             const float* rays_d, int R, int V, int S, float far_dist,
                                             int white_bkgd, int chunk, const float* g_rgb, const float* g_disp,
                                             const float* g_acc, const float* g_depth, const float* g_weights,
                                             const float* g_regs, float* d_raw_alpha_s, float* d_raw_rgb_s,
                                             float* d_raw_alpha_d, float* d_raw_rgb_d, void* stream) {
  if (!raw_alpha_s || !raw_rgb_s || !raw_alpha_d || !raw_rgb_d || !z_vals || !rays_d || !d_raw_alpha_s ||
      !d_raw_rgb_s || !d_raw_alpha_d || !d_raw_rgb_d)
    return STAR_E_NULL;
  if (R < 1 || S < 1 || S > 32768 || V < 1 || V > STAR_MAX_V || chunk < 1) return STAR_E_BAD_SHAPE;
  int blocks, threads;
  size_t smem;
  warp_per_ray_cfg(R, sizeof(float) * S, blocks, threads, smem);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(composite_multi_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  composite_multi_bwd_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>(
      raw_alpha_s, raw_rgb_s, raw_alpha_d, raw_rgb_d, z_vals, rays_d, R, V, S, far_dist, white_bkgd, chunk, g_rgb,
      g_disp, g_acc, g_depth, g_weights, g_regs, d_raw_alpha_s, d_raw_rgb_s, d_raw_alpha_d, d_raw_rgb_d);
  return star_check_launch();
}
