// fp32 tier of the mip-NeRF field (models/mipnerf.py:53-100, models/star_mipnerf.py:200-260
// -> nerfstudio NeRFField(use_integrated_encoding=True)), forward and backward, on CUDA cores:
//   rigid transform of the ray (origin, direction) into the object frame (star_mipnerf.py:206-214)
//   -> conical frustum -> Gaussian (nerfstudio conical_frustum_to_gaussian, pixel_area = 1: star_mipnerf.py:267)
//   -> integrated positional encoding (nerfstudio NeRFEncoding with covariances) -> mlp_base (8 x 256, skip at 4)
//   -> density head; mlp_head (2 x 128) on cat[encoded_dir, base_out] -> rgb head.
// Returns RAW density / rgb (pre-softplus / pre-sigmoid): the activations live in the compositing kernels
// (mip_render.cu), exactly like the vanilla path.  Same tile machinery as mlp_f32.cu (mlp_f32_device.cuh).
#include "star_common.cuh"
#include "mlp_f32_device.cuh"
#include "mip_layout.h"

#define MIP_TWO_PI 6.2831854820251465f   // float32(2 * pi)
#define MIP_PIO2 1.5707963705062866f     // float32(pi / 2)

struct MipGeom {
  float o[3], d[3];        // object-frame ray origin and direction
  float tmean, dvar, rvar; // conical frustum moments along / across the ray
};

__device__ __forceinline__ MipGeom mip_geom(const float* __restrict__ origins, const float* __restrict__ dirs,
                                            const float* __restrict__ pose12, const float* __restrict__ bins,
                                            int64_t gi, int S, float radius) {
  MipGeom g;
  const int64_t r = gi / S;
  const int s = (int)(gi - r * S);
  const float ox = origins[r * 3 + 0], oy = origins[r * 3 + 1], oz = origins[r * 3 + 2];
  const float dx = dirs[r * 3 + 0], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
  if (pose12 != nullptr) {   // o' = R o + t, d' = R d   (pp.SE3.Act / pp.SO3.Act, star_mipnerf.py:207-212)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g.o[i] = pose12[i * 3 + 0] * ox + pose12[i * 3 + 1] * oy + pose12[i * 3 + 2] * oz + pose12[9 + i];
      g.d[i] = pose12[i * 3 + 0] * dx + pose12[i * 3 + 1] * dy + pose12[i * 3 + 2] * dz;
    }
  } else {
    g.o[0] = ox; g.o[1] = oy; g.o[2] = oz;
    g.d[0] = dx; g.d[1] = dy; g.d[2] = dz;
  }
  // conical_frustum_to_gaussian, one rounding per reference op (no FMA contraction): the encoding multiplies these
  // numbers by up to 2 pi 2^24, so the phase of the surviving frequencies follows the reference's roundings
  const float st = bins[r * (S + 1) + s], en = bins[r * (S + 1) + s + 1];
  const float mu = __fdiv_rn(__fadd_rn(st, en), 2.f), hw = __fdiv_rn(__fsub_rn(en, st), 2.f);
  const float mu2 = __fmul_rn(mu, mu), hw2 = __fmul_rn(hw, hw), hw4 = __fmul_rn(hw2, hw2);
  const float den = __fadd_rn(__fmul_rn(3.f, mu2), hw2);
  g.tmean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.f, mu), hw2), den));
  g.dvar = __fsub_rn(__fdiv_rn(hw2, 3.f),
                     __fmul_rn(4.f / 15.f, __fdiv_rn(__fmul_rn(hw4, __fsub_rn(__fmul_rn(12.f, mu2), hw2)), __fmul_rn(den, den))));
  g.rvar = __fmul_rn(radius * radius,
                     __fsub_rn(__fadd_rn(__fdiv_rn(mu2, 4.f), __fmul_rn(5.f / 12.f, hw2)), __fdiv_rn(__fmul_rn(4.f / 15.f, hw4), den)));
  return g;
}

// mean / diagonal covariance of the Gaussian along axis c (compute_3d_gaussian), same rounding policy
__device__ __forceinline__ float mip_mean(const MipGeom& g, int c) { return __fadd_rn(g.o[c], __fmul_rn(g.d[c], g.tmean)); }
__device__ __forceinline__ float mip_diag(const MipGeom& g, int c, float mag) {
  return __fadd_rn(__fmul_rn(g.dvar, __fmul_rn(g.d[c], g.d[c])),
                   __fmul_rn(g.rvar, __fsub_rn(1.f, __fmul_rn(g.d[c], __fdiv_rn(g.d[c], mag)))));
}

__device__ __forceinline__ float mip_mag(const MipGeom& g) {
  return fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(g.d[0], g.d[0]), __fmul_rn(g.d[1], g.d[1])), __fmul_rn(g.d[2], g.d[2])), 1e-10f);
}

// ================================================================================ forward kernel
template <bool STASH>
__global__ void __launch_bounds__(NTHREADS, 1)
mip_fwd_f32_kernel(MipLayout lay, const float* __restrict__ packed, const float* __restrict__ origins,
                   const float* __restrict__ dirs, const float* __restrict__ pose12, const float* __restrict__ bins,
                   const float* __restrict__ freqs, float radius, int S, int64_t M, float* __restrict__ raw_sigma,
                   float* __restrict__ raw_rgb, int64_t ray_stride, float* __restrict__ stash) {
  extern __shared__ __align__(16) float smem[];
  float* E = smem;                       // [160][AS] integrated positional encoding
  float* Ed = E + MIP_KXP * AS;          // [32][AS]  encoded direction
  float* A0 = Ed + MIP_KDP * AS;         // [256][AS]
  float* A1 = A0 + MIP_W * AS;           // [256][AS]
  float* Wbuf = A1 + MIP_W * AS;         // [2][KS][256]
  __shared__ float s_f[MIP_FREQ_FLOATS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TM - 1) / TM;
  if (tid < MIP_FREQ_FLOATS) s_f[tid] = freqs[tid];
  __syncthreads();

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * TM;
    {
      const int m = tid & 63, part = tid >> 6;
      const int64_t gi = m0 + m;
      if (gi < M) {
        const MipGeom g = mip_geom(origins, dirs, pose12, bins, gi, S, radius);
        const float mag = mip_mag(g);
        float mean[3], diag[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mean[c] = mip_mean(g, c);
          diag[c] = mip_diag(g, c, mag);
        }
        for (int j = part; j < 3 * MIP_NF; j += 4) {
          const int c = j / MIP_NF, k = j - c * MIP_NF;
          const float mc = c == 0 ? mean[0] : (c == 1 ? mean[1] : mean[2]);
          const float dc = c == 0 ? diag[0] : (c == 1 ? diag[1] : diag[2]);
          const float a = __fmul_rn(__fmul_rn(MIP_TWO_PI, mc), s_f[k]);
          const float e = expf(-0.5f * __fmul_rn(dc, s_f[MIP_NF + k]));
          float v0 = 0.f, v1 = 0.f;
          if (e != 0.f) {   // fully damped features are exactly 0 (and skip sinf's huge-argument path)
            v0 = e * sinf(a);
            v1 = e * sinf(__fadd_rn(a, MIP_PIO2));
          }
          E[j * AS + m] = v0;
          E[(3 * MIP_NF + j) * AS + m] = v1;
        }
        if (part == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            E[(6 * MIP_NF + c) * AS + m] = mean[c];
            Ed[(6 * MIP_NFD + c) * AS + m] = g.d[c];
          }
        }
        for (int j = part; j < 3 * MIP_NFD; j += 4) {
          const int c = j / MIP_NFD, k = j - c * MIP_NFD;
          const float dc = c == 0 ? g.d[0] : (c == 1 ? g.d[1] : g.d[2]);
          const float a = __fmul_rn(__fmul_rn(MIP_TWO_PI, dc), s_f[2 * MIP_NF + k]);
          Ed[j * AS + m] = sinf(a);
          Ed[(3 * MIP_NFD + j) * AS + m] = sinf(__fadd_rn(a, MIP_PIO2));
        }
      } else {
        for (int j = part; j < MIP_KX; j += 4) E[j * AS + m] = 0.f;
        for (int j = part; j < MIP_KD; j += 4) Ed[j * AS + m] = 0.f;
      }
      for (int j = MIP_KX + part; j < MIP_KXP; j += 4) E[j * AS + m] = 0.f;
      for (int j = MIP_KD + part; j < MIP_KDP; j += 4) Ed[j * AS + m] = 0.f;
    }
    __syncthreads();
    if (STASH) {
      float* de = stash + lay.s_enc * M;
      for (int i = tid; i < TM * MIP_KXP; i += NTHREADS) {
        const int m = i / MIP_KXP, j = i - m * MIP_KXP;
        if (m0 + m < M) de[(m0 + m) * MIP_KXP + j] = E[j * AS + m];
      }
      float* dd = stash + lay.s_dir * M;
      for (int i = tid; i < TM * MIP_KDP; i += NTHREADS) {
        const int m = i >> 5, j = i & 31;
        if (m0 + m < M) dd[(m0 + m) * MIP_KDP + j] = Ed[j * AS + m];
      }
    }

    float* cur = A0;
    float* nxt = A1;
    for (int l = 0; l < MIP_NBASE; ++l) {
      float acc[8][8];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) acc[mi][nj] = 0.f;
      const float* wt = packed + lay.p_wt[l];
      if (l == 0) {
        gemm_tile<2>(E, wt, MIP_W, MIP_KXP, acc, Wbuf, tid);
      } else if (l == MIP_SKIP) {   // cat[enc, x]
        gemm_tile<2>(E, wt, MIP_W, MIP_KXP, acc, Wbuf, tid);
        gemm_tile<2>(cur, wt + (int64_t)MIP_KXP * MIP_W, MIP_W, MIP_W, acc, Wbuf, tid);
      } else {
        gemm_tile<2>(cur, wt, MIP_W, MIP_W, acc, Wbuf, tid);
      }
      const float* bias = packed + lay.p_b[l];
      float bv[8];
#pragma unroll
      for (int nj = 0; nj < 8; ++nj) bv[nj] = bias[col_of(lane, nj)];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) {
          acc[mi][nj] = fmaxf(acc[mi][nj] + bv[nj], 0.f);
          nxt[col_of(lane, nj) * AS + warp * 8 + mi] = acc[mi][nj];
        }
      if (STASH) {
        float* dst = stash + (l + 1 < MIP_NBASE ? lay.s_in[l + 1] : lay.s_base) * M;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
          const int64_t gi = m0 + warp * 8 + mi;
          if (gi < M) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              *reinterpret_cast<float4*>(dst + gi * MIP_W + lane * 4 + 128 * j) =
                  make_float4(acc[mi][4 * j], acc[mi][4 * j + 1], acc[mi][4 * j + 2], acc[mi][4 * j + 3]);
          }
        }
      }
      if (l == MIP_NBASE - 1) {   // density head on base_out (DensityFieldHead, pre-softplus)
        const float* dw = packed + lay.p_dw;
        float wv[8];
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) wv[nj] = dw[col_of(lane, nj)];
        const float db = packed[lay.p_db];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
          float p = 0.f;
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) p = fmaf(acc[mi][nj], wv[nj], p);
          p = warp_sum(p);
          const int64_t gi = m0 + warp * 8 + mi;
          if (lane == 0 && gi < M) raw_sigma[(gi / S) * ray_stride + (gi % S)] = p + db;
        }
      }
      float* t = cur; cur = nxt; nxt = t;
      __syncthreads();
    }
    // ---- mlp_head layer 0 on cat[encoded_dir, base_out]
    {
      float acc[8][4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[mi][nj] = 0.f;
      const float* wt = packed + lay.p_h0t;
      gemm_tile<1>(Ed, wt, MIP_WH, MIP_KDP, acc, Wbuf, tid);
      gemm_tile<1>(cur, wt + (int64_t)MIP_KDP * MIP_WH, MIP_WH, MIP_W, acc, Wbuf, tid);
      const float* bias = packed + lay.p_h0b;
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        const float b = bias[lane * 4 + nj];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
          acc[mi][nj] = fmaxf(acc[mi][nj] + b, 0.f);
          nxt[(lane * 4 + nj) * AS + warp * 8 + mi] = acc[mi][nj];
        }
      }
      if (STASH) {
        float* dst = stash + lay.s_h0 * M;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
          const int64_t gi = m0 + warp * 8 + mi;
          if (gi < M)
            *reinterpret_cast<float4*>(dst + gi * MIP_WH + lane * 4) =
                make_float4(acc[mi][0], acc[mi][1], acc[mi][2], acc[mi][3]);
        }
      }
      __syncthreads();
    }
    // ---- mlp_head layer 1 + rgb head (pre-sigmoid)
    {
      float acc[8][4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[mi][nj] = 0.f;
      gemm_tile<1>(nxt, packed + lay.p_h1t, MIP_WH, MIP_WH, acc, Wbuf, tid);
      const float* bias = packed + lay.p_h1b;
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        const float b = bias[lane * 4 + nj];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) acc[mi][nj] = fmaxf(acc[mi][nj] + b, 0.f);
      }
      if (STASH) {
        float* dst = stash + lay.s_h1 * M;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
          const int64_t gi = m0 + warp * 8 + mi;
          if (gi < M)
            *reinterpret_cast<float4*>(dst + gi * MIP_WH + lane * 4) =
                make_float4(acc[mi][0], acc[mi][1], acc[mi][2], acc[mi][3]);
        }
      }
      const float* rw = packed + lay.p_rw;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float wv[4];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) wv[nj] = rw[c * MIP_WH + lane * 4 + nj];
        const float rb = packed[lay.p_rb + c];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
          float p = 0.f;
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) p = fmaf(acc[mi][nj], wv[nj], p);
          p = warp_sum(p);
          const int64_t gi = m0 + warp * 8 + mi;
          if (lane == 0 && gi < M) raw_rgb[((gi / S) * ray_stride + (gi % S)) * 3 + c] = p + rb;
        }
      }
    }
    __syncthreads();
  }
}

// ================================================================================ backward (dX chain)
// Walks the layers in reverse for one tile of 64 samples, producing G_l = dL/d(pre-activation of layer l) for
// every GEMM layer into `gst` (consumed by the dW GEMMs) and, for dynamic objects, the Euclidean pose accumulators
// of include/star_b200.h (sum g, sum g o^T, sum h d^T with g = dL/do', h = dL/dd').
__global__ void __launch_bounds__(NTHREADS, 1)
mip_bwd_f32_kernel(MipLayout lay, const float* __restrict__ packed, const float* __restrict__ origins,
                   const float* __restrict__ dirs, const float* __restrict__ pose12, const float* __restrict__ bins,
                   const float* __restrict__ freqs, float radius, int S, int64_t M,
                   const float* __restrict__ d_raw_sigma, const float* __restrict__ d_raw_rgb, int64_t ray_stride,
                   const float* __restrict__ stash, float* __restrict__ gst, float* __restrict__ pose_acc) {
  extern __shared__ __align__(16) float smem[];
  float* G0 = smem;                      // [256][AS]
  float* G1 = G0 + MIP_W * AS;           // [256][AS]
  float* Wbuf = G1 + MIP_W * AS;         // [2][KS][256]
  float* s_dr = Wbuf + 2 * KS * 256;     // [64][4]: d_raw_rgb (3), d_raw_sigma
  float* Den = s_dr + TM * 4;            // [160][AS] dL/d(encoding)   (objects only)
  float* Ded = Den + MIP_KXP * AS;       // [32][AS]  dL/d(encoded dir) (objects only)
  __shared__ float s_f[MIP_FREQ_FLOATS];
  __shared__ float s_pose[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TM - 1) / TM;
  const bool want_pose = pose12 != nullptr;
  if (tid < MIP_FREQ_FLOATS) s_f[tid] = freqs[tid];
  float pacc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) pacc[i] = 0.f;
  __syncthreads();

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * TM;
    if (tid < TM) {
      const int64_t gi = m0 + tid;
      float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (gi < M) {
        const int64_t o = (gi / S) * ray_stride + (gi % S);
        a = d_raw_sigma[o];
        c0 = d_raw_rgb[o * 3 + 0]; c1 = d_raw_rgb[o * 3 + 1]; c2 = d_raw_rgb[o * 3 + 2];
      }
      s_dr[tid * 4 + 0] = c0; s_dr[tid * 4 + 1] = c1; s_dr[tid * 4 + 2] = c2; s_dr[tid * 4 + 3] = a;
    }
    __syncthreads();
    // ---- rgb head + ReLU of head layer 1: G_h1[m][n] = (sum_c d_rgb[m][c] Wr[c][n]) * [h1 > 0]  -> G0 rows [0,128)
    {
      const float* rw = packed + lay.p_rw;
      const float* h1s = stash + lay.s_h1 * M;
      float* gdst = gst + lay.g_h1 * M;
      float w0[4], w1[4], w2[4];
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        w0[nj] = rw[0 * MIP_WH + lane * 4 + nj];
        w1[nj] = rw[1 * MIP_WH + lane * 4 + nj];
        w2[nj] = rw[2 * MIP_WH + lane * 4 + nj];
      }
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int m = warp * 8 + mi;
        const int64_t gi = m0 + m;
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < M) h = *reinterpret_cast<const float4*>(h1s + gi * MIP_WH + lane * 4);
        const float hv[4] = {h.x, h.y, h.z, h.w};
        float g[4];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          const float v = s_dr[m * 4 + 0] * w0[nj] + s_dr[m * 4 + 1] * w1[nj] + s_dr[m * 4 + 2] * w2[nj];
          g[nj] = hv[nj] > 0.f ? v : 0.f;
          G0[(lane * 4 + nj) * AS + m] = g[nj];
        }
        if (gi < M) *reinterpret_cast<float4*>(gdst + gi * MIP_WH + lane * 4) = make_float4(g[0], g[1], g[2], g[3]);
      }
    }
    __syncthreads();
    // ---- G_h0 = (G_h1 * W_h1) * [h0 > 0]  -> G1 rows [0,128)
    {
      float acc[8][4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[mi][nj] = 0.f;
      gemm_tile<1>(G0, packed + lay.p_h1w, MIP_WH, MIP_WH, acc, Wbuf, tid);
      const float* h0s = stash + lay.s_h0 * M;
      float* gdst = gst + lay.g_h0 * M;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int m = warp * 8 + mi;
        const int64_t gi = m0 + m;
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < M) h = *reinterpret_cast<const float4*>(h0s + gi * MIP_WH + lane * 4);
        const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          if (!(hv[nj] > 0.f)) acc[mi][nj] = 0.f;
          G1[(lane * 4 + nj) * AS + m] = acc[mi][nj];
        }
        if (gi < M)
          *reinterpret_cast<float4*>(gdst + gi * MIP_WH + lane * 4) = make_float4(acc[mi][0], acc[mi][1], acc[mi][2], acc[mi][3]);
      }
    }
    __syncthreads();
    float acc[8][8];
    auto zero_acc = [&]() {
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) acc[mi][nj] = 0.f;
    };
    // multiply acc by [stashed activation > 0], store to smem (transposed) for the next GEMM and to gst for dW
    auto mask_put = [&](float* Gs, int64_t s_off, int64_t g_off) {
      const float* src = stash + s_off * M;
      float* gdst = gst + g_off * M;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int64_t gi = m0 + warp * 8 + mi;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gi < M) h = *reinterpret_cast<const float4*>(src + gi * MIP_W + lane * 4 + 128 * j);
          if (!(h.x > 0.f)) acc[mi][4 * j + 0] = 0.f;
          if (!(h.y > 0.f)) acc[mi][4 * j + 1] = 0.f;
          if (!(h.z > 0.f)) acc[mi][4 * j + 2] = 0.f;
          if (!(h.w > 0.f)) acc[mi][4 * j + 3] = 0.f;
        }
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) Gs[col_of(lane, nj) * AS + warp * 8 + mi] = acc[mi][nj];
        if (gi < M) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<float4*>(gdst + gi * MIP_W + lane * 4 + 128 * j) =
                make_float4(acc[mi][4 * j], acc[mi][4 * j + 1], acc[mi][4 * j + 2], acc[mi][4 * j + 3]);
        }
      }
    };
    // ---- dL/d(encoded dir) = G_h0 * W_h0[:, 0:32]   (objects only)  -> Ded
    if (want_pose) {
      const int m = tid & 63, j0 = (tid >> 6) * 8;
      float de[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) de[j] = 0.f;
      const float* wb = packed + lay.p_h0bd;   // [128][32]
      for (int n = 0; n < MIP_WH; ++n) {
        const float g = G1[n * AS + m];
#pragma unroll
        for (int j = 0; j < 8; ++j) de[j] = fmaf(g, wb[n * MIP_KDP + j0 + j], de[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Ded[(j0 + j) * AS + m] = de[j];
    }
    // ---- G_7 = (G_h0 * W_h0[:, 32:] + d_sigma * w_density) * [base_out > 0]  -> G0
    zero_acc();
    gemm_tile<2>(G1, packed + lay.p_h0bx, MIP_W, MIP_WH, acc, Wbuf, tid);
    {
      const float* dw = packed + lay.p_dw;
#pragma unroll
      for (int nj = 0; nj < 8; ++nj) {
        const float w = dw[col_of(lane, nj)];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) acc[mi][nj] = fmaf(s_dr[(warp * 8 + mi) * 4 + 3], w, acc[mi][nj]);
      }
    }
    mask_put(G0, lay.s_base, lay.g_base[MIP_NBASE - 1]);
    __syncthreads();
    // ---- base layers in reverse:  G_{l-1} = (G_l * W_l[x part]) * [In_l > 0]
    float* gc = G0;
    float* gn = G1;
    for (int l = MIP_NBASE - 1; l >= 1; --l) {
      if (l == MIP_SKIP && want_pose) {   // encoding part of the skip layer -> Den
        zero_acc();
        gemm_tile<2>(gc, packed + lay.p_wbe[1], MIP_W, MIP_W, acc, Wbuf, tid);
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) {
            const int col = col_of(lane, nj);
            if (col < MIP_KXP) Den[col * AS + warp * 8 + mi] = acc[mi][nj];
          }
      }
      zero_acc();
      gemm_tile<2>(gc, packed + lay.p_wb[l], MIP_W, MIP_W, acc, Wbuf, tid);
      mask_put(gn, lay.s_in[l], lay.g_base[l - 1]);
      float* t = gc; gc = gn; gn = t;
      __syncthreads();
    }
    // ---- pose gradients (objects only)
    if (want_pose) {
      zero_acc();
      gemm_tile<2>(gc, packed + lay.p_wbe[0], MIP_W, MIP_W, acc, Wbuf, tid);
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) {
          const int col = col_of(lane, nj);
          if (col < MIP_KXP) Den[col * AS + warp * 8 + mi] += acc[mi][nj];
        }
      __syncthreads();
      if (tid < TM) {
        const int m = tid;
        const int64_t gi = m0 + m;
        if (gi < M) {
          const MipGeom g = mip_geom(origins, dirs, pose12, bins, gi, S, radius);
          const float mag = mip_mag(g);
          const float dd = g.d[0] * g.d[0] + g.d[1] * g.d[1] + g.d[2] * g.d[2];
          float gmean[3], gdiag[3], h[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float mean = mip_mean(g, c);
            const float diag = mip_diag(g, c, mag);
            float gm = Den[(6 * MIP_NF + c) * AS + m], gd = 0.f;
            for (int k = 0; k < MIP_NF; ++k) {
              const float f = s_f[k], f2 = s_f[MIP_NF + k];
              const float e = expf(-0.5f * __fmul_rn(diag, f2));
              if (e != 0.f) {
                const float a = __fmul_rn(__fmul_rn(MIP_TWO_PI, mean), f), a2 = __fadd_rn(a, MIP_PIO2);
                const float g1 = Den[(c * MIP_NF + k) * AS + m], g2 = Den[(3 * MIP_NF + c * MIP_NF + k) * AS + m];
                float s1, c1, s2, c2;
                sincosf(a, &s1, &c1);
                sincosf(a2, &s2, &c2);
                gm += e * (g1 * c1 + g2 * c2) * (MIP_TWO_PI * f);
                gd += e * (g1 * s1 + g2 * s2) * (-0.5f * f2);
              }
            }
            gmean[c] = gm;
            gdiag[c] = gd;
            // encoded direction
            float hd = Ded[(6 * MIP_NFD + c) * AS + m];
            for (int k = 0; k < MIP_NFD; ++k) {
              const float f = s_f[2 * MIP_NF + k];
              const float a = __fmul_rn(__fmul_rn(MIP_TWO_PI, g.d[c]), f), a2 = __fadd_rn(a, MIP_PIO2);
              hd += (Ded[(c * MIP_NFD + k) * AS + m] * cosf(a) + Ded[(3 * MIP_NFD + c * MIP_NFD + k) * AS + m] * cosf(a2)) *
                    (MIP_TWO_PI * f);
            }
            h[c] = hd + gm * g.tmean;
          }
          // diag_c = dvar d_c^2 + rvar (1 - d_c^2 / mag),  mag = max(|d|^2, 1e-10)
          float cross = 0.f;   // sum_c gdiag_c rvar d_c^2 / mag^2  (d mag / d d_j = 2 d_j when not clamped)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            h[c] += gdiag[c] * 2.f * g.d[c] * (g.dvar - g.rvar / mag);
            cross += gdiag[c] * g.rvar * g.d[c] * g.d[c] / (mag * mag);
          }
          if (dd >= 1e-10f) {
#pragma unroll
            for (int c = 0; c < 3; ++c) h[c] += cross * 2.f * g.d[c];
          }
          const int64_t r = gi / S;
          const float ow[3] = {origins[r * 3 + 0], origins[r * 3 + 1], origins[r * 3 + 2]};
          const float dw[3] = {dirs[r * 3 + 0], dirs[r * 3 + 1], dirs[r * 3 + 2]};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            pacc[i] += gmean[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              pacc[3 + i * 3 + j] += gmean[i] * ow[j];
              pacc[15 + i * 3 + j] += h[i] * dw[j];
            }
          }
          pacc[12] += g.o[1] * gmean[2] - g.o[2] * gmean[1];
          pacc[13] += g.o[2] * gmean[0] - g.o[0] * gmean[2];
          pacc[14] += g.o[0] * gmean[1] - g.o[1] * gmean[0];
          pacc[24] += g.d[1] * h[2] - g.d[2] * h[1];
          pacc[25] += g.d[2] * h[0] - g.d[0] * h[2];
          pacc[26] += g.d[0] * h[1] - g.d[1] * h[0];
        }
      }
    }
    __syncthreads();
  }
  if (want_pose) {
    if (tid < 32) s_pose[tid] = 0.f;
    __syncthreads();
    if (tid < TM) {
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const float v = warp_sum(pacc[i]);
        if (lane == 0) atomicAdd(&s_pose[i], v);
      }
    }
    __syncthreads();
    if (tid < 27) atomicAdd(&pose_acc[tid], s_pose[tid]);
  }
}

// Head gradients: d w_density[k] = sum_m d_sigma[m] base_out[m][k]; d w_rgb[c][n] = sum_m d_rgb[m][c] h1[m][n]; biases.
__global__ void __launch_bounds__(256)
mip_head_grad_kernel(MipLayout lay, const float* __restrict__ stash, const float* __restrict__ d_raw_sigma,
                     const float* __restrict__ d_raw_rgb, int64_t ray_stride, int S, int64_t M,
                     float* __restrict__ grad_flat) {
  const int tid = threadIdx.x;
  const int64_t per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t mb = (int64_t)blockIdx.x * per, me = (mb + per < M) ? mb + per : M;
  const float* x = stash + lay.s_base * M;
  const float* h1 = stash + lay.s_h1 * M;
  float aw = 0.f, ab = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  const int n2 = tid & 127;
  for (int64_t m = mb; m < me; ++m) {
    const int64_t o = (m / S) * ray_stride + (m % S);
    const float da = d_raw_sigma[o];
    aw = fmaf(da, x[m * MIP_W + tid], aw);
    if (tid == 0) ab += da;
    if (tid < 128) {
      const float c0 = d_raw_rgb[o * 3 + 0], c1 = d_raw_rgb[o * 3 + 1], c2 = d_raw_rgb[o * 3 + 2];
      const float v = h1[m * MIP_WH + n2];
      r0 = fmaf(c0, v, r0); r1 = fmaf(c1, v, r1); r2 = fmaf(c2, v, r2);
      if (tid == 0) { b0 += c0; b1 += c1; b2 += c2; }
    }
  }
  atomicAdd(&grad_flat[lay.m_w[MIP_L_DENS] + tid], aw);
  if (tid < 128) {
    atomicAdd(&grad_flat[lay.m_w[MIP_L_RGB] + 0 * MIP_WH + n2], r0);
    atomicAdd(&grad_flat[lay.m_w[MIP_L_RGB] + 1 * MIP_WH + n2], r1);
    atomicAdd(&grad_flat[lay.m_w[MIP_L_RGB] + 2 * MIP_WH + n2], r2);
  }
  if (tid == 0) {
    atomicAdd(&grad_flat[lay.m_b[MIP_L_DENS]], ab);
    atomicAdd(&grad_flat[lay.m_b[MIP_L_RGB] + 0], b0);
    atomicAdd(&grad_flat[lay.m_b[MIP_L_RGB] + 1], b1);
    atomicAdd(&grad_flat[lay.m_b[MIP_L_RGB] + 2], b2);
  }
}

// ================================================================================ weight packing
__global__ void mip_pack_f32_kernel(MipLayout lay, const float* __restrict__ master, float* __restrict__ packed) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= lay.n_packed) return;
  float v = 0.f;
  // master column of padded input row kp of a layer whose input is [enc (147 -> 160) | x (256)] / [dirs (27 -> 32) | x]
  auto col_skip = [](int kp, int kenc, int kencp) -> int { return kp < kencp ? (kp < kenc ? kp : -1) : kenc + (kp - kencp); };
  bool done = false;
  for (int l = 0; l < MIP_NBASE && !done; ++l) {
    const int K = lay.K[l];
    const int Kp = l == 0 ? MIP_KXP : (l == MIP_SKIP ? MIP_KXP + MIP_W : MIP_W);
    if (i >= lay.p_wt[l] && i < lay.p_wt[l] + (int64_t)Kp * MIP_W) {          // W^T [Kp][256]
      const int64_t j = i - lay.p_wt[l];
      const int kp = (int)(j / MIP_W), n = (int)(j % MIP_W);
      const int k = (l == 0 || l == MIP_SKIP) ? col_skip(kp, MIP_KX, MIP_KXP) : kp;
      v = (k >= 0 && k < K) ? master[lay.m_w[l] + (int64_t)n * K + k] : 0.f;
      done = true;
    } else if (l >= 1 && i >= lay.p_wb[l] && i < lay.p_wb[l] + (int64_t)MIP_W * MIP_W) {   // x part of W [256][256]
      const int64_t j = i - lay.p_wb[l];
      const int n = (int)(j / MIP_W), kx = (int)(j % MIP_W);
      v = master[lay.m_w[l] + (int64_t)n * K + (l == MIP_SKIP ? MIP_KX : 0) + kx];
      done = true;
    } else if (i >= lay.p_b[l] && i < lay.p_b[l] + MIP_W) {
      v = master[lay.m_b[l] + (i - lay.p_b[l])];
      done = true;
    }
  }
  for (int e = 0; e < 2 && !done; ++e) {
    if (i >= lay.p_wbe[e] && i < lay.p_wbe[e] + (int64_t)MIP_W * MIP_W) {      // enc part [256][256 (147 used)]
      const int l = e == 0 ? 0 : MIP_SKIP;
      const int64_t j = i - lay.p_wbe[e];
      const int n = (int)(j / MIP_W), k = (int)(j % MIP_W);
      v = k < MIP_KX ? master[lay.m_w[l] + (int64_t)n * lay.K[l] + k] : 0.f;
      done = true;
    }
  }
  if (!done) {
    const int K0 = MIP_KD + MIP_W;
    if (i >= lay.p_h0t && i < lay.p_h0bx) {                                     // head 0 W^T [32 + 256][128]
      const int64_t j = i - lay.p_h0t;
      const int kp = (int)(j / MIP_WH), n = (int)(j % MIP_WH);
      const int k = col_skip(kp, MIP_KD, MIP_KDP);
      v = k >= 0 ? master[lay.m_w[MIP_L_H0] + (int64_t)n * K0 + k] : 0.f;
    } else if (i >= lay.p_h0bx && i < lay.p_h0bd) {                             // head 0 x part [128][256]
      const int64_t j = i - lay.p_h0bx;
      v = master[lay.m_w[MIP_L_H0] + (j / MIP_W) * K0 + MIP_KD + (j % MIP_W)];
    } else if (i >= lay.p_h0bd && i < lay.p_h0b) {                              // head 0 dir part [128][32]
      const int64_t j = i - lay.p_h0bd;
      const int n = (int)(j / MIP_KDP), k = (int)(j % MIP_KDP);
      v = k < MIP_KD ? master[lay.m_w[MIP_L_H0] + (int64_t)n * K0 + k] : 0.f;
    } else if (i >= lay.p_h0b && i < lay.p_h1t) {
      v = master[lay.m_b[MIP_L_H0] + (i - lay.p_h0b)];
    } else if (i >= lay.p_h1t && i < lay.p_h1w) {                               // head 1 W^T [128][128]
      const int64_t j = i - lay.p_h1t;
      v = master[lay.m_w[MIP_L_H1] + (j % MIP_WH) * MIP_WH + (j / MIP_WH)];
    } else if (i >= lay.p_h1w && i < lay.p_h1b) {
      v = master[lay.m_w[MIP_L_H1] + (i - lay.p_h1w)];
    } else if (i >= lay.p_h1b && i < lay.p_dw) {
      v = master[lay.m_b[MIP_L_H1] + (i - lay.p_h1b)];
    } else if (i >= lay.p_dw && i < lay.p_db) {
      v = master[lay.m_w[MIP_L_DENS] + (i - lay.p_dw)];
    } else if (i == lay.p_db) {
      v = master[lay.m_b[MIP_L_DENS]];
    } else if (i >= lay.p_rw && i < lay.p_rb) {
      v = master[lay.m_w[MIP_L_RGB] + (i - lay.p_rw)];
    } else if (i >= lay.p_rb && i < lay.p_rb + 3) {
      v = master[lay.m_b[MIP_L_RGB] + (i - lay.p_rb)];
    }
  }
  packed[i] = v;
}

// ================================================================================ host side
static const size_t MIP_FWD_SMEM = sizeof(float) * ((MIP_KXP + MIP_KDP + 2 * MIP_W) * AS + 2 * KS * 256);
static const size_t MIP_BWD_SMEM = sizeof(float) * (2 * MIP_W * AS + 2 * KS * 256 + TM * 4 + (MIP_KXP + MIP_KDP) * AS);

static int mip_grid_for_tiles(int64_t ntiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(ntiles < sms ? ntiles : sms);
}

int star_mip_f32_pack(const MipLayout& lay, const float* master, void* packed, cudaStream_t st) {
  const int threads = 256;
  const int blocks = (int)((lay.n_packed + threads - 1) / threads);
  mip_pack_f32_kernel<<<blocks, threads, 0, st>>>(lay, master, (float*)packed);
  return star_check_launch();
}

int star_mip_f32_forward(const MipLayout& lay, const void* packed, const float* origins, const float* dirs,
                         const float* pose12, const float* bins, const float* freqs, float radius, int R, int S,
                         float* raw_sigma, float* raw_rgb, int64_t ray_stride, void* stash, cudaStream_t st) {
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TM - 1) / TM;
  const int grid = mip_grid_for_tiles(ntiles);
  if (stash != nullptr) {
    cudaFuncSetAttribute(mip_fwd_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIP_FWD_SMEM);
    mip_fwd_f32_kernel<true><<<grid, NTHREADS, MIP_FWD_SMEM, st>>>(lay, (const float*)packed, origins, dirs, pose12,
                                                                   bins, freqs, radius, S, M, raw_sigma, raw_rgb,
                                                                   ray_stride, (float*)stash);
  } else {
    cudaFuncSetAttribute(mip_fwd_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIP_FWD_SMEM);
    mip_fwd_f32_kernel<false><<<grid, NTHREADS, MIP_FWD_SMEM, st>>>(lay, (const float*)packed, origins, dirs, pose12,
                                                                    bins, freqs, radius, S, M, raw_sigma, raw_rgb,
                                                                    ray_stride, nullptr);
  }
  return star_check_launch();
}

int star_mip_f32_backward(const MipLayout& lay, const void* packed, const float* origins, const float* dirs,
                          const float* pose12, const float* bins, const float* freqs, float radius, int R, int S,
                          const float* d_raw_sigma, const float* d_raw_rgb, int64_t ray_stride, const void* stash,
                          void* workspace, float* grad_flat, float* pose_acc, cudaStream_t st) {
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TM - 1) / TM;
  float* gst = (float*)workspace;
  const float* stf = (const float*)stash;
  cudaFuncSetAttribute(mip_bwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIP_BWD_SMEM);
  mip_bwd_f32_kernel<<<mip_grid_for_tiles(ntiles), NTHREADS, MIP_BWD_SMEM, st>>>(
      lay, (const float*)packed, origins, dirs, pose12, bins, freqs, radius, S, M, d_raw_sigma, d_raw_rgb, ray_stride,
      stf, gst, pose_acc);
  int rc = star_check_launch();
  if (rc) return rc;
  for (int l = 0; l < MIP_NBASE && !rc; ++l) {
    const float* G = gst + lay.g_base[l] * M;
    float* dW = grad_flat + lay.m_w[l];
    float* db = grad_flat + lay.m_b[l];
    const int K = lay.K[l];
    if (l == 0) {
      rc = star_f32_dw(G, MIP_W, stf + lay.s_enc * M, MIP_KXP, MIP_KX, M, dW, K, db, st);
    } else if (l == MIP_SKIP) {
      rc = star_f32_dw(G, MIP_W, stf + lay.s_enc * M, MIP_KXP, MIP_KX, M, dW, K, db, st);
      if (!rc) rc = star_f32_dw(G, MIP_W, stf + lay.s_in[l] * M, MIP_W, MIP_W, M, dW + MIP_KX, K, nullptr, st);
    } else {
      rc = star_f32_dw(G, MIP_W, stf + lay.s_in[l] * M, MIP_W, MIP_W, M, dW, K, db, st);
    }
  }
  if (rc) return rc;
  {
    const float* G = gst + lay.g_h0 * M;
    float* dW = grad_flat + lay.m_w[MIP_L_H0];
    const int K = MIP_KD + MIP_W;
    rc = star_f32_dw(G, MIP_WH, stf + lay.s_dir * M, MIP_KDP, MIP_KD, M, dW, K, grad_flat + lay.m_b[MIP_L_H0], st);
    if (!rc) rc = star_f32_dw(G, MIP_WH, stf + lay.s_base * M, MIP_W, MIP_W, M, dW + MIP_KD, K, nullptr, st);
    if (!rc)
      rc = star_f32_dw(gst + lay.g_h1 * M, MIP_WH, stf + lay.s_h0 * M, MIP_WH, MIP_WH, M, grad_flat + lay.m_w[MIP_L_H1],
                       MIP_WH, grad_flat + lay.m_b[MIP_L_H1], st);
  }
  if (rc) return rc;
  int hb = (int)((M + 2047) / 2048);
  if (hb > 296) hb = 296;
  mip_head_grad_kernel<<<hb, 256, 0, st>>>(lay, stf, d_raw_sigma, d_raw_rgb, ray_stride, S, M, grad_flat);
  return star_check_launch();
}
