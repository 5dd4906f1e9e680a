// fp32 tier of the per-sample NeRF MLP (models/nerf.py:112-179, models/resnet.py:51-59,103-110):
// fused  pose transform (star__.py:160-199) -> positional encoding (embedder.py:81-112) ->
// ResNet-FC trunk -> heads, forward and backward, on CUDA cores (no tensor cores: this tier is the
// 1e-4-absolute path of the north star and the on-device check for the bf16 tcgen05 tier).
//
// One CTA = one tile of 64 samples, 256 threads.  Activations live in shared memory transposed
// ([feature][sample], row stride 68) so that the 8 rows a warp needs are one broadcast float4 pair;
// weights are streamed from L2 in 16-deep K slabs with cp.async double buffering; every thread owns
// an 8 (samples) x 8 (features) register tile; the residual stream stays in registers across layers.
#include "star_common.cuh"
#include "mlp_layout.h"

#include "mlp_f32_device.cuh"

// ------------------------------------------------------------------------------ encoding helpers
struct SampleGeom {
  float p[3], d[3];   // object-frame point and view direction
};

__device__ __forceinline__ SampleGeom load_geom(const StarPtsSrc& pts, const float* __restrict__ viewdirs,
                                                const float* __restrict__ pose12, int64_t gi, int S) {
  SampleGeom g;
  const int64_t r = gi / S;
  float px, py, pz;
  star_load_pt(pts, gi, r, px, py, pz);
  const float dx = viewdirs[r * 3 + 0], dy = viewdirs[r * 3 + 1], dz = viewdirs[r * 3 + 2];
  if (pose12 != nullptr) {   // p' = R p + t, d' = R d   (star__.py:165-180 / pypose Act :191-196)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g.p[i] = pose12[i * 3 + 0] * px + pose12[i * 3 + 1] * py + pose12[i * 3 + 2] * pz + pose12[9 + i];
      g.d[i] = pose12[i * 3 + 0] * dx + pose12[i * 3 + 1] * dy + pose12[i * 3 + 2] * dz;
    }
  } else {
    g.p[0] = px; g.p[1] = py; g.p[2] = pz;
    g.d[0] = dx; g.d[1] = dy; g.d[2] = dz;
  }
  return g;
}

// ================================================================================ forward kernel
template <bool STASH>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_fwd_f32_kernel(MlpLayout lay, const float* __restrict__ packed, const StarPtsSrc pts,
                   const float* __restrict__ viewdirs, const float* __restrict__ pose12,
                   const float* __restrict__ sc_xyz, const float* __restrict__ sc_dir, int S, int64_t M,
                   float* __restrict__ raw_alpha, float* __restrict__ raw_rgb, int64_t ray_stride,
                   float* __restrict__ stash) {
  extern __shared__ __align__(16) float smem[];
  float* A0 = smem;
  float* A1 = A0 + A_ROWS * AS;
  float* Ed = A1 + A_ROWS * AS;        // [32][AS] encoded view direction
  float* Wbuf = Ed + 32 * AS;          // [2][KS][256]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TM - 1) / TM;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * TM;
    // ---- pose transform + positional encoding into A0 rows [0,64) and Ed rows [0,32)
    {
      const int m = tid & 63, part = tid >> 6;
      const int64_t gi = m0 + m;
      if (gi < M) {
        const SampleGeom g = load_geom(pts, viewdirs, pose12, gi, S);
        if (part == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            A0[c * AS + m] = g.p[c] * (sc_xyz ? sc_xyz[c] : 1.f);
            Ed[c * AS + m] = g.d[c] * (sc_dir ? sc_dir[c] : 1.f);
          }
        }
        for (int k = part; k < lay.L_xyz; k += 4) {
          const float f = exp2f((float)k);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(g.p[c] * f, &sn, &cs);
            const int j = 3 + 6 * k + c;
            A0[j * AS + m] = sn * (sc_xyz ? sc_xyz[j] : 1.f);
            A0[(j + 3) * AS + m] = cs * (sc_xyz ? sc_xyz[j + 3] : 1.f);
          }
        }
        for (int k = part; k < lay.L_dir; k += 4) {
          const float f = exp2f((float)k);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(g.d[c] * f, &sn, &cs);
            const int j = 3 + 6 * k + c;
            Ed[j * AS + m] = sn * (sc_dir ? sc_dir[j] : 1.f);
            Ed[(j + 3) * AS + m] = cs * (sc_dir ? sc_dir[j + 3] : 1.f);
          }
        }
      } else {
        for (int j = part; j < 64; j += 4) A0[j * AS + m] = 0.f;
        for (int j = part; j < 32; j += 4) Ed[j * AS + m] = 0.f;
      }
      // zero padding rows
      for (int j = lay.in_xyz + part; j < 64; j += 4) A0[j * AS + m] = 0.f;
      for (int j = lay.in_dir + part; j < 32; j += 4) Ed[j * AS + m] = 0.f;
    }
    __syncthreads();
    if (STASH) {   // In_0 = enc_xyz [M][64]
      float* dst = stash + lay.L[0].s_in * M;
      for (int i = tid; i < TM * 64; i += NTHREADS) {
        const int m = i >> 6, j = i & 63;
        if (m0 + m < M) dst[(m0 + m) * 64 + j] = A0[j * AS + m];
      }
    }

    float x[8][8];   // residual stream
    float* cur = A0;
    float* nxt = A1;
    for (int l = 0; l < lay.n_layers; ++l) {
      const MlpLayer& ly = lay.L[l];
      const float* bias = packed + ly.p_b;
      if (ly.kind != LK_VIEWS) {
        float acc[8][8];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) acc[mi][nj] = 0.f;
        gemm_tile<2>(cur, packed + ly.p_wt, STAR_W, ly.Kpad, acc, Wbuf, tid);
        float bv[8];
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) bv[nj] = bias[col_of(lane, nj)];
        float outv[8][8];
        if (ly.kind == LK_IN) {
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) { x[mi][nj] = acc[mi][nj] + bv[nj]; outv[mi][nj] = fmaxf(x[mi][nj], 0.f); }
        } else if (ly.kind == LK_FC0) {
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) outv[mi][nj] = fmaxf(acc[mi][nj] + bv[nj], 0.f);
        } else if (ly.kind == LK_FC1) {
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) { x[mi][nj] += acc[mi][nj] + bv[nj]; outv[mi][nj] = fmaxf(x[mi][nj], 0.f); }
        } else {  // LK_OUT, LK_FEAT: affine only
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) outv[mi][nj] = acc[mi][nj] + bv[nj];
        }
        // next layer's input: smem (transposed) + stash (row-major, coalesced float4)
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) nxt[col_of(lane, nj) * AS + warp * 8 + mi] = outv[mi][nj];
        if (STASH) {
          const MlpLayer& nl = lay.L[l + 1];
          float* dst = stash + nl.s_in * M;
#pragma unroll
          for (int mi = 0; mi < 8; ++mi) {
            const int64_t gi = m0 + warp * 8 + mi;
            if (gi < M) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                *reinterpret_cast<float4*>(dst + gi * nl.Kpad + lane * 4 + 128 * j) =
                    make_float4(outv[mi][4 * j], outv[mi][4 * j + 1], outv[mi][4 * j + 2], outv[mi][4 * j + 3]);
            }
          }
        }
        if (ly.kind == LK_OUT) {   // alpha head (nerf.py:151)
          const float* aw = packed + lay.p_alpha_w;
          float awv[8];
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) awv[nj] = aw[col_of(lane, nj)];
          const float ab = packed[lay.p_alpha_b];
#pragma unroll
          for (int mi = 0; mi < 8; ++mi) {
            float p = 0.f;
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) p = fmaf(outv[mi][nj], awv[nj], p);
            p = warp_sum(p);
            const int64_t gi = m0 + warp * 8 + mi;
            if (lane == 0 && gi < M) raw_alpha[(gi / S) * ray_stride + (gi % S)] = p + ab;
          }
        }
        if (ly.kind == LK_FEAT) {  // append encoded dirs (nerf.py:153): rows 256..287
          for (int i = tid; i < 32 * TM; i += NTHREADS) {
            const int j = i >> 6, m = i & 63;
            nxt[(STAR_W + j) * AS + m] = Ed[j * AS + m];
          }
          if (STASH) {
            const MlpLayer& nl = lay.L[l + 1];
            float* dst = stash + nl.s_in * M;
            for (int i = tid; i < 32 * TM; i += NTHREADS) {
              const int m = i >> 5, j = i & 31;
              if (m0 + m < M) dst[(m0 + m) * nl.Kpad + STAR_W + j] = Ed[j * AS + m];
            }
          }
        }
      } else {
        // views layer 283 -> 128 with ReLU, then rgb head (nerf.py:155-159)
        float acc[8][4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) acc[mi][nj] = 0.f;
        gemm_tile<1>(cur, packed + ly.p_wt, STAR_WV, ly.Kpad, acc, Wbuf, tid);
        const float* rw = packed + lay.p_rgb_w;
        float h2[8][4];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          const float b = bias[lane * 4 + nj];
#pragma unroll
          for (int mi = 0; mi < 8; ++mi) h2[mi][nj] = fmaxf(acc[mi][nj] + b, 0.f);
        }
        if (STASH) {
          float* dst = stash + lay.s_h2 * M;
#pragma unroll
          for (int mi = 0; mi < 8; ++mi) {
            const int64_t gi = m0 + warp * 8 + mi;
            if (gi < M)
              *reinterpret_cast<float4*>(dst + gi * STAR_WV + lane * 4) =
                  make_float4(h2[mi][0], h2[mi][1], h2[mi][2], h2[mi][3]);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float wv[4];
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) wv[nj] = rw[c * STAR_WV + lane * 4 + nj];
          const float rb = packed[lay.p_rgb_b + c];
#pragma unroll
          for (int mi = 0; mi < 8; ++mi) {
            float p = 0.f;
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) p = fmaf(h2[mi][nj], wv[nj], p);
            p = warp_sum(p);
            const int64_t gi = m0 + warp * 8 + mi;
            if (lane == 0 && gi < M) raw_rgb[((gi / S) * ray_stride + (gi % S)) * 3 + c] = p + rb;
          }
        }
      }
      float* t = cur; cur = nxt; nxt = t;
      __syncthreads();
    }
  }
}

// ================================================================================ backward (dX chain)
// Walks the layers in reverse for one tile of 64 samples, producing G_l = dL/d(output of layer l)
// for every GEMM layer into `gst` ([layer][M][N]) -- consumed by the dW kernel -- and, for dynamic
// objects, the pose accumulators (see star_b200.h).  ReLU masks come from the stashed activations.
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_bwd_f32_kernel(MlpLayout lay, const float* __restrict__ packed, const StarPtsSrc pts,
                   const float* __restrict__ viewdirs, const float* __restrict__ pose12,
                   const float* __restrict__ sc_xyz, const float* __restrict__ sc_dir, int S, int64_t M,
                   const float* __restrict__ d_raw_alpha, const float* __restrict__ d_raw_rgb, int64_t ray_stride,
                   const float* __restrict__ stash, float* __restrict__ gst, float* __restrict__ pose_acc) {
  extern __shared__ __align__(16) float smem[];
  float* G0 = smem;                    // [256][AS]
  float* G1 = G0 + STAR_W * AS;        // [256][AS]
  float* Wbuf = G1 + STAR_W * AS;      // [2][KS][256]
  float* s_dr = Wbuf + 2 * KS * 256;   // d_raw_rgb [64][4] (c<3) + d_raw_alpha in [m][3]
  __shared__ float s_pose[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TM - 1) / TM;
  const int NL = lay.n_layers;
  const MlpLayer& lv = lay.L[NL - 1];   // views
  const MlpLayer& lf = lay.L[NL - 2];   // feature
  const MlpLayer& lo = lay.L[NL - 3];   // lin_out
  float pacc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) pacc[i] = 0.f;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * TM;
    if (tid < TM) {
      const int64_t gi = m0 + tid;
      float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (gi < M) {
        const int64_t o = (gi / S) * ray_stride + (gi % S);
        a = d_raw_alpha[o];
        c0 = d_raw_rgb[o * 3 + 0]; c1 = d_raw_rgb[o * 3 + 1]; c2 = d_raw_rgb[o * 3 + 2];
      }
      s_dr[tid * 4 + 0] = c0; s_dr[tid * 4 + 1] = c1; s_dr[tid * 4 + 2] = c2; s_dr[tid * 4 + 3] = a;
    }
    __syncthreads();
    // ---- rgb head + ReLU of the views layer:  G_views[m][n] = (sum_c d_rgb[m][c] Wr[c][n]) * [h2 > 0]
    {
      const float* rw = packed + lay.p_rgb_w;
      const float* h2s = stash + lay.s_h2 * M;
      float* gdst = gst + lv.g_out * M;
      float w0[4], w1[4], w2[4];
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        w0[nj] = rw[0 * STAR_WV + lane * 4 + nj];
        w1[nj] = rw[1 * STAR_WV + lane * 4 + nj];
        w2[nj] = rw[2 * STAR_WV + lane * 4 + nj];
      }
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int m = warp * 8 + mi;
        const int64_t gi = m0 + m;
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < M) h = *reinterpret_cast<const float4*>(h2s + gi * STAR_WV + lane * 4);
        const float hv[4] = {h.x, h.y, h.z, h.w};
        float g[4];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          const float v = s_dr[m * 4 + 0] * w0[nj] + s_dr[m * 4 + 1] * w1[nj] + s_dr[m * 4 + 2] * w2[nj];
          g[nj] = hv[nj] > 0.f ? v : 0.f;
          G0[(lane * 4 + nj) * AS + m] = g[nj];
        }
        if (gi < M) *reinterpret_cast<float4*>(gdst + gi * STAR_WV + lane * 4) = make_float4(g[0], g[1], g[2], g[3]);
      }
    }
    __syncthreads();
    // ---- d_feat = G_views * Wv[:, 0:256]     (reduction over the 128 view units)
    float acc[8][8];
    auto zero_acc = [&]() {
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) acc[mi][nj] = 0.f;
    };
    // store a gradient tile: smem (transposed) for the next GEMM and gstash (row-major) for dW
    auto put_grad = [&](float* Gs, const MlpLayer& ly) {
      float* gdst = gst + ly.g_out * M;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) Gs[col_of(lane, nj) * AS + warp * 8 + mi] = acc[mi][nj];
        const int64_t gi = m0 + warp * 8 + mi;
        if (gi < M) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<float4*>(gdst + gi * STAR_W + lane * 4 + 128 * j) =
                make_float4(acc[mi][4 * j], acc[mi][4 * j + 1], acc[mi][4 * j + 2], acc[mi][4 * j + 3]);
        }
      }
    };
    // multiply acc by [stashed activation > 0]
    auto relu_mask = [&](const MlpLayer& ly_in) {
      const float* src = stash + ly_in.s_in * M;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int64_t gi = m0 + warp * 8 + mi;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gi < M) h = *reinterpret_cast<const float4*>(src + gi * ly_in.Kpad + lane * 4 + 128 * j);
          if (!(h.x > 0.f)) acc[mi][4 * j + 0] = 0.f;
          if (!(h.y > 0.f)) acc[mi][4 * j + 1] = 0.f;
          if (!(h.z > 0.f)) acc[mi][4 * j + 2] = 0.f;
          if (!(h.w > 0.f)) acc[mi][4 * j + 3] = 0.f;
        }
      }
    };

    zero_acc();
    gemm_tile<2>(G0, packed + lv.p_wb, lv.Kpad, STAR_WV, acc, Wbuf, tid);
    // direction-encoding gradient d_ed[m][0:32] = G_views * Wv[:, 256:288]  (only objects need it)
    float d_ed[8];  // this thread: sample m = tid & 63, columns (tid >> 6) * 8 .. +8
    if (pose12 != nullptr) {
      const int m = tid & 63, j0 = (tid >> 6) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) d_ed[j] = 0.f;
      const float* wb = packed + lv.p_wb;
      for (int n = 0; n < STAR_WV; ++n) {
        const float g = G0[n * AS + m];
#pragma unroll
        for (int j = 0; j < 8; ++j) d_ed[j] = fmaf(g, wb[(int64_t)n * lv.Kpad + STAR_W + j0 + j], d_ed[j]);
      }
    }
    __syncthreads();
    put_grad(G1, lf);                       // G_feature = d_feat
    __syncthreads();
    // ---- d_h = G_feature * Wf + d_alpha * w_alpha
    zero_acc();
    gemm_tile<2>(G1, packed + lf.p_wb, lf.Kpad, STAR_W, acc, Wbuf, tid);
    {
      const float* aw = packed + lay.p_alpha_w;
#pragma unroll
      for (int nj = 0; nj < 8; ++nj) {
        const float w = aw[col_of(lane, nj)];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) acc[mi][nj] = fmaf(s_dr[(warp * 8 + mi) * 4 + 3], w, acc[mi][nj]);
      }
    }
    put_grad(G0, lo);                       // G_lin_out = d_h
    __syncthreads();
    // ---- d_x = (G_lin_out * Wo) * [a_o > 0]
    zero_acc();
    gemm_tile<2>(G0, packed + lo.p_wb, lo.Kpad, STAR_W, acc, Wbuf, tid);
    relu_mask(lo);
    float dx[8][8];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int nj = 0; nj < 8; ++nj) dx[mi][nj] = acc[mi][nj];
    // ---- residual blocks in reverse
    for (int b = lay.n_blocks - 1; b >= 0; --b) {
      const MlpLayer& l0 = lay.L[1 + 2 * b];
      const MlpLayer& l1 = lay.L[2 + 2 * b];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) acc[mi][nj] = dx[mi][nj];
      put_grad(G1, l1);                     // G_fc1 = d_x
      __syncthreads();
      zero_acc();
      gemm_tile<2>(G1, packed + l1.p_wb, l1.Kpad, STAR_W, acc, Wbuf, tid);
      relu_mask(l1);                        // * [r_b > 0]
      put_grad(G0, l0);                     // G_fc0 = d_net
      __syncthreads();
      zero_acc();
      gemm_tile<2>(G0, packed + l0.p_wb, l0.Kpad, STAR_W, acc, Wbuf, tid);
      relu_mask(l0);                        // * [a_b > 0]
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) dx[mi][nj] += acc[mi][nj];
    }
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int nj = 0; nj < 8; ++nj) acc[mi][nj] = dx[mi][nj];
    put_grad(G1, lay.L[0]);                 // G_lin_in = d_x0
    __syncthreads();

    // ---- pose gradients (objects only): d_e = G_lin_in * W_in  -> encoding Jacobian -> accumulators
    if (pose12 != nullptr) {
      // d_ed partial sums live in 4 thread groups (columns j0..j0+8): park them in G0 rows [0,32)
      {
        const int m = tid & 63, j0 = (tid >> 6) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) G0[(j0 + j) * AS + m] = d_ed[j];
      }
      // d_e[m][0:64]: thread (m = tid & 63, quarter q = tid >> 6) computes columns q*16 .. +16 -> G0 rows [32,96)
      {
        const int m = tid & 63, q = tid >> 6;
        float de[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) de[j] = 0.f;
        const float* wb = packed + lay.L[0].p_wb;   // [256][64]
        for (int n = 0; n < STAR_W; ++n) {
          const float g = G1[n * AS + m];
          const float4* w4 = reinterpret_cast<const float4*>(wb + n * 64 + q * 16);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w = w4[j4];
            de[4 * j4 + 0] = fmaf(g, w.x, de[4 * j4 + 0]);
            de[4 * j4 + 1] = fmaf(g, w.y, de[4 * j4 + 1]);
            de[4 * j4 + 2] = fmaf(g, w.z, de[4 * j4 + 2]);
            de[4 * j4 + 3] = fmaf(g, w.w, de[4 * j4 + 3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) G0[(32 + q * 16 + j) * AS + m] = de[j];
      }
      __syncthreads();
      if (tid < TM) {
        const int m = tid;
        const int64_t gi = m0 + m;
        if (gi < M) {
          const SampleGeom sg = load_geom(pts, viewdirs, pose12, gi, S);
          const int64_t r = gi / S;
          float g[3], h[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float a = G0[(32 + c) * AS + m] * (sc_xyz ? sc_xyz[c] : 1.f);
            for (int k = 0; k < lay.L_xyz; ++k) {
              const float f = exp2f((float)k);
              float sn, cs;
              sincosf(sg.p[c] * f, &sn, &cs);
              const int j = 3 + 6 * k + c;
              a += f * (cs * G0[(32 + j) * AS + m] * (sc_xyz ? sc_xyz[j] : 1.f) -
                        sn * G0[(32 + j + 3) * AS + m] * (sc_xyz ? sc_xyz[j + 3] : 1.f));
            }
            g[c] = a;
            float bq = G0[c * AS + m] * (sc_dir ? sc_dir[c] : 1.f);
            for (int k = 0; k < lay.L_dir; ++k) {
              const float f = exp2f((float)k);
              float sn, cs;
              sincosf(sg.d[c] * f, &sn, &cs);
              const int j = 3 + 6 * k + c;
              bq += f * (cs * G0[j * AS + m] * (sc_dir ? sc_dir[j] : 1.f) -
                         sn * G0[(j + 3) * AS + m] * (sc_dir ? sc_dir[j + 3] : 1.f));
            }
            h[c] = bq;
          }
          float p0, p1, p2;
          star_load_pt(pts, gi, gi / S, p0, p1, p2);
          const float pw[3] = {p0, p1, p2};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            pacc[i] += g[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) pacc[3 + i * 3 + j] += g[i] * pw[j];
          }
          pacc[12] += sg.p[1] * g[2] - sg.p[2] * g[1];
          pacc[13] += sg.p[2] * g[0] - sg.p[0] * g[2];
          pacc[14] += sg.p[0] * g[1] - sg.p[1] * g[0];
          // the view direction is per ray: its gradient is the sum over the ray's samples, so each
          // sample contributes its own h (dL/dd' of that sample) against the same d.
          const float dw[3] = {viewdirs[r * 3 + 0], viewdirs[r * 3 + 1], viewdirs[r * 3 + 2]};
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) pacc[15 + i * 3 + j] += h[i] * dw[j];
          pacc[24] += sg.d[1] * h[2] - sg.d[2] * h[1];
          pacc[25] += sg.d[2] * h[0] - sg.d[0] * h[2];
          pacc[26] += sg.d[0] * h[1] - sg.d[1] * h[0];
        }
      }
    }
    __syncthreads();
  }
  if (pose12 != nullptr) {
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

// ================================================================================ dW / db
// dW[n][k] += sum_m G[m][n] * In[m][k] over this CTA's slice of samples; db[n] += sum_m G[m][n].
// Grid: (Kpad/64, N/64, splits).  64x64 output tile, 256 threads, 4x4 per thread, 32-sample slabs.
__global__ void __launch_bounds__(256)
dw_f32_kernel(const float* __restrict__ G, int N, const float* __restrict__ In, int Kpad, int K, int64_t M,
              float* __restrict__ dW, int ldw, float* __restrict__ db) {
  __shared__ float Gs[32][64 + 4];
  __shared__ float Xs[32][64 + 4];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tn = (tid >> 4) * 4, tk = (tid & 15) * 4;
  const int64_t per = ((M + gridDim.z - 1) / gridDim.z + 31) / 32 * 32;
  const int64_t mb = (int64_t)blockIdx.z * per, me = (mb + per < M) ? mb + per : M;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;   // threads with tid < 64 and blockIdx.x == 0 accumulate db[n0 + tid]
  for (int64_t ms = mb; ms < me; ms += 32) {
    for (int i = tid; i < 32 * 16; i += 256) {
      const int row = i >> 4, c4 = (i & 15) * 4;
      const int64_t m = ms + row;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), xv = g;
      if (m < me) {
        g = *reinterpret_cast<const float4*>(G + m * N + n0 + c4);
        if (k0 + c4 < Kpad) xv = *reinterpret_cast<const float4*>(In + m * Kpad + k0 + c4);
      }
      Gs[row][c4] = g.x; Gs[row][c4 + 1] = g.y; Gs[row][c4 + 2] = g.z; Gs[row][c4 + 3] = g.w;
      Xs[row][c4] = xv.x; Xs[row][c4 + 1] = xv.y; Xs[row][c4 + 2] = xv.z; Xs[row][c4 + 3] = xv.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int mm = 0; mm < 32; ++mm) {
      const float4 g = *reinterpret_cast<const float4*>(&Gs[mm][tn]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[mm][tk]);
      const float gv[4] = {g.x, g.y, g.z, g.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xx[j], acc[i][j]);
    }
    if (db != nullptr && blockIdx.x == 0 && tid < 64) {
#pragma unroll 8
      for (int mm = 0; mm < 32; ++mm) bsum += Gs[mm][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + i, k = k0 + tk + j;
      if (k < K) atomicAdd(&dW[(int64_t)n * ldw + k], acc[i][j]);
    }
  if (db != nullptr && blockIdx.x == 0 && tid < 64) atomicAdd(&db[n0 + tid], bsum);
}

// dW[n][k] (row stride ldw) += G^T In for one GEMM layer (or one K-segment of it); db (may be NULL) += column sums of G
int star_f32_dw(const float* G, int N, const float* In, int Kpad, int K, int64_t M, float* dW, int ldw, float* db,
                cudaStream_t st) {
  int splits = (int)((M + 8191) / 8192);
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  dim3 grid(Kpad / 64 + (Kpad % 64 ? 1 : 0), N / 64, splits);
  dw_f32_kernel<<<grid, 256, 0, st>>>(G, N, In, Kpad, K, M, dW, ldw, db);
  return star_check_launch();
}

// Head gradients: d alpha_w[k] = sum_m d_alpha[m] h[m][k]; d rgb_w[c][n] = sum_m d_rgb[m][c] h2[m][n]; biases.
__global__ void __launch_bounds__(256)
head_grad_f32_kernel(MlpLayout lay, const float* __restrict__ stash, const float* __restrict__ d_raw_alpha,
                     const float* __restrict__ d_raw_rgb, int64_t ray_stride, int S, int64_t M,
                     float* __restrict__ grad_flat) {
  const int tid = threadIdx.x;
  const int64_t per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t mb = (int64_t)blockIdx.x * per, me = (mb + per < M) ? mb + per : M;
  const float* h = stash + lay.L[lay.n_layers - 2].s_in * M;   // input of feature_linear = h
  const float* h2 = stash + lay.s_h2 * M;
  float aw = 0.f, ab = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  const int n2 = tid & 127;
  for (int64_t m = mb; m < me; ++m) {
    const int64_t o = (m / S) * ray_stride + (m % S);
    const float da = d_raw_alpha[o];
    aw = fmaf(da, h[m * STAR_W + tid], aw);
    if (tid == 0) ab += da;
    if (tid < 128) {
      const float c0 = d_raw_rgb[o * 3 + 0], c1 = d_raw_rgb[o * 3 + 1], c2 = d_raw_rgb[o * 3 + 2];
      const float v = h2[m * STAR_WV + n2];
      r0 = fmaf(c0, v, r0); r1 = fmaf(c1, v, r1); r2 = fmaf(c2, v, r2);
      if (tid == 0) { b0 += c0; b1 += c1; b2 += c2; }
    }
  }
  atomicAdd(&grad_flat[lay.m_alpha_w + tid], aw);
  if (tid < 128) {
    atomicAdd(&grad_flat[lay.m_rgb_w + 0 * STAR_WV + n2], r0);
    atomicAdd(&grad_flat[lay.m_rgb_w + 1 * STAR_WV + n2], r1);
    atomicAdd(&grad_flat[lay.m_rgb_w + 2 * STAR_WV + n2], r2);
  }
  if (tid == 0) {
    atomicAdd(&grad_flat[lay.m_alpha_b], ab);
    atomicAdd(&grad_flat[lay.m_rgb_b + 0], b0);
    atomicAdd(&grad_flat[lay.m_rgb_b + 1], b1);
    atomicAdd(&grad_flat[lay.m_rgb_b + 2], b2);
  }
}

// ================================================================================ weight packing
__global__ void pack_f32_kernel(MlpLayout lay, const float* __restrict__ master, float* __restrict__ packed) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= lay.n_packed) return;
  float v = 0.f;
  bool done = false;
  for (int l = 0; l < lay.n_layers && !done; ++l) {
    const MlpLayer& ly = lay.L[l];
    if (i >= ly.p_wt && i < ly.p_wb) {           // W^T [Kpad][N]
      const int64_t j = i - ly.p_wt;
      const int k = (int)(j / ly.N), n = (int)(j % ly.N);
      v = (k < ly.K) ? master[ly.m_w + (int64_t)n * ly.K + k] : 0.f;
      done = true;
    } else if (i >= ly.p_wb && i < ly.p_b) {     // W padded [N][Kpad]
      const int64_t j = i - ly.p_wb;
      const int n = (int)(j / ly.Kpad), k = (int)(j % ly.Kpad);
      v = (k < ly.K) ? master[ly.m_w + (int64_t)n * ly.K + k] : 0.f;
      done = true;
    } else if (i >= ly.p_b && i < ly.p_b + ly.N) {
      v = master[ly.m_b + (i - ly.p_b)];
      done = true;
    }
  }
  if (!done) {
    if (i >= lay.p_alpha_w && i < lay.p_alpha_w + STAR_W) v = master[lay.m_alpha_w + (i - lay.p_alpha_w)];
    else if (i == lay.p_alpha_b) v = master[lay.m_alpha_b];
    else if (i >= lay.p_rgb_w && i < lay.p_rgb_w + 3 * STAR_WV) v = master[lay.m_rgb_w + (i - lay.p_rgb_w)];
    else if (i >= lay.p_rgb_b && i < lay.p_rgb_b + 3) v = master[lay.m_rgb_b + (i - lay.p_rgb_b)];
  }
  packed[i] = v;
}

// ================================================================================ host side
static const size_t FWD_SMEM = sizeof(float) * (2 * A_ROWS * AS + 32 * AS + 2 * KS * 256);
static const size_t BWD_SMEM = sizeof(float) * (2 * STAR_W * AS + 2 * KS * 256 + TM * 4);

int star_f32_pack(const MlpLayout& lay, const float* master, void* packed, cudaStream_t st) {
  const int threads = 256;
  const int blocks = (int)((lay.n_packed + threads - 1) / threads);
  pack_f32_kernel<<<blocks, threads, 0, st>>>(lay, master, (float*)packed);
  return star_check_launch();
}

static int grid_for_tiles(int64_t ntiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(ntiles < sms ? ntiles : sms);
}

int star_f32_forward(const MlpLayout& lay, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                     const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, float* raw_alpha,
                     float* raw_rgb, int64_t ray_stride, void* stash, cudaStream_t st) {
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TM - 1) / TM;
  const int grid = grid_for_tiles(ntiles);
  if (stash != nullptr) {
    cudaFuncSetAttribute(mlp_fwd_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    mlp_fwd_f32_kernel<true><<<grid, NTHREADS, FWD_SMEM, st>>>(lay, (const float*)packed, pts, viewdirs, pose12,
                                                               sc_xyz, sc_dir, S, M, raw_alpha, raw_rgb, ray_stride,
                                                               (float*)stash);
  } else {
    cudaFuncSetAttribute(mlp_fwd_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    mlp_fwd_f32_kernel<false><<<grid, NTHREADS, FWD_SMEM, st>>>(lay, (const float*)packed, pts, viewdirs, pose12,
                                                                sc_xyz, sc_dir, S, M, raw_alpha, raw_rgb, ray_stride,
                                                                nullptr);
  }
  return star_check_launch();
}

int star_f32_backward(const MlpLayout& lay, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                      const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S,
                      const float* d_raw_alpha, const float* d_raw_rgb, int64_t ray_stride, const void* stash,
                      void* workspace, float* grad_flat, float* pose_acc, cudaStream_t st) {
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TM - 1) / TM;
  float* gst = (float*)workspace;
  cudaFuncSetAttribute(mlp_bwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
  mlp_bwd_f32_kernel<<<grid_for_tiles(ntiles), NTHREADS, BWD_SMEM, st>>>(
      lay, (const float*)packed, pts, viewdirs, pose12, sc_xyz, sc_dir, S, M, d_raw_alpha, d_raw_rgb, ray_stride,
      (const float*)stash, gst, pose_acc);
  int rc = star_check_launch();
  if (rc) return rc;
  for (int l = 0; l < lay.n_layers; ++l) {
    const MlpLayer& ly = lay.L[l];
    rc = star_f32_dw(gst + ly.g_out * M, ly.N, (const float*)stash + ly.s_in * M, ly.Kpad, ly.K, M,
                     grad_flat + ly.m_w, ly.K, grad_flat + ly.m_b, st);
    if (rc) return rc;
  }
  int hb = (int)((M + 2047) / 2048);
  if (hb > 296) hb = 296;
  head_grad_f32_kernel<<<hb, 256, 0, st>>>(lay, (const float*)stash, d_raw_alpha, d_raw_rgb, ray_stride, S, M,
                                           grad_flat);
  return star_check_launch();
}
