// what sits either side of the render path inside one training step.
//   - photometric loss  nn.MSELoss(rgb0, target) + nn.MSELoss(rgb, target) and mse2psnr, with the gradients of both
//     rgb maps written by the same pass                     (train_online__.py:158-166, models/rendering__.py:18-23)
//   - DS-NeRF depth / sigma losses                                                      (models/loss.py:4-87)
//   - gradient clipping by global norm + Adam over flat parameter runs
//     (train_online__.py:333-353 torch.optim.Adam(betas=(0.9, 0.999)); Trainer(gradient_clip_val=1.0) :1170)
// All of it is HBM-bound streaming work: 16-byte accesses, fp64 accumulation of the reductions, and a deterministic
// two-level reduction (per-block partials added in block order by the last block to finish).
#include "star_common.cuh"

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 148 * 8;
constexpr int RED_MAX_NV = 2;
// workspace: [0,4) ticket counter, [8,16) running fp64 total, [16, ...) per-block partials
constexpr size_t WS_PARTIALS_OFF = 16;
constexpr size_t WS_BYTES = WS_PARTIALS_OFF + sizeof(double) * RED_MAX_NV * RED_MAX_BLOCKS;

// Every block leaves its partial sums in `partials`; the block that draws the last ticket adds them in block order.
// Returns true in thread 0 of that block with the totals in v[].  `counter` must be 0 at launch and is left at 0.
template <int NV>
__device__ bool grid_reduce(double (&v)[NV], double* partials, unsigned* counter) {
  __shared__ double sh[NV][RED_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = warp_sum_d(v[i]);
    if (lane == 0) sh[i][w] = v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int k = 0; k < RED_THREADS / 32; ++k) s += sh[i][k];
      partials[(size_t)blockIdx.x * NV + i] = s;
    }
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += RED_THREADS) s += __ldcg(partials + (size_t)b * NV + i);
    s = warp_sum_d(s);
    if (lane == 0) sh[i][w] = s;
  }
  __syncthreads();
  if (threadIdx.x != 0) return false;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int k = 0; k < RED_THREADS / 32; ++k) s += sh[i][k];
    v[i] = s;
  }
  *counter = 0;
  return true;
}

inline int red_blocks(int64_t items_per_thread_unit) {
  int64_t b = (items_per_thread_unit + RED_THREADS - 1) / RED_THREADS;
  if (b < 1) b = 1;
  if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
  return (int)b;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// ------------------------------------------------------------------------------------------ photometric loss
// out = [mse0, mse, psnr0, psnr, mse0 + mse]; g_rgb0 / g_rgb = d(mse0 + mse)/d(rgb0 | rgb) = 2 (x - t) / n.
__device__ __forceinline__ float psnr_of(float mse) { return __fdiv_rn(__fmul_rn(-10.f, logf(mse)), logf(10.f)); }

__global__ void __launch_bounds__(RED_THREADS)
photometric_kernel(const float* __restrict__ rgb0, const float* __restrict__ rgb, const float* __restrict__ target,
                   int64_t n, float* __restrict__ g0, float* __restrict__ g1, double* partials, unsigned* counter,
                   float* __restrict__ out) {
  double v[2] = {0.0, 0.0};
  const float sc = __fdiv_rn(2.f, (float)n);
  for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RED_THREADS) {
    const float t = target[i];
    if (rgb0 != nullptr) {
      const float d = __fsub_rn(rgb0[i], t);
      v[0] += (double)__fmul_rn(d, d);
      if (g0 != nullptr) g0[i] = __fmul_rn(d, sc);
    }
    const float d = __fsub_rn(rgb[i], t);
    v[1] += (double)__fmul_rn(d, d);
    if (g1 != nullptr) g1[i] = __fmul_rn(d, sc);
  }
  if (grid_reduce<2>(v, partials, counter)) {
    const float m0 = (float)(v[0] / (double)n), m1 = (float)(v[1] / (double)n);
    out[0] = m0;
    out[1] = m1;
    out[2] = rgb0 != nullptr ? psnr_of(m0) : 0.f;
    out[3] = psnr_of(m1);
    out[4] = rgb0 != nullptr ? __fadd_rn(m0, m1) : m1;
  }
}

// ------------------------------------------------------------------------------------------ DS-NeRF depth loss
// models/loss.py:4-10: mask = near < gt < far; mean over the masked rays of ((depth - gt) / gt)^2.  out = [loss, count]
// (an empty mask gives 0/0 = NaN, as torch.mean of an empty tensor does).
__global__ void __launch_bounds__(RED_THREADS)
depth_loss_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ gt, int64_t R, float near_, float far_,
                      double* partials, unsigned* counter, float* __restrict__ out) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < R; i += (int64_t)gridDim.x * RED_THREADS) {
    const float g = gt[i];
    if (g < far_ && g > near_) {
      const float q = __fdiv_rn(__fsub_rn(depth[i], g), g);
      v[0] += (double)__fmul_rn(q, q);
      v[1] += 1.0;
    }
  }
  if (grid_reduce<2>(v, partials, counter)) {
    out[0] = (float)(v[0] / v[1]);
    out[1] = (float)v[1];
  }
}

__global__ void depth_loss_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ gt, int64_t R,
                                      float near_, float far_, const float* __restrict__ out2,
                                      const float* __restrict__ g_out, float* __restrict__ g_depth) {
  const float sc = __fdiv_rn(g_out[0], out2[1]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = gt[i];
    float r = 0.f;
    if (g < far_ && g > near_) {
      const float q = __fdiv_rn(__fsub_rn(depth[i], g), g);
      r = __fmul_rn(__fdiv_rn(__fmul_rn(2.f, q), g), sc);
    }
    g_depth[i] = r;
  }
}

// ------------------------------------------------------------------------------------------ DS-NeRF sigma loss
// models/loss.py:13-66 (and :70-87 for the per-ray variant): per masked ray
//   sum_s -log(w'_s) * exp(-(z_s - D)^2 / (2 err)) * dists_s,   w' = w <= 0 ? eps : w,
// averaged over the masked rays.  One warp per ray, 16-byte loads when S % 4 == 0.
__device__ __forceinline__ float sigma_term(float w, float z, float dist, float D, float two_err) {
  const float wp = w <= 0.f ? STAR_EPS_F32 : w;
  const float dz = __fsub_rn(z, D);
  const float e = expf(__fdiv_rn(-__fmul_rn(dz, dz), two_err));
  return __fmul_rn(__fmul_rn(-logf(wp), e), dist);
}
__device__ __forceinline__ float sigma_grad(float w, float z, float dist, float D, float two_err, float sc) {
  if (w <= 0.f) return 0.f;   // torch.where routes the gradient of the replaced entries to the constant
  const float dz = __fsub_rn(z, D);
  const float e = expf(__fdiv_rn(-__fmul_rn(dz, dz), two_err));
  return __fmul_rn(__fdiv_rn(-__fmul_rn(e, dist), w), sc);
}

template <bool VEC>
__global__ void __launch_bounds__(RED_THREADS)
sigma_loss_fwd_kernel(const float* __restrict__ w, const float* __restrict__ z, const float* __restrict__ dists,
                      const float* __restrict__ depths, int64_t R, int S, float near_, float far_, float two_err,
                      double* partials, unsigned* counter, float* __restrict__ out, float* __restrict__ per_ray) {
  double v[2] = {0.0, 0.0};
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)RED_THREADS + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * RED_THREADS) >> 5;
  for (int64_t r = warp0; r < R; r += nwarps) {
    const float D = depths[r];
    const bool in = D < far_ && D > near_;
    if (!in && per_ray == nullptr) continue;
    float s = 0.f;
    const int64_t base = r * (int64_t)S;
    if (VEC) {
      const float4* w4 = reinterpret_cast<const float4*>(w + base);
      const float4* z4 = reinterpret_cast<const float4*>(z + base);
      const float4* d4 = reinterpret_cast<const float4*>(dists + base);
      const int n4 = S / 4;
      for (int k = lane; k < n4; k += 64) {      // two 16-byte loads per array in flight
        const bool two = k + 32 < n4;
        const float4 a = __ldcs(w4 + k), b = __ldcs(z4 + k), c = __ldcs(d4 + k);
        float4 a2 = make_float4(1.f, 1.f, 1.f, 1.f), b2 = a2, c2 = make_float4(0.f, 0.f, 0.f, 0.f);   // log(1) * 0 = 0
        if (two) a2 = __ldcs(w4 + k + 32), b2 = __ldcs(z4 + k + 32), c2 = __ldcs(d4 + k + 32);
        s += sigma_term(a.x, b.x, c.x, D, two_err);
        s += sigma_term(a.y, b.y, c.y, D, two_err);
        s += sigma_term(a.z, b.z, c.z, D, two_err);
        s += sigma_term(a.w, b.w, c.w, D, two_err);
        if (two) {
          s += sigma_term(a2.x, b2.x, c2.x, D, two_err);
          s += sigma_term(a2.y, b2.y, c2.y, D, two_err);
          s += sigma_term(a2.z, b2.z, c2.z, D, two_err);
          s += sigma_term(a2.w, b2.w, c2.w, D, two_err);
        }
      }
    } else {
      for (int k = lane; k < S; k += 32) s += sigma_term(w[base + k], z[base + k], dists[base + k], D, two_err);
    }
    s = warp_sum(s);
    if (lane == 0) {
      if (per_ray != nullptr) per_ray[r] = s;
      if (in) {
        v[0] += (double)s;
        v[1] += 1.0;
      }
    }
  }
  if (grid_reduce<2>(v, partials, counter)) {
    out[0] = (float)(v[0] / v[1]);
    out[1] = (float)v[1];
  }
}

template <bool VEC>
__global__ void sigma_loss_bwd_kernel(const float* __restrict__ w, const float* __restrict__ z,
                                      const float* __restrict__ dists, const float* __restrict__ depths, int64_t R,
                                      int S, float near_, float far_, float two_err, const float* __restrict__ out2,
                                      const float* __restrict__ g_out, bool g_per_ray, float* __restrict__ g_w) {
  const float sc0 = g_per_ray ? 0.f : __fdiv_rn(g_out[0], out2[1]);
  const int per = VEC ? S / 4 : S;
  const int64_t total = R * (int64_t)per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / per;
    const float D = depths[r];
    const bool in = g_per_ray || (D < far_ && D > near_);
    const float sc = g_per_ray ? g_out[r] : sc0;
    if (VEC) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(w) + i), b = __ldcs(reinterpret_cast<const float4*>(z) + i),
                     c = __ldcs(reinterpret_cast<const float4*>(dists) + i);
        o.x = sigma_grad(a.x, b.x, c.x, D, two_err, sc);
        o.y = sigma_grad(a.y, b.y, c.y, D, two_err, sc);
        o.z = sigma_grad(a.z, b.z, c.z, D, two_err, sc);
        o.w = sigma_grad(a.w, b.w, c.w, D, two_err, sc);
      }
      __stcs(reinterpret_cast<float4*>(g_w) + i, o);
    } else {
      g_w[i] = in ? sigma_grad(w[i], z[i], dists[i], D, two_err, sc) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------ clip + Adam
constexpr int ADAM_MAX_SEGS = 16;
constexpr int ADAM_CHUNK = 4096;   // elements of the virtual concatenation handled by one block iteration

struct AdamBatch {
  StarAdamSeg s[ADAM_MAX_SEGS];
  long long start[ADAM_MAX_SEGS + 1];   // prefix sums of s[].n
  int count;
};

// Applies f1(i) / f4(i) (scalar / 4 consecutive elements, i = element index within the segment) to [a, b) of segment
// `sg`; the 16-byte path is taken where every listed pointer shares the alignment of the first.
template <typename F1, typename F4>
__device__ __forceinline__ void for_piece(long long a, long long b, bool same_align, unsigned mis, F1 f1, F4 f4) {
  long long head = b - a;
  long long nvec = 0;
  if (same_align) {
    head = (4 - (long long)((mis + (unsigned)(a & 3)) & 3)) & 3;
    if (head > b - a) head = b - a;
    nvec = (b - a - head) >> 2;
  }
  for (long long i = a + threadIdx.x; i < a + head; i += blockDim.x) f1(i);
  for (long long k = threadIdx.x; k < nvec; k += blockDim.x) f4(a + head + 4 * k);
  for (long long i = a + head + 4 * nvec + threadIdx.x; i < b; i += blockDim.x) f1(i);
}

template <typename BODY>
__device__ __forceinline__ void for_chunks(const AdamBatch& t, BODY body) {
  const long long total = t.start[t.count];
  const long long nchunks = (total + ADAM_CHUNK - 1) / ADAM_CHUNK;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    long long lo = c * ADAM_CHUNK, hi = lo + ADAM_CHUNK;
    if (hi > total) hi = total;
    int sg = 0;
    while (t.start[sg + 1] <= lo) ++sg;
    while (lo < hi) {
      const long long e = t.start[sg + 1] < hi ? t.start[sg + 1] : hi;
      body(sg, lo - t.start[sg], e - t.start[sg]);
      lo = e;
      ++sg;
    }
  }
}

__device__ __forceinline__ unsigned mis_of(const void* p) { return (unsigned)(((uintptr_t)p >> 2) & 3); }

__global__ void __launch_bounds__(RED_THREADS)
grad_sqnorm_kernel(const __grid_constant__ AdamBatch t, double* partials, unsigned* counter, double* total) {
  double v[1] = {0.0};
  for_chunks(t, [&](int sg, long long a, long long b) {
    const float* g = t.s[sg].grad;
    for_piece(a, b, ((uintptr_t)g & 3) == 0, mis_of(g),
              [&](long long i) { const float x = g[i]; v[0] += (double)x * (double)x; },
              [&](long long i) {
                const float4 x = *reinterpret_cast<const float4*>(g + i);
                v[0] += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
              });
  });
  if (grid_reduce<1>(v, partials, counter)) *total += v[0];
}

// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6))
__device__ __forceinline__ float clip_coef(const double* sqnorm, float max_norm) {
  if (sqnorm == nullptr) return 1.f;
  const float total = (float)sqrt(*sqnorm);
  const float c = __fdiv_rn(max_norm, __fadd_rn(total, 1e-6f));
  return c > 1.f ? 1.f : c;    // NaN propagates, as torch.clamp(max=1) does
}

__global__ void __launch_bounds__(RED_THREADS)
grad_scale_kernel(const __grid_constant__ AdamBatch t, const double* __restrict__ sqnorm, float max_norm) {
  const float coef = clip_coef(sqnorm, max_norm);
  for_chunks(t, [&](int sg, long long a, long long b) {
    float* g = t.s[sg].grad;
    for_piece(a, b, ((uintptr_t)g & 3) == 0, mis_of(g), [&](long long i) { g[i] = __fmul_rn(g[i], coef); },
              [&](long long i) {
                float4 x = *reinterpret_cast<float4*>(g + i);
                x.x = __fmul_rn(x.x, coef), x.y = __fmul_rn(x.y, coef), x.z = __fmul_rn(x.z, coef),
                x.w = __fmul_rn(x.w, coef);
                *reinterpret_cast<float4*>(g + i) = x;
              });
  });
}

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), one rounding per torch op:
//   exp_avg.lerp_(g, 1 - b1); exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2);
//   denom = exp_avg_sq.sqrt() / sqrt(1 - b2^t) + eps;  p.addcdiv_(exp_avg, denom, value=-lr / (1 - b1^t))
struct AdamHyper {
  float w1, b2, w2, eps;
};
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float coef, const AdamHyper& h,
                                      float step_size, float bc2s) {
  g = __fmul_rn(g, coef);
  m = fmaf(h.w1, __fsub_rn(g, m), m);
  v = fmaf(h.w2, __fmul_rn(g, g), __fmul_rn(v, h.b2));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), h.eps);
  p = fmaf(-step_size, __fdiv_rn(m, denom), p);
}

template <bool WRITE_G>
__global__ void __launch_bounds__(RED_THREADS)
adam_kernel(const __grid_constant__ AdamBatch t, AdamHyper h, const double* __restrict__ sqnorm, float max_norm) {
  const float coef = clip_coef(sqnorm, max_norm);
  for_chunks(t, [&](int sg, long long a, long long b) {
    const StarAdamSeg& s = t.s[sg];
    float* p = s.param;
    float* g = s.grad;
    float* m = s.exp_avg;
    float* v = s.exp_avg_sq;
    const float step_size = s.step_size, bc2s = s.bc2_sqrt;
    const unsigned mis = mis_of(p);
    const bool same = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 3) == 0 && mis_of(g) == mis &&
                      mis_of(m) == mis && mis_of(v) == mis;
    for_piece(a, b, same, mis,
              [&](long long i) {
                float pp = p[i], mm = m[i], vv = v[i];
                const float gg = g[i];
                adam1(pp, gg, mm, vv, coef, h, step_size, bc2s);
                p[i] = pp, m[i] = mm, v[i] = vv;
                if (WRITE_G) g[i] = __fmul_rn(gg, coef);
              },
              [&](long long i) {
                float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
                       vv = *reinterpret_cast<float4*>(v + i);
                float4 gg = *reinterpret_cast<const float4*>(g + i);
                adam1(pp.x, gg.x, mm.x, vv.x, coef, h, step_size, bc2s);
                adam1(pp.y, gg.y, mm.y, vv.y, coef, h, step_size, bc2s);
                adam1(pp.z, gg.z, mm.z, vv.z, coef, h, step_size, bc2s);
                adam1(pp.w, gg.w, mm.w, vv.w, coef, h, step_size, bc2s);
                *reinterpret_cast<float4*>(p + i) = pp;
                *reinterpret_cast<float4*>(m + i) = mm;
                *reinterpret_cast<float4*>(v + i) = vv;
                if (WRITE_G) {
                  gg.x = __fmul_rn(gg.x, coef), gg.y = __fmul_rn(gg.y, coef), gg.z = __fmul_rn(gg.z, coef),
                  gg.w = __fmul_rn(gg.w, coef);
                  *reinterpret_cast<float4*>(g + i) = gg;
                }
              });
  });
}

// Splits the host segment list into launches of <= ADAM_MAX_SEGS segments.
template <typename LAUNCH>
int for_batches(const StarAdamSeg* segs, int n_segs, bool need_state, LAUNCH launch) {
  if (n_segs < 0) return STAR_E_BAD_SHAPE;
  if (n_segs > 0 && !segs) return STAR_E_NULL;
  for (int i = 0; i < n_segs; ++i) {
    if (segs[i].n < 0) return STAR_E_BAD_SHAPE;
    if (segs[i].n > 0 && (!segs[i].grad || (need_state && (!segs[i].param || !segs[i].exp_avg || !segs[i].exp_avg_sq))))
      return STAR_E_NULL;
  }
  for (int i0 = 0; i0 < n_segs; i0 += ADAM_MAX_SEGS) {
    AdamBatch t;
    t.count = 0;
    t.start[0] = 0;
    for (int i = i0; i < n_segs && i < i0 + ADAM_MAX_SEGS; ++i) {
      if (segs[i].n == 0) continue;
      t.s[t.count] = segs[i];
      t.start[t.count + 1] = t.start[t.count] + segs[i].n;
      ++t.count;
    }
    if (t.count == 0) continue;
    long long blocks = (t.start[t.count] + ADAM_CHUNK - 1) / ADAM_CHUNK;
    if (blocks > RED_MAX_BLOCKS) blocks = RED_MAX_BLOCKS;
    launch(t, (int)blocks);
    const int rc = star_check_launch();
    if (rc) return rc;
  }
  return STAR_OK;
}

int ws_reset(void* ws, cudaStream_t st) {
  const cudaError_t e = cudaMemsetAsync(ws, 0, WS_PARTIALS_OFF, st);
  if (e != cudaSuccess) {
    g_star_last_cuda_error = (int)e;
    return STAR_E_CUDA;
  }
  return STAR_OK;
}

}  // namespace

extern "C" size_t star_train_ws_bytes(void) { return WS_BYTES; }

extern "C" int star_photometric_loss(const float* rgb0, const float* rgb, const float* target, int64_t n, float* out5,
                                     float* g_rgb0, float* g_rgb, void* ws, void* stream) {
  if (!rgb || !target || !out5 || !ws) return STAR_E_NULL;
  if (n < 1) return STAR_E_BAD_SHAPE;
  if (!aligned16(ws)) return STAR_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ws_reset(ws, st)) return rc;
  photometric_kernel<<<red_blocks(n), RED_THREADS, 0, st>>>(rgb0, rgb, target, n, g_rgb0, g_rgb,
                                                            (double*)((char*)ws + WS_PARTIALS_OFF), (unsigned*)ws, out5);
  return star_check_launch();
}

extern "C" int star_depth_loss_forward(const float* depth, const float* gt_depth, int64_t R, float near_, float far_,
                                       float* out2, void* ws, void* stream) {
  if (!depth || !gt_depth || !out2 || !ws) return STAR_E_NULL;
  if (R < 1) return STAR_E_BAD_SHAPE;
  if (!aligned16(ws)) return STAR_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ws_reset(ws, st)) return rc;
  depth_loss_fwd_kernel<<<red_blocks(R), RED_THREADS, 0, st>>>(depth, gt_depth, R, near_, far_,
                                                               (double*)((char*)ws + WS_PARTIALS_OFF), (unsigned*)ws, out2);
  return star_check_launch();
}

extern "C" int star_depth_loss_backward(const float* depth, const float* gt_depth, int64_t R, float near_, float far_,
                                        const float* out2, const float* g_out, float* g_depth, void* stream) {
  if (!depth || !gt_depth || !out2 || !g_out || !g_depth) return STAR_E_NULL;
  if (R < 1) return STAR_E_BAD_SHAPE;
  depth_loss_bwd_kernel<<<red_blocks(R), RED_THREADS, 0, (cudaStream_t)stream>>>(depth, gt_depth, R, near_, far_, out2,
                                                                                 g_out, g_depth);
  return star_check_launch();
}

extern "C" int star_sigma_loss_forward(const float* weights, const float* z_vals, const float* dists, const float* depths,
                                       int64_t R, int S, float near_, float far_, float err, float* out2, float* per_ray,
                                       void* ws, void* stream) {
  if (!weights || !z_vals || !dists || !depths || !out2 || !ws) return STAR_E_NULL;
  if (R < 1 || S < 1) return STAR_E_BAD_SHAPE;
  if (!aligned16(ws)) return STAR_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ws_reset(ws, st)) return rc;
  const bool vec = S % 4 == 0 && aligned16(weights) && aligned16(z_vals) && aligned16(dists);
  const int blocks = red_blocks(R * 32);
  double* partials = (double*)((char*)ws + WS_PARTIALS_OFF);
  const float two_err = 2.f * err;
  if (vec)
    sigma_loss_fwd_kernel<true><<<blocks, RED_THREADS, 0, st>>>(weights, z_vals, dists, depths, R, S, near_, far_, two_err,
                                                                partials, (unsigned*)ws, out2, per_ray);
  else
    sigma_loss_fwd_kernel<false><<<blocks, RED_THREADS, 0, st>>>(weights, z_vals, dists, depths, R, S, near_, far_,
                                                                 two_err, partials, (unsigned*)ws, out2, per_ray);
  return star_check_launch();
}

extern "C" int star_sigma_loss_backward(const float* weights, const float* z_vals, const float* dists,
                                        const float* depths, int64_t R, int S, float near_, float far_, float err,
                                        const float* out2, const float* g_out, int g_per_ray, float* g_weights,
                                        void* stream) {
  if (!weights || !z_vals || !dists || !depths || !g_out || !g_weights || (!out2 && !g_per_ray)) return STAR_E_NULL;
  if (R < 1 || S < 1) return STAR_E_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = S % 4 == 0 && aligned16(weights) && aligned16(z_vals) && aligned16(dists) && aligned16(g_weights);
  const int64_t total = R * (int64_t)(vec ? S / 4 : S);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const float two_err = 2.f * err;
  if (vec)
    sigma_loss_bwd_kernel<true><<<(int)blocks, 256, 0, st>>>(weights, z_vals, dists, depths, R, S, near_, far_, two_err,
                                                             out2, g_out, g_per_ray != 0, g_weights);
  else
    sigma_loss_bwd_kernel<false><<<(int)blocks, 256, 0, st>>>(weights, z_vals, dists, depths, R, S, near_, far_, two_err,
                                                              out2, g_out, g_per_ray != 0, g_weights);
  return star_check_launch();
}

extern "C" int star_grad_sqnorm(const StarAdamSeg* segs, int n_segs, void* ws, void* stream) {
  if (!ws) return STAR_E_NULL;
  if (!aligned16(ws)) return STAR_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ws_reset(ws, st)) return rc;
  return for_batches(segs, n_segs, false, [&](const AdamBatch& t, int blocks) {
    grad_sqnorm_kernel<<<blocks, RED_THREADS, 0, st>>>(t, (double*)((char*)ws + WS_PARTIALS_OFF), (unsigned*)ws,
                                                       (double*)((char*)ws + 8));
  });
}

extern "C" const double* star_grad_sqnorm_result(const void* ws) {
  return ws ? (const double*)((const char*)ws + 8) : nullptr;
}

extern "C" int star_grad_scale(const StarAdamSeg* segs, int n_segs, const double* sqnorm, float max_norm, void* stream) {
  if (!sqnorm) return STAR_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  return for_batches(segs, n_segs, false, [&](const AdamBatch& t, int blocks) {
    grad_scale_kernel<<<blocks, RED_THREADS, 0, st>>>(t, sqnorm, max_norm);
  });
}

extern "C" int star_adam_step(const StarAdamSeg* segs, int n_segs, double beta1, double beta2, double eps,
                              const double* sqnorm, float max_norm, int write_back_grads, void* stream) {
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) return STAR_E_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  // torch forms 1 - beta in double and rounds once when the scalar meets the fp32 tensor
  AdamHyper h;
  h.w1 = (float)(1.0 - beta1), h.b2 = (float)beta2, h.w2 = (float)(1.0 - beta2), h.eps = (float)eps;
  const bool wb = write_back_grads != 0 && sqnorm != nullptr;
  return for_batches(segs, n_segs, true, [&](const AdamBatch& t, int blocks) {
    if (wb)
      adam_kernel<true><<<blocks, RED_THREADS, 0, st>>>(t, h, sqnorm, max_norm);
    else
      adam_kernel<false><<<blocks, RED_THREADS, 0, st>>>(t, h, sqnorm, max_norm);
  });
}
