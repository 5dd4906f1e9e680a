// Backward pass of the tensor-core MLP tier (sm_100a): gradients of (raw_alpha, raw_rgb) w.r.t. every MLP weight
// and, for object nets, the pose accumulators (models/nerf.py:112-179, models/resnet.py:51-59, models/star__.py:160-199).
//
// Inputs: the activation stash written by the training forward (mlp_tc.cu; 16 KB swizzled blocks, mlp_tc_layout.h).
//   1. mlp_bwd_tc_kernel  -- the dX chain, one persistent CTA per SM over 128-sample tiles, same roles as the
//      forward: G_l = dL/d(output of GEMM layer l) flows backwards through tcgen05 GEMMs against the TRANSPOSED
//      weight stream; ReLU masks come from the stash; the residual gradient dx lives in TMEM columns 0..255 (fp32)
//      and is read-modified-written by the epilogue; every G_l is written to the gradient stash in the same block
//      format.  Object nets also run G_in * W_in and G_views * W_views[:, dirs] and fold the result through the
//      positional-encoding Jacobian into the 27 pose accumulators of include/star_b200.h.
//   2. dw_tc_kernel       -- dW_l = G_l^T A_l and db_l = G_l^T 1 as tcgen05 GEMMs whose K dimension is the sample
//      index: stash blocks are consumed directly as MN-major operands (a block is [samples][64 features]), split over
//      CTAs by (layer, 128-row half, sample range), partial results reduced into the flat fp32 gradient with atomics.
//   3. head_grad_tc_kernel -- alpha_linear / rgb_linear gradients on CUDA cores (tiny).
#include <cuda_fp16.h>
#include "mlp_tc_device.cuh"
#include "mip_tc_layout.h"

#define BK_PRE 0     // G_views from d_rgb, rgb_linear and the relu(h2) mask (no MMA before it)
#define BK_FEAT 1    // G_feat = T                     (+ dirs part of the pose gradient)
#define BK_OUT 2     // G_out  = X + d_alpha * w_alpha
#define BK_X0 3      // dx = T * mask                  -> X, G_fc1(last block)
#define BK_FC1 4     // G_fc0 = T * mask
#define BK_FC0 5     // dx = X + T * mask              -> X, G_fc1(previous block) / G_in
#define BK_IN 6      // pose gradient through the xyz encoding Jacobian (objects only)

struct BwdStage {    // one weight K-block = one ring stage = one group of MMAs
  uint32_t w_off, bytes;
  int a_kb, N, tcol, accum, nk, last;   // A K-block, MMA N, TMEM column, accumulate flag of the first MMA, K-steps, last of phase
};
struct BwdPhase {
  int kind, nch;       // epilogue kind, number of 64-column chunks it produces (A K-blocks)
  int mask_blk, g_blk; // stash block of the mask source / gradient-stash block of the output (-1: none)
  int mask_layer;      // >= 0: the mask comes from the bit masks of this forward layer (TC_MASK_BYTES each) instead
  int n_wait;          // A K-blocks the following MMA group must wait for (= nch)
};
#define BWD_MAX_STAGES 64
#define BWD_MAX_PHASES 24

struct BwdSmem {
  uint32_t A, W, small, tab, bars, tmem_ptr, pacc, total;
};
#define BWD_PACC_FLOATS 32      // per epilogue warp: [0, 15) translation / rotation sums of the positions, [16, 28) of the directions
__host__ __device__ static inline BwdSmem bwd_smem_layout(uint32_t small_bytes) {
  BwdSmem s;
  uint32_t o = 0;
  s.A = o; o += 4 * TC_KB_BYTES;
  s.W = o; o += TC_NS * TC_STAGE_BYTES;
  s.small = o; o += small_bytes;
  s.tab = o; o += BWD_MAX_STAGES * sizeof(BwdStage) + BWD_MAX_PHASES * sizeof(BwdPhase) + 16;
  s.bars = o; o += 40 * 8;
  s.tmem_ptr = o; o += 16;
  s.pacc = o; o += TC_EPI_WARPS * BWD_PACC_FLOATS * sizeof(float);
  s.total = o + 1024;
  return s;
}

// Sums 16 per-lane values across the warp with 16 shuffles (each butterfly step halves the values a lane carries): the
// lane pair (2 i, 2 i + 1) returns the total of v[i].
__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
  float a[8], b[4], c[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (h16 ? v[8 + i] : v[i]) + __shfl_xor_sync(STAR_FULL_MASK, h16 ? v[i] : v[8 + i], 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = (h8 ? a[4 + i] : a[i]) + __shfl_xor_sync(STAR_FULL_MASK, h8 ? a[i] : a[4 + i], 8);
#pragma unroll
  for (int i = 0; i < 2; ++i) c[i] = (h4 ? b[2 + i] : b[i]) + __shfl_xor_sync(STAR_FULL_MASK, h4 ? b[i] : b[2 + i], 4);
  float d = (h2 ? c[1] : c[0]) + __shfl_xor_sync(STAR_FULL_MASK, h2 ? c[0] : c[1], 2);
  d += __shfl_xor_sync(STAR_FULL_MASK, d, 1);
  return d;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// d(encoding column)/d(input component) for the 16 columns [C0, C0+16): column col < 3 -> component col, 1;
// sin(2^k x_c) -> 2^k cos; cos(2^k x_c) -> -2^k sin.  comp[j] = component index, dv[j] = derivative (0 for padding).
template <int C0, int NV>
__device__ __forceinline__ void encode_slice_grad(const float (&p)[3], const float* __restrict__ sc, float (&dv)[16]) {
  constexpr int LAST = (C0 + 15 < NV) ? C0 + 15 : NV - 1;
  constexpr int K_LO = (C0 < 3) ? 0 : (C0 - 3) / 6;
  constexpr int K_HI = (LAST < 3) ? -1 : (LAST - 3) / 6;
#pragma unroll
  for (int j = 0; j < 16; ++j) dv[j] = (C0 + j < 3) ? 1.f : 0.f;
  if (K_HI >= K_LO) {
    float sn[3], cs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sincosf(p[c] * (float)(1 << K_LO), &sn[c], &cs[c]);
#pragma unroll
    for (int k = K_LO; k <= K_HI; ++k) {
      const float f = (float)(1 << k);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = C0 + j;
        if (col >= 3 && col < NV && (col - 3) / 6 == k) {
          const int w = (col - 3) % 6;
          dv[j] = (w < 3) ? f * cs[w % 3] : -f * sn[w % 3];
        }
      }
      if (k < K_HI) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float s2 = 2.f * sn[c] * cs[c], c2 = 1.f - 2.f * sn[c] * sn[c];
          sn[c] = s2; cs[c] = c2;
        }
      }
    }
  }
  if (sc != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dv[j] *= sc[C0 + j];
  }
}
// component (0..2) of encoding column col
__host__ __device__ constexpr int enc_comp(int col) { return col < 3 ? col : ((col - 3) % 6) % 3; }

template <int C0, int NV>
__device__ __forceinline__ void fold_jacobian(const uint32_t (&r)[16], const float (&p)[3], const float* sc, float (&g)[3]) {
  float dv[16];
  encode_slice_grad<C0, NV>(p, sc, dv);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (C0 + j < NV) g[enc_comp(C0 + j)] = fmaf(__uint_as_float(r[j]), dv[j], g[enc_comp(C0 + j)]);
}

// ============================================================================================ dX chain
// Operand formats of the backward pass.  kind::f16 does not take mixed formats (bf16 x fp16 raises an illegal-instruction
// fault on B200: tools/umma_probe.cu), so the gradients G_l share the format of the weights / stashed activations they are
// multiplied with: bf16 on the bf16 tier; fp16 on the fp16 tier.  dL/d(activation) of a 4096-ray batch reaches 1e-8,
// below fp16's range, so the fp16 tier back-propagates SCALED gradients: scale = 2^k (exact) with k chosen per call from
// the largest |upstream gradient| (grad_absmax_kernel; no host round trip) such that the scaled maximum lies in [64, 128) --
// 2^9 of headroom to fp16's 65504 for growth along the chain (conversions saturate instead of producing inf), and every
// value down to 2^-20 of the maximum stays a normal fp16 number with 11 significant bits (bf16 keeps 8 everywhere).  The dW
// and pose epilogues multiply by 2^-k.  F16 = false: bf16, scale 1.
__device__ __forceinline__ float grad_scale_of(const float* __restrict__ absmax) {
  const uint32_t bits = __float_as_uint(*absmax);
  const int e = (int)((bits >> 23) & 0xffu);            // biased exponent of max |g| (floor(log2) + 127)
  if (e == 0 || e == 255) return 1.f;                   // zero / subnormal / non-finite upstream gradient: leave it alone
  int k = 6 + 127 - e;                                  // 2^k * max in [64, 128)
  k = k > 96 ? 96 : (k < -96 ? -96 : k);
  return __uint_as_float((uint32_t)(127 + k) << 23);
}

__global__ void grad_absmax_kernel(const float* __restrict__ d_alpha, const float* __restrict__ d_rgb, int S, int64_t M,
                                   int64_t ray_stride, float* __restrict__ absmax) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / S, o = r * ray_stride + (i - r * S);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(d_alpha[o]), fabsf(d_rgb[o * 3])), fmaxf(fabsf(d_rgb[o * 3 + 1]), fabsf(d_rgb[o * 3 + 2]))));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(absmax), __float_as_uint(m));
}

#ifndef TC_WSHARE_DX
// 1: the dX chain's (transposed) weight stream shared by the two CTAs of a cluster, as in the forward kernels.  Built, GPU
// tests green with it, measured NEUTRAL (dX + dW + heads 4.73-4.76 against 4.74 ms per 4096-ray step,
// profiles/r2zz_ab_wshare_dx.txt): the serial MMA / epilogue alternation of this kernel does not wait for weights.
#define TC_WSHARE_DX 0
#endif
#ifndef TC_DX_DEFER_WAIT_ST
#define TC_DX_DEFER_WAIT_ST 0  // 1: one tcgen05.wait::st per phase instead of one per chunk -- measured no gain (bwd 4.68 / 4.82 vs 4.65 / 4.70 ms, profiles/r2z_ab_defer_wait_st.txt)
#endif
#ifndef TC_DX_LD_DEPTH
#define TC_DX_LD_DEPTH 2       // accumulator chunk loads in flight per epilogue thread of the dX chain (1 = load, wait, use)
#endif
template <bool F16, bool POSE, bool WSHARE = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_bwd_tc_kernel(const TcLayout lay, const uint8_t* __restrict__ packed, const StarPtsSrc pts,
                  const float* __restrict__ viewdirs, const float* __restrict__ pose12,
                  const float* __restrict__ sc_xyz, const float* __restrict__ sc_dir, int S, int64_t M,
                  const float* __restrict__ d_raw_alpha, const float* __restrict__ d_raw_rgb, int64_t ray_stride,
                  const uint8_t* __restrict__ stash, uint8_t* __restrict__ gstash, float* __restrict__ pose_acc,
                  const float* __restrict__ absmax, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  tc_mark_begin(dbg);
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  const BwdSmem sl = bwd_smem_layout(lay.small_bytes);
  const uint32_t sA = base + sl.A, sW = base + sl.W, sBars = base + sl.bars;
  float* s_small = reinterpret_cast<float*>(gbase + sl.small);
  BwdStage* stages = reinterpret_cast<BwdStage*>(gbase + sl.tab);
  BwdPhase* phases = reinterpret_cast<BwdPhase*>(gbase + sl.tab + BWD_MAX_STAGES * sizeof(BwdStage));
  int* counts = reinterpret_cast<int*>(gbase + sl.tab + BWD_MAX_STAGES * sizeof(BwdStage) + BWD_MAX_PHASES * sizeof(BwdPhase));
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(gbase + sl.tmem_ptr);
  auto bar = [&](int i) -> uint32_t { return sBars + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // WSHARE (mlp_tc.cu: the two CTAs of a cluster share the weight stream): an even number of trips per cluster; a tile beyond
  // the launch is a ghost (zero gradients, its blocks land in the spare tile of the even-sized stashes)
  const int64_t ntiles = WSHARE ? (((M + TC_M - 1) / TC_M + 1) & ~(int64_t)1) : (M + TC_M - 1) / TC_M;
  const uint32_t crank = WSHARE ? cluster_ctarank() : 0u;
  constexpr bool has_pose = POSE;   // object nets: pose accumulators and two extra GEMMs (compiled out for the static net)
  // accumulator chunk loads in flight per epilogue thread: the second buffer (measured neutral on the static net) would push
  // the object nets' epilogue past the 96 registers 576 threads leave, into spills inside the chunk loop
  constexpr int LD_DEPTH = POSE ? 1 : TC_DX_LD_DEPTH;

  // ---- the per-tile program (identical for every tile): built once by one thread
  if (tid == 0) {
    const int nl = lay.n_layers, V = nl - 1, F = nl - 2, O = nl - 3;
    int ns = 0, np = 0;
    auto stage = [&](uint32_t w_off, uint32_t bytes, int a_kb, int N, int tcol, int accum, int nk, int last) {
      BwdStage& s = stages[ns++];
      s.w_off = w_off; s.bytes = bytes; s.a_kb = a_kb; s.N = N; s.tcol = tcol; s.accum = accum; s.nk = nk; s.last = last;
    };
    auto phase = [&](int kind, int nch, int mask_blk, int g_blk, int mask_layer = -1) {
      BwdPhase& p = phases[np++];
      p.kind = kind; p.nch = nch; p.mask_blk = mask_blk; p.g_blk = g_blk; p.n_wait = nch; p.mask_layer = mask_layer;
    };
    // full layer: N = 256 input features, K = n_out / 64 blocks of the transposed stream
    auto gemm = [&](int l, int tcol) {
      const int nkb = lay.L[l].N / 64;
      for (int kb = 0; kb < nkb; ++kb)
        stage(lay.stream_bytes + lay.L[l].wt_off + (uint32_t)kb * 32768u, 32768u, kb, 256, tcol, kb > 0, 4, kb == nkb - 1);
    };
    phase(BK_PRE, 2, lay.L[V].s_out, lay.L[V].g_out);
    // views: T <- G_v * Wv[:, :256]^T (+ X[0:32] <- G_v * Wv[:, 256:288]^T for objects)
    for (int kb = 0; kb < 2; ++kb)
      stage(lay.stream_bytes + lay.L[V].wt_off + (uint32_t)kb * 32768u, 32768u, kb, 256, 256, kb > 0, 4, !has_pose && kb == 1);
    if (has_pose)
      for (int kb = 0; kb < 2; ++kb)
        stage(lay.stream_bytes + lay.wt_dirs_off + (uint32_t)kb * 8192u, 8192u, kb, 32, 0, kb > 0, 4, kb == 1);
    phase(BK_FEAT, 4, -1, lay.L[F].g_out);
    gemm(F, 0);
    phase(BK_OUT, 4, -1, lay.L[O].g_out);
    gemm(O, 256);
    // the rectified operand of layer x was written by the epilogue of layer x - 1: its bit masks are slot x - 1
    phase(BK_X0, 4, -1, lay.L[O - 1].g_out, O - 1);           // mask relu(x_B) > 0; G of the last fc_1 (or lin_in)
    for (int b = lay.n_blocks - 1; b >= 0; --b) {
      const int l0 = 1 + 2 * b, l1 = 2 + 2 * b;
      gemm(l1, 256);
      phase(BK_FC1, 4, -1, lay.L[l0].g_out, l1 - 1);          // mask relu(net_b) > 0
      gemm(l0, 256);
      phase(BK_FC0, 4, -1, lay.L[l0 - 1].g_out, l0 - 1);      // mask relu(x_b) > 0; G of fc_1(b-1) or lin_in
    }
    if (has_pose) {   // d enc_xyz = G_in * W_in: N = 64 input features, K = 256
      for (int kb = 0; kb < 4; ++kb)
        stage(lay.stream_bytes + lay.L[0].wt_off + (uint32_t)kb * 8192u, 8192u, kb, 64, 256, kb > 0, 4, kb == 3);
      phase(BK_IN, 0, -1, -1);
    }
    counts[0] = ns; counts[1] = np;
  }
  if (warp == TC_EPI_WARPS && lane == 0) {
    for (int i = 0; i < TC_NS; ++i) { mbar_init(bar(BAR_W_FULL(i)), 1); mbar_init(bar(BAR_W_EMPTY(i)), WSHARE ? 2 : 1); }
    for (int i = 0; i < 5; ++i) mbar_init(bar(BAR_A_READY(i)), TC_EPI_WARPS);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_STASH_DONE), 1);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(base + sl.tmem_ptr, TC_TMEM_COLS);
  for (int i = tid; i < lay.small_floats; i += TC_THREADS) s_small[i] = reinterpret_cast<const float*>(packed)[i];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (WSHARE) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int n_stages = counts[0], n_phases = counts[1];

  if (warp == TC_EPI_WARPS) {
    // ======================================================================== weight producer
    // The same thread writes the gradient stash: an A block in shared memory IS the 16 KB gstash block, so once all
    // 16 epilogue warps have published block kb of phase g (a_ready[kb]) one bulk store moves it to global memory.
    // The stores of phase g are issued before the weights of MMA group g + 1 are requested (its first ring stage frees
    // only after an MMA of group g, which itself waits for every a_ready of phase g); stash_done then tells the
    // epilogue of phase g + 1 that the A blocks have been read and may be overwritten.
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, s_par = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint8_t* gs_tile = gstash + (size_t)tile * (size_t)lay.gstash_blocks * TC_BLOCK_BYTES;
        auto stash_phase = [&](int g) {
          const int nch = phases[g].nch, g_blk = phases[g].g_blk;
          if (nch == 0) return;
          for (int kb = 0; kb < nch; ++kb) {
            mbar_wait(bar(BAR_A_READY(kb)), (s_par >> kb) & 1u, dbg, 6);
            s_par ^= 1u << kb;
            if (g_blk >= 0) {
              bulk_s2g(gs_tile + (size_t)(g_blk + kb) * TC_BLOCK_BYTES, sA + (uint32_t)kb * TC_KB_BYTES, TC_KB_BYTES);
              bulk_commit_group();
            }
          }
          bulk_wait_group_read0();
          mbar_arrive(bar(BAR_STASH_DONE));
        };
        int grp = 0;
        bool group_start = true;
        for (int i = 0; i < n_stages; ++i) {
          const BwdStage& s = stages[i];
          if (group_start && grp >= 1) stash_phase(grp - 1);
          group_start = false;
          mbar_wait(bar(BAR_W_EMPTY(stage)), phase ^ 1u, dbg, 1);
          mbar_arrive_expect_tx(bar(BAR_W_FULL(stage)), s.bytes);
          if (WSHARE) {
            const uint32_t half = s.bytes >> 1;
            bulk_g2s_mcast(sW + stage * TC_STAGE_BYTES + crank * half, packed + lay.small_bytes + s.w_off + crank * half, half,
                           bar(BAR_W_FULL(stage)), (uint16_t)3);
          } else {
            bulk_g2s(sW + stage * TC_STAGE_BYTES, packed + lay.small_bytes + s.w_off, s.bytes, bar(BAR_W_FULL(stage)));
          }
          if (++stage == TC_NS) { stage = 0; phase ^= 1u; }
          if (s.last) { ++grp; group_start = true; }
        }
        for (int g = grp - 1; g < n_phases; ++g) stash_phase(g);   // the phases after the last MMA group of the tile
      }
      bulk_wait_group0();
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    // warp-uniform control flow (all lanes wait on the barriers), one elected lane issues the tcgen05 instructions
    {
      uint32_t stage = 0, phase = 0, a_par = 0;
      const uint64_t desc_a0 = umma_desc_sw128(sA), desc_w0 = umma_desc_sw128(sW);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int ph = 0;              // index of the epilogue phase that produced the A operand of the current group
        bool group_start = true;
        for (int i = 0; i < n_stages; ++i) {
          const BwdStage s = stages[i];
          if (group_start) {
            // the A operand (and the free accumulator) of this group need ALL chunks of the previous epilogue
            for (int kb = 0; kb < phases[ph].n_wait; ++kb) {
              mbar_wait(bar(BAR_A_READY(kb)), (a_par >> kb) & 1u, dbg, 2);
              a_par ^= 1u << kb;
            }
            group_start = false;
          }
          mbar_wait(bar(BAR_W_FULL(stage)), phase, dbg, 3);
          tc_fence_after();
          const uint32_t idesc = umma_idesc_16(TC_M, s.N, F16 ? 0 : 1);
          const uint64_t a0 = desc_a0 + (uint64_t)(s.a_kb * (TC_KB_BYTES >> 4));
          const uint64_t b0 = desc_w0 + (uint64_t)(stage * (TC_STAGE_BYTES >> 4));
          if (elect_one_sync()) {
            tc_mma_kblock<4>(tmem_base + (uint32_t)s.tcol, a0, b0, idesc, s.accum ? 1u : 0u);   // every stage has 4 K-steps
            if (WSHARE) tc_commit_mcast(bar(BAR_W_EMPTY(stage)), (uint16_t)3);
            else tc_commit(bar(BAR_W_EMPTY(stage)));
            if (s.last) tc_commit(bar(BAR_ACC_FULL));
          }
          __syncwarp();
          if (++stage == TC_NS) { stage = 0; phase ^= 1u; }
          if (s.last) {
            ++ph;
            group_start = true;
          }
        }
        // the tile's last epilogue phase (G_in for static nets) also arrives on a_ready[]: consume those phases so
        // that the parity bookkeeping stays in step with the barriers across tiles
        for (int kb = 0; ph < n_phases && kb < phases[ph].n_wait; ++kb) {
          mbar_wait(bar(BAR_A_READY(kb)), (a_par >> kb) & 1u, dbg, 5);
          a_par ^= 1u << kb;
        }
      }
    }
  } else {
    // ======================================================================== epilogue warps
    const int q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    uint32_t acc_par = 0, stash_par = 0;
    // Pose sums: reduced over the warp once per tile and kept per WARP in shared memory.  (27 per-thread running sums plus the
    // transformed point / direction kept live across the tile cost this kernel 560 bytes of spills in its chunk loops, and an
    // object net's dX chain -- 8 GEMM groups -- took as long as the static net's 12: profiles/r2m_c4_launches.md.)
    float* s_pacc = reinterpret_cast<float*>(gbase + sl.pacc) + warp * BWD_PACC_FLOATS;
    if (has_pose) s_pacc[lane] = 0.f;
    const float gscale = F16 ? grad_scale_of(absmax) : 1.f;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gi = tile * TC_M + row;
      const bool valid = gi < M;
      float da = 0.f, dc0 = 0.f, dc1 = 0.f, dc2 = 0.f;
      if (valid) {
        const int64_t ray = gi / S;
        const int64_t o = ray * ray_stride + (gi - ray * S);
        da = d_raw_alpha[o] * gscale;         // (exact: a power of two)
        dc0 = d_raw_rgb[o * 3 + 0] * gscale; dc1 = d_raw_rgb[o * 3 + 1] * gscale; dc2 = d_raw_rgb[o * 3 + 2] * gscale;
      }
      const uint8_t* st_tile = stash + (size_t)tile * (size_t)lay.stash_blocks * TC_BLOCK_BYTES;
      uint8_t* gs_tile = gstash + (size_t)tile * (size_t)lay.gstash_blocks * TC_BLOCK_BYTES;
      const uint32_t x_sw = (uint32_t)row & 7u;

      for (int pi = 0; pi < n_phases; ++pi) {
        const BwdPhase P = phases[pi];
        // mask sources of the whole phase are fetched BEFORE waiting for the accumulator (the global-memory latency
        // hides behind the MMAs): the hidden layers' ReLU masks are one 64-bit word per thread (bit 16 kb + j), written
        // by the forward epilogue; relu(h2) of the view layer is tested on its stashed 16-bit activations
        // (> 0 <=> bits != 0 for a ReLU output)
        uint4 mq[2][2];
        uint2 mbits = make_uint2(~0u, ~0u);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          mq[kb][0] = make_uint4(~0u, ~0u, ~0u, ~0u);
          mq[kb][1] = mq[kb][0];
          if (P.mask_blk >= 0 && kb < P.nch) {
            const uint8_t* mb = st_tile + (size_t)(P.mask_blk + kb) * TC_BLOCK_BYTES;
            mq[kb][0] = __ldg(reinterpret_cast<const uint4*>(mb + (uint32_t)row * 128u + ((((uint32_t)(cg * 2)) ^ x_sw) << 4)));
            mq[kb][1] = __ldg(reinterpret_cast<const uint4*>(mb + (uint32_t)row * 128u + ((((uint32_t)(cg * 2 + 1)) ^ x_sw) << 4)));
          }
        }
        if (P.mask_layer >= 0)
          mbits = __ldg(reinterpret_cast<const uint2*>(st_tile + (size_t)lay.mask_blk0 * TC_BLOCK_BYTES +
                                                       (size_t)P.mask_layer * TC_MASK_BYTES) + (cg * TC_M + row));
        if (P.kind != BK_PRE) {
          mbar_wait(bar(BAR_ACC_FULL), acc_par, dbg, 4);
          acc_par ^= 1u;
          tc_fence_after();
        }
        if (P.nch > 0 && !(pi == 0 && tile == (int64_t)blockIdx.x)) {   // the previous phase's A blocks are in the gstash
          mbar_wait(bar(BAR_STASH_DONE), stash_par, dbg, 7);
          stash_par ^= 1u;
        }
        const uint32_t tX = tmem_base + lane_addr + (uint32_t)(cg * TC_CPT);
        const uint32_t tT = tX + 256u;
        if (has_pose && P.kind == BK_IN) {       // d enc_xyz in T[0:64]: 16 columns per thread
          uint32_t r[16];
          tmem_ld16(tT, r);
          tmem_wait_ld();
          // the sample's position and its image in the object frame are formed here, where they are used (not kept live
          // across the tile's phases)
          float pw[3] = {0.f, 0.f, 0.f}, po[3] = {0.f, 0.f, 0.f};
          if (valid) {
            star_load_pt(pts, gi, gi / S, pw[0], pw[1], pw[2]);
#pragma unroll
            for (int i = 0; i < 3; ++i)
              po[i] = pose12[i * 3 + 0] * pw[0] + pose12[i * 3 + 1] * pw[1] + pose12[i * 3 + 2] * pw[2] + pose12[9 + i];
          }
          float g[3] = {0.f, 0.f, 0.f};
          if (cg == 0) fold_jacobian<0, 63>(r, po, sc_xyz, g);
          else if (cg == 1) fold_jacobian<16, 63>(r, po, sc_xyz, g);
          else if (cg == 2) fold_jacobian<32, 63>(r, po, sc_xyz, g);
          else fold_jacobian<48, 63>(r, po, sc_xyz, g);
          float c[16];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            c[i] = valid ? g[i] : 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) c[3 + i * 3 + j] = valid ? g[i] * pw[j] : 0.f;
          }
          c[12] = valid ? po[1] * g[2] - po[2] * g[1] : 0.f;
          c[13] = valid ? po[2] * g[0] - po[0] * g[2] : 0.f;
          c[14] = valid ? po[0] * g[1] - po[1] * g[0] : 0.f;
          c[15] = 0.f;
          const float tot = warp_reduce16(c, lane);
          if ((lane & 1) == 0) s_pacc[lane >> 1] += tot;
          tc_fence_before();
          continue;
        }
        if (P.kind == BK_FEAT && has_pose && cg < 2) {   // d enc_dir in X[0:32]
          uint32_t r[16];
          tmem_ld16(tX, r);
          tmem_wait_ld();
          float dw[3] = {0.f, 0.f, 0.f}, dob[3] = {0.f, 0.f, 0.f};
          if (valid) {
            const int64_t ray = gi / S;
            dw[0] = viewdirs[ray * 3 + 0]; dw[1] = viewdirs[ray * 3 + 1]; dw[2] = viewdirs[ray * 3 + 2];
#pragma unroll
            for (int i = 0; i < 3; ++i) dob[i] = pose12[i * 3 + 0] * dw[0] + pose12[i * 3 + 1] * dw[1] + pose12[i * 3 + 2] * dw[2];
          }
          float h[3] = {0.f, 0.f, 0.f};
          if (cg == 0) fold_jacobian<0, 27>(r, dob, sc_dir, h);
          else fold_jacobian<16, 27>(r, dob, sc_dir, h);
          float c[16];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) c[i * 3 + j] = valid ? h[i] * dw[j] : 0.f;
          c[9] = valid ? dob[1] * h[2] - dob[2] * h[1] : 0.f;
          c[10] = valid ? dob[2] * h[0] - dob[0] * h[2] : 0.f;
          c[11] = valid ? dob[0] * h[1] - dob[1] * h[0] : 0.f;
          c[12] = c[13] = c[14] = c[15] = 0.f;
          const float tot = warp_reduce16(c, lane);
          if ((lane & 1) == 0) s_pacc[16 + (lane >> 1)] += tot;
        }
        // publish one 16-column chunk of this phase's output: zero rows beyond the launch, 16-bit conversion into A block kb
        // (fp16: saturating -- a gradient that outgrows the headroom clamps to 65504 instead of becoming inf), proxy fence,
        // arrival.  (The gstash copy of the block is a bulk store issued by the producer warp once the block is complete.)
        auto publish = [&](const int kb, float (&v)[16]) {
          if (!valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          store_row16<F16, false, true>(sA + (uint32_t)kb * TC_KB_BYTES, row, cg * 2, v);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_A_READY(kb)));
        };
        if (P.kind == BK_PRE) {
          // G of relu(h2) from d_rgb and the rgb head (2 chunks of the 128 view-layer features), masked by the stashed
          // relu(h2) itself (> 0 <=> bits != 0 for a ReLU output)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const int col0 = kb * 64 + cg * TC_CPT;
            const uint4 m0 = mq[kb][0], m1 = mq[kb][1];
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            const float* rw = s_small + lay.off_rgb_w + col0;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[j] = dc0 * rw[j] + dc1 * rw[STAR_WV + j] + dc2 * rw[2 * STAR_WV + j];
              const uint32_t bits = (j & 1) ? (mw[j >> 1] >> 16) : (mw[j >> 1] & 0xffffu);
              if (bits == 0u) v[j] = 0.f;
            }
            publish(kb, v);
          }
          continue;
        }
        // TMEM loads are double buffered: the accumulator chunk of K-block kb + 1 is requested as soon as chunk kb has
        // arrived, so its latency hides behind chunk kb's masking / conversion / stores (fc_0 also fetches the residual
        // gradient's chunk kb then, and waits for both before the add)
        const uint32_t t_acc = (P.kind == BK_OUT) ? tX : tT;
        uint32_t racc[2][16];
        if (LD_DEPTH >= 2) tmem_ld16(t_acc, racc[0]);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb >= P.nch) break;
          const int col0 = kb * 64 + cg * TC_CPT;
          const uint32_t mb16 = ((kb < 2 ? mbits.x : mbits.y) >> ((kb & 1) * 16)) & 0xffffu;
          if (LD_DEPTH < 2) tmem_ld16(t_acc + 64u * (uint32_t)kb, racc[kb & 1]);
          tmem_wait_ld();            // (all of this thread's outstanding loads: chunk kb)
          if (LD_DEPTH >= 2 && kb + 1 < P.nch) tmem_ld16(t_acc + 64u * (uint32_t)(kb + 1), racc[(kb + 1) & 1]);
          uint32_t rx[16];
          if (P.kind == BK_FC0) tmem_ld16(tX + 64u * (uint32_t)kb, rx);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(racc[kb & 1][j]);
          if (P.kind == BK_OUT) {
            const float* aw = s_small + lay.off_alpha_w + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(da, aw[j], v[j]);
          }
          if (P.mask_layer >= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (!((mb16 >> j) & 1u)) v[j] = 0.f;
          }
          if (P.kind == BK_FC0) {      // dx += masked product
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(rx[j]);
          }
          if (P.kind == BK_FC0 || P.kind == BK_X0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) rx[j] = __float_as_uint(v[j]);
            tmem_st16(tX + 64u * (uint32_t)kb, rx);
            // the residual gradient is next touched (by this thread, or by the next tile's MMAs after this phase's last
            // arrival) after the phase: one wait for the phase's stores, before its last block is published
            if (!TC_DX_DEFER_WAIT_ST || kb == P.nch - 1) tmem_wait_st();
          }
          publish(kb, v);
        }
      }
    }
    if (has_pose) {
      const float inv = 1.f / gscale;
      __syncwarp();
      if (lane < 27) atomicAdd(&pose_acc[lane], s_pacc[lane < 15 ? lane : lane + 1] * inv);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (WSHARE) cluster_sync_all();      // no CTA leaves while its peer's commits / copies can still land in it
  tc_mark_end(dbg);
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ============================================================================================ dX chain, pipelined
// Round 2.  In the kernel above every GEMM writes the same accumulator region T (the other one, X, holds the residual
// gradient), so the MMAs of group g + 1 cannot start before the epilogue of group g has read ALL of T, and that epilogue
// cannot start before ALL MMAs of group g are done: MMA (~2.8 k cycles) and epilogue (~3 k) of the 11 groups of a tile
// alternate strictly, which is why the tensor pipe is active 29 % of the time.  Here every N = 256 GEMM is issued as two
// N = 128 HALVES with their own completion barriers (half a -> T[0:128), half b -> T[128:256)); the epilogue consumes half a
// (chunks 0, 1) while the tensor core still computes half b, and the next group's half a starts as soon as chunks 0 and 1
// have been read (T[0:128) is free again) and the operand blocks it needs have been published -- K-block by K-block, like
// the forward kernel.  Extra synchronisation this needs: an operand block may be overwritten by the next epilogue only
// after the LAST MMA that reads it (half b of the same K-block: b_done[kb], a tcgen05.commit per K-block) and after its
// copy to the gradient stash has left shared memory (stash_done[kb], per block).  The weight ring holds 8 stages of
// 16 KB (one half K-block each).
#define BWD2_MAX_STAGES 112
#define BAR2_ACC_B (2 * TC_MAX_NS + 15)
#define BAR2_B_DONE(kb) (2 * TC_MAX_NS + 16 + (kb))
#define BWD2_NS 8
#define BWD2_STAGE_BYTES 16384

struct Bwd2Stage {
  uint32_t w_off, bytes;
  int a_kb, N, tcol, accum;
  int wait0, wait1;      // a_ready barriers to wait for before this stage (-1: none)
  int sig_acc;           // after this stage commit: 1 = accumulator half a complete, 2 = half b (or the whole group) complete
  int sig_bdone;         // >= 0: this is the last stage of the group that reads A block sig_bdone
};
struct Bwd2Phase {
  int kind, nch, mask_blk, g_blk, mask_layer;
  int acc_a, acc_b;      // the preceding group signals half a / half b (phase waits for a before chunk 0, for b before chunk 2)
  int reads_mask;        // A blocks the FOLLOWING group reads (bit kb): they carry a b_done commit
};
struct Bwd2Smem {
  uint32_t A, W, small, tab, bars, tmem_ptr, total;
};
__host__ __device__ static inline Bwd2Smem bwd2_smem_layout(uint32_t small_bytes) {
  Bwd2Smem s;
  uint32_t o = 0;
  s.A = o; o += 4 * TC_KB_BYTES;
  s.W = o; o += BWD2_NS * BWD2_STAGE_BYTES;
  s.small = o; o += small_bytes;
  s.tab = o; o += BWD2_MAX_STAGES * sizeof(Bwd2Stage) + BWD_MAX_PHASES * sizeof(Bwd2Phase) + 16;
  s.bars = o; o += 40 * 8;
  s.tmem_ptr = o; o += 16;
  s.total = o + 1024;
  return s;
}

template <bool F16, bool POSE>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_bwd2_tc_kernel(const TcLayout lay, const uint8_t* __restrict__ packed, const StarPtsSrc pts,
                   const float* __restrict__ viewdirs, const float* __restrict__ pose12,
                   const float* __restrict__ sc_xyz, const float* __restrict__ sc_dir, int S, int64_t M,
                   const float* __restrict__ d_raw_alpha, const float* __restrict__ d_raw_rgb, int64_t ray_stride,
                   const uint8_t* __restrict__ stash, uint8_t* __restrict__ gstash, float* __restrict__ pose_acc,
                   const float* __restrict__ absmax, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  const Bwd2Smem sl = bwd2_smem_layout(lay.small_bytes);
  const uint32_t sA = base + sl.A, sW = base + sl.W, sBars = base + sl.bars;
  float* s_small = reinterpret_cast<float*>(gbase + sl.small);
  Bwd2Stage* stages = reinterpret_cast<Bwd2Stage*>(gbase + sl.tab);
  Bwd2Phase* phases = reinterpret_cast<Bwd2Phase*>(gbase + sl.tab + BWD2_MAX_STAGES * sizeof(Bwd2Stage));
  int* counts = reinterpret_cast<int*>(gbase + sl.tab + BWD2_MAX_STAGES * sizeof(Bwd2Stage) + BWD_MAX_PHASES * sizeof(Bwd2Phase));
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(gbase + sl.tmem_ptr);
  auto bar = [&](int i) -> uint32_t { return sBars + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  constexpr bool has_pose = POSE;

  // ---- the per-tile program: phase p (epilogue) writes the operand blocks that group p (MMAs) reads
  if (tid == 0) {
    const int nl = lay.n_layers, V = nl - 1, F = nl - 2, O = nl - 3;
    int ns = 0, np = 0;
    auto stage = [&](uint32_t w_off, uint32_t bytes, int a_kb, int N, int tcol, int accum, int wait0, int wait1, int sig_acc,
                     int sig_bdone) {
      Bwd2Stage& s = stages[ns++];
      s.w_off = w_off; s.bytes = bytes; s.a_kb = a_kb; s.N = N; s.tcol = tcol; s.accum = accum;
      s.wait0 = wait0; s.wait1 = wait1; s.sig_acc = sig_acc; s.sig_bdone = sig_bdone;
    };
    auto phase = [&](int kind, int nch, int mask_blk, int g_blk, int mask_layer, int acc_a, int acc_b) {
      Bwd2Phase& p = phases[np++];
      p.kind = kind; p.nch = nch; p.mask_blk = mask_blk; p.g_blk = g_blk; p.mask_layer = mask_layer;
      p.acc_a = acc_a; p.acc_b = acc_b; p.reads_mask = 0;
    };
    // N = 256 GEMM of layer l over nkb operand blocks into [tcol, tcol + 256) as two N = 128 halves.  The B operand of a
    // half = rows [128 h, 128 h + 128) of the transposed K-block = a contiguous 16 KB half of its 32 KB image.
    // extra_b: more stages follow in the group (dirs rows of the view layer): they carry the "half b complete" commit.
    auto gemm2 = [&](uint32_t wt_off, int nkb, int tcol, bool extra_b) {
      for (int h = 0; h < 2; ++h)
        for (int kb = 0; kb < nkb; ++kb) {
          const bool last = kb == nkb - 1;
          int w0 = -1, w1 = -1;
          if (h == 0) {                       // operand block kb published; before the first MMA also: T[0:128) is free,
            w0 = kb;                          // i.e. the previous epilogue has read its chunks 0 and 1
            if (kb == 0 && nkb > 1) w1 = 1;
          }
          stage(lay.stream_bytes + wt_off + (uint32_t)kb * 32768u + (uint32_t)h * 16384u, 16384u, kb, 128, tcol + 128 * h,
                kb > 0, w0, w1, last ? (h == 0 ? 1 : (extra_b ? 0 : 2)) : 0, (h == 1 && !extra_b) ? kb : -1);
          phases[np - 1].reads_mask |= 1 << kb;
        }
    };
    phase(BK_PRE, 2, lay.L[V].s_out, lay.L[V].g_out, -1, 0, 0);
    gemm2(lay.L[V].wt_off, 2, 256, has_pose);      // views: T <- G_v Wv[:, :256]^T
    if (has_pose)                                   // (+ X[0:32] <- G_v Wv[:, 256:288]^T for objects)
      for (int kb = 0; kb < 2; ++kb)
        stage(lay.stream_bytes + lay.wt_dirs_off + (uint32_t)kb * 8192u, 8192u, kb, 32, 0, kb > 0, -1, -1, kb == 1 ? 2 : 0, kb);   // (the last readers of blocks 0, 1)
    phase(BK_FEAT, 4, -1, lay.L[F].g_out, -1, 1, 1);
    gemm2(lay.L[F].wt_off, 4, 0, false);
    phase(BK_OUT, 4, -1, lay.L[O].g_out, -1, 1, 1);
    gemm2(lay.L[O].wt_off, 4, 256, false);
    phase(BK_X0, 4, -1, lay.L[O - 1].g_out, O - 1, 1, 1);
    for (int b = lay.n_blocks - 1; b >= 0; --b) {
      const int l0 = 1 + 2 * b, l1 = 2 + 2 * b;
      gemm2(lay.L[l1].wt_off, 4, 256, false);
      phase(BK_FC1, 4, -1, lay.L[l0].g_out, l1 - 1, 1, 1);
      gemm2(lay.L[l0].wt_off, 4, 256, false);
      phase(BK_FC0, 4, -1, lay.L[l0 - 1].g_out, l0 - 1, 1, 1);
    }
    if (has_pose) {   // d enc_xyz = G_in W_in: N = 64, one "half"
      for (int kb = 0; kb < 4; ++kb) {
        stage(lay.stream_bytes + lay.L[0].wt_off + (uint32_t)kb * 8192u, 8192u, kb, 64, 256, kb > 0, kb, -1, kb == 3 ? 1 : 0, kb);
        phases[np - 1].reads_mask |= 1 << kb;
      }
      phase(BK_IN, 0, -1, -1, -1, 1, 0);
    }
    counts[0] = ns; counts[1] = np;
  }
  if (warp == TC_EPI_WARPS && lane == 0) {
    for (int i = 0; i < BWD2_NS; ++i) { mbar_init(bar(BAR_W_FULL(i)), 1); mbar_init(bar(BAR_W_EMPTY(i)), 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(BAR_A_READY(i)), TC_EPI_WARPS);
      mbar_init(bar(BAR_STASH_DONE_KB(i)), 1);
      mbar_init(bar(BAR2_B_DONE(i)), 1);
    }
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR2_ACC_B), 1);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(base + sl.tmem_ptr, TC_TMEM_COLS);
  for (int i = tid; i < lay.small_floats; i += TC_THREADS) s_small[i] = reinterpret_cast<const float*>(packed)[i];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int n_stages = counts[0], n_phases = counts[1];

  if (warp == TC_EPI_WARPS) {
    // ======================================================================== weight producer + gradient-stash writer
    // per group: request its weight stages (each waits for its ring slot, i.e. trails the MMAs of the previous group), then
    // copy the operand blocks of the phase that feeds it to the gradient stash as they are published
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, s_par = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint8_t* gs_tile = gstash + (size_t)tile * (size_t)lay.gstash_blocks * TC_BLOCK_BYTES;
        auto stash_phase = [&](int g) {
          const int nch = phases[g].nch, g_blk = phases[g].g_blk;
          if (nch == 0 || g_blk < 0) return;
          for (int kb = 0; kb < nch; ++kb) {
            mbar_wait(bar(BAR_A_READY(kb)), (s_par >> kb) & 1u, dbg, 6);
            s_par ^= 1u << kb;
            bulk_s2g(gs_tile + (size_t)(g_blk + kb) * TC_BLOCK_BYTES, sA + (uint32_t)kb * TC_KB_BYTES, TC_KB_BYTES);
            bulk_commit_group();
            if (kb >= 1) {                     // groups complete in order: block kb - 1 has been read
              bulk_wait_group_read1();
              mbar_arrive(bar(BAR_STASH_DONE_KB(kb - 1)));
            }
          }
          bulk_wait_group_read0();
          mbar_arrive(bar(BAR_STASH_DONE_KB(nch - 1)));
        };
        int i = 0;
        for (int g = 0; g < n_phases; ++g) {
          // the stages of group g: up to (and including) the one that signals "group complete" (sig_acc == 2, or == 1 for a
          // single-half group followed by a phase without half b)
          const bool has_group = i < n_stages;
          if (has_group) {
            const int want = phases[g + 1 < n_phases ? g + 1 : g].acc_b ? 2 : 1;
            for (;;) {
              const Bwd2Stage& s = stages[i++];
              mbar_wait(bar(BAR_W_EMPTY(stage)), phase ^ 1u, dbg, 1);
              mbar_arrive_expect_tx(bar(BAR_W_FULL(stage)), s.bytes);
              bulk_g2s(sW + stage * BWD2_STAGE_BYTES, packed + lay.small_bytes + s.w_off, s.bytes, bar(BAR_W_FULL(stage)));
              if (++stage == BWD2_NS) { stage = 0; phase ^= 1u; }
              if (s.sig_acc == want) break;
            }
          }
          stash_phase(g);
        }
      }
      bulk_wait_group0();
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    uint32_t stage = 0, phase = 0, a_par = 0;
    const uint64_t desc_a0 = umma_desc_sw128(sA), desc_w0 = umma_desc_sw128(sW);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int i = 0; i < n_stages; ++i) {
        const Bwd2Stage s = stages[i];
        if (s.wait0 >= 0) {
          mbar_wait(bar(BAR_A_READY(s.wait0)), (a_par >> s.wait0) & 1u, dbg, 2);
          a_par ^= 1u << s.wait0;
        }
        if (s.wait1 >= 0) {      // (waited for without consuming: its own stage consumes it)
          mbar_wait(bar(BAR_A_READY(s.wait1)), (a_par >> s.wait1) & 1u, dbg, 2);
        }
        mbar_wait(bar(BAR_W_FULL(stage)), phase, dbg, 3);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_16(TC_M, s.N, F16 ? 0 : 1);
        const uint64_t a0 = desc_a0 + (uint64_t)(s.a_kb * (TC_KB_BYTES >> 4));
        const uint64_t b0 = desc_w0 + (uint64_t)(stage * (BWD2_STAGE_BYTES >> 4));
        if (elect_one_sync()) {
          tc_mma_kblock<4>(tmem_base + (uint32_t)s.tcol, a0, b0, idesc, s.accum ? 1u : 0u);
          tc_commit(bar(BAR_W_EMPTY(stage)));
          if (s.sig_bdone >= 0) tc_commit(bar(BAR2_B_DONE(s.sig_bdone)));
          if (s.sig_acc == 1) tc_commit(bar(BAR_ACC_FULL));
          if (s.sig_acc == 2) tc_commit(bar(BAR2_ACC_B));
        }
        __syncwarp();
        if (++stage == BWD2_NS) { stage = 0; phase ^= 1u; }
      }
      // the tile's last phase publishes operand blocks that no MMA reads (static nets: G of lin_in, for dW only): consume
      // their a_ready phases to keep the parities in step across tiles
      const Bwd2Phase& PL = phases[n_phases - 1];
      if (PL.reads_mask == 0)
        for (int kb = 0; kb < PL.nch; ++kb) {
          mbar_wait(bar(BAR_A_READY(kb)), (a_par >> kb) & 1u, dbg, 5);
          a_par ^= 1u << kb;
        }
    }
  } else {
    // ======================================================================== epilogue warps
    const int q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    uint32_t acc_par = 0, accb_par = 0;
    uint32_t st_par = 0, st_pend = 0;      // stash_done[kb]: parity of the next wait / a store of block kb is outstanding
    uint32_t bd_par = 0, bd_pend = 0;      // b_done[kb]: the same for "the last MMA that reads block kb"
    float pacc[POSE ? 27 : 1];
#pragma unroll
    for (int i = 0; i < (POSE ? 27 : 1); ++i) pacc[i] = 0.f;
    const float gscale = F16 ? grad_scale_of(absmax) : 1.f;
    // operand block kb is about to be overwritten: its last reader (MMA) and its copy to the gradient stash must be done
    auto claim_block = [&](int kb) {
      const uint32_t bit = 1u << kb;
      if (bd_pend & bit) {
        mbar_wait(bar(BAR2_B_DONE(kb)), (bd_par & bit) ? 1u : 0u, dbg, 8);
        bd_par ^= bit; bd_pend &= ~bit;
      }
      if (st_pend & bit) {
        mbar_wait(bar(BAR_STASH_DONE_KB(kb)), (st_par & bit) ? 1u : 0u, dbg, 7);
        st_par ^= bit; st_pend &= ~bit;
      }
    };
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gi = tile * TC_M + row;
      const bool valid = gi < M;
      float da = 0.f, dc0 = 0.f, dc1 = 0.f, dc2 = 0.f;
      float pw[3] = {0.f, 0.f, 0.f}, po[3] = {0.f, 0.f, 0.f}, dw[3] = {0.f, 0.f, 0.f}, dob[3] = {0.f, 0.f, 0.f};
      if (valid) {
        const int64_t r = gi / S;
        const int64_t o = r * ray_stride + (gi - r * S);
        da = d_raw_alpha[o] * gscale;
        dc0 = d_raw_rgb[o * 3 + 0] * gscale; dc1 = d_raw_rgb[o * 3 + 1] * gscale; dc2 = d_raw_rgb[o * 3 + 2] * gscale;
        if (has_pose) {
          star_load_pt(pts, gi, r, pw[0], pw[1], pw[2]);
          dw[0] = viewdirs[r * 3 + 0]; dw[1] = viewdirs[r * 3 + 1]; dw[2] = viewdirs[r * 3 + 2];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            po[i] = pose12[i * 3 + 0] * pw[0] + pose12[i * 3 + 1] * pw[1] + pose12[i * 3 + 2] * pw[2] + pose12[9 + i];
            dob[i] = pose12[i * 3 + 0] * dw[0] + pose12[i * 3 + 1] * dw[1] + pose12[i * 3 + 2] * dw[2];
          }
        }
      }
      const uint8_t* st_tile = stash + (size_t)tile * (size_t)lay.stash_blocks * TC_BLOCK_BYTES;
      const uint32_t x_sw = (uint32_t)row & 7u;

      for (int pi = 0; pi < n_phases; ++pi) {
        const Bwd2Phase P = phases[pi];
        // masks of the whole phase first (global-memory latency hides behind the MMAs)
        uint4 mq[2][2];
        uint2 mbits = make_uint2(~0u, ~0u);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          mq[kb][0] = make_uint4(~0u, ~0u, ~0u, ~0u);
          mq[kb][1] = mq[kb][0];
          if (P.mask_blk >= 0 && kb < P.nch) {
            const uint8_t* mb = st_tile + (size_t)(P.mask_blk + kb) * TC_BLOCK_BYTES;
            mq[kb][0] = __ldg(reinterpret_cast<const uint4*>(mb + (uint32_t)row * 128u + ((((uint32_t)(cg * 2)) ^ x_sw) << 4)));
            mq[kb][1] = __ldg(reinterpret_cast<const uint4*>(mb + (uint32_t)row * 128u + ((((uint32_t)(cg * 2 + 1)) ^ x_sw) << 4)));
          }
        }
        if (P.mask_layer >= 0)
          mbits = __ldg(reinterpret_cast<const uint2*>(st_tile + (size_t)lay.mask_blk0 * TC_BLOCK_BYTES +
                                                       (size_t)P.mask_layer * TC_MASK_BYTES) + (cg * TC_M + row));
        if (P.acc_a) {
          mbar_wait(bar(BAR_ACC_FULL), acc_par, dbg, 4);
          acc_par ^= 1u;
          tc_fence_after();
        }
        auto wait_half_b = [&]() {
          mbar_wait(bar(BAR2_ACC_B), accb_par, dbg, 4);
          accb_par ^= 1u;
          tc_fence_after();
        };
        const uint32_t tX = tmem_base + lane_addr + (uint32_t)(cg * TC_CPT);
        const uint32_t tT = tX + 256u;
        bool half_b_seen = false;
        if (P.kind == BK_FEAT && has_pose) {
          // d enc_dir sits in X[0:32], written by the stages that also complete half b; the next group (feature_linear^T)
          // writes X from its first MMA on, i.e. as soon as chunks 0 and 1 of THIS phase are published: read it first
          wait_half_b();
          half_b_seen = true;
          if (cg < 2) {
            uint32_t r[16];
            tmem_ld16(tX, r);
            tmem_wait_ld();
            float h[3] = {0.f, 0.f, 0.f};
            if (cg == 0) fold_jacobian<0, 27>(r, dob, sc_dir, h);
            else fold_jacobian<16, 27>(r, dob, sc_dir, h);
            if (valid) {
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) pacc[15 + i * 3 + j] += h[i] * dw[j];
              pacc[24] += dob[1] * h[2] - dob[2] * h[1];
              pacc[25] += dob[2] * h[0] - dob[0] * h[2];
              pacc[26] += dob[0] * h[1] - dob[1] * h[0];
            }
          }
          tc_fence_before();
        }
        if (P.kind == BK_IN) {       // d enc_xyz in T[0:64]: 16 columns per thread
          uint32_t r[16];
          tmem_ld16(tT, r);
          tmem_wait_ld();
          float g[3] = {0.f, 0.f, 0.f};
          if (cg == 0) fold_jacobian<0, 63>(r, po, sc_xyz, g);
          else if (cg == 1) fold_jacobian<16, 63>(r, po, sc_xyz, g);
          else if (cg == 2) fold_jacobian<32, 63>(r, po, sc_xyz, g);
          else fold_jacobian<48, 63>(r, po, sc_xyz, g);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              pacc[i] += g[i];
#pragma unroll
              for (int j = 0; j < 3; ++j) pacc[3 + i * 3 + j] += g[i] * pw[j];
            }
            pacc[12] += po[1] * g[2] - po[2] * g[1];
            pacc[13] += po[2] * g[0] - po[0] * g[2];
            pacc[14] += po[0] * g[1] - po[1] * g[0];
          }
          tc_fence_before();
          continue;
        }
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb >= P.nch) break;
          if (kb == 2 && P.acc_b && !half_b_seen) wait_half_b();
          const int col0 = kb * 64 + cg * TC_CPT;
          const uint4 m0 = mq[kb & 1][0], m1 = mq[kb & 1][1];
          const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          const uint32_t mb16 = ((kb < 2 ? mbits.x : mbits.y) >> ((kb & 1) * 16)) & 0xffffu;
          float v[16];
          if (P.kind == BK_PRE) {
            const float* rw = s_small + lay.off_rgb_w + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = dc0 * rw[j] + dc1 * rw[STAR_WV + j] + dc2 * rw[2 * STAR_WV + j];
          } else {
            uint32_t r[16];
            tmem_ld16(((P.kind == BK_OUT) ? tX : tT) + 64u * (uint32_t)kb, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (P.kind == BK_OUT) {
            const float* aw = s_small + lay.off_alpha_w + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(da, aw[j], v[j]);
          }
          if (P.mask_blk >= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t bits = (j & 1) ? (mw[j >> 1] >> 16) : (mw[j >> 1] & 0xffffu);
              if (bits == 0u) v[j] = 0.f;
            }
          }
          if (P.mask_layer >= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (!((mb16 >> j) & 1u)) v[j] = 0.f;
          }
          if (P.kind == BK_FC0) {      // dx += masked product
            uint32_t rx[16];
            tmem_ld16(tX + 64u * (uint32_t)kb, rx);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(rx[j]);
          }
          if (P.kind == BK_FC0 || P.kind == BK_X0) {
            uint32_t rx[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) rx[j] = __float_as_uint(v[j]);
            tmem_st16(tX + 64u * (uint32_t)kb, rx);
            tmem_wait_st();
          }
          if (!valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          claim_block(kb);
          store_row16<F16, false, true>(sA + (uint32_t)kb * TC_KB_BYTES, row, cg * 2, v);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_A_READY(kb)));
          if (P.g_blk >= 0) st_pend |= 1u << kb;
        }
        if (P.acc_b && P.nch <= 2 && !half_b_seen) wait_half_b();           // (keep the half-b barrier in step)
        bd_pend |= (uint32_t)P.reads_mask;                   // the following group commits b_done for the blocks it reads
      }
    }
    if (has_pose) {
      const float inv = 1.f / gscale;
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const float s = warp_sum(pacc[i]);
        if (lane == 0) atomicAdd(&pose_acc[i], s * inv);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ============================================================================================ dW / db
// One CTA = (layer l, 128-row half of its outputs, sample-tile range).  Per tile: 2 gradient blocks (A operand,
// MN-major: M = output features) and up to 4 activation blocks (B operand, MN-major: N = input features), K = 128
// samples = 8 MMAs of K = 16; a constant "ones" block gives db as 16 extra accumulator columns.
#define DW_THREADS 192          // warp 0: producer, warp 1: MMA, warps 2..5: epilogue (TMEM lane quadrants 2,3,0,1)
#define DW_STAGE_BLOCKS 6
#define DW_NSTAGE 2

struct DwItem {
  int g_blk, a_blk, n_a;     // first gradient-stash block (2 blocks), first stash block, number of stash blocks
  int64_t w_off, b_off;      // flat-gradient offsets of W[half*128][0] (+k0) and b[half*128] (-1: no bias here)
  int K, k0, k_valid;        // row length of W, first column written, valid columns of this item
  int ipe_perm;              // mip field: operand column c holds master column ipe_master_col(c) (-1: padding)
};
#define DW_MAX_ITEMS 40
struct DwPlan {
  int n_items, splits;
  DwItem it[DW_MAX_ITEMS];
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // between 64-element groups along M / N
  d |= (uint64_t)(1024 >> 4) << 32;                    // between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// F16: operand format (both the gradient and the activation blocks; see the note on formats above).  The fp16 tier's
// gradients are scaled by 2^k: the epilogue multiplies the accumulators by 2^-k.
template <bool F16>
__global__ void __launch_bounds__(DW_THREADS, 1)
dw_tc_kernel(const DwPlan plan, int stash_blocks, int gstash_blocks, const uint8_t* __restrict__ stash,
             const uint8_t* __restrict__ gstash, int64_t ntiles, float* __restrict__ grad_flat,
             const float* __restrict__ absmax, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  tc_mark_begin(dbg);
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  const uint32_t sStage = base, sOnes = base + DW_NSTAGE * DW_STAGE_BLOCKS * TC_BLOCK_BYTES;
  const uint32_t sBars = sOnes + TC_BLOCK_BYTES;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(gbase + (sBars - base) + 64);
  auto bar = [&](int i) -> uint32_t { return sBars + 8u * (uint32_t)i; };   // 0,1 full; 2,3 empty; 4 done; 5,6 converted
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item_id = blockIdx.x / plan.splits, split = blockIdx.x % plan.splits;
  const DwItem it = plan.it[item_id];
  const int64_t per = (ntiles + plan.splits - 1) / plan.splits;
  const int64_t t0 = (int64_t)split * per, t1 = (t0 + per < ntiles) ? t0 + per : ntiles;

  if (tid == 0) {
    for (int i = 0; i < 2 * DW_NSTAGE; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(4), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(sBars + 64, 512);
  // ones block: logical column 0 of every row = 1.0 (16-byte chunk 0 ^ (row & 7), element 0), zero elsewhere
  for (int i = tid; i < TC_BLOCK_BYTES / 16; i += DW_THREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c == (r & 7)) v.x = F16 ? 0x3C00u : 0x3F80u;     // 1.0
    reinterpret_cast<uint4*>(gbase + (sOnes - base))[i] = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t stage_bytes = (uint32_t)(2 + it.n_a) * TC_BLOCK_BYTES;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int64_t t = t0; t < t1; ++t) {
        mbar_wait(bar(2 + st), ph ^ 1u, dbg, 1);
        mbar_arrive_expect_tx(bar(st), stage_bytes);
        const uint32_t dst = sStage + st * DW_STAGE_BLOCKS * TC_BLOCK_BYTES;
        bulk_g2s(dst, gstash + ((size_t)t * gstash_blocks + it.g_blk) * TC_BLOCK_BYTES, 2 * TC_BLOCK_BYTES, bar(st));
        bulk_g2s(dst + 2 * TC_BLOCK_BYTES, stash + ((size_t)t * stash_blocks + it.a_blk) * TC_BLOCK_BYTES,
                 (uint32_t)it.n_a * TC_BLOCK_BYTES, bar(st));
        if (++st == DW_NSTAGE) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // warp-uniform control flow, one elected lane issues the tcgen05 instructions (see elect_one_sync)
    {
      const int N = 64 * it.n_a;
      const uint32_t idesc = umma_idesc_16(128, N, F16 ? 0 : 1) | (1u << 15) | (1u << 16);      // both operands MN-major
      const uint32_t idesc1 = umma_idesc_16(128, 16, F16 ? 0 : 1) | (1u << 15) | (1u << 16);
      const bool with_bias = it.b_off >= 0;
      const uint64_t d1 = umma_desc_mn_sw128(sOnes, TC_BLOCK_BYTES);
      uint32_t st = 0, ph = 0;
      uint32_t acc = 0u;
      for (int64_t t = t0; t < t1; ++t) {
        mbar_wait(bar(st), ph, dbg, 2);
        tc_fence_after();
        const uint32_t g_addr = sStage + st * DW_STAGE_BLOCKS * TC_BLOCK_BYTES, a_addr = g_addr + 2 * TC_BLOCK_BYTES;
        const uint64_t dg = umma_desc_mn_sw128(g_addr, TC_BLOCK_BYTES), da = umma_desc_mn_sw128(a_addr, TC_BLOCK_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {     // +2048 bytes per K-step of 16 samples = +128 in the 16-byte address field
            const uint32_t a_ = (ks == 0) ? acc : 1u;
            tc_mma_bf16(tmem_base, dg + (uint64_t)(128 * ks), da + (uint64_t)(128 * ks), idesc, a_);
            if (with_bias) tc_mma_bf16(tmem_base + 256u, dg + (uint64_t)(128 * ks), d1 + (uint64_t)(128 * ks), idesc1, a_);
          }
          tc_commit(bar(2 + st));
          if (t == t1 - 1) tc_commit(bar(4));
        }
        __syncwarp();
        acc = 1u;
        if (++st == DW_NSTAGE) { st = 0; ph ^= 1u; }
      }
      if (t1 <= t0) {
        if (elect_one_sync()) tc_commit(bar(4));
        __syncwarp();
      }
    }
  } else {
    // epilogue: TMEM -> registers -> atomics into the flat gradient
    const int q = warp & 3;
    const int n = q * 32 + lane;                      // output feature within the half
    const float unscale = F16 ? 1.f / grad_scale_of(absmax) : 1.f;
    if (t1 > t0) {
      mbar_wait(bar(4), 0, dbg, 3);
      tc_fence_after();
      const uint32_t tl = tmem_base + (((uint32_t)(q * 32)) << 16);
      float* wrow = grad_flat + it.w_off + (int64_t)n * it.K + it.k0;
      for (int c0 = 0; c0 < 64 * it.n_a; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tl + (uint32_t)c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (it.ipe_perm) {
            const int mc = ipe_master_col(c0 + j);
            if (mc >= 0) atomicAdd(wrow + mc, __uint_as_float(r[j]) * unscale);
          } else if (c0 + j < it.k_valid) {
            atomicAdd(wrow + c0 + j, __uint_as_float(r[j]) * unscale);
          }
      }
      if (it.b_off >= 0) {
        uint32_t r[16];
        tmem_ld16(tl + 256u, r);
        tmem_wait_ld();
        atomicAdd(grad_flat + it.b_off + n, __uint_as_float(r[0]) * unscale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_mark_end(dbg);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ============================================================================================ head gradients
// d alpha_w[k] = sum_m d_alpha[m] h[m][k]  (h = stashed input of feature_linear), d rgb_w[c][n] = sum_m d_rgb[m][c] relu(h2)[m][n].
// 256 threads = 8 row groups x 32 sixteen-byte chunks: thread (rg, ch) owns 8 consecutive columns of h (and of
// relu(h2) when ch < 16) for rows rg*16..rg*16+15 of every tile; partial sums stay in registers across tiles.
__device__ __forceinline__ void unpack8(const uint4& v, bool fp16, float (&x)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      x[2 * i] = f.x; x[2 * i + 1] = f.y;
    } else {
      x[2 * i] = __uint_as_float(w[i] << 16);
      x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}

struct HeadGradCfg {
  int stash_blocks;          // blocks per tile of the activation stash
  int blk_h, blk_h2;         // first block of the 256-wide input of the density head / of the 128-wide input of the rgb head
  int64_t m_aw, m_ab, m_rw, m_rb;   // flat-gradient offsets: density head weight [256], bias; rgb head weight [3][128], bias [3]
};

// Rows of a tile are taken in two halves of 64 (8 rows per thread and half: 8 + 4 sixteen-byte loads in flight), and the
// 128-wide relu(h2) is split over BOTH half-warps -- lane ch owns chunk ch & 15 for the rows of parity ch >> 4 -- so that no
// lane idles through the rgb part: 2 CTAs per SM (<= 128 registers) instead of one at 199 registers with half of every warp
// masked off in 2/3 of its instructions.
#ifndef HEAD_GRAD_OLD
__global__ void __launch_bounds__(256, 2)
head_grad_tc_kernel(const HeadGradCfg hc, const uint8_t* __restrict__ stash, const float* __restrict__ d_raw_alpha,
                    const float* __restrict__ d_raw_rgb, int64_t ray_stride, int S, int64_t M, int fp16,
                    float* __restrict__ grad_flat) {
  const int tid = threadIdx.x, rg = tid >> 5, ch = tid & 31;
  const int par = ch >> 4;
  const int64_t ntiles = (M + 127) / 128;
  float aw[8], rw[3][8], ab = 0.f, rb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) { aw[j] = 0.f; rw[0][j] = 0.f; rw[1][j] = 0.f; rw[2][j] = 0.f; }
  __shared__ float s_d[128][4];
  __shared__ float s_red[8][32][33];
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    __syncthreads();
    if (tid < 128) {
      const int64_t gi = t * 128 + tid;
      float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (gi < M) {
        const int64_t r = gi / S, o = r * ray_stride + (gi - r * S);
        a = d_raw_alpha[o]; c0 = d_raw_rgb[o * 3]; c1 = d_raw_rgb[o * 3 + 1]; c2 = d_raw_rgb[o * 3 + 2];
      }
      s_d[tid][0] = a; s_d[tid][1] = c0; s_d[tid][2] = c1; s_d[tid][3] = c2;
    }
    __syncthreads();
    const uint8_t* tile = stash + (size_t)t * hc.stash_blocks * TC_BLOCK_BYTES;
    const uint8_t* hb = tile + (size_t)(hc.blk_h + (ch >> 3)) * TC_BLOCK_BYTES;
    const uint8_t* h2b = tile + (size_t)(hc.blk_h2 + ((ch & 15) >> 3)) * TC_BLOCK_BYTES;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m0 = half * 64 + rg * 8;
      uint4 qh[8], q2[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + i;
        qh[i] = __ldg(reinterpret_cast<const uint4*>(hb + (uint32_t)m * 128u + ((((uint32_t)ch & 7u) ^ ((uint32_t)m & 7u)) << 4)));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + 2 * i + par;
        q2[i] = __ldg(reinterpret_cast<const uint4*>(h2b + (uint32_t)m * 128u + ((((uint32_t)ch & 7u) ^ ((uint32_t)m & 7u)) << 4)));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = s_d[m0 + i][0];
        float x[8];
        unpack8(qh[i], fp16 != 0, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) aw[j] = fmaf(d, x[j], aw[j]);
        if (ch == 31) {
          const float4 dd = *reinterpret_cast<const float4*>(&s_d[m0 + i][0]);
          ab += dd.x; rb[0] += dd.y; rb[1] += dd.z; rb[2] += dd.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 d = *reinterpret_cast<const float4*>(&s_d[m0 + 2 * i + par][0]);
        float x[8];
        unpack8(q2[i], fp16 != 0, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          rw[0][j] = fmaf(d.y, x[j], rw[0][j]);
          rw[1][j] = fmaf(d.z, x[j], rw[1][j]);
          rw[2][j] = fmaf(d.w, x[j], rw[2][j]);
        }
      }
    }
  }
  // reduce the 8 row groups (and, for the rgb head, the two row parities) through shared memory, then one atomic per
  // output element and CTA
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) s_red[rg][ch][j] = aw[j];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_red[rg][ch][8 + c * 8 + j] = rw[c][j];
  s_red[rg][ch][32] = (ch == 31) ? ab : 0.f;
  __syncthreads();
  for (int e = tid; e < 32 * 32; e += 256) {
    const int c2 = e >> 5, k = e & 31;     // chunk, slot
    if (k < 8) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += s_red[g][c2][k];
      atomicAdd(&grad_flat[hc.m_aw + c2 * 8 + k], v);
    } else if (c2 < 16) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += s_red[g][c2][k] + s_red[g][c2 + 16][k];
      atomicAdd(&grad_flat[hc.m_rw + ((k - 8) >> 3) * STAR_WV + c2 * 8 + ((k - 8) & 7)], v);
    }
  }
  if (tid == 0) {
    float v = 0.f;
    for (int g = 0; g < 8; ++g) v += s_red[g][31][32];
    atomicAdd(&grad_flat[hc.m_ab], v);
  }
  // rgb bias: sum of d_rgb
  __syncthreads();
  if (ch == 31) { s_red[rg][0][0] = rb[0]; s_red[rg][0][1] = rb[1]; s_red[rg][0][2] = rb[2]; }
  __syncthreads();
  if (tid < 3) {
    float v = 0.f;
    for (int g = 0; g < 8; ++g) v += s_red[g][0][tid];
    atomicAdd(&grad_flat[hc.m_rb + tid], v);
  }
}
#define HEAD_GRAD_CTAS_PER_SM 2
#else
__global__ void __launch_bounds__(256)
head_grad_tc_kernel(const HeadGradCfg hc, const uint8_t* __restrict__ stash, const float* __restrict__ d_raw_alpha,
                    const float* __restrict__ d_raw_rgb, int64_t ray_stride, int S, int64_t M, int fp16,
                    float* __restrict__ grad_flat) {
  const int tid = threadIdx.x, rg = tid >> 5, ch = tid & 31;
  const int64_t ntiles = (M + 127) / 128;
  float aw[8], rw[3][8], ab = 0.f, rb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) { aw[j] = 0.f; rw[0][j] = 0.f; rw[1][j] = 0.f; rw[2][j] = 0.f; }
  __shared__ float s_d[128][4];
  __shared__ float s_red[8][32][33];
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    __syncthreads();
    if (tid < 128) {
      const int64_t gi = t * 128 + tid;
      float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (gi < M) {
        const int64_t r = gi / S, o = r * ray_stride + (gi - r * S);
        a = d_raw_alpha[o]; c0 = d_raw_rgb[o * 3]; c1 = d_raw_rgb[o * 3 + 1]; c2 = d_raw_rgb[o * 3 + 2];
      }
      s_d[tid][0] = a; s_d[tid][1] = c0; s_d[tid][2] = c1; s_d[tid][3] = c2;
    }
    __syncthreads();
    const uint8_t* tile = stash + (size_t)t * hc.stash_blocks * TC_BLOCK_BYTES;
    const uint8_t* hb = tile + (size_t)(hc.blk_h + (ch >> 3)) * TC_BLOCK_BYTES;
    const uint8_t* h2b = tile + (size_t)(hc.blk_h2 + ((ch & 15) >> 3)) * TC_BLOCK_BYTES;
    // all 16 (+16) 16-byte loads of the tile are issued before the first use: the kernel is a pure HBM reader
    uint4 qh[16], q2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int m = rg * 16 + i;
      const uint32_t off = (uint32_t)m * 128u + ((((uint32_t)ch & 7u) ^ ((uint32_t)m & 7u)) << 4);
      qh[i] = __ldg(reinterpret_cast<const uint4*>(hb + off));
      if (ch < 16) q2[i] = __ldg(reinterpret_cast<const uint4*>(h2b + off));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int m = rg * 16 + i;
      const float4 d = *reinterpret_cast<const float4*>(&s_d[m][0]);
      float x[8];
      unpack8(qh[i], fp16 != 0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) aw[j] = fmaf(d.x, x[j], aw[j]);
      if (ch < 16) {
        unpack8(q2[i], fp16 != 0, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          rw[0][j] = fmaf(d.y, x[j], rw[0][j]);
          rw[1][j] = fmaf(d.z, x[j], rw[1][j]);
          rw[2][j] = fmaf(d.w, x[j], rw[2][j]);
        }
      }
      if (ch == 31) { ab += d.x; rb[0] += d.y; rb[1] += d.z; rb[2] += d.w; }
    }
  }
  // reduce the 8 row groups through shared memory, then one atomic per output element and CTA
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) s_red[rg][ch][j] = aw[j];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_red[rg][ch][8 + c * 8 + j] = rw[c][j];
  s_red[rg][ch][32] = (ch == 31) ? ab : 0.f;
  __syncthreads();
  for (int e = tid; e < 32 * 32; e += 256) {
    const int c2 = e >> 5, k = e & 31;     // chunk, slot
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += s_red[g][c2][k];
    if (k < 8) atomicAdd(&grad_flat[hc.m_aw + c2 * 8 + k], v);
    else if (c2 < 16) atomicAdd(&grad_flat[hc.m_rw + ((k - 8) >> 3) * STAR_WV + c2 * 8 + ((k - 8) & 7)], v);
  }
  if (tid == 0) {
    float v = 0.f;
    for (int g = 0; g < 8; ++g) v += s_red[g][31][32];
    atomicAdd(&grad_flat[hc.m_ab], v);
  }
  // rgb bias: sum of d_rgb
  __syncthreads();
  if (ch == 31) { s_red[rg][0][0] = rb[0]; s_red[rg][0][1] = rb[1]; s_red[rg][0][2] = rb[2]; }
  __syncthreads();
  if (tid < 3) {
    float v = 0.f;
    for (int g = 0; g < 8; ++g) v += s_red[g][0][tid];
    atomicAdd(&grad_flat[hc.m_rb + tid], v);
  }
}

#define HEAD_GRAD_CTAS_PER_SM 4
#endif

// ============================================================================================ transposed weight stream
// dX GEMM of layer l: out[m][j] = sum_n G[m][n] W[n][j]: B operand rows j (input features, padded to 64 / 256),
// K-blocks over the output features n, swizzled like every other block.
__global__ void pack_tc_tstream_kernel(TcLayout tl, MlpLayout ml, const float* __restrict__ master,
                                       uint16_t* __restrict__ tstream, int fp16) {
  const uint32_t n_elems = tl.tstream_bytes / 2;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += gridDim.x * blockDim.x) {
    const uint32_t byte = e * 2;
    float v = 0.f;
    int rows, l, kb, j, nn;
    uint32_t off;
    if (byte >= tl.wt_dirs_off) {                 // view layer, encoded-dirs rows: [64 rows j][64 n] x 2
      l = tl.n_layers - 1; rows = 64; off = byte - tl.wt_dirs_off;
    } else {
      l = 0;
      while (l + 1 < tl.n_layers && byte >= tl.L[l + 1].wt_off) ++l;
      rows = (tl.L[l].nkb < 4 ? tl.L[l].nkb : 4) * 64; off = byte - tl.L[l].wt_off;
    }
    const uint32_t kb_bytes = (uint32_t)rows * 128u;
    kb = (int)(off / kb_bytes);
    const uint32_t rem = off % kb_bytes;
    j = (int)(rem >> 7);
    const uint32_t inrow = rem & 127u;
    const int chunk = (int)((inrow >> 4) ^ ((uint32_t)j & 7u));
    nn = kb * 64 + chunk * 8 + (int)((inrow & 15u) >> 1);      // output feature
    const int K = ml.L[l].K;
    const int jj = (byte >= tl.wt_dirs_off) ? STAR_W + j : j;    // input feature
    if (nn < tl.L[l].N && jj < K && (byte >= tl.wt_dirs_off ? j < K - STAR_W : true))
      v = master[ml.L[l].m_w + (int64_t)nn * K + jj];
    tstream[e] = fp16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// ============================================================================================ host side
int star_tc_pack_tstream(const TcLayout& tl, const MlpLayout& ml, const float* master, void* tstream, int fp16,
                         cudaStream_t st) {
  pack_tc_tstream_kernel<<<148 * 4, 256, 0, st>>>(tl, ml, master, (uint16_t*)tstream, fp16);
  return star_check_launch();
}

size_t star_tc_gstash_bytes(const TcLayout& tl, int64_t n_samples) {
  // (+ 256 bytes at the end: the |upstream gradient| maximum of the call, grad_absmax_kernel)
  return (size_t)(((n_samples + 127) / 128 + 1) / 2 * 2) * (size_t)tl.gstash_blocks * TC_BLOCK_BYTES + 256;
}

int star_tc_backward(const TcLayout& tl, const MlpLayout& ml, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                     const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, const float* d_raw_alpha,
                     const float* d_raw_rgb, int64_t ray_stride, const void* stash, void* gstash, float* grad_flat,
                     float* pose_acc, int fp16, int serial_dx, cudaStream_t st) {
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // ---- 0. fp16 tier: largest |upstream gradient| of the call -> the power-of-two gradient scale (read on the device)
  float* absmax = reinterpret_cast<float*>((uint8_t*)gstash + star_tc_gstash_bytes(tl, M) - 256);
  if (fp16) {
    cudaError_t e = cudaMemsetAsync(absmax, 0, 4, st);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    int64_t b = (M + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    grad_absmax_kernel<<<(int)b, 256, 0, st>>>(d_raw_alpha, d_raw_rgb, S, M, ray_stride, absmax);
    int rc = star_check_launch();
    if (rc) return rc;
  }
  // ---- 1. dX chain: the pipelined kernel (N = 128 halves), or the serial one (A/B flag)
  if (!serial_dx) {
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const Bwd2Smem sl = bwd2_smem_layout(tl.small_bytes);
    auto kern = pose12 != nullptr ? (fp16 ? mlp_bwd2_tc_kernel<true, true> : mlp_bwd2_tc_kernel<false, true>)
                                  : (fp16 ? mlp_bwd2_tc_kernel<true, false> : mlp_bwd2_tc_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    kern<<<grid, TC_THREADS, sl.total, st>>>(tl, (const uint8_t*)packed, pts, viewdirs, pose12, sc_xyz, sc_dir, S, M,
                                              d_raw_alpha, d_raw_rgb, ray_stride, (const uint8_t*)stash, (uint8_t*)gstash,
                                              pose_acc, absmax, star_watchdog_dev(STAR_WD_DX2));
    int rc = star_check_launch();
    if (rc) return rc;
  } else {
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const BwdSmem sl = bwd_smem_layout(tl.small_bytes);
    // TC_WSHARE_DX: clusters of 2 share the (transposed) weight stream as in the forward kernels (mlp_tc.cu)
    const bool wshare = TC_WSHARE_DX && ntiles >= 2 * (int64_t)sms;
    using KernDx = void (*)(const TcLayout, const uint8_t*, const StarPtsSrc, const float*, const float*, const float*,
                            const float*, int, int64_t, const float*, const float*, int64_t, const uint8_t*, uint8_t*, float*,
                            const float*, int*);
    KernDx kern = wshare ? (pose12 != nullptr ? (fp16 ? mlp_bwd_tc_kernel<true, true, TC_WSHARE_DX != 0> : mlp_bwd_tc_kernel<false, true, TC_WSHARE_DX != 0>)
                                              : (fp16 ? mlp_bwd_tc_kernel<true, false, TC_WSHARE_DX != 0> : mlp_bwd_tc_kernel<false, false, TC_WSHARE_DX != 0>))
                         : (pose12 != nullptr ? (fp16 ? mlp_bwd_tc_kernel<true, true> : mlp_bwd_tc_kernel<false, true>)
                                              : (fp16 ? mlp_bwd_tc_kernel<true, false> : mlp_bwd_tc_kernel<false, false>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    if (wshare) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(grid & ~1), 1, 1);
      cfg.blockDim = dim3(TC_THREADS, 1, 1);
      cfg.dynamicSmemBytes = sl.total;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tl, (const uint8_t*)packed, pts, viewdirs, pose12, sc_xyz, sc_dir, S, M,
                                                d_raw_alpha, d_raw_rgb, ray_stride, (const uint8_t*)stash, (uint8_t*)gstash,
                                                pose_acc, (const float*)absmax, star_watchdog_dev(STAR_WD_DX));
      if (le != cudaSuccess) { g_star_last_cuda_error = (int)le; return STAR_E_CUDA; }
    } else {
      kern<<<grid, TC_THREADS, sl.total, st>>>(tl, (const uint8_t*)packed, pts, viewdirs, pose12, sc_xyz, sc_dir, S, M,
                                                d_raw_alpha, d_raw_rgb, ray_stride, (const uint8_t*)stash, (uint8_t*)gstash,
                                                pose_acc, absmax, star_watchdog_dev(STAR_WD_DX));
    }
    int rc = star_check_launch();
    if (rc) return rc;
  }
  // ---- 2. dW / db
  {
    DwPlan plan;
    int n = 0;
    for (int l = 0; l < tl.n_layers; ++l) {
      const TcLayer& L = tl.L[l];
      const int halves = L.N / 128, K = ml.L[l].K;
      for (int h = 0; h < halves; ++h) {
        DwItem& it = plan.it[n++];
        it.g_blk = L.g_out + 2 * h;
        it.a_blk = L.s_in;
        it.n_a = L.nkb < 4 ? L.nkb : 4;
        it.K = K; it.k0 = 0;
        it.k_valid = K < 64 * it.n_a ? K : 64 * it.n_a;
        it.w_off = ml.L[l].m_w + (int64_t)h * 128 * K;
        it.b_off = ml.L[l].m_b + h * 128;
        it.ipe_perm = 0;
        if (L.kind == LK_VIEWS) {          // encoded-dirs columns of the view layer: separate item, no bias
          DwItem& d2 = plan.it[n++];
          d2 = it;
          d2.a_blk = L.s_in + 4; d2.n_a = 1; d2.k0 = STAR_W; d2.k_valid = K - STAR_W; d2.b_off = -1;
        }
      }
    }
    plan.n_items = n;
    int splits = sms / n;
    if (splits < 1) splits = 1;
    if ((int64_t)splits > ntiles) splits = (int)ntiles;
    plan.splits = splits;
    const size_t smem = (size_t)(DW_NSTAGE * DW_STAGE_BLOCKS + 1) * TC_BLOCK_BYTES + 256 + 1024;
    auto kern = fp16 ? dw_tc_kernel<true> : dw_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    kern<<<n * splits, DW_THREADS, smem, st>>>(plan, tl.stash_blocks, tl.gstash_blocks, (const uint8_t*)stash,
                                               (const uint8_t*)gstash, ntiles, grad_flat, absmax, star_watchdog_dev(STAR_WD_DW));
    int rc = star_check_launch();
    if (rc) return rc;
  }
  // ---- 3. heads
  {
    int blocks = (int)(ntiles < HEAD_GRAD_CTAS_PER_SM * sms ? ntiles : HEAD_GRAD_CTAS_PER_SM * sms);
    const HeadGradCfg hc{tl.stash_blocks, tl.L[tl.n_layers - 2].s_in, tl.L[tl.n_layers - 1].s_out, ml.m_alpha_w, ml.m_alpha_b,
                         ml.m_rgb_w, ml.m_rgb_b};
    head_grad_tc_kernel<<<blocks, 256, 0, st>>>(hc, (const uint8_t*)stash, d_raw_alpha, d_raw_rgb, ray_stride, S, M, fp16,
                                                grad_flat);
    return star_check_launch();
  }
}

// ================================================================================================================
// mip-NeRF field (models/mipnerf.py:53-100 -> nerfstudio NeRFField): tensor-core backward to the
// WEIGHTS.  Same machine as the dX chain above (table-driven producer / issuer, bulk-stored gradient stash, the same
// dw_tc_kernel and head-gradient kernel), with the mip layer program: no residual stream, ReLU after every layer (the
// masks are read off the stashed rectified outputs: bits != 0), the density head on the rectified base output.
// Gradients to the ray (pose of an object field: through the integrated positional encoding and its covariance) are NOT
// produced here -- a pass that needs them runs on the fp32 tier (mip_f32.cu), which the host selects.
#define MK_B_PRE 0     // G_h1 = (d_rgb . W_rgb) (.) mask(relu(h1))           (no MMA before it)
#define MK_B_T 1       // G = acc (.) mask
#define MK_B_BASE 2    // G_7 = (acc + d_sigma w_dens) (.) mask(relu(base_out))

template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
mip_bwd_tc_kernel(const MipTcLayout lay, const uint8_t* __restrict__ packed, int S, int64_t M,
                  const float* __restrict__ d_raw_sigma, const float* __restrict__ d_raw_rgb, int64_t ray_stride,
                  const uint8_t* __restrict__ stash, uint8_t* __restrict__ gstash, const float* __restrict__ absmax,
                  int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  const BwdSmem sl = bwd_smem_layout(lay.small_bytes);
  const uint32_t sA = base + sl.A, sW = base + sl.W, sBars = base + sl.bars;
  float* s_small = reinterpret_cast<float*>(gbase + sl.small);
  BwdStage* stages = reinterpret_cast<BwdStage*>(gbase + sl.tab);
  BwdPhase* phases = reinterpret_cast<BwdPhase*>(gbase + sl.tab + BWD_MAX_STAGES * sizeof(BwdStage));
  int* counts = reinterpret_cast<int*>(gbase + sl.tab + BWD_MAX_STAGES * sizeof(BwdStage) + BWD_MAX_PHASES * sizeof(BwdPhase));
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(gbase + sl.tmem_ptr);
  auto bar = [&](int i) -> uint32_t { return sBars + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;

  // ---- the per-tile program: phase p's epilogue reads the accumulator of MMA group p - 1 (BwdPhase.mask_layer holds
  // the TMEM column of that accumulator here) and writes the A operand of group p
  if (tid == 0) {
    int ns = 0, np = 0;
    auto stage = [&](uint32_t w_off, uint32_t bytes, int a_kb, int N, int tcol, int accum, int last) {
      BwdStage& s = stages[ns++];
      s.w_off = lay.stream_bytes + w_off; s.bytes = bytes; s.a_kb = a_kb; s.N = N; s.tcol = tcol; s.accum = accum; s.nk = 4;
      s.last = last;
    };
    auto phase = [&](int kind, int nch, int mask_blk, int g_blk, int tsrc) {
      BwdPhase& p = phases[np++];
      p.kind = kind; p.nch = nch; p.mask_blk = mask_blk; p.g_blk = g_blk; p.n_wait = nch; p.mask_layer = tsrc;
    };
    phase(MK_B_PRE, 2, MIP_S_H1, MIP_G_H1, 0);
    for (int kb = 0; kb < 2; ++kb) stage(lay.wt_h1 + (uint32_t)kb * 16384u, 16384u, kb, 128, 256, kb > 0, kb == 1);
    phase(MK_B_T, 2, MIP_S_H0, MIP_G_H0, 256);
    for (int kb = 0; kb < 2; ++kb) stage(lay.wt_h0 + (uint32_t)kb * 32768u, 32768u, kb, 256, 0, kb > 0, kb == 1);
    phase(MK_B_BASE, 4, MIP_S_OUT(MIP_NBASE - 1), MIP_G_OUT(MIP_NBASE - 1), 0);
    int tcol = 256;
    for (int l = MIP_NBASE - 1; l >= 1; --l) {
      for (int kb = 0; kb < 4; ++kb) stage(lay.wt_off[l] + (uint32_t)kb * 32768u, 32768u, kb, 256, tcol, kb > 0, kb == 3);
      phase(MK_B_T, 4, MIP_S_OUT(l - 1), MIP_G_OUT(l - 1), tcol);
      tcol ^= 256;
    }
    counts[0] = ns; counts[1] = np;
  }
  if (warp == TC_EPI_WARPS && lane == 0) {
    for (int i = 0; i < TC_NS; ++i) { mbar_init(bar(BAR_W_FULL(i)), 1); mbar_init(bar(BAR_W_EMPTY(i)), 1); }
    for (int i = 0; i < 5; ++i) mbar_init(bar(BAR_A_READY(i)), TC_EPI_WARPS);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_STASH_DONE), 1);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(base + sl.tmem_ptr, TC_TMEM_COLS);
  for (int i = tid; i < lay.small_floats; i += TC_THREADS) s_small[i] = reinterpret_cast<const float*>(packed)[i];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int n_stages = counts[0], n_phases = counts[1];

  if (warp == TC_EPI_WARPS) {
    // ======================================================================== weight producer + gradient-stash writer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, s_par = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint8_t* gs_tile = gstash + (size_t)tile * MIP_GSTASH_BLOCKS * TC_BLOCK_BYTES;
        auto stash_phase = [&](int g) {
          const int nch = phases[g].nch, g_blk = phases[g].g_blk;
          for (int kb = 0; kb < nch; ++kb) {
            mbar_wait(bar(BAR_A_READY(kb)), (s_par >> kb) & 1u, dbg, 6);
            s_par ^= 1u << kb;
            bulk_s2g(gs_tile + (size_t)(g_blk + kb) * TC_BLOCK_BYTES, sA + (uint32_t)kb * TC_KB_BYTES, TC_KB_BYTES);
            bulk_commit_group();
          }
          bulk_wait_group_read0();
          mbar_arrive(bar(BAR_STASH_DONE));
        };
        int grp = 0;
        bool group_start = true;
        for (int i = 0; i < n_stages; ++i) {
          const BwdStage& s = stages[i];
          if (group_start && grp >= 1) stash_phase(grp - 1);
          group_start = false;
          mbar_wait(bar(BAR_W_EMPTY(stage)), phase ^ 1u, dbg, 1);
          mbar_arrive_expect_tx(bar(BAR_W_FULL(stage)), s.bytes);
          bulk_g2s(sW + stage * TC_STAGE_BYTES, packed + lay.small_bytes + s.w_off, s.bytes, bar(BAR_W_FULL(stage)));
          if (++stage == TC_NS) { stage = 0; phase ^= 1u; }
          if (s.last) { ++grp; group_start = true; }
        }
        for (int g = grp - 1; g < n_phases; ++g) stash_phase(g);
      }
      bulk_wait_group0();
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    uint32_t stage = 0, phase = 0, a_par = 0;
    const uint64_t desc_a0 = umma_desc_sw128(sA), desc_w0 = umma_desc_sw128(sW);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int ph = 0;
      bool group_start = true;
      for (int i = 0; i < n_stages; ++i) {
        const BwdStage s = stages[i];
        if (group_start) {
          for (int kb = 0; kb < phases[ph].n_wait; ++kb) {
            mbar_wait(bar(BAR_A_READY(kb)), (a_par >> kb) & 1u, dbg, 2);
            a_par ^= 1u << kb;
          }
          group_start = false;
        }
        mbar_wait(bar(BAR_W_FULL(stage)), phase, dbg, 3);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_16(TC_M, s.N, F16 ? 0 : 1);
        const uint64_t a0 = desc_a0 + (uint64_t)(s.a_kb * (TC_KB_BYTES >> 4));
        const uint64_t b0 = desc_w0 + (uint64_t)(stage * (TC_STAGE_BYTES >> 4));
        if (elect_one_sync()) {
          tc_mma_kblock<4>(tmem_base + (uint32_t)s.tcol, a0, b0, idesc, s.accum ? 1u : 0u);
          tc_commit(bar(BAR_W_EMPTY(stage)));
          if (s.last) tc_commit(bar(BAR_ACC_FULL));
        }
        __syncwarp();
        if (++stage == TC_NS) { stage = 0; phase ^= 1u; }
        if (s.last) { ++ph; group_start = true; }
      }
      // the tile's last phase (G_0) also arrives on a_ready[]: consume it to keep the parities in step across tiles
      for (int kb = 0; ph < n_phases && kb < phases[ph].n_wait; ++kb) {
        mbar_wait(bar(BAR_A_READY(kb)), (a_par >> kb) & 1u, dbg, 5);
        a_par ^= 1u << kb;
      }
    }
  } else {
    // ======================================================================== epilogue warps
    const int q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    uint32_t acc_par = 0, stash_par = 0;
    const float gscale = F16 ? grad_scale_of(absmax) : 1.f;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gi = tile * TC_M + row;
      const bool valid = gi < M;
      float ds = 0.f, dc0 = 0.f, dc1 = 0.f, dc2 = 0.f;
      if (valid) {
        const int64_t r = gi / S;
        const int64_t o = r * ray_stride + (gi - r * S);
        ds = d_raw_sigma[o] * gscale;
        dc0 = d_raw_rgb[o * 3 + 0] * gscale; dc1 = d_raw_rgb[o * 3 + 1] * gscale; dc2 = d_raw_rgb[o * 3 + 2] * gscale;
      }
      const uint8_t* st_tile = stash + (size_t)tile * MIP_STASH_BLOCKS * TC_BLOCK_BYTES;
      const uint32_t x_sw = (uint32_t)row & 7u;
      for (int pi = 0; pi < n_phases; ++pi) {
        const BwdPhase P = phases[pi];
        // the masks of the whole phase (the stashed rectified outputs: > 0 <=> bits != 0) are fetched BEFORE the accumulator
        // wait: their global-memory latency hides behind the MMAs
        uint4 mq[4][2];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb < P.nch) {
            const uint8_t* mb = st_tile + (size_t)(P.mask_blk + kb) * TC_BLOCK_BYTES + (uint32_t)row * 128u;
            mq[kb][0] = __ldg(reinterpret_cast<const uint4*>(mb + ((((uint32_t)(cg * 2)) ^ x_sw) << 4)));
            mq[kb][1] = __ldg(reinterpret_cast<const uint4*>(mb + ((((uint32_t)(cg * 2 + 1)) ^ x_sw) << 4)));
          }
        }
        if (P.kind != MK_B_PRE) {
          mbar_wait(bar(BAR_ACC_FULL), acc_par, dbg, 4);
          acc_par ^= 1u;
          tc_fence_after();
        }
        if (!(pi == 0 && tile == (int64_t)blockIdx.x)) {   // the previous phase's A blocks have been copied to the gstash
          mbar_wait(bar(BAR_STASH_DONE), stash_par, dbg, 7);
          stash_par ^= 1u;
        }
        const uint32_t tsrc = tmem_base + lane_addr + (uint32_t)P.mask_layer + (uint32_t)(cg * TC_CPT);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb >= P.nch) break;
          const int col0 = kb * 64 + cg * TC_CPT;
          float v[16];
          if (P.kind == MK_B_PRE) {
            const float* rw = s_small + lay.off_rw + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = dc0 * rw[j] + dc1 * rw[MIP_WH + j] + dc2 * rw[2 * MIP_WH + j];
          } else {
            uint32_t r[16];
            tmem_ld16(tsrc + 64u * (uint32_t)kb, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (P.kind == MK_B_BASE) {
            const float* dw = s_small + lay.off_dw + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(ds, dw[j], v[j]);
          }
          const uint32_t mw[8] = {mq[kb][0].x, mq[kb][0].y, mq[kb][0].z, mq[kb][0].w,
                                  mq[kb][1].x, mq[kb][1].y, mq[kb][1].z, mq[kb][1].w};
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t bits = (j & 1) ? (mw[j >> 1] >> 16) : (mw[j >> 1] & 0xffffu);
            if (bits == 0u || !valid) v[j] = 0.f;
          }
          store_row16<F16, false, true>(sA + (uint32_t)kb * TC_KB_BYTES, row, cg * 2, v);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_A_READY(kb)));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// transposed weight stream of the mip field (layout: mip_tc_layout.h, wt_off / wt_h0 / wt_h1)
__global__ void mip_pack_tc_tstream_kernel(MipTcLayout tl, MipLayout ml, const float* __restrict__ master,
                                           uint16_t* __restrict__ tstream, int fp16) {
  const uint32_t n_elems = tl.tstream_bytes / 2;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += gridDim.x * blockDim.x) {
    const uint32_t byte = e * 2;
    int ml_idx, rows, koff;
    uint32_t off;
    if (byte >= tl.wt_h1) { ml_idx = MIP_L_H1; rows = MIP_WH; koff = 0; off = byte - tl.wt_h1; }
    else if (byte >= tl.wt_h0) { ml_idx = MIP_L_H0; rows = MIP_W; koff = MIP_KD; off = byte - tl.wt_h0; }
    else {
      int l = 1;
      while (l + 1 < MIP_NBASE && byte >= tl.wt_off[l + 1]) ++l;
      ml_idx = l; rows = MIP_W; koff = (l == MIP_SKIP) ? MIP_KX : 0; off = byte - tl.wt_off[l];
    }
    const uint32_t kb_bytes = (uint32_t)rows * 128u;
    const int kb = (int)(off / kb_bytes);
    const uint32_t rem = off % kb_bytes;
    const int j = (int)(rem >> 7);                                   // input feature (B operand row)
    const uint32_t inrow = rem & 127u;
    const int chunk = (int)((inrow >> 4) ^ ((uint32_t)j & 7u));
    const int nn = kb * 64 + chunk * 8 + (int)((inrow & 15u) >> 1);  // output feature
    const int K = ml.K[ml_idx], N = ml.N[ml_idx];
    const float v = (nn < N) ? master[ml.m_w[ml_idx] + (int64_t)nn * K + koff + j] : 0.f;
    tstream[e] = fp16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int star_mip_tc_pack_tstream(const MipTcLayout& tl, const MipLayout& ml, const float* master, void* tstream, int fp16,
                             cudaStream_t st) {
  mip_pack_tc_tstream_kernel<<<148 * 4, 256, 0, st>>>(tl, ml, master, (uint16_t*)tstream, fp16);
  return star_check_launch();
}

size_t star_mip_tc_stash_bytes(int64_t n_samples) {
  return (size_t)((n_samples + 127) / 128) * MIP_STASH_BLOCKS * TC_BLOCK_BYTES;
}
size_t star_mip_tc_gstash_bytes(int64_t n_samples) {
  return (size_t)((n_samples + 127) / 128) * MIP_GSTASH_BLOCKS * TC_BLOCK_BYTES + 256;
}

int star_mip_tc_backward(const void* packed, int R, int S, const float* d_raw_sigma, const float* d_raw_rgb,
                         int64_t ray_stride, const void* stash, void* gstash, float* grad_flat, int fp16,
                         cudaStream_t st) {
  MipTcLayout tl;
  MipLayout ml;
  star_make_mip_tc_layout(&tl);
  star_make_mip_layout(&ml);
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* absmax = reinterpret_cast<float*>((uint8_t*)gstash + star_mip_tc_gstash_bytes(M) - 256);
  if (fp16) {
    cudaError_t e = cudaMemsetAsync(absmax, 0, 4, st);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    int64_t b = (M + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    grad_absmax_kernel<<<(int)b, 256, 0, st>>>(d_raw_sigma, d_raw_rgb, S, M, ray_stride, absmax);
    int rc = star_check_launch();
    if (rc) return rc;
  }
  {   // ---- 1. dX chain
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    const BwdSmem sl = bwd_smem_layout(tl.small_bytes);
    auto kern = fp16 ? mip_bwd_tc_kernel<true> : mip_bwd_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    kern<<<grid, TC_THREADS, sl.total, st>>>(tl, (const uint8_t*)packed, S, M, d_raw_sigma, d_raw_rgb, ray_stride,
                                              (const uint8_t*)stash, (uint8_t*)gstash, absmax, star_watchdog_dev(STAR_WD_MIP_DX));
    int rc = star_check_launch();
    if (rc) return rc;
  }
  {   // ---- 2. dW / db through the shared dW kernel
    DwPlan plan;
    int n = 0;
    auto item = [&](int g_blk, int a_blk, int n_a, int ml_idx, int half, int k0, int k_valid, bool bias, int perm) {
      DwItem& it = plan.it[n++];
      it.g_blk = g_blk + 2 * half; it.a_blk = a_blk; it.n_a = n_a;
      it.K = ml.K[ml_idx]; it.k0 = k0; it.k_valid = k_valid;
      it.w_off = ml.m_w[ml_idx] + (int64_t)half * 128 * it.K;
      it.b_off = bias ? ml.m_b[ml_idx] + half * 128 : -1;
      it.ipe_perm = perm;
    };
    for (int h = 0; h < 2; ++h) {
      item(MIP_G_OUT(0), MIP_S_IPE, 3, 0, h, 0, MIP_KX, true, 1);                       // layer 0: the encoding
      for (int l = 1; l < MIP_NBASE; ++l) {
        const int k0 = (l == MIP_SKIP) ? MIP_KX : 0;
        item(MIP_G_OUT(l), MIP_S_OUT(l - 1), 4, l, h, k0, MIP_W, true, 0);              // x part
        if (l == MIP_SKIP) item(MIP_G_OUT(l), MIP_S_IPE, 3, l, h, 0, MIP_KX, false, 1);  // re-concatenated encoding
      }
    }
    item(MIP_G_H0, MIP_S_OUT(MIP_NBASE - 1), 4, MIP_L_H0, 0, MIP_KD, MIP_W, true, 0);   // head 0: base part
    item(MIP_G_H0, MIP_S_DIRS, 1, MIP_L_H0, 0, 0, MIP_KD, false, 0);                     //         encoded dirs
    item(MIP_G_H1, MIP_S_H0, 2, MIP_L_H1, 0, 0, MIP_WH, true, 0);                        // head 1
    plan.n_items = n;
    int splits = sms / n;
    if (splits < 1) splits = 1;
    if ((int64_t)splits > ntiles) splits = (int)ntiles;
    plan.splits = splits;
    const size_t smem = (size_t)(DW_NSTAGE * DW_STAGE_BLOCKS + 1) * TC_BLOCK_BYTES + 256 + 1024;
    auto kern = fp16 ? dw_tc_kernel<true> : dw_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
    kern<<<n * splits, DW_THREADS, smem, st>>>(plan, MIP_STASH_BLOCKS, MIP_GSTASH_BLOCKS, (const uint8_t*)stash,
                                               (const uint8_t*)gstash, ntiles, grad_flat, absmax, star_watchdog_dev(STAR_WD_DW));
    int rc = star_check_launch();
    if (rc) return rc;
  }
  {   // ---- 3. density / rgb heads
    int blocks = (int)(ntiles < HEAD_GRAD_CTAS_PER_SM * sms ? ntiles : HEAD_GRAD_CTAS_PER_SM * sms);
    const HeadGradCfg hc{MIP_STASH_BLOCKS, MIP_S_OUT(MIP_NBASE - 1), MIP_S_H1, ml.m_w[MIP_L_DENS], ml.m_b[MIP_L_DENS],
                         ml.m_w[MIP_L_RGB], ml.m_b[MIP_L_RGB]};
    head_grad_tc_kernel<<<blocks, 256, 0, st>>>(hc, (const uint8_t*)stash, d_raw_sigma, d_raw_rgb, ray_stride, S, M, fp16,
                                                grad_flat);
    return star_check_launch();
  }
}
