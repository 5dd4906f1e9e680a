// Tensor-core tier (tcgen05 / TMEM, bf16 or fp16 operands) of the mip-NeRF field forward pass (SURVEY.md row a12:
// models/mipnerf.py:53-100 -> nerfstudio NeRFField with integrated positional encoding).  Same machine as
// mlp_tc.cu: one persistent CTA per SM walks tiles of 128 samples with 16 epilogue warps, one weight-producer warp
// (1-D bulk copies of pre-swizzled K-blocks into a 4-stage ring) and one MMA-issuer thread; accumulators alternate
// between two 256-column TMEM regions so the epilogue of layer L overlaps the MMAs of layer L+1 K-block by K-block.
//
// Layer program per tile (K-blocks of 64 features; N = output width):
//   encode IPE (147 -> 192) into A[0..2], encoded dirs (27 -> 32) into AD
//   L0 : A[0..2]            N=256 ReLU                         L1..L3 : A[0..3] N=256 ReLU
//   L4 : A[0..3] (x part), then -- once those MMAs are done (x_done) -- the epilogue warps put the IPE back into
//        A[0..2] (each thread kept its 24 packed words) and 3 more K-blocks accumulate the skip connection's
//        encoding part (nerfstudio MLP: cat[enc, x]); this saves the 48 KB of shared memory the weight ring needs
//   L5..L7 : A[0..3] N=256 ReLU; L7's epilogue also takes the density head (fp32 dot on the rectified base_out)
//   H0 : A[0..3] + AD       N=128 ReLU -> A[0..1]              H1 : A[0..1] N=128 ReLU + rgb head (fp32 dot)
// Biases are fp32 and added in the epilogue; the two heads run in fp32 on un-rounded activations (as in mlp_tc.cu).
// Training (STASH): every GEMM input of the tile (IPE blocks, encoded dirs, the rectified output of each layer) is also
// written to the activation stash in the same swizzled 16 KB block image (mip_tc_layout.h) for the tensor-core backward
// pass (mip_tc_bwd.cu: dX chain, dW through dw_tc_kernel, head gradients).
#include "star_common.cuh"
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "mip_layout.h"
#include "mlp_tc_device.cuh"

#include "mip_tc_layout.h"

#define MIP_TWO_PI 6.2831854820251465f
#define MIP_PIO2 1.5707963705062866f

// sin for arguments up to ~1e5 at 16-bit-operand accuracy: two-term Cody-Waite reduction by 2 pi, then the SFU
__device__ __forceinline__ float sin_reduced(float a) {
  const float n = rintf(a * 0.15915494309189535f);
  float r = fmaf(-n, 6.28318548202514648f, a);
  r = fmaf(-n, -1.74845553146951715e-7f, r);
  return __sinf(r);
}

struct MipTcGeom {
  float mean[3], diag[3], d[3];
};

__device__ __forceinline__ MipTcGeom mip_tc_geom(const float* __restrict__ origins, const float* __restrict__ dirs,
                                                 const float* __restrict__ pose12, const float* __restrict__ bins,
                                                 int64_t gi, int S, float radius) {
  MipTcGeom g;
  const int64_t r = gi / S;
  const int s = (int)(gi - r * S);
  const float ox = origins[r * 3 + 0], oy = origins[r * 3 + 1], oz = origins[r * 3 + 2];
  const float dx = dirs[r * 3 + 0], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
  float o[3];
  if (pose12 != nullptr) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      o[i] = pose12[i * 3 + 0] * ox + pose12[i * 3 + 1] * oy + pose12[i * 3 + 2] * oz + pose12[9 + i];
      g.d[i] = pose12[i * 3 + 0] * dx + pose12[i * 3 + 1] * dy + pose12[i * 3 + 2] * dz;
    }
  } else {
    o[0] = ox; o[1] = oy; o[2] = oz;
    g.d[0] = dx; g.d[1] = dy; g.d[2] = dz;
  }
  // conical_frustum_to_gaussian with the reference's rounding sequence (see mip_f32.cu)
  const float st = bins[r * (S + 1) + s], en = bins[r * (S + 1) + s + 1];
  const float mu = __fdiv_rn(__fadd_rn(st, en), 2.f), hw = __fdiv_rn(__fsub_rn(en, st), 2.f);
  const float mu2 = __fmul_rn(mu, mu), hw2 = __fmul_rn(hw, hw), hw4 = __fmul_rn(hw2, hw2);
  const float den = __fadd_rn(__fmul_rn(3.f, mu2), hw2);
  const float tmean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.f, mu), hw2), den));
  const float dvar = __fsub_rn(__fdiv_rn(hw2, 3.f),
                               __fmul_rn(4.f / 15.f, __fdiv_rn(__fmul_rn(hw4, __fsub_rn(__fmul_rn(12.f, mu2), hw2)), __fmul_rn(den, den))));
  const float rvar = __fmul_rn(radius * radius,
                               __fsub_rn(__fadd_rn(__fdiv_rn(mu2, 4.f), __fmul_rn(5.f / 12.f, hw2)), __fdiv_rn(__fmul_rn(4.f / 15.f, hw4), den)));
  const float mag = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(g.d[0], g.d[0]), __fmul_rn(g.d[1], g.d[1])), __fmul_rn(g.d[2], g.d[2])), 1e-10f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    g.mean[c] = __fadd_rn(o[c], __fmul_rn(g.d[c], tmean));
    g.diag[c] = __fadd_rn(__fmul_rn(dvar, __fmul_rn(g.d[c], g.d[c])),
                          __fmul_rn(rvar, __fsub_rn(1.f, __fmul_rn(g.d[c], __fdiv_rn(g.d[c], mag)))));
  }
  return g;
}

__device__ __forceinline__ void ipe_pair(const MipTcGeom& g, const float* __restrict__ s_f, int p, float& f0, float& f1) {
  const int c = p / MIP_NF, k = p - c * MIP_NF;
  const float mc = c == 0 ? g.mean[0] : (c == 1 ? g.mean[1] : g.mean[2]);
  const float dc = c == 0 ? g.diag[0] : (c == 1 ? g.diag[1] : g.diag[2]);
  const float hv = -0.5f * __fmul_rn(dc, s_f[MIP_NF + k]);
  f0 = 0.f; f1 = 0.f;
  if (hv >= -104.f) {                    // below: exp underflows to 0 in fp32, the reference's features are exactly 0
    const float a = __fmul_rn(__fmul_rn(MIP_TWO_PI, mc), s_f[k]);
    const float n = rintf(a * 0.15915494309189535f);
    float r = fmaf(-n, 6.28318548202514648f, a);
    r = fmaf(-n, -1.74845553146951715e-7f, r);
    float sn, cs;
    __sincosf(r, &sn, &cs);
    const float e = __expf(hv);
    f0 = e * sn;
    f1 = e * cs;
  }
}

// 16 fp32 values -> 8 packed 16-bit pairs
template <bool FP16>
__device__ __forceinline__ void pack_row16(const float (&v)[16], uint32_t (&w)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack_16x2<FP16, false>(v[2 * i], v[2 * i + 1]);
}
__device__ __forceinline__ void store_words16(uint32_t kblock_saddr, int row, int ch0, const uint32_t (&w)[8],
                                              uint8_t* gblock = nullptr) {
  const uint32_t x = (uint32_t)row & 7u;
  const uint32_t o0 = (uint32_t)row * 128u + ((((uint32_t)ch0) ^ x) << 4), o1 = (uint32_t)row * 128u + ((((uint32_t)(ch0 + 1)) ^ x) << 4);
  st_shared_v4(kblock_saddr + o0, w[0], w[1], w[2], w[3]);
  st_shared_v4(kblock_saddr + o1, w[4], w[5], w[6], w[7]);
  if (gblock != nullptr) {     // training: the same chunks into the stash block at that global address
    *reinterpret_cast<uint4*>(gblock + o0) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(gblock + o1) = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// Encodes this thread's 3 x 16 columns, stores them into A[0..2] and keeps the packed words for the skip layer
template <bool FP16>
__device__ __forceinline__ void encode_ipe_blocks(const MipTcGeom& g, const float* __restrict__ s_f, uint32_t sA, int row,
                                                  int cg, uint32_t (&cache)[3][8], uint8_t* st_ipe) {
#pragma unroll
  for (int kb = 0; kb < 3; ++kb) {
    float e[16];
    const int col0 = kb * 64 + cg * 16;
    if (col0 + 16 <= 6 * MIP_NF) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ipe_pair(g, s_f, (col0 >> 1) + j, e[2 * j], e[2 * j + 1]);
    } else {                             // the block that holds the last pairs, the raw mean and the padding
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = col0 + 2 * j;
        e[2 * j] = 0.f; e[2 * j + 1] = 0.f;
        if (col + 1 < 6 * MIP_NF) ipe_pair(g, s_f, col >> 1, e[2 * j], e[2 * j + 1]);
        else {
          if (col == 6 * MIP_NF) { e[2 * j] = g.mean[0]; e[2 * j + 1] = g.mean[1]; }
          if (col == 6 * MIP_NF + 2) e[2 * j] = g.mean[2];
        }
      }
    }
    pack_row16<FP16>(e, cache[kb]);
    store_words16(sA + (uint32_t)kb * TC_KB_BYTES, row, cg * 2, cache[kb],
                  st_ipe != nullptr ? st_ipe + (size_t)kb * TC_BLOCK_BYTES : nullptr);
  }
}

// ---------------------------------------------------------------------------------------------- epilogue
struct MipEpi {
  uint32_t sA, a_ready0, tcol;
  const float* bias;
  const float* head_w;
  uint8_t* stash_out;     // training: first stash block of this layer's rectified output, else NULL
  int row, cg, lane;
};

template <int KIND, bool FP16>
__device__ __forceinline__ void mip_epilogue(const MipEpi& c, float (&h)[3]) {
  constexpr int NCH = (KIND == MK_H0 || KIND == MK_H1) ? 2 : 4;
  // two TMEM chunk loads in flight (measured best in mlp_tc.cu: 1 and NCH are slower)
  uint32_t r[NCH][TC_CPT];
#pragma unroll
  for (int kb = 0; kb < NCH && kb < 2; ++kb) tmem_ld16(c.tcol + 64u * (uint32_t)kb, r[kb]);
#pragma unroll
  for (int kb = 0; kb < NCH; ++kb) {
    const int col0 = kb * 64 + c.cg * TC_CPT;
    float4 bq[TC_CPT / 4];
#pragma unroll
    for (int j4 = 0; j4 < TC_CPT / 4; ++j4) bq[j4] = *reinterpret_cast<const float4*>(c.bias + col0 + 4 * j4);
    tmem_wait_ld();
    if (kb + 2 < NCH) tmem_ld16(c.tcol + 64u * (uint32_t)(kb + 2), r[kb + 2]);
    float v[TC_CPT];
#pragma unroll
    for (int j4 = 0; j4 < TC_CPT / 4; ++j4) {
      v[4 * j4 + 0] = __uint_as_float(r[kb][4 * j4 + 0]);
      v[4 * j4 + 1] = __uint_as_float(r[kb][4 * j4 + 1]);
      v[4 * j4 + 2] = __uint_as_float(r[kb][4 * j4 + 2]);
      v[4 * j4 + 3] = __uint_as_float(r[kb][4 * j4 + 3]);
      add_f32x2(v[4 * j4 + 0], v[4 * j4 + 1], bq[j4].x, bq[j4].y);
      add_f32x2(v[4 * j4 + 2], v[4 * j4 + 3], bq[j4].z, bq[j4].w);
    }
    if (KIND == MK_BASE_OUT) {       // density head on the rectified base_out (fp32)
#pragma unroll
      for (int j = 0; j < TC_CPT; ++j) h[0] = fmaf(fmaxf(v[j], 0.f), c.head_w[col0 + j], h[0]);
    }
    if (KIND == MK_H1) {             // rgb head on the rectified head output (fp32)
#pragma unroll
      for (int j = 0; j < TC_CPT; ++j) {
        const float x = fmaxf(v[j], 0.f);
        h[0] = fmaf(x, c.head_w[col0 + j], h[0]);
        h[1] = fmaf(x, c.head_w[MIP_WH + col0 + j], h[1]);
        h[2] = fmaf(x, c.head_w[2 * MIP_WH + col0 + j], h[2]);
      }
      if (c.stash_out != nullptr) store_row16<FP16, true>(0u, c.row, c.cg * 2, v, c.stash_out + (size_t)kb * TC_BLOCK_BYTES);
    } else {
      store_row16<FP16, true>(c.sA + (uint32_t)kb * TC_KB_BYTES, c.row, c.cg * 2, v,
                              c.stash_out != nullptr ? c.stash_out + (size_t)kb * TC_BLOCK_BYTES : nullptr);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (c.lane == 0) mbar_arrive(c.a_ready0 + 8u * (uint32_t)kb);
    }
  }
}

// ============================================================================================ forward
template <bool FP16, bool STASH>
__global__ void __launch_bounds__(TC_THREADS, 1)
mip_fwd_tc_kernel(const MipTcLayout lay, const uint8_t* __restrict__ packed, const float* __restrict__ origins,
                  const float* __restrict__ dirs, const float* __restrict__ pose12, const float* __restrict__ bins,
                  float radius, int S, int64_t M, float* __restrict__ raw_sigma, float* __restrict__ raw_rgb,
                  int64_t ray_stride, uint8_t* __restrict__ stash, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  const TcSmem sl = tc_smem_layout(lay.small_bytes);
  const uint32_t sA = base + sl.A, sAD = base + sl.AD, sW = base + sl.W, sBars = base + sl.bars;
  float* s_small = reinterpret_cast<float*>(gbase + sl.small);
  const float* s_f = s_small + lay.off_freq;
  auto part = [&](int r, int g) -> float* {   // head partial sums of column group g: unused half of the dirs block
    return reinterpret_cast<float*>(gbase + sl.part + (uint32_t)r * 128u + ((((uint32_t)(3 + g)) ^ ((uint32_t)r & 7u)) << 4));
  };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(gbase + sl.tmem_ptr);
  auto bar = [&](int i) -> uint32_t { return sBars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;

  if (warp == TC_EPI_WARPS && lane == 0) {
    for (int i = 0; i < TC_NS; ++i) { mbar_init(bar(BAR_W_FULL(i)), 1); mbar_init(bar(BAR_W_EMPTY(i)), 1); }
    for (int i = 0; i < 5; ++i) mbar_init(bar(BAR_A_READY(i)), TC_EPI_WARPS);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_X_DONE), 1);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(base + sl.tmem_ptr, TC_TMEM_COLS);
  for (int i = tid; i < lay.small_floats; i += TC_THREADS) s_small[i] = reinterpret_cast<const float*>(packed)[i];
  for (int i = tid; i < TC_KB_BYTES / 16; i += TC_THREADS)
    reinterpret_cast<uint4*>(gbase + sl.AD)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == TC_EPI_WARPS) {
    // ======================================================================== weight producer
    if (lane == 0) {
      const uint8_t* wstream = packed + lay.small_bytes;
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = 0; l < MIP_TC_NL; ++l) {
          const uint32_t bytes = (uint32_t)lay.L[l].N * 128u;
          const int n = lay.L[l].nkb + lay.L[l].nkb_extra;
          for (int kb = 0; kb < n; ++kb) {
            mbar_wait(bar(BAR_W_EMPTY(stage)), phase ^ 1u, dbg, 1);
            mbar_arrive_expect_tx(bar(BAR_W_FULL(stage)), bytes);
            bulk_g2s(sW + stage * TC_STAGE_BYTES, wstream + lay.L[l].w_off + (uint32_t)kb * bytes, bytes,
                     bar(BAR_W_FULL(stage)));
            if (++stage == TC_NS) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    // warp-uniform control flow (all lanes wait on the barriers), one elected lane issues the tcgen05 instructions
    {
      uint32_t stage = 0, phase = 0, a_par = 0;
      const uint64_t desc_a0 = umma_desc_sw128(sA), desc_ad = umma_desc_sw128(sAD), desc_w0 = umma_desc_sw128(sW);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = 0; l < MIP_TC_NL; ++l) {
          const int kind = lay.L[l].kind, nkb = lay.L[l].nkb, n = nkb + lay.L[l].nkb_extra;
          const uint32_t d_tmem = tmem_base + (lay.L[l].region ? 256u : 0u);
          const uint32_t idesc = umma_idesc_16(TC_M, lay.L[l].N, FP16 ? 0 : 1);
          // one K-block kb of the layer reading operand block idx (4 = encoded dirs, 2 of its 4 K-steps)
          auto kblock = [&](const int kb, const int idx) {
            mbar_wait(bar(BAR_A_READY(idx)), (a_par >> idx) & 1u, dbg, 2);
            a_par ^= 1u << idx;
            mbar_wait(bar(BAR_W_FULL(stage)), phase, dbg, 3);
            tc_fence_after();
            const uint64_t a0 = idx == 4 ? desc_ad : desc_a0 + (uint64_t)(idx * (TC_KB_BYTES >> 4));
            const uint64_t b0 = desc_w0 + (uint64_t)(stage * (TC_STAGE_BYTES >> 4));
            const uint32_t acc0 = kb > 0 ? 1u : 0u;
            if (elect_one_sync()) {
              if (idx == 4) tc_mma_kblock<2>(d_tmem, a0, b0, idesc, acc0);
              else tc_mma_kblock<4>(d_tmem, a0, b0, idesc, acc0);
              tc_commit(bar(BAR_W_EMPTY(stage)));
              if (kb == n - 1) tc_commit(bar(BAR_ACC_FULL));
            }
            __syncwarp();
            if (++stage == TC_NS) { stage = 0; phase ^= 1u; }
          };
          if (n == 4) {              // plain hidden layer, unrolled: K-block indices become immediates
            kblock(0, 0); kblock(1, 1); kblock(2, 2); kblock(3, 3);
          } else {
            for (int kb = 0; kb < nkb; ++kb) kblock(kb, kb);
            if (n > nkb) {
              if (kind == MK_H0) {
                kblock(nkb, 4);                         // encoded-dirs block
              } else {                                  // skip layer: x-part MMAs done -> A may be re-encoded
                if (elect_one_sync()) tc_commit(bar(BAR_X_DONE));
                __syncwarp();
                for (int e = 0; e < n - nkb; ++e) kblock(nkb + e, e);
              }
            }
          }
        }
      }
    }
  } else {
    // ======================================================================== epilogue warps
    const int q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t acc_par = 0, x_par = 0;
    MipEpi ctx;
    ctx.sA = sA; ctx.a_ready0 = bar(BAR_A_READY(0));
    ctx.row = row; ctx.cg = cg; ctx.lane = lane;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gi = tile * TC_M + row;
      const bool valid = gi < M;
      int64_t out_idx = 0;
      MipTcGeom g;
      if (valid) {
        const int64_t r = gi / S;
        out_idx = r * ray_stride + (gi - r * S);
        g = mip_tc_geom(origins, dirs, pose12, bins, gi, S, radius);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { g.mean[c] = 0.f; g.diag[c] = 1e30f; g.d[c] = 0.f; }
      }
      // ---- inputs: IPE -> A[0..2]; encoded dirs -> AD (columns 0..31: groups 0 and 1)
      uint32_t enc_cache[3][8];
      uint8_t* st_tile = STASH ? stash + (size_t)tile * MIP_STASH_BLOCKS * TC_BLOCK_BYTES : nullptr;
      encode_ipe_blocks<FP16>(g, s_f, sA, row, cg, enc_cache, STASH ? st_tile + (size_t)MIP_S_IPE * TC_BLOCK_BYTES : nullptr);
      if (cg < 2) {
        float e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = cg * 16 + j;
          float v = 0.f;
          if (col < 6 * MIP_NFD) {
            const int jj = col < 3 * MIP_NFD ? col : col - 3 * MIP_NFD;
            const int c = jj / MIP_NFD, k = jj - c * MIP_NFD;
            const float dc = c == 0 ? g.d[0] : (c == 1 ? g.d[1] : g.d[2]);
            float a = __fmul_rn(__fmul_rn(MIP_TWO_PI, dc), s_f[2 * MIP_NF + k]);
            if (col >= 3 * MIP_NFD) a = __fadd_rn(a, MIP_PIO2);
            v = sin_reduced(a);
          } else if (col < MIP_KD) {
            v = col == 6 * MIP_NFD ? g.d[0] : (col == 6 * MIP_NFD + 1 ? g.d[1] : g.d[2]);
          }
          e[j] = valid ? v : 0.f;
        }
        store_row16<FP16, false>(sAD, row, cg * 2, e, STASH ? st_tile + (size_t)MIP_S_DIRS * TC_BLOCK_BYTES : nullptr);
      } else if (STASH) {      // columns 32..63 of the stashed dirs block: zeros (the dW GEMM reads the whole block)
        const float z[16] = {0.f};
        store_row16<FP16, false>(0u, row, cg * 2, z, st_tile + (size_t)MIP_S_DIRS * TC_BLOCK_BYTES);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(BAR_A_READY(0)));
        mbar_arrive(bar(BAR_A_READY(1)));
        mbar_arrive(bar(BAR_A_READY(2)));
        mbar_arrive(bar(BAR_A_READY(4)));
      }
      // ---- layers
      for (int l = 0; l < MIP_TC_NL; ++l) {
        const MipTcLayer& L = lay.L[l];
        if (L.kind != MK_H0 && L.nkb_extra > 0) {
          // skip layer: once its x-part MMAs are complete the A blocks are free -> the encoding (kept as 24 packed
          // words per thread since the tile's first encode) goes back into A[0..2]
          mbar_wait(bar(BAR_X_DONE), x_par, dbg, 6);
          x_par ^= 1u;
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) store_words16(sA + (uint32_t)kb * TC_KB_BYTES, row, cg * 2, enc_cache[kb]);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(BAR_A_READY(0)));
            mbar_arrive(bar(BAR_A_READY(1)));
            mbar_arrive(bar(BAR_A_READY(2)));
          }
        }
        mbar_wait(bar(BAR_ACC_FULL), acc_par, dbg, 4);
        acc_par ^= 1u;
        tc_fence_after();
        ctx.tcol = tmem_base + (((uint32_t)(q * 32)) << 16) + (L.region ? 256u : 0u) + (uint32_t)(cg * TC_CPT);
        ctx.bias = s_small + L.bias_off;
        ctx.stash_out = STASH ? st_tile + (size_t)(l < MIP_NBASE ? MIP_S_OUT(l) : (l == 8 ? MIP_S_H0 : MIP_S_H1)) * TC_BLOCK_BYTES
                              : nullptr;
        float h[3] = {0.f, 0.f, 0.f};
        if (L.kind == MK_HID) {
          mip_epilogue<MK_HID, FP16>(ctx, h);
        } else if (L.kind == MK_BASE_OUT) {
          ctx.head_w = s_small + lay.off_dw;
          mip_epilogue<MK_BASE_OUT, FP16>(ctx, h);
          if (cg != 0) part(row, cg)[0] = h[0];
          named_bar_sync(1, TC_EPI_THREADS);
          if (cg == 0 && valid)
            raw_sigma[out_idx] = h[0] + part(row, 1)[0] + part(row, 2)[0] + part(row, 3)[0] + s_small[lay.off_db];
        } else if (L.kind == MK_H0) {
          mip_epilogue<MK_H0, FP16>(ctx, h);
        } else {
          ctx.head_w = s_small + lay.off_rw;
          mip_epilogue<MK_H1, FP16>(ctx, h);
          if (cg != 0) {
            float* d = part(row, cg);
            d[1] = h[0]; d[2] = h[1]; d[3] = h[2];
          }
          tc_fence_before();
          named_bar_sync(1, TC_EPI_THREADS);
          if (cg == 0 && valid) {
            float* o = raw_rgb + out_idx * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
              o[ch] = h[ch] + part(row, 1)[1 + ch] + part(row, 2)[1 + ch] + part(row, 3)[1 + ch] + s_small[lay.off_rb + ch];
          }
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

// ============================================================================================ packing
__global__ void mip_pack_tc_small_kernel(MipTcLayout tl, MipLayout ml, const float* __restrict__ master,
                                         const float* __restrict__ freqs, float* __restrict__ small) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tl.small_floats; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    bool done = false;
    for (int l = 0; l < MIP_TC_NL && !done; ++l) {
      const int o = i - tl.L[l].bias_off;
      if (o >= 0 && o < MIP_W) {
        done = true;
        const int ml_idx = l < MIP_NBASE ? l : (l == 8 ? MIP_L_H0 : MIP_L_H1);
        if (o < tl.L[l].N) v = master[ml.m_b[ml_idx] + o];
      }
    }
    if (!done) {
      if (i >= tl.off_dw && i < tl.off_dw + MIP_W) v = master[ml.m_w[MIP_L_DENS] + (i - tl.off_dw)];
      else if (i == tl.off_db) v = master[ml.m_b[MIP_L_DENS]];
      else if (i >= tl.off_rw && i < tl.off_rw + 3 * MIP_WH) v = master[ml.m_w[MIP_L_RGB] + (i - tl.off_rw)];
      else if (i >= tl.off_rb && i < tl.off_rb + 3) v = master[ml.m_b[MIP_L_RGB] + (i - tl.off_rb)];
      else if (i >= tl.off_freq && i < tl.off_freq + MIP_FREQ_FLOATS) v = freqs[i - tl.off_freq];
    }
    small[i] = v;
  }
}

__global__ void mip_pack_tc_stream_kernel(MipTcLayout tl, MipLayout ml, const float* __restrict__ master,
                                          uint16_t* __restrict__ stream, int fp16) {
  const uint32_t n_elems = tl.stream_bytes / 2;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += gridDim.x * blockDim.x) {
    const uint32_t byte = e * 2;
    int l = 0;
    while (l + 1 < MIP_TC_NL && byte >= tl.L[l + 1].w_off) ++l;
    const MipTcLayer& L = tl.L[l];
    const uint32_t off = byte - L.w_off, kb_bytes = (uint32_t)L.N * 128u;
    const int kb = (int)(off / kb_bytes);
    const uint32_t rem = off % kb_bytes;
    const int n = (int)(rem >> 7);
    const uint32_t inrow = rem & 127u;
    const int chunk = (int)((inrow >> 4) ^ ((uint32_t)n & 7u));   // undo the 128-byte swizzle
    const int kk = chunk * 8 + (int)((inrow & 15u) >> 1);
    const int ml_idx = l < MIP_NBASE ? l : (l == 8 ? MIP_L_H0 : MIP_L_H1);
    const int K = ml.K[ml_idx];
    int k = -1;   // master column of this element, -1 = zero padding
    if (l == 0) {
      k = ipe_master_col(kb * 64 + kk);                 // kernel-internal K order of the encoding
    } else if (l == MIP_SKIP) {
      if (kb < 4) k = MIP_KX + kb * 64 + kk;            // x part first (issue order), encoding part after x_done
      else k = ipe_master_col((kb - 4) * 64 + kk);
    } else if (l == 8) {
      if (kb < 4) k = MIP_KD + kb * 64 + kk;            // base_out part, then the encoded-dirs block
      else { k = kk; if (k >= MIP_KD) k = -1; }
    } else {
      k = kb * 64 + kk;
    }
    const float v = (k >= 0 && k < K) ? master[ml.m_w[ml_idx] + (int64_t)n * K + k] : 0.f;
    stream[e] = fp16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// ============================================================================================ host side
int star_mip_tc_pack_tstream(const MipTcLayout& tl, const MipLayout& ml, const float* master, void* tstream, int fp16,
                             cudaStream_t st);

size_t star_mip_tc_packed_bytes() {
  MipTcLayout tl;
  star_make_mip_tc_layout(&tl);
  return (size_t)tl.small_bytes + tl.stream_bytes + tl.tstream_bytes;
}

int star_mip_tc_pack(const float* master, const float* freqs, void* packed, int fp16, cudaStream_t st) {
  MipTcLayout tl;
  MipLayout ml;
  star_make_mip_tc_layout(&tl);
  star_make_mip_layout(&ml);
  mip_pack_tc_small_kernel<<<8, 256, 0, st>>>(tl, ml, master, freqs, (float*)packed);
  int rc = star_check_launch();
  if (rc) return rc;
  mip_pack_tc_stream_kernel<<<148 * 4, 256, 0, st>>>(tl, ml, master, (uint16_t*)((uint8_t*)packed + tl.small_bytes), fp16);
  rc = star_check_launch();
  if (rc) return rc;
  return star_mip_tc_pack_tstream(tl, ml, master, (uint8_t*)packed + tl.small_bytes + tl.stream_bytes, fp16, st);
}

int star_mip_tc_forward(const void* packed, const float* origins, const float* dirs, const float* pose12,
                        const float* bins, float radius, int R, int S, float* raw_sigma, float* raw_rgb,
                        int64_t ray_stride, void* stash, int fp16, cudaStream_t st) {
  MipTcLayout tl;
  star_make_mip_tc_layout(&tl);
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  const TcSmem sl = tc_smem_layout(tl.small_bytes);
  if (((uintptr_t)packed & 15) != 0) return STAR_E_ALIGN;
  auto kern = stash != nullptr ? (fp16 ? mip_fwd_tc_kernel<true, true> : mip_fwd_tc_kernel<false, true>)
                               : (fp16 ? mip_fwd_tc_kernel<true, false> : mip_fwd_tc_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
  if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
  kern<<<grid, TC_THREADS, sl.total, st>>>(tl, (const uint8_t*)packed, origins, dirs, pose12, bins, radius, S, M,
                                            raw_sigma, raw_rgb, ray_stride, (uint8_t*)stash, star_watchdog_dev(STAR_WD_MIP_FWD));
  return star_check_launch();
}
