// Device-side pieces shared by the tensor-core MLP kernels (forward: mlp_tc.cu, backward: mlp_tc_bwd.cu):
// tile / role constants, the positional-encoding slice generator and the swizzled 16-bit row store.
#pragma once
#include "star_common.cuh"
#include "tc_common.cuh"
#include "mlp_tc_layout.h"

#define TC_M 128
#define TC_NS 4
#define TC_STAGE_BYTES 32768
#define TC_KB_BYTES 16384          // one A K-block: 128 rows x 128 B
#define TC_EPI_WARPS 16            // 4 per TMEM lane quadrant: each thread owns 1 row x 16 of the 64 columns of a K-block
#define TC_EPI_THREADS (32 * TC_EPI_WARPS)
#define TC_THREADS (TC_EPI_THREADS + 64)
#define TC_TMEM_COLS 512
#define TC_CPT 16                  // columns per thread per K-block

struct TcSmem {
  uint32_t A, AD, W, small, part, bars, tmem_ptr;   // byte offsets from the 1024-aligned base
  uint32_t total;
};
#define TC_MAX_NS 8                 // barrier index stride (the pipelined dX kernel runs a ring of 8 half-size stages)
__host__ __device__ static inline TcSmem tc_smem_layout(uint32_t small_bytes) {
  TcSmem s;
  uint32_t o = 0;
  s.A = o; o += 4 * TC_KB_BYTES;
  s.AD = o; o += TC_KB_BYTES;
  s.W = o; o += (uint32_t)TC_NS * TC_STAGE_BYTES;
  s.small = o; o += small_bytes;
  s.part = s.AD;    // head partial sums live in the never-read half (columns 32..63) of the dirs block
  s.bars = o; o += 40 * 8;
  s.tmem_ptr = o; o += 16;
  s.total = o + 1024;   // slack for aligning the dynamic smem base
  return s;
}

// barrier indices inside the bars block
#define BAR_W_FULL(i) (i)
#define BAR_W_EMPTY(i) (TC_MAX_NS + (i))
#define BAR_A_READY(i) (2 * TC_MAX_NS + (i))      // 0..3: A K-blocks, 4: encoded-dirs block
#define BAR_ACC_FULL (2 * TC_MAX_NS + 5)
#define BAR_STASH_DONE (2 * TC_MAX_NS + 6)   // training: the bulk store of A block kb has read shared memory (one barrier per block)
#define BAR_STASH_DONE_KB(kb) (BAR_STASH_DONE + (kb))
#define BAR_S_READY (2 * TC_MAX_NS + 13)     // forward kernel: lin_in's operand (ring slot S) written by all 16 epilogue warps
#define BAR_ACC_FULL_T (2 * TC_MAX_NS + 14)  // forward kernel: the T region's own completion barrier (BAR_ACC_FULL = X's)

// ---------------------------------------------------------------------------------------------- encode
// 16 consecutive columns [C0, C0+16) of the positional encoding [x, sin(2^k x), cos(2^k x)]_k (embedder.py:90-97;
// NV = 3 + 6 L valid columns, zero beyond).  The lowest octave of the slice comes from sincosf, the following
// ones from the double-angle recurrence (error doubles per octave from ~6e-8: < 1e-6 here, far below the
// 16-bit operand resolution).  Everything is resolved at compile time after unrolling.
template <int C0, int NV>
__device__ __forceinline__ void encode_slice(const float (&p)[3], const float* __restrict__ sc, float (&e)[16]) {
  constexpr int LAST = (C0 + 15 < NV) ? C0 + 15 : NV - 1;
  constexpr int K_LO = (C0 < 3) ? 0 : (C0 - 3) / 6;
  constexpr int K_HI = (LAST < 3) ? -1 : (LAST - 3) / 6;
#pragma unroll
  for (int j = 0; j < 16; ++j) e[j] = (C0 + j < 3) ? p[(C0 + j) % 3] : 0.f;
  if (K_HI >= K_LO) {
    float sn[3], cs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sincosf(p[c] * (float)(1 << K_LO), &sn[c], &cs[c]);
#pragma unroll
    for (int k = K_LO; k <= K_HI; ++k) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = C0 + j;
        if (col >= 3 && col < NV && (col - 3) / 6 == k) {
          const int w = (col - 3) % 6;
          e[j] = (w < 3) ? sn[w % 3] : cs[w % 3];
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
    for (int j = 0; j < 16; ++j) e[j] *= sc[C0 + j];
  }
}

// write 16 fp32 values as 16-bit operands into 16-byte chunks ch0, ch0+1 of row `row` of a SW128 K-block
// `gblock` (nullable): the same two chunks also go to the 16 KB stash block at that global address (same
// swizzled offsets), for the backward pass.  kblock_saddr == 0: global copy only.
template <bool FP16, bool RELU, bool SAT = false>
__device__ __forceinline__ void store_row16(uint32_t kblock_saddr, int row, int ch0, const float (&v)[16],
                                            uint8_t* gblock = nullptr) {
  const uint32_t x = (uint32_t)row & 7u;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)(ch0 + c)) ^ x) << 4);
    const uint32_t p0 = pack_16x2<FP16, RELU, SAT>(v[8 * c + 0], v[8 * c + 1]), p1 = pack_16x2<FP16, RELU, SAT>(v[8 * c + 2], v[8 * c + 3]),
                   p2 = pack_16x2<FP16, RELU, SAT>(v[8 * c + 4], v[8 * c + 5]), p3 = pack_16x2<FP16, RELU, SAT>(v[8 * c + 6], v[8 * c + 7]);
    if (kblock_saddr != 0u) st_shared_v4(kblock_saddr + off, p0, p1, p2, p3);
    if (gblock != nullptr) *reinterpret_cast<uint4*>(gblock + off) = make_uint4(p0, p1, p2, p3);
  }
}

