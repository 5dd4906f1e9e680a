// Layout of the tensor-core tier of the mip-NeRF field (csrc/mip_tc.cu forward, csrc/mip_tc_bwd.cu backward): the layer
// program, the packed weight image (small fp32 section | forward weight stream | transposed stream for the dX GEMMs) and
// the per-tile activation / gradient stashes of a training pass.
#pragma once
#include "mip_layout.h"
#include "mlp_tc_layout.h"

#define MIP_TC_NL 10
#define BAR_X_DONE (2 * TC_MAX_NS + 14)

enum MipTcKind { MK_HID = 0, MK_BASE_OUT = 1, MK_H0 = 2, MK_H1 = 3 };

struct MipTcLayer {
  int nkb;          // K-blocks issued before the (optional) mid-layer barrier
  int nkb_extra;    // layer 4: 3 K-blocks of the re-encoded input; H0: 1 K-block of encoded dirs
  int N, region, kind;
  int bias_off;     // float offset in the small section
  uint32_t w_off;   // byte offset of the first K-block in the weight stream
};

struct MipTcLayout {
  MipTcLayer L[MIP_TC_NL];
  int off_dw, off_db, off_rw, off_rb, off_freq;   // float offsets: density head, rgb head, frequency table
  int small_floats;
  uint32_t small_bytes, stream_bytes;
  // transposed weight stream (dX GEMMs: B operand rows = INPUT features j, K-blocks over the output features n):
  //   wt_off[l], l = 1..7: x part of base layer l, 4 K-blocks of [256 rows][64] (32 KB);  wt_h0: base part of head
  //   layer 0, 2 K-blocks of [256][64];  wt_h1: head layer 1, 2 K-blocks of [128][64] (16 KB)
  uint32_t wt_off[MIP_NBASE], wt_h0, wt_h1, tstream_bytes;
};

// Training stashes, per 128-sample tile, in 16 KB blocks of the swizzled operand image (TC_BLOCK_BYTES):
//   activations: [IPE 3][dirs 1][relu(out_l) 4 each, l = 0..7][relu(h0) 2][relu(h1) 2]         = 40 blocks
//   gradients  : [G_l 4 each, l = 0..7][G_h0 2][G_h1 2]   (dL/d(pre-activation output of a layer)) = 36 blocks
#define MIP_S_IPE 0
#define MIP_S_DIRS 3
#define MIP_S_OUT(l) (4 + 4 * (l))
#define MIP_S_H0 36
#define MIP_S_H1 38
#define MIP_STASH_BLOCKS 40
#define MIP_G_OUT(l) (4 * (l))
#define MIP_G_H0 32
#define MIP_G_H1 34
#define MIP_GSTASH_BLOCKS 36

static inline void star_make_mip_tc_layout(MipTcLayout* o) {
  int fo = 0;
  uint32_t wo = 0;
  auto add = [&](int i, int nkb, int extra, int N, int region, int kind) {
    MipTcLayer& l = o->L[i];
    l.nkb = nkb; l.nkb_extra = extra; l.N = N; l.region = region; l.kind = kind;
    l.bias_off = fo; fo += MIP_W;
    l.w_off = wo; wo += (uint32_t)(nkb + extra) * (uint32_t)N * 128u;
  };
  add(0, 3, 0, MIP_W, 0, MK_HID);
  for (int l = 1; l < MIP_NBASE; ++l)
    add(l, 4, l == MIP_SKIP ? 3 : 0, MIP_W, l & 1, l == MIP_NBASE - 1 ? MK_BASE_OUT : MK_HID);
  add(8, 4, 1, MIP_WH, 0, MK_H0);
  add(9, 2, 0, MIP_WH, 1, MK_H1);
  o->off_dw = fo; fo += MIP_W;
  o->off_db = fo; fo += 4;
  o->off_rw = fo; fo += 3 * MIP_WH;
  o->off_rb = fo; fo += 4;
  o->off_freq = fo; fo += MIP_FREQ_FLOATS;
  o->small_floats = fo;
  o->small_bytes = ((uint32_t)fo * 4u + TC_SMALL_ALIGN - 1) / TC_SMALL_ALIGN * TC_SMALL_ALIGN;
  o->stream_bytes = wo;
  uint32_t wt = 0;
  o->wt_off[0] = 0;
  for (int l = 1; l < MIP_NBASE; ++l) { o->wt_off[l] = wt; wt += 4u * 32768u; }
  o->wt_h0 = wt; wt += 2u * 32768u;
  o->wt_h1 = wt; wt += 2u * 16384u;
  o->tstream_bytes = wt;
}


// Kernel-internal K order of the integrated positional encoding (the packed weights are permuted to match,
// mip_pack_tc_stream_kernel): column 2 p + t, p = c * 24 + k the (axis, frequency) pair, t = 0: e sin(a), t = 1:
// e sin(a + pi/2); columns 144..146 the raw mean; zero padding up to 192.  A thread's 16 columns are 8 whole pairs, so
// the damping factor and the range reduction are shared by the two features of a pair (the reference evaluates
// sin(fl(a + pi/2)); cos of the reduced argument differs by < ulp(a)/2, far below the 16-bit operand resolution).
__host__ __device__ __forceinline__ int ipe_master_col(int col) {   // kernel column -> reference feature index, -1 = padding
  if (col < 6 * MIP_NF) return (col & 1) * 3 * MIP_NF + (col >> 1);
  return col < MIP_KX ? col : -1;
}

