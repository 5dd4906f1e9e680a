// Packed weight image of one NeRF MLP for the tcgen05 (bf16) tier.
//
//   [ small section, fp32 ]  per-layer epilogue bias vectors (256 floats each; for fc_1 layers the
//                            CUMULATIVE bias  b_in + sum_{i<=j} b1_i, because the residual stream lives in
//                            TMEM without biases and fc_1 accumulates onto it), alpha_linear weight/bias,
//                            rgb_linear weight/bias.  Copied to shared memory once per CTA.
//   [ weight stream, bf16 ]  the GEMM layers in execution order, each cut into K-blocks of 64 input
//                            features; one K-block = [N rows][64] bf16 already in the 128-byte-swizzled
//                            K-major shared-memory image that tcgen05.mma reads (N*128 bytes), so that one
//                            1-D bulk copy per K-block lands it ready to use.
// Layers (models/nerf.py:150-163, models/resnet.py:103-110):  lin_in (K 63->64), {fc_0, fc_1} x n_blocks,
// lin_out, feature_linear, views_linears.0 (K = 256 features + 27 encoded dirs padded to 64, N = 128).
#pragma once
#include "mlp_layout.h"

#define TC_SMALL_ALIGN 1024

struct TcLayer {
  int kind;            // LayerKind
  int nkb;             // K-blocks of 64
  int N;               // 256 or 128
  int region;          // accumulator region in TMEM: 0 = X (columns 0..255), 1 = T (columns 256..511)
  int bias_off;        // float offset of the epilogue bias vector in the small section
  uint32_t w_off;      // byte offset of the first K-block in the weight stream
};

struct TcLayout {
  int n_layers, n_blocks;
  TcLayer L[STAR_MAX_LAYERS];
  int off_alpha_w, off_alpha_b, off_rgb_w, off_rgb_b;   // float offsets in the small section
  int small_floats;
  uint32_t small_bytes;     // padded to TC_SMALL_ALIGN
  uint32_t stream_bytes;
  int n_stages;             // K-blocks per tile pass
};

static inline int star_make_tc_layout(const StarNetDesc* d, TcLayout* o) {
  if (d->n_blocks < 1 || 2 * d->n_blocks + 4 > STAR_MAX_LAYERS) return STAR_E_UNSUPPORTED;
  if (d->L_xyz != 10 || d->L_dir != 4) return STAR_E_UNSUPPORTED;   // tensor-core tier: reference configs only
  o->n_blocks = d->n_blocks;
  int n = 0, fo = 0, stages = 0;
  uint32_t wo = 0;
  auto add = [&](int kind, int nkb, int N, int region) {
    TcLayer& l = o->L[n++];
    l.kind = kind; l.nkb = nkb; l.N = N; l.region = region;
    l.bias_off = fo; fo += STAR_W;
    l.w_off = wo; wo += (uint32_t)nkb * (uint32_t)N * 128u;
    stages += nkb;
  };
  add(LK_IN, 1, STAR_W, 0);
  for (int b = 0; b < d->n_blocks; ++b) {
    add(LK_FC0, 4, STAR_W, 1);
    add(LK_FC1, 4, STAR_W, 0);
  }
  add(LK_OUT, 4, STAR_W, 1);
  add(LK_FEAT, 4, STAR_W, 0);
  add(LK_VIEWS, 5, STAR_WV, 1);
  o->n_layers = n;
  o->off_alpha_w = fo; fo += STAR_W;
  o->off_alpha_b = fo; fo += 4;
  o->off_rgb_w = fo; fo += 3 * STAR_WV;
  o->off_rgb_b = fo; fo += 4;
  o->small_floats = fo;
  o->small_bytes = ((uint32_t)fo * 4u + TC_SMALL_ALIGN - 1) / TC_SMALL_ALIGN * TC_SMALL_ALIGN;
  o->stream_bytes = wo;
  o->n_stages = stages;
  return STAR_OK;
}
