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
  // training only (activation stash, see TC_BLOCK_BYTES below): first block of this layer's INPUT operand
  // (nkb blocks; the view layer's 5th block, the encoded dirs, is s_in + 4) and of its epilogue OUTPUT
  // (the next layer's input; for the view layer relu(h2), 2 blocks), and of dL/d(output) (N/64 blocks).
  int s_in, s_out, g_out;
  uint32_t wt_off;     // byte offset of the transposed K-blocks (dX GEMM operand) in the backward weight stream
};

struct TcLayout {
  int n_layers, n_blocks;
  TcLayer L[STAR_MAX_LAYERS];
  int off_alpha_w, off_alpha_b, off_rgb_w, off_rgb_b;   // float offsets in the small section
  int small_floats;
  uint32_t small_bytes;     // padded to TC_SMALL_ALIGN
  uint32_t stream_bytes;
  int n_stages;             // K-blocks per tile pass
  // training: blocks per 128-sample tile of the activation stash and of the gradient stash, and the size of
  // the backward (transposed) weight stream that follows the forward one in the packed image
  int stash_blocks, gstash_blocks;
  int mask_blk0;            // first block of the tile's ReLU bit masks (TC_MASK_BYTES per layer, see below)
  uint32_t wt_dirs_off;     // transposed view-layer rows of the encoded dirs (objects only): 2 K-blocks x [64 rows][128 B]
  uint32_t tstream_bytes;
};

// Stash format: per tile of 128 samples a sequence of 16 KB blocks, each [128 sample rows][64 features] of
// 16-bit values in exactly the 128-byte-swizzled image the kernels keep in shared memory (row r at r*128,
// 16-byte chunk c at position c ^ (r & 7)).  A block is therefore at once a K-major operand (K = features)
// for the forward / dX GEMMs and an MN-major operand (K = samples) for the dW GEMMs, and moves with one bulk copy.
#define TC_BLOCK_BYTES 16384
// ReLU masks for the dX pass: for every layer l whose epilogue output is rectified (lin_in, fc_0, fc_1) the forward also
// leaves one bit per output feature, [4 column groups][128 rows] 64-bit words (bit 16 kb + j of word (cg, row) =
// feature 64 kb + 16 cg + j is positive), at byte l * TC_MASK_BYTES of the tile's mask blocks: the dX pass then reads
// 4 KB instead of 64 KB of 16-bit activations per layer and tile.
#define TC_MASK_BYTES 4096

static inline int star_make_tc_layout(const StarNetDesc* d, TcLayout* o) {
  if (d->n_blocks < 1 || 2 * d->n_blocks + 4 > STAR_MAX_LAYERS) return STAR_E_UNSUPPORTED;
  if (d->L_xyz != 10 || d->L_dir != 4) return STAR_E_UNSUPPORTED;   // tensor-core tier: reference configs only
  o->n_blocks = d->n_blocks;
  int n = 0, fo = 0, stages = 0, sb = 0, gb = 0;
  uint32_t wo = 0, wto = 0;
  auto add = [&](int kind, int nkb, int N, int region) {
    TcLayer& l = o->L[n++];
    l.kind = kind; l.nkb = nkb; l.N = N; l.region = region;
    l.bias_off = fo; fo += STAR_W;
    l.w_off = wo; wo += (uint32_t)nkb * (uint32_t)N * 128u;
    stages += nkb;
    l.s_in = sb; sb += nkb;                 // input operand blocks (written by the previous epilogue / the encoder)
    l.g_out = gb; gb += N / 64;
    // dX GEMM of this layer: out[m][j] = sum_n G[m][n] W[n][j]:  B operand rows j (K_in padded to 64 nkb... the
    // first min(nkb,4)*64 input features), K = n in N/64 blocks
    l.wt_off = wto; wto += (uint32_t)(N / 64) * (uint32_t)((nkb < 4 ? nkb : 4) * 64) * 128u;
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
  for (int i = 0; i + 1 < n; ++i) o->L[i].s_out = o->L[i + 1].s_in;   // epilogue output = next layer's input
  o->L[n - 1].s_out = sb; sb += 2;                                     // relu(h2), 128 columns
  o->mask_blk0 = sb;
  sb += (n * TC_MASK_BYTES + TC_BLOCK_BYTES - 1) / TC_BLOCK_BYTES;
  o->stash_blocks = sb;
  o->gstash_blocks = gb;
  o->wt_dirs_off = wto; wto += 2u * 64u * 128u;
  o->tstream_bytes = wto;
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
