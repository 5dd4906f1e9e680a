// Host+device description of one NeRF MLP as a chain of GEMM layers (models/nerf.py:66-97,150-163;
// models/resnet.py:51-59,103-110) and of the buffers derived from it.
#pragma once
#include <stdint.h>
#include "../../include/star_b200.h"

#define STAR_W 256         // netwidth (all reference configs)
#define STAR_WV 128        // view-branch width W/2
#define STAR_MAX_LAYERS 16

enum LayerKind { LK_IN = 0, LK_FC0 = 1, LK_FC1 = 2, LK_OUT = 3, LK_FEAT = 4, LK_VIEWS = 5 };

struct MlpLayer {
  int kind;
  int K, Kpad, N;          // in features (true / padded), out features
  int64_t m_w, m_b;        // offsets (floats) of weight [N,K] and bias [N] in the flat master vector
  int64_t p_wt, p_wb, p_b; // fp32 packed image: W^T [Kpad][N], W padded [N][Kpad], bias [N]
  int64_t s_in;            // stash: per-sample column offset of this layer's INPUT (width Kpad)
  int64_t g_out;           // gstash: per-sample column offset of dL/d(output) (width N)
};

struct MlpLayout {
  int n_blocks, n_layers, L_xyz, L_dir, in_xyz, in_dir;
  MlpLayer L[STAR_MAX_LAYERS];
  int64_t m_alpha_w, m_alpha_b, m_rgb_w, m_rgb_b;  // master offsets of the two heads
  int64_t p_alpha_w, p_alpha_b, p_rgb_w, p_rgb_b;  // packed offsets
  int64_t s_h2;                                   // stash column offset of h2 [128]
  int64_t n_master, n_packed;                     // float counts
  int64_t stash_cols, g_cols;                     // floats per sample
};

static inline int star_make_layout(const StarNetDesc* d, MlpLayout* o) {
  if (d->n_blocks < 1 || 2 * d->n_blocks + 4 > STAR_MAX_LAYERS) return STAR_E_UNSUPPORTED;
  if (d->L_xyz < 1 || d->L_xyz > 10 || d->L_dir < 1 || d->L_dir > 4) return STAR_E_UNSUPPORTED;
  o->n_blocks = d->n_blocks;
  o->L_xyz = d->L_xyz;
  o->L_dir = d->L_dir;
  o->in_xyz = 3 + 6 * d->L_xyz;
  o->in_dir = 3 + 6 * d->L_dir;
  int n = 0;
  int64_t m = 0, p = 0, s = 0, g = 0;
  auto add = [&](int kind, int K, int Kpad, int N) {
    MlpLayer& l = o->L[n++];
    l.kind = kind; l.K = K; l.Kpad = Kpad; l.N = N;
    l.m_w = m; m += (int64_t)N * K;
    l.m_b = m; m += N;
    l.p_wt = p; p += (int64_t)Kpad * N;
    l.p_wb = p; p += (int64_t)N * Kpad;
    l.p_b = p; p += N;
    l.s_in = s; s += Kpad;
    l.g_out = g; g += N;
  };
  add(LK_IN, o->in_xyz, 64, STAR_W);
  for (int b = 0; b < d->n_blocks; ++b) {
    add(LK_FC0, STAR_W, STAR_W, STAR_W);
    add(LK_FC1, STAR_W, STAR_W, STAR_W);
  }
  add(LK_OUT, STAR_W, STAR_W, STAR_W);
  // master order: ..., lin_out, alpha_linear, feature_linear, views_linears.0, rgb_linear
  o->m_alpha_w = m; m += STAR_W;
  o->m_alpha_b = m; m += 1;
  add(LK_FEAT, STAR_W, STAR_W, STAR_W);
  add(LK_VIEWS, STAR_W + o->in_dir, STAR_W + 32, STAR_WV);
  o->m_rgb_w = m; m += 3 * STAR_WV;
  o->m_rgb_b = m; m += 3;
  o->p_alpha_w = p; p += STAR_W;
  o->p_alpha_b = p; p += 4;
  o->p_rgb_w = p; p += 3 * STAR_WV;
  o->p_rgb_b = p; p += 4;
  o->s_h2 = s; s += STAR_WV;
  o->n_layers = n;
  o->n_master = m;
  o->n_packed = p;
  o->stash_cols = s;
  o->g_cols = g;
  return STAR_OK;
}
