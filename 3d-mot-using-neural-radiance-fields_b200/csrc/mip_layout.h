// Host+device description of one mip-NeRF field (models/mipnerf.py:53-100 -> nerfstudio NeRFField with
// use_integrated_encoding): integrated positional encoding (3 x 24 frequencies x {sin, sin(.+pi/2)} + raw = 147,
// padded to 160) -> mlp_base: 8 x 256 ReLU, the encoding re-concatenated IN FRONT of layer 4's input (K = 147 + 256)
// -> density head 256 -> 1;  mlp_head on cat[encoded_dir (27, padded to 32), base_out (256)]: 2 x 128 ReLU -> rgb
// head 128 -> 3.  708 352-MAC vanilla net vs 587 264 MAC here.
#pragma once
#include <stdint.h>
#include "../../include/star_b200.h"

#define MIP_KX 147
#define MIP_KXP 160
#define MIP_KD 27
#define MIP_KDP 32
#define MIP_W 256
#define MIP_WH 128
#define MIP_NF 24
#define MIP_NFD 4
#define MIP_NBASE 8
#define MIP_SKIP 4
#define MIP_FREQ_FLOATS 64   // freq table from the host: f[24] | f^2[24] | f_dir[4] | pad

// indices into m_w / m_b
#define MIP_L_DENS 8
#define MIP_L_H0 9
#define MIP_L_H1 10
#define MIP_L_RGB 11

struct MipLayout {
  // flat fp32 master vector (each nn.Linear as weight [out,in] row-major then bias [out]), in the order
  //   field.mlp_base.layers.0..7, field.field_output_density.net, field.mlp_head.layers.0, .1, field.field_heads.0.net
  int64_t m_w[12], m_b[12];
  int K[12], N[12];
  int64_t n_master;
  // fp32 packed image
  int64_t p_wt[MIP_NBASE];   // base layer l: W^T [Kp][256]; Kp = 160 (l = 0), 160 + 256 (l = 4: [enc | x]), 256 otherwise
  int64_t p_wb[MIP_NBASE];   // base layer l >= 1: the x part of W, [256][256]
  int64_t p_wbe[2];          // the encoding part of W zero-padded to [256][256]: [0] layer 0, [1] layer 4
  int64_t p_b[MIP_NBASE];
  int64_t p_h0t, p_h0bx, p_h0bd, p_h0b;   // head 0: W^T [32 + 256][128] (dirs first), x part [128][256], dir part [128][32], bias
  int64_t p_h1t, p_h1w, p_h1b;            // head 1: W^T [128][128], W [128][128], bias
  int64_t p_dw, p_db, p_rw, p_rb;         // density head w[256], b[1 (+3 pad)]; rgb head w[3][128], b[3 (+1 pad)]
  int64_t n_packed;
  // activation stash (training), each block [M][cols] fp32; offsets in floats-per-sample
  int64_t s_enc, s_dir, s_in[MIP_NBASE], s_base, s_h0, s_h1, stash_cols;
  // gradient stash dL/d(output of a GEMM layer)
  int64_t g_base[MIP_NBASE], g_h0, g_h1, g_cols;
};

static inline void star_make_mip_layout(MipLayout* o) {
  int64_t m = 0, p = 0, s = 0, g = 0;
  auto lin = [&](int idx, int K, int N) {
    o->K[idx] = K; o->N[idx] = N;
    o->m_w[idx] = m; m += (int64_t)N * K;
    o->m_b[idx] = m; m += N;
  };
  for (int l = 0; l < MIP_NBASE; ++l) lin(l, l == 0 ? MIP_KX : (l == MIP_SKIP ? MIP_KX + MIP_W : MIP_W), MIP_W);
  lin(MIP_L_DENS, MIP_W, 1);
  lin(MIP_L_H0, MIP_KD + MIP_W, MIP_WH);
  lin(MIP_L_H1, MIP_WH, MIP_WH);
  lin(MIP_L_RGB, MIP_WH, 3);
  o->n_master = m;
  for (int l = 0; l < MIP_NBASE; ++l) {
    const int Kp = l == 0 ? MIP_KXP : (l == MIP_SKIP ? MIP_KXP + MIP_W : MIP_W);
    o->p_wt[l] = p; p += (int64_t)Kp * MIP_W;
    o->p_wb[l] = p; if (l >= 1) p += (int64_t)MIP_W * MIP_W;
    o->p_b[l] = p; p += MIP_W;
  }
  o->p_wbe[0] = p; p += (int64_t)MIP_W * MIP_W;
  o->p_wbe[1] = p; p += (int64_t)MIP_W * MIP_W;
  o->p_h0t = p; p += (int64_t)(MIP_KDP + MIP_W) * MIP_WH;
  o->p_h0bx = p; p += (int64_t)MIP_WH * MIP_W;
  o->p_h0bd = p; p += (int64_t)MIP_WH * MIP_KDP;
  o->p_h0b = p; p += MIP_WH;
  o->p_h1t = p; p += (int64_t)MIP_WH * MIP_WH;
  o->p_h1w = p; p += (int64_t)MIP_WH * MIP_WH;
  o->p_h1b = p; p += MIP_WH;
  o->p_dw = p; p += MIP_W;
  o->p_db = p; p += 4;
  o->p_rw = p; p += 3 * MIP_WH;
  o->p_rb = p; p += 4;
  o->n_packed = p;
  o->s_enc = s; s += MIP_KXP;
  o->s_dir = s; s += MIP_KDP;
  o->s_in[0] = o->s_enc;
  for (int l = 1; l < MIP_NBASE; ++l) { o->s_in[l] = s; s += MIP_W; }
  o->s_base = s; s += MIP_W;
  o->s_h0 = s; s += MIP_WH;
  o->s_h1 = s; s += MIP_WH;
  o->stash_cols = s;
  for (int l = 0; l < MIP_NBASE; ++l) { o->g_base[l] = g; g += MIP_W; }
  o->g_h0 = g; g += MIP_WH;
  o->g_h1 = g; g += MIP_WH;
  o->g_cols = g;
}
