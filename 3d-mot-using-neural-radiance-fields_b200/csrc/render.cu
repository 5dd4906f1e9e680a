// star_render_forward: the whole coarse -> fine render (models/rendering__.py:115-149, :249-298; models/star__.py:119-225)
// as ONE C-ABI call that queues every kernel on the caller's stream: ray generation, depth sampling, the V + 1 coarse
// MLPs (positions formed in-kernel, object raws written in place into the [R,V,S] layout), compositing, inverse-CDF
// sampling + merge, the V + 1 fine MLPs, compositing.  No host synchronisation, no allocation: scratch comes from the
// caller's workspace.
#include "star_common.cuh"
#include "mlp_layout.h"
#include "mlp_tc_layout.h"

int star_tc_forward(const TcLayout& tl, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                    const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, float* raw_alpha,
                    float* raw_rgb, int64_t ray_stride, void* stash, int* status, int fp16, cudaStream_t st);
int star_f32_forward(const MlpLayout& lay, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                     const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, float* raw_alpha,
                     float* raw_rgb, int64_t ray_stride, void* stash, cudaStream_t st);

namespace {

struct Scratch {        // float offsets into the workspace
  size_t ra_s, rc_s, ra_d, rc_d, z0, rays_o, rays_d, viewdirs, regws, total;
};

Scratch scratch_layout(const StarRenderCfg& c) {
  Scratch s;
  const size_t R = (size_t)c.R, Nf = (size_t)(c.Nc + (c.Ni > 0 ? c.Ni : 0)), V = (size_t)c.V;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t at = o; o += (n + 3) / 4 * 4; return at; };   // 16-byte aligned pieces
  s.ra_s = take(R * Nf);
  s.rc_s = take(R * Nf * 3);
  s.ra_d = take(R * V * Nf);
  s.rc_d = take(R * V * Nf * 3);
  s.z0 = take(R * (size_t)c.Nc);
  s.rays_o = take(R * 3);
  s.rays_d = take(R * 3);
  s.viewdirs = take(R * 3);
  s.regws = take(star_composite_multi_ws_bytes(c.R > 0 ? c.R : 1) / sizeof(float));
  s.total = o;
  return s;
}

__global__ void normalize_dirs_kernel(const float* __restrict__ d, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = d[i * 3 + 0], y = d[i * 3 + 1], z = d[i * 3 + 2];
    // torch.norm(rays_d, dim=-1) then a division per component (train_app_init__.py:56)
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    out[i * 3 + 0] = __fdiv_rn(x, nrm);
    out[i * 3 + 1] = __fdiv_rn(y, nrm);
    out[i * 3 + 2] = __fdiv_rn(z, nrm);
  }
}

int run_net(const StarRenderCfg& c, int n_blocks, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
            const float* pose12, const float* sc_xyz, const float* sc_dir, int S, float* raw_alpha, float* raw_rgb,
            int64_t ray_stride, int* status, cudaStream_t st) {
  StarNetDesc d{n_blocks, c.L_xyz, c.L_dir, c.precision};
  MlpLayout lay;
  int rc = star_make_layout(&d, &lay);
  if (rc) return rc;
  const int prec = c.precision & 0xff;
  if (prec == STAR_PREC_F32)
    return star_f32_forward(lay, packed, pts, viewdirs, pose12, sc_xyz, sc_dir, c.R, S, raw_alpha, raw_rgb, ray_stride,
                            nullptr, st);
  if (prec == STAR_PREC_BF16 || prec == STAR_PREC_F16) {
    if (c.precision & STAR_PREC_FLAG_RETIRED) return STAR_E_UNSUPPORTED;
    TcLayout tl;
    rc = star_make_tc_layout(&d, &tl);
    if (rc) return rc;
    return star_tc_forward(tl, packed, pts, viewdirs, pose12, sc_xyz, sc_dir, c.R, S, raw_alpha, raw_rgb, ray_stride,
                           nullptr, status, (prec == STAR_PREC_F16 ? 1 : 0) | ((c.precision & STAR_PREC_FLAG_NO_WSHARE) ? 2 : 0), st);
  }
  return STAR_E_UNSUPPORTED;
}

// one pass (coarse or fine): V + 1 field MLPs, then compositing
int run_pass(const StarRenderCfg& c, const StarRenderIn& in, bool coarse, const StarPtsSrc& pts, const float* viewdirs,
             const float* rays_d, const float* z, int S, float* ws, const Scratch& sc, const StarMultiOut& o,
             float* dists, int* status, cudaStream_t st) {
  const void* ps = coarse ? in.packed_static_coarse : in.packed_static_fine;
  int rc = run_net(c, c.n_blocks_static, ps, pts, viewdirs, nullptr, nullptr, nullptr, S, ws + sc.ra_s, ws + sc.rc_s, S,
                   status, st);                                                        // star__.py:144 (never BARF)
  if (rc) return rc;
  if (c.V == 0)
    return star_composite_single_forward(ws + sc.ra_s, ws + sc.rc_s, z, rays_d, c.R, S, c.far_dist, c.white_bkgd, o.rgb,
                                         o.disp, o.acc, o.depth, o.weights, dists, st);
  const void* const* pd = coarse ? in.packed_dynamic_coarse : in.packed_dynamic_fine;
  for (int v = 0; v < c.V; ++v) {                                                       // star__.py:201-210
    rc = run_net(c, c.n_blocks_dynamic, pd[v], pts, viewdirs, in.pose12 + 12 * v, in.enc_scale_xyz, in.enc_scale_dir, S,
                 ws + sc.ra_d + (size_t)v * S, ws + sc.rc_d + (size_t)v * S * 3, (int64_t)c.V * S, status, st);
    if (rc) return rc;
  }
  return star_composite_multi_forward(ws + sc.ra_s, ws + sc.rc_s, ws + sc.ra_d, ws + sc.rc_d, z, rays_d, c.R, c.V, S,
                                      c.far_dist, c.white_bkgd, c.chunk, &o, ws + sc.regws, st);
}

}  // namespace

extern "C" size_t star_render_workspace_bytes(const StarRenderCfg* cfg) {
  if (!cfg || cfg->R < 0 || cfg->Nc < 1 || cfg->V < 0 || cfg->V > STAR_MAX_V) return 0;
  return scratch_layout(*cfg).total * sizeof(float) + 256;
}

extern "C" int star_render_forward(const StarRenderCfg* cfg, const StarRenderIn* in, const StarRenderOut* out,
                                   void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
  if (!cfg) return STAR_E_NULL;
  if (cfg->R == 0) return STAR_OK;     /* an empty batch carries no pointers (and needs no workspace) */
  if (!in || !out || !workspace) return STAR_E_NULL;
  const StarRenderCfg& c = *cfg;
  if (c.R < 0 || c.Nc < 1 || c.V < 0 || c.V > STAR_MAX_V || c.chunk < 1) return STAR_E_BAD_SHAPE;
  if (c.Ni > 0 && c.Nc < 3) return STAR_E_BAD_SHAPE;
  if (!in->packed_static_coarse || (c.Ni > 0 && !in->packed_static_fine)) return STAR_E_NULL;
  if (c.V > 0 && (!in->pose12 || !in->packed_dynamic_coarse || (c.Ni > 0 && !in->packed_dynamic_fine))) return STAR_E_NULL;
  if (c.Ni > 0 && !in->u && !in->u_det && !in->z_samples) return STAR_E_NULL;
  if (c.Ni > 0 && (!out->z_vals || !out->z_std)) return STAR_E_NULL;
  if (workspace_bytes < star_render_workspace_bytes(cfg)) return STAR_E_WORKSPACE;
  if (((uintptr_t)workspace & 15) != 0) return STAR_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const Scratch sc = scratch_layout(c);
  int rc;

  // ---- rays
  const float* rays_o = in->rays_o;
  const float* rays_d = in->rays_d;
  const float* viewdirs = in->viewdirs;
  if (rays_o == nullptr) {
    if (!in->c2w || in->nrows * in->W != c.R) return STAR_E_BAD_SHAPE;
    float* ro = out->rays_o ? out->rays_o : ws + sc.rays_o;
    float* rd = out->rays_d ? out->rays_d : ws + sc.rays_d;
    float* vd = out->viewdirs ? out->viewdirs : ws + sc.viewdirs;
    rc = star_get_rays(in->H, in->W, in->fx, in->fy, in->cx, in->cy, in->c2w, in->row0, in->nrows, ro, rd, vd, st);
    if (rc) return rc;
    rays_o = ro; rays_d = rd; viewdirs = vd;
  } else {
    if (!rays_d) return STAR_E_NULL;
    if (viewdirs == nullptr) {
      float* vd = out->viewdirs ? out->viewdirs : ws + sc.viewdirs;
      int blocks = (c.R + 255) / 256;
      if (blocks > 148 * 8) blocks = 148 * 8;
      normalize_dirs_kernel<<<blocks, 256, 0, st>>>(rays_d, c.R, vd);
      rc = star_check_launch();
      if (rc) return rc;
      viewdirs = vd;
    }
  }

  // ---- coarse depths
  const float* z0 = in->z_vals;
  const float* pts0 = in->z_vals ? in->pts : nullptr;
  if (z0 == nullptr) {
    if (!in->t_vals) return STAR_E_NULL;
    float* zo = out->z_vals0 ? out->z_vals0 : ws + sc.z0;
    rc = star_sample_pts(nullptr, nullptr, in->t_vals, in->t_rand, c.near_, c.far_, c.R, c.Nc, c.lindisp, nullptr, zo, st);
    if (rc) return rc;
    z0 = zo;
  }

  // ---- coarse pass
  const StarPtsSrc src0{pts0, rays_o, rays_d, z0};
  // single field, fine samples drawn here: the coarse compositing and the hierarchical step are ONE kernel per ray
  // (ray_fused.cu; weights and depths stay in shared memory between the two).  Measured (profiles/r2j_hbm_kernels.txt):
  // 64 + 128 samples 0.336 ms against 0.421 ms for the two kernels (160 000 rays); 256 + 256: 0.285 against 0.275 ms (the
  // larger per-warp shared memory costs occupancy) -- so only up to 128 coarse samples.
  bool fused_tail = c.V == 0 && c.Ni > 0 && in->z_samples == nullptr && (c.Nc & 1) == 0 && c.Nc >= 4 && c.Nc <= 128 &&
                    out->z_samples != nullptr;
  if (fused_tail) {
    rc = run_net(c, c.n_blocks_static, in->packed_static_coarse, src0, viewdirs, nullptr, nullptr, nullptr, c.Nc,
                 ws + sc.ra_s, ws + sc.rc_s, c.Nc, status, st);
    if (rc) return rc;
    const StarMultiOut& o = out->coarse;
    rc = star_composite_hier_forward(ws + sc.ra_s, ws + sc.rc_s, z0, rays_d, in->u, in->u_det, c.R, c.Nc, c.Ni, c.far_dist,
                                     c.white_bkgd, o.rgb, o.disp, o.acc, o.depth, o.weights, out->dists0, out->z_samples,
                                     out->z_vals, out->z_std, st);
    if (rc == STAR_E_UNSUPPORTED) {        // misaligned caller arrays: the two stand-alone kernels
      fused_tail = false;
      rc = star_composite_single_forward(ws + sc.ra_s, ws + sc.rc_s, z0, rays_d, c.R, c.Nc, c.far_dist, c.white_bkgd, o.rgb,
                                         o.disp, o.acc, o.depth, o.weights, out->dists0, st);
    }
    if (rc) return rc;
  } else {
    rc = run_pass(c, *in, true, src0, viewdirs, rays_d, z0, c.Nc, ws, sc, out->coarse, out->dists0, status, st);
    if (rc || c.Ni <= 0) return rc;
  }

  // ---- hierarchical step: z_mid, sample_pdf(weights[1:-1]), sort(cat), std   (:128-144 / :271-296)
  if (fused_tail) {
    rc = STAR_OK;
  } else if (in->z_samples != nullptr) {
    rc = star_merge_samples(z0, in->z_samples, nullptr, nullptr, c.R, c.Nc, c.Ni, out->z_vals, out->z_std, nullptr, st);
  } else {
    if (!out->z_samples) return STAR_E_NULL;
    rc = star_hierarchical(z0, out->coarse.weights, in->u, in->u_det, nullptr, nullptr, c.R, c.Nc, c.Ni, out->z_samples,
                           out->z_vals, out->z_std, nullptr, st);
  }
  if (rc) return rc;

  // ---- fine pass on all Nc + Ni samples with the fine nets
  const StarPtsSrc src1{nullptr, rays_o, rays_d, out->z_vals};
  return run_pass(c, *in, false, src1, viewdirs, rays_d, out->z_vals, c.Nc + c.Ni, ws, sc, out->fine, out->dists, status,
                  st);
}
