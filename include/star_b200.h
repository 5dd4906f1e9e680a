/*
 * star_b200.h -- C ABI of the B200-native (sm_100a) STaR / NeRF render hot path.
 *
 * The reference (burakcuhadar/3D-MOT-using-Neural-Radiance-Fields) is pure Python/PyTorch and has
 * no FFI of its own; each entry point below replaces the reference Python function cited next to
 * it (paths relative to the reference root).  The host side that binds this header is
 * 3d-mot-using-neural-radiance-fields_b200/_capi.py (ctypes); INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to densely packed row-major data unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on that stream,
 *     allocate nothing, and keep no global mutable state (re-entrant across streams / devices);
 *   - workspaces are supplied by the caller (sizes from the *_bytes queries);
 *   - return value: 0 = OK, otherwise a STAR_E_* code (star_error_string() gives the text).
 *     Nothing throws or aborts.  There is no CPU fallback.
 */
#ifndef STAR_B200_H_
#define STAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAR_ABI_VERSION 2

enum {
  STAR_OK = 0,
  STAR_E_BAD_SHAPE = 1,    /* a size is out of the supported range                       */
  STAR_E_UNSUPPORTED = 2,  /* W / depth / L / precision combination not compiled          */
  STAR_E_NULL = 3,         /* a required pointer is NULL                                  */
  STAR_E_ALIGN = 4,        /* a pointer violates the documented alignment                 */
  STAR_E_WORKSPACE = 5,    /* workspace too small                                         */
  STAR_E_CUDA = 6          /* a CUDA runtime call / launch failed (see star_last_cuda_error) */
};

/* precision tiers of the MLP (north star: 1e-4 abs for fp32, 2e-3 abs for the 16-bit tensor-core MLP).
 *   F32  : CUDA-core fp32 kernels (forward, backward).
 *   BF16 : tcgen05 tensor cores, bf16 operands, fp32 accumulation in TMEM (forward and backward).  Random-init nets
 *          deviate by up to 9e-3 on composited rgb (every one of the 11 chained GEMMs contributes
 *          2-3e-3 on its own), i.e. this tier does NOT meet the 2e-3 bound in the max norm.
 *   F16  : the same kernels with IEEE fp16 activations and weights (2^-12 operand rounding instead of 2^-9, same
 *          speed): max error < 1e-3, the tier that meets the bound and the default of the tensor-core path.  The
 *          back-propagated gradients are fp16 too, scaled per launch by a power of two taken from the device-side
 *          maximum of the incoming gradient and un-scaled in the dW / pose epilogues.  Range
 *          guard: activations beyond 65504 overflow to inf and surface as non-finite raw outputs, which
 *          star_mlp_forward reports through its `status` word. */
enum { STAR_PREC_F32 = 0, STAR_PREC_BF16 = 1, STAR_PREC_F16 = 2 };
/* Bits 0x100 (CTA-pair / cta_group::2 forward) and 0x200 (activation stash written by the epilogue threads) selected two
 * A/B variants of the forward kernels that were built, verified bit-identical and measured SLOWER in round 2 (C2 render 2.05 M
 * against 2.66 M rays/s; stash forward 3.04 against 2.19 ms per step -- DESIGN.md section 7.0, profiles/r2b_*).  The variants
 * were removed when the forward kernel was re-scheduled across tile boundaries; the bits are RETIRED: an entry point that
 * receives one returns STAR_E_UNSUPPORTED. */
#define STAR_PREC_FLAG_RETIRED 0x300
/* backward, A/B variant: the pipelined dX chain (every N = 256 GEMM as two N = 128 halves with their own completion barriers,
 * so that the epilogue of a group overlaps the MMAs of the next) instead of the serial one.  Same gradients; measured SLOWER
 * (dX + dW + heads 5.09 against 4.79 ms per 4096-ray step: twice as many stages and N = 128 MMAs make the single issuing thread
 * the bottleneck), so it is opt-in. */
#define STAR_PREC_FLAG_DX_PIPELINED 0x400
/* inference forward, opt-OUT of the weight-sharing cluster launch: by default (launches of at least two tiles per SM) the two
 * CTAs of a cluster of 2 each fetch half of every weight K-block from L2 and multicast it into both CTAs' shared-memory ring,
 * which halves the L2 -> SM weight traffic (1.6 MB per tile and SM); with this bit every CTA streams its own weights.  Same
 * arithmetic, bit-identical outputs (tests/test_gpu_bf16.py); measured 2.81 -> 2.88 M rays/s on the C2 render
 * (profiles/r2zz_ab_wshare.txt). */
#define STAR_PREC_FLAG_NO_WSHARE 0x800

/* One NeRF radiance MLP (models/nerf.py:34-110, models/resnet.py:62-110).  W is fixed at 256,
 * the view branch at 128 (all 15 reference configs agree). */
typedef struct StarNetDesc {
  int32_t n_blocks; /* ResnetFC blocks: netdepth/2 = 4 (static), netdepth/4 = 2 (dynamic)  */
  int32_t L_xyz;    /* multires        (10) -> 3 + 6*L_xyz = 63 input dims                 */
  int32_t L_dir;    /* multires_views  (4)  -> 27 dims                                     */
  int32_t precision; /* STAR_PREC_*                                                         */
} StarNetDesc;

int star_abi_version(void);
const char* star_error_string(int code);
int star_last_cuda_error(void);
/* Watchdog of the tensor-core kernels: every mbarrier wait is bounded (2 s of SM cycles); a wait that times out stores
 * (wait code << 16 | CTA index) in a word of mapped host memory and traps, which surfaces as a CUDA launch failure.  The
 * word survives the dead context: family 0 = MLP forward, 1 = dX chain, 2 = dW, 3 = pipelined dX, 4 = mip forward,
 * 5 = mip dX.  0 = that family never timed out. */
int star_watchdog_word(int family);

/* Number of fp32 elements of the flat master parameter vector of one net, in this order
 * (each nn.Linear as weight [out,in] row-major then bias [out]):
 *   pts_net.lin_in, {pts_net.blocks.b.fc_0, pts_net.blocks.b.fc_1} b<n_blocks, pts_net.lin_out,
 *   alpha_linear, feature_linear, views_linears.0, rgb_linear
 * = 711300 for n_blocks=4, 448132 for n_blocks=2 (reference state_dict). */
size_t star_net_param_count(const StarNetDesc* d);

/* Packed (kernel-ready) weight image derived from the flat master vector. */
size_t star_packed_bytes(const StarNetDesc* d);
int star_pack_weights(const StarNetDesc* d, const float* flat_master, void* packed, void* stream);

/* ---- a1: models/rendering__.py:75-112  sample_pts --------------------------------------------
 * t_vals[Nc] = torch.linspace(0,1,Nc) from the host (its symmetric formula is not i/(n-1)).
 * t_rand[R*Nc] or NULL: injected stratified jitter (perturb > 0 and is_train).
 * Outputs pts[R,Nc,3] (NULL: depths only -- the MLP kernels can form the positions themselves), z_vals[R,Nc].
 * Bit-exact vs the reference ops (no FMA contraction). */
int star_sample_pts(const float* rays_o, const float* rays_d, const float* t_vals, const float* t_rand,
                    float near_, float far_, int R, int Nc, int lindisp, float* pts, float* z_vals,
                    void* stream);

/* ---- a2: models/rendering__.py:41-55  get_rays ---------------------------------
 * Pinhole rays of the pixel rows [row0, row0 + nrows) of an H x W view: intrinsics (fx, fy, cx, cy) = (K[0][0], K[1][1],
 * K[0][2], K[1][2]), c2w = DEVICE pointer to the [3,4] camera-to-world matrix (row-major).  Outputs rays_o, rays_d
 * [nrows*W, 3] (rays_d un-normalised, bit-exact vs the reference ops) and, if viewdirs != NULL, rays_d / ||rays_d||.
 * Lets a full-view render start from 12 floats instead of 24 B/ray of host-generated rays. */
int star_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int row0, int nrows,
                  float* rays_o, float* rays_d, float* viewdirs, void* stream);

/* ---- a3: models/embedder.py:81-112  Embedder.forward -----------------------------------------
 * x[M,3] -> out[M,3+6L].  scale[3+6L] or NULL: per-element BARF mask (w[j mod L] quirk, :32). */
int star_embed(const float* x, int M, int L, const float* scale, float* out, void* stream);

/* ---- a4 (+K2 pose transform): models/nerf.py:112-179, models/star__.py:160-199 ---------------
 * pts[R,S,3], viewdirs[R,3]; pose12 = NULL (static net) or 12 floats row-major [R(3x3) | t] :
 * p' = R p + t, d' = R d (object frame).  enc_scale_xyz / enc_scale_dir: BARF masks or NULL.
 * raw_alpha [R*S] and raw_rgb [R*S*3] are written with element strides so that the multi-field
 * layout [R,V,S] of raw2outputs_star can be filled in place:
 *   raw_alpha[r*alpha_ray_stride + s], raw_rgb[(r*alpha_ray_stride + s)*3 + c].
 * stash (NULL for inference): per-GEMM input activations kept for the backward pass,
 * star_stash_bytes(d, R*S) bytes.
 * Sample positions: pts[R,S,3], or pts = NULL and (rays_o[R,3], rays_d[R,3], z_vals[R,S]): the kernel forms
 * p = rays_o + rays_d * z with the reference's two roundings (rendering__.py:108-110), bit-identical to a materialised
 * pts, so the fused render path never writes / reads 12 B per sample of positions.
 * status (NULL or a device-accessible int32, e.g. mapped pinned host memory): set to 1 when a raw output is not
 * finite (range guard of the fp16 tier; never cleared by the library). */
size_t star_stash_bytes(const StarNetDesc* d, int64_t n_samples);
int star_mlp_forward(const StarNetDesc* d, const void* packed, const float* pts, const float* rays_o,
                     const float* rays_d, const float* z_vals, const float* viewdirs,
                     const float* pose12, const float* enc_scale_xyz, const float* enc_scale_dir,
                     int R, int S, float* raw_alpha, float* raw_rgb, int64_t alpha_ray_stride,
                     void* stash, int32_t* status, void* stream);

/* Backward of star_mlp_forward.  d_raw_alpha / d_raw_rgb use the same strides as the forward.
 * grad_flat: fp32 [star_net_param_count] in the flat master order, ACCUMULATED into (caller zeroes).
 * pose_acc: 32 floats ACCUMULATED into (NULL when pose12 is NULL):
 *   [0:3]  sum g            [3:12]  sum g p^T (row-major)   [12:15] sum p' x g
 *   [15:24] sum h d^T       [24:27] sum d' x h              (g = dL/dp', h = dL/dd')
 * from which the host forms dL/dM (4x4 pose, star__.py:160-180) or the pypose left-tangent
 * gradient [sum g, sum p' x g + sum d' x h, 0] (7-vector pose, star__.py:182-199).
 * workspace: star_mlp_backward_workspace_bytes(d, R*S) bytes. */
size_t star_mlp_backward_workspace_bytes(const StarNetDesc* d, int64_t n_samples);
int star_mlp_backward(const StarNetDesc* d, const void* packed, const float* flat_master, const float* pts,
                      const float* rays_o, const float* rays_d, const float* z_vals,
                      const float* viewdirs, const float* pose12, const float* enc_scale_xyz,
                      const float* enc_scale_dir, int R, int S, const float* d_raw_alpha,
                      const float* d_raw_rgb, int64_t alpha_ray_stride, const void* stash, void* workspace,
                      float* grad_flat, float* pose_acc, void* stream);

/* ---- a5/a6: models/rendering__.py:301-379  raw2outputs ---------------------------------------
 * Outputs: rgb[R,3], disp[R], acc[R], depth[R], weights[R,S], dists[R,S]. */
int star_composite_single_forward(const float* raw_alpha, const float* raw_rgb, const float* z_vals,
                                  const float* rays_d, int R, int S, float far_dist, int white_bkgd,
                                  float* rgb, float* disp, float* acc, float* depth, float* weights,
                                  float* dists, void* stream);
/* Gradients of (rgb, disp, acc, depth, weights) -> (raw_alpha, raw_rgb); any g_* may be NULL. */
int star_composite_single_backward(const float* raw_alpha, const float* raw_rgb, const float* z_vals,
                                   const float* rays_d, int R, int S, float far_dist, int white_bkgd,
                                   const float* g_rgb, const float* g_disp, const float* g_acc,
                                   const float* g_depth, const float* g_weights, float* d_raw_alpha,
                                   float* d_raw_rgb, void* stream);

/* ---- a7/a8: models/rendering__.py:383-576 raw2outputs_star + :612-715 regularisers ------------
 * raw_alpha_d [R,V,S], raw_rgb_d [R,V,S,3] (reference layout).  `chunk` reproduces
 * STaR.forward's "mean within a ray chunk, summed over chunks" (star__.py:84-112) for the five
 * scalar regularisers.  reg_partial: workspace of 5*gridDim floats (star_composite_multi_ws_bytes);
 * regs[5] = alpha_entropy, dynamic_vs_static, ray_reg, static_reg, dynamic_reg.
 * rgb_dynamic_all may be NULL (train mode: "test" is False, star__.py:224). */
typedef struct StarMultiOut {
  float* rgb;                   /* [R,3]   */
  float* disp;                  /* [R]     */
  float* acc;                   /* [R]     */
  float* depth;                 /* [R]     */
  float* weights;               /* [R,S]   */
  float* rgb_static;            /* [R,3]   */
  float* depth_static;          /* [R]     */
  float* rgb_dynamic;           /* [R,V,3] */
  float* depth_dynamic;         /* [R,V]   */
  float* dynamic_transmittance; /* [R,V]   */
  float* rgb_dynamic_all;       /* [R,3] or NULL */
  float* regs;                  /* [5]     */
} StarMultiOut;
size_t star_composite_multi_ws_bytes(int R);
int star_composite_multi_forward(const float* raw_alpha_s, const float* raw_rgb_s, const float* raw_alpha_d,
                                 const float* raw_rgb_d, const float* z_vals, const float* rays_d, int R,
                                 int V, int S, float far_dist, int white_bkgd, int chunk,
                                 const StarMultiOut* out, void* workspace, void* stream);
/* g_regs: DEVICE pointer to 5 floats (upstream grads of the scalars) or NULL. */
int star_composite_multi_backward(const float* raw_alpha_s, const float* raw_rgb_s, const float* raw_alpha_d,
                                  const float* raw_rgb_d, const float* z_vals, const float* rays_d, int R,
                                  int V, int S, float far_dist, int white_bkgd, int chunk,
                                  const float* g_rgb, const float* g_disp, const float* g_acc,
                                  const float* g_depth, const float* g_weights, const float* g_regs,
                                  float* d_raw_alpha_s, float* d_raw_rgb_s, float* d_raw_alpha_d,
                                  float* d_raw_rgb_d, void* stream);

/* ---- a9: models/rendering__.py:719-761  sample_pdf -------------------------------------------
 * bins[R,nb] (row stride bins_stride), weights[R,nb-1] (row stride w_stride; lets the caller pass
 * the weights[...,1:-1] view without a copy).  u[R,Ni] or NULL (det: u = u_det[Ni], a host
 * torch.linspace(0,1,Ni) copied to the device).  Outputs samples[R,Ni]; optional int64
 * inds/below/above[R,Ni] and cdf[R,nb] (NULL to skip) for the bit-exactness tests.
 * Arithmetic (defined): normaliser = exactly rounded fp32 sum (fp64 accumulate),
 * cdf = fp64 prefix sum rounded per element (== torch CPU cumsum), no FMA contraction. */
int star_sample_pdf(const float* bins, int64_t bins_stride, const float* weights, int64_t w_stride,
                    const float* u, const float* u_det, int R, int nb, int Ni, float* samples,
                    int64_t* inds, int64_t* below, int64_t* above, float* cdf, void* stream);
/* cdf supplied by the caller (kernel-level parity: identical cdf,u -> identical inds, bit-exact). */
int star_invert_cdf(const float* bins, const float* cdf, const float* u, int R, int nb, int Ni,
                    float* samples, int64_t* inds, int64_t* below, int64_t* above, void* stream);

/* ---- a10: models/rendering__.py:128-144 / :271-296  hierarchical step -------------------------
 * z_vals[R,Nc], weights[R,Nc] -> z_mid, sample_pdf(z_mid, weights[...,1:-1]) -> z_samples[R,Ni],
 * z_all[R,Nc+Ni] = sort(cat(z_vals, z_samples)), z_std[R] = std(z_samples, unbiased=False),
 * pts_fine[R,Nc+Ni,3] = o + d*z_all (NULL to skip). */
int star_hierarchical(const float* z_vals, const float* weights, const float* u, const float* u_det,
                      const float* rays_o, const float* rays_d, int R, int Nc, int Ni, float* z_samples,
                      float* z_all, float* z_std, float* pts_fine, void* stream);

/* The merge half of star_hierarchical on caller-supplied fine samples z_samples[R,Ni] (the
 * reference's `sort(cat([z_vals, z_samples]))`, std and `pts = o + d*z`, :136-144): used when the
 * caller injects the sample positions (stage-wise parity of the fine pass; sample_pdf is
 * ill-conditioned where the coarse pdf is small). */
int star_merge_samples(const float* z_vals, const float* z_samples, const float* rays_o, const float* rays_d,
                       int R, int Nc, int Ni, float* z_all, float* z_std, float* pts_fine, void* stream);

/* ---- a6 + a9 + a10 fused (north-star bullet 1): models/rendering__.py:307-379 (raw2outputs of the coarse pass),
 *      :719-761 (sample_pdf), :128-144 (z_mid, sort(cat), std) ---------------------------------------------------------
 * One kernel, one warp per ray: compositing of the Nc coarse samples, the CDF of weights[1:-1], the inverse-CDF draw,
 * z_std and the merge; weights and depths of the ray pass from the compositing part to the sampling part in shared
 * memory.  Outputs = those of star_composite_single_forward (rgb[R,3], disp, acc, depth [R], weights, dists [R,Nc];
 * the last two may be NULL) + those of star_hierarchical without pts (z_samples[R,Ni], z_all[R,Nc+Ni], z_std[R]), bit
 * for bit.  Nc must be even and the [R,Nc] arrays 8-byte aligned: otherwise STAR_E_UNSUPPORTED (use the two entries). */
int star_composite_hier_forward(const float* raw_alpha, const float* raw_rgb, const float* z_vals, const float* rays_d,
                                const float* u, const float* u_det, int R, int Nc, int Ni, float far_dist,
                                int white_bkgd, float* rgb, float* disp, float* acc, float* depth, float* weights,
                                float* dists, float* z_samples, float* z_all, float* z_std, void* stream);

/* ---- a10 + a11 in ONE call: models/rendering__.py:115-149 (render_star_appinit), :249-298 (render_star_online),
 *      models/star__.py:119-225 (STaR.forward_chunk) -----------------------------------------------------------
 * star_render_forward runs the whole coarse -> fine render of R rays on `stream` with no host round trip:
 *   [ray generation from a pinhole camera (get_rays, :41-55)] -> [stratified depths (sample_pts, :87-106)] ->
 *   coarse MLPs of the static field and of the V object fields (positions formed in-kernel from rays + depths, or read
 *   from a caller-supplied pts; samples moved into each object frame by pose12[v]; raw outputs written in place into the
 *   [R,V,S] layout) -> compositing + regularisers -> inverse-CDF sampling, merge, z_std -> fine MLPs on all Nc + Ni
 *   samples -> compositing.  2 V + 6 (+2) launches for the vanilla path, e.g. 18 for V = 5.
 * Inference only (no activations are kept); training goes through the per-stage entries and their backward twins.
 * Pointers are device pointers unless noted; any output pointer may be NULL except rgb / disp / acc / depth / weights.
 * With V == 0 only the first five StarMultiOut fields are used (+ dists). */
typedef struct StarRenderCfg {
  int32_t R, Nc, Ni, V;
  int32_t precision;                 /* STAR_PREC_* (| STAR_PREC_FLAG_*) of every field MLP                         */
  int32_t n_blocks_static, n_blocks_dynamic, L_xyz, L_dir;
  int32_t white_bkgd, lindisp, test; /* test: eval extras of raw2outputs_star (rgb_dynamic_all), star__.py:224      */
  int32_t chunk;                     /* ray-chunk length of the regulariser means (star__.py:84-112)                 */
  float near_, far_, far_dist;
} StarRenderCfg;

typedef struct StarRenderIn {
  /* rays: explicit [R,3] arrays, or rays_o == NULL and a camera: rays of pixel rows [row0, row0 + nrows) of an H x W
   * view (R == nrows * W) are generated into the workspace (star_get_rays arithmetic)                               */
  const float* rays_o; const float* rays_d; const float* viewdirs;
  int32_t H, W, row0, nrows; float fx, fy, cx, cy; const float* c2w;
  /* coarse depths: z_vals [R,Nc] given by the caller (optionally with pts [R,Nc,3]), or z_vals == NULL and
   * t_vals [Nc] (= torch.linspace(0,1,Nc)) + optional t_rand [R,Nc] jitter: sampled here                          */
  const float* z_vals; const float* pts; const float* t_vals; const float* t_rand;
  /* inverse-CDF draws: u [R,Ni], or u_det [Ni] (= torch.linspace(0,1,Ni), deterministic / eval); z_samples [R,Ni]
   * injects the fine samples themselves (skips the inversion)                                                      */
  const float* u; const float* u_det; const float* z_samples;
  const float* pose12;               /* [V,12] row-major [R | t] per object (V > 0)                                  */
  const float* enc_scale_xyz; const float* enc_scale_dir;   /* BARF masks of the object nets, or NULL              */
  const void* packed_static_coarse; const void* packed_static_fine;
  const void* const* packed_dynamic_coarse;  /* HOST arrays of V device pointers (star_pack_weights images)        */
  const void* const* packed_dynamic_fine;
} StarRenderIn;

typedef struct StarRenderOut {
  StarMultiOut coarse, fine;         /* per-pass outputs ("...0" keys / plain keys of the reference dict)            */
  float* dists0; float* dists;       /* [R,Nc] / [R,Nc+Ni], V == 0 only (raw2outputs' `dists`), may be NULL         */
  float* z_vals0;                    /* [R,Nc]: the coarse depths when sampled here (NULL if the caller gave z_vals) */
  float* z_vals;                     /* [R,Nc+Ni] merged depths of the fine pass                                     */
  float* z_samples;                  /* [R,Ni] (NULL when injected) */
  float* z_std;                      /* [R] */
  float* rays_o; float* rays_d; float* viewdirs;   /* camera mode: where the generated rays go ([R,3] each, or NULL) */
} StarRenderOut;

size_t star_render_workspace_bytes(const StarRenderCfg* cfg);
int star_render_forward(const StarRenderCfg* cfg, const StarRenderIn* in, const StarRenderOut* out, void* workspace,
                        size_t workspace_bytes, int32_t* status, void* stream);

/* ================================================================================================
 * a12: the mip-NeRF / integrated-positional-encoding variant
 * (models/star_mipnerf.py:99-357, models/rendering_starmip.py:32-175, models/mipnerf.py:53-100; the arithmetic
 * is nerfstudio's -- un-vendored and un-pinned by the reference: "parity unpinned", see oracle/mip_oracle.py).
 * ================================================================================================ */

/* ---- nerfstudio UniformSampler (star_mipnerf.py:75-77,271): Nc+1 frustum edges per ray --------
 * lin[Nc+1] = torch.linspace(0,1,Nc+1) from the host; t_rand[R,Nc+1] or NULL (training-mode stratified jitter).
 * Outputs spacing[R,Nc+1] (in [0,1]) and euclid[R,Nc+1] = spacing*far + (1-spacing)*near. */
int star_mip_uniform_bins(const float* lin, const float* t_rand, float near_, float far_, int R, int Nc,
                          float* spacing, float* euclid, void* stream);

/* ---- nerfstudio PDFSampler(include_original=False, histogram_padding=0.01) (star_mipnerf.py:286-288) ----
 * spacing_bins[R,Nc+1], weights[R,Nc] (row stride w_stride) -> Ni+1 new edges per ray.
 * u_base[Ni+1]: host-side linspace(0, 1-1/nb, nb) (+ 1/(2 nb) in eval mode); u_rand[R,Ni+1] or NULL: training jitter
 * (u = u_base + u_rand/nb).  Optional inds[R,Ni+1] (int64, searchsorted side="right") and cdf[R,Nc+1] for the
 * bit-exactness tests.  Same defined arithmetic as star_sample_pdf. */
int star_mip_pdf_sample(const float* spacing_bins, const float* weights, int64_t w_stride, const float* u_base,
                        const float* u_rand, float near_, float far_, int R, int Nc, int Ni, float* spacing_out,
                        float* euclid_out, int64_t* inds, float* cdf, void* stream);

/* ---- the field: models/mipnerf.py:89-100 -> nerfstudio NeRFField(use_integrated_encoding=True) ---------
 * Flat master order (each nn.Linear as weight [out,in] then bias [out]): field.mlp_base.layers.0..7,
 * field.field_output_density.net, field.mlp_head.layers.0, .1, field.field_heads.0.net  (589 572 floats).
 * origins[R,3], dirs[R,3] (normalised view directions), bins[R,S+1] euclidean frustum edges; pose12 = NULL (static
 * field) or [R(3x3) | t]: o' = R o + t, d' = R d (star_mipnerf.py:206-214).  freqs[64]: host table
 * [2**linspace(0,24,24) | its square | 2**linspace(0,4,4) | pad].  radius = sqrt(pixel_area)/sqrt(pi) (= 0.5642:
 * the reference passes pixel_area = 1, star_mipnerf.py:267).
 * Outputs are RAW (pre-softplus density, pre-sigmoid rgb), written with the strides of star_mlp_forward.
 * Precision tiers: STAR_PREC_F32 (CUDA cores, forward + backward incl. the pose accumulators) and STAR_PREC_BF16 /
 * STAR_PREC_F16 (tcgen05 tensor cores; the frequency table is baked into the packed image): forward, and -- with a stash --
 * a backward pass to the WEIGHTS (star_mip_field_backward with pose_acc == NULL; with pose_acc != NULL it returns
 * STAR_E_UNSUPPORTED: the gradient of the ray through the integrated positional encoding exists on the fp32 tier only). */
size_t star_mip_param_count(void);
size_t star_mip_packed_bytes(int precision);
int star_mip_pack_weights(int precision, const float* flat_master, const float* freqs, void* packed, void* stream);
size_t star_mip_stash_bytes(int precision, int64_t n_samples);
size_t star_mip_backward_workspace_bytes(int precision, int64_t n_samples);
int star_mip_field_forward(int precision, const void* packed, const float* origins, const float* dirs,
                           const float* pose12, const float* bins, const float* freqs, float radius, int R, int S,
                           float* raw_sigma, float* raw_rgb, int64_t ray_stride, void* stash, void* stream);
/* grad_flat: ACCUMULATED into; pose_acc[32] as in star_mlp_backward with p = the ray origin
 * ([0:3] sum g, [3:12] sum g o^T, [12:15] sum o' x g, [15:24] sum h d^T, [24:27] sum d' x h). */
int star_mip_field_backward(int precision, const void* packed, const float* origins, const float* dirs,
                            const float* pose12, const float* bins, const float* freqs, float radius, int R, int S,
                            const float* d_raw_sigma, const float* d_raw_rgb, int64_t ray_stride, const void* stash,
                            void* workspace, float* grad_flat, float* pose_acc, void* stream);

/* ---- density-space compositing: models/rendering_starmip.py:32-91 (single field) -----------------------
 * sigma = softplus(raw), colour = sigmoid(raw); depth = nerfstudio median DepthRenderer.
 * Outputs rgb[R,3], acc[R], depth[R], weights[R,S]. */
int star_mip_composite_single_forward(const float* raw_sigma, const float* raw_rgb, const float* bins, int R, int S,
                                      float* rgb, float* acc, float* depth, float* weights, void* stream);
int star_mip_composite_single_backward(const float* raw_sigma, const float* raw_rgb, const float* bins, int R, int S,
                                       const float* g_rgb, const float* g_acc, const float* g_weights,
                                       float* d_raw_sigma, float* d_raw_rgb, void* stream);

/* ---- models/rendering_starmip.py:112-175 (static + V dynamic fields, regularisers) ----------------------
 * regs[5] = alpha_entropy, dynamic_vs_static, ray_reg, static_reg, dynamic_reg with the mip shapes' quirks
 * (ray_reg has no max over samples; static_reg is identically 0 -- oracle/mip_oracle.py online_outputs). */
typedef struct StarMipMultiOut {
  float* rgb;                   /* [R,3]   */
  float* acc;                   /* [R]     */
  float* depth;                 /* [R]     */
  float* weights;               /* [R,S]   */
  float* rgb_static;            /* [R,3]   */
  float* depth_static;          /* [R]     */
  float* rgb_dynamic;           /* [R,V,3] */
  float* depth_dynamic;         /* [R,V]   */
  float* dynamic_transmittance; /* [R,V]   */
  float* regs;                  /* [5]     */
} StarMipMultiOut;
size_t star_mip_composite_multi_ws_bytes(int R);
int star_mip_composite_multi_forward(const float* raw_sigma_s, const float* raw_rgb_s, const float* raw_sigma_d,
                                     const float* raw_rgb_d, const float* bins, int R, int V, int S, int chunk,
                                     const StarMipMultiOut* out, void* workspace, void* stream);
int star_mip_composite_multi_backward(const float* raw_sigma_s, const float* raw_rgb_s, const float* raw_sigma_d,
                                      const float* raw_rgb_d, const float* bins, int R, int V, int S, int chunk,
                                      const float* g_rgb, const float* g_acc, const float* g_weights,
                                      const float* g_regs, float* d_raw_sigma_s, float* d_raw_rgb_s,
                                      float* d_raw_sigma_d, float* d_raw_rgb_d, void* stream);

/* ==== SURVEY.md 8(f) rows 2-3: loss and optimiser step either side of the render path =========
 * All reductions accumulate in fp64 and are deterministic (per-block partials added in block order).
 * `ws` = caller workspace of star_train_ws_bytes() bytes, 16-byte aligned; a call may overwrite it. */
size_t star_train_ws_bytes(void);

/* train_online__.py:158-166 / train_app_init__.py: loss = MSELoss(rgb0, target) + MSELoss(rgb, target) and
 * models/rendering__.py:22-23 mse2psnr.  n = number of elements (R*3).  out5 = [mse0, mse, psnr0, psnr, mse0 + mse];
 * g_rgb0 / g_rgb (nullable) receive d(mse0 + mse)/d(rgb0 | rgb) = 2 (x - target) / n.  rgb0 may be NULL
 * (N_importance == 0): mse0 = psnr0 = 0. */
int star_photometric_loss(const float* rgb0, const float* rgb, const float* target, int64_t n, float* out5,
                          float* g_rgb0, float* g_rgb, void* ws, void* stream);

/* models/loss.py:4-10 compute_depth_loss.  out2 = [loss, number of rays with near < gt < far] (loss is NaN when no
 * ray qualifies, like torch.mean of an empty tensor).  backward: g_depth[R] = g_out[0] * d loss / d depth. */
int star_depth_loss_forward(const float* depth, const float* gt_depth, int64_t R, float near_, float far_, float* out2,
                            void* ws, void* stream);
int star_depth_loss_backward(const float* depth, const float* gt_depth, int64_t R, float near_, float far_,
                             const float* out2, const float* g_out, float* g_depth, void* stream);

/* models/loss.py:13-66 compute_sigma_loss (weights, z_vals, dists [R,S]; depths [R]); per_ray (nullable, [R]) receives
 * the un-masked per-ray sums of compute_sigma_loss_per_ray (:70-87).  backward: gradient w.r.t. weights only (z_vals
 * and dists carry no gradient on the render path: z_samples are detached, rendering__.py:135); g_out = the scalar
 * upstream gradient, or with g_per_ray != 0 one upstream gradient per ray for the per-ray variant (mask ignored). */
int star_sigma_loss_forward(const float* weights, const float* z_vals, const float* dists, const float* depths, int64_t R,
                            int S, float near_, float far_, float err, float* out2, float* per_ray, void* ws,
                            void* stream);
int star_sigma_loss_backward(const float* weights, const float* z_vals, const float* dists, const float* depths,
                             int64_t R, int S, float near_, float far_, float err, const float* out2, const float* g_out,
                             int g_per_ray, float* g_weights, void* stream);

/* One contiguous run of fp32 parameters with its gradient and Adam moments (all DEVICE pointers, 4-byte aligned; the
 * 16-byte path is taken where the four share their alignment).  step_size = lr / (1 - beta1^t) and
 * bc2_sqrt = sqrt(1 - beta2^t) are formed by the host in double, as torch.optim.Adam does. */
typedef struct StarAdamSeg {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  float step_size;
  float bc2_sqrt;
} StarAdamSeg;

/* `segs` is a HOST array.  star_grad_sqnorm leaves sum(grad^2) over all segments as a double at
 * star_grad_sqnorm_result(ws) (a device address inside ws).  star_grad_scale = torch.nn.utils.clip_grad_norm_
 * (Trainer(gradient_clip_val=1.0), train_online__.py:1170): grad *= min(1, max_norm / (sqrt(sqnorm) + 1e-6)).
 * star_adam_step = torch.optim.Adam(betas, eps, amsgrad=False, weight_decay=0) (train_online__.py:333-353) on every
 * segment, with the clip coefficient applied to the gradient on the fly when sqnorm != NULL (write_back_grads != 0
 * also stores the clipped gradient, as the reference's in-place clip leaves it). */
int star_grad_sqnorm(const StarAdamSeg* segs, int n_segs, void* ws, void* stream);
const double* star_grad_sqnorm_result(const void* ws);
int star_grad_scale(const StarAdamSeg* segs, int n_segs, const double* sqnorm, float max_norm, void* stream);
int star_adam_step(const StarAdamSeg* segs, int n_segs, double beta1, double beta2, double eps, const double* sqnorm,
                   float max_norm, int write_back_grads, void* stream);

/* ==== SURVEY.md 8(f) row 4, device side: utils/metrics.py:527-550 compute_2d_iou ================
 * dyn_t [R, V] = dynamic_transmittance of render_star_online; sem [R] bytes (non-zero = vehicle pixel of the semantic
 * mask).  pred (nullable) [V, R] bytes <- (dyn_t[:, v] < thres); counts[2] (int64) <- {|sem AND any_v pred|,
 * |sem OR any_v pred|}; the host forms iou = counts[0] / counts[1] (0 when the union is empty).  Bit-exact. */
int star_iou2d(const float* dyn_t, const uint8_t* sem, int64_t R, int V, float thres, uint8_t* pred, int64_t* counts,
               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STAR_B200_H_ */
