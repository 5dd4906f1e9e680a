// extern "C" entry points of libstar_b200.so that are not kernel-file local (see include/star_b200.h).
#include "star_common.cuh"
#include "mlp_layout.h"
#include <stdlib.h>
#include "mlp_tc_layout.h"
#include "mip_layout.h"

int g_star_last_cuda_error = 0;
int g_star_sync_launches = []() { const char* e = getenv("STAR_B200_SYNC_LAUNCHES"); return (e && e[0] == '1') ? 1 : 0; }();

static inline int star_prec(const StarNetDesc* d) { return d->precision & 0xff; }   // (upper bits: STAR_PREC_FLAG_*)

size_t star_tc_packed_bytes(const TcLayout& tl);
int star_tc_pack(const TcLayout& tl, const MlpLayout& ml, const float* master, void* packed, int fp16,
                 cudaStream_t st);
int star_tc_forward(const TcLayout& tl, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                    const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, float* raw_alpha,
                    float* raw_rgb, int64_t ray_stride, void* stash, int* status, int fp16, cudaStream_t st);

size_t star_tc_gstash_bytes(const TcLayout& tl, int64_t n_samples);
int star_tc_backward(const TcLayout& tl, const MlpLayout& ml, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                     const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, const float* d_raw_alpha,
                     const float* d_raw_rgb, int64_t ray_stride, const void* stash, void* gstash, float* grad_flat,
                     float* pose_acc, int fp16, int serial_dx, cudaStream_t st);

int star_f32_pack(const MlpLayout& lay, const float* master, void* packed, cudaStream_t st);
int star_f32_forward(const MlpLayout& lay, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                     const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, float* raw_alpha,
                     float* raw_rgb, int64_t ray_stride, void* stash, cudaStream_t st);
int star_f32_backward(const MlpLayout& lay, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                      const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S,
                      const float* d_raw_alpha, const float* d_raw_rgb, int64_t ray_stride, const void* stash,
                      void* workspace, float* grad_flat, float* pose_acc, cudaStream_t st);

extern "C" int star_abi_version(void) { return STAR_ABI_VERSION; }
extern "C" int star_last_cuda_error(void) { return g_star_last_cuda_error; }

// Watchdog words of the tensor-core kernels: mapped host memory (readable after the context died), one int per kernel family
// (STAR_WD_*).  A barrier wait that exceeds STAR_TC_WATCHDOG_CYCLES stores (wait code << 16 | CTA) there and traps.
static int* g_wd_host = nullptr;
static int* g_wd_dev = nullptr;
int* star_watchdog_dev(int family) {
  if (g_wd_host == nullptr) {
    if (cudaHostAlloc((void**)&g_wd_host, 256, cudaHostAllocMapped) != cudaSuccess) { g_wd_host = nullptr; return nullptr; }
    for (int i = 0; i < 64; ++i) g_wd_host[i] = 0;
    if (cudaHostGetDevicePointer((void**)&g_wd_dev, g_wd_host, 0) != cudaSuccess) return nullptr;
  }
  return g_wd_dev ? g_wd_dev + family : nullptr;
}
extern "C" int star_watchdog_word(int family) {
  return (g_wd_host != nullptr && family >= 0 && family < 64) ? ((volatile int*)g_wd_host)[family] : 0;
}

extern "C" const char* star_error_string(int code) {
  switch (code) {
    case STAR_OK: return "ok";
    case STAR_E_BAD_SHAPE: return "size out of the supported range";
    case STAR_E_UNSUPPORTED: return "unsupported width / depth / encoding / precision";
    case STAR_E_NULL: return "required pointer is NULL";
    case STAR_E_ALIGN: return "pointer alignment violated";
    case STAR_E_WORKSPACE: return "workspace too small";
    case STAR_E_CUDA: return "CUDA runtime error (see star_last_cuda_error)";
    default: return "unknown error code";
  }
}

extern "C" size_t star_net_param_count(const StarNetDesc* d) {
  MlpLayout lay;
  if (!d || star_make_layout(d, &lay)) return 0;
  return (size_t)lay.n_master;
}

extern "C" size_t star_packed_bytes(const StarNetDesc* d) {
  MlpLayout lay;
  if (!d || star_make_layout(d, &lay)) return 0;
  if (star_prec(d) == STAR_PREC_F32) return sizeof(float) * (size_t)lay.n_packed;
  if (star_prec(d) == STAR_PREC_BF16 || star_prec(d) == STAR_PREC_F16) {
    TcLayout tl;
    if (star_make_tc_layout(d, &tl)) return 0;
    return star_tc_packed_bytes(tl);
  }
  return 0;
}

extern "C" int star_pack_weights(const StarNetDesc* d, const float* flat_master, void* packed, void* stream) {
  if (!d || !flat_master || !packed) return STAR_E_NULL;
  MlpLayout lay;
  int rc = star_make_layout(d, &lay);
  if (rc) return rc;
  if (star_prec(d) == STAR_PREC_F32) return star_f32_pack(lay, flat_master, packed, (cudaStream_t)stream);
  if (star_prec(d) == STAR_PREC_BF16 || star_prec(d) == STAR_PREC_F16) {
    TcLayout tl;
    rc = star_make_tc_layout(d, &tl);
    if (rc) return rc;
    return star_tc_pack(tl, lay, flat_master, packed, star_prec(d) == STAR_PREC_F16, (cudaStream_t)stream);
  }
  return STAR_E_UNSUPPORTED;
}

extern "C" size_t star_stash_bytes(const StarNetDesc* d, int64_t n_samples) {
  MlpLayout lay;
  if (!d || star_make_layout(d, &lay) || n_samples < 0) return 0;
  if (star_prec(d) == STAR_PREC_F32) return sizeof(float) * (size_t)lay.stash_cols * (size_t)n_samples;
  if (star_prec(d) == STAR_PREC_BF16 || star_prec(d) == STAR_PREC_F16) {
    TcLayout tl;
    if (star_make_tc_layout(d, &tl)) return 0;
    // (an even number of tiles: the CTA-pair kernels walk pairs of tiles and may write a phantom last tile)
    return (size_t)(((n_samples + 127) / 128 + 1) / 2 * 2) * (size_t)tl.stash_blocks * TC_BLOCK_BYTES;
  }
  return 0;
}

extern "C" size_t star_mlp_backward_workspace_bytes(const StarNetDesc* d, int64_t n_samples) {
  MlpLayout lay;
  if (!d || star_make_layout(d, &lay) || n_samples < 0) return 0;
  if (star_prec(d) == STAR_PREC_F32) return sizeof(float) * (size_t)lay.g_cols * (size_t)n_samples;
  if (star_prec(d) == STAR_PREC_BF16 || star_prec(d) == STAR_PREC_F16) {
    TcLayout tl;
    if (star_make_tc_layout(d, &tl)) return 0;
    return star_tc_gstash_bytes(tl, n_samples);
  }
  return 0;
}

extern "C" int star_mlp_forward(const StarNetDesc* d, const void* packed, const float* pts_or_null,
                                const float* rays_o, const float* rays_d, const float* z_vals, const float* viewdirs,
                                const float* pose12, const float* enc_scale_xyz, const float* enc_scale_dir, int R,
                                int S, float* raw_alpha, float* raw_rgb, int64_t alpha_ray_stride, void* stash,
                                int32_t* status, void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!d || !packed || !viewdirs || !raw_alpha || !raw_rgb) return STAR_E_NULL;
  if (!pts_or_null && (!rays_o || !rays_d || !z_vals)) return STAR_E_NULL;
  const StarPtsSrc pts{pts_or_null, rays_o, rays_d, z_vals};
  if (R < 0 || S < 1 || alpha_ray_stride < S) return STAR_E_BAD_SHAPE;
  MlpLayout lay;
  int rc = star_make_layout(d, &lay);
  if (rc) return rc;
  if (star_prec(d) == STAR_PREC_F32)
    return star_f32_forward(lay, packed, pts, viewdirs, pose12, enc_scale_xyz, enc_scale_dir, R, S, raw_alpha,
                            raw_rgb, alpha_ray_stride, stash, (cudaStream_t)stream);
  if (star_prec(d) == STAR_PREC_BF16 || star_prec(d) == STAR_PREC_F16) {
    if (d->precision & STAR_PREC_FLAG_RETIRED) return STAR_E_UNSUPPORTED;
    TcLayout tl;
    rc = star_make_tc_layout(d, &tl);
    if (rc) return rc;
    return star_tc_forward(tl, packed, pts, viewdirs, pose12, enc_scale_xyz, enc_scale_dir, R, S, raw_alpha, raw_rgb,
                           alpha_ray_stride, stash, status,
                           (star_prec(d) == STAR_PREC_F16 ? 1 : 0) | ((d->precision & STAR_PREC_FLAG_NO_WSHARE) ? 2 : 0),
                           (cudaStream_t)stream);
  }
  return STAR_E_UNSUPPORTED;
}

extern "C" int star_mlp_backward(const StarNetDesc* d, const void* packed, const float* flat_master,
                                 const float* pts_or_null, const float* rays_o, const float* rays_d,
                                 const float* z_vals, const float* viewdirs, const float* pose12,
                                 const float* enc_scale_xyz, const float* enc_scale_dir, int R, int S,
                                 const float* d_raw_alpha, const float* d_raw_rgb, int64_t alpha_ray_stride,
                                 const void* stash, void* workspace, float* grad_flat, float* pose_acc,
                                 void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  (void)flat_master;
  if (!d || !packed || !viewdirs || !d_raw_alpha || !d_raw_rgb || !stash || !workspace || !grad_flat)
    return STAR_E_NULL;
  if (!pts_or_null && (!rays_o || !rays_d || !z_vals)) return STAR_E_NULL;
  const StarPtsSrc pts{pts_or_null, rays_o, rays_d, z_vals};
  if (pose12 && !pose_acc) return STAR_E_NULL;
  if (R < 0 || S < 1 || alpha_ray_stride < S) return STAR_E_BAD_SHAPE;
  MlpLayout lay;
  int rc = star_make_layout(d, &lay);
  if (rc) return rc;
  if (star_prec(d) == STAR_PREC_F32)
    return star_f32_backward(lay, packed, pts, viewdirs, pose12, enc_scale_xyz, enc_scale_dir, R, S, d_raw_alpha,
                             d_raw_rgb, alpha_ray_stride, stash, workspace, grad_flat, pose_acc,
                             (cudaStream_t)stream);
  if (star_prec(d) == STAR_PREC_BF16 || star_prec(d) == STAR_PREC_F16) {
    TcLayout tl;
    rc = star_make_tc_layout(d, &tl);
    if (rc) return rc;
    return star_tc_backward(tl, lay, packed, pts, viewdirs, pose12, enc_scale_xyz, enc_scale_dir, R, S, d_raw_alpha,
                            d_raw_rgb, alpha_ray_stride, stash, workspace, grad_flat, pose_acc,
                            star_prec(d) == STAR_PREC_F16, (d->precision & STAR_PREC_FLAG_DX_PIPELINED) == 0,
                            (cudaStream_t)stream);
  }
  return STAR_E_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------- a12 mip field
int star_mip_f32_pack(const MipLayout& lay, const float* master, void* packed, cudaStream_t st);
int star_mip_f32_forward(const MipLayout& lay, const void* packed, const float* origins, const float* dirs,
                         const float* pose12, const float* bins, const float* freqs, float radius, int R, int S,
                         float* raw_sigma, float* raw_rgb, int64_t ray_stride, void* stash, cudaStream_t st);
int star_mip_f32_backward(const MipLayout& lay, const void* packed, const float* origins, const float* dirs,
                          const float* pose12, const float* bins, const float* freqs, float radius, int R, int S,
                          const float* d_raw_sigma, const float* d_raw_rgb, int64_t ray_stride, const void* stash,
                          void* workspace, float* grad_flat, float* pose_acc, cudaStream_t st);

size_t star_mip_tc_packed_bytes();
int star_mip_tc_pack(const float* master, const float* freqs, void* packed, int fp16, cudaStream_t st);
int star_mip_tc_forward(const void* packed, const float* origins, const float* dirs, const float* pose12,
                        const float* bins, float radius, int R, int S, float* raw_sigma, float* raw_rgb,
                        int64_t ray_stride, void* stash, int fp16, cudaStream_t st);
size_t star_mip_tc_stash_bytes(int64_t n_samples);
size_t star_mip_tc_gstash_bytes(int64_t n_samples);
int star_mip_tc_backward(const void* packed, int R, int S, const float* d_raw_sigma, const float* d_raw_rgb,
                         int64_t ray_stride, const void* stash, void* gstash, float* grad_flat, int fp16,
                         cudaStream_t st);

extern "C" size_t star_mip_param_count(void) {
  MipLayout lay;
  star_make_mip_layout(&lay);
  return (size_t)lay.n_master;
}

extern "C" size_t star_mip_packed_bytes(int precision) {
  MipLayout lay;
  star_make_mip_layout(&lay);
  if (precision == STAR_PREC_F32) return sizeof(float) * (size_t)lay.n_packed;
  if (precision == STAR_PREC_BF16 || precision == STAR_PREC_F16) return star_mip_tc_packed_bytes();
  return 0;
}

extern "C" int star_mip_pack_weights(int precision, const float* flat_master, const float* freqs, void* packed,
                                     void* stream) {
  if (!flat_master || !packed) return STAR_E_NULL;
  MipLayout lay;
  star_make_mip_layout(&lay);
  if (precision == STAR_PREC_F32) return star_mip_f32_pack(lay, flat_master, packed, (cudaStream_t)stream);
  if (precision == STAR_PREC_BF16 || precision == STAR_PREC_F16) {
    if (!freqs) return STAR_E_NULL;
    return star_mip_tc_pack(flat_master, freqs, packed, precision == STAR_PREC_F16, (cudaStream_t)stream);
  }
  return STAR_E_UNSUPPORTED;
}

extern "C" size_t star_mip_stash_bytes(int precision, int64_t n_samples) {
  MipLayout lay;
  star_make_mip_layout(&lay);
  if (n_samples < 0) return 0;
  if (precision == STAR_PREC_BF16 || precision == STAR_PREC_F16) return star_mip_tc_stash_bytes(n_samples);
  if (precision != STAR_PREC_F32) return 0;
  return sizeof(float) * (size_t)lay.stash_cols * (size_t)n_samples;
}

extern "C" size_t star_mip_backward_workspace_bytes(int precision, int64_t n_samples) {
  MipLayout lay;
  star_make_mip_layout(&lay);
  if (n_samples < 0) return 0;
  if (precision == STAR_PREC_BF16 || precision == STAR_PREC_F16) return star_mip_tc_gstash_bytes(n_samples);
  if (precision != STAR_PREC_F32) return 0;
  return sizeof(float) * (size_t)lay.g_cols * (size_t)n_samples;
}

extern "C" int star_mip_field_forward(int precision, const void* packed, const float* origins, const float* dirs,
                                      const float* pose12, const float* bins, const float* freqs, float radius, int R,
                                      int S, float* raw_sigma, float* raw_rgb, int64_t ray_stride, void* stash,
                                      void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!packed || !origins || !dirs || !bins || !freqs || !raw_sigma || !raw_rgb) return STAR_E_NULL;
  if (R < 0 || S < 1 || ray_stride < S) return STAR_E_BAD_SHAPE;
  MipLayout lay;
  star_make_mip_layout(&lay);
  if (precision == STAR_PREC_F32)
    return star_mip_f32_forward(lay, packed, origins, dirs, pose12, bins, freqs, radius, R, S, raw_sigma, raw_rgb,
                                ray_stride, stash, (cudaStream_t)stream);
  if (precision == STAR_PREC_BF16 || precision == STAR_PREC_F16) {
    return star_mip_tc_forward(packed, origins, dirs, pose12, bins, radius, R, S, raw_sigma, raw_rgb, ray_stride, stash,
                               precision == STAR_PREC_F16, (cudaStream_t)stream);
  }
  return STAR_E_UNSUPPORTED;
}

extern "C" int star_mip_field_backward(int precision, const void* packed, const float* origins, const float* dirs,
                                       const float* pose12, const float* bins, const float* freqs, float radius, int R,
                                       int S, const float* d_raw_sigma, const float* d_raw_rgb, int64_t ray_stride,
                                       const void* stash, void* workspace, float* grad_flat, float* pose_acc,
                                       void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!packed || !origins || !dirs || !bins || !freqs || !d_raw_sigma || !d_raw_rgb || !stash || !workspace ||
      !grad_flat)
    return STAR_E_NULL;
  if (pose12 && !pose_acc && precision == STAR_PREC_F32) return STAR_E_NULL;
  if (R < 0 || S < 1 || ray_stride < S) return STAR_E_BAD_SHAPE;
  MipLayout lay;
  star_make_mip_layout(&lay);
  if (precision == STAR_PREC_F32)
    return star_mip_f32_backward(lay, packed, origins, dirs, pose12, bins, freqs, radius, R, S, d_raw_sigma, d_raw_rgb,
                                 ray_stride, stash, workspace, grad_flat, pose_acc, (cudaStream_t)stream);
  if (precision == STAR_PREC_BF16 || precision == STAR_PREC_F16) {
    // tensor-core backward: gradients to the WEIGHTS only; a pass that needs the gradient of the ray (object pose)
    // must run on the fp32 tier
    if (pose_acc != nullptr) return STAR_E_UNSUPPORTED;
    return star_mip_tc_backward(packed, R, S, d_raw_sigma, d_raw_rgb, ray_stride, stash, workspace, grad_flat,
                                precision == STAR_PREC_F16, (cudaStream_t)stream);
  }
  return STAR_E_UNSUPPORTED;
}
