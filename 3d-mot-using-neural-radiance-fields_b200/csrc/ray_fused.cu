// Fused coarse-pass tail of a single-field render (north-star bullet 1): ONE kernel per ray does the alpha compositing of
// the coarse samples (raw2outputs, models/rendering__.py:307-379: warp-shuffle prefix product of the transmittance), the
// sample-PDF CDF of weights[1:-1] (sample_pdf :719-761: fp64 warp prefix sums), the inverse-CDF draw, z_std and the
// sort(cat(z_vals, z_samples)) merge (:128-144) -- the weights and the depths of the ray go from the compositing part to the
// sampling part through shared memory instead of a global round trip and a second launch.  Same per-ray device functions
// as the stand-alone kernels (ray_device.cuh), so star_composite_single_forward + star_hierarchical give the same bits.
#include "ray_device.cuh"

__global__ void __launch_bounds__(128) composite_hier_kernel(
    const float* __restrict__ raw_alpha, const float* __restrict__ raw_rgb, const float* __restrict__ z_vals,
    const float* __restrict__ rays_d, const float* __restrict__ u, const float* __restrict__ u_det, int R, int Nc, int Ni,
    int P, float far_dist, int white_bkgd, float* __restrict__ rgb_o, float* __restrict__ disp_o, float* __restrict__ acc_o,
    float* __restrict__ depth_o, float* __restrict__ weights_o, float* __restrict__ dists_o, float* z_samples,
    float* __restrict__ z_all, float* __restrict__ z_std) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nb = Nc - 1, Nf = Nc + Ni;
  // per warp: the hierarchical step's regions, then the ray's Nc coarse weights
  float* za = smem + (size_t)warp * (hier_smem_floats(Nc, Ni, P) + ((Nc + 3) & ~3));
  float* zs = za + ((Nf + 3) & ~3);
  float* zc = zs + P;
  float* cdf = zc + Nc;
  float* sb = cdf + nb;
  int* gs = reinterpret_cast<int*>(sb + nb);
  float* wsm = za + hier_smem_floats(Nc, Ni, P);
  const bool vec4 = ((Nf & 3) == 0) && ((uintptr_t)z_all & 15) == 0;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    composite_single_ray_x2(raw_alpha, raw_rgb, z_vals, rays_d, r, Nc, far_dist, white_bkgd, rgb_o, disp_o, acc_o, depth_o,
                            weights_o, dists_o, wsm, zc, lane);
    __syncwarp();
    hier_ray<false, true>(zc, wsm, u, u_det, nullptr, nullptr, r, Nc, Ni, P, z_samples, z_all, z_std, nullptr, za, zs, zc,
                          cdf, sb, gs, vec4, lane);
  }
}

// Coarse compositing + hierarchical step of R rays in one launch.  Nc even, >= 4 (the two-samples-per-lane form), 8-byte
// aligned arrays; everything else: STAR_E_UNSUPPORTED and the caller takes the two stand-alone entries.  weights / dists
// may be NULL (a caller that only wants the maps and the fine depths saves their 8 B per sample).
extern "C" int star_composite_hier_forward(const float* raw_alpha, const float* raw_rgb, const float* z_vals,
                                           const float* rays_d, const float* u, const float* u_det, int R, int Nc, int Ni,
                                           float far_dist, int white_bkgd, float* rgb, float* disp, float* acc, float* depth,
                                           float* weights, float* dists, float* z_samples, float* z_all, float* z_std,
                                           void* stream) {
  if (R == 0) return STAR_OK;     /* an empty batch carries no pointers (torch: data_ptr() == 0 for numel() == 0) */
  if (!raw_alpha || !raw_rgb || !z_vals || !rays_d || !rgb || !disp || !acc || !depth || !z_samples || !z_all || !z_std ||
      (!u && !u_det))
    return STAR_E_NULL;
  if (R < 0 || Nc < 3 || Ni < 1 || Nc + Ni > 8192) return STAR_E_BAD_SHAPE;
  if ((Nc & 1) != 0 || Nc < 4) return STAR_E_UNSUPPORTED;
  if ((((uintptr_t)raw_alpha | (uintptr_t)raw_rgb | (uintptr_t)z_vals | (uintptr_t)weights | (uintptr_t)dists) & 7) != 0)
    return STAR_E_UNSUPPORTED;
  int P = 2;
  while (P < Ni) P <<= 1;
  const size_t per_warp = sizeof(float) * ((size_t)hier_smem_floats(Nc, Ni, P) + ((Nc + 3) & ~3));
  int wpb = 4;
  while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
  if (per_warp * wpb > 200 * 1024) return STAR_E_BAD_SHAPE;
  const size_t smem = per_warp * wpb;
  int64_t blocks = ((int64_t)R + wpb - 1) / wpb;
#ifndef STAR_FUSED_GRID_CAP
#define STAR_FUSED_GRID_CAP 0x3fffffff     // grid over all rays (profiles/r2y_ab_grid_cap.txt)
#endif
  if (blocks > STAR_FUSED_GRID_CAP) blocks = STAR_FUSED_GRID_CAP;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(composite_hier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  composite_hier_kernel<<<(int)blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
      raw_alpha, raw_rgb, z_vals, rays_d, u, u_det, R, Nc, Ni, P, far_dist, white_bkgd, rgb, disp, acc, depth, weights, dists,
      z_samples, z_all, z_std);
  return star_check_launch();
}
