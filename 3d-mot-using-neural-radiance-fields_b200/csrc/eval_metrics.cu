// SURVEY.md section 8(f) row 4 (device side of the online-tracking evaluation): utils/metrics.py:527-550 compute_2d_iou
// -- per-object masks dynamic_transmittance[:, v] < thres, their union over the objects, and the intersection / union
// pixel counts against the semantic mask -- in one pass over the [R, V] transmittances instead of V device->host copies
// and numpy loops.  Integer work: bit-exact.
#include "star_common.cuh"

namespace {

__global__ void iou2d_kernel(const float* __restrict__ dyn_t, const uint8_t* __restrict__ sem, int64_t R, int V, float thres,
                             uint8_t* __restrict__ pred, unsigned long long* __restrict__ counts) {
  unsigned long long inter = 0ull, uni = 0ull;
  for (int64_t r0 = blockIdx.x * (int64_t)blockDim.x; r0 < R; r0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = r0 + threadIdx.x;
    bool any = false, s = false;
    if (r < R) {
      for (int v = 0; v < V; ++v) {
        const bool m = dyn_t[r * V + v] < thres;     // NaN compares false, as in torch
        if (pred != nullptr) pred[(int64_t)v * R + r] = m ? 1 : 0;
        any |= m;
      }
      s = sem[r] != 0;
    }
    inter += __popc(__ballot_sync(STAR_FULL_MASK, s && any));
    uni += __popc(__ballot_sync(STAR_FULL_MASK, s || any));
  }
  if ((threadIdx.x & 31) == 0) {
    if (inter) atomicAdd(counts + 0, inter);
    if (uni) atomicAdd(counts + 1, uni);
  }
}

}  // namespace

// dyn_t [R, V] fp32 (dynamic_transmittance of render_star_online), sem [R] bytes (non-zero = vehicle pixel),
// pred (nullable) [V, R] bytes <- per-object masks, counts[2] <- {intersection, union} pixel counts (int64).
extern "C" int star_iou2d(const float* dyn_t, const uint8_t* sem, int64_t R, int V, float thres, uint8_t* pred,
                          int64_t* counts, void* stream) {
  if (!counts || (R != 0 && (!dyn_t || !sem))) return STAR_E_NULL;     /* R == 0: empty arrays carry no pointers */
  if (R < 0 || V < 1 || V > STAR_MAX_V) return STAR_E_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const cudaError_t e = cudaMemsetAsync(counts, 0, 2 * sizeof(int64_t), st);
  if (e != cudaSuccess) {
    g_star_last_cuda_error = (int)e;
    return STAR_E_CUDA;
  }
  if (R == 0) return STAR_OK;
  int64_t blocks = (R + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  iou2d_kernel<<<(int)blocks, 256, 0, st>>>(dyn_t, sem, R, V, thres, pred, reinterpret_cast<unsigned long long*>(counts));
  return star_check_launch();
}
