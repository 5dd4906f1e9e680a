// Micro-benchmark: TMEM -> register read bandwidth (tcgen05.ld 32x32b.x32) with 4 / 8 / 16 warps per SM.
#include <cstdio>
#include "../3d-mot-using-neural-radiance-fields_b200/csrc/tc_common.cuh"

__global__ void __launch_bounds__(512, 1) tmem_kernel(int niter, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t base = t + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * 32 % 512);
    for (int i = 0; i < niter; ++i) {
      uint32_t r[32];
      tmem_ld32(base + (uint32_t)((i * 64) & 255), r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(t, 512); }
}

int main() {
  long long* d_out; uint32_t* sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&sink, 4096);
  for (int nw : {1, 4, 8, 16}) {
    const int niter = 4000;
    tmem_kernel<<<148, 512>>>(niter, nw, d_out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("warps=%2d: %lld cycles for %d x LDTM.x32 per warp -> %.1f cycles per LDTM per warp, %.1f B/clk/SM (%s)\n", nw, cyc,
           niter, (double)cyc / niter, 4096.0 * nw * niter / cyc, cudaGetErrorString(e));
  }
  return 0;
}
