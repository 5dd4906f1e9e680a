// Micro-benchmark: how fast does a 16 KB bulk store (cp.async.bulk.global.shared::cta) release its shared-memory
// source, alone and while the same thread also streams 32 KB weight K-blocks from L2 into a ring the way the
// training forward kernel does?  One CTA per SM, one issuing thread.  Reports, for CTA 0 and the slowest CTA:
//   lat     cycles from issue to wait_group.read 0 of ONE store (all SMs storing at once, paced by `gap` cycles)
//   rate    cycles per 16 KB block with `depth` stores in flight (wait_group.read depth-1), back to back
// Modes: 0 = stores only, 1 = stores + one 32 KB bulk load from a (L2-resident) 1.6 MB stream per store.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store_rate bulk_store_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../3d-mot-using-neural-radiance-fields_b200/csrc/tc_common.cuh"

template <int DEPTH>
__device__ __forceinline__ void wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH) : "memory");
}

template <int DEPTH, int MODE>
__global__ void __launch_bounds__(128, 1) store_kernel(uint8_t* dst, const uint8_t* wstream, int niter, int gap,
                                                        long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = i;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t* my = dst + (size_t)blockIdx.x * (size_t)niter * 16384u;
    const uint32_t ring = base + 65536u;
    uint32_t ph[4] = {0, 0, 0, 0};
    long long worst = 0, t0 = clock64();
    for (int i = 0; i < niter; ++i) {
      if (MODE == 1) {
        const int s = i & 3;
        if (i >= 4) { mbar_wait(smem_u32(&bar[s]), ph[s], nullptr, 0); ph[s] ^= 1u; }
        mbar_arrive_expect_tx(smem_u32(&bar[s]), 32768u);
        bulk_g2s(ring + (uint32_t)s * 32768u, wstream + (size_t)(i % 48) * 32768u, 32768u, smem_u32(&bar[s]));
      }
      const long long a = clock64();
      bulk_s2g(my + (size_t)i * 16384u, base + (uint32_t)(i & 3) * 16384u, 16384u);
      bulk_commit_group();
      wait_read<DEPTH - 1>();
      const long long b = clock64();
      if (b - a > worst) worst = b - a;
      if (gap > 0) { while (clock64() - a < gap) {} }
    }
    wait_read<0>();
    const long long t1 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = worst;
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (MODE == 1)
      for (int s = 0; s < 4 && s < niter; ++s) mbar_wait(smem_u32(&bar[s]), ph[s], nullptr, 0);
  }
  __syncthreads();
}

template <int DEPTH, int MODE>
static void run(const char* name, uint8_t* dst, const uint8_t* w, int niter, int gap, long long* d_out, int sms) {
  const int smem = 65536 + 4 * 32768 + 2048;
  cudaFuncSetAttribute(store_kernel<DEPTH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  store_kernel<DEPTH, MODE><<<sms, 128, smem>>>(dst, w, niter, gap, d_out);
  cudaEventRecord(e0);
  store_kernel<DEPTH, MODE><<<sms, 128, smem>>>(dst, w, niter, gap, d_out);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long h[2 * 160];
  cudaMemcpy(h, d_out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
  long long mx = 0, wmx = 0;
  for (int i = 0; i < sms; ++i) { if (h[2 * i] > mx) mx = h[2 * i]; if (h[2 * i + 1] > wmx) wmx = h[2 * i + 1]; }
  printf("%-44s depth %d gap %5d: %7.1f cycles / 16 KB block (slowest SM %7.1f), worst issue->read-done %6lld, %6.2f TB/s written\n",
         name, DEPTH, gap, (double)h[0] / niter, (double)mx / niter, wmx, (double)sms * niter * 16384.0 / (ms * 1e-3) / 1e12);
}

int main() {
  int sms = 148, niter = 2048;
  uint8_t *dst, *w;
  long long* d_out;
  cudaMalloc(&dst, (size_t)sms * niter * 16384);
  cudaMalloc(&w, 48 * 32768);
  cudaMemset(w, 1, 48 * 32768);
  cudaMalloc(&d_out, sizeof(long long) * 2 * 160);
  run<1, 0>("stores only", dst, w, niter, 0, d_out, sms);
  run<2, 0>("stores only", dst, w, niter, 0, d_out, sms);
  run<4, 0>("stores only", dst, w, niter, 0, d_out, sms);
  run<1, 0>("stores only, paced (1 per 700 cycles)", dst, w, niter, 700, d_out, sms);
  run<1, 0>("stores only, paced (1 per 1000 cycles)", dst, w, niter, 1000, d_out, sms);
  run<1, 0>("stores only, paced (1 per 1500 cycles)", dst, w, niter, 1500, d_out, sms);
  run<2, 1>("stores + 32 KB L2 loads", dst, w, niter, 0, d_out, sms);
  run<1, 1>("stores + 32 KB L2 loads, paced 700", dst, w, niter, 700, d_out, sms);
  run<1, 1>("stores + 32 KB L2 loads, paced 1000", dst, w, niter, 1000, d_out, sms);
  run<1, 1>("stores + 32 KB L2 loads, paced 1500", dst, w, niter, 1500, d_out, sms);
  run<1, 0>("stores only, ONE SM", dst, w, niter, 0, d_out, 1);
  return 0;
}
