// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, M=128) from shared-memory operands.
// One CTA per SM; one thread issues NITER MMAs on (garbage) smem operands, commits, and the cycle count is
// reported.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../3d-mot-using-neural-radiance-fields_b200/csrc/tc_common.cuh"

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int niter, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr), 512);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_16(128, N, 1);
    const uint32_t a = base, b = base + 16384;
    long long t0 = clock64();
    for (int i = 0; i < niter; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(t + ((i & 1) ? 256u : 0u), umma_desc_sw128(a + 32u * k), umma_desc_sw128(b + 32u * k), idesc, 1u);
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(t, 512); }
}

// Same issue loop, but the operands walk through distinct shared-memory blocks the way the MLP kernel does:
// A = 4 K-blocks of 16 KB, B = a ring of 4 stages of 32 KB, 4 K-steps (+32 B) inside each block.
template <int N>
__global__ void __launch_bounds__(128, 1) stream_kernel(int niter, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr), 512);
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_16(128, N, 1);
    long long t0 = clock64();
    for (int i = 0; i < niter; ++i) {
      const uint32_t a = base + (uint32_t)(i & 3) * 16384u, b = base + 65536u + (uint32_t)(i & 3) * 32768u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(t + ((i & 4) ? 256u : 0u), umma_desc_sw128(a + 32u * k), umma_desc_sw128(b + 32u * k), idesc, 1u);
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(t, 512); }
}

// The MLP kernel's issuer loop, piece by piece: per K-block [VAR & 1: two try_waits on already-completed barriers]
// [VAR & 2: tcgen05.fence::after_thread_sync] 4 MMAs [VAR & 4: tcgen05.commit to a per-stage barrier]
template <int VAR>
__global__ void __launch_bounds__(128, 1) loop_kernel(int niter, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar, done_bar[2], stage_bar[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&done_bar[0]), 1); mbar_init(smem_u32(&done_bar[1]), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&stage_bar[i]), 1);
    fence_mbar_init();
    mbar_arrive(smem_u32(&done_bar[0]));   // phase 0 of both complete: try_wait(parity 0) succeeds immediately
    mbar_arrive(smem_u32(&done_bar[1]));
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr), 512);
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_16(128, 256, 1);
    long long t0 = clock64();
    for (int i = 0; i < niter; ++i) {
      const uint32_t a = base + (uint32_t)(i & 3) * 16384u, b = base + 65536u + (uint32_t)(i & 3) * 32768u;
      if (VAR & 1) {
        mbar_wait(smem_u32(&done_bar[0]), 0, nullptr, 0);
        mbar_wait(smem_u32(&done_bar[1]), 0, nullptr, 0);
      }
      if (VAR & 2) tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(t + ((i & 4) ? 256u : 0u), umma_desc_sw128(a + 32u * k), umma_desc_sw128(b + 32u * k), idesc, 1u);
      if (VAR & 4) tc_commit(smem_u32(&stage_bar[i & 3]));
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(t, 512); }
}

template <int VAR>
void run_loop(int niter, long long* d_out) {
  cudaFuncSetAttribute(loop_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  loop_kernel<VAR><<<148, 128, 200 * 1024>>>(niter, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("issuer loop variant %d (1: 2 try_waits, 2: fence, 4: commit per K-block): %.1f cycles per K-block of 4 MMAs (ideal 512)  (%s)\n",
         VAR, (double)cyc / niter, cudaGetErrorString(e));
}

// MMA rate under interference: one issuing thread streams N=256 MMAs (as stream_kernel) while 16 other warps run
// (VAR & 1) tcgen05.ld x16 loops on the OTHER accumulator region, (VAR & 2) 16-byte st.shared loops into a scratch
// block, (VAR & 4) ld.shared loops -- the epilogue's traffic classes.  Reports cycles per MMA.
template <int VAR>
__global__ void __launch_bounds__(576, 1) interf_kernel(int niter, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); stop = 0; }
  if (warp == 17) tmem_alloc(smem_u32(&tmem_ptr), 512);
  for (int i = threadIdx.x; i < 208 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp == 17) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16(128, 256, 1);
      long long t0 = clock64();
      if (VAR & 32) {
        while (clock64() - t0 < (long long)niter * 512) { }
      } else {
        for (int i = 0; i < niter; ++i) {
          const uint32_t a = base + (uint32_t)(i & 3) * 16384u, b = base + 65536u + (uint32_t)(i & 3) * 32768u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(t, umma_desc_sw128(a + 32u * k), umma_desc_sw128(b + 32u * k), idesc, 1u);
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, nullptr, 0);
      }
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    }
  } else if (warp < 16) {
    const uint32_t taddr = t + 256u + (((uint32_t)((warp & 3) * 32)) << 16) + (uint32_t)((warp >> 2) * 16);
    const uint32_t saddr = base + 196608u + (uint32_t)threadIdx.x * 16u;   // scratch: 8 KB past the operand blocks
    uint32_t acc = 0;
    long long iters = 0;
    while (!stop) {
      ++iters;
      if (VAR & 1) {
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_wait_ld();
        acc += r[0] + r[15];
      }
      if (VAR & 2) {
        st_shared_v4(saddr, acc, acc, acc, acc);
        st_shared_v4(saddr + 8192u < base + 208u * 1024u ? saddr : saddr, acc, acc, acc, acc);
      }
      if (VAR & 4) {
        uint32_t v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr));
        acc += v;
      }
      if (VAR & 8) { fence_proxy_async_smem(); }
    }
    if (acc == 0x12345678u) out[1] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[3] = iters;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) { tc_fence_after(); tmem_dealloc(t, 512); }
}

template <int VAR>
void run_interf(int niter, long long* d_out) {
  cudaFuncSetAttribute(interf_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  interf_kernel<VAR><<<148, 576, 212 * 1024>>>(niter, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0, o4[4] = {0, 0, 0, 0};
  cudaMemcpy(o4, d_out, 32, cudaMemcpyDeviceToHost);
  cyc = o4[0];
  if (VAR & 2) printf("   st.shared traffic of the 16 warps during the run: %.1f B/clk/SM\n", (double)o4[3] * 2 * 16 * 512 / (double)cyc);
  printf("interference %2d (1: LDTM, 2: st.shared, 4: ld.shared, 8: proxy fence; 16 warps): %.1f cycles/MMA (ideal 128)  (%s)\n",
         VAR, (double)cyc / (niter * 4), cudaGetErrorString(e));
}

// How far can the issuing thread run ahead?  Cycles spent INSIDE the issue loop of n MMAs (N=256) starting from an
// idle tensor pipe, before any completion wait.
__global__ void __launch_bounds__(128, 1) depth_kernel(int n, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr), 512);
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_16(128, 256, 1);
    const uint64_t a0 = umma_desc_sw128(base), b0 = umma_desc_sw128(base + 65536u);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one_sync()) {
      t0 = clock64();
      for (int i = 0; i < n; ++i) tc_mma_bf16(t, a0 + (uint64_t)(2 * (i & 3)), b0 + (uint64_t)(2 * (i & 3)), idesc, 1u);
      t1 = clock64();
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, nullptr, 0);
      t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(t, 512); }
}

// The epilogue's per-chunk chain in isolation: 16 warps loop over [VAR&1: tcgen05.ld x16 + wait] [VAR&16: 4 x ld.shared.v4
// (bias)] [math: 8 packed adds + 8 converts] [VAR&2: 2 x st.shared.v4] [VAR&4: fence.proxy.async] [VAR&8: tcgen05 fence +
// syncwarp + one mbarrier arrive per warp].  Reports cycles per iteration (= per chunk) of warp 0.
template <int VAR>
__global__ void __launch_bounds__(576, 1) chunk_kernel(int niter, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bars[4];
  __shared__ __align__(16) float bias[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 16); fence_mbar_init(); }
  if (threadIdx.x < 256) bias[threadIdx.x] = 0.5f;
  if (warp == 17) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp < 16) {
    const int q = warp & 3, cg = warp >> 2, row = q * 32 + lane;
    const uint32_t taddr = t + (((uint32_t)(q * 32)) << 16) + (uint32_t)(cg * 16);
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < niter; ++i) {
      const int kb = i & 3;
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(acc + (float)j);
      if (VAR & 1) { tmem_ld16(taddr + 64u * kb, r); tmem_wait_ld(); }
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      if (VAR & 16) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = *reinterpret_cast<const float4*>(&bias[kb * 64 + cg * 16 + 4 * j4]);
          add_f32x2(v[4 * j4], v[4 * j4 + 1], b.x, b.y);
          add_f32x2(v[4 * j4 + 2], v[4 * j4 + 3], b.z, b.w);
        }
      }
      uint32_t p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = pack_16x2<false, true>(v[2 * j], v[2 * j + 1]);
      if (VAR & 2) {
        const uint32_t x = (uint32_t)row & 7u;
        const uint32_t o0 = base + (uint32_t)kb * 16384u + (uint32_t)row * 128u + ((((uint32_t)(cg * 2)) ^ x) << 4);
        const uint32_t o1 = base + (uint32_t)kb * 16384u + (uint32_t)row * 128u + ((((uint32_t)(cg * 2 + 1)) ^ x) << 4);
        st_shared_v4(o0, p[0], p[1], p[2], p[3]);
        st_shared_v4(o1, p[4], p[5], p[6], p[7]);
      } else {
        acc += __uint_as_float(p[0] ^ p[7]);
      }
      if (VAR & 4) fence_proxy_async_smem();
      if (VAR & 8) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[kb]));
      }
      acc += v[3];
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 1234.5f) out[2] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) { tc_fence_after(); tmem_dealloc(t, 512); }
}

template <int VAR>
void run_chunk(int niter, long long* d_out) {
  cudaFuncSetAttribute(chunk_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  chunk_kernel<VAR><<<148, 576, 80 * 1024>>>(niter, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("epilogue chunk variant %2d (1: LDTM+wait, 16: bias LDS, 2: STS, 4: proxy fence, 8: fence+syncwarp+arrive): %.1f cycles per chunk  (%s)\n",
         VAR, (double)cyc / niter, cudaGetErrorString(e));
}

template <int N>
void run_stream(int grid, int niter, long long* d_out) {
  cudaFuncSetAttribute(stream_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  stream_kernel<N><<<grid, 128, 200 * 1024>>>(niter, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("streaming operands N=%3d grid=%3d: %.1f cycles/MMA, %.0f MAC/clk/SM  (%s)\n", N, grid,
         (double)cyc / (niter * 4), 128.0 * N * 16 * niter * 4 / cyc, cudaGetErrorString(e));
}

template <int N>
void run(int grid, int niter, long long* d_out) {
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  rate_kernel<N><<<grid, 128, 64 * 1024>>>(niter, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d grid=%3d: %lld cycles for %d MMAs (K=16) -> %.1f cycles/MMA, %.0f MAC/clk/SM  (%s)\n", N, grid, cyc,
         niter * 4, (double)cyc / (niter * 4), 128.0 * N * 16 * niter * 4 / cyc, cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  for (int grid : {1, 148}) {
    run<256>(grid, 2000, d_out);
    run<128>(grid, 2000, d_out);
    run<64>(grid, 2000, d_out);
  }
  for (int grid : {1, 148}) {
    run_stream<256>(grid, 2000, d_out);
    run_stream<128>(grid, 2000, d_out);
  }
  run_loop<0>(2000, d_out); run_loop<1>(2000, d_out); run_loop<2>(2000, d_out); run_loop<4>(2000, d_out);
  run_loop<3>(2000, d_out); run_loop<7>(2000, d_out);
  run_interf<0>(1000, d_out); run_interf<1>(1000, d_out); run_interf<2>(1000, d_out); run_interf<34>(1000, d_out);
  run_interf<4>(1000, d_out); run_interf<8>(1000, d_out); run_interf<3>(1000, d_out); run_interf<15>(1000, d_out);
  for (int n : {1, 2, 3, 4, 6, 8, 12, 16, 32}) {
    cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    depth_kernel<<<148, 128, 200 * 1024>>>(n, d_out);
    cudaDeviceSynchronize();
    long long c2[2] = {0, 0};
    cudaMemcpy(c2, d_out, 16, cudaMemcpyDeviceToHost);
    printf("issue depth: %2d MMAs issued in %5lld cycles, all complete after %5lld cycles (ideal %d)\n", n, c2[0], c2[1], n * 128);
  }
  run_chunk<0>(2000, d_out); run_chunk<1>(2000, d_out); run_chunk<17>(2000, d_out); run_chunk<19>(2000, d_out);
  run_chunk<23>(2000, d_out); run_chunk<31>(2000, d_out); run_chunk<27>(2000, d_out); run_chunk<30>(2000, d_out);
  // commit -> mbarrier latency: 0, 1, 2, 4 groups of 4 MMAs (N=256: 512 cycles per group) then commit + wait
  for (int n : {0, 1, 2, 4, 8}) {
    cudaFuncSetAttribute(rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    rate_kernel<256><<<148, 128, 64 * 1024>>>(n, d_out);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("latency: %d MMA groups (ideal %d cycles) + commit + wait = %lld cycles\n", n, n * 512, cyc);
  }
  return 0;
}
