// Device helpers shared by the fp32 (CUDA-core) MLP kernels: mlp_f32.cu (vanilla ResNet-FC NeRF) and mip_f32.cu
// (mip-NeRF field).  One CTA = one tile of 64 samples, 256 threads; activations live in shared memory transposed
// ([feature][sample], row stride 68); weights are streamed from L2 in 16-deep K slabs with cp.async double buffering;
// every thread owns an 8 (samples) x 8 (features) register tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TM 64
#define AS 68
#define KS 16
#define A_ROWS 288
#define NTHREADS 256

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Column owned by register slot nj of this lane: lane*4 + (nj&3) + 128*(nj>>2)
__device__ __forceinline__ int col_of(int lane, int nj) { return lane * 4 + (nj & 3) + 128 * (nj >> 2); }

// acc[8][4*NJ] += As^T[64 x K] * Wg[K x 128*NJ]   (As: smem [K][AS], Wg: global row-major, ld = ldw)
template <int NJ>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ As, const float* __restrict__ Wg, int ldw,
                                          int K, float (&acc)[8][4 * NJ], float* __restrict__ Wbuf, int tid) {
  constexpr int NCOL = 128 * NJ;
  constexpr int CHUNKS = KS * NCOL / 4;          // float4 chunks per slab
  const int warp = tid >> 5, lane = tid & 31;
  const int nslab = K / KS;
  auto load_slab = [&](int s, int buf) {
    float* dst = Wbuf + buf * (KS * 256);
    const float* src = Wg + (int64_t)s * KS * ldw;
#pragma unroll
    for (int c = tid; c < CHUNKS; c += NTHREADS) {
      const int row = c / (NCOL / 4), col4 = c % (NCOL / 4);
      cp_async16(dst + row * NCOL + col4 * 4, src + (int64_t)row * ldw + col4 * 4);
    }
    cp_async_commit();
  };
  load_slab(0, 0);
  for (int s = 0; s < nslab; ++s) {
    if (s + 1 < nslab) {
      load_slab(s + 1, (s + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Wb = Wbuf + (s & 1) * (KS * 256);
    const float* Ab = As + (s * KS) * AS + warp * 8;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(Ab + kk * AS);
      const float4 a1 = *reinterpret_cast<const float4*>(Ab + kk * AS + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[4 * NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float4 bv = *reinterpret_cast<const float4*>(Wb + kk * NCOL + lane * 4 + 128 * j);
        b[4 * j + 0] = bv.x; b[4 * j + 1] = bv.y; b[4 * j + 2] = bv.z; b[4 * j + 3] = bv.w;
      }
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4 * NJ; ++nj) acc[mi][nj] = fmaf(a[mi], b[nj], acc[mi][nj]);
    }
    __syncthreads();
  }
}


// dW[n][k] (row stride ldw) += G^T In for one GEMM layer or one K-segment of it; db (may be NULL) += column sums of G
int star_f32_dw(const float* G, int N, const float* In, int Kpad, int K, int64_t M, float* dW, int ldw, float* db,
                cudaStream_t st);
