// Hardware probes for two tcgen05 questions the MLP kernels depend on (run on a B200; prints PASS / FAIL lines):
//   1. mixed operand formats inside kind::f16: A = bf16 (a_format 1), B = fp16 (b_format 0) -- the gradient GEMMs of the
//      fp16 tier multiply bf16 gradients (range) with fp16 weights / activations (precision);
//   2. cta_group::2 with shared-memory operands: a CTA pair issues ONE M = 256 MMA; each CTA supplies its own 128 rows of
//      A and HALF of the N rows of B (rows [0, N/2) in the leader, [N/2, N) in the peer) at identical smem offsets, and
//      receives its 128 x N accumulator in its own TMEM; tcgen05.commit multicast arrives on both CTAs' barriers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../3d-mot-using-neural-radiance-fields_b200/csrc/tc_common.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__host__ __device__ constexpr uint32_t idesc_ab(int M, int N, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ probe 1
// A [128][64] (a_fmt), B [N][64] (b_fmt), both K-major SW128 images prepared by the host; D [128][N] fp32.
__global__ void __launch_bounds__(128, 1) mixed_kernel(const uint8_t* a_img, const uint8_t* b_img, int N, int afmt, int bfmt,
                                                        float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr), 256);
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (int i = threadIdx.x; i < N * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(g + 16384)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = idesc_ab(128, N, afmt, bfmt);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(t, umma_desc_sw128(base) + (uint64_t)(2 * k), umma_desc_sw128(base + 16384) + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
      tc_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0, nullptr, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tmem_ld16(t + (((uint32_t)(warp * 32)) << 16) + (uint32_t)c, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) out[row * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(t, 256); }
}

// ------------------------------------------------------------------------------------------------ probe 2
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// a_img: [2 CTAs][128][64] SW128 images; b_img: [N][64] as N/128... stored as two halves of N/2 rows each (SW128 image per
// half); out: [2][128][N].  niter > 1: timing loop (accumulating), cycles -> cyc[0].
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const uint8_t* a_img, const uint8_t* b_img, int N, float* out, int niter, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar_done, bar_ready;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int half_bytes = (N / 2) * 128;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    mbar_init(smem_u32(&bar_ready), 2);     // leader: own operands staged + the peer's
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(a_img + (size_t)rank * 16384)[i];
  for (int i = threadIdx.x; i < half_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(g + 16384)[i] = reinterpret_cast<const uint4*>(b_img + (size_t)rank * half_bytes)[i];
  fence_proxy_async_smem();
  cluster_sync_all();                        // barrier inits visible cluster-wide before anyone arrives remotely
  if (warp == 1) tmem_alloc2(smem_u32(&tmem_ptr), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_ptr;
  if (threadIdx.x == 0) mbar_arrive_remote(smem_u32(&bar_ready), 0);   // "my operands are in shared memory" -> leader
  if (warp == 0 && rank == 0) {
    mbar_wait(smem_u32(&bar_ready), 0, nullptr, 0);
    tc_fence_after();
    const long long t0 = clock64();
    if (elect_one_sync()) {
      const uint32_t idesc = idesc_ab(256, N, 1, 1);
      for (int it = 0; it < niter; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma2(t, umma_desc_sw128(base) + (uint64_t)(2 * k), umma_desc_sw128(base + 16384) + (uint64_t)(2 * k), idesc,
                  (k || it) ? 1u : 0u);
      }
      tc_commit2(smem_u32(&bar_done), 3);
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar_done), 0, nullptr, 0);
    if (lane == 0 && cyc) cyc[0] = clock64() - t0;
  }
  mbar_wait(smem_u32(&bar_done), 0, nullptr, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tmem_ld16(t + (((uint32_t)(warp * 32)) << 16) + (uint32_t)c, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + row) * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc2(t, 256); }
}

// ------------------------------------------------------------------------------------------------ host
static uint16_t to16(float v, int fmt) {   // fmt 1 = bf16, 0 = fp16
  if (fmt == 1) { __nv_bfloat16 b = __float2bfloat16(v); return *reinterpret_cast<uint16_t*>(&b); }
  __half h = __float2half(v); return *reinterpret_cast<uint16_t*>(&h);
}
static float from16(uint16_t u, int fmt) {
  if (fmt == 1) { __nv_bfloat16 b; *reinterpret_cast<uint16_t*>(&b) = u; return __bfloat162float(b); }
  __half h; *reinterpret_cast<uint16_t*>(&h) = u; return __half2float(h);
}
// rows x 64 matrix -> SW128 K-major image
static void make_img(const std::vector<float>& m, int rows, int fmt, std::vector<uint8_t>& img, std::vector<float>& rounded) {
  img.assign((size_t)rows * 128, 0);
  rounded.resize((size_t)rows * 64);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < 64; ++k) {
      const uint16_t u = to16(m[(size_t)r * 64 + k], fmt);
      rounded[(size_t)r * 64 + k] = from16(u, fmt);
      *reinterpret_cast<uint16_t*>(&img[sw128_off(r, k)]) = u;
    }
}

int main(int argc, char** argv) {
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  // ---- probe 1 (cfg 0 = bf16 x bf16 sanity; cfg 1, 2 = mixed formats, only with --mixed: on B200 they raise an
  // illegal-instruction fault, which kills the context -- measured 2026-10-18)
  const bool mixed = argc > 1 && std::string(argv[1]) == "--mixed";
  for (int cfg = 0; cfg < (mixed ? 3 : 1); ++cfg) {
    const int afmt = cfg == 0 ? 1 : (cfg == 1 ? 1 : 0), bfmt = cfg == 0 ? 1 : (cfg == 1 ? 0 : 1), N = 64;
    std::vector<float> A(128 * 64), B((size_t)N * 64), Ar, Br;
    for (auto& v : A) v = rnd() * (afmt == 1 ? 1e-6f : 1.f);      // bf16 operand: tiny values that fp16 would flush
    for (auto& v : B) v = rnd() * (bfmt == 1 ? 1e-6f : 1.f);
    std::vector<uint8_t> ai, bi;
    make_img(A, 128, afmt, ai, Ar);
    make_img(B, N, bfmt, bi, Br);
    uint8_t *da, *db; float* dout;
    CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dout, 128 * N * 4));
    CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    mixed_kernel<<<1, 128, 64 * 1024>>>(da, db, N, afmt, bfmt, dout);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * N);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxrel = 0, scale = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)Ar[r * 64 + k] * Br[n * 64 + k];
        scale = fmax(scale, fabs(ref));
        maxrel = fmax(maxrel, fabs(out[r * N + n] - ref));
      }
    printf("probe1 a_fmt=%s b_fmt=%s: max abs err %.3e of scale %.3e -> %s\n", afmt ? "bf16" : "fp16", bfmt ? "bf16" : "fp16",
           maxrel, scale, maxrel < 1e-5 * scale + 1e-30 ? "PASS" : "FAIL");
    cudaFree(da); cudaFree(db); cudaFree(dout);
  }
  // ---- probe 2
  for (int N : {256, 128}) {
    std::vector<float> A(256 * 64), B((size_t)N * 64), Ar, Br;
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    std::vector<uint8_t> ai(2 * 16384), bi((size_t)N * 128), tmp;
    std::vector<float> r0, r1;
    for (int c = 0; c < 2; ++c) {
      std::vector<float> sub(A.begin() + c * 128 * 64, A.begin() + (c + 1) * 128 * 64), rr;
      make_img(sub, 128, 1, tmp, rr);
      memcpy(&ai[c * 16384], tmp.data(), 16384);
      Ar.insert(Ar.end(), rr.begin(), rr.end());
      std::vector<float> subb(B.begin() + (size_t)c * (N / 2) * 64, B.begin() + (size_t)(c + 1) * (N / 2) * 64);
      make_img(subb, N / 2, 1, tmp, rr);
      memcpy(&bi[(size_t)c * (N / 2) * 128], tmp.data(), (size_t)(N / 2) * 128);
      Br.insert(Br.end(), rr.begin(), rr.end());
    }
    uint8_t *da, *db; float* dout; long long* dcyc;
    CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dout, 256 * N * 4)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, 256 * N * 4));
    CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    pair_kernel<<<2, 128, 64 * 1024>>>(da, db, N, dout, 1, dcyc);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(256 * N);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, scale = 0;
    for (int r = 0; r < 256; ++r)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)Ar[r * 64 + k] * Br[(size_t)n * 64 + k];
        scale = fmax(scale, fabs(ref));
        maxerr = fmax(maxerr, fabs(out[(size_t)r * N + n] - ref));
      }
    printf("probe2 cta_group::2 M=256 N=%d: max abs err %.3e of scale %.3e -> %s\n", N, maxerr, scale,
           maxerr < 1e-5 * scale ? "PASS" : "FAIL");
    // issue rate: 256 K-blocks of 4 MMAs
    pair_kernel<<<2, 128, 64 * 1024>>>(da, db, N, dout, 256, dcyc);
    CK(cudaDeviceSynchronize());
    long long cyc;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("probe2 N=%d: %.1f cycles per M=256 MMA (1024 MMAs back to back)\n", N, (double)cyc / 1024.0);
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dcyc);
  }
  return 0;
}
