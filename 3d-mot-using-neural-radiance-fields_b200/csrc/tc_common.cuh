// sm_100a primitives used by the tensor-core MLP kernels: mbarrier, 1-D bulk async copy (TMA engine,
// SASS UBLKCP), tcgen05 MMA / TMEM load / commit / alloc, UMMA shared-memory and instruction descriptors.
// Inline PTX only; nothing here is portable to other architectures.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

// Debug watchdog for mbarrier waits: a wait that lasts longer than this many clock64 ticks records
// (code, blockIdx) in the caller-supplied debug word and traps instead of hanging the GPU.
#ifndef STAR_TC_WATCHDOG_CYCLES
#define STAR_TC_WATCHDOG_CYCLES (4000000000LL)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TC_WAIT_HINT_NS > 0: try_wait carries a suspend-time hint (the thread sleeps until the phase completes or the hint
// expires instead of returning after the implementation's default time limit): fewer polling instructions while the 16
// epilogue warps wait for an accumulator.  Measured on the power-capped C2 render (tools/ab_bench.sh, one box, interleaved):
// 1 us and 20 us hints are 0.5 % SLOWER than the default (2.777 / 2.778 M against 2.791 M rays/s, clock unchanged) -> off.
#ifndef TC_WAIT_HINT_NS
#define TC_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
#if TC_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"((uint32_t)TC_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* dbg, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > STAR_TC_WATCHDOG_CYCLES) {
      if (dbg != nullptr) {      // (mapped host memory: make sure the word has landed before the context dies)
        atomicExch(dbg, (code << 16) | (int)(blockIdx.x & 0xffff));
        __threadfence_system();
        if (*reinterpret_cast<volatile int*>(dbg) == 0) __threadfence_system();
      }
      __trap();
    }
  }
}

// launch markers next to a family's watchdog word (mapped host memory): word + 16 counts kernel starts, word + 32 kernel ends
// (CTA 0 only) -- after a CUDA fault, begin != end names the family that was running (star_watchdog_word(16 + f) / (32 + f))
__device__ __forceinline__ void tc_mark_begin(int* dbg) {
  if (dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(dbg + 16, 1);
}
__device__ __forceinline__ void tc_mark_end(int* dbg) {
  if (dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(dbg + 32, 1);
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// global -> shared bulk copy on the TMA engine; completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// same, multicast to the CTAs of the cluster selected by cta_mask (same smem / barrier offsets in each)
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}

// shared -> global bulk copy (TMA engine), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst_global, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the most recently committed group have read their source (groups complete in order)
__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all committed groups are complete (global writes performed)
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M x N x 16 per instruction
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- CTA pairs (cluster of 2, cta_group::2): ONE tcgen05.mma of the leader CTA drives both SMs' tensor cores: M = 256
// (each CTA's own 128 rows of A from its own shared memory, results in its own TMEM), B split by rows -- the leader holds
// rows [0, N/2), the peer rows [N/2, N) of the [N][64] K-block at the same shared-memory offset (tools/umma_probe.cu
// checks this on the hardware: 128 cycles per 256 x 256 x 16 MMA).  Per SM this halves the weight bytes fetched from
// L2, written to and read from shared memory.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the leader's MMAs have completed
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrive (release, cluster scope) on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait with cluster-scope acquire: pairs with mbar_arrive_remote (the peer's shared-memory writes become visible)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int* dbg, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > STAR_TC_WATCHDOG_CYCLES) {
      if (dbg != nullptr) {      // (mapped host memory: make sure the word has landed before the context dies)
        atomicExch(dbg, (code << 16) | (int)(blockIdx.x & 0xffff));
        __threadfence_system();
        if (*reinterpret_cast<volatile int*>(dbg) == 0) __threadfence_system();
      }
      __trap();
    }
  }
}
template <int NK>
__device__ __forceinline__ void tc_mma2_kblock(uint32_t d_tmem, uint64_t a_desc0, uint64_t b_desc0, uint32_t idesc,
                                               uint32_t acc0) {
#pragma unroll
  for (int k = 0; k < NK; ++k)
    tc_mma2_bf16(d_tmem, a_desc0 + (uint64_t)(2 * k), b_desc0 + (uint64_t)(2 * k), idesc, k == 0 ? acc0 : 1u);
}

// One lane of a fully active warp (elect.sync).  The MMA-issuing warp runs its control flow on all 32 lanes and guards
// only the tcgen05 instructions with this predicate: behind a plain `lane == 0` branch the compiler has to treat the
// uniform-datapath operands of UTCHMMA / UTCBAR as divergent and wraps every one of them in a vote / elect / branch loop
// plus vector->uniform register moves, which made the issuer thread (~120 dependent instructions per K-block) slower
// than the tensor core it feeds.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// 4 (or NK) K-steps of one K-block: descriptors advance by 32 bytes (+2 in the 16-byte address field) per K-step.
// ACC0: accumulate flag of the first K-step (the others always accumulate).
template <int NK>
__device__ __forceinline__ void tc_mma_kblock(uint32_t d_tmem, uint64_t a_desc0, uint64_t b_desc0, uint32_t idesc,
                                              uint32_t acc0) {
#pragma unroll
  for (int k = 0; k < NK; ++k)
    tc_mma_bf16(d_tmem, a_desc0 + (uint64_t)(2 * k), b_desc0 + (uint64_t)(2 * k), idesc, k == 0 ? acc0 : 1u);
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base + t), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row
// groups 1024 B apart (SBO); the tile base must be 1024-byte aligned; advancing K by 16 elements inside the
// 64-wide swizzle atom = +32 bytes on the start address.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)  [46,48)
  d |= (uint64_t)2 << 61;                    // layout type SWIZZLE_128B  [61,64)
  return d;
}
// instruction descriptor for kind::f16: A and B both bf16 (fmt = 1) or both fp16 (fmt = 0), K-major,
// D = fp32, shape M x N
__host__ __device__ constexpr uint32_t umma_idesc_16(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// (A and B must share the format: kind::f16 with a_format != b_format -- bf16 x fp16 -- raises an illegal-instruction
//  fault on B200, tools/umma_probe.cu)

// byte offset of element (row, k) inside one [rows][64] bf16 SW128 K-block
__host__ __device__ __forceinline__ uint32_t sw128_off(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

// two fp32 -> packed 16-bit pair (lo in the low half); FP16 selects IEEE half instead of bfloat16;
// RELU fuses max(x, 0) into the conversion (cvt.rn.relu)
// SAT (fp16 only): saturate to +-65504 instead of producing inf (F2FP.SATFINITE: no extra instruction)
template <bool FP16, bool RELU = false, bool SAT = false>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
  uint32_t r;
  if (FP16 && SAT) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else if (FP16) {
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}
// packed fp32x2 add (one issue slot for two lanes-worth of adds on sm_100): (x0, x1) += (b0, b1)
__device__ __forceinline__ void add_f32x2(float& x0, float& x1, float b0, float b1) {
  uint64_t a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(c));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
