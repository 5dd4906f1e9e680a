// 16-bit (bf16 / fp16 operands) tensor-core tiers of the per-sample NeRF MLP (models/nerf.py:112-179,
// models/resnet.py:51-59,103-110) for sm_100a: pose transform (star__.py:160-199) -> positional encoding
// (embedder.py:81-112) -> ResNet-FC trunk -> heads as ONE persistent, warp-specialised kernel per launch.
//
// One CTA per SM walks tiles of 128 samples.  Roles (576 threads):
//   warps 0-15 "epilogue" warps: encode the tile's inputs into shared memory, and after every layer move the fp32
//              accumulator TMEM -> registers, add bias / ReLU, convert to 16 bits and write the next layer's A operand
//              (warp w owns TMEM lanes 32*(w%4).., 16 of every 64 columns: group w/4); alpha / rgb heads are fp32
//              dot products in these registers.
//   warp 16    weight producer: streams the pre-swizzled weight K-blocks (32 KB each) from L2 into a 4-stage
//              shared-memory ring with 1-D bulk async copies (TMA engine) + mbarrier transaction counts.
//   warp 17    MMA issuer: one thread issues tcgen05.mma (M=128, N=256|128, K=16, 16-bit operands -> fp32 in TMEM).
//
// TMEM (512 columns): X = columns 0..255 holds the residual stream x (fp32, biases kept separately),
// T = columns 256..511 holds the other accumulator.  fc_1 ACCUMULATES onto X, which performs the residual
// add inside the tensor core.  Layers alternate X/T, so the epilogue of layer L (reading one region and
// producing A K-block by K-block) overlaps the MMA of layer L+1 (writing the other region, consuming A
// K-block by K-block).  An operand block of layer L+1 is announced on the SAME barrier as the weight K-block it will
// be multiplied with: w_full[stage] expects 1 arrival + the copy's bytes from the producer and 16 arrivals from the
// epilogue warps (count 17), so the issuer -- whose loop must stay under the 512 tensor cycles of a K-block -- waits
// on one barrier per K-block instead of two.
// RULE for that shared barrier: the epilogue warps may arrive on w_full[stage] only when the stage's previous phase is
// certain to be complete, i.e. when the stage's previous user (4 ring slots earlier) belongs to a layer whose MMAs are
// all done.  That holds for K-blocks 0..3 of a layer (an epilogue runs after its layer's accumulator is complete), NOT
// for a 5th K-block (it wraps onto the stage of the same layer's K-block 0) and NOT for operands written ahead of time.
// Those get barriers of their own and the producer thread supplies all 17 arrivals of their K-block's w_full, in program
// order: the encoded-dirs block (a_ready[4]; the view layer's LAST K-block) and lin_in's operand of the next tile
// (s_ready).  Breaking the rule is an arrival-count underflow = a launch failure on some GPUs (profiles/r2h_ring_race.md).
// (Training additionally arrives on a_ready[kb] for the stash writer.)
#include "star_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "tc_common.cuh"
#include "mlp_tc_layout.h"

#include "mlp_tc_device.cuh"

// Training (STASH) variant: who moves a finished A block to the activation stash?
//   0 (default) the weight producer thread, with one bulk store (cp.async.bulk shared -> global) per block;
//   1 a COPY WARP of its own (warp 18; the kernel then runs 19 warps): ld.shared.v4 / st.global.v4, 512 contiguous bytes per
//     warp instruction, release of the block (stash_done) as soon as its loads have landed.
// Measured (tools/bulk_store_rate.cu, profiles/r2v_bulk_store_rate.txt): the bulk-copy engine of an SM drains a 16 KB store at
// ~25 B/clk (644 cycles per block even with ONE SM active; 148 SMs together = the 7.3 TB/s HBM write peak), 825 cycles per
// block when the same engine also fetches a 32 KB weight K-block -- 4 x 825 cycles of engine time per layer, and weight loads
// queue behind stores (no-stores debug mode: -6.4 k of 62.8 k cycles per tile).  The copy warp takes the stores off that engine
// and is bit-identical (GPU suite), but SLOWER: 68.0 k cycles per tile (its copies cost 10 k: the loads / stores share the
// LSU path with the 16 epilogue warps, the kernel's critical path), stash forward 2.14 against 2.09 ms per 4096-ray step
// (profiles/r2w_ab_copy_warp.txt, r2w_stash_debug_modes.txt).  Kept as a compile-time A/B only.  Also measured: requesting the
// weight K-block BEFORE the store of the same iteration (so that it cannot queue behind it in the engine): 62.8 k cycles per
// tile either way (profiles/r2x_ab_store_after_load.txt) -- the order of the requests is not what the stores cost.
#ifndef TC_STASH_COPY_WARP
#define TC_STASH_COPY_WARP 0
#endif
// Inference variant: two CTAs of a cluster SHARE the weight stream (TC_WSHARE = 1): each fetches half of every weight K-block
// from L2 and multicasts it into both CTAs' ring stage, so the L2 -> SM weight traffic (1.6 MB per tile and SM, 9.4 TB/s over
// the chip) halves.  Everything else stays per CTA (own tiles, own MMAs / TMEM / epilogue -- no cta_group::2, no cross-CTA
// hand-off on the layer chain): the only coupling is the ring itself -- a stage is refilled when BOTH CTAs' MMAs have released
// it (w_empty counts 2: every issuer's commit is multicast to both CTAs).
// Why the shared ring is race-free (the argument the barrier counts rest on):
//  * A stage use n of CTA A is loaded only after A's w_empty[stage] saw BOTH commits of use n - 1.  The peer's commit follows
//    the peer's MMAs of use n - 1, which waited for the peer's w_full[stage] phase n - 1: so when A's half of use n lands in the
//    peer (bytes + complete_tx on the peer's w_full[stage]) the peer's phase n - 1 is complete and the bytes count into phase n
//    -- possibly before the peer's own expect_tx of phase n (a negative transaction count is legal; the phase cannot complete
//    before the peer's producer arrival), and the peer's MMAs of use n - 1 no longer read the stage.  The same holds for the
//    uses whose w_full arrivals all come from the producer (slot S, the dirs K-block): the producer posts them, in program
//    order, before it requests the next weights, so a later commit of the peer implies they have been posted.
//  * w_empty[stage] receives exactly two commits per use (both CTAs run the same layer program over the same number of
//    tiles); use n + 1 cannot be committed by either CTA before BOTH producers have observed phase n (its weights are
//    requested after that observation), so no waiter -- the producer, or the epilogue warps that own slot S -- can miss a phase.
//  * Cluster barriers after the mbarrier initialisation and before exit keep the peer's copies / commits inside the lifetime
//    of the shared memory they target.
#ifndef TC_WSHARE
#define TC_WSHARE 1      // A/B on one box (profiles/r2zz_ab_wshare.txt): C2 render 2.808 -> 2.880 M rays/s, 1668 -> 1684 MHz under the power cap
#endif
// the same for the training (stash) forward -- A/B switch (the stash is sized for an even number of tiles, star_stash_bytes, so a
// ghost tile's blocks land in the spare one)
#ifndef TC_WSHARE_STASH
#define TC_WSHARE_STASH 1      // A/B (profiles/r2zz_ab_wshare_stash.txt): stash forward 2.085 -> 2.027 ms per 4096-ray step
#endif
#define TC_THREADS_STASH (TC_THREADS + (TC_STASH_COPY_WARP ? 32 : 0))

#ifndef TC_LD_DEPTH
#define TC_LD_DEPTH 2        // TMEM chunk loads in flight per epilogue thread (measured: 2 beats 1 and 4)
#endif

// ---------------------------------------------------------------------------------------------- epilogue
struct EpiCtx {
  uint32_t sA, a_ready0;       // smem address of A K-block 0, of barrier a_ready[0]
  uint32_t tcol;               // TMEM address: lane base of this warp | accumulator region | 16 * column group
  const float* bias;           // smem, epilogue bias vector of this layer
  const float* head_w;         // smem, alpha_linear / rgb_linear weights (OUT / VIEWS layers)
  uint8_t* stash_out;          // global: first stash block of this layer's epilogue output for this tile, or NULL
  uint8_t* mask_out;           // global: this layer's ReLU bit masks for this tile (training), or NULL
  bool no_mask;                // debug (timing experiments only)
  uint32_t stash_done0;        // smem address of barrier stash_done[0]
  bool publishes;              // training: this layer's output blocks are bulk-stored to the stash afterwards
  bool no_stash_wait;          // debug (timing experiments only)
  uint32_t w_full0;            // smem address of barrier w_full[0]
  uint32_t next_stage0;        // ring stage of the NEXT layer's K-block 0 (its K-block kb uses (next_stage0 + kb) % NS)
  uint32_t ns_mask;            // NS - 1
  int row, cg, lane;
};

// One layer's accumulator -> next layer's A operand for this thread's row: per K-block kb, columns
// [64 kb + 16 cg, +16): TMEM -> registers (double buffered) -> + bias -> (ReLU fused into the 16-bit
// conversion) -> swizzled smem -> proxy fence -> one mbarrier arrival per warp.
// KIND: LK_IN / LK_FC0 / LK_FC1 (ReLU), LK_OUT (affine + alpha head partial in h[0]), LK_FEAT (affine),
// LK_VIEWS (ReLU, N = 128, rgb head partials in h[0..2], no A output).
// Stash bookkeeping (training): bit kb of `pend` = a bulk store that READS A block kb has been issued
// by the producer warp since this thread last waited for that block (every block strictly alternates "written and
// published" -> "stored" -> "waited for" -> "overwritten"; the layer program is static, so the flag needs no
// communication); `par` holds the parity of the next wait per block.
template <int KIND, bool FP16, bool STASH>
__device__ __forceinline__ void epilogue_layer(const EpiCtx& c, float (&h)[3], uint32_t& stash_par, uint32_t& stash_pend) {
  constexpr int NCH = (KIND == LK_VIEWS) ? 2 : 4;
  constexpr bool RELU = (KIND == LK_IN || KIND == LK_FC0 || KIND == LK_FC1);
  // All TMEM loads of the layer are issued up front (TC_LD_DEPTH chunks in flight): a tcgen05.ld takes several hundred
  // cycles to return, which with one chunk of look-ahead made every chunk of the serial epilogue chain as long as
  // that latency.
  uint32_t r[NCH][TC_CPT];
  uint32_t mlo = 0u, mhi = 0u;   // training: one bit per rectified output of this row (dX pass of the backward)
#pragma unroll
  for (int kb = 0; kb < NCH && kb < TC_LD_DEPTH; ++kb) tmem_ld16(c.tcol + 64u * (uint32_t)kb, r[kb]);
#pragma unroll
  for (int kb = 0; kb < NCH; ++kb) {
    const int col0 = kb * 64 + c.cg * TC_CPT;
    float4 bq[TC_CPT / 4];
#pragma unroll
    for (int j4 = 0; j4 < TC_CPT / 4; ++j4) bq[j4] = *reinterpret_cast<const float4*>(c.bias + col0 + 4 * j4);
    // tcgen05.wait::ld waits for every outstanding load of the thread: with the whole layer in flight there is one
    // wait (before chunk 0); with a shallower depth the next load is issued after each wait
    if (TC_LD_DEPTH >= NCH) {
      if (kb == 0) tmem_wait_ld();
    } else {
      tmem_wait_ld();
      if (kb + TC_LD_DEPTH < NCH) tmem_ld16(c.tcol + 64u * (uint32_t)(kb + TC_LD_DEPTH), r[kb + TC_LD_DEPTH]);
    }
    float v[TC_CPT];
#pragma unroll
    for (int j4 = 0; j4 < TC_CPT / 4; ++j4) {
      v[4 * j4 + 0] = __uint_as_float(r[kb][4 * j4 + 0]);
      v[4 * j4 + 1] = __uint_as_float(r[kb][4 * j4 + 1]);
      v[4 * j4 + 2] = __uint_as_float(r[kb][4 * j4 + 2]);
      v[4 * j4 + 3] = __uint_as_float(r[kb][4 * j4 + 3]);
      add_f32x2(v[4 * j4 + 0], v[4 * j4 + 1], bq[j4].x, bq[j4].y);
      add_f32x2(v[4 * j4 + 2], v[4 * j4 + 3], bq[j4].z, bq[j4].w);
    }
    if (KIND == LK_VIEWS) {          // rgb head (nerf.py:159) on the rectified fp32 h2: 4 independent partial sums per channel
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j4 = 0; j4 < TC_CPT / 4; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(c.head_w + ch * STAR_WV + col0 + 4 * j4);
          a[0] = fmaf(fmaxf(v[4 * j4 + 0], 0.f), w.x, a[0]);
          a[1] = fmaf(fmaxf(v[4 * j4 + 1], 0.f), w.y, a[1]);
          a[2] = fmaf(fmaxf(v[4 * j4 + 2], 0.f), w.z, a[2]);
          a[3] = fmaf(fmaxf(v[4 * j4 + 3], 0.f), w.w, a[3]);
        }
        h[ch] += (a[0] + a[1]) + (a[2] + a[3]);
      }
      if (STASH) store_row16<FP16, true>(0u, c.row, c.cg * 2, v, c.stash_out + kb * TC_KB_BYTES);
    } else {
      // (training: the stash copy of this block is a bulk store issued by the producer warp once the block is complete;
      //  the previous layer's copy of block kb must have left shared memory before it is overwritten)
      if (STASH) {
        const uint32_t bit = 1u << kb;
        if (stash_pend & bit) {
          if (!c.no_stash_wait) mbar_wait(c.stash_done0 + 8u * (uint32_t)kb, (stash_par & bit) ? 1u : 0u, nullptr, 7);
          stash_par ^= bit;
          stash_pend &= ~bit;
        }
        if (c.publishes) stash_pend |= bit;
      }
      store_row16<FP16, RELU>(c.sA + (uint32_t)kb * TC_KB_BYTES, c.row, c.cg * 2, v);
      if (STASH && RELU && !c.no_mask) {
        uint32_t b = 0u;
#pragma unroll
        for (int j = 0; j < TC_CPT; ++j) b |= (v[j] > 0.f ? 1u : 0u) << j;
        if (kb & 1) b <<= 16;
        if (kb < 2) mlo |= b; else mhi |= b;
      }
      fence_proxy_async_smem();      // this thread's A writes -> async proxy (tcgen05.mma operand reads)
      tc_fence_before();
      __syncwarp();
      if (c.lane == 0) {
        mbar_arrive(c.w_full0 + 8u * ((c.next_stage0 + (uint32_t)kb) & c.ns_mask));
        if (STASH) mbar_arrive(c.a_ready0 + 8u * (uint32_t)kb);
      }
      if (KIND == LK_OUT) {          // alpha head (nerf.py:151) on the fp32 h -- AFTER the block is published: off the
        float a[4] = {0.f, 0.f, 0.f, 0.f};   // path the next layer's MMAs wait on; 4 independent partial sums
#pragma unroll
        for (int j4 = 0; j4 < TC_CPT / 4; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(c.head_w + col0 + 4 * j4);
          a[0] = fmaf(v[4 * j4 + 0], w.x, a[0]);
          a[1] = fmaf(v[4 * j4 + 1], w.y, a[1]);
          a[2] = fmaf(v[4 * j4 + 2], w.z, a[2]);
          a[3] = fmaf(v[4 * j4 + 3], w.w, a[3]);
        }
        h[0] += (a[0] + a[1]) + (a[2] + a[3]);
      }
    }
  }
  if (STASH && RELU && !c.no_mask)
    *reinterpret_cast<uint2*>(c.mask_out + ((uint32_t)(c.cg * TC_M + c.row) << 3)) = make_uint2(mlo, mhi);
}

// ---------------------------------------------------------------------------------------------- input encoders
// Deliberately NOT inlined: each runs once per tile, and inlined into the layer loop their sincosf slow paths and the
// registers they need made the compiler re-materialise addresses in every epilogue chunk (35 -> 55 instructions per chunk,
// measured: 2.84 -> 2.73 M rays/s).  As calls they cost a few register saves per tile.
struct TcSample { float p[3]; float dv[3]; };
// pose transform of sample gi and of its ray direction (star__.py:160-199); zeros beyond the launch
__device__ __noinline__ TcSample tc_load_sample(StarPtsSrc pts, const float* __restrict__ viewdirs,
                                                const float* __restrict__ pose12, int64_t gi, int64_t M, int S) {
  TcSample o;
  o.p[0] = o.p[1] = o.p[2] = 0.f;
  o.dv[0] = o.dv[1] = o.dv[2] = 0.f;
  if (gi < M) {
    const int64_t r = gi / S;
    float px, py, pz;
    star_load_pt(pts, gi, r, px, py, pz);
    const float dx = viewdirs[r * 3 + 0], dy = viewdirs[r * 3 + 1], dz = viewdirs[r * 3 + 2];
    if (pose12 != nullptr) {   // p' = R p + t, d' = R d
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        o.p[i] = pose12[i * 3 + 0] * px + pose12[i * 3 + 1] * py + pose12[i * 3 + 2] * pz + pose12[9 + i];
        o.dv[i] = pose12[i * 3 + 0] * dx + pose12[i * 3 + 1] * dy + pose12[i * 3 + 2] * dz;
      }
    } else {
      o.p[0] = px; o.p[1] = py; o.p[2] = pz;
      o.dv[0] = dx; o.dv[1] = dy; o.dv[2] = dz;
    }
  }
  return o;
}
// 16 columns [16 cg, 16 cg + 16) of the encoded xyz of one row -> chunks 2 cg, 2 cg + 1 of the row in the SW128 block at
// shared address `dst` (and the same chunks of the stash block `gblock`, training)
template <bool FP16>
__device__ __noinline__ void tc_encode_xyz(float px, float py, float pz, const float* __restrict__ sc_xyz, int cg, int row,
                                           uint32_t dst, uint8_t* gblock) {
  const float p[3] = {px, py, pz};
  float e[16];
  if (cg == 0) encode_slice<0, 63>(p, sc_xyz, e);
  else if (cg == 1) encode_slice<16, 63>(p, sc_xyz, e);
  else if (cg == 2) encode_slice<32, 63>(p, sc_xyz, e);
  else encode_slice<48, 63>(p, sc_xyz, e);
  store_row16<FP16, false>(dst, row, cg * 2, e, gblock);
}
// the encoded ray direction of one row (27 columns: column groups 1 and 2 write 16 each) into the dirs block
template <bool FP16, bool STASH>
__device__ __noinline__ void tc_encode_dirs(float dx, float dy, float dz, const float* __restrict__ sc_dir, int cg, int row,
                                            uint32_t dst, uint8_t* gblock) {
  const float dv[3] = {dx, dy, dz};
  float e[16];
  if (cg == 1) { encode_slice<0, 27>(dv, sc_dir, e); store_row16<FP16, false>(dst, row, 0, e, gblock); }
  if (cg == 2) { encode_slice<16, 27>(dv, sc_dir, e); store_row16<FP16, false>(dst, row, 2, e, gblock); }
  if (STASH && (cg == 0 || cg == 3)) {   // zero the unused half of the stashed dirs block once per tile
    const float z[16] = {0.f};
    store_row16<FP16, false>(0u, row, cg == 0 ? 4 : 6, z, gblock);
  }
}

// ============================================================================================ forward
// Ring slots of one tile, in order (producer, issuer and epilogue warps all walk the same sequence):
//   [S]  the tile's lin_in OPERAND (encoded xyz, 16 KB) -- written by the epilogue warps, not the producer
//   [lin_in K-block 0] [fc_0 K-blocks 0..3] ... [view layer K-blocks 0..4]   weights (bulk copies by the producer)
// Giving lin_in's operand a ring slot of its own (instead of A block 0, which still holds feature_linear's output while the
// view layer runs) lets the epilogue warps encode tile t + 1 while tile t's view-layer MMAs run, and lets lin_in's MMAs
// of tile t + 1 run while the view layer's epilogue of tile t (rgb head, raw outputs) is busy: the accumulator regions have
// one completion barrier each (X: lin_in / fc_1 / feature_linear, T: fc_0 / lin_out / view layer), so lin_in (t + 1) -> X
// can complete before the view layer's accumulator T has been read.
template <bool FP16, bool STASH, bool WSHARE = false>
__global__ void __launch_bounds__(STASH ? TC_THREADS_STASH : TC_THREADS, 1)
mlp_fwd_tc_kernel(const TcLayout lay, const uint8_t* __restrict__ packed, const StarPtsSrc pts,
                  const float* __restrict__ viewdirs, const float* __restrict__ pose12,
                  const float* __restrict__ sc_xyz, const float* __restrict__ sc_dir, int S, int64_t M,
                  float* __restrict__ raw_alpha, float* __restrict__ raw_rgb, int64_t ray_stride,
                  uint8_t* __restrict__ stash, int* __restrict__ status, int* dbg, int dbg_mode_arg) {
  // status (nullable): word set to 1 when a raw output is not finite -- the range guard of the fp16-operand tier (an
  // activation beyond 65504 becomes +inf in the 16-bit operand and reaches the heads as inf / NaN) and, for any tier,
  // the sign of non-finite inputs or weights.
#ifdef STAR_TC_DEBUG
  const int dbg_mode = dbg_mode_arg;
#else
  constexpr int dbg_mode = 0;    // the bottleneck-experiment switches below exist in -DSTAR_TC_DEBUG builds only
  (void)dbg_mode_arg;
#endif
  // stash (NULL for inference): [tiles][lay.stash_blocks] blocks of TC_BLOCK_BYTES (mlp_tc_layout.h).
  // dbg_mode (bottleneck experiments only; results are garbage): bit 0 = epilogue skips TMEM load / math /
  // A store, bit 1 = no weight streaming (MMA reads whatever is in the ring), bit 2 = MMA issuer skips the MMAs,
  // training variant: bit 3 = no stash bulk stores, bit 4 = no ReLU bit masks, bit 5 = no stash_done waits
  extern __shared__ uint8_t smem_raw[];
  tc_mark_begin(dbg);
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  constexpr int NS = TC_NS;
  constexpr uint32_t STAGE_BYTES = TC_STAGE_BYTES;
  const TcSmem sl = tc_smem_layout(lay.small_bytes);
  const uint32_t sA = base + sl.A, sAD = base + sl.AD, sW = base + sl.W, sBars = base + sl.bars;
  float* s_small = reinterpret_cast<float*>(gbase + sl.small);
  // partial head sums of column group g (1..3) of a row: 4 floats in 16-byte chunk (3 + g) of the row of the
  // dirs block (logical columns 32..63, which the MMA never reads: only 2 of its 4 K-steps are issued)
  auto part = [&](int r, int g) -> float* {
    return reinterpret_cast<float*>(gbase + sl.part + (uint32_t)r * 128u + ((((uint32_t)(3 + g)) ^ ((uint32_t)r & 7u)) << 4));
  };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(gbase + sl.tmem_ptr);
  auto bar = [&](int i) -> uint32_t { return sBars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // (WSHARE: the two CTAs of a cluster walk the ring in lock step and must make the same number of trips: the tile count is
  //  rounded up to even -- the grid is even -- and a tile beyond the launch is a ghost: zero inputs, no outputs)
  const int64_t ntiles = WSHARE ? (((M + TC_M - 1) / TC_M + 1) & ~(int64_t)1) : (M + TC_M - 1) / TC_M;
  const int64_t tile0 = (int64_t)blockIdx.x, tile_step = (int64_t)gridDim.x;
  const uint32_t crank = WSHARE ? cluster_ctarank() : 0u;
  const uint32_t slots_per_tile = 1u + (uint32_t)lay.n_stages;
  const long long t_start = clock64();
#ifdef STAR_TC_TIMELINE
  // -DSTAR_TC_TIMELINE (debug build only, see tools/tc_timeline.sh): CTA 0 records clock64() stamps of its second tile
  // into dbg[8..]:  issuer, slot 16 + 4 l: {a_ready[0] seen, last K-block issued + committed}; slots 240 + 3 kb: layer
  // TL_LAYER per K-block {a_ready seen, w_full seen, issued};  epilogue warp 0, slot 80 + 8 l: {wait start, accumulator
  // seen, layer done}; slots 200 + w: layer-1 epilogue done per warp; slots 8 / 9: encode (xyz) start / done
  long long* tl = (dbg != nullptr && blockIdx.x == 0) ? reinterpret_cast<long long*>(dbg) : nullptr;
  const int64_t tl_tile = tile0 + tile_step;
#define TL_STAMP(cond, slot) do { if (tl != nullptr && (cond)) tl[slot] = clock64(); } while (0)
#ifndef TL_LAYER
#define TL_LAYER 2             // the layer whose K-blocks the issuer stamps one by one
#endif
#else
#define TL_STAMP(cond, slot) do { } while (0)
#endif

  // ---- one-time setup
  if (warp == TC_EPI_WARPS && lane == 0) {
    // w_full: the producer (+ its bytes) and the 16 epilogue warps (operand block of the same K-block)
    for (int i = 0; i < NS; ++i) { mbar_init(bar(BAR_W_FULL(i)), 1 + TC_EPI_WARPS); mbar_init(bar(BAR_W_EMPTY(i)), WSHARE ? 2 : 1); }
    for (int i = 0; i < 5; ++i) mbar_init(bar(BAR_A_READY(i)), TC_EPI_WARPS);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_ACC_FULL_T), 1);
    mbar_init(bar(BAR_S_READY), TC_EPI_WARPS);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_STASH_DONE_KB(i)), 1);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS + 1) tmem_alloc(base + sl.tmem_ptr, TC_TMEM_COLS);
  constexpr int NTHREADS = STASH ? TC_THREADS_STASH : TC_THREADS;
  for (int i = tid; i < lay.small_floats; i += NTHREADS) s_small[i] = reinterpret_cast<const float*>(packed)[i];
  for (int i = tid; i < TC_KB_BYTES / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(gbase + sl.AD)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (WSHARE) cluster_sync_all();      // the peer's barriers are initialised before anything of this CTA reaches them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == TC_EPI_WARPS) {
    // ======================================================================== weight producer
    // Training (STASH): the same thread also writes the activation stash.  A shared-memory A block IS the 16 KB stash
    // block (same swizzled image), so once all 16 epilogue warps have published output block kb of layer ls
    // (a_ready[kb]) one bulk store moves it to global memory -- instead of two 16-byte stores per thread and chunk.
    // The store of (ls, kb) is issued just before the weights of (ls + 2, kb) are requested (it cannot be needed
    // earlier than that: the weight stage frees only after MMA (ls + 1, kb), which itself waits for a_ready[kb]);
    // stash_done tells the epilogue of layer ls + 1 that the A blocks have been read and may be overwritten.
    if (lane == 0) {
      const uint8_t* wstream = packed + lay.small_bytes;
      uint32_t stage = 0, phase = 0, s_par = 0;
      for (int64_t tile = tile0; tile < ntiles; tile += tile_step) {
        uint8_t* st_tile = STASH ? stash + (size_t)tile * (size_t)lay.stash_blocks * TC_BLOCK_BYTES : nullptr;
        auto stash_chunk = [&](int ls, int kb) {
          mbar_wait(bar(BAR_A_READY(kb)), (s_par >> kb) & 1u, dbg, 6);
          s_par ^= 1u << kb;
          if (!(dbg_mode & 8)) {
            bulk_s2g(st_tile + (size_t)(lay.L[ls].s_out + kb) * TC_BLOCK_BYTES, sA + (uint32_t)kb * TC_KB_BYTES, TC_KB_BYTES);
            bulk_commit_group();
          }
          // block kb - 1 may be overwritten once its own store has read it (groups complete in order): the epilogue that
          // next writes it waits per block, so only the last block's store can still be in flight by then
          if (kb >= 1) {
            bulk_wait_group_read1();
            mbar_arrive(bar(BAR_STASH_DONE_KB(kb - 1)));
          }
          if (kb == 3) {
            bulk_wait_group_read0();
            mbar_arrive(bar(BAR_STASH_DONE_KB(3)));
          }
        };
        // slot S belongs to the epilogue warps (they wait for the same w_empty phase before they write it and announce it
        // on lin_in's w_full); the producer only keeps the slot's own barriers in step with the ring
        mbar_wait(bar(BAR_W_EMPTY(stage)), phase ^ 1u, dbg, 1);
        mbar_arrive_n(bar(BAR_W_FULL(stage)), 1 + TC_EPI_WARPS);
        if (++stage == NS) { stage = 0; phase ^= 1u; }
        for (int l = 0; l < lay.n_layers; ++l) {
          const uint32_t kb_bytes = (uint32_t)lay.L[l].N * 128u;
          for (int kb = 0; kb < lay.L[l].nkb; ++kb) {
            if (STASH && !TC_STASH_COPY_WARP && l >= 2 && kb < 4) stash_chunk(l - 2, kb);
            mbar_wait(bar(BAR_W_EMPTY(stage)), phase ^ 1u, dbg, 1);
            // operands with barriers of their own (the 16 epilogue warps do not arrive on w_full for them): the dirs block
            // (a_ready[4]) and lin_in's operand in slot S (s_ready) -- both are written far ahead of their K-block's turn,
            // when this stage's previous phase may still be open
            if ((lay.L[l].kind == LK_VIEWS && kb == 4) || l == 0) mbar_arrive_n(bar(BAR_W_FULL(stage)), TC_EPI_WARPS);
            if (dbg_mode & 2) {
              mbar_arrive(bar(BAR_W_FULL(stage)));
            } else {
              mbar_arrive_expect_tx(bar(BAR_W_FULL(stage)), kb_bytes);
              if (WSHARE) {      // this CTA's half of the K-block, into both CTAs' stage (each half signals both w_full barriers)
                const uint32_t half = kb_bytes >> 1;
                bulk_g2s_mcast(sW + stage * STAGE_BYTES + crank * half,
                               wstream + lay.L[l].w_off + (uint32_t)kb * kb_bytes + crank * half, half,
                               bar(BAR_W_FULL(stage)), (uint16_t)3);
              } else {
                bulk_g2s(sW + stage * STAGE_BYTES, wstream + lay.L[l].w_off + (uint32_t)kb * kb_bytes, kb_bytes,
                         bar(BAR_W_FULL(stage)));
              }
            }
            if (++stage == NS) { stage = 0; phase ^= 1u; }
          }
        }
        if (STASH && !TC_STASH_COPY_WARP)
          for (int kb = 0; kb < 4; ++kb) stash_chunk(lay.n_layers - 2, kb);   // feature_linear's output blocks
      }
      if (STASH && !TC_STASH_COPY_WARP) bulk_wait_group0();
    }
  } else if (STASH && TC_STASH_COPY_WARP && warp == TC_EPI_WARPS + 2) {
    // ======================================================================== stash copy warp (training)
    // Output block kb of every layer but the last is at once the next layer's A operand and a 16 KB block of the activation
    // stash (same swizzled image).  Once the 16 epilogue warps have published it (a_ready[kb]: their generic-proxy stores are
    // visible to this warp's loads) the warp copies it, 8 x 512 bytes in flight, and releases it (stash_done[kb]) as soon as the
    // loads have landed in registers -- the epilogue of the next layer waits for that before it overwrites the block.
    uint32_t s_par = 0;
    for (int64_t tile = tile0; tile < ntiles; tile += tile_step) {
      uint8_t* st_tile = stash + (size_t)tile * (size_t)lay.stash_blocks * TC_BLOCK_BYTES;
      for (int ls = 0; ls + 1 < lay.n_layers; ++ls) {
        uint4* dst_l = reinterpret_cast<uint4*>(st_tile + (size_t)lay.L[ls].s_out * TC_BLOCK_BYTES) + lane;
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(BAR_A_READY(kb)), (s_par >> kb) & 1u, dbg, 6);
          s_par ^= 1u << kb;
          const uint4* src = reinterpret_cast<const uint4*>(gbase + sl.A + (uint32_t)kb * TC_KB_BYTES) + lane;
          uint4* dst = dst_l + (size_t)kb * (TC_BLOCK_BYTES / 16);
          if (!(dbg_mode & 8)) {
#pragma unroll
            for (int it = 0; it < 32; it += 8) {
              uint4 v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = src[(it + j) * 32];
#pragma unroll
              for (int j = 0; j < 8; ++j) __stcs(dst + (it + j) * 32, v[j]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_STASH_DONE_KB(kb)));
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    // warp-uniform control flow (all lanes wait on the barriers), one elected lane issues the tcgen05 instructions
    uint32_t stage = 0, phase = 0, a_par = 0;
    const bool no_mma = (dbg_mode & 4) != 0;
    const uint64_t desc_a0 = umma_desc_sw128(sA), desc_ad = umma_desc_sw128(sAD), desc_w0 = umma_desc_sw128(sW);
    for (int64_t tile = tile0; tile < ntiles; tile += tile_step) {
      const uint32_t stage_s = stage;             // slot S: lin_in's operand
      if (++stage == NS) { stage = 0; phase ^= 1u; }
      for (int l = 0; l < lay.n_layers; ++l) {
        const int kind = lay.L[l].kind, nkb = lay.L[l].nkb;
        const uint32_t d_tmem = tmem_base + (lay.L[l].region ? 256u : 0u);
        const uint32_t acc_bar = bar(lay.L[l].region ? BAR_ACC_FULL_T : BAR_ACC_FULL);
        const uint32_t idesc = umma_idesc_16(TC_M, lay.L[l].N, FP16 ? 0 : 1);
        // The view layer multiplies the encoded-dirs block LAST (K-block 4).  It must be the dirs block that wraps around the
        // 4-stage ring onto the stage of the layer's K-block 0: its w_full arrivals all come from the producer thread, in
        // order behind K-block 0's.  A feature block in that position would have the 16 epilogue warps arrive on the barrier
        // while K-block 0's phase can still be open (the producer lags behind when stash stores back up) -- a count
        // underflow that the hardware answers with a launch failure (seen with the dirs block first, profiles/r2h_*).
        // one K-block: wait for its operand block + weights, issue 4 (2 for the dirs block) MMAs, release the weight stage
        // (SRC: 0 = A block kb_a, 1 = the dirs block, 2 = slot S)
        auto kblock = [&](const int kb, const int kb_a, const int src) {
          if (src == 1) {
            mbar_wait(bar(BAR_A_READY(4)), a_par & 1u, dbg, 2);
            a_par ^= 1u;
          }
          if (src == 2) {
            mbar_wait(bar(BAR_S_READY), (a_par >> 1) & 1u, dbg, 2);
            a_par ^= 2u;
          }
          TL_STAMP(tile == tl_tile && lane == 0 && kb == 0, 16 + 4 * l);
          TL_STAMP(tile == tl_tile && lane == 0 && l == TL_LAYER && kb < 5, 240 + 3 * kb);
          mbar_wait(bar(BAR_W_FULL(stage)), phase, dbg, 3);
          tc_fence_after();
          TL_STAMP(tile == tl_tile && lane == 0 && l == TL_LAYER && kb < 5, 241 + 3 * kb);
          const uint64_t a0 = src == 1 ? desc_ad
                            : src == 2 ? desc_w0 + (uint64_t)(stage_s * (STAGE_BYTES >> 4))
                                       : desc_a0 + (uint64_t)(kb_a * (TC_KB_BYTES >> 4));
          const uint64_t b0 = desc_w0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
          const uint32_t acc0 = (kind == LK_FC1 || kb > 0) ? 1u : 0u;
          if (elect_one_sync()) {
            if (!no_mma) {
              if (src == 1) tc_mma_kblock<2>(d_tmem, a0, b0, idesc, acc0);
              else tc_mma_kblock<4>(d_tmem, a0, b0, idesc, acc0);
            }
            if (WSHARE) {        // a stage is free when BOTH CTAs' MMAs have read it
              if (src == 2) tc_commit_mcast(bar(BAR_W_EMPTY(stage_s)), (uint16_t)3);
              tc_commit_mcast(bar(BAR_W_EMPTY(stage)), (uint16_t)3);
            } else {
              if (src == 2) tc_commit(bar(BAR_W_EMPTY(stage_s)));
              tc_commit(bar(BAR_W_EMPTY(stage)));
            }
            if (kb == nkb - 1) tc_commit(acc_bar);
          }
          __syncwarp();
          TL_STAMP(tile == tl_tile && lane == 0 && kb == nkb - 1, 17 + 4 * l);
          TL_STAMP(tile == tl_tile && lane == 0 && l == TL_LAYER && kb < 5, 242 + 3 * kb);
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        };
        if (nkb == 4) {            // the common case, unrolled: K-block indices become immediates
          kblock(0, 0, 0); kblock(1, 1, 0); kblock(2, 2, 0); kblock(3, 3, 0);
        } else if (nkb == 1) {     // lin_in
          kblock(0, 0, 2);
        } else if (kind == LK_VIEWS && nkb == 5) {
          kblock(0, 0, 0); kblock(1, 1, 0); kblock(2, 2, 0); kblock(3, 3, 0); kblock(4, 0, 1);
        } else {
          for (int kb = 0; kb < nkb; ++kb) kblock(kb, kb, (kind == LK_VIEWS && kb == 4) ? 1 : 0);
        }
      }
    }
  } else if (warp < TC_EPI_WARPS) {
    // ======================================================================== epilogue warps
    const int q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t acc_par = 0, stash_par = 0, stash_pend = 0;
    bool nonfinite = false;
    uint32_t kslot = 0;           // ring slot S of the current tile (same sequence as producer / issuer)
    EpiCtx ctx;
    ctx.sA = sA; ctx.a_ready0 = bar(BAR_A_READY(0)); ctx.w_full0 = bar(BAR_W_FULL(0));
    ctx.row = row; ctx.cg = cg; ctx.lane = lane;
    ctx.stash_out = nullptr;
    ctx.mask_out = nullptr;
    ctx.no_mask = (dbg_mode & 16) != 0;
    ctx.stash_done0 = bar(BAR_STASH_DONE_KB(0));
    ctx.publishes = false;
    ctx.ns_mask = NS - 1;
    ctx.no_stash_wait = (dbg_mode & 32) != 0;
    // encoded xyz of tile `tile_n` (4 threads per row) -> ring slot `slot` (= that tile's slot S), announced on s_ready (NOT on
    // the w_full barrier of lin_in's weights: this runs a whole layer before that stage's turn, while the barrier's phase of
    // the stage's previous K-block can still be open); dv = the rotated ray direction of this thread's row
    auto encode_xyz = [&](int64_t tile_n, uint32_t slot, float (&dv)[3]) {
      TL_STAMP(tile_n == tl_tile && tid == 0, 8);
      const TcSample sm = tc_load_sample(pts, viewdirs, pose12, tile_n * TC_M + row, M, S);
      dv[0] = sm.dv[0]; dv[1] = sm.dv[1]; dv[2] = sm.dv[2];
      const uint32_t st = slot & (NS - 1);
      mbar_wait(bar(BAR_W_EMPTY(st)), ((slot / NS) & 1u) ^ 1u, dbg, 8);     // the MMAs that last read this stage are done
      uint8_t* st_tile = STASH ? stash + (size_t)tile_n * (size_t)lay.stash_blocks * TC_BLOCK_BYTES : nullptr;
      tc_encode_xyz<FP16>(sm.p[0], sm.p[1], sm.p[2], sc_xyz, cg, row, sW + st * STAGE_BYTES,
                          STASH ? st_tile + (size_t)lay.L[0].s_in * TC_BLOCK_BYTES : nullptr);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_S_READY));
      TL_STAMP(tile_n == tl_tile && tid == 0, 9);
    };
    float dv_next[3];
    if (tile0 < ntiles) encode_xyz(tile0, 0u, dv_next);
    for (int64_t tile = tile0; tile < ntiles; tile += tile_step) {
      const int64_t gi = tile * TC_M + row;
      const bool valid = gi < M;
      int64_t out_idx = 0;
      if (valid) {
        const int64_t r = gi / S;
        out_idx = r * ray_stride + (gi - r * S);
      }
      float dv[3] = {dv_next[0], dv_next[1], dv_next[2]};
      uint32_t kstage = (kslot + 1u) & (NS - 1);     // stage of the current layer's K-block 0
      // ---- layers
      for (int l = 0; l < lay.n_layers; ++l) {
        const TcLayer& L = lay.L[l];
        // everything that does not depend on the accumulator is set up BEFORE the wait: the stretch from "accumulator
        // complete" to "first operand block of the next layer published" is tensor-idle time of every layer
        const int kind = L.kind;
        ctx.tcol = tmem_base + (((uint32_t)(q * 32)) << 16) + (L.region ? 256u : 0u) + (uint32_t)(cg * TC_CPT);
        ctx.bias = s_small + L.bias_off;
        if (STASH) {
          ctx.stash_out = stash + ((size_t)tile * (size_t)lay.stash_blocks + (size_t)L.s_out) * TC_BLOCK_BYTES;
          ctx.mask_out = stash + ((size_t)tile * (size_t)lay.stash_blocks + (size_t)lay.mask_blk0) * TC_BLOCK_BYTES +
                         (size_t)l * TC_MASK_BYTES;
        }
        float h[3] = {0.f, 0.f, 0.f};
        kstage = (kstage + (uint32_t)L.nkb) & (NS - 1);     // now the stage of the NEXT layer's K-block 0
        ctx.next_stage0 = kstage;
        ctx.publishes = STASH && l + 1 < lay.n_layers;
        if (kind == LK_VIEWS && tile + tile_step < ntiles)   // while the view layer's MMAs run: the next tile's lin_in operand
          encode_xyz(tile + tile_step, kslot + slots_per_tile, dv_next);
        TL_STAMP(tile == tl_tile && tid == 0, 80 + 8 * l);
        mbar_wait(bar(L.region ? BAR_ACC_FULL_T : BAR_ACC_FULL), (acc_par >> L.region) & 1u, dbg, 4);
        acc_par ^= 1u << L.region;
        tc_fence_after();
        TL_STAMP(tile == tl_tile && tid == 0, 81 + 8 * l);
        if (dbg_mode & 1) {
          for (int kb = 0; kb < (L.N >> 6) && kind != LK_VIEWS; ++kb) {
            fence_proxy_async_smem(); tc_fence_before(); __syncwarp();
            if (lane == 0) {
              mbar_arrive(bar(BAR_W_FULL((ctx.next_stage0 + (uint32_t)kb) & (NS - 1))));
              if (STASH) mbar_arrive(bar(BAR_A_READY(kb)));
            }
          }
        } else if (kind == LK_FC0 || kind == LK_FC1 || kind == LK_IN) {
          epilogue_layer<LK_FC0, FP16, STASH>(ctx, h, stash_par, stash_pend);
        } else if (kind == LK_FEAT) {
          epilogue_layer<LK_FEAT, FP16, STASH>(ctx, h, stash_par, stash_pend);
        } else if (kind == LK_OUT) {
          ctx.head_w = s_small + lay.off_alpha_w;
          epilogue_layer<LK_OUT, FP16, STASH>(ctx, h, stash_par, stash_pend);
          // combine the 4 column groups of each row: groups 1..3 park their partial, group 0 finishes
          if (cg != 0) part(row, cg)[0] = h[0];
          named_bar_sync(1, TC_EPI_THREADS);
          if (cg == 0 && valid) {
            const float ra = h[0] + part(row, 1)[0] + part(row, 2)[0] + part(row, 3)[0] + s_small[lay.off_alpha_b];
            raw_alpha[out_idx] = ra;
            nonfinite |= !(fabsf(ra) <= 3.4028235e38f);
          }
        } else {   // LK_VIEWS
          ctx.head_w = s_small + lay.off_rgb_w;
          epilogue_layer<LK_VIEWS, FP16, STASH>(ctx, h, stash_par, stash_pend);
          if (cg != 0) {
            float* d = part(row, cg);
            d[1] = h[0]; d[2] = h[1]; d[3] = h[2];
          }
          tc_fence_before();
          named_bar_sync(1, TC_EPI_THREADS);
          if (cg == 0 && valid) {
            float* o = raw_rgb + out_idx * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              const float rc = h[ch] + part(row, 1)[1 + ch] + part(row, 2)[1 + ch] + part(row, 3)[1 + ch] +
                               s_small[lay.off_rgb_b + ch];
              o[ch] = rc;
              nonfinite |= !(fabsf(rc) <= 3.4028235e38f);
            }
          }
        }
        TL_STAMP(tile == tl_tile && tid == 0, 82 + 8 * l);
        TL_STAMP(tile == tl_tile && lane == 0 && l == 1, 200 + warp);
        if (l == 0) {
          // the dirs block is only read by the view layer (and the previous tile's view layer is complete: its accumulator
          // was read before lin_in's): encoded after lin_in's epilogue, while fc_0's MMAs run
          uint8_t* st_tile = STASH ? stash + (size_t)tile * (size_t)lay.stash_blocks * TC_BLOCK_BYTES : nullptr;
          uint8_t* st_dirs = STASH ? st_tile + (size_t)(lay.L[lay.n_layers - 1].s_in + 4) * TC_BLOCK_BYTES : nullptr;
          tc_encode_dirs<FP16, STASH>(dv[0], dv[1], dv[2], sc_dir, cg, row, sAD, st_dirs);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_A_READY(4)));
        }
      }
      kslot += slots_per_tile;
    }
    if (status != nullptr && nonfinite) *reinterpret_cast<volatile int*>(status) = 1;
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (WSHARE) cluster_sync_all();      // no CTA leaves while its peer's commits / copies can still land in it
  tc_mark_end(dbg);
#ifdef STAR_TC_DEBUG
  if (dbg != nullptr && tid == 0 && blockIdx.x == 0) {   // debug only: cycles of CTA 0 (see star_tc_forward)
    reinterpret_cast<long long*>(dbg)[1] = clock64() - t_start;
  }
#else
  (void)t_start;
#endif
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ============================================================================================ packing
__global__ void pack_tc_small_kernel(TcLayout tl, MlpLayout ml, const float* __restrict__ master,
                                     float* __restrict__ small) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tl.small_floats; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    bool done = false;
    for (int l = 0; l < tl.n_layers && !done; ++l) {
      const int o = i - tl.L[l].bias_off;
      if (o >= 0 && o < STAR_W) {
        done = true;
        if (o < tl.L[l].N) {
          v = master[ml.L[l].m_b + o];
          if (tl.L[l].kind == LK_FC1) {   // cumulative bias of the residual stream up to this block
            v += master[ml.L[0].m_b + o];
            for (int j = 2; j < l; j += 2) v += master[ml.L[j].m_b + o];
          }
        }
      }
    }
    if (!done) {
      if (i >= tl.off_alpha_w && i < tl.off_alpha_w + STAR_W) v = master[ml.m_alpha_w + (i - tl.off_alpha_w)];
      else if (i == tl.off_alpha_b) v = master[ml.m_alpha_b];
      else if (i >= tl.off_rgb_w && i < tl.off_rgb_w + 3 * STAR_WV) v = master[ml.m_rgb_w + (i - tl.off_rgb_w)];
      else if (i >= tl.off_rgb_b && i < tl.off_rgb_b + 3) v = master[ml.m_rgb_b + (i - tl.off_rgb_b)];
    }
    small[i] = v;
  }
}

__global__ void pack_tc_stream_kernel(TcLayout tl, MlpLayout ml, const float* __restrict__ master,
                                      uint16_t* __restrict__ stream, int fp16) {
  const uint32_t n_elems = tl.stream_bytes / 2;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += gridDim.x * blockDim.x) {
    const uint32_t byte = e * 2;
    int l = 0;
    while (l + 1 < tl.n_layers && byte >= tl.L[l + 1].w_off) ++l;
    const TcLayer& L = tl.L[l];
    const uint32_t off = byte - L.w_off, kb_bytes = (uint32_t)L.N * 128u;
    const int kb = (int)(off / kb_bytes);
    const uint32_t rem = off % kb_bytes;
    const int n = (int)(rem >> 7);
    const uint32_t inrow = rem & 127u;
    const int chunk = (int)((inrow >> 4) ^ ((uint32_t)n & 7u));   // undo the 128-byte swizzle
    const int kk = chunk * 8 + (int)((inrow & 15u) >> 1);
    const int K = ml.L[l].K;                                     // true input width (63, 256, 283)
    float v = 0.f;
    if (L.kind == LK_VIEWS) {          // K-blocks 0..3 = the feature blocks, 4 = the encoded dirs
      const int k = (kb < 4) ? kb * 64 + kk : STAR_W + kk;
      if (kb < 4 || kk < K - STAR_W) v = master[ml.L[l].m_w + (int64_t)n * K + k];
    } else {
      const int k = kb * 64 + kk;
      if (k < K) v = master[ml.L[l].m_w + (int64_t)n * K + k];
    }
    stream[e] = fp16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// ============================================================================================ host side
int star_tc_pack_tstream(const TcLayout& tl, const MlpLayout& ml, const float* master, void* tstream, int fp16,
                         cudaStream_t st);

// packed image = [small section][forward weight stream][transposed (backward) weight stream]
size_t star_tc_packed_bytes(const TcLayout& tl) { return (size_t)tl.small_bytes + tl.stream_bytes + tl.tstream_bytes; }

int star_tc_pack(const TcLayout& tl, const MlpLayout& ml, const float* master, void* packed, int fp16,
                 cudaStream_t st) {
  pack_tc_small_kernel<<<8, 256, 0, st>>>(tl, ml, master, (float*)packed);
  int rc = star_check_launch();
  if (rc) return rc;
  pack_tc_stream_kernel<<<148 * 4, 256, 0, st>>>(tl, ml, master, (uint16_t*)((uint8_t*)packed + tl.small_bytes),
                                                 fp16);
  rc = star_check_launch();
  if (rc) return rc;
  return star_tc_pack_tstream(tl, ml, master, (uint8_t*)packed + tl.small_bytes + tl.stream_bytes, fp16, st);
}

int star_tc_forward(const TcLayout& tl, const void* packed, const StarPtsSrc& pts, const float* viewdirs,
                    const float* pose12, const float* sc_xyz, const float* sc_dir, int R, int S, float* raw_alpha,
                    float* raw_rgb, int64_t ray_stride, void* stash, int* status, int fp16_flags, cudaStream_t st) {
  const bool fp16 = (fp16_flags & 1) != 0, no_wshare = (fp16_flags & 2) != 0;   // bit 1: STAR_PREC_FLAG_NO_WSHARE
  const int64_t M = (int64_t)R * S;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (((uintptr_t)packed & 15) != 0) return STAR_E_ALIGN;
  const bool with_stash = stash != nullptr;
  const TcSmem sl = tc_smem_layout(tl.small_bytes);
  const int grid = (int)(ntiles < sms ? ntiles : sms);       // persistent: one CTA per SM over the tiles
  using Kern = void (*)(const TcLayout, const uint8_t*, const StarPtsSrc, const float*, const float*, const float*,
                        const float*, int, int64_t, float*, float*, int64_t, uint8_t*, int*, int*, int);
  // (a device / partition that refuses the cluster launch -- a synchronous launch error, nothing has run -- gets the plain
  //  launch of the same kernel family from then on, with one line on stderr: a launch shape, not a numerical fallback)
  static bool s_cluster_refused = false;
  const bool wshare = TC_WSHARE && !no_wshare && !s_cluster_refused && (!with_stash || TC_WSHARE_STASH) &&
                      ntiles >= 2 * (int64_t)sms;
  Kern kern_plain = with_stash ? (fp16 ? mlp_fwd_tc_kernel<true, true> : mlp_fwd_tc_kernel<false, true>)
                               : (fp16 ? mlp_fwd_tc_kernel<true, false> : mlp_fwd_tc_kernel<false, false>);
  Kern kern = with_stash ? (wshare && TC_WSHARE_STASH ? (fp16 ? mlp_fwd_tc_kernel<true, true, TC_WSHARE_STASH != 0>
                                                              : mlp_fwd_tc_kernel<false, true, TC_WSHARE_STASH != 0>)
                                                      : (fp16 ? mlp_fwd_tc_kernel<true, true> : mlp_fwd_tc_kernel<false, true>))
              : wshare   ? (fp16 ? mlp_fwd_tc_kernel<true, false, true> : mlp_fwd_tc_kernel<false, false, true>)
                         : (fp16 ? mlp_fwd_tc_kernel<true, false> : mlp_fwd_tc_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
  if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
  if (kern_plain != kern) {
    e = cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
    if (e != cudaSuccess) { g_star_last_cuda_error = (int)e; return STAR_E_CUDA; }
  }
  auto launch = [&](int* dbg, int dbg_mode) -> int {
    if (wshare) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(grid & ~1), 1, 1);
      cfg.blockDim = dim3(with_stash ? TC_THREADS_STASH : TC_THREADS, 1, 1);
      cfg.dynamicSmemBytes = sl.total;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tl, (const uint8_t*)packed, pts, viewdirs, pose12, sc_xyz, sc_dir, S, M,
                                                raw_alpha, raw_rgb, ray_stride, (uint8_t*)stash, status, dbg, dbg_mode);
      if (le == cudaSuccess) return STAR_OK;
      (void)cudaGetLastError();      // the refused launch leaves a (non-sticky) error behind
      s_cluster_refused = true;
      fprintf(stderr, "[star_b200] cluster launch of the MLP forward refused (%s): one CTA per SM streams its own weights\n",
              cudaGetErrorString(le));
    }
    kern_plain<<<grid, with_stash ? TC_THREADS_STASH : TC_THREADS, sl.total, st>>>(tl, (const uint8_t*)packed, pts, viewdirs, pose12, sc_xyz, sc_dir, S, M,
                                             raw_alpha, raw_rgb, ray_stride, (uint8_t*)stash, status, dbg, dbg_mode);
    return STAR_OK;
  };
#ifndef STAR_TC_DEBUG
  { const int lrc = launch(star_watchdog_dev(STAR_WD_FWD), 0); if (lrc) return lrc; }
#else
  // -DSTAR_TC_DEBUG builds only (tools/tc_debug_modes.sh, tools/tc_timeline.sh): STAR_TC_DEBUG_MODE switches parts of the
  // kernel off (results are garbage), STAR_TC_DEBUG_CYCLES=1 makes the launch synchronous and prints CTA 0's cycles
  static int dbg_mode = -1;
  if (dbg_mode < 0) { const char* e2 = getenv("STAR_TC_DEBUG_MODE"); dbg_mode = e2 ? atoi(e2) : 0; }
  // STAR_TC_DEBUG_CYCLES=1 (bottleneck experiments only): synchronous launch that prints the SM cycles CTA 0 spent
  static int dbg_cycles = -1;
  static long long* d_dbg = nullptr;
  if (dbg_cycles < 0) {
    dbg_cycles = getenv("STAR_TC_DEBUG_CYCLES") ? 1 : 0;
    if (dbg_cycles) { cudaMalloc(&d_dbg, 4096); cudaMemset(d_dbg, 0, 4096); }
  }
  { const int lrc = launch((int*)d_dbg, dbg_mode); if (lrc) return lrc; }
  if (dbg_cycles) {
    static long long h[512];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, 4096, cudaMemcpyDeviceToHost);
    const double t0 = (double)((ntiles + grid - 1) / grid);
    fprintf(stderr, "[star_tc] mode %d: CTA0 %.0f cycles/tile\n", dbg_mode, h[1] / t0);
#ifdef STAR_TC_TIMELINE
    if (ntiles > 2 * (int64_t)grid && h[8] != 0) {
      const long long z = h[8];
      fprintf(stderr, "[star_tc] timeline of CTA 0, tile 2 (cycles from encode start): encode done %lld\n", h[9] - z);
      for (int l = 0; l < tl.n_layers; ++l)
        fprintf(stderr, "[star_tc]  L%-2d kind %d  issuer: a_ready0 %6lld  last kb issued %6lld | epilogue: wait-start %6lld  "
                        "acc seen %6lld  done %6lld\n", l, tl.L[l].kind, h[16 + 4 * l] - z, h[17 + 4 * l] - z,
                h[80 + 8 * l] - z, h[81 + 8 * l] - z, h[82 + 8 * l] - z);
      fprintf(stderr, "[star_tc]  L1 epilogue done per warp:");
      for (int w = 0; w < 16; ++w) fprintf(stderr, " %lld", h[200 + w] - z);
      fprintf(stderr, "\n[star_tc]  L%d issuer per K-block (a_ready seen, w_full seen, issued):", TL_LAYER);
      for (int kb = 0; kb < tl.L[TL_LAYER].nkb; ++kb) fprintf(stderr, "  kb%d %lld %lld %lld", kb, h[240 + 3 * kb] - z, h[241 + 3 * kb] - z, h[242 + 3 * kb] - z);
      fprintf(stderr, "\n");
      cudaMemset(d_dbg, 0, 4096);
    }
#endif
  }
#endif
  return star_check_launch();
}
