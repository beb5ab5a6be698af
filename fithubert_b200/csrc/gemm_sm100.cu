// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma
// (UMMA M=128, N<=256, K=16, accumulators in TMEM, double buffered) -> tcgen05.ld -> fused epilogue
// -> 128B-swizzled smem slab -> TMA store (TMA reduce-add for split-K).
//
// One CTA per SM, 352 threads:
//   warps 0-7  epilogue: warp w owns TMEM lanes 32(w%4)..+31 (= tile rows) and column half (w/4) of
//              every 128-byte output slab
//   warp  8    TMA producer (one elected lane)
//   warp  9    MMA issuer (one elected lane) + TMEM allocation
//   warp 10    store warp (one elected lane): TMA stores of finished output slabs, refills of the epilogue-input ring
// Four pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), staging-buffer full/empty
// (epilogue <-> store warp: the epilogue warps meet at no block barrier inside a tile), static
// round-robin tile schedule (tile = blockIdx.x + i * gridDim.x, n fastest so CTAs of one wave
// share the A tile in L2).
// PAIR MODE (K-major operands, p.pair = 1): the grid is launched as clusters of two CTAs that work on two vertically
// adjacent row blocks of the SAME n-block.  Each CTA loads its own A tile and HALF of the shared B tile, multicast into
// both CTAs' shared memory (cp.async.bulk.tensor ... .multicast::cluster): the L2 -> SM operand traffic of B halves
// (48 -> 32 KB per k-block and CTA at BN = 256).  A stage is free again once BOTH CTAs' MMAs have consumed it
// (tcgen05.commit ... .multicast::cluster onto empty barriers of count 2).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "fhb_common.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kMaxBN = 256;
constexpr int kStages = 6;   // barrier slots; the pipeline depth is p.stages (<= 4 with whole B tiles, <= 6 with cta_group::2 half tiles)
constexpr int kAccStages = 2;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 96;  // + TMA producer warp, MMA warp, store warp
constexpr int kMaxOut = 4;                  // most output staging buffers (p.n_out)
constexpr uint32_t kABytes = kBM * kBK * 2;     // 16 KiB
constexpr uint32_t kBBytes = kMaxBN * kBK * 2;  // 32 KiB
constexpr uint32_t kStoreBytes = kBM * 128;  // one output slab: 128 rows x 128 bytes (64 bf16 / 32 fp32 columns)
constexpr uint32_t kBiasBytes = 2 * kMaxBN * 4;
constexpr int kInRing = 6;  // most epilogue-input slabs in flight (the ring depth is p.n_in: 3, or 4 with cta_group::2)
constexpr uint32_t kBarBytes = 256;
// 14 units of 16 KiB: 3 per pipeline stage + 2 output slabs (+ 2 second-output slabs) (+ 3 epilogue-input slabs)
constexpr uint32_t kSmemBytes = 14 * kStoreBytes + kBiasBytes + kBarBytes;

#ifdef FHB_GEMM_TRACE
// tools/gemm_trace.py: clock64 stamps of the epilogue phases of CTA 0 (thread 0 and thread 255), 8 per slab
__device__ long long* g_gemm_trace = nullptr;
#define FHB_TRACE(slot)                                                                              \
  do {                                                                                               \
    if (trace_on && slab_ctr < 96u) g_gemm_trace[(trace_who * 96 + slab_ctr) * 8 + (slot)] = clock64(); \
  } while (0)
// whole-kernel timeline of CTA 0 (one stamp per slot, written by whichever thread passes the point): [2 * 96 * 8 + slot]
#define FHB_TL(slot)                                                                                  \
  do {                                                                                                \
    if (g_gemm_trace != nullptr && blockIdx.x == 0) g_gemm_trace[2 * 96 * 8 + (slot)] = clock64();   \
  } while (0)
// start / end of EVERY CTA in nanoseconds (%globaltimer): [2 * 96 * 8 + 32 + 2 * blockIdx.x + {0, 1}], up to 512 CTAs
__device__ __forceinline__ long long fhb_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FHB_CTA_STAMP(which)                                                                                    \
  do {                                                                                                          \
    if (g_gemm_trace != nullptr && blockIdx.x < 512) g_gemm_trace[2 * 96 * 8 + 32 + 2 * blockIdx.x + (which)] = fhb_globaltimer(); \
  } while (0)
// log of every launch (CTA 0: start / end in ns, shape): tools/gemm_steplog.py reads it after a few training steps
__device__ long long g_gemm_log[6 * 8192];
__device__ unsigned int g_gemm_log_n = 0;
#else
#define FHB_TRACE(slot) do {} while (0)
#define FHB_TL(slot) do {} while (0)
#define FHB_CTA_STAMP(which) do {} while (0)
#endif

struct GemmParams {
  int m, n, k, bn;
  int num_m_blk, num_n_blk, num_ob, ob_mod, split_k;
  int kb_per_cb, kb_total, kb_per_split;
  int a_lo_c0, a_hi_c2, a_lo_c2, a_cb_c2;
  int b_lo_c0, b_hi_c2, b_lo_c2, b_cb_c2;
  int a_c1_off, b_c1_off;
  long long d_ld, d_hi_stride, d_lo_stride, bias_hi_stride;
  void* d;
  const float* bias;
  const void* residual;  // fp16, bf16 (FHB_EPI_RES_BF16) or fp32 (FHB_EPI_RES_F32)
  const __half* aux_in;  // always fp16 (a forward-pass quantity: a saved gelu' or pre-activation)
  void* aux_out;         // fp16 (with FHB_EPI_AUX_DGELU: always; else the type of D)
  const int* row_valid;
  const __half* loss_target;
  float* loss_acc;
  float loss_weight, grad_scale, alpha;
  uint32_t drop_seed, drop_thr;
  float drop_scale;
  int flags;
  uint32_t stage_tx_bytes;
  int total_tiles;
  int use_tma_store;
  int stages;    // smem pipeline depth: 4, 3 or 2 depending on how many epilogue staging buffers are needed
  int n_out;     // 2 .. 4 staging buffers for the TMA stores of D (deeper = more stores in flight behind the epilogue)
  int n_auxout;  // 0 / n_out staging buffers for the second output
  int n_in;      // 0 / 3 / 4 staging buffers (<= kInRing) for the TMA-loaded epilogue input (residual or aux_in)
  int pair;      // 1: clusters of two CTAs share the B tile by TMA multicast (num_m_blk then counts PAIRS of row blocks)
};

struct Tile {
  int ob_hi, ob_lo, m0, n0, kb_begin, kb_end;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load multicast to the CTAs of `mask`: lands at the same shared-memory offset in each and completes bytes on the
// barrier at the same offset in each
__device__ __forceinline__ void tma_load_3d_mc(const void* desc, uint64_t* bar, void* smem, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem)), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---- cta_group::2 (p.pair == 2): the two CTAs of a cluster run ONE UMMA of M = 256: each holds its 128 rows of A and
// HALF of the B tile in its own shared memory (the instruction reads both), each accumulates its 128 rows in its own
// TMEM.  Per CTA and k-block the tensor pipe then reads 16 KB A + bn/2 x 128 B of B instead of the whole B tile, and TMA
// writes as little: the shared-memory bandwidth that caps a one-CTA UMMA at ~2/3 of the tensor peak is no longer the bound.
// Only the leader (cluster rank 0) issues MMAs; both CTAs' loads complete on the LEADER's full barrier.
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;  // shared::cluster address -> same offset in the even CTA of the pair
__device__ __forceinline__ void tma_load_3d_cg2(const void* desc, uint32_t bar_addr, void* smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(desc), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_mma_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int tile, int pair_rank = 0) {
  Tile t;
  const int nb = tile % p.num_n_blk;
  tile /= p.num_n_blk;
  const int mb = tile % p.num_m_blk;
  tile /= p.num_m_blk;
  const int sp = tile % p.split_k;
  const int ob = tile / p.split_k;
  t.ob_hi = ob / p.ob_mod;
  t.ob_lo = ob % p.ob_mod;
  t.m0 = (p.pair ? 2 * mb + pair_rank : mb) * kBM;  // pair mode: mb counts pairs of row blocks (a row block past m is all clipped)
  t.n0 = nb * p.bn;
  t.kb_begin = sp * p.kb_per_split;
  t.kb_end = min(t.kb_begin + p.kb_per_split, p.kb_total);
  return t;
}

// EPI_IN = 1: the epilogue reads [m][n] operands (residual / aux_in / loss target) laid out like D
// CG2 = 1: the cta_group::2 instantiation.  A kernel that contains cta_group::2 instructions can only be launched as
// clusters of two ("cluster misconfiguration" otherwise), so the pair-of-SMs path is a separate instantiation.
// FLAGS_CT >= 0: the epilogue flag set is a compile-time constant (the hot flag sets of the distillation step get their own
// instantiation: the epilogue's ~15 runtime flag branches and the 25 KB of code behind them disappear); -1: p.flags at run time
template <int A_MN, int B_MN, int EPI_IN, int CG2 = 0, int FLAGS_CT = -1>
__global__ void __launch_bounds__(kThreads, 1)
fhb_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_aux,
                const __grid_constant__ CUtensorMap tm_in, const GemmParams p) {
  // pair mode: tm_b's box is HALF the B tile (bn / 2 rows); this CTA's rank in its cluster picks the half it loads
  const int pair_rank = p.pair ? (int)cluster_ctarank() : 0;
  const int first_tile = p.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = p.pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool cg2 = CG2 != 0;  // (the host launches this instantiation iff p.pair == 2)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nstages = p.stages;
  uint8_t* smem_a = smem;
  // cta_group::2: a stage holds A and HALF a B tile (32 KiB instead of 48): the freed shared memory deepens the pipeline -
  // 4 x 48 KiB in flight cover ~1.1 us of L2 latency at the full MMA rate, 6 x 32 KiB at half the bytes per CTA ~1.6 us
  constexpr uint32_t kBSlot = CG2 ? kBBytes / 2 : kBBytes;
  uint8_t* smem_b = smem + nstages * kABytes;
  uint8_t* smem_out = smem + nstages * (kABytes + kBSlot);     // 2 x 16 KiB staging buffers for TMA stores of D
  uint8_t* smem_auxo = smem_out + p.n_out * kStoreBytes;       // 0 / n_out buffers for the second output
  uint8_t* smem_in = smem_auxo + p.n_auxout * kStoreBytes;     // 0 / 3 buffers for the TMA-loaded epilogue input
  uint8_t* smem_tail = smem + 14 * kStoreBytes;
  float* bias_s = reinterpret_cast<float*>(smem_tail);  // [2][kMaxBN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_tail + kBiasBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = bars + 2 * kStages + kAccStages;
  uint64_t* in_full = bars + 2 * kStages + 2 * kAccStages;  // [kInRing]
  uint64_t* out_full = in_full + kInRing;     // [kMaxOut] 256 arrivals: every epilogue thread has written its share of the slab
  uint64_t* out_empty = out_full + kMaxOut;   // [kMaxOut] the store warp: the TMA store of that buffer has finished reading it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_empty + kMaxOut);
  static_assert((2 * kStages + 2 * kAccStages + kInRing + 2 * kMaxOut) * 8 + 4 <= kBarBytes, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FHB_TL(0);  // kernel entry
  if (threadIdx.x == 0) FHB_CTA_STAMP(0);
#ifdef FHB_GEMM_TRACE
  unsigned int log_slot = 0xFFFFFFFFu;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    log_slot = atomicAdd(&g_gemm_log_n, 1u);
    if (log_slot < 8192u) {
      g_gemm_log[6 * log_slot] = fhb_globaltimer();
      g_gemm_log[6 * log_slot + 2] = ((long long)p.m << 32) | (unsigned int)p.n;
      g_gemm_log[6 * log_slot + 3] = ((long long)p.k << 32) | (unsigned int)p.flags;
      g_gemm_log[6 * log_slot + 4] = clock64();
    }
  }
#endif

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();  // 128B-swizzle atoms need a 1024-byte aligned base
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_d);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], p.pair == 1 ? 2 : 1);  // multicast pairs: both CTAs' MMAs release a stage (the peer writes into it)
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], p.pair == 2 ? 2 : kEpiThreads);  // cta_group::2: one arrival per CTA of the pair (see the epilogue)
    }
    for (int i = 0; i < kInRing; ++i) mbar_init(&in_full[i], 1);
    for (int i = 0; i < kMaxOut; ++i) {
      mbar_init(&out_full[i], kEpiThreads);
      mbar_init(&out_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    if constexpr (CG2) tmem_alloc_cg2(tmem_slot, 512);  // collective: one warp of each CTA of the pair
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pair) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  if (threadIdx.x == 0) FHB_TL(1);  // barriers initialised, TMEM allocated
  pdl_sync();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) FHB_TL(2);  // predecessor grid complete

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_stride) {
        const Tile t = decode_tile(p, tile, pair_rank);
        const int a_c2 = t.ob_hi * p.a_hi_c2 + t.ob_lo * p.a_lo_c2;
        const int b_c2 = t.ob_hi * p.b_hi_c2 + t.ob_lo * p.b_lo_c2;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if constexpr (CG2) {
            // both CTAs' bytes complete on the leader's barrier; the leader arms it with the sum
            const uint32_t lbar = smem_u32(&full[stage]) & kLeaderMask;
            if (pair_rank == 0) mbar_expect_tx(&full[stage], p.stage_tx_bytes + kABytes);  // 2 x A + the two halves of B
            tma_load_3d_cg2(&tm_a, lbar, smem_a + stage * kABytes, kb * kBK + t.ob_lo * p.a_lo_c0, t.m0,
                            t.ob_hi * p.a_hi_c2 + t.ob_lo * p.a_lo_c2);
            const int n_eff2 = min(p.bn, (p.n - t.n0 + 15) & ~15) >> 1;  // this n-block's rows of B per CTA
            tma_load_3d_cg2(&tm_b, lbar, smem_b + stage * kBSlot, kb * kBK + t.ob_lo * p.b_lo_c0, t.n0 + pair_rank * n_eff2,
                            t.ob_hi * p.b_hi_c2 + t.ob_lo * p.b_lo_c2);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_expect_tx(&full[stage], p.stage_tx_bytes);
          const int cb = kb / p.kb_per_cb;
          const int kr = (kb - cb * p.kb_per_cb) * kBK;
          uint8_t* sa = smem_a + stage * kABytes;
          uint8_t* sb = smem_b + stage * kBSlot;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i)
              tma_load_3d(&tm_a, &full[stage], sa + i * (kBK * 128), t.m0 + i * 64 + t.ob_lo * p.a_lo_c0, kr + p.a_c1_off,
                          a_c2 + cb * p.a_cb_c2);
          } else {
            tma_load_3d(&tm_a, &full[stage], sa, kb * kBK + t.ob_lo * p.a_lo_c0, t.m0, a_c2);
          }
          if (B_MN) {
            const int atoms = (p.bn + 63) >> 6;
            for (int i = 0; i < atoms; ++i)
              tma_load_3d(&tm_b, &full[stage], sb + i * (kBK * 128), t.n0 + i * 64 + t.ob_lo * p.b_lo_c0, kr + p.b_c1_off,
                          b_c2 + cb * p.b_cb_c2);
          } else if (p.pair == 1) {
            const int half_rows = p.bn >> 1;
            tma_load_3d_mc(&tm_b, &full[stage], sb + pair_rank * half_rows * 128, kb * kBK + t.ob_lo * p.b_lo_c0,
                           t.n0 + pair_rank * half_rows, b_c2, (uint16_t)3);
          } else {
            tma_load_3d(&tm_b, &full[stage], sb, kb * kBK + t.ob_lo * p.b_lo_c0, t.n0, b_c2);
          }
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer (cta_group::2: the leader CTA only)
    if (elect_one() && !(cg2 && pair_rank != 0)) {
      // K-major: 8-row groups 1024 B apart (SBO); MN-major: 64-element atoms kBK*128 B apart (LBO),
      // 8-k-row groups 1024 B apart (SBO).
      const uint32_t a_lbo = A_MN ? kBK * 128 : 0, b_lbo = B_MN ? kBK * 128 : 0;
      const uint32_t a_kstep = A_MN ? 16 * 128 : 32, b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_stride, ++it) {
        const Tile t = decode_tile(p, tile, pair_rank);
        // the last n-block of a row may be narrower: issue only the columns that exist (multiple of 16)
        const int n_eff = min(p.bn, (p.n - t.n0 + 15) & ~15);
        const uint32_t idesc = umma_idesc_16(cg2 ? 2 * kBM : kBM, (uint32_t)n_eff, A_MN, B_MN, (p.flags & FHB_GEMM_A_BF16) ? 1u : 0u,
                                             (p.flags & FHB_GEMM_B_BF16) ? 1u : 0u);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kMaxBN;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          if (kb == t.kb_begin && it < 4) FHB_TL(4 + 4 * it);  // first stage of tile `it` has landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * kBSlot);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t db = umma_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            if constexpr (CG2) tc_mma_cg2(tmem_d, da, db, idesc, (kb > t.kb_begin || k > 0) ? 1u : 0u);
            else tc_mma_bf16(tmem_d, da, db, idesc, (kb > t.kb_begin || k > 0) ? 1u : 0u);
          }
          if constexpr (CG2) tc_commit_cg2(&empty[stage], (uint16_t)3);        // both CTAs' stage slots were read by these MMAs
          else if (p.pair) tc_commit_mc(&empty[stage], (uint16_t)3);  // ... in BOTH CTAs: the peer's B half lands in this slot too
          else tc_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (CG2) tc_commit_cg2(&acc_full[as], (uint16_t)3);  // each CTA's epilogue drains its own 128 rows
        else tc_commit(&acc_full[as]);
        if (it < 4) FHB_TL(5 + 4 * it);  // every MMA of tile `it` issued
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ store warp (one elected lane)
    // Takes every TMA out of the epilogue warps' path: it waits until all 256 epilogue threads have written (and fenced)
    // their share of an output slab, issues the TMA store(s) of that staging buffer, refills the epilogue-input ring slot
    // the slab has just released, and hands the buffer back once the store has finished reading it.  The epilogue warps
    // therefore never meet at a block barrier inside a tile: a warp that is done with slab s starts slab s + 1 (the other
    // buffer) at once.  (Measured before, profiles/r03_gemm_epilogue_trace.txt: per slab every epilogue thread waited
    // 125 - 370 clk at the barrier plus 350 clk for thread 0 to issue the store.)
    if (elect_one()) {
      const int flags = FLAGS_CT >= 0 ? FLAGS_CT : p.flags;
      const bool out_f32 = (flags & FHB_EPI_OUT_F32) != 0;
      const bool two_out = (flags & FHB_EPI_STORE_PREACT) != 0;
      const bool in_tma = EPI_IN && p.n_in > 0;
      const int slab_cols = out_f32 ? 32 : 64;
      const uint32_t out_u32 = smem_u32(smem_out), auxo_u32 = smem_u32(smem_auxo), in_u32 = smem_u32(smem_in);
      auto tile_slabs = [&](const Tile& t) {
        return p.use_tma_store ? (min(p.bn, p.n - t.n0) + slab_cols - 1) / slab_cols : 0;
      };
      // ---- input-ring prefetcher: walks the (tile, slab) sequence p.n_in slabs ahead of the consumers
      int pf_tile = first_tile, pf_sidx = 0, pf_ns = 0;
      uint32_t pf_slot = 0;
      Tile pf_t = {0, 0, 0, 0, 0, 0};
      auto prefetch_one = [&]() {
        if (pf_tile >= p.total_tiles) return;
        const uint32_t slot = pf_slot;
        if (++pf_slot == (uint32_t)p.n_in) pf_slot = 0;
        mbar_expect_tx(&in_full[slot], kStoreBytes);
        tma_load_4d(&tm_in, &in_full[slot], in_u32 + slot * kStoreBytes, pf_t.n0 + pf_sidx * slab_cols, pf_t.m0, pf_t.ob_lo,
                    pf_t.ob_hi);
        if (++pf_sidx == pf_ns) {
          pf_sidx = 0;
          pf_tile += tile_stride;
          if (pf_tile < p.total_tiles) {
            pf_t = decode_tile(p, pf_tile, pair_rank);
            pf_ns = tile_slabs(pf_t);
          }
        }
      };
      if (in_tma) {
        tma_prefetch_desc(&tm_in);
        if (pf_tile < p.total_tiles) {
          pf_t = decode_tile(p, pf_tile, pair_rank);
          pf_ns = tile_slabs(pf_t);
        }
        for (int i = 0; i < p.n_in; ++i) prefetch_one();  // all p.n_in slots start out free
      }
      uint32_t obuf = 0, ouse = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_stride) {
        const Tile t = decode_tile(p, tile, pair_rank);
        const int n_slabs = tile_slabs(t);
        for (int sidx = 0; sidx < n_slabs; ++sidx) {
          mbar_wait(&out_full[obuf], ouse & 1u);
          const uint32_t dbuf = out_u32 + obuf * kStoreBytes;
          const uint32_t abuf = auxo_u32 + obuf * kStoreBytes;
          const int cc = t.n0 + sidx * slab_cols;
          if (flags & FHB_EPI_ATOMIC_ADD) {
            asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&tm_d), "r"(dbuf), "r"(cc), "r"(t.m0), "r"(t.ob_lo), "r"(t.ob_hi) : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&tm_d), "r"(dbuf), "r"(cc), "r"(t.m0), "r"(t.ob_lo), "r"(t.ob_hi) : "memory");
          }
          if (two_out)
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&tm_aux), "r"(abuf), "r"(cc), "r"(t.m0), "r"(t.ob_lo), "r"(t.ob_hi) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          // every epilogue thread has read this slab's ring slot (it arrived after doing so): refill it
          if (in_tma) prefetch_one();
          // hand the staging buffer(s) back as soon as the store has read them (the epilogue warps are busy with the
          // next slab in the other buffer meanwhile)
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&out_empty[obuf]);
          if (++obuf == (uint32_t)p.n_out) {
            obuf = 0;
            ++ouse;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0-7)
    const int flags = FLAGS_CT >= 0 ? FLAGS_CT : p.flags;
    const bool out_f32 = (flags & FHB_EPI_OUT_F32) != 0;
    const bool out_f16 = (flags & FHB_EPI_OUT_BF16) == 0;                            // 16-bit D: fp16 unless flagged bf16
    const bool aux_f16 = out_f16 || (flags & FHB_EPI_AUX_DGELU) != 0;                 // a saved gelu' is always fp16
    const bool two_out = (flags & FHB_EPI_STORE_PREACT) != 0;
    const bool in_tma = EPI_IN && p.n_in > 0;
    // which [m][n] operand travels through the TMA input ring: aux_in if it is used, else the residual
    const bool ring_is_aux = (flags & (FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX)) != 0;
    const int slab_cols = out_f32 ? 32 : 64;       // columns per 128-byte slab
    const int my_cols = slab_cols >> 1;            // this warp's half of the slab: 16 (fp32) or 32 (bf16)
    const int quarter = warp & 3, half = warp >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t rsw = (uint32_t)(row_in_tile & 7);
    const uint32_t out_u32 = smem_u32(smem_out), auxo_u32 = smem_u32(smem_auxo), in_u32 = smem_u32(smem_in);
    auto tile_slabs = [&](const Tile& t) {
      return p.use_tma_store ? (min(p.bn, p.n - t.n0) + slab_cols - 1) / slab_cols : 0;
    };
    uint32_t in_slot = 0, in_phase = 0;  // consumer side of the input ring (the store warp refills it)
    int it = 0;
    uint32_t slab_ctr = 0, obuf = 0, ouse = 0;  // obuf: staging buffer of the current slab (round robin over p.n_out)
#ifdef FHB_GEMM_TRACE
    const bool trace_on = g_gemm_trace != nullptr && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 255);
    const int trace_who = threadIdx.x == 0 ? 0 : 1;
#endif
    float loss_local = 0.f;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_stride, ++it) {
      const Tile t = decode_tile(p, tile, pair_rank);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // stage this tile's bias slice (double buffered by tile parity; the slab barriers order it)
      float* bs = bias_s + as * kMaxBN;
      if (flags & FHB_EPI_BIAS) {
        const int c = threadIdx.x;
        if (c < p.bn) bs[c] = (t.n0 + c < p.n) ? __ldg(p.bias + (long long)t.ob_hi * p.bias_hi_stride + t.n0 + c) : 0.f;
      }
      FHB_TRACE(6);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      FHB_TRACE(7);
      if (threadIdx.x == 0 && it < 4) FHB_TL(6 + 4 * it);  // accumulator of tile `it` complete
      const int row = t.m0 + row_in_tile;
      const bool row_ok = row < p.m;
      bool zero_row = false;
      if ((flags & FHB_EPI_ROWZERO) && row_ok) zero_row = row >= p.row_valid[t.ob_hi];
      const long long off = (long long)t.ob_hi * p.d_hi_stride + (long long)t.ob_lo * p.d_lo_stride +
                            (long long)row * p.d_ld + t.n0;
      const uint32_t taddr = tmem_base + as * kMaxBN + ((uint32_t)(quarter * 32) << 16);
      const int n_slabs = tile_slabs(t);
      const int tile_cols = min(p.bn, p.n - t.n0);

      // fused epilogue math on 16 consecutive columns starting at tile column c.  `ring`: this thread's 16
      // values of the TMA-staged input operand (nullptr = read every [m][n] operand from global memory).
      auto math16 = [&](float* v, float* pre, int c, const uint32_t* ring) {
        if (flags & FHB_EPI_ALPHA) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= p.alpha;
        }
        if (flags & FHB_EPI_BIAS) {
          const float4* bp = reinterpret_cast<const float4*>(bs + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = bp[j];
            v[4 * j] += b4.x;
            v[4 * j + 1] += b4.y;
            v[4 * j + 2] += b4.z;
            v[4 * j + 3] += b4.w;
          }
        }
        if ((flags & (FHB_EPI_GELU | FHB_EPI_AUX_DGELU)) == (FHB_EPI_GELU | FHB_EPI_AUX_DGELU)) {
#pragma unroll
#ifdef FHB_GELU_SCALAR
          for (int j = 0; j < 16; ++j) gelu_erf_both(v[j], v[j], pre[j]);
#else
          for (int j = 0; j < 16; j += 2) gelu_erf_both2(v[j], v[j + 1], v[j], v[j + 1], pre[j], pre[j + 1]);
#endif
        } else {
          if (flags & FHB_EPI_STORE_PREACT) {
            if (flags & FHB_EPI_AUX_DGELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pre[j] = gelu_erf_grad(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pre[j] = v[j];
            }
          }
          if (flags & FHB_EPI_GELU) {
#pragma unroll
#ifdef FHB_GELU_SCALAR
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
#else
            for (int j = 0; j < 16; j += 2) gelu_erf2(v[j], v[j + 1], v[j], v[j + 1]);
#endif
          }
        }
        if (flags & FHB_EPI_DROPOUT) {
          // element index = ((ob * m) + row) * n + column; one hash per column pair
          const uint32_t pair0 = (uint32_t)((((long long)(t.ob_hi * p.ob_mod + t.ob_lo) * p.m + row) * p.n + t.n0 + c) >> 1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float m0, m1;
            dropout_pair(p.drop_seed, pair0 + j, p.drop_thr, p.drop_scale, m0, m1);
            v[2 * j] *= m0;
            v[2 * j + 1] *= m1;
            if (flags & FHB_EPI_AUX_DGELU) {
              pre[2 * j] *= m0;
              pre[2 * j + 1] *= m1;
            }
          }
        }
        if (EPI_IN) {
          const bool col_ok = row_ok && (t.n0 + c < p.n);
          if (flags & (FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX)) {
            uint32_t uu[8];
            if (ring) {
#pragma unroll
              for (int j = 0; j < 8; ++j) uu[j] = ring[j];
            } else if (col_ok) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.aux_in + off + c);
              const uint4 u0 = __ldg(ap), u1 = __ldg(ap + 1);
              uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w;
              uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) uu[j] = 0u;
            }
            if (flags & FHB_EPI_MUL_DGELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f = unpack_f16(uu[j]);
                v[2 * j] *= gelu_erf_grad(f.x);
                v[2 * j + 1] *= gelu_erf_grad(f.y);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f = unpack_f16(uu[j]);
                v[2 * j] *= f.x;
                v[2 * j + 1] *= f.y;
              }
            }
          }
          if (flags & FHB_EPI_RESIDUAL) {
            if (flags & FHB_EPI_RES_F32) {
              // fp32 residual stream: 16 floats per thread, through the ring when D is fp32 too (same slab geometry)
              if (ring && !ring_is_aux) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(ring[j]);
              } else if (col_ok) {
                const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + off + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 f = __ldg(rp + j);
                  v[4 * j] += f.x;
                  v[4 * j + 1] += f.y;
                  v[4 * j + 2] += f.z;
                  v[4 * j + 3] += f.w;
                }
              }
            } else {
              uint32_t uu[8];
              if (ring && !ring_is_aux) {
#pragma unroll
                for (int j = 0; j < 8; ++j) uu[j] = ring[j];
              } else if (col_ok) {
                const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + off + c);
                const uint4 u0 = __ldg(rp), u1 = __ldg(rp + 1);
                uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w;
                uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) uu[j] = 0u;
              }
              const bool res_f16 = (flags & FHB_EPI_RES_BF16) == 0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f = unpack16(uu[j], res_f16);
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
              }
            }
          }
          if ((flags & FHB_EPI_SQDIFF) && col_ok) {
            const uint4* tp = reinterpret_cast<const uint4*>(p.loss_target + off + c);
            const uint4 u0 = __ldg(tp), u1 = __ldg(tp + 1);
            const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 f = unpack_f16(uu[j]);
              const float d0 = v[2 * j] - f.x, d1 = v[2 * j + 1] - f.y;
              loss_local += d0 * d0 + d1 * d1;
              v[2 * j] = d0 * p.grad_scale;
              v[2 * j + 1] = d1 * p.grad_scale;
            }
          }
        }
        if (zero_row) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
      };

      // ---- 128-byte slabs through shared memory + TMA (the last slab of a row of tiles may hang over the
      //      tensor edge: TMA clips the store and zero-fills the load)
      for (int sidx = 0; sidx < n_slabs; ++sidx) {
        // staging buffers (round robin): D in out[obuf], the optional second output in auxo[obuf].  The store warp hands a
        // buffer back once the TMA store of its previous slab has read it (parity trick: the first use passes at once).
        const uint32_t mybuf = obuf;
        const uint32_t dbuf = out_u32 + mybuf * kStoreBytes;
        const uint32_t abuf = auxo_u32 + mybuf * kStoreBytes;
        mbar_wait(&out_empty[mybuf], (ouse & 1u) ^ 1u);
        if (++obuf == (uint32_t)p.n_out) {
          obuf = 0;
          ++ouse;
        }
        const int c0 = sidx * slab_cols + half * my_cols;  // first tile column of this warp's share
        uint32_t r[32];
        FHB_TRACE(0);
        tmem_ld16(taddr + c0, r);
        if (!out_f32) tmem_ld16(taddr + c0 + 16, r + 16);
        uint32_t ring[16];
        if (in_tma) {
          const uint32_t slot = in_slot;
          mbar_wait(&in_full[slot], in_phase);
          if (++in_slot == (uint32_t)p.n_in) {
            in_slot = 0;
            in_phase ^= 1u;
          }
          const uint32_t irow = in_u32 + slot * kStoreBytes + row_in_tile * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t ch = (uint32_t)(half * 4 + q);
            ld_shared_v4(irow + ((ch ^ rsw) << 4), ring[4 * q], ring[4 * q + 1], ring[4 * q + 2], ring[4 * q + 3]);
          }
        }
        tmem_ld_wait();
        FHB_TRACE(1);
        const uint32_t drow = dbuf + row_in_tile * 128;
        const uint32_t arow = abuf + row_in_tile * 128;
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          if (gq == 1 && out_f32) break;
          float v[16], pre[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[gq * 16 + j]);
          math16(v, pre, c0 + gq * 16, in_tma ? ring + gq * 8 : nullptr);
          if (out_f32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 4 chunks of 4 floats; chunk index within the 128-byte row
              const uint32_t ch = (uint32_t)(half * 4 + q);
              st_shared_v4(drow + ((ch ^ rsw) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                           __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
            }
          } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {  // 2 chunks of 8 bf16
              const uint32_t ch = (uint32_t)(half * 4 + gq * 2 + q);
              st_shared_v4(drow + ((ch ^ rsw) << 4), pack16(v[8 * q], v[8 * q + 1], out_f16), pack16(v[8 * q + 2], v[8 * q + 3], out_f16),
                           pack16(v[8 * q + 4], v[8 * q + 5], out_f16), pack16(v[8 * q + 6], v[8 * q + 7], out_f16));
              if (two_out)
                st_shared_v4(arow + ((ch ^ rsw) << 4), pack16(pre[8 * q], pre[8 * q + 1], aux_f16),
                             pack16(pre[8 * q + 2], pre[8 * q + 3], aux_f16), pack16(pre[8 * q + 4], pre[8 * q + 5], aux_f16),
                             pack16(pre[8 * q + 6], pre[8 * q + 7], aux_f16));
            }
          }
        }
        FHB_TRACE(2);
        fence_async_shared();  // this thread's shared-memory writes are visible to the async proxy (the TMA store) ...
        FHB_TRACE(3);
        mbar_arrive(&out_full[mybuf]);  // ... before the store warp sees the slab complete
        FHB_TRACE(4);
        FHB_TRACE(5);
        ++slab_ctr;
      }

      // ---- TMA store disabled (FHB_GEMM_DIRECT_STORE): direct stores, 16-column groups dealt round-robin to
      //      the two warps of a lane quarter
      for (int c = n_slabs * slab_cols + half * 16; c < tile_cols; c += 32) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        tmem_ld_wait();
        float v[16], pre[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        math16(v, pre, c, nullptr);
        if (!row_ok || t.n0 + c >= p.n) continue;
        if (flags & FHB_EPI_STORE_PREACT) {
          uint4* ap = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.aux_out) + off + c);
          ap[0] = make_uint4(pack16(pre[0], pre[1], aux_f16), pack16(pre[2], pre[3], aux_f16), pack16(pre[4], pre[5], aux_f16),
                             pack16(pre[6], pre[7], aux_f16));
          ap[1] = make_uint4(pack16(pre[8], pre[9], aux_f16), pack16(pre[10], pre[11], aux_f16), pack16(pre[12], pre[13], aux_f16),
                             pack16(pre[14], pre[15], aux_f16));
        }
        if (out_f32) {
          float* dp = reinterpret_cast<float*>(p.d) + off + c;
          if (flags & FHB_EPI_ATOMIC_ADD) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dp + j, v[j]);
          } else {
            float4* d4 = reinterpret_cast<float4*>(dp);
#pragma unroll
            for (int j = 0; j < 4; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.d) + off + c);
          dp[0] = make_uint4(pack16(v[0], v[1], out_f16), pack16(v[2], v[3], out_f16), pack16(v[4], v[5], out_f16),
                             pack16(v[6], v[7], out_f16));
          dp[1] = make_uint4(pack16(v[8], v[9], out_f16), pack16(v[10], v[11], out_f16), pack16(v[12], v[13], out_f16),
                             pack16(v[14], v[15], out_f16));
        }
      }
      if (threadIdx.x == 0 && it < 4) FHB_TL(7 + 4 * it);  // epilogue of tile `it` done (this thread)
      tc_fence_before();
      if (cg2) {
        // the leader's MMA warp waits for both CTAs' epilogues: ONE (remote) arrival per CTA behind a block barrier instead of
        // 256 arrivals across the cluster
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) mbar_arrive_cluster(smem_u32(&acc_empty[as]) & kLeaderMask);
      } else {
        mbar_arrive(&acc_empty[as]);
      }
    }
    if (EPI_IN && (flags & FHB_EPI_SQDIFF)) {
      loss_local = warp_sum(loss_local);
      if (lane == 0 && loss_local != 0.f) atomicAdd(p.loss_acc, loss_local * p.loss_weight);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FHB_TL(3);  // every warp done (the store warp has issued and drained its last store)
  if (threadIdx.x == 0) FHB_CTA_STAMP(1);
#ifdef FHB_GEMM_TRACE
  if (log_slot < 8192u) {
    g_gemm_log[6 * log_slot + 1] = fhb_globaltimer();
    g_gemm_log[6 * log_slot + 5] = clock64();
  }
#endif
  if (p.pair) cluster_sync_all();  // the peer may still be arriving on this CTA's empty barriers
  if (warp == 9) {
    tc_fence_after();
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// cuTensorMapEncodeTiled costs several microseconds of driver time; a training step issues the same few
// hundred (pointer, shape) combinations every iteration (the caching allocator returns the same blocks), so
// encoded maps are memoised.  Key = every input of the encode call.
struct TmapKey {
  const void* ptr;
  long long d[4];
  long long s[3];
  unsigned box0, box1, kind;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&k);
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmap_cache.find(k);
  if (it == g_tmap_cache.end()) return false;
  *out = it->second;
  return true;
}
void tmap_insert(const TmapKey& k, const CUtensorMap& v) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
  g_tmap_cache.emplace(k, v);
}

int make_tmap(CUtensorMap* tm, const fhb_tensor3& t, uint32_t box0, uint32_t box1, const char* name) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = t.ptr;
  key.d[0] = t.dim[0]; key.d[1] = t.dim[1]; key.d[2] = t.dim[2];
  key.s[0] = t.stride[0]; key.s[1] = t.stride[1];
  key.box0 = box0; key.box1 = box1; key.kind = 1;
  if (tmap_lookup(key, tm)) return 0;
  EncodeTiledFn enc = get_encode_fn();
  FHB_ARG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  FHB_ARG_CHECK(t.ptr != nullptr && (reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "gemm: %s pointer must be 16B aligned", name);
  FHB_ARG_CHECK(t.dim[0] > 0 && t.dim[1] > 0 && t.dim[2] > 0, "gemm: %s has an empty dimension", name);
  FHB_ARG_CHECK(t.stride[0] % 8 == 0 && t.stride[0] > 0, "gemm: %s row stride %lld must be a positive multiple of 8 elements",
                name, (long long)t.stride[0]);
  cuuint64_t dims[3] = {(cuuint64_t)t.dim[0], (cuuint64_t)t.dim[1], (cuuint64_t)t.dim[2]};
  int64_t s2 = t.stride[1];
  if (t.dim[2] == 1 && (s2 <= 0 || s2 % 8 != 0)) s2 = t.stride[0] * t.dim[1];
  FHB_ARG_CHECK(s2 % 8 == 0 && s2 > 0, "gemm: %s batch stride %lld must be a positive multiple of 8 elements", name,
                (long long)s2);
  cuuint64_t strides[2] = {(cuuint64_t)t.stride[0] * 2, (cuuint64_t)s2 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(t.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FHB_ARG_CHECK(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %lld,%lld,%lld strides %lld,%lld box %u,%u)",
                name, (int)r, (long long)t.dim[0], (long long)t.dim[1], (long long)t.dim[2], (long long)t.stride[0],
                (long long)s2, box0, box1);
  tmap_insert(key, *tm);
  return 0;
}

// Output tensor map: {columns, rows, ob_lo, ob_hi}; box = one 128-byte slab x 128 rows, 128B swizzle.
int make_out_tmap(CUtensorMap* tm, void* base, bool f32, int n, int m, int ob_mod, int n_hi, long long ld,
                  long long lo_stride, long long hi_stride, const char* name) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = base;
  key.d[0] = n; key.d[1] = m; key.d[2] = ob_mod; key.d[3] = n_hi;
  key.s[0] = ld; key.s[1] = lo_stride; key.s[2] = hi_stride;
  key.kind = f32 ? 3 : 2;
  if (tmap_lookup(key, tm)) return 0;
  EncodeTiledFn enc = get_encode_fn();
  FHB_ARG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const int elt = f32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)m, (cuuint64_t)ob_mod, (cuuint64_t)n_hi};
  long long s1 = ld, s2 = lo_stride > 0 ? lo_stride : ld * m, s3 = hi_stride > 0 ? hi_stride : s2 * ob_mod;
  cuuint64_t strides[3] = {(cuuint64_t)s1 * elt, (cuuint64_t)s2 * elt, (cuuint64_t)s3 * elt};
  cuuint32_t box[4] = {(cuuint32_t)(128 / elt), (cuuint32_t)kBM, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FHB_ARG_CHECK(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled(%s) failed with CUresult %d (n=%d m=%d ld=%lld lo=%lld hi=%lld)",
                name, (int)r, n, m, ld, s2, s3);
  tmap_insert(key, *tm);
  return 0;
}

// Tile width.  One n-block when the row fits (any multiple of 16); otherwise a multiple of 64 so that the
// 128-byte store slabs of neighbouring n-blocks never overlap (only the last block of a row is clipped by
// TMA).  Among {64,128,192,256} pick the width that minimises (rounds over the SMs) x (tile width + a fixed
// per-tile cost worth ~48 columns): N = 480 on 98 row blocks runs as 192+192+96 (294 tiles, 2 rounds)
// instead of 256+224 (196 tiles, still 2 rounds of wider tiles).
// `step`: the granularity of a multi-block tile width = the column count of one 128-byte output slab (64 for 16-bit
// outputs; 32 for fp32 outputs, which lets N = 480 run as three EQUAL 160-column tiles: every CTA of the two-round
// schedule then owns 320 columns instead of up to 384).
int pick_bn(int n, long long row_tiles, bool split_k, int step = 64) {
  const int n16 = (n + 15) / 16 * 16;
  if (n16 <= 64) return n16;
  static const int forced = getenv("FHB_GEMM_BN") ? atoi(getenv("FHB_GEMM_BN")) : 0;  // tuning sweeps only
  if (forced >= 64 && forced % 64 == 0 && forced <= kMaxBN) return n16 <= forced ? n16 : forced;
  const int sms = fhb_num_sms();
  int best = kMaxBN;
  double best_cost = 1e30;
  for (int bn = kMaxBN; bn >= 64; bn -= step) {
    const int nblk = (n + bn - 1) / bn;
    const int w = nblk == 1 ? n16 : bn;
    const long long tiles = row_tiles * nblk;
    const double rounds = split_k ? (double)nblk : (double)((tiles + sms - 1) / sms);
    const double cost = rounds * (w + 48);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = w;
    }
  }
  return best;
}

template <int A_MN, int B_MN, int EPI_IN, int CG2 = 0, int FLAGS_CT = -1>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx,
            const CUtensorMap& ti, const GemmParams& p, cudaStream_t s) {
  FHB_ONCE_PER_DEVICE(FHB_CUDA_CHECK(cudaFuncSetAttribute(fhb_gemm_kernel<A_MN, B_MN, EPI_IN, CG2, FLAGS_CT>,
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
  int grid = p.total_tiles < fhb_num_sms() ? p.total_tiles : fhb_num_sms();
  if (p.pair) {
    // clusters of two CTAs: as many as can be co-resident (one CTA per SM; a GPC with an odd number of free SMs strands one)
    static int max_clusters[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& mc = max_clusters[dev & 63];
    if (mc == 0) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(2 * 148);
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = kSmemBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fhb_gemm_kernel<A_MN, B_MN, EPI_IN, CG2, FLAGS_CT>, &q) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = 64;
      }
      mc = n;
      if (getenv("FHB_GEMM_DEBUG")) fprintf(stderr, "fhb_gemm: %d co-resident 2-CTA clusters\n", n);
    }
    int clusters = mc < fhb_num_sms() / 2 ? mc : fhb_num_sms() / 2;
    if (clusters > p.total_tiles) clusters = p.total_tiles;
    grid = 2 * clusters;
  }
  // "small" = at most ~200 k-blocks per SM: the student's GEMMs and the teacher's encoder GEMMs, not the conv stacks
  // (swept on B200, profiles/r01y_pdl_ab.log: 7 104 -> 23.61 ms, 30 000 -> 23.55 ms, 150 000 -> 23.68 ms, none -> 23.94 ms)
  static const long long pdl_limit = getenv("FHB_PDL_GEMM_LIMIT") ? atoll(getenv("FHB_PDL_GEMM_LIMIT")) : 30000;
  fhb_pdl_hint((long long)p.total_tiles * (p.kb_per_split < 1 ? 1 : p.kb_per_split) <= pdl_limit);
  if (p.pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = fhb_pdl_enabled() ? 2 : 1;
    FHB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fhb_gemm_kernel<A_MN, B_MN, EPI_IN, CG2, FLAGS_CT>, ta, tb, td, tx, ti, p));
  } else {
    FHB_CUDA_CHECK(fhb_launch((fhb_gemm_kernel<A_MN, B_MN, EPI_IN, CG2, FLAGS_CT>), dim3(grid), dim3(kThreads), kSmemBytes, s, ta, tb, td, tx, ti, p));
  }
  FHB_LAUNCH_CHECK();
  return 0;
}

constexpr int kEpiInFlags = FHB_EPI_RESIDUAL | FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX | FHB_EPI_SQDIFF;
template <int A_MN, int B_MN, int CG2, int F>
int launch_spec(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx,
                const CUtensorMap& ti, const GemmParams& p, cudaStream_t s) {
  return launch2<A_MN, B_MN, (F & kEpiInFlags) != 0, CG2, F>(ta, tb, td, tx, ti, p, s);
}
#define FHB_SPEC_CASE(A, B, C, F) \
  case F: return launch_spec<A, B, C, F>(ta, tb, td, tx, ti, p, s)

template <int A_MN, int B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx,
           const CUtensorMap& ti, const GemmParams& p, cudaStream_t s) {
  // The flag sets the distillation step spends its time in (profiles/r03final_gemm_table.txt) run flag-specialised
  // instantiations; every other combination takes the generic kernels below.  FHB_GEMM_NOSPEC=1: generic only.
  static const bool nospec = getenv("FHB_GEMM_NOSPEC") != nullptr;
  if (!nospec) {
    constexpr int B_ = FHB_EPI_BIAS, G_ = FHB_EPI_GELU, R_ = FHB_EPI_RESIDUAL, P_ = FHB_EPI_STORE_PREACT, O32 = FHB_EPI_OUT_F32,
                  AT = FHB_EPI_ATOMIC_ADD, DG = FHB_EPI_AUX_DGELU, MA = FHB_EPI_MUL_AUX, DR = FHB_EPI_DROPOUT, R32 = FHB_EPI_RES_F32,
                  AL = FHB_EPI_ALPHA;
    if constexpr (A_MN == 0 && B_MN == 0) {
      if (p.pair == 2) {
        switch (p.flags) {
          FHB_SPEC_CASE(0, 0, 1, 0);        // positional-conv GEMMs (K = 4224 / 6336)
          FHB_SPEC_CASE(0, 0, 1, G_);       // teacher conv stack (k = 3: K = 1536)
          FHB_SPEC_CASE(0, 0, 1, B_ | R_);  // teacher fc2 + residual
          default: break;
        }
      } else if (p.pair == 0) {
        switch (p.flags) {
          FHB_SPEC_CASE(0, 0, 0, 0);                       // plain products (k = 2 conv dgrads, head chain rule)
          FHB_SPEC_CASE(0, 0, 0, B_);                      // q | k | v projections, heads
          FHB_SPEC_CASE(0, 0, 0, B_ | G_);                 // teacher fc1
          FHB_SPEC_CASE(0, 0, 0, G_);                      // conv layers, inference
          FHB_SPEC_CASE(0, 0, 0, G_ | P_ | DG);            // student conv layers, training (saved gelu')
          FHB_SPEC_CASE(0, 0, 0, B_ | G_ | P_ | DG);       // student fc1, p = 0
          FHB_SPEC_CASE(0, 0, 0, B_ | G_ | P_ | DG | DR);  // student fc1 with activation dropout
          FHB_SPEC_CASE(0, 0, 0, B_ | R_);                 // teacher out_proj + fp16 residual
          FHB_SPEC_CASE(0, 0, 0, B_ | R_ | R32 | O32);       // student out_proj / fc2, fp32 stream
          FHB_SPEC_CASE(0, 0, 0, B_ | R_ | R32 | O32 | DR);  // ... with dropout
          FHB_SPEC_CASE(0, 0, 0, MA);                      // k = 1 conv dgrad x saved gelu'
          default: break;
        }
      }
    } else if constexpr (A_MN == 0 && B_MN == 1) {
      if (p.pair == 0) {
        switch (p.flags) {
          FHB_SPEC_CASE(0, 1, 0, 0);               // plain dgrads
          FHB_SPEC_CASE(0, 1, 0, MA);              // dgrad x saved gelu'
          FHB_SPEC_CASE(0, 1, 0, R_ | R32 | O32);  // dgrad + fp32 residual gradient
          FHB_SPEC_CASE(0, 1, 0, MA | R_ | R32 | O32);
          default: break;
        }
      }
    } else {
      if (p.pair == 0) {
        switch (p.flags) {
          FHB_SPEC_CASE(1, 1, 0, AT | O32);  // weight gradients (split-K, fp32 reduce-add)
          FHB_SPEC_CASE(1, 1, 0, O32);       // positional-conv weight gradient (one split)
          FHB_SPEC_CASE(1, 1, 0, AL);        // folded-head weight gradient (fp16, scaled)
          FHB_SPEC_CASE(1, 1, 0, AT | O32 | AL);
          default: break;
        }
      }
    }
  }
  const bool epi_in = (p.flags & kEpiInFlags) != 0;
  if constexpr (A_MN == 0 && B_MN == 0) {
    if (p.pair == 2) return epi_in ? launch2<0, 0, 1, 1>(ta, tb, td, tx, ti, p, s) : launch2<0, 0, 0, 1>(ta, tb, td, tx, ti, p, s);
  }
  if (epi_in) return launch2<A_MN, B_MN, 1>(ta, tb, td, tx, ti, p, s);
  return launch2<A_MN, B_MN, 0>(ta, tb, td, tx, ti, p, s);
}

}  // namespace

// external-linkage wrapper for the other translation units (attention_tc.cu)
int fhb_make_tmap_bf16_3d(CUtensorMap* tm, const void* ptr, const int64_t dim[3], const int64_t stride[2],
                          uint32_t box0, uint32_t box1, const char* name) {
  fhb_tensor3 t;
  t.ptr = ptr;
  for (int i = 0; i < 3; ++i) t.dim[i] = dim[i];
  t.stride[0] = stride[0];
  t.stride[1] = stride[1];
  return make_tmap(tm, t, box0, box1, name);
}

// 16-bit 4-D map {d0, d1, d2, d3} (strides in elements for d1..d3), box {box0, 1, box2, 1}, 128B swizzle.  box0 may
// exceed d0: the out-of-range columns are zero-filled by TMA (attention_tc.cu loads 40-wide heads into 64-wide tiles).
int fhb_make_tmap_bf16_4d(CUtensorMap* tm, const void* ptr, const int64_t dim[4], const int64_t stride[3], uint32_t box0,
                          uint32_t box2, const char* name) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  for (int i = 0; i < 4; ++i) key.d[i] = dim[i];
  for (int i = 0; i < 3; ++i) key.s[i] = stride[i];
  key.box0 = box0; key.box1 = box2; key.kind = 5;
  if (tmap_lookup(key, tm)) return 0;
  EncodeTiledFn enc = get_encode_fn();
  FHB_ARG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  FHB_ARG_CHECK(ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && stride[0] % 8 == 0 && stride[1] % 8 == 0 &&
                    stride[2] % 8 == 0 && box0 * 2 <= 128 && box2 <= 256,
                "tensor map %s: 16-bit maps need 16-byte aligned base / strides and a box row of at most 128 bytes", name);
  cuuint64_t dims[4] = {(cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2], (cuuint64_t)dim[3]};
  cuuint64_t strides[3] = {(cuuint64_t)stride[0] * 2, (cuuint64_t)stride[1] * 2, (cuuint64_t)stride[2] * 2};
  cuuint32_t box[4] = {box0, 1, box2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FHB_ARG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", name, (int)r);
  tmap_insert(key, *tm);
  return 0;
}

extern "C" int fhb_gemm(const fhb_gemm_args* a, fhb_stream_t stream) {
  FHB_ARG_CHECK(a != nullptr, "gemm: null args");
  FHB_ARG_CHECK(a->m > 0 && a->n > 0 && a->k > 0, "gemm: m,n,k must be positive (got %d,%d,%d)", a->m, a->n, a->k);
  FHB_ARG_CHECK(a->n % 16 == 0, "gemm: n=%d must be a multiple of 16", a->n);
  FHB_ARG_CHECK(a->a_major == 0 || a->a_major == 1, "gemm: a_major must be 0 or 1");
  FHB_ARG_CHECK(a->b_major == 0 || a->b_major == 1, "gemm: b_major must be 0 or 1");
  FHB_ARG_CHECK(!(a->a_major == 1 && a->b_major == 0), "gemm: (A MN-major, B K-major) is not instantiated");
  FHB_ARG_CHECK(a->d != nullptr && a->d_ld > 0, "gemm: null output / bad d_ld");
  const int num_ob = a->num_ob > 0 ? a->num_ob : 1;
  const int ob_mod = a->ob_mod > 0 ? a->ob_mod : 1;
  const int num_cb = a->num_cb > 0 ? a->num_cb : 1;
  FHB_ARG_CHECK(num_cb == 1 || a->b_major == 1, "gemm: num_cb > 1 needs MN-major operands");
  const int flags = a->flags;
  FHB_ARG_CHECK(!(flags & FHB_EPI_BIAS) || a->bias, "gemm: FHB_EPI_BIAS without bias");
  FHB_ARG_CHECK(!(flags & FHB_EPI_RESIDUAL) || a->residual, "gemm: FHB_EPI_RESIDUAL without residual");
  FHB_ARG_CHECK(!(flags & (FHB_EPI_RES_F32 | FHB_EPI_RES_BF16)) || (flags & FHB_EPI_RESIDUAL), "gemm: FHB_EPI_RES_* needs FHB_EPI_RESIDUAL");
  FHB_ARG_CHECK((flags & (FHB_EPI_RES_F32 | FHB_EPI_RES_BF16)) != (FHB_EPI_RES_F32 | FHB_EPI_RES_BF16), "gemm: one residual type");
  FHB_ARG_CHECK((flags & (FHB_EPI_OUT_F32 | FHB_EPI_OUT_BF16)) != (FHB_EPI_OUT_F32 | FHB_EPI_OUT_BF16), "gemm: one output type");
  FHB_ARG_CHECK(!(flags & FHB_GEMM_A_BF16) == !(flags & FHB_GEMM_B_BF16),
                "gemm: A and B must share one 16-bit format (tcgen05 kind::f16 faults on fp16 x bf16)");
  FHB_ARG_CHECK(!(flags & FHB_EPI_STORE_PREACT) || !(flags & FHB_EPI_OUT_F32), "gemm: a second output needs a bf16 D");
  FHB_ARG_CHECK(!(flags & FHB_EPI_ROWZERO) || a->row_valid, "gemm: FHB_EPI_ROWZERO without row_valid");
  FHB_ARG_CHECK(!(flags & FHB_EPI_STORE_PREACT) || a->aux_out, "gemm: FHB_EPI_STORE_PREACT without aux_out");
  FHB_ARG_CHECK(!(flags & (FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX)) || a->aux_in, "gemm: FHB_EPI_MUL_DGELU/MUL_AUX without aux_in");
  FHB_ARG_CHECK((flags & (FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX)) != (FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX),
                "gemm: FHB_EPI_MUL_DGELU and FHB_EPI_MUL_AUX are exclusive");
  FHB_ARG_CHECK(!(flags & FHB_EPI_AUX_DGELU) || (flags & FHB_EPI_STORE_PREACT), "gemm: FHB_EPI_AUX_DGELU needs FHB_EPI_STORE_PREACT");
  FHB_ARG_CHECK(!(flags & FHB_EPI_SQDIFF) || (a->loss_target && a->loss_acc), "gemm: FHB_EPI_SQDIFF without target/acc");
  FHB_ARG_CHECK(!(flags & FHB_EPI_ATOMIC_ADD) || (flags & FHB_EPI_OUT_F32), "gemm: atomic accumulate needs fp32 output");
  const int elt = (flags & FHB_EPI_OUT_F32) ? 4 : 2;
  FHB_ARG_CHECK((reinterpret_cast<uintptr_t>(a->d) & 15) == 0 && (a->d_ld * elt) % 16 == 0 &&
                    (a->d_hi_stride * elt) % 16 == 0 && (a->d_lo_stride * elt) % 16 == 0,
                "gemm: output must be 16B aligned in every stride");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m = a->m;
  p.n = a->n;
  p.k = a->k;
  p.num_m_blk = (a->m + kBM - 1) / kBM;
  static const bool bn32 = getenv("FHB_GEMM_BN32") == nullptr || atoi(getenv("FHB_GEMM_BN32")) != 0;  // A/B switch
  const bool f32_slabs = bn32 && (flags & FHB_EPI_OUT_F32) && !(flags & (FHB_EPI_STORE_PREACT | FHB_EPI_ATOMIC_ADD));
  p.bn = pick_bn(a->n, (long long)p.num_m_blk * num_ob, (flags & FHB_EPI_ATOMIC_ADD) != 0, f32_slabs ? 32 : 64);
  p.num_n_blk = (a->n + p.bn - 1) / p.bn;
  // Pair modes (see the kernel header): forward-shaped GEMMs (both operands K-major, no split-K) with at least two row
  // blocks run as clusters of two CTAs.  FHB_GEMM_PAIR=1: the B tile is shared by TMA multicast (L2 -> SM traffic of B
  // halved); FHB_GEMM_PAIR=2: cta_group::2 - one UMMA of M = 256 over the pair, B split between the two CTAs, and the
  // shared memory that frees buys a 6-stage pipeline; 0: one CTA per tile.  Default (unset): cta_group::2 for LONG
  // contractions only (K >= 1536).  Measured per shape inside the step (profiles/r02zz_gemm_pair_ab.txt): cta_group::2 wins
  // where the main loop dominates - fc2 K = 3072: 111.6 -> 100.1 us, the teacher's k=3 conv layers K = 1536: 1001 -> 899 us,
  // pos-conv K = 4224 / 6336: -25 % - because a 6 x 32 KB ring covers ~1.6 us of L2 latency at the full MMA rate where
  // 4 x 48 KB cover 1.1 us; it loses 5-10 % on short contractions and heavy epilogues (fc1 + GELU N = 3072, out_proj K = 768,
  // every HBM-bound conv layer), where the pair's lock step and cluster barriers cost more than the ring buys.  TMA
  // multicast alone changes nothing: operand bandwidth is not the bound.
  static const int pair_env = getenv("FHB_GEMM_PAIR") ? atoi(getenv("FHB_GEMM_PAIR")) : -1;
  int pair_mode = pair_env < 0 ? 2 : pair_env;
  if (pair_env < 0 && (a->k + kBK - 1) / kBK < 24) pair_mode = 0;
  const bool pair_on = pair_mode == 1 || pair_mode == 2;
  const int m_blocks = p.num_m_blk;
  if (pair_on && a->a_major == 0 && a->b_major == 0 && !(flags & FHB_EPI_ATOMIC_ADD) && a->split_k <= 1 && m_blocks >= 2 &&
      p.bn % 16 == 0 && (long long)((m_blocks + 1) / 2) * p.num_n_blk * num_ob >= fhb_num_sms() / 2) {
    p.pair = pair_mode;
    // cta_group::2 needs 32 <= N <= 256 for every n-block (the last one may be narrower)
    const int last_n = a->n - (p.num_n_blk - 1) * p.bn;
    if (p.pair == 2 && (((last_n + 15) & ~15) < 32 || p.bn < 32)) p.pair = 1;
    p.num_m_blk = (m_blocks + 1) / 2;  // decode_tile counts PAIRS of row blocks
  }
  static const bool dbg = getenv("FHB_GEMM_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "fhb_gemm m=%d n=%d k=%d bn=%d a_major=%d b_major=%d ob=%d pair=%d\n", a->m, a->n, a->k, p.bn, a->a_major, a->b_major, num_ob, p.pair);
  p.num_ob = num_ob;
  p.ob_mod = ob_mod;
  p.kb_per_cb = (a->k + kBK - 1) / kBK;
  p.kb_total = p.kb_per_cb * num_cb;
  int split = a->split_k;
  const int base_tiles = p.num_m_blk * p.num_n_blk * num_ob;
  if (split <= 0) {
    split = 1;
    if (flags & FHB_EPI_ATOMIC_ADD) {  // wgrad: small output, long contraction -> about one wave of work,
      split = fhb_num_sms() / base_tiles;  // each split long enough (>= 8 k-blocks) to amortise its reduce-add
      if (split > p.kb_total / 8) split = p.kb_total / 8;
      if (split < 1) split = 1;
    }
  }
  FHB_ARG_CHECK(split == 1 || (flags & FHB_EPI_ATOMIC_ADD), "gemm: split_k > 1 needs FHB_EPI_ATOMIC_ADD");
  if (split > p.kb_total) split = p.kb_total;
  p.kb_per_split = (p.kb_total + split - 1) / split;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.total_tiles = base_tiles * p.split_k;
  p.a_lo_c0 = a->a_lo_c0; p.a_hi_c2 = a->a_hi_c2; p.a_lo_c2 = a->a_lo_c2; p.a_cb_c2 = a->a_cb_c2;
  p.b_lo_c0 = a->b_lo_c0; p.b_hi_c2 = a->b_hi_c2; p.b_lo_c2 = a->b_lo_c2; p.b_cb_c2 = a->b_cb_c2;
  p.a_c1_off = a->a_c1_off; p.b_c1_off = a->b_c1_off;
  p.d_ld = a->d_ld; p.d_hi_stride = a->d_hi_stride; p.d_lo_stride = a->d_lo_stride;
  p.d = a->d;
  p.bias_hi_stride = a->bias_hi_stride;
  p.bias = a->bias;
  p.residual = a->residual;
  p.aux_in = static_cast<const __half*>(a->aux_in);
  p.aux_out = a->aux_out;
  p.row_valid = a->row_valid;
  p.loss_target = static_cast<const __half*>(a->loss_target);
  p.loss_acc = a->loss_acc;
  p.loss_weight = a->loss_weight;
  p.grad_scale = a->grad_scale;
  p.alpha = a->alpha;
  if (flags & FHB_EPI_DROPOUT) {
    FHB_ARG_CHECK(a->drop_p >= 0.f && a->drop_p < 1.f, "gemm: drop_p=%f must be in [0, 1)", (double)a->drop_p);
    FHB_ARG_CHECK((long long)num_ob * a->m * a->n < (1LL << 32) && a->n % 2 == 0, "gemm: dropout index space exceeds 32 bits");
    p.drop_seed = a->drop_seed;
    p.drop_thr = fhb_dropout_thr16(a->drop_p);
    p.drop_scale = fhb_dropout_scale(a->drop_p);
  }
  p.flags = flags;

  CUtensorMap ta, tb;
  int rc;
  const int b_atoms = (p.bn + 63) / 64;
  if (a->a_major == 0) {
    if ((rc = make_tmap(&ta, a->a, kBK, kBM, "A")) != 0) return rc;
  } else {
    if ((rc = make_tmap(&ta, a->a, 64, kBK, "A")) != 0) return rc;
  }
  if (a->b_major == 0) {
    if ((rc = make_tmap(&tb, a->b, kBK, (uint32_t)(p.pair ? p.bn / 2 : p.bn), "B")) != 0) return rc;
    p.stage_tx_bytes = kABytes + (uint32_t)p.bn * kBK * 2;
  } else {
    if ((rc = make_tmap(&tb, a->b, 64, kBK, "B")) != 0) return rc;
    p.stage_tx_bytes = kABytes + (uint32_t)b_atoms * 64 * kBK * 2;
  }
  // TMA store path: needs 16B-aligned strides (already checked) and every (ob_lo, ob_hi) offset expressible as
  // a tensor-map stride; otherwise (or when FHB_GEMM_DIRECT_STORE is set) the epilogue stores directly.
  CUtensorMap td, tx, ti;
  memset(&td, 0, sizeof(td));
  memset(&tx, 0, sizeof(tx));
  memset(&ti, 0, sizeof(ti));
  static const bool force_direct = getenv("FHB_GEMM_DIRECT_STORE") != nullptr;
  p.use_tma_store = force_direct ? 0 : 1;
  const bool out_f32 = (flags & FHB_EPI_OUT_F32) != 0;
  // the [m][n] epilogue input (aux_in if used, else the residual) is staged through a TMA ring when D is bf16
  const bool ring_aux = (flags & (FHB_EPI_MUL_DGELU | FHB_EPI_MUL_AUX)) != 0;
  const void* ring_src = ring_aux ? a->aux_in : ((flags & FHB_EPI_RESIDUAL) ? a->residual : nullptr);
  // the ring's slabs have the geometry of D's: usable when the staged operand has D's element type (aux_in is always
  // bf16; the residual is bf16 or, with FHB_EPI_RES_F32, fp32); otherwise that operand is read from global memory
  const bool ring_f32 = !ring_aux && (flags & FHB_EPI_RES_F32) != 0;
  // staging depth of the output path: FHB_GEMM_NOUT = 2 | 3 | 4 (tuning; the pipeline keeps at least 2 stages)
  static const int nout_env = getenv("FHB_GEMM_NOUT") ? atoi(getenv("FHB_GEMM_NOUT")) : 2;
  p.n_out = nout_env < 2 ? 2 : (nout_env > 4 ? 4 : nout_env);
  p.n_auxout = (p.use_tma_store && (flags & FHB_EPI_STORE_PREACT)) ? p.n_out : 0;
  // ring depth: 3 slabs (2 in flight); cta_group::2 stages are 2 units, so a 4th slab fits beside 4 pipeline stages - the
  // teacher's out_proj (K = 768: a tile's 64 KB of residual against 3.2 us of MMA) is bound by this ring
  static const int ring_cg2 = getenv("FHB_GEMM_RING") ? atoi(getenv("FHB_GEMM_RING")) : 4;
  p.n_in = (p.use_tma_store && ring_src && ring_f32 == out_f32) ? (p.pair == 2 ? (ring_cg2 < 3 ? 3 : (ring_cg2 > kInRing ? kInRing : ring_cg2)) : 3) : 0;
  const int unit_per_stage = p.pair == 2 ? 2 : 3;  // cta_group::2: half B tiles, 2 units of 16 KiB per stage
  while (p.n_out > 2 && (14 - p.n_out - p.n_auxout - p.n_in) / unit_per_stage < 2) {
    --p.n_out;
    if (p.n_auxout) p.n_auxout = p.n_out;
  }
  p.stages = (14 - p.n_out - p.n_auxout - p.n_in) / unit_per_stage;
  if (p.stages > (p.pair == 2 ? kStages : 4)) p.stages = p.pair == 2 ? kStages : 4;
  if (p.use_tma_store) {
    const int n_hi = (num_ob + ob_mod - 1) / ob_mod;
    if ((rc = make_out_tmap(&td, a->d, out_f32, a->n, a->m, ob_mod, n_hi, a->d_ld, a->d_lo_stride, a->d_hi_stride,
                            "D")) != 0)
      return rc;
    tx = td;
    ti = td;
    if (flags & FHB_EPI_STORE_PREACT) {
      if ((rc = make_out_tmap(&tx, a->aux_out, false, a->n, a->m, ob_mod, n_hi, a->d_ld, a->d_lo_stride, a->d_hi_stride,
                              "aux_out")) != 0)
        return rc;
    }
    if (p.n_in) {
      if ((rc = make_out_tmap(&ti, const_cast<void*>(ring_src), ring_f32, a->n, a->m, ob_mod, n_hi, a->d_ld, a->d_lo_stride,
                              a->d_hi_stride, "epilogue input")) != 0)
        return rc;
    }
  } else {
    td = ta;
    tx = ta;
    ti = ta;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->a_major == 0 && a->b_major == 0) return launch<0, 0>(ta, tb, td, tx, ti, p, s);
  if (a->a_major == 0 && a->b_major == 1) return launch<0, 1>(ta, tb, td, tx, ti, p, s);
  return launch<1, 1>(ta, tb, td, tx, ti, p, s);
}

#ifdef FHB_GEMM_TRACE
// debug builds only (-DFHB_GEMM_TRACE, tools/gemm_trace.py): device buffer of 2 x 96 x 8 int64 clock stamps, or NULL
extern "C" int fhb_gemm_steplog_read(long long* host, int max_entries, int reset) {
  unsigned int n = 0;
  FHB_CUDA_CHECK(cudaDeviceSynchronize());
  FHB_CUDA_CHECK(cudaMemcpyFromSymbol(&n, g_gemm_log_n, sizeof(n)));
  if (n > 8192u) n = 8192u;
  if ((int)n > max_entries) n = (unsigned int)max_entries;
  if (host && n) FHB_CUDA_CHECK(cudaMemcpyFromSymbol(host, g_gemm_log, sizeof(long long) * 6 * n));
  if (reset) {
    const unsigned int zero = 0;
    FHB_CUDA_CHECK(cudaMemcpyToSymbol(g_gemm_log_n, &zero, sizeof(zero)));
  }
  return (int)n;
}
extern "C" int fhb_gemm_set_trace_buffer(long long* buf) {
  FHB_CUDA_CHECK(cudaMemcpyToSymbol(g_gemm_trace, &buf, sizeof(buf)));
  return 0;
}
#endif
