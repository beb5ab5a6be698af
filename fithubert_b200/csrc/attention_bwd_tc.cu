// K7 backward on 5th-gen tensor cores (head_dim 40: FitHuBERT student; 64 also instantiated).
// Gradient of fairseq MultiheadAttention's bmm -> masked softmax -> (dropout) -> bmm chain
// (modules/module.py:558-564), one fused kernel for dQ, dK and dV:
//
// One CTA per (128-key tile, head, sample), 544 threads, 1 CTA per SM, loop over 128-query tiles i:
//   warp 16 (one elected lane): TMA loads (K, V once; Q_i, dO_i double buffered) and every tcgen05.mma:
//        S^T_i  = K  Q_i^T      (128 keys x 128 queries, contraction over d)        -> TMEM
//        dP^T_i = V  dO_i^T                                                          -> TMEM
//        dV    += Pd^T_i dO_i   (A = Pd^T from smem, K-major; B = dO_i MN-major)     -> TMEM, accumulates over i
//        dK    += dS^T_i Q_i                                                         -> TMEM, accumulates over i
//        dQ_i   = dS_i K        (A = the SAME dS^T smem tile read MN-major)          -> TMEM, fresh per i
//   warps 0-15: thread = (key row, 32-query quarter).  S^T / dP^T out of TMEM in one batch, then
//        P^T = 2^(S^T*scale*log2e - lse_q), Pd^T = P^T o dropmask, dS^T = P^T o (dP^T o dropmask - delta_q)  [the
//        softmax scale is applied once to dK at the final store and to dQ in dq_convert, not per element],
//        both written as fp16 A operands (128B-swizzled);  dQ_i is drained TMEM -> registers -> red.global.add.v4.f32
//        into a fp32 [B, T, H*d] workspace (each key tile contributes its partial dQ), converted to fp16 afterwards.
// S^T_{i+1} / dP^T_{i+1} are issued as soon as the compute warps hold tile i in registers, so the tensor pipe
// works ahead of the exponentials.  Ragged edges are trimmed: the last query tile issues N = ceil16(valid
// queries) and contracts over that many queries only; warps whose queries or keys are all out of range skip
// the exponentials (they write the zeros the MMAs need).  Keys >= valid[b] give P = 0; key tiles beyond valid[b] write zeros and exit;
// out-of-range query rows are zero-filled by TMA and carry lse = +inf.  Dropout masks are regenerated from the
// forward's (seed, index) hash.  head_dim 40: the 8 pad columns of K and V are zeroed in smem once per CTA.
#include "fhb_common.cuh"

namespace {

constexpr int kT = 128;
constexpr uint32_t kTileBytes = kT * 64 * 2;  // 16 KiB
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kCT = 512;  // compute threads (16 warps)
// Phase trace (tools/attn_bwd_trace.py): clock64 stamps of lane 0 of every compute warp ([warp * 64 + tile * 16 + slot])
// and of the MMA thread ([1024 + tile * 16 + slot]) of CTA (0, 0, 0), compiled in only under -DFHB_BWD_TRACE (a separate
// build of the library; exports fhb_attn_bwd_trace_read)
#ifdef FHB_BWD_TRACE
__device__ long long g_bwd_trace[2048];
#define BWD_TR(slot) do { if (trace_on) g_bwd_trace[trace_base + (slot)] = clock64(); } while (0)
#else
#define BWD_TR(slot) do { } while (0)
#endif
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int HD>
struct Smem {
  static constexpr uint32_t kK = 0, kV = kTileBytes, kQ = 2 * kTileBytes, kDO = 4 * kTileBytes, kP = 6 * kTileBytes,
                            kDS = 8 * kTileBytes;
  static constexpr uint32_t kStats = 10 * kTileBytes;   // lse[2][128], delta[2][128] floats
  static constexpr uint32_t kBars = kStats + 4 * kT * 4;
  static constexpr uint32_t kTotal = kBars + 128;
};

template <int HD, bool DROP>
__global__ void __launch_bounds__(kCT + 32, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   float* __restrict__ dq_ws, const int* __restrict__ valid,
                   const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                   int T, int H, float scale, uint32_t drop_seed, uint32_t drop_thr, float drop_scale) {
  constexpr int DK = (HD + 15) / 16 * 16;
  using S = Smem<HD>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* kv_full = bars;         // TMA: K, V landed
  uint64_t* kv_ready = bars + 1;    // kCT: pad columns zeroed (HD % 16 != 0)
  uint64_t* qd_full = bars + 2;     // [2] TMA: Q_i, dO_i landed
  uint64_t* s_full = bars + 4;      // S^T_i, dP^T_i in TMEM
  uint64_t* s_free = bars + 5;      // 256: both copied to registers
  uint64_t* p_full = bars + 6;      // 256: Pd^T_i, dS^T_i in smem
  uint64_t* mma_done = bars + 7;    // dV, dK, dQ_i MMAs of tile i retired (smem P/dS + Q/dO stage reusable, dQ_i ready)
  uint64_t* dq_free = bars + 8;     // 256: dQ_i copied out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* lse_s = reinterpret_cast<float*>(smem + S::kStats);
  float* delta_s = lse_s + 2 * kT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
#ifdef FHB_BWD_TRACE
  const bool trace_on = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0;
  const int trace_base = warp * 64;
  if (threadIdx.x == 0 && trace_on) g_bwd_trace[2047] = clock64();
#endif
  const int HDall = H * HD;
  const long long ld = 3LL * HDall;
  pdl_sync();  // before the first global read (valid[]) and before the early exit below
  int nvalid = valid ? valid[b] : T;
  nvalid = max(1, min(nvalid, T));
  const int nq = (T + kT - 1) / kT;

  if (k0 >= nvalid) {
    // every key of this tile is masked: P = 0 -> dK = dV = 0 (block-uniform early exit, no TMEM needed)
    for (int i = threadIdx.x; i < kT * (HD / 8); i += blockDim.x) {
      const int r = i / (HD / 8), c = i - r * (HD / 8);
      if (k0 + r < T) {
        __nv_bfloat16* base = dqkv + ((long long)b * T + k0 + r) * ld + h * HD + c * 8;
        *reinterpret_cast<uint4*>(base + HDall) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(base + 2 * HDall) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    return;
  }

  if (warp == 16 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    mbar_init(kv_ready, kCT);
    mbar_init(&qd_full[0], 1);
    mbar_init(&qd_full[1], 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, kCT);
    mbar_init(p_full, kCT);
    mbar_init(mma_done, 1);
    mbar_init(dq_free, kCT);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320,
                 tm_dqa = tmem_base + 384;

  if (warp == 16) {
    if (elect_one()) {
      auto load_qd = [&](int i) {
        const int st = i & 1;
        mbar_expect_tx(&qd_full[st], 2 * kTileBytes);
        tma_load_3d(&tm_qkv, &qd_full[st], smem + S::kQ + st * kTileBytes, h * HD, i * kT, b);
        tma_load_3d(&tm_do, &qd_full[st], smem + S::kDO + st * kTileBytes, h * HD, i * kT, b);
      };
      mbar_expect_tx(kv_full, 2 * kTileBytes);
      tma_load_3d(&tm_qkv, kv_full, smem + S::kK, HDall + h * HD, k0, b);
      tma_load_3d(&tm_qkv, kv_full, smem + S::kV, 2 * HDall + h * HD, k0, b);
      load_qd(0);
      if (nq > 1) load_qd(1);
      // every operand is fp16 (tcgen05 kind::f16 needs A and B in ONE format: mixing fp16 with bf16 is an illegal
      // instruction on sm_100a)
      const uint32_t id_dv = umma_idesc_16(128, DK, 0, 1, 0, 0);   // dV / dK: A K-major (queries), B MN-major
      const uint32_t id_dk = id_dv;
      const uint32_t id_dq = umma_idesc_16(128, DK, 1, 1, 0, 0);   // dQ: A = dS^T read MN-major, B = K MN-major
      const uint32_t ka = smem_u32(smem + S::kK), va = smem_u32(smem + S::kV);
      const uint32_t pa = smem_u32(smem + S::kP), dsa = smem_u32(smem + S::kDS);
      auto nq16 = [&](int i) { return (min(kT, T - i * kT) + 15) & ~15; };  // valid queries of tile i, rounded to 16
      auto issue_s = [&](int i) {
        const uint32_t qa = smem_u32(smem + S::kQ + (i & 1) * kTileBytes);
        const uint32_t da = smem_u32(smem + S::kDO + (i & 1) * kTileBytes);
        const uint32_t id_s = umma_idesc_16(128, (uint32_t)nq16(i), 0, 0, 0, 0);   // both operands K-major over d
        const uint32_t id_p = umma_idesc_16(128, (uint32_t)nq16(i), 0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
          tc_mma_bf16(tm_s, umma_desc_sw128(ka + k * 32, 0, 1024), umma_desc_sw128(qa + k * 32, 0, 1024), id_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
          tc_mma_bf16(tm_dp, umma_desc_sw128(va + k * 32, 0, 1024), umma_desc_sw128(da + k * 32, 0, 1024), id_p, k > 0 ? 1u : 0u);
        tc_commit(s_full);
      };
      if (HD % 16) mbar_wait(kv_ready, 0); else mbar_wait(kv_full, 0);
      mbar_wait(&qd_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        BWD_TR(16 * i + 0);
        if (i + 1 < nq) {
          mbar_wait(&qd_full[st ^ 1], ((i + 1) >> 1) & 1);
          BWD_TR(16 * i + 1);
          mbar_wait(s_free, i & 1);
          BWD_TR(16 * i + 2);
          tc_fence_after();
          issue_s(i + 1);
          BWD_TR(16 * i + 3);
        }
        mbar_wait(p_full, i & 1);
        BWD_TR(16 * i + 4);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + S::kQ + st * kTileBytes);
        const uint32_t da = smem_u32(smem + S::kDO + st * kTileBytes);
        const int ksteps = nq16(i) / 16;  // contraction over the valid queries of this tile only
        // descriptors: base + constant (the 14-bit address field cannot overflow: shared memory ends below 256 KB), so
        // the unrolled, predicated issue loops are a handful of instructions per MMA
        const uint64_t d_p = umma_desc_sw128(pa, 0, 1024), d_ds = umma_desc_sw128(dsa, 0, 1024);
        const uint64_t d_q = umma_desc_sw128(qa, 0, 1024), d_do = umma_desc_sw128(da, 0, 1024);
#pragma unroll
        for (int k = 0; k < kT / 16; ++k) {  // A atoms of 64 queries, B 16 query rows = 2 KiB per step
          const uint32_t aoff = (k >> 2) * kTileBytes + (k & 3) * 32;
          if (k < ksteps) tc_mma_bf16(tm_dv, d_p + (aoff >> 4), d_do + ((k * 2048) >> 4), id_dv, (i > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < kT / 16; ++k) {
          const uint32_t aoff = (k >> 2) * kTileBytes + (k & 3) * 32;
          if (k < ksteps) tc_mma_bf16(tm_dk, d_ds + (aoff >> 4), d_q + ((k * 2048) >> 4), id_dk, (i > 0 || k > 0) ? 1u : 0u);
        }
        BWD_TR(16 * i + 5);
        if (i > 0) {
          mbar_wait(dq_free, (i - 1) & 1);
          tc_fence_after();
        }
        BWD_TR(16 * i + 6);
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)  // contraction over the 128 keys: dS^T rows are the K rows (MN-major A)
          tc_mma_bf16(tm_dqa, umma_desc_sw128(dsa + k * 2048, kTileBytes, 1024), umma_desc_sw128(ka + k * 2048, 0, 1024),
                      id_dq, k > 0 ? 1u : 0u);
        tc_commit(mma_done);
        BWD_TR(16 * i + 7);
        if (i + 2 < nq) {
          mbar_wait(mma_done, i & 1);
          BWD_TR(16 * i + 8);
          load_qd(i + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ compute warps
    const int quarter = warp & 3, quad = warp >> 2;  // TMEM lane quarter (key rows), 32-query column group
    const int row = quarter * 32 + lane;              // key row of this thread within the tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t rsw = (uint32_t)(row & 7);
    const int key = k0 + row;
    const float keymask = key < nvalid ? 1.f : 0.f;
    const bool warp_keys_live = (k0 + quarter * 32) < nvalid;  // warp-uniform: any valid key in this warp's rows
    const bool masked_keys = (k0 + quarter * 32 + 32) > nvalid;  // warp-uniform: some key row of this warp is masked
    const int tid = threadIdx.x;
    if (HD % 16) {
      // pad columns HD..DK-1 of K (threads 0-127) and V (128-255) hold the next head's values: zero them
      mbar_wait(kv_full, 0);
      if (tid < 256) {
        const uint32_t tile = smem_u32(smem + (tid >= 128 ? S::kV : S::kK));
        st_shared_v4(tile + row * 128 + ((((uint32_t)HD >> 3) ^ rsw) << 4), 0u, 0u, 0u, 0u);
        fence_async_shared();
      }
      mbar_arrive(kv_ready);
    }
    const float sc = scale * kLog2e;
    const uint32_t col_off = (uint32_t)(quad >> 1) * kTileBytes + row * 128;  // atom of 64 queries + this row
    const uint32_t prow = smem_u32(smem + S::kP) + col_off;
    const uint32_t dsrow = smem_u32(smem + S::kDS) + col_off;
    const uint32_t ch0 = (uint32_t)(quad & 1) * 4;  // first 16-byte chunk of this thread's 32 queries inside the atom
    // dQ drain / final dK, dV store geometry: thread owns tile row `row`, column group [16 quad, 16 quad + 16)
    const int c16 = quad * 16;
    auto drain_dq = [&](int j) {
      // dQ_j (this key tile's partial): TMEM -> registers -> vector reductions red.global.add.v4.f32 straight into the
      // fp32 [B, T, H*d] workspace (thread = query row, 16 columns).  No shared-memory staging, no block barrier, no TMA:
      // the staged TMA reduce-add this replaces cost every warp ~850 clk per tile (two block barriers around it,
      // profiles/r05_attnbwd_trace_after.txt)
      tc_fence_after();
      if (c16 < DK) {
        uint32_t r[16];
        tmem_ld16(tm_dqa + lane_off + c16, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_free);
        BWD_TR(16 * ((j + 1) & 3) + 11);
        if constexpr (HD > 48) {
          // 4 x 4 transpose of the 16-byte chunks inside every group of four lanes (two butterfly steps of shuffles): lane
          // l then holds chunk l & 3 of the four rows 4 (l >> 2) .. + 3, so one reduction instruction of the warp covers
          // 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes.  d = 64, T = 779: 411 -> 389 us; at d = 40 (160-byte
          // rows, 10 chunks) it measured 2 us SLOWER than the plain form below (profiles/r05_attnbwd_red_ab.txt)
#pragma unroll
          for (int bit = 0; bit < 2; ++bit) {
            const bool up = (lane >> bit) & 1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if ((c >> bit) & 1) continue;
              const int lo = c, hi = c | (1 << bit);
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const uint32_t send = up ? r[4 * lo + w] : r[4 * hi + w];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1 << bit);
                if (up) r[4 * lo + w] = recv;
                else r[4 * hi + w] = recv;
              }
            }
          }
          const int cq = c16 + 4 * (lane & 3);                  // first column of this lane's chunk
          const int qb = j * kT + quarter * 32 + (lane & ~3);   // first of its four query rows
          float* dst = dq_ws + ((long long)b * T + qb) * HDall + h * HD + cq;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (qb + k < T && cq < HD)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (long long)k * HDall),
                           "f"(__uint_as_float(r[4 * k])), "f"(__uint_as_float(r[4 * k + 1])), "f"(__uint_as_float(r[4 * k + 2])),
                           "f"(__uint_as_float(r[4 * k + 3])) : "memory");
          }
        } else {
          const int q = j * kT + row;
          if (q < T) {
            float* dst = dq_ws + ((long long)b * T + q) * HDall + h * HD + c16;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              if (c16 + 4 * q4 < HD)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q4), "f"(__uint_as_float(r[4 * q4])),
                             "f"(__uint_as_float(r[4 * q4 + 1])), "f"(__uint_as_float(r[4 * q4 + 2])),
                             "f"(__uint_as_float(r[4 * q4 + 3])) : "memory");
            }
          }
        }
      } else {
        tc_fence_before();
        mbar_arrive(dq_free);
      }
      BWD_TR(16 * ((j + 1) & 3) + 12);
    };
    // lse / delta of a query tile: fetched as RAW values from a clamped address well ahead of their use (nothing
    // depends on the load before the block barrier that follows it, so its latency hides under a tile's math); negated /
    // scaled / masked when they are stored to shared memory at the END of the previous tile, in front of the block
    // barriers of the dQ drain - the tile loop has no barrier of its own.  The inner loop reads them straight into
    // packed FFMA2 addends.
    const float* stat_src = (tid < 128 ? lse : delta) + ((long long)b * H + h) * T;
    auto load_stat = [&](int i) -> float {
      return tid < 256 ? stat_src[min(i * kT + (tid & 127), T - 1)] : 0.f;
    };
    auto store_stat = [&](int i, float raw) {  // buffer i & 1: last read by the math of tile i - 2
      if (tid < 128) lse_s[(i & 1) * kT + tid] = i * kT + tid < T ? -raw * kLog2e : -INFINITY;
      else if (tid < 256) delta_s[(i & 1) * kT + tid - 128] = i * kT + tid - 128 < T ? -raw : 0.f;
    };
    float stat = load_stat(0);
    store_stat(0, stat);
    if (nq > 1) stat = load_stat(1);
    bar_compute();
    for (int i = 0; i < nq; ++i) {
      const int q0 = i * kT;
      float* ls = lse_s + (i & 1) * kT;
      float* dl = delta_s + (i & 1) * kT;
      BWD_TR(16 * i + 0);
      const int nq16v = (min(kT, T - q0) + 15) & ~15;      // query columns the MMAs of this tile touch
      const bool cols_live = quad * 32 < nq16v;             // warp-uniform: this warp's queries exist
      const bool work = cols_live && warp_keys_live;
      BWD_TR(16 * i + 1);
      mbar_wait(s_full, i & 1);
      BWD_TR(16 * i + 2);
      tc_fence_after();
      // Two passes of 16 query columns: 32 live accumulator registers instead of 64 (the 17-warp block caps a thread at
      // 96 registers), so the hash constants stay in registers and nothing spills.  S^T / dP^T are released to the MMA
      // warp after the second pass's loads.
      const f32x2_t sc2 = pack2(sc, sc);
      // Dropout bits: one 32-bit hash serves the key pair (2m, 2m + 1) of a query (fhb_common.cuh), and lanes 2m / 2m + 1
      // of this warp own exactly those two key rows - so each lane hashes every second query of its 32 and the
      // partner lane supplies the other half through one shuffle (the forward's pair index q * ceil(T / 2) + key / 2,
      // advanced by a constant per query pair).  Even key rows read the low 16 bits, odd ones the high 16.
      const bool odd = lane & 1;
      const uint32_t psel = odd ? 0x3276u : 0x5410u, thrhi = drop_thr << 16;
      const uint32_t T2h = (uint32_t)((T + 1) >> 1);
      const uint32_t xh = ((uint32_t)((b * H + h) * T + q0 + quad * 32 + (int)odd) * T2h + ((uint32_t)key >> 1)) * 0x9E3779B1u + drop_seed;
      const uint32_t xstep = 2u * T2h * 0x9E3779B1u;
      const float4* ls4 = reinterpret_cast<const float4*>(ls + quad * 32);
      const float4* dl4 = reinterpret_cast<const float4*>(dl + quad * 32);
      uint32_t held[8];
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        uint32_t sr[16], dr[16];
        if (work) {
          tmem_ld16(tm_s + lane_off + quad * 32 + hp * 16, sr);
          tmem_ld16(tm_dp + lane_off + quad * 32 + hp * 16, dr);
          tmem_ld_wait();
        }
        BWD_TR(16 * i + 3 + 3 * hp);
        if (hp == 1) {
          tc_fence_before();
          mbar_arrive(s_free);
        }
        if (work) {
          // P^T = 2^(S^T sc - lse), Pd^T = P^T o mask, dS^T = P^T o (dP^T o mask - delta)   (x scale folded into the dQ
          // conversion and the dK store).  Two queries per FFMA2 / FMUL2; lse and delta come as 16-byte smem vectors.
#pragma unroll
          for (int c8 = 0; c8 < 16; c8 += 8) {
            const int c = hp * 16 + c8;  // first query column (of this thread's 32) of the group of eight
            const float4 la = ls4[c >> 2], lb = ls4[(c >> 2) + 1], da = dl4[c >> 2], db = dl4[(c >> 2) + 1];
            const float nl[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
            const float nd[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
            float pd[8], ds[8];
            if (hp == 0 && c8 == 8) {
              // the previous tile's dV / dK / dQ MMAs still read the P / dS tiles while the first eight columns are being
              // computed: wait for them only here, then store the eight columns held back in registers
              if (i > 0) mbar_wait(mma_done, (i - 1) & 1);
              BWD_TR(16 * i + 4);
              const uint32_t chp = ch0 + (uint32_t)((hp * 16) >> 3);
              st_shared_v4(prow + ((chp ^ rsw) << 4), held[0], held[1], held[2], held[3]);
              st_shared_v4(dsrow + ((chp ^ rsw) << 4), held[4], held[5], held[6], held[7]);
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              float t0, t1;
              unpack2(fma2(pack2(__uint_as_float(sr[c8 + e]), __uint_as_float(sr[c8 + e + 1])), sc2, pack2(nl[e], nl[e + 1])), t0, t1);
              f32x2_t p2 = pack2(ex2_approx(t0), ex2_approx(t1));
              if (masked_keys) p2 = mul2(p2, pack2(keymask, keymask));
              f32x2_t dp2 = pack2(__uint_as_float(dr[c8 + e]), __uint_as_float(dr[c8 + e + 1]));
              f32x2_t pd2 = p2;
              if (DROP) {
                const uint32_t mine = fhb_hash32(xh + (uint32_t)((c + e) >> 1) * xstep);  // query c + e + odd
                const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);                // query c + e + 1 - odd
                // one byte permute gathers the two 16-bit decisions of this lane: (query c + e | query c + e + 1) =
                // even key rows (low halves): (mine | other), odd key rows (high halves): (other | mine)
                const uint32_t w = __byte_perm(mine, other, psel);
                const f32x2_t mk2 = pack2((w << 16) >= thrhi ? drop_scale : 0.f, w >= thrhi ? drop_scale : 0.f);
                pd2 = mul2(p2, mk2);
                dp2 = mul2(dp2, mk2);
              }
              unpack2(pd2, pd[e], pd[e + 1]);
              unpack2(mul2(p2, add2(dp2, pack2(nd[e], nd[e + 1]))), ds[e], ds[e + 1]);
            }
            const uint32_t ch = ch0 + (uint32_t)(c >> 3);
            if (hp == 0 && c8 == 0) {
              held[0] = pack_f16(pd[0], pd[1]); held[1] = pack_f16(pd[2], pd[3]);
              held[2] = pack_f16(pd[4], pd[5]); held[3] = pack_f16(pd[6], pd[7]);
              held[4] = pack_f16(ds[0], ds[1]); held[5] = pack_f16(ds[2], ds[3]);
              held[6] = pack_f16(ds[4], ds[5]); held[7] = pack_f16(ds[6], ds[7]);
            } else {
              st_shared_v4(prow + ((ch ^ rsw) << 4), pack_f16(pd[0], pd[1]), pack_f16(pd[2], pd[3]), pack_f16(pd[4], pd[5]),
                           pack_f16(pd[6], pd[7]));
              st_shared_v4(dsrow + ((ch ^ rsw) << 4), pack_f16(ds[0], ds[1]), pack_f16(ds[2], ds[3]), pack_f16(ds[4], ds[5]),
                           pack_f16(ds[6], ds[7]));
            }
          }
        } else if (hp == 0) {
          if (i > 0) mbar_wait(mma_done, (i - 1) & 1);  // every thread: the drain below reads dQ_{i-1} out of TMEM
          if (cols_live) {
            // every key of this warp is masked: P = dS = 0 (the dQ contraction reads these rows)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              st_shared_v4(prow + (((ch0 + c) ^ rsw) << 4), 0u, 0u, 0u, 0u);
              st_shared_v4(dsrow + (((ch0 + c) ^ rsw) << 4), 0u, 0u, 0u, 0u);
            }
          }
        }
        BWD_TR(16 * i + 5 + 3 * hp);
      }
      tc_fence_before();
      fence_async_shared();
      mbar_arrive(p_full);
      BWD_TR(16 * i + 9);
      if (i + 1 < nq) {
        store_stat(i + 1, stat);
        if (i + 2 < nq) stat = load_stat(i + 2);
      }
      if (i > 0) drain_dq(i - 1);
      bar_compute();  // the one block barrier of a tile: publishes the statistics stored above
      BWD_TR(16 * i + 10);
    }
    mbar_wait(mma_done, (nq - 1) & 1);
    drain_dq(nq - 1);
    // ---- dK, dV: TMEM -> fp16 -> global.  Thread = key row, columns [16 quad, 16 quad + 16)
    tc_fence_after();
    if (c16 < DK) {
      __nv_bfloat16* krow = dqkv + ((long long)b * T + key) * ld + HDall + h * HD;
      __nv_bfloat16* vrow = krow + HDall;
      uint32_t rk[16], rv[16];
      tmem_ld16(tm_dk + lane_off + c16, rk);  // warp-collective: every lane takes part, stores are predicated
      tmem_ld16(tm_dv + lane_off + c16, rv);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) rk[e] = __float_as_uint(__uint_as_float(rk[e]) * scale);  // dS^T was kept un-scaled
      if (key >= nvalid) {  // masked keys: exact zeros (their P rows were never computed)
#pragma unroll
        for (int e = 0; e < 16; ++e) rk[e] = rv[e] = 0u;
      }
#pragma unroll
      for (int q8 = 0; q8 < 2; ++q8) {
        if (key < T && c16 + 8 * q8 < HD) {
          const uint32_t* a = rk + 8 * q8;
          const uint32_t* v = rv + 8 * q8;
          *reinterpret_cast<uint4*>(krow + c16 + 8 * q8) =
              make_uint4(pack_f16(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_f16(__uint_as_float(a[2]), __uint_as_float(a[3])),
                         pack_f16(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_f16(__uint_as_float(a[6]), __uint_as_float(a[7])));
          *reinterpret_cast<uint4*>(vrow + c16 + 8 * q8) =
              make_uint4(pack_f16(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_f16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                         pack_f16(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_f16(__uint_as_float(v[6]), __uint_as_float(v[7])));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dqkv[row][0:E] = fp16(scale * dq_acc[row][0:E])   (the softmax scale the fused kernel left out of dS)
__global__ void __launch_bounds__(256)
dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, long long rows, int E8, long long ld,
                  float scale) {
  pdl_sync();
  const long long total = rows * E8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / E8;
    const int c = (int)(i - r * E8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(acc + r * (E8 * 8) + c);
    const float4 b2 = *reinterpret_cast<const float4*>(acc + r * (E8 * 8) + c + 4);
    *reinterpret_cast<uint4*>(dqkv + r * ld + c) =
        make_uint4(pack_f16(a.x * scale, a.y * scale), pack_f16(a.z * scale, a.w * scale),
                   pack_f16(b2.x * scale, b2.y * scale), pack_f16(b2.z * scale, b2.w * scale));
  }
}

template <int HD>
int launch_bwd(const void* qkv, const int32_t* valid, const void* dout, const float* lse, const float* delta, void* dqkv,
               float* dq_ws, int32_t B, int32_t T, int32_t H, float scale, uint32_t drop_seed, float drop_p,
               cudaStream_t s) {
  using S = Smem<HD>;
  const int64_t E = (int64_t)H * HD;
  CUtensorMap tq, td;
  {
    const int64_t dim[3] = {3 * E, T, B}, stride[2] = {3 * E, 3 * E * T};
    int rc = fhb_make_tmap_bf16_3d(&tq, qkv, dim, stride, 64, kT, "qkv");
    if (rc) return rc;
  }
  {
    const int64_t dim[3] = {E, T, B}, stride[2] = {E, E * T};
    int rc = fhb_make_tmap_bf16_3d(&td, dout, dim, stride, 64, kT, "dout");
    if (rc) return rc;
  }
  FHB_CUDA_CHECK(cudaMemsetAsync(dq_ws, 0, sizeof(float) * (size_t)B * T * E, s));
  FHB_ONCE_PER_DEVICE({
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  });
  dim3 grid((T + kT - 1) / kT, H, B);
  if (drop_p > 0.f)
    FHB_CUDA_CHECK(fhb_launch((attn_bwd_tc_kernel<HD, true>), dim3(grid), dim3(kCT + 32), S::kTotal, s, tq, td, dq_ws, valid, lse, delta, static_cast<__nv_bfloat16*>(dqkv), T,
                                                             H, scale, drop_seed, fhb_dropout_thr16(drop_p),
                                                             fhb_dropout_scale(drop_p)));
  else
    FHB_CUDA_CHECK(fhb_launch((attn_bwd_tc_kernel<HD, false>), dim3(grid), dim3(kCT + 32), S::kTotal, s, tq, td, dq_ws, valid, lse, delta, static_cast<__nv_bfloat16*>(dqkv),
                                                              T, H, scale, 0u, 0u, 1.f));
  FHB_LAUNCH_CHECK();
  const long long rows = (long long)B * T;
  long long blocks = (rows * (E / 8) + 255) / 256;
  if (blocks > 8LL * fhb_num_sms()) blocks = 8LL * fhb_num_sms();
  fhb_pdl_hint(true);
  FHB_CUDA_CHECK(fhb_launch(dq_convert_kernel, dim3((unsigned)blocks), dim3(256), 0, s, dq_ws, static_cast<__nv_bfloat16*>(dqkv), rows, (int)(E / 8), 3 * E, scale));
  FHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

#ifdef FHB_BWD_TRACE
extern "C" int fhb_attn_bwd_trace_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_bwd_trace, sizeof(long long) * (size_t)(n < 2048 ? n : 2048));
}
#endif

// Called by fhb_attn_bwd (attention.cu) for head_dim 40 / 64 when a dQ workspace is supplied.
int fhb_attn_bwd_tc(const void* qkv, const int32_t* valid, const void* dout, const float* lse, const float* delta,
                    void* dqkv, float* dq_ws, int32_t B, int32_t T, int32_t H, int32_t d, float scale, uint32_t drop_seed,
                    float drop_p, cudaStream_t s) {
  if (d == 64) return launch_bwd<64>(qkv, valid, dout, lse, delta, dqkv, dq_ws, B, T, H, scale, drop_seed, drop_p, s);
  return launch_bwd<40>(qkv, valid, dout, lse, delta, dqkv, dq_ws, B, T, H, scale, drop_seed, drop_p, s);
}
