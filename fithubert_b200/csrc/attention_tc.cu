// K7 forward on 5th-gen tensor cores, head_dim 64 (HuBERT-Base / wav2vec2-Base teacher) and head_dim 40
// (FitHuBERT student).  Replaces the bmm -> masked_fill(-inf) -> fp32 softmax -> bmm chain of fairseq
// MultiheadAttention (reached from modules/module.py:558-564; the teacher runs the same code in fairseq).
//
// Persistent CTAs (2 per SM, 288 threads, 113 KB smem, 256 TMEM columns each) walk a static list of work items
// (128-query tile, head, sample) - the full query tiles first, the ragged last tile of every (sample, head) at the end, so
// the cheap items level the tail - and run ONE software pipeline across item boundaries: the Q / K / V loads and the
// S = Q K^T product of the next item are in flight while the softmax warps finish the current one.
//   warp 8 (one elected lane): TMA loads (4-D map {d, 3H slots, T, B}: a 40-wide head lands in a 64-wide swizzled tile
//                              with the pad columns zero-filled by TMA; K/V double buffered) and all tcgen05.mma issue:
//                              S_j = Q K_j^T (128 x 128 x d), O_half += P_j,half V_j,half (128 x d x 64) per key half
//   warps 0-7: thread = (query row, 64-key half).  The two halves of a row are INDEPENDENT online softmaxes with their
//              own running maximum, row sum and O accumulator in TMEM (O_lo over keys [0,64) of every tile, O_hi over
//              [64,128)); they only meet once per item, in the epilogue:  O = (a_lo O_lo + a_hi O_hi) / (a_lo l_lo +
//              a_hi l_hi),  a = 2^(m_half - max(m_lo, m_hi)).  No per-tile exchange, no block barrier inside the key
//              loop: each warp runs tcgen05.ld -> max (FMNMX3) -> exp2 (FFMA2 / MUFU / FADD2) -> fp16 P -> swizzled
//              smem at its own pace, so the MUFU phase of one warp overlaps the TMEM / ALU phases of the others.
//   The running maximum of a half is only moved - and its O rescaled in TMEM - when it grows by more than 2^8 (the
//   probabilities then stay below 2^8: in fp16 range); l and the LSE stay exact because P, l and O share the reference.
//   Ragged edges are trimmed: the last key tile issues N = ceil16(valid keys) and contracts over that many keys
//   only; warps whose 32 query rows are all >= T skip the exponentials.
// Keys >= valid[b] are masked (P = 0) and key tiles entirely beyond valid[b] are skipped; padded QUERY
// rows are computed like any other.
#include "fhb_common.cuh"

namespace {

constexpr int kTQ = 128, kTK = 128;
constexpr uint32_t kTileBytes = kTQ * 64 * 2;  // 16 KiB: 128 rows x 64 fp16 (one 128-byte swizzle row each)
constexpr uint32_t kSmem = kTileBytes /*Q*/ + 2 * kTileBytes /*K*/ + 2 * kTileBytes /*V*/ + 2 * kTileBytes /*P halves*/ + 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One work item = (128-query tile, head, sample).  Item order: every full query tile first ((sample, head) slow, tile
// fast: neighbouring CTAs share K / V in L2), then the ragged last tile of every (sample, head).
struct Item {
  int q0, h, b, nvalid, nt;
};
struct ItemSpace {
  int total, nq, nfull, BH, H, T;
  const int* valid;
};
__device__ __forceinline__ bool decode_item(const ItemSpace& sp, int w, Item& it) {
  if (w >= sp.total) return false;
  int qt, bh;
  if (w < sp.nfull * sp.BH) {
    bh = w / sp.nfull;
    qt = w - bh * sp.nfull;
  } else {
    bh = w - sp.nfull * sp.BH;
    qt = sp.nq - 1;
  }
  it.q0 = qt * kTQ;
  it.b = bh / sp.H;
  it.h = bh - it.b * sp.H;
  int nv = sp.valid ? sp.valid[it.b] : sp.T;
  nv = max(1, min(nv, sp.T));
  it.nvalid = nv;
  it.nt = (nv + kTK - 1) / kTK;
  return true;
}
// position in the flat (item, key tile) sequence of one CTA
struct Cursor {
  Item it;
  int w, j;
  bool ok;
  __device__ __forceinline__ void init(const ItemSpace& sp, int w0) {
    w = w0;
    j = 0;
    ok = decode_item(sp, w, it);
  }
  __device__ __forceinline__ void next(const ItemSpace& sp, int stride) {
    if (++j == it.nt) {
      j = 0;
      w += stride;
      ok = decode_item(sp, w, it);
    }
  }
  __device__ __forceinline__ int nk16() const { return (min(kTK, it.nvalid - j * kTK) + 15) & ~15; }  // keys the MMAs touch
};

// HD: logical head dim (64 or 40); DK = HD rounded up to the UMMA k-step / n-step of 16
template <int HD, bool DROP>
__global__ void __launch_bounds__(288, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const int* __restrict__ valid,
                   __half* __restrict__ out, float* __restrict__ lse, int T, int H, int B, float scale,
                   uint32_t drop_seed, uint32_t drop_thr, float drop_scale) {
  constexpr int DK = (HD + 15) / 16 * 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 3 * kTileBytes;
  uint8_t* sP = smem + 5 * kTileBytes;  // [half][128 rows][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;     // 256 arrivals: S_j has been copied to registers
  uint64_t* p_full = bars + 7;     // [2] 128 arrivals each: this half of P_j is in smem (and its O has been rescaled if needed)
  uint64_t* o_done = bars + 9;     // [2] PV_j of this half retired: its P smem is reusable, its O readable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_done[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // 128 fp32 columns
  const uint32_t tmem_o = tmem_base + 128;  // 2 x DK fp32 columns: O_lo, O_hi
  pdl_sync();  // the prologue above overlapped the previous kernel's tail; global memory only from here on
  ItemSpace sp;
  sp.nq = (T + kTQ - 1) / kTQ;
  sp.nfull = (sp.nq > 1 && (T % kTQ) != 0) ? sp.nq - 1 : sp.nq;
  sp.BH = B * H;
  sp.H = H;
  sp.T = T;
  sp.total = sp.nq * sp.BH;
  sp.valid = valid;
  const int stride = gridDim.x;

  if (warp == 8) {
    if (elect_one()) {
      Cursor L, S, P;  // K/V load, S issue and P V issue positions (L runs 2 tiles ahead of P, S one)
      L.init(sp, blockIdx.x);
      S = L;
      P = L;
      auto load_kv = [&](const Cursor& c, int g) {
        const int st = g & 1;
        mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
        tma_load_4d(&tm_qkv, &kv_full[st], smem_u32(sK + st * kTileBytes), 0, H + c.it.h, c.j * kTK, c.it.b);
        tma_load_4d(&tm_qkv, &kv_full[st], smem_u32(sV + st * kTileBytes), 0, 2 * H + c.it.h, c.j * kTK, c.it.b);
      };
      auto load_q = [&](const Item& it) {
        mbar_expect_tx(q_full, kTileBytes);
        tma_load_4d(&tm_qkv, q_full, smem_u32(sQ), 0, it.h, it.q0, it.b);
      };
      const uint32_t idesc_o = umma_idesc_16(128, DK, 0, 1, 0, 0);  // O: A = P (K-major), B = V (MN-major), both fp16
      const uint32_t qa = smem_u32(sQ), pa = smem_u32(sP);
      auto issue_s = [&](const Cursor& c, int g) {
        const uint32_t ka = smem_u32(sK + (g & 1) * kTileBytes);
        const uint32_t idesc_s = umma_idesc_16(128, (uint32_t)c.nk16(), 0, 0, 0, 0);  // S: A = Q, B = K, both K-major, fp16
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
          tc_mma_bf16(tmem_s, umma_desc_sw128(qa + k * 32, 0, 1024), umma_desc_sw128(ka + k * 32, 0, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
        tc_commit(s_full);
      };
      if (P.ok) {
        load_q(S.it);
        load_kv(L, 0);
        L.next(sp, stride);
        if (L.ok) {
          load_kv(L, 1);
          L.next(sp, stride);
        }
        int qc = 0;  // Q loads consumed so far (parity of q_full)
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_s(S, 0);
        bool q_stale = (S.j == S.it.nt - 1);  // the S just issued was the last product against this Q
        S.next(sp, stride);
        for (int g = 0; P.ok; ++g) {
          if (q_stale) {  // S_g has to retire before the next item's Q may overwrite the tile
            q_stale = false;
            if (S.ok) {
              mbar_wait(s_full, g & 1);
              ++qc;
              load_q(S.it);
            }
          }
          if (S.ok) {  // S_{g+1} as soon as the softmax warps have S_g in registers: overlaps their exponentials
            if (S.j == 0) mbar_wait(q_full, qc & 1);
            mbar_wait(&kv_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
            mbar_wait(s_free, g & 1);
            tc_fence_after();
            issue_s(S, g + 1);
            q_stale = (S.j == S.it.nt - 1);
            S.next(sp, stride);
          }
          const uint32_t va = smem_u32(sV + (g & 1) * kTileBytes);
          const int nk = P.nk16();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&p_full[hf], g & 1);
            tc_fence_after();
            const int ksteps = max(0, min(64, nk - hf * 64)) >> 4;
            for (int k = 0; k < ksteps; ++k)  // P half: one 64-key atom of 16 KiB; V: 16 key rows = 2 KiB per step
              tc_mma_bf16(tmem_o + hf * DK, umma_desc_sw128(pa + hf * kTileBytes + k * 32, 0, 1024),
                          umma_desc_sw128(va + (hf * 4 + k) * 2048, 0, 1024), idesc_o, (P.j > 0 || k > 0) ? 1u : 0u);
            tc_commit(&o_done[hf]);
          }
          tc_commit(&kv_empty[g & 1]);
          if (L.ok) {  // refill this K/V stage once S_g / PV_g have retired
            mbar_wait(&kv_empty[g & 1], (g >> 1) & 1);
            load_kv(L, g + 2);
            L.next(sp, stride);
          }
          P.next(sp, stride);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax: thread = (query row, 64-key half)
    const int quarter = warp & 3, half = warp >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t rsw = (uint32_t)(row_in_tile & 7);
    const float sc = scale * kLog2e;
    const uint32_t prow = smem_u32(sP) + half * kTileBytes + row_in_tile * 128;
    const uint32_t tmem_oh = tmem_o + half * DK + lane_off;  // this half's accumulator, this warp's lanes
    float2* xch = reinterpret_cast<float2*>(sP);              // (m, l) of both halves, exchanged once per item (P is idle then)
    // O columns stored by this thread: half 0 -> [0, 32), half 1 -> [32, DK)
    const int oc_begin = half ? 32 : 0, oc_end = half ? DK : 32;
    int n = 0;  // key tiles processed so far by this CTA (parity of the per-tile barriers)
    Item it;
    for (int w = blockIdx.x; decode_item(sp, w, it); w += stride) {
      const bool warp_live = it.q0 + quarter * 32 < T;  // warp-uniform: some query row of this warp exists
      float m_i = -INFINITY, l_i = 0.f;                 // this half's reference maximum (log2 units) and row sum
      bool have = false;                                // this half's accumulator has been written
      // dropout pair index of (b, h, q, k): ((b*H + h)*T + q) * ceil(T/2) + (k >> 1)
      const uint32_t drop_row = (uint32_t)(((it.b * H + it.h) * T + it.q0 + row_in_tile) * ((T + 1) >> 1));
      for (int j = 0; j < it.nt; ++j, ++n) {
        const int nk = it.nvalid - j * kTK;           // valid keys in this tile (>= 1; >= 128 for all but the last tile)
        const int nkh = nk - half * 64;               // ... of which fall after this half's first column
        const int n16 = ((min(kTK, nk) + 15) & ~15) - half * 64;  // columns of this half the MMAs touch (<= 0: none)
        const bool work = warp_live && n16 > 0;
        mbar_wait(s_full, n & 1);
        tc_fence_after();
        // Two passes over this half's 64 columns of S in TMEM, 32 at a time (64 fp32 scores + 16 probabilities live at
        // once would not fit the 96 registers two CTAs per SM leave): pass 1 finds the row maximum, pass 2 re-reads the
        // columns and exponentiates; S is released for the next product once the last columns are in registers.
        uint32_t r[32];
        bool need = false;
        float m_new = m_i;
        if (work) {
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c2 = 0; c2 < 64; c2 += 32) {
            if (c2 < n16) {
              tmem_ld32(tmem_s + lane_off + half * 64 + c2, r);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; c += 16) {
                if (c2 + c < n16) {
                  if (c2 + c + 16 <= nkh) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2)
                      mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(r[c + i]), __uint_as_float(r[c + i + 1]));
                  } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      mx4[i & 3] = fmaxf(mx4[i & 3], c2 + c + i < nkh ? __uint_as_float(r[c + i]) : -INFINITY);
                  }
                }
              }
            }
          }
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          m_new = fmaxf(m_i, mx * sc);
          if (!have) m_i = m_new;
          else need = (m_new - m_i) > kRescaleThreshold;
        }
        if (n > 0) mbar_wait(&o_done[half], (n - 1) & 1);  // this half's previous PV retired: its P smem is free, its O complete
        if (__any_sync(0xffffffffu, need)) {
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_i - m_new) : 1.f;
          if (need) m_i = m_new;
          l_i *= alpha;
          for (int c = 0; c < DK; c += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_oh + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_oh + c, o);
          }
          tmem_st_wait();
        }
        // probabilities -> fp16 (at most 2^8 after the lazy rescale: in range) -> swizzled smem (K-major A operand: this half's 64-key atom)
        bool released = false;
        if (work) {
          f32x2_t ls2 = pack2(0.f, 0.f), ls2b = pack2(0.f, 0.f);
          const f32x2_t sc2 = pack2(sc, sc), nm2 = pack2(-m_i, -m_i);
          const float neg_m = -m_i;
#pragma unroll
          for (int c2 = 0; c2 < 64; c2 += 32) {
            if (c2 < n16) {
              tmem_ld32(tmem_s + lane_off + half * 64 + c2, r);
              tmem_ld_wait();
              if (c2 + 32 >= n16) {  // the last columns this thread reads
                tc_fence_before();
                mbar_arrive(s_free);
                released = true;
              }
#pragma unroll
              for (int c = 0; c < 32; c += 16) {
                if (c2 + c < n16) {
                  float pv[16];
                  if (c2 + c + 16 <= nkh) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                      float x0, x1;
                      unpack2(fma2(pack2(__uint_as_float(r[c + i]), __uint_as_float(r[c + i + 1])), sc2, nm2), x0, x1);
                      pv[i] = ex2_approx(x0);
                      pv[i + 1] = ex2_approx(x1);
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      pv[i] = (c2 + c + i < nkh) ? ex2_approx(fmaf(__uint_as_float(r[c + i]), sc, neg_m)) : 0.f;
                  }
#pragma unroll
                  for (int i = 0; i < 16; i += 4) {
                    ls2 = add2(ls2, pack2(pv[i], pv[i + 1]));
                    ls2b = add2(ls2b, pack2(pv[i + 2], pv[i + 3]));
                  }
                  if (DROP) {  // attention dropout on the probabilities (the row sum l stays un-dropped)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      float m0, m1;
                      dropout_pair(drop_seed, drop_row + (uint32_t)((j * kTK + half * 64 + c2 + c) >> 1) + i, drop_thr,
                                   drop_scale, m0, m1);
                      pv[2 * i] *= m0;
                      pv[2 * i + 1] *= m1;
                    }
                  }
                  const uint32_t ch = (uint32_t)((c2 + c) >> 3);
                  st_shared_v4(prow + (((ch) ^ rsw) << 4), pack_f16(pv[0], pv[1]), pack_f16(pv[2], pv[3]),
                               pack_f16(pv[4], pv[5]), pack_f16(pv[6], pv[7]));
                  st_shared_v4(prow + (((ch + 1) ^ rsw) << 4), pack_f16(pv[8], pv[9]), pack_f16(pv[10], pv[11]),
                               pack_f16(pv[12], pv[13]), pack_f16(pv[14], pv[15]));
                }
              }
            }
          }
          float a0, a1, b0, b1;
          unpack2(ls2, a0, a1);
          unpack2(ls2b, b0, b1);
          l_i += (a0 + a1) + (b0 + b1);
          have = true;
        }
        if (!released) {
          tc_fence_before();
          mbar_arrive(s_free);
        }
        tc_fence_before();
        fence_async_shared();
        mbar_arrive(&p_full[half]);
      }
      // ---- epilogue of the item: both accumulators complete, the halves meet
      mbar_wait(&o_done[0], (n - 1) & 1);
      mbar_wait(&o_done[1], (n - 1) & 1);
      tc_fence_after();
      xch[half * kTQ + row_in_tile] = make_float2(m_i, l_i);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 other = xch[(half ^ 1) * kTQ + row_in_tile];
      const bool two = it.nvalid > 64;  // the high half has seen keys (uniform over the CTA)
      const float m_lo = half ? other.x : m_i, l_lo = half ? other.y : l_i;
      const float m_hi = half ? m_i : other.x, l_hi = half ? l_i : other.y;
      const float m_row = two ? fmaxf(m_lo, m_hi) : m_lo;
      const float a_lo = ex2_approx(m_lo - m_row), a_hi = two ? ex2_approx(m_hi - m_row) : 0.f;
      const float l_row = fmaf(l_lo, a_lo, l_hi * a_hi);
      const float inv = 1.f / l_row;
      const float w_lo = a_lo * inv, w_hi = a_hi * inv;
      const int row = it.q0 + row_in_tile;
      __half* orow = out + ((long long)it.b * T + row) * (H * HD) + it.h * HD;
      for (int c = oc_begin; c < oc_end; c += 16) {
        uint32_t o[16], o2[16];
        tmem_ld16(tmem_o + lane_off + c, o);
        if (two) tmem_ld16(tmem_o + DK + lane_off + c, o2);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(o[i]) * w_lo;
        if (two) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(o2[i]), w_hi, v[i]);
        }
        if (row < T) {
          uint4* op = reinterpret_cast<uint4*>(orow + c);
          op[0] = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
          if (c + 8 < HD)
            op[1] = make_uint4(pack_f16(v[8], v[9]), pack_f16(v[10], v[11]), pack_f16(v[12], v[13]), pack_f16(v[14], v[15]));
        }
      }
      if (lse && half == 0 && row < T) lse[((long long)it.b * H + it.h) * T + row] = (m_row + log2f(l_row)) * kLn2;
      // every thread has read both accumulators and the exchange slots: the next item's P / O writes may start
      tc_fence_before();
      asm volatile("bar.sync 2, 256;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int HD>
int launch_fwd(const void* qkv, const int32_t* valid, void* out, float* lse, int32_t B, int32_t T, int32_t H, float scale,
               uint32_t drop_seed, float drop_p, cudaStream_t s) {
  CUtensorMap tm;
  // {d within the head, slot (q heads | k heads | v heads), frame, sample}: a box 64 columns wide reads one head and
  // zero-fills the columns beyond HD
  const int64_t dim[4] = {HD, 3LL * H, T, B};
  const int64_t stride[3] = {HD, 3LL * H * HD, 3LL * H * HD * T};
  int rc = fhb_make_tmap_bf16_4d(&tm, qkv, dim, stride, 64, kTQ, "qkv");
  if (rc) return rc;
  FHB_ONCE_PER_DEVICE({
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  });
  const long long items = (long long)((T + kTQ - 1) / kTQ) * H * B;
  const long long slots = 2LL * fhb_num_sms();
  dim3 grid((unsigned)(items < slots ? items : slots));
  if (drop_p > 0.f)
    FHB_CUDA_CHECK(fhb_launch((attn_fwd_tc_kernel<HD, true>), dim3(grid), dim3(288), kSmem, s, tm, valid, static_cast<__half*>(out), lse, T, H, B, scale,
                                                          drop_seed, fhb_dropout_thr16(drop_p), fhb_dropout_scale(drop_p)));
  else
    FHB_CUDA_CHECK(fhb_launch((attn_fwd_tc_kernel<HD, false>), dim3(grid), dim3(288), kSmem, s, tm, valid, static_cast<__half*>(out), lse, T, H, B, scale, 0u,
                                                           0u, 1.f));
  FHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Called by fhb_attn_fwd (attention.cu) for head_dim 64 / 40.
int fhb_attn_fwd_tc(const void* qkv, const int32_t* valid, void* out, float* lse, int32_t B, int32_t T, int32_t H,
                    int32_t d, float scale, uint32_t drop_seed, float drop_p, cudaStream_t s) {
  if (d == 64) return launch_fwd<64>(qkv, valid, out, lse, B, T, H, scale, drop_seed, drop_p, s);
  return launch_fwd<40>(qkv, valid, out, lse, B, T, H, scale, drop_seed, drop_p, s);
}
