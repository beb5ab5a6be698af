// K7 forward on 5th-gen tensor cores, head_dim 64 (HuBERT-Base / wav2vec2-Base teacher) and head_dim 40
// (FitHuBERT student: 3 k-steps of 16 with the 8 pad columns of Q zeroed in shared memory, N = 48 for P V).
// Replaces the bmm -> masked_fill(-inf) -> fp32 softmax -> bmm chain of fairseq MultiheadAttention
// (reached from modules/module.py:558-564; the teacher runs the same code in fairseq).
//
// One CTA per (128-query tile, head, sample), 288 threads, 2 CTAs per SM (113 KB smem, 256 TMEM columns
// each) so one CTA's softmax overlaps the other's MMAs:
//   warp 8 (one elected lane): TMA loads of Q / K_j / V_j (128B swizzle, double-buffered K/V) and all
//                              tcgen05.mma issue:  S_j = Q K_j^T (128x128xd)  and  O += P_j V_j (128xdx128)
//   warps 0-7: thread = (query row, 64-key half).  S_j is pulled out of TMEM in one batch of tcgen05.ld (which
//              frees the S columns for S_{j+1} while the exponentials run), online softmax in registers - the two
//              halves of a row only exchange their maxima through smem once per key tile - and P_j -> fp16 ->
//              swizzled smem (A operand of the second MMA).  16 softmax warps per SM hide the TMEM / MUFU latency.
//   Ragged edges are trimmed: the last key tile issues N = ceil16(valid keys) and contracts over that many keys
//   only; warps whose 32 query rows are all >= T skip the exponentials.
//   O stays in TMEM for the whole key loop (accumulating MMAs).  The running maximum is only moved - and O
//   rescaled in TMEM (tcgen05.ld / st) - when it grows by more than 2^8, so the rescale is off the
//   steady-state path; l and the LSE stay exact because P, l and O share the same reference maximum.
// Keys >= valid[b] are masked (P = 0) and key tiles entirely beyond valid[b] are skipped; padded QUERY
// rows are computed like any other.
#include "fhb_common.cuh"

namespace {

constexpr int kTQ = 128, kTK = 128;
constexpr uint32_t kTileBytes = kTQ * 64 * 2;  // 16 KiB: 128 rows x 64 bf16 (one 128-byte swizzle row each)
constexpr uint32_t kSmem = kTileBytes /*Q*/ + 2 * kTileBytes /*K*/ + 2 * kTileBytes /*V*/ + 2 * kTileBytes /*P*/ + 512 /*xch*/ + 128;
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

// HD: logical head dim (64 or 40); DK = HD rounded up to the UMMA k-step / n-step of 16
template <int HD, bool DROP>
__global__ void __launch_bounds__(288, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const int* __restrict__ valid,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale,
                   uint32_t drop_seed, uint32_t drop_thr, float drop_scale) {
  constexpr int DK = (HD + 15) / 16 * 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 3 * kTileBytes;
  uint8_t* sP = smem + 5 * kTileBytes;
  // [2][128] row maxima of the two halves, exchanged as bf16: both halves use the same ROUNDED pair, so they agree
  // exactly on the reference maximum (which only has to stay within 2^8 of the true one)
  __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(smem + 7 * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes + 512);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;     // 256 arrivals: S_j has been copied to registers
  uint64_t* p_full = bars + 7;     // 256 arrivals: P_j is in smem (and O has been rescaled if needed)
  uint64_t* o_done = bars + 8;     // PV_j retired: P smem reusable, O readable
  uint64_t* q_ready = bars + 9;    // 256 arrivals: pad columns of Q zeroed (HD % 16 != 0 only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTQ, h = blockIdx.y, b = blockIdx.z;

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(o_done, 1);
    mbar_init(q_ready, 256);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // 128 fp32 columns
  const uint32_t tmem_o = tmem_base + 128;  // DK fp32 columns
  pdl_sync();  // the prologue above overlapped the previous kernel's tail; global memory only from here on
  int nvalid = valid ? valid[b] : T;
  nvalid = max(1, min(nvalid, T));
  const int nt = (nvalid + kTK - 1) / kTK;

  if (warp == 8) {
    if (elect_one()) {
      const int HD_all = H * HD;
      auto load_kv = [&](int j) {
        const int st = j & 1;
        mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
        tma_load_3d(&tm_qkv, &kv_full[st], sK + st * kTileBytes, HD_all + h * HD, j * kTK, b);
        tma_load_3d(&tm_qkv, &kv_full[st], sV + st * kTileBytes, 2 * HD_all + h * HD, j * kTK, b);
      };
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(&tm_qkv, q_full, sQ, h * HD, q0, b);
      load_kv(0);
      if (nt > 1) load_kv(1);
      const uint32_t idesc_o = umma_idesc_16(128, DK, 0, 1, 0, 0);   // O: A = P (K-major), B = V (MN-major), both fp16
      const uint32_t qa = smem_u32(sQ), pa = smem_u32(sP);
      auto nk16 = [&](int j) { return (min(kTK, nvalid - j * kTK) + 15) & ~15; };  // valid keys of tile j, rounded
      auto issue_s = [&](int j) {
        const uint32_t ka = smem_u32(sK + (j & 1) * kTileBytes);
        const uint32_t idesc_s = umma_idesc_16(128, (uint32_t)nk16(j), 0, 0, 0, 0);  // S: A = Q, B = K, both K-major, fp16
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
          tc_mma_bf16(tmem_s, umma_desc_sw128(qa + k * 32, 0, 1024), umma_desc_sw128(ka + k * 32, 0, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
        tc_commit(s_full);
      };
      if (HD % 16) mbar_wait(q_ready, 0); else mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1;
        if (j + 1 < nt) {  // S_{j+1} as soon as the softmax warps have S_j in registers: overlaps their exponentials
          mbar_wait(&kv_full[st ^ 1], ((j + 1) >> 1) & 1);
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t va = smem_u32(sV + st * kTileBytes);
        const int ksteps = nk16(j) / 16;
        for (int k = 0; k < ksteps; ++k)  // P: two 64-key atoms of 16 KiB; V: 16 key rows = 2 KiB per step
          tc_mma_bf16(tmem_o, umma_desc_sw128(pa + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024),
                      umma_desc_sw128(va + k * 2048, 0, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        tc_commit(o_done);
        tc_commit(&kv_empty[st]);
        if (j + 2 < nt) {  // refill this K/V stage once S_j / PV_j have retired
          mbar_wait(&kv_empty[st], (j >> 1) & 1);
          load_kv(j + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax: thread = (query row, 64-key half)
    const int quarter = warp & 3, half = warp >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t rsw = (uint32_t)(row_in_tile & 7);
    const bool warp_live = q0 + quarter * 32 < T;  // warp-uniform: some query row of this warp exists
    if constexpr (HD % 16 != 0) {
      // columns HD..DK-1 of this Q row hold the next head's values: zero them so they drop out of Q K^T
      mbar_wait(q_full, 0);
      static_assert(HD % 8 == 0 && DK - HD == 8, "pad is one 16-byte chunk");
      if (half == 0) {
        st_shared_v4(smem_u32(sQ) + row_in_tile * 128 + ((((uint32_t)HD >> 3) ^ rsw) << 4), 0u, 0u, 0u, 0u);
        fence_async_shared();
      }
      mbar_arrive(q_ready);
    }
    const float sc = scale * kLog2e;
    float m_i = -INFINITY, l_i = 0.f;  // m_i: reference maximum of the whole row (shared by both halves), l_i: this half
    const uint32_t prow = smem_u32(sP) + half * kTileBytes + row_in_tile * 128;
    // dropout pair index of (b, h, q, k): ((b*H + h)*T + q) * ceil(T/2) + (k >> 1)
    const uint32_t drop_row = (uint32_t)(((b * H + h) * T + q0 + row_in_tile) * ((T + 1) >> 1));
    // O columns owned by this thread for rescale / final store: half 0 -> [0, 32), half 1 -> [32, DK)
    const int oc_begin = half ? 32 : 0, oc_end = half ? DK : 32;
    for (int j = 0; j < nt; ++j) {
      const int nk = nvalid - j * kTK;             // valid keys in this tile (>= 1; >= 128 for all but the last tile)
      const int nkh = nk - half * 64;              // ... of which fall after this half's first column
      const int n16 = ((min(kTK, nk) + 15) & ~15) - half * 64;  // columns of this half the MMAs touch (<= 0: none)
      const bool work = warp_live && n16 > 0;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t r[64];
      if (work) {
        tmem_ld32(tmem_s + lane_off + half * 64, r);
        if (n16 > 32) tmem_ld32(tmem_s + lane_off + half * 64 + 32, r + 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(s_free);
      // 8 independent max chains over this half's columns
      float mx = -INFINITY;
      if (work) {
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          if (c < n16) {
            if (c + 16 <= nkh) {
#pragma unroll
              for (int i = 0; i < 16; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(r[c + i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], c + i < nkh ? __uint_as_float(r[c + i]) : -INFINITY);
            }
          }
        }
        mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      }
      // the two halves of a row agree on the row maximum through smem (one named barrier per key tile)
      const __nv_bfloat16 mxr = __float2bfloat16_ru(mx);
      xch[half * kTQ + row_in_tile] = mxr;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(__bfloat162float(mxr), __bfloat162float(xch[(half ^ 1) * kTQ + row_in_tile]));
      const float m_new = fmaxf(m_i, mx * sc);
      if (j == 0) {
        m_i = m_new;
      } else {
        const bool need = (m_new - m_i) > kRescaleThreshold;
        const bool any = __any_sync(0xffffffffu, need);
        mbar_wait(o_done, (j - 1) & 1);  // PV_{j-1} retired: P smem is free, O is complete up to tile j-1
        if (any) {
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_i - m_new) : 1.f;
          if (need) m_i = m_new;
          l_i *= alpha;
          for (int c = oc_begin; c < oc_end; c += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_off + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_o + lane_off + c, o);
          }
          tmem_st_wait();
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // xch may be rewritten by the next tile only after both halves read it
      // probabilities -> fp16 (at most 2^8 after the lazy rescale: in range) -> swizzled smem (K-major A operand: this half's 64-key atom)
      if (work) {
        float ls4[4] = {0.f, 0.f, 0.f, 0.f};
        const float neg_m = -m_i;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          if (c < n16) {
            float pv[16];
            if (c + 16 <= nkh) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(r[c + i]), sc, neg_m));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                pv[i] = (c + i < nkh) ? ex2_approx(fmaf(__uint_as_float(r[c + i]), sc, neg_m)) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) ls4[i & 3] += pv[i];
            if (DROP) {  // attention dropout on the probabilities (the row sum l stays un-dropped)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float m0, m1;
                dropout_pair(drop_seed, drop_row + (uint32_t)((j * kTK + half * 64 + c) >> 1) + i, drop_thr, drop_scale, m0,
                             m1);
                pv[2 * i] *= m0;
                pv[2 * i + 1] *= m1;
              }
            }
            const uint32_t ch = (uint32_t)(c >> 3);
            st_shared_v4(prow + (((ch) ^ rsw) << 4), pack_f16(pv[0], pv[1]), pack_f16(pv[2], pv[3]),
                         pack_f16(pv[4], pv[5]), pack_f16(pv[6], pv[7]));
            st_shared_v4(prow + (((ch + 1) ^ rsw) << 4), pack_f16(pv[8], pv[9]), pack_f16(pv[10], pv[11]),
                         pack_f16(pv[12], pv[13]), pack_f16(pv[14], pv[15]));
          }
        }
        l_i += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      }
      tc_fence_before();
      fence_async_shared();
      mbar_arrive(p_full);
    }
    mbar_wait(o_done, (nt - 1) & 1);
    tc_fence_after();
    // row sum of both halves, exchanged through the (now idle) P tile
    float* lx = reinterpret_cast<float*>(sP);
    lx[half * kTQ + row_in_tile] = l_i;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_row = l_i + lx[(half ^ 1) * kTQ + row_in_tile];
    const int row = q0 + row_in_tile;
    const float inv = 1.f / l_row;
    __nv_bfloat16* orow = out + ((long long)b * T + row) * (H * HD) + h * HD;
    for (int c = oc_begin; c < oc_end; c += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_o + lane_off + c, o);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(o[i]) * inv;
      if (row < T) {
        uint4* op = reinterpret_cast<uint4*>(orow + c);
        op[0] = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
        if (c + 8 < HD)
          op[1] = make_uint4(pack_f16(v[8], v[9]), pack_f16(v[10], v[11]), pack_f16(v[12], v[13]), pack_f16(v[14], v[15]));
      }
    }
    if (lse && half == 0 && row < T) lse[((long long)b * H + h) * T + row] = (m_i + log2f(l_row)) * kLn2;
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
  const int64_t dim[3] = {3LL * H * HD, T, B};
  const int64_t stride[2] = {3LL * H * HD, 3LL * H * HD * T};
  int rc = fhb_make_tmap_bf16_3d(&tm, qkv, dim, stride, 64, kTQ, "qkv");
  if (rc) return rc;
  FHB_ONCE_PER_DEVICE({
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  });
  dim3 grid((T + kTQ - 1) / kTQ, H, B);
  if (drop_p > 0.f)
    FHB_CUDA_CHECK(fhb_launch((attn_fwd_tc_kernel<HD, true>), dim3(grid), dim3(288), kSmem, s, tm, valid, static_cast<__nv_bfloat16*>(out), lse, T, H, scale,
                                                          drop_seed, fhb_dropout_thr16(drop_p), fhb_dropout_scale(drop_p)));
  else
    FHB_CUDA_CHECK(fhb_launch((attn_fwd_tc_kernel<HD, false>), dim3(grid), dim3(288), kSmem, s, tm, valid, static_cast<__nv_bfloat16*>(out), lse, T, H, scale, 0u,
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
