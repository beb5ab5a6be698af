// K7 forward on 5th-gen tensor cores (head_dim 64: the HuBERT-Base / wav2vec2-Base teacher).
// Replaces the bmm -> masked_fill(-inf) -> fp32 softmax -> bmm chain of fairseq MultiheadAttention
// (reached from modules/module.py:558-564; the teacher runs the same code in fairseq).
//
// One CTA per (128-query tile, head, sample), 160 threads, 2 CTAs per SM (112 KB smem, 256 TMEM columns
// each) so one CTA's softmax overlaps the other's MMAs:
//   warp 4 (one elected lane): TMA loads of Q / K_j / V_j (128B swizzle, double-buffered K/V) and all
//                              tcgen05.mma issue:  S_j = Q K_j^T (128x128x64)  and  O_j = P_j V_j (128x64x128)
//   warps 0-3: one query row per thread.  tcgen05.ld S_j from TMEM, online softmax in registers (no
//              cross-thread reductions), P_j -> bf16 -> swizzled smem (A operand of the second MMA), then
//              O_j from TMEM, accumulated and rescaled in registers (o = (o + O_{j-1}) * alpha_j).
// Keys >= valid[b] are masked (P = 0) and key tiles entirely beyond valid[b] are skipped; padded QUERY
// rows are computed like any other.
#include "fhb_common.cuh"

namespace {

constexpr int kD = 64;
constexpr int kTQ = 128, kTK = 128;
constexpr uint32_t kTileBytes = kTQ * kD * 2;  // 16 KiB
constexpr uint32_t kSmem = kTileBytes /*Q*/ + 2 * kTileBytes /*K*/ + 2 * kTileBytes /*V*/ + 2 * kTileBytes /*P*/ + 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__global__ void __launch_bounds__(160, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const int* __restrict__ valid,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 3 * kTileBytes;
  uint8_t* sP = smem + 5 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTQ, h = blockIdx.y, b = blockIdx.z;
  int nvalid = valid ? valid[b] : T;
  nvalid = max(1, min(nvalid, T));
  const int nt = (nvalid + kTK - 1) / kTK;

  if (warp == 4 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // 128 fp32 columns
  const uint32_t tmem_o = tmem_base + 128;  // 64 fp32 columns

  if (warp == 4) {
    if (elect_one()) {
      const int HD = H * kD;
      auto load_kv = [&](int j) {
        const int st = j & 1;
        mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
        tma_load_3d(&tm_qkv, &kv_full[st], sK + st * kTileBytes, HD + h * kD, j * kTK, b);
        tma_load_3d(&tm_qkv, &kv_full[st], sV + st * kTileBytes, 2 * HD + h * kD, j * kTK, b);
      };
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(&tm_qkv, q_full, sQ, h * kD, q0, b);
      load_kv(0);
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S: A = Q (K-major), B = K (K-major)
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // O: A = P (K-major), B = V (MN-major)
      mbar_wait(q_full, 0);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          tc_mma_bf16(tmem_s, umma_desc_sw128(qa + k * 32, 0, 1024), umma_desc_sw128(ka + k * 32, 0, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
        tc_commit(s_full);
        if (j + 1 < nt) {  // prefetch the next K/V tile into the other stage once PV_{j-1} has released it
          mbar_wait(&kv_empty[(j + 1) & 1], (((j + 1) >> 1) & 1) ^ 1);
          load_kv(j + 1);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < kTK / 16; ++k)  // P: two 64-key atoms of 16 KiB; V: 16 key rows = 2 KiB per step
          tc_mma_bf16(tmem_o, umma_desc_sw128(pa + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024),
                      umma_desc_sw128(va + k * 2048, 0, 1024), idesc_o, k > 0 ? 1u : 0u);
        tc_commit(o_full);
        tc_commit(&kv_empty[st]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / accumulate (one row per thread)
    const int row_in_tile = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sc = scale * kLog2e;
    float m_i = -INFINITY, l_i = 0.f;
    float o_acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o_acc[i] = 0.f;
    const uint32_t rsw = (uint32_t)(row_in_tile & 7);
    const uint32_t prow = smem_u32(sP) + row_in_tile * 128;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j > 0) {  // fold in the previous tile's P V product before the running max moves
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kD; c += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_o + lane_off + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o_acc[c + i] += __uint_as_float(r[i]);
        }
      }
      const int k0 = j * kTK;
      const int nk = min(kTK, nvalid - k0);  // valid keys in this tile (>= 1)
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kTK; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_s + lane_off + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c + i < nk) ? __uint_as_float(r[i]) : -INFINITY);
      }
      const float m_new = fmaxf(m_i, mx * sc);
      const float alpha = ex2_approx(m_i - m_new);
      m_i = m_new;
      l_i *= alpha;
#pragma unroll
      for (int i = 0; i < kD; ++i) o_acc[i] *= alpha;
      // pass 2: probabilities -> bf16 -> swizzled smem (K-major A operand: two 64-key atoms)
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < kTK; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_s + lane_off + c, r);
        tmem_ld_wait();
        float pv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pe = ex2_approx(fmaf(__uint_as_float(r[i]), sc, -m_new));
          pv[i] = (c + i < nk) ? pe : 0.f;
          lsum += pv[i];
        }
        const uint32_t atom = (uint32_t)(c >> 6) * kTileBytes;
        const uint32_t ch = (uint32_t)((c & 63) >> 3);
        st_shared_v4(prow + atom + (((ch) ^ rsw) << 4), pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]),
                     pack_bf16(pv[4], pv[5]), pack_bf16(pv[6], pv[7]));
        st_shared_v4(prow + atom + (((ch + 1) ^ rsw) << 4), pack_bf16(pv[8], pv[9]), pack_bf16(pv[10], pv[11]),
                     pack_bf16(pv[12], pv[13]), pack_bf16(pv[14], pv[15]));
      }
      l_i += lsum;
      tc_fence_before();
      fence_async_shared();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (nt - 1) & 1);
    tc_fence_after();
    const int row = q0 + row_in_tile;
    const float inv = 1.f / l_i;
    __nv_bfloat16* orow = out + ((long long)b * T + row) * (H * kD) + h * kD;
#pragma unroll
    for (int c = 0; c < kD; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_o + lane_off + c, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (o_acc[c + i] + __uint_as_float(r[i])) * inv;
      if (row < T) {
        uint4* op = reinterpret_cast<uint4*>(orow + c);
        op[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        op[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
      }
    }
    if (lse && row < T) lse[((long long)b * H + h) * T + row] = (m_i + log2f(l_i)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// Called by fhb_attn_fwd (attention.cu) when head_dim == 64.
int fhb_attn_fwd_tc64(const void* qkv, const int32_t* valid, void* out, float* lse, int32_t B, int32_t T, int32_t H,
                      float scale, cudaStream_t s) {
  CUtensorMap tm;
  const int64_t dim[3] = {3LL * H * kD, T, B};
  const int64_t stride[2] = {3LL * H * kD, 3LL * H * kD * T};
  int rc = fhb_make_tmap_bf16_3d(&tm, qkv, dim, stride, kD, kTQ, "qkv");
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    FHB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  dim3 grid((T + kTQ - 1) / kTQ, H, B);
  attn_fwd_tc_kernel<<<grid, 160, kSmem, s>>>(tm, valid, static_cast<__nv_bfloat16*>(out), lse, T, H, scale);
  FHB_LAUNCH_CHECK();
  return 0;
}
