// K7: flash-style masked self-attention, forward + backward, fp16 in / fp32 softmax statistics.
// Reference: fairseq MultiheadAttention manual path reached from modules/module.py:558-564
// (bmm QK^T -> masked_fill(-inf on padded keys) -> fp32 softmax -> bmm PV); the T x T score matrix is
// never materialised here.  Padded QUERY rows are computed like any other row (their outputs enter the
// un-masked loss); only padded KEYS are masked, and key tiles entirely beyond valid[b] are
// skipped.
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 (bf16) with ldmatrix-fed fragments, 64-query x
// 64-key tiles, 4 warps per CTA.  Attention is ~8 % of the step's FLOPs; moving it to
// tcgen05 with S/P in TMEM is the planned next step.
#include <stdlib.h>

#include "fhb_common.cuh"

namespace {

constexpr int kTile = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// F16: fp16 operands (every 16-bit tensor of the library; gradients carry a loss scale); false: bf16
template <bool F16 = true>
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Load a [64 x d] bf16 tile (rows row0.., row stride ld) into smem [64][DP+8]; rows >= T zero-filled.
template <int DP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int row0, int T, int d) {
  const int chunks = d >> 3;
  for (int i = threadIdx.x; i < kTile * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i - r * chunks;
    const bool ok = row0 + r < T;
    const __nv_bfloat16* src = g + (long long)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(s + r * (DP + 8) + c * 8, src, ok);
  }
}
template <int DP>
__device__ __forceinline__ void zero_pad_cols(__nv_bfloat16* s, int d) {
  // columns [d, DP) are never written by load_tile: clear them once
  const int padc = DP - d;
  if (padc <= 0) return;
  for (int i = threadIdx.x; i < kTile * padc; i += blockDim.x) {
    const int r = i / padc, c = d + (i - r * padc);
    s[r * (DP + 8) + c] = __float2bfloat16(0.f);  // all-zero bits in either 16-bit format
  }
}

// A-operand fragments (16 rows x DP) of this warp's rows from a [64][DP+8] tile
template <int DP>
__device__ __forceinline__ void load_a_frags(const __nv_bfloat16* s, int warp, int lane, uint32_t (*f)[4]) {
#pragma unroll
  for (int kk = 0; kk < DP / 16; ++kk) {
    const uint32_t addr = smem_u32(s + (warp * 16 + (lane & 15)) * (DP + 8) + kk * 16 + (lane >> 4) * 8);
    ldsm_x4(addr, f[kk][0], f[kk][1], f[kk][2], f[kk][3]);
  }
}
// acc[8][4] (16 x 64) += A(frags, 16 x DP) * B^T where B tile is [64 rows (n)][DP (k)] in smem
template <int DP, bool F16 = true>
__device__ __forceinline__ void mma_a_bt(float (*acc)[4], const uint32_t (*af)[4], const __nv_bfloat16* bs, int lane) {
#pragma unroll
  for (int kk = 0; kk < DP / 16; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // pairs of n8 tiles
      uint32_t b0, b1, b2, b3;
      // matrices: (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
      const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int col = kk * 16 + ((lane >> 3) & 1) * 8;
      ldsm_x4(smem_u32(bs + row * (DP + 8) + col), b0, b1, b2, b3);
      mma16816<F16>(acc[2 * np], af[kk], b0, b1);
      mma16816<F16>(acc[2 * np + 1], af[kk], b2, b3);
    }
  }
}
// acc[DP/8][4] (16 x DP) += P(frags pf[4][4], 16 x 64) * B where B tile is [64 rows (k)][DP (n)] in smem
template <int DP, bool F16 = true>
__device__ __forceinline__ void mma_p_b(float (*acc)[4], const uint32_t (*pf)[4], const __nv_bfloat16* bs, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < DP / 16; ++np) {
      uint32_t b0, b1, b2, b3;
      // .trans matrices: (k 0-7,n 0-7) (k 8-15,n 0-7) (k 0-7,n 8-15) (k 8-15,n 8-15)
      const int row = kk * 16 + (lane & 15);
      const int col = np * 16 + (lane >> 4) * 8;
      ldsm_x4_t(smem_u32(bs + row * (DP + 8) + col), b0, b1, b2, b3);
      mma16816<F16>(acc[2 * np], pf[kk], b0, b1);
      mma16816<F16>(acc[2 * np + 1], pf[kk], b2, b3);
    }
  }
}
template <bool F16 = true>
__device__ __forceinline__ void acc_to_frags(const float (*s)[4], uint32_t (*pf)[4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    pf[kk][0] = pack16(s[2 * kk][0], s[2 * kk][1], F16);
    pf[kk][1] = pack16(s[2 * kk][2], s[2 * kk][3], F16);
    pf[kk][2] = pack16(s[2 * kk + 1][0], s[2 * kk + 1][1], F16);
    pf[kk][3] = pack16(s[2 * kk + 1][2], s[2 * kk + 1][3], F16);
  }
}

// ------------------------------------------------------------------ forward
template <int DP>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ valid, __nv_bfloat16* __restrict__ out,
                float* __restrict__ lse, int T, int H, int d, float scale, uint32_t drop_seed, uint32_t drop_thr,
                float drop_scale) {
  __shared__ __align__(16) __nv_bfloat16 Qs[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Ks[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Vs[kTile * (DP + 8)];
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const long long ld = 3LL * H * d;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld;
  int nvalid = valid ? valid[b] : T;
  nvalid = max(1, min(nvalid, T));
  zero_pad_cols<DP>(Qs, d);
  zero_pad_cols<DP>(Ks, d);
  zero_pad_cols<DP>(Vs, d);
  load_tile<DP>(Qs, base + h * d, ld, q0, T, d);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[DP / 16][4];
  load_a_frags<DP>(Qs, warp, lane, qf);
  float o[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
  const float sc = scale * kLog2e;
  const int ntiles = (nvalid + kTile - 1) / kTile;
  for (int kt = 0; kt < ntiles; ++kt) {
    const int k0 = kt * kTile;
    __syncthreads();
    load_tile<DP>(Ks, base + (long long)H * d + h * d, ld, k0, T, d);
    load_tile<DP>(Vs, base + 2LL * H * d + h * d, ld, k0, T, d);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_a_bt<DP>(s, qf, Ks, lane);  // fp16 q, k
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = k0 + i * 8 + 2 * tq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = (col + (e & 1)) < nvalid;
        s[i][e] = ok ? s[i][e] * sc : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[i][e]);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_i[r], mx[r]);
      alpha[r] = ex2_approx(m_i[r] - m_new);
      m_i[r] = m_new;
      l_i[r] *= alpha[r];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[i][e] = ex2_approx(s[i][e] - m_i[e >> 1]);
        l_i[e >> 1] += s[i][e];
      }
    }
    if (drop_thr) {  // attention dropout on the probabilities (l stays un-dropped); pairs = adjacent keys
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t q = (uint32_t)((b * H + h) * T + q0 + warp * 16 + g + r * 8);
          float m0, m1;
          dropout_pair(drop_seed, q * (uint32_t)((T + 1) >> 1) + (uint32_t)((k0 + i * 8 + 2 * tq) >> 1), drop_thr, drop_scale,
                       m0, m1);
          s[i][2 * r] *= m0;
          s[i][2 * r + 1] *= m1;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      o[i][0] *= alpha[0];
      o[i][1] *= alpha[0];
      o[i][2] *= alpha[1];
      o[i][3] *= alpha[1];
    }
    uint32_t pf[4][4];
    acc_to_frags(s, pf);  // probabilities -> fp16
    mma_p_b<DP>(o, pf, Vs, lane);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
  const int HD = H * d;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= T) continue;
    const float inv = 1.f / l_i[r];
    __nv_bfloat16* orow = out + ((long long)b * T + row) * HD + h * d;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * tq;
      if (col < d) *reinterpret_cast<uint32_t*>(orow + col) = pack_f16(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
    }
    if (lse && tq == 0) lse[((long long)b * H + h) * T + row] = (m_i[r] + log2f(l_i[r])) * kLn2;
  }
}

// ------------------------------------------------------------------ backward
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int B, int T, int H, int d) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, t, h)
  if (idx >= (long long)B * T * H) return;
  const int h = idx % H;
  const long long bt = idx / H;
  const int t = bt % T;
  const int b = bt / T;
  const __nv_bfloat16* po = o + bt * H * d + h * d;
  const __nv_bfloat16* pd = dout + bt * H * d + h * d;
  float s = 0.f;
  for (int c = 0; c < d; c += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(po + c), e = *reinterpret_cast<const uint4*>(pd + c);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, eu[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_f16(au[j]), y = unpack_f16(eu[j]);
      s += x.x * y.x + x.y * y.y;
    }
  }
  delta[((long long)b * H + h) * T + t] = s;
}

// dK, dV: one CTA per (64-key tile, head, sample); warp owns 16 keys; loops over query tiles.
template <int DP>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ valid,
                    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int H, int d, float scale,
                    uint32_t drop_seed, uint32_t drop_thr, float drop_scale) {
  __shared__ __align__(16) __nv_bfloat16 Ks[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Vs[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Qs[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Ds[kTile * (DP + 8)];
  __shared__ float lse_s[kTile], delta_s[kTile];
  const int k0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const long long ld = 3LL * H * d;
  const int HD = H * d;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld;
  const __nv_bfloat16* dob = dout + (long long)b * T * HD + h * d;
  __nv_bfloat16* dbase = dqkv + (long long)b * T * ld;
  int nvalid = valid ? valid[b] : T;
  nvalid = max(1, min(nvalid, T));
  float dk[DP / 8][4], dv[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  if (k0 < nvalid) {  // block-uniform: tiles of fully masked keys get zero gradients
    zero_pad_cols<DP>(Ks, d);
    zero_pad_cols<DP>(Vs, d);
    zero_pad_cols<DP>(Qs, d);
    zero_pad_cols<DP>(Ds, d);
    load_tile<DP>(Ks, base + (long long)H * d + h * d, ld, k0, T, d);
    load_tile<DP>(Vs, base + 2LL * H * d + h * d, ld, k0, T, d);
    cp_async_wait_all();
    __syncthreads();
    uint32_t kf[DP / 16][4], vf[DP / 16][4];
    load_a_frags<DP>(Ks, warp, lane, kf);
    load_a_frags<DP>(Vs, warp, lane, vf);
    const float sc = scale * kLog2e;
    const int key_row[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
    const int nq = (T + kTile - 1) / kTile;
    for (int qt = 0; qt < nq; ++qt) {
      const int q0 = qt * kTile;
      __syncthreads();
      load_tile<DP>(Qs, base + h * d, ld, q0, T, d);
      load_tile<DP>(Ds, dob, HD, q0, T, d);
      if (threadIdx.x < kTile) {
        const int q = q0 + threadIdx.x;
        lse_s[threadIdx.x] = q < T ? lse[((long long)b * H + h) * T + q] * kLog2e : INFINITY;
        delta_s[threadIdx.x] = q < T ? delta[((long long)b * H + h) * T + q] : 0.f;
      }
      cp_async_wait_all();
      __syncthreads();
      float st[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
      mma_a_bt<DP>(st, kf, Qs, lane);  // S^T[key][q]
      mma_a_bt<DP>(dp, vf, Ds, lane);        // dP^T[key][q] = V dO^T
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qc = i * 8 + 2 * tq + (e & 1);
          const bool ok = key_row[e >> 1] < nvalid;
          const float p = ok ? ex2_approx(st[i][e] * sc - lse_s[qc]) : 0.f;
          float mk = 1.f;  // dropout multiplier of P[q][key] (0 or 1/(1-p)); same hash as the forward
          if (drop_thr) {
            const uint32_t q = (uint32_t)((b * H + h) * T + q0 + qc);
            mk = dropout_one(drop_seed, q * (uint32_t)(2 * ((T + 1) >> 1)) + (uint32_t)key_row[e >> 1], drop_thr, drop_scale);
          }
          st[i][e] = p * mk;                                        // dropped P -> dV
          dp[i][e] = p * (dp[i][e] * mk - delta_s[qc]) * scale;   // dS = P o (dP_dropped o mask - delta)
        }
      }
      uint32_t pf[4][4];
      acc_to_frags(st, pf);
      mma_p_b<DP>(dv, pf, Ds, lane);  // dV += P^T dO
      acc_to_frags(dp, pf);
      mma_p_b<DP>(dk, pf, Qs, lane);  // dK += dS^T Q 
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = k0 + warp * 16 + g + r * 8;
    if (row >= T) continue;
    __nv_bfloat16* krow = dbase + (long long)row * ld + HD + h * d;
    __nv_bfloat16* vrow = dbase + (long long)row * ld + 2 * HD + h * d;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * tq;
      if (col < d) {
        *reinterpret_cast<uint32_t*>(krow + col) = pack_f16(dk[i][2 * r], dk[i][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(vrow + col) = pack_f16(dv[i][2 * r], dv[i][2 * r + 1]);
      }
    }
  }
}

// dQ: one CTA per (64-query tile, head, sample); warp owns 16 queries; loops over valid key tiles.
template <int DP>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ valid,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int H, int d, float scale,
                   uint32_t drop_seed, uint32_t drop_thr, float drop_scale) {
  __shared__ __align__(16) __nv_bfloat16 Qs[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Ds[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Ks[kTile * (DP + 8)];
  __shared__ __align__(16) __nv_bfloat16 Vs[kTile * (DP + 8)];
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const long long ld = 3LL * H * d;
  const int HD = H * d;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld;
  int nvalid = valid ? valid[b] : T;
  nvalid = max(1, min(nvalid, T));
  zero_pad_cols<DP>(Qs, d);
  zero_pad_cols<DP>(Ds, d);
  zero_pad_cols<DP>(Ks, d);
  zero_pad_cols<DP>(Vs, d);
  load_tile<DP>(Qs, base + h * d, ld, q0, T, d);
  load_tile<DP>(Ds, dout + (long long)b * T * HD + h * d, HD, q0, T, d);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[DP / 16][4], dof[DP / 16][4];
  load_a_frags<DP>(Qs, warp, lane, qf);
  load_a_frags<DP>(Ds, warp, lane, dof);
  float lse_r[2], delta_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    lse_r[r] = row < T ? lse[((long long)b * H + h) * T + row] * kLog2e : INFINITY;
    delta_r[r] = row < T ? delta[((long long)b * H + h) * T + row] : 0.f;
  }
  float dq[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const float sc = scale * kLog2e;
  const int ntiles = (nvalid + kTile - 1) / kTile;
  for (int kt = 0; kt < ntiles; ++kt) {
    const int k0 = kt * kTile;
    __syncthreads();
    load_tile<DP>(Ks, base + (long long)H * d + h * d, ld, k0, T, d);
    load_tile<DP>(Vs, base + 2LL * H * d + h * d, ld, k0, T, d);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    mma_a_bt<DP>(s, qf, Ks, lane);            // S = Q K^T
    mma_a_bt<DP>(dp, dof, Vs, lane);   // dP = dO V^T 
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = (k0 + i * 8 + 2 * tq + (e & 1)) < nvalid;
        const float p = ok ? ex2_approx(s[i][e] * sc - lse_r[e >> 1]) : 0.f;
        float mk = 1.f;
        if (drop_thr) {
          const uint32_t q = (uint32_t)((b * H + h) * T + q0 + warp * 16 + g + (e >> 1) * 8);
          mk = dropout_one(drop_seed, q * (uint32_t)(2 * ((T + 1) >> 1)) + (uint32_t)(k0 + i * 8 + 2 * tq + (e & 1)), drop_thr,
                           drop_scale);
        }
        dp[i][e] = p * (dp[i][e] * mk - delta_r[e >> 1]) * scale;
      }
    }
    uint32_t pf[4][4];
    acc_to_frags(dp, pf);
    mma_p_b<DP>(dq, pf, Ks, lane);  // dQ += dS K 
  }
  __nv_bfloat16* dbase = dqkv + (long long)b * T * ld;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= T) continue;
    __nv_bfloat16* qrow = dbase + (long long)row * ld + h * d;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * tq;
      if (col < d) *reinterpret_cast<uint32_t*>(qrow + col) = pack_f16(dq[i][2 * r], dq[i][2 * r + 1]);
    }
  }
}

int check_shape(int B, int T, int H, int d) {
  FHB_ARG_CHECK(B > 0 && T > 0 && H > 0, "attn: empty problem");
  FHB_ARG_CHECK(d % 8 == 0 && d >= 8 && d <= 64, "attn: head_dim %d must be a multiple of 8 in [8, 64]", d);
  return 0;
}

}  // namespace

#define FHB_ATTN_DISPATCH(DPV, CALL) \
  if (DPV <= 16) { constexpr int DP = 16; CALL; } \
  else if (DPV <= 32) { constexpr int DP = 32; CALL; } \
  else if (DPV <= 48) { constexpr int DP = 48; CALL; } \
  else { constexpr int DP = 64; CALL; }

int fhb_attn_fwd_tc(const void* qkv, const int32_t* valid, void* out, float* lse, int32_t B, int32_t T, int32_t H,
                    int32_t d, float scale, uint32_t drop_seed, float drop_p, cudaStream_t s);  // attention_tc.cu (tcgen05 path, head_dim 64 / 40)

int check_drop(int32_t B, int32_t T, int32_t H, float drop_p) {
  FHB_ARG_CHECK(drop_p >= 0.f && drop_p < 1.f, "attention: drop_p=%f must be in [0, 1)", (double)drop_p);
  FHB_ARG_CHECK(drop_p == 0.f || (long long)B * H * T * (T + 1) < (1LL << 32), "attention: dropout index space exceeds 32 bits");
  return 0;
}

extern "C" int fhb_attn_fwd(const void* qkv, const int32_t* valid, void* out, float* lse, int32_t B, int32_t T,
                            int32_t H, int32_t d, float scale, uint32_t drop_seed, float drop_p, fhb_stream_t stream) {
  int rc = check_shape(B, T, H, d);
  if (rc) return rc;
  if ((rc = check_drop(B, T, H, drop_p)) != 0) return rc;
  const uint32_t thr = drop_p > 0.f ? fhb_dropout_thr16(drop_p) : 0u;
  const float dsc = fhb_dropout_scale(drop_p);
  FHB_ARG_CHECK(qkv && out, "attn_fwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool no_tc = getenv("FHB_ATTN_NO_TC") != nullptr;
  if ((d == 64 || d == 40) && !no_tc) return fhb_attn_fwd_tc(qkv, valid, out, lse, B, T, H, d, scale, drop_seed, drop_p, s);
  dim3 grid((T + kTile - 1) / kTile, H, B);
  FHB_ATTN_DISPATCH(d, (attn_fwd_kernel<DP><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), valid,
                                                                   static_cast<__nv_bfloat16*>(out), lse, T, H, d, scale,
                                                                   drop_seed, thr, dsc)));
  FHB_LAUNCH_CHECK();
  return 0;
}

int fhb_attn_bwd_tc(const void* qkv, const int32_t* valid, const void* dout, const float* lse, const float* delta,
                    void* dqkv, float* dq_ws, int32_t B, int32_t T, int32_t H, int32_t d, float scale, uint32_t drop_seed,
                    float drop_p, cudaStream_t s);  // attention_bwd_tc.cu (tcgen05 path, head_dim 40 / 64)

extern "C" int fhb_attn_bwd(const void* qkv, const int32_t* valid, const void* out, const void* dout, const float* lse,
                            void* dqkv, float* delta_ws, float* dq_ws, int32_t B, int32_t T, int32_t H, int32_t d,
                            float scale, uint32_t drop_seed, float drop_p, fhb_stream_t stream) {
  int rc = check_shape(B, T, H, d);
  if (rc) return rc;
  if ((rc = check_drop(B, T, H, drop_p)) != 0) return rc;
  const uint32_t thr = drop_p > 0.f ? fhb_dropout_thr16(drop_p) : 0u;
  const float dsc = fhb_dropout_scale(drop_p);
  FHB_ARG_CHECK(qkv && out && dout && lse && dqkv && delta_ws, "attn_bwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * T * H;
  fhb_pdl_hint(true);
  FHB_CUDA_CHECK(fhb_launch(attn_delta_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, static_cast<const __nv_bfloat16*>(out),
                                                                static_cast<const __nv_bfloat16*>(dout), delta_ws, B, T, H, d));
  FHB_LAUNCH_CHECK();
  static const bool no_tc = getenv("FHB_ATTN_NO_TC") != nullptr;
  if ((d == 64 || d == 40) && dq_ws && !no_tc)
    return fhb_attn_bwd_tc(qkv, valid, dout, lse, delta_ws, dqkv, dq_ws, B, T, H, d, scale, drop_seed, drop_p, s);
  dim3 grid((T + kTile - 1) / kTile, H, B);
  FHB_ATTN_DISPATCH(d, (attn_bwd_dkv_kernel<DP><<<grid, 128, 0, s>>>(
                           static_cast<const __nv_bfloat16*>(qkv), valid, static_cast<const __nv_bfloat16*>(dout), lse,
                           delta_ws, static_cast<__nv_bfloat16*>(dqkv), T, H, d, scale, drop_seed, thr, dsc)));
  FHB_LAUNCH_CHECK();
  FHB_ATTN_DISPATCH(d, (attn_bwd_dq_kernel<DP><<<grid, 128, 0, s>>>(
                           static_cast<const __nv_bfloat16*>(qkv), valid, static_cast<const __nv_bfloat16*>(dout), lse,
                           delta_ws, static_cast<__nv_bfloat16*>(dqkv), T, H, d, scale, drop_seed, thr, dsc)));
  FHB_LAUNCH_CHECK();
  return 0;
}
