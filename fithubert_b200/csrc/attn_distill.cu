// Attention-map / value-relation distillation: the LAST encoder layer's un-normalised attention logits
// and value-relation maps of student and teacher, their transfer losses and the gradient back into q, k, v.
//
// Reference: utils/utils.py:190-258 (`rtrn_attn_forward`: fairseq MultiheadAttention with before_softmax=True hands back
// the masked logits bmm(q * scaling, k^T) and the value heads; v_rel = bmm(v * scaling, v^T)), train.py:327-368 (the
// losses).  Only the last layer's maps enter the loss, so this path materialises T x T maps for that one layer
// (fp32 [B*H][T][pitch]); every other layer keeps the flash kernels, which never do.
//
//   fhb_attn_scores      S[bh][i][j] = scale * <a[b, i, h, :], b[b, j, h, :]>, -inf at keys j >= valid[b]
//   fhb_attn_map_loss    mse over the keys neither side masks (train.py:331-341) or KL(softmax(t) || softmax(s)) per query
//                        row (:342-349, :357-364); writes the loss-scaled fp16 gradient wrt the student map
//   fhb_attn_scores_bwd  out[b, r, h, :] (+)= alpha * sum_c G[bh][r][c] * m[b, c, h, :]      (trans = 0: dQ = dS K)
//                        out[b, r, h, :] (+)= alpha * sum_c G[bh][c][r] * m[b, c, h, :]      (trans = 1: dK = dS^T Q)
//
// The contractions are warp-level mma.sync.m16n8k16 (fp16 in, fp32 accumulate) on 64 x 64 tiles: the maps are written /
// read once per step for one layer (HBM-bound: 4 T^2 bytes per head against 2 T^2 d flops), not a tensor-pipe problem.
#include <math.h>
#include <stdlib.h>

#include "fhb_common.cuh"

namespace {

constexpr int kTile = 64;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
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

// [64 rows][W cols] fp16 tile into smem rows of W + 8 elements; source rows row0.. (stride ld), columns col0..; rows >=
// n_rows and 8-element chunks starting at or beyond n_cols are zero-filled
template <int W>
__device__ __forceinline__ void load_tile(__half* s, const __half* g, long long ld, int row0, int n_rows, int col0, int n_cols) {
  constexpr int chunks = W / 8;
  for (int i = threadIdx.x; i < kTile * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i - r * chunks;
    const bool ok = row0 + r < n_rows && col0 + c * 8 < n_cols;
    const __half* src = g + (long long)(ok ? row0 + r : 0) * ld + (ok ? col0 + c * 8 : 0);
    cp_async16(s + r * (W + 8) + c * 8, src, ok);
  }
}

// ------------------------------------------------------------------ scores
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One CTA = one 64-query block of one (sample, head): the query fragments stay in registers while the key tiles stream
// through a double-buffered shared tile (cp.async), each 64 x 64 product is scaled, masked and stored straight from the
// accumulator fragments (32-byte row pieces: measured faster than staging the tile for 256-byte rows).
template <int DP>
__global__ void __launch_bounds__(128) attn_scores_kernel(const __half* __restrict__ a, const __half* __restrict__ b, long long ld,
                                                          const int* __restrict__ valid, float* __restrict__ out,
                                                          long long pitch, int T, int H, int d, float scale) {
  __shared__ __align__(16) __half sa[kTile * (DP + 8)];
  __shared__ __align__(16) __half sb[2][kTile * (DP + 8)];
  pdl_sync();
  const int m0 = blockIdx.x * kTile, bh = blockIdx.y;
  const int bi = bh / H, h = bh - bi * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* ga = a + (long long)bi * T * ld + h * d;
  const __half* gb = b + (long long)bi * T * ld + h * d;
  const int nkt = (T + kTile - 1) / kTile;
  // d is a multiple of 8: whole 16-byte chunks; chunks at or beyond d are zero-filled (padding of the contraction)
  load_tile<DP>(sa, ga, ld, m0, T, 0, d);
  load_tile<DP>(sb[0], gb, ld, 0, T, 0, d);
  cp_async_commit();
  const int nvalid = valid ? valid[bi] : T;
  const float ninf = -INFINITY;
  float* o = out + (long long)bh * T * pitch;
  uint32_t af[DP / 16][4];
  for (int kt = 0; kt < nkt; ++kt) {
    if (kt + 1 < nkt) {
      load_tile<DP>(sb[(kt + 1) & 1], gb, ld, (kt + 1) * kTile, T, 0, d);
      cp_async_commit();
      cp_async_wait_group<1>();  // everything but the tile just requested has landed
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk)
        ldsm_x4(smem_u32(sa + (warp * 16 + (lane & 15)) * (DP + 8) + kk * 16 + (lane >> 4) * 8), af[kk][0], af[kk][1],
                af[kk][2], af[kk][3]);
    }
    const __half* sbt = sb[kt & 1];
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DP / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int col = kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sbt + row * (DP + 8) + col), b0, b1, b2, b3);
        mma16816(acc[2 * np], af[kk], b0, b1);
        mma16816(acc[2 * np + 1], af[kk], b2, b3);
      }
    }
    const int n0 = kt * kTile;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = n0 + nt * 8 + 2 * (lane & 3);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = m0 + warp * 16 + (lane >> 2) + hh * 8;
        if (row >= T || col >= T) continue;
        const float v0 = col < nvalid ? acc[nt][2 * hh] * scale : ninf;
        const float v1 = col + 1 < nvalid ? acc[nt][2 * hh + 1] * scale : ninf;
        float* p = o + (long long)row * pitch + col;
        if (col + 1 < T) *reinterpret_cast<float2*>(p) = make_float2(v0, v1);  // pitch and col are even
        else p[0] = v0;
      }
    }
    __syncthreads();  // every warp is done with sb[kt & 1] before the next iteration's prefetch overwrites it
  }
}

// ------------------------------------------------------------------ losses (one warp per query row)
// mode 0: mse over the keys j < min(vs, vt) (the others are inf / nan in the reference and zeroed, train.py:336-340)
// mode 1: sum_j p_j (log p_j - log q_j), p = softmax(t over j < vt), q = softmax(s over j < vs), over the keys both keep
//         (a student-masked key with teacher mass is +inf in the reference and zeroed, :348).  Gradient wrt s_k, k < vs:
//         q_k * P - p_k with P = the teacher mass on the kept keys.
// ds: fp16 [rows][pitch], every column written (zero beyond the kept keys and in the row padding)
__global__ void __launch_bounds__(256) attn_map_loss_kernel(const float* __restrict__ s, const float* __restrict__ t, long long pitch,
                                                            const int* __restrict__ valid_s, const int* __restrict__ valid_t,
                                                            __half* __restrict__ ds, float* __restrict__ loss, long long rows,
                                                            int T, int H, int mode, float loss_mult, float grad_mult) {
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + warp;
  float local = 0.f;
  if (row < rows) {
    const int bi = (int)(row / ((long long)T * H));
    const int vs = valid_s ? min(valid_s[bi], T) : T;
    const int vt = valid_t ? min(valid_t[bi], T) : T;
    const int vk = min(vs, vt);
    const float* sr = s + row * pitch;
    const float* tr = t + row * pitch;
    __half* dr = ds + row * pitch;
    if (mode == 0) {
      for (int j = lane; j < (int)pitch; j += 32) {
        float g = 0.f;
        if (j < vk) {
          const float diff = sr[j] - tr[j];
          local += diff * diff;
          g = 2.f * diff * grad_mult;
        }
        dr[j] = __float2half_rn(g);
      }
    } else {
      float ms = -INFINITY, mt = -INFINITY;
      for (int j = lane; j < vs; j += 32) ms = fmaxf(ms, sr[j]);
      for (int j = lane; j < vt; j += 32) mt = fmaxf(mt, tr[j]);
      ms = warp_max(ms);
      mt = warp_max(mt);
      float zs = 0.f, zt = 0.f, ek = 0.f, et = 0.f;
      for (int j = lane; j < max(vs, vt); j += 32) {
        if (j < vs) zs += expf(sr[j] - ms);
        if (j < vt) {
          const float e = expf(tr[j] - mt);
          zt += e;
          if (j < vk) {
            ek += e;
            et += e * (tr[j] - sr[j]);
          }
        }
      }
      zs = warp_sum(zs);
      zt = warp_sum(zt);
      ek = warp_sum(ek);
      et = warp_sum(et);
      const float inv_zs = 1.f / zs, inv_zt = 1.f / zt;
      const float pk = ek * inv_zt;
      // sum_j p_j ((t_j - mt - log zt) - (s_j - ms - log zs)) over the kept keys
      if (lane == 0) local = et * inv_zt + pk * (ms + logf(zs) - mt - logf(zt));
      for (int j = lane; j < (int)pitch; j += 32) {
        float g = 0.f;
        if (j < vs) {
          const float q = expf(sr[j] - ms) * inv_zs;
          const float p = j < vt ? expf(tr[j] - mt) * inv_zt : 0.f;
          g = (q * pk - p) * grad_mult;
        }
        dr[j] = __float2half_rn(g);
      }
    }
  }
  __shared__ float part[8];
  local = warp_sum(local);
  if (lane == 0) part[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += part[i];
    if (v != 0.f) atomicAdd(loss, v * loss_mult);
  }
}

// KL mode with the whole row in registers (T <= 32 * NT): one read of s and t, one write of ds.  exp through ex2.approx on
// pre-scaled arguments (2 ulp), the same arithmetic as the streaming kernel above otherwise.
template <int NT>
__global__ void __launch_bounds__(256) attn_map_kl_rows_kernel(const float* __restrict__ s, const float* __restrict__ t, long long pitch,
                                                               const int* __restrict__ valid_s, const int* __restrict__ valid_t,
                                                               __half* __restrict__ ds, float* __restrict__ loss, long long rows,
                                                               int T, int H, float loss_mult, float grad_mult) {
  pdl_sync();
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + warp;
  float local = 0.f;
  if (row < rows) {
    const int bi = (int)(row / ((long long)T * H));
    const int vs = valid_s ? min(valid_s[bi], T) : T;
    const int vt = valid_t ? min(valid_t[bi], T) : T;
    const int vk = min(vs, vt);
    const float* sr = s + row * pitch;
    const float* tr = t + row * pitch;
    float sv[NT], tv[NT];
    float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int j = lane + 32 * i;
      sv[i] = j < vs ? __ldg(sr + j) : -INFINITY;
      tv[i] = j < vt ? __ldg(tr + j) : -INFINITY;
      ms = fmaxf(ms, sv[i]);
      mt = fmaxf(mt, tv[i]);
    }
    ms = warp_max(ms);
    mt = warp_max(mt);
    float zs = 0.f, zt = 0.f, ek = 0.f, et = 0.f;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int j = lane + 32 * i;
      const float d = tv[i] - sv[i];                        // only used where both are finite (j < vk)
      const float es = ex2_approx((sv[i] - ms) * kLog2e);   // exp(-inf) = 0 at masked keys
      const float e = ex2_approx((tv[i] - mt) * kLog2e);
      zs += es;
      zt += e;
      if (j < vk) {
        ek += e;
        et += e * d;
      }
      sv[i] = es;
      tv[i] = e;
    }
    zs = warp_sum(zs);
    zt = warp_sum(zt);
    ek = warp_sum(ek);
    et = warp_sum(et);
    const float inv_zs = 1.f / zs, inv_zt = 1.f / zt;
    const float pk = ek * inv_zt;
    if (lane == 0) local = et * inv_zt + pk * (ms - mt + kLn2 * (log2f(zs) - log2f(zt)));
    __half* dr = ds + row * pitch;
    const float a = inv_zs * pk * grad_mult, b = inv_zt * grad_mult;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int j = lane + 32 * i;
      if (j < (int)pitch) dr[j] = __float2half_rn(j < vs ? sv[i] * a - tv[i] * b : 0.f);  // tv = 0 beyond vt
    }
  }
  __shared__ float part[8];
  local = warp_sum(local);
  if (lane == 0) part[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += part[i];
    if (v != 0.f) atomicAdd(loss, v * loss_mult);
  }
}

// ------------------------------------------------------------------ gradient back into the heads
// acc [64 r][DP] = sum_c A[r][c] * M[c][:], A = G tile (TRANS = 0) or its transpose (TRANS = 1)
template <int DP, bool TRANS>
__global__ void __launch_bounds__(128) attn_scores_bwd_kernel(const __half* __restrict__ g, long long pitch, const __half* __restrict__ m,
                                                              long long ld_m, __half* __restrict__ out, long long ld_out, int T, int H,
                                                              int d, float alpha, int accumulate) {
  __shared__ __align__(16) __half sg[2][kTile * (kTile + 8)];
  __shared__ __align__(16) __half sm[2][kTile * (DP + 8)];
  pdl_sync();
  const int r0 = blockIdx.x * kTile, bh = blockIdx.y;
  const int bi = bh / H, h = bh - bi * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* gb = g + (long long)bh * T * pitch;
  const __half* mb = m + (long long)bi * T * ld_m + h * d;
  float acc[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // contraction tiles stream through two shared buffers: tile c + 1 is in flight (cp.async) while tile c is multiplied.
  // The loss kernel wrote every column of g up to the pitch (zeros in the padding): whole 16-byte chunks are readable.
  auto load = [&](int buf, int c0) {
    if (TRANS) load_tile<kTile>(sg[buf], gb, pitch, c0, T, r0, (int)pitch);  // sg[c][r]
    else load_tile<kTile>(sg[buf], gb, pitch, r0, T, c0, (int)pitch);        // sg[r][c]
    load_tile<DP>(sm[buf], mb, ld_m, c0, T, 0, d);                            // sm[c][:]
    cp_async_commit();
  };
  const int nct = (T + kTile - 1) / kTile;
  load(0, 0);
  for (int ct = 0; ct < nct; ++ct) {
    if (ct + 1 < nct) {
      load((ct + 1) & 1, (ct + 1) * kTile);
      cp_async_wait_group<1>();
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();
    const __half* sgt = sg[ct & 1];
    const __half* smt = sm[ct & 1];
#pragma unroll
    for (int kk = 0; kk < kTile / 16; ++kk) {
      uint32_t af[4];
      if (TRANS) {
        // A[r][c] = sg[c][r]: 8x8 blocks (r 0-7, c 0-7) (r 8-15, c 0-7) (r 0-7, c 8-15) (r 8-15, c 8-15), transposed on load
        const int c = kk * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int r = warp * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(smem_u32(sgt + c * (kTile + 8) + r), af[0], af[1], af[2], af[3]);
      } else {
        ldsm_x4(smem_u32(sgt + (warp * 16 + (lane & 15)) * (kTile + 8) + kk * 16 + (lane >> 4) * 8), af[0], af[1], af[2], af[3]);
      }
#pragma unroll
      for (int np = 0; np < DP / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        // .trans matrices: (k 0-7, n 0-7) (k 8-15, n 0-7) (k 0-7, n 8-15) (k 8-15, n 8-15)
        const int row = kk * 16 + (lane & 15);
        const int col = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(smt + row * (DP + 8) + col), b0, b1, b2, b3);
        mma16816(acc[2 * np], af, b0, b1);
        mma16816(acc[2 * np + 1], af, b2, b3);
      }
    }
    __syncthreads();  // both buffers of tile ct are free before the prefetch two iterations ahead overwrites them
  }
  __half* ob = out + (long long)bi * T * ld_out + h * d;
#pragma unroll
  for (int nt = 0; nt < DP / 8; ++nt) {
    const int col = nt * 8 + 2 * (lane & 3);
    if (col >= d) continue;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = r0 + warp * 16 + (lane >> 2) + hh * 8;
      if (row >= T) continue;
      __half2* p = reinterpret_cast<__half2*>(ob + (long long)row * ld_out + col);
      float v0 = acc[nt][2 * hh] * alpha, v1 = acc[nt][2 * hh + 1] * alpha;
      if (accumulate) {
        const float2 old = __half22float2(*p);
        v0 += old.x;
        v1 += old.y;
      }
      // saturating conversion, like every other 16-bit store of the library
      v0 = fminf(fmaxf(v0, -65504.f), 65504.f);
      v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
      *p = __floats2half2_rn(v0, v1);
    }
  }
}

#define FHB_DP_DISPATCH(DPV, CALL)                      \
  if (DPV <= 16) { constexpr int DP = 16; CALL; }       \
  else if (DPV <= 32) { constexpr int DP = 32; CALL; }  \
  else if (DPV <= 48) { constexpr int DP = 48; CALL; }  \
  else { constexpr int DP = 64; CALL; }

int check_heads(const char* who, int32_t B, int32_t T, int32_t H, int32_t d) {
  FHB_ARG_CHECK(B > 0 && T > 0 && H > 0, "%s: B, T, H must be positive (got %d, %d, %d)", who, B, T, H);
  FHB_ARG_CHECK(d > 0 && d <= 64 && d % 8 == 0, "%s: head dim %d must be a multiple of 8, at most 64", who, d);
  FHB_ARG_CHECK((long long)B * H <= 65535, "%s: B * H = %lld exceeds the grid limit", who, (long long)B * H);
  return 0;
}

}  // namespace

extern "C" int fhb_attn_scores(const void* a, const void* b, int64_t ld, const int32_t* valid, float* out, int64_t pitch,
                               int32_t B, int32_t T, int32_t H, int32_t d, float scale, fhb_stream_t stream) {
  int rc = check_heads("attn_scores", B, T, H, d);
  if (rc) return rc;
  FHB_ARG_CHECK(a && b && out, "attn_scores: null pointer");
  FHB_ARG_CHECK(ld % 8 == 0 && ld >= (int64_t)H * d, "attn_scores: row stride %lld must be a multiple of 8 covering H * d", (long long)ld);
  FHB_ARG_CHECK(pitch >= T && pitch % 8 == 0, "attn_scores: pitch %lld must be a multiple of 8, at least T", (long long)pitch);
  FHB_ARG_CHECK(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)out & 15) == 0, "attn_scores: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((T + kTile - 1) / kTile, B * H);
  FHB_DP_DISPATCH(d, FHB_CUDA_CHECK(fhb_launch(attn_scores_kernel<DP>, grid, dim3(128), 0, s, static_cast<const __half*>(a),
                                               static_cast<const __half*>(b), (long long)ld, valid, out, (long long)pitch, T, H, d, scale)));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_attn_map_loss(const float* s, const float* t, int64_t pitch, const int32_t* valid_s, const int32_t* valid_t,
                                 void* ds, float* loss, int32_t B, int32_t T, int32_t H, int32_t mode, float loss_mult,
                                 float grad_mult, fhb_stream_t stream) {
  FHB_ARG_CHECK(s && t && ds && loss, "attn_map_loss: null pointer");
  FHB_ARG_CHECK(B > 0 && T > 0 && H > 0, "attn_map_loss: B, T, H must be positive");
  FHB_ARG_CHECK(mode == 0 || mode == 1, "attn_map_loss: mode must be 0 (mse) or 1 (kldiv)");
  FHB_ARG_CHECK(pitch >= T && pitch % 8 == 0, "attn_map_loss: pitch %lld must be a multiple of 8, at least T", (long long)pitch);
  const long long rows = (long long)B * H * T;
  FHB_ARG_CHECK((rows + 7) / 8 < (1LL << 31), "attn_map_loss: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 1 && pitch <= 32 * 32) {
    // rows of up to 1024 keys (every BASELINE configuration short of cfg-5's 30 s utterances) stay in registers
    const dim3 grid((unsigned)((rows + 7) / 8));
    if (pitch <= 32 * 8) {
      FHB_CUDA_CHECK(fhb_launch(attn_map_kl_rows_kernel<8>, grid, dim3(256), 0, st, s, t, (long long)pitch, valid_s, valid_t,
                                static_cast<__half*>(ds), loss, rows, T, H, loss_mult, grad_mult));
    } else if (pitch <= 32 * 16) {
      FHB_CUDA_CHECK(fhb_launch(attn_map_kl_rows_kernel<16>, grid, dim3(256), 0, st, s, t, (long long)pitch, valid_s, valid_t,
                                static_cast<__half*>(ds), loss, rows, T, H, loss_mult, grad_mult));
    } else {
      FHB_CUDA_CHECK(fhb_launch(attn_map_kl_rows_kernel<32>, grid, dim3(256), 0, st, s, t, (long long)pitch, valid_s, valid_t,
                                static_cast<__half*>(ds), loss, rows, T, H, loss_mult, grad_mult));
    }
    FHB_LAUNCH_CHECK();
    return 0;
  }
  FHB_CUDA_CHECK(fhb_launch(attn_map_loss_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, s, t, (long long)pitch, valid_s,
                            valid_t, static_cast<__half*>(ds), loss, rows, T, H, mode, loss_mult, grad_mult));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_attn_scores_bwd(const void* g, int64_t pitch, const void* m, int64_t ld_m, void* out, int64_t ld_out,
                                   int32_t B, int32_t T, int32_t H, int32_t d, float alpha, int32_t trans, int32_t accumulate,
                                   fhb_stream_t stream) {
  int rc = check_heads("attn_scores_bwd", B, T, H, d);
  if (rc) return rc;
  FHB_ARG_CHECK(g && m && out, "attn_scores_bwd: null pointer");
  FHB_ARG_CHECK(pitch >= T && pitch % 8 == 0, "attn_scores_bwd: pitch %lld must be a multiple of 8, at least T", (long long)pitch);
  FHB_ARG_CHECK(ld_m % 8 == 0 && ld_m >= (int64_t)H * d && ld_out % 2 == 0 && ld_out >= (int64_t)H * d,
                "attn_scores_bwd: row strides must cover H * d (operand: multiple of 8, output: even)");
  FHB_ARG_CHECK(((uintptr_t)g & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)out & 3) == 0, "attn_scores_bwd: misaligned pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((T + kTile - 1) / kTile, B * H);
  if (trans) {
    FHB_DP_DISPATCH(d, FHB_CUDA_CHECK(fhb_launch((attn_scores_bwd_kernel<DP, true>), grid, dim3(128), 0, s, static_cast<const __half*>(g),
                                                 (long long)pitch, static_cast<const __half*>(m), (long long)ld_m,
                                                 static_cast<__half*>(out), (long long)ld_out, T, H, d, alpha, accumulate)));
  } else {
    FHB_DP_DISPATCH(d, FHB_CUDA_CHECK(fhb_launch((attn_scores_bwd_kernel<DP, false>), grid, dim3(128), 0, s, static_cast<const __half*>(g),
                                                 (long long)pitch, static_cast<const __half*>(m), (long long)ld_m,
                                                 static_cast<__half*>(out), (long long)ld_out, T, H, d, alpha, accumulate)));
  }
  FHB_LAUNCH_CHECK();
  return 0;
}
