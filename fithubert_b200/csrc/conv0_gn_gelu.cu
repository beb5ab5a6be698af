// K1: waveform conv (1 -> C channels, k=10, s=5, no bias) + GroupNorm(C groups) + GELU, fused.
// Reference: modules/module.py:46,65-71,99-100 layer 0 (Conv1d -> Fp32GroupNorm(dim,dim) -> GELU).
//
// HBM plan.  GroupNorm needs per-(sample, channel) mean/var of the conv output over ALL T0 frames
// (zero padding included).  Because the conv has one input channel, those statistics are
// a function of 65 numbers per sample: S_j = sum_t x[5t+j] and R_jj' = sum_t x[5t+j] x[5t+j']:
//   mean_c = w_c . S / T0,   E[y^2]_c = w_c^T R w_c / T0.
// So pass 1 reads only the waveform (1 MB/sample), never the C x T0 conv output; pass 2 reads the
// waveform again and writes the bf16 channel-last output once.  Backward needs dW, dgamma, dbeta only
// (the waveform has no gradient) and is ONE pass over dY using the same algebra (see bwd kernel).
#include <stdlib.h>

#include "fhb_common.cuh"

namespace {

constexpr int kK = 10, kS = 5;
constexpr int kNR = 55;            // upper triangle of the 10x10 autocorrelation
constexpr int kNStat = kK + kNR;   // 65

// ---------------------------------------------------------------- pass 1: S_j and R_jj'
__global__ void __launch_bounds__(256) conv0_stats_kernel(const float* __restrict__ wave, long long ld, int T0,
                                                          int frames_per_block, double* __restrict__ stat) {
  pdl_sync();
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(t_begin + frames_per_block, T0);
  const float* x = wave + (long long)b * ld;
  float acc[kNStat];
#pragma unroll
  for (int i = 0; i < kNStat; ++i) acc[i] = 0.f;
  for (int t = t_begin + threadIdx.x; t < t_end; t += blockDim.x) {
    float v[kK];
#pragma unroll
    for (int j = 0; j < kK; ++j) v[j] = __ldg(x + (long long)t * kS + j);
    int r = kK;
#pragma unroll
    for (int j = 0; j < kK; ++j) {
      acc[j] += v[j];
#pragma unroll
      for (int q = j; q < kK; ++q) acc[r++] += v[j] * v[q];
    }
  }
  __shared__ float red[8][kNStat];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kNStat; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < kNStat) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[w][threadIdx.x];
    atomicAdd(stat + (long long)b * kNStat + threadIdx.x, s);
  }
}

__device__ __forceinline__ double quad_form(const float* w, const double* R) {
  // w^T R w with R stored as the packed upper triangle
  double q = 0.0;
  int r = 0;
  for (int j = 0; j < kK; ++j)
    for (int k2 = j; k2 < kK; ++k2) {
      const double term = (double)w[j] * (double)w[k2] * R[r++];
      q += (j == k2) ? term : 2.0 * term;
    }
  return q;
}

__global__ void conv0_finalize_stats_kernel(const double* __restrict__ stat, const float* __restrict__ weight, int C,
                                            int T0, float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_sync();
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* S = stat + (long long)b * kNStat;
  float w[kK];
  double m = 0.0;
  for (int j = 0; j < kK; ++j) {
    w[j] = weight[c * kK + j];
    m += (double)w[j] * S[j];
  }
  m /= (double)T0;
  const double ey2 = quad_form(w, S + kK) / (double)T0;
  const double var = fmax(ey2 - m * m, 0.0);
  mean[b * C + c] = (float)m;
  rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---------------------------------------------------------------- pass 2: normalise + GELU + store
template <int CPT>  // channels per thread (8 -> one 16-byte bf16 store)
__global__ void __launch_bounds__(256)
conv0_fwd_kernel(const float* __restrict__ wave, long long ld, int T0, int C, int frames_per_block,
                 const float* __restrict__ weight, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ mean, const float* __restrict__ rstd, __nv_bfloat16* __restrict__ out,
                 __nv_bfloat16* __restrict__ gp_out) {
  pdl_sync();
  extern __shared__ float xs[];  // frames_per_block*5 + 5 samples
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * frames_per_block;
  const int nframes = min(frames_per_block, T0 - t_begin);
  const float* x = wave + (long long)b * ld + (long long)t_begin * kS;
  const int nsamp = nframes * kS + (kK - kS);
  for (int i = threadIdx.x; i < nsamp; i += blockDim.x) xs[i] = __ldg(x + i);
  const int tcols = C / CPT;
  const int tx = threadIdx.x % tcols, ty = threadIdx.x / tcols, fy = blockDim.x / tcols;
  const int c0 = tx * CPT;
  float w[CPT][kK], sc[CPT], sh[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int j = 0; j < kK; ++j) w[i][j] = __ldg(weight + (c0 + i) * kK + j);
    const float r = rstd[b * C + c0 + i], g = __ldg(gamma + c0 + i);
    sc[i] = r * g;
    sh[i] = __ldg(beta + c0 + i) - mean[b * C + c0 + i] * r * g;
  }
  __syncthreads();
  if (ty >= fy) return;
  __nv_bfloat16* o = out + ((long long)b * T0 + t_begin) * C + c0;
  for (int t = ty; t < nframes; t += fy) {
    float v[kK];
#pragma unroll
    for (int j = 0; j < kK; ++j) v[j] = xs[t * kS + j];
    float y[CPT], gp[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < kK; ++j) a = fmaf(w[i][j], v[j], a);
      if (gp_out) gelu_erf_both(fmaf(a, sc[i], sh[i]), y[i], gp[i]);
      else y[i] = gelu_erf(fmaf(a, sc[i], sh[i]));
    }
    uint32_t pk[CPT / 2];
#pragma unroll
    for (int i = 0; i < CPT / 2; ++i) pk[i] = pack_f16(y[2 * i], y[2 * i + 1]);
    if (CPT == 8) {
      *reinterpret_cast<uint4*>(o + (long long)t * C) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    } else {
      *reinterpret_cast<uint2*>(o + (long long)t * C) = make_uint2(pk[0], pk[1]);
    }
    if (gp_out) {  // gelu'(z): the backward multiplier (applied by the dgrad epilogue of the next layer)
#pragma unroll
      for (int i = 0; i < CPT / 2; ++i) pk[i] = pack_f16(gp[2 * i], gp[2 * i + 1]);
      __nv_bfloat16* og = gp_out + ((long long)b * T0 + t_begin) * C + c0 + (long long)t * C;
      if (CPT == 8) *reinterpret_cast<uint4*>(og) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      else *reinterpret_cast<uint2*>(og) = make_uint2(pk[0], pk[1]);
    }
  }
}

// ---------------------------------------------------------------- pass 2 on the tensor pipe (C = 64 * 2^k <= 512)
// The direct kernel above spends 10 of its ~23 instructions per output on the convolution FMAs and is FP32-issue
// bound at 2 TB/s (ncu r01n_ncu_full_conv0_fwd_t).  Here the k = 10 taps become the contraction of a
// mma.sync.m16n8k8 TF32 product (16 frames x 8 channels per instruction, K = 10 padded to 16), with
//   * the waveform split into two TF32 terms (x = hi + lo, both multiplied in): exact in the samples; the scaled
//     weights are rounded to TF32 once (2^-11 relative - the fp16 the reference's own AMP conv uses has the same
//     mantissa, and the bf16 rounding of the output is 4x coarser),
//   * the GroupNorm scale folded into the B operand and the shift used as the accumulator's initial value,
//   * channels permuted inside each group of 4 n-tiles so that a thread ends up with 8 CONSECUTIVE channels of
//     a frame: one 16-byte store per (thread, frame), 64 contiguous bytes per quad.
// What is left per output is the GELU (8 instructions), half a pack and 1/8 of a store.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_tf32_k4(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
// first product of a chain: D = A B + C with C in its own registers (no accumulator-initialising MOVs)
__device__ __forceinline__ void mma_tf32_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %11, %10, %11};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
}

template <bool GP>
__global__ void __launch_bounds__(256, 2)
conv0_fwd_mma_kernel(const float* __restrict__ wave, long long ld, int T0, int C, int frames_per_block,
                     const float* __restrict__ weight, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ mean, const float* __restrict__ rstd, __nv_bfloat16* __restrict__ out,
                     __nv_bfloat16* __restrict__ gp_out) {
  pdl_sync();
  extern __shared__ float xs[];  // frames_per_block * 5 + 16 samples (tail zero-filled: the padded taps read it)
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * frames_per_block;
  const int nframes = min(frames_per_block, T0 - t_begin);
  const float* x = wave + (long long)b * ld + (long long)t_begin * kS;
  const int nsamp = nframes * kS + (kK - kS);
  const int nsm = frames_per_block * kS + 16;
  for (int i0 = threadIdx.x; i0 < nsm; i0 += 4 * blockDim.x) {  // 4 loads in flight per thread
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < nsamp ? __ldg(x + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < nsm) xs[i] = v[u];
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ngroups = C >> 6;            // 64-channel groups: 1, 2, 4 or 8
  const int wpg = 8 / ngroups;           // warps sharing a channel group (they split the frames)
  const int cbase = (warp % ngroups) * 64;
  const int flane = warp / ngroups;
  // B operand: n-tile j (0..7), column n = g  <->  channel cbase + 32 (j >> 2) + 8 (g >> 1) + 2 (j & 3) + (g & 1);
  // k-step 0 (m16n8k8) holds taps t and t + 4, k-step 1 (m16n8k4) tap t + 8 (< 10 for t < 2, else padding)
  uint32_t bh[8][3];
  float sh[8][2];  // accumulator start = GroupNorm shift of this thread's two output columns (2t, 2t + 1) per n-tile
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cbase + 32 * (j >> 2) + 8 * (g >> 1) + 2 * (j & 3) + (g & 1);
    const float sc = rstd[b * C + ch] * __ldg(gamma + ch);
    const float w0 = __ldg(weight + ch * kK + t) * sc, w1 = __ldg(weight + ch * kK + t + 4) * sc;
    const float w2 = t < 2 ? __ldg(weight + ch * kK + t + 8) * sc : 0.f;
    bh[j][0] = to_tf32(w0);
    bh[j][1] = to_tf32(w1);
    bh[j][2] = to_tf32(w2);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int co = cbase + 32 * (j >> 2) + 8 * t + 2 * (j & 3) + e;  // output column 2t + e of n-tile j
      sh[j][e] = __ldg(beta + co) - mean[b * C + co] * rstd[b * C + co] * __ldg(gamma + co);
    }
  }
  __syncthreads();
  // this thread's output pointer for frame f0 + g; rows g + 8 sit `row8` elements further, the next m-tile `step`
  const long long row8 = 8LL * C, step = (long long)wpg * 16 * C;
  __nv_bfloat16* op = out + ((long long)b * T0 + t_begin + flane * 16 + g) * C + cbase + 8 * t;
  __nv_bfloat16* gpp = GP ? gp_out + ((long long)b * T0 + t_begin + flane * 16 + g) * C + cbase + 8 * t : nullptr;
  for (int f0 = flane * 16; f0 < nframes; f0 += wpg * 16, op += step, gpp += GP ? step : 0) {
    // A fragments (frames f0 + g and f0 + g + 8): k-step 0 = taps (t, t + 4), k-step 1 = tap t + 8
    const float* x0 = xs + (f0 + g) * kS;
    const float* x1 = x0 + 8 * kS;
    const float av[6] = {x0[t], x1[t], x0[t + 4], x1[t + 4], t < 2 ? x0[t + 8] : 0.f, t < 2 ? x1[t + 8] : 0.f};
    // x = hi + lo: hi keeps the 10 mantissa bits the tensor pipe reads (the low 13 bits are ignored by the MMA),
    // lo = x - hi is exact in fp32 and itself truncated to TF32 by the hardware (error 2^-21 |x|)
    uint32_t ah0[4], al0[4], ah1[2], al1[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ah0[i] = __float_as_uint(av[i]) & 0xFFFFE000u;
      al0[i] = __float_as_uint(av[i] - __uint_as_float(ah0[i]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ah1[i] = __float_as_uint(av[4 + i]) & 0xFFFFE000u;
      al1[i] = __float_as_uint(av[4 + i] - __uint_as_float(ah1[i]));
    }
    const bool r0 = f0 + g < nframes, r1 = f0 + g + 8 < nframes;
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
      float y[2][8], gp[2][8];  // [row g / g + 8][8 consecutive channels]
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = grp * 4 + jj;
        float d[4];
        mma_tf32_c(d, al0, bh[j][0], bh[j][1], sh[j][0], sh[j][1]);
        mma_tf32_k4(d, al1[0], al1[1], bh[j][2]);
        mma_tf32(d, ah0, bh[j][0], bh[j][1]);
        mma_tf32_k4(d, ah1[0], ah1[1], bh[j][2]);
#pragma unroll
        for (int row = 0; row < 2; ++row) {  // (d[0], d[1]) = row g, (d[2], d[3]) = row g + 8: adjacent channels
          if (GP) gelu_erf_both2(d[2 * row], d[2 * row + 1], y[row][2 * jj], y[row][2 * jj + 1], gp[row][2 * jj], gp[row][2 * jj + 1]);
          else gelu_erf2(d[2 * row], d[2 * row + 1], y[row][2 * jj], y[row][2 * jj + 1]);
        }
      }
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        if (row == 0 ? r0 : r1) {
          const long long off = row * row8 + 32 * grp;
          *reinterpret_cast<uint4*>(op + off) = make_uint4(pack_f16(y[row][0], y[row][1]), pack_f16(y[row][2], y[row][3]),
                                                              pack_f16(y[row][4], y[row][5]), pack_f16(y[row][6], y[row][7]));
          if (GP)
            *reinterpret_cast<uint4*>(gpp + off) = make_uint4(pack_f16(gp[row][0], gp[row][1]), pack_f16(gp[row][2], gp[row][3]),
                                                                pack_f16(gp[row][4], gp[row][5]), pack_f16(gp[row][6], gp[row][7]));
        }
      }
    }
  }
}

// Fold the per-thread accumulators of the fy frame lanes (same channels, different frames) into red[(i, q)][tx]
// (bank-conflict free, zero-initialised by the caller) one lane at a time: plain read-modify-writes between
// barriers - shared-memory float atomics are CAS spin loops and the lanes collide on every address.
template <int CPT>
__device__ __forceinline__ void block_fold(float* red, const float (&acc)[CPT][2 + kK], int tx, int ty, int fy, int tcols) {
  __syncthreads();
  for (int w = 0; w < fy; ++w) {
    if (ty == w) {
#pragma unroll
      for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int q = 0; q < 2 + kK; ++q) red[(i * (2 + kK) + q) * tcols + tx] += acc[i][q];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- backward (single pass over dY)
// Per (b, c) accumulate  A0 = sum_t dz,  A1 = sum_t dz*xhat,  P_j = sum_t dz*x[5t+j]   (dz = dY*gelu'(z)).
constexpr int kNAcc = 2 + kK;  // 12
template <int CPT>             // 4
__global__ void __launch_bounds__(256)
conv0_bwd_kernel(const float* __restrict__ wave, long long ld, int T0, int C, int frames_per_block,
                 const float* __restrict__ weight, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ mean, const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dy,
                 float* __restrict__ acc_out /*[B][C][12]*/, int dy_is_dz) {
  pdl_sync();
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * frames_per_block;
  const int nframes = min(frames_per_block, T0 - t_begin);
  const int nsamp = nframes * kS + (kK - kS);
  float* xs = smem;
  float* red = smem + frames_per_block * kS + (kK - kS);  // [C][12]
  const float* x = wave + (long long)b * ld + (long long)t_begin * kS;
  for (int i = threadIdx.x; i < nsamp; i += blockDim.x) xs[i] = __ldg(x + i);
  for (int i = threadIdx.x; i < C * kNAcc; i += blockDim.x) red[i] = 0.f;
  const int tcols = C / CPT;
  const int tx = threadIdx.x % tcols, ty = threadIdx.x / tcols, fy = blockDim.x / tcols;
  const int c0 = tx * CPT;
  float w[CPT][kK], mu[CPT], rs[CPT], g[CPT], be[CPT], acc[CPT][kNAcc];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int j = 0; j < kK; ++j) w[i][j] = __ldg(weight + (c0 + i) * kK + j);
    mu[i] = mean[b * C + c0 + i];
    rs[i] = rstd[b * C + c0 + i];
    g[i] = __ldg(gamma + c0 + i);
    be[i] = __ldg(beta + c0 + i);
#pragma unroll
    for (int q = 0; q < kNAcc; ++q) acc[i][q] = 0.f;
  }
  __syncthreads();
  if (ty < fy) {
    const __nv_bfloat16* d = dy + ((long long)b * T0 + t_begin) * C + c0;
    for (int t = ty; t < nframes; t += fy) {
      float v[kK];
#pragma unroll
      for (int j = 0; j < kK; ++j) v[j] = xs[t * kS + j];
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(d + (long long)t * C));
      const float2 d01 = unpack_f16(raw.x), d23 = unpack_f16(raw.y);
      const float dv[4] = {d01.x, d01.y, d23.x, d23.y};
      if (dy_is_dz) {
        // dy already carries gelu'(z) (saved by the forward, multiplied in by the producing dgrad epilogue):
        // only A0 and P_j are accumulated; A1 = rstd * (w . P - mean * A0) follows algebraically below
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          acc[i][0] += dv[i];
#pragma unroll
          for (int j = 0; j < kK; ++j) acc[i][2 + j] = fmaf(dv[i], v[j], acc[i][2 + j]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          float a = 0.f;
#pragma unroll
          for (int j = 0; j < kK; ++j) a = fmaf(w[i][j], v[j], a);
          const float xhat = (a - mu[i]) * rs[i];
          const float dz = dv[i] * gelu_erf_grad(fmaf(xhat, g[i], be[i]));
          acc[i][0] += dz;
          acc[i][1] += dz * xhat;
#pragma unroll
          for (int j = 0; j < kK; ++j) acc[i][2 + j] = fmaf(dz, v[j], acc[i][2 + j]);
        }
      }
    }
    if (dy_is_dz) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        float wp = 0.f;
#pragma unroll
        for (int j = 0; j < kK; ++j) wp = fmaf(w[i][j], acc[i][2 + j], wp);
        acc[i][1] = rs[i] * (wp - mu[i] * acc[i][0]);
      }
    }
  }
  block_fold<CPT>(red, acc, tx, ty, fy, tcols);
  float* o = acc_out + (long long)b * C * kNAcc;
  for (int i = threadIdx.x; i < C * kNAcc; i += blockDim.x) {
    const int c = i / kNAcc, q = i - c * kNAcc;
    atomicAdd(o + i, red[((c % CPT) * kNAcc + q) * tcols + c / CPT]);
  }
}

// Training path: dy already carries gelu'(z) (saved by the forward as gp_out and multiplied in by the dgrad
// epilogue that produced dy), so per (b, c) only A0 = sum_t dz and P_j = sum_t dz x[5t+j] are streamed
// (11 FMAs per element, HBM-bound read of dz); A1 = rstd (w . P - mean A0) follows algebraically.
// 4 frames are loaded ahead of the FMAs so that every thread keeps 4 x 8-byte loads in flight.
template <int CPT>  // 4
__global__ void __launch_bounds__(256)
conv0_bwd_dz_kernel(const float* __restrict__ wave, long long ld, int T0, int C, int frames_per_block,
                    const float* __restrict__ weight, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const __nv_bfloat16* __restrict__ dz, float* __restrict__ acc_out /*[B][C][12]*/) {
  pdl_sync();
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * frames_per_block;
  const int nframes = min(frames_per_block, T0 - t_begin);
  const int nsamp = nframes * kS + (kK - kS);
  float* xs = smem;
  float* red = smem + frames_per_block * kS + (kK - kS);  // [C][12]
  const float* x = wave + (long long)b * ld + (long long)t_begin * kS;
  for (int i = threadIdx.x; i < nsamp; i += blockDim.x) xs[i] = __ldg(x + i);
  for (int i = threadIdx.x; i < C * kNAcc; i += blockDim.x) red[i] = 0.f;
  const int tcols = C / CPT;
  const int tx = threadIdx.x % tcols, ty = threadIdx.x / tcols, fy = blockDim.x / tcols;
  const int c0 = tx * CPT;
  float acc[CPT][kNAcc];
#pragma unroll
  for (int i = 0; i < CPT; ++i)
#pragma unroll
    for (int q = 0; q < kNAcc; ++q) acc[i][q] = 0.f;
  __syncthreads();
  if (ty < fy) {
    const __nv_bfloat16* d = dz + ((long long)b * T0 + t_begin) * C + c0;
    constexpr int U = 4;
    for (int t0 = ty; t0 < nframes; t0 += U * fy) {
      uint2 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + u * fy;
        raw[u] = t < nframes ? __ldg(reinterpret_cast<const uint2*>(d + (long long)t * C)) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = min(t0 + u * fy, nframes - 1);  // out-of-range frames carry dz = 0
        float v[kK];
#pragma unroll
        for (int j = 0; j < kK; ++j) v[j] = xs[t * kS + j];
        const float2 d01 = unpack_f16(raw[u].x), d23 = unpack_f16(raw[u].y);
        const float dv[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          acc[i][0] += dv[i];
#pragma unroll
          for (int j = 0; j < kK; ++j) acc[i][2 + j] = fmaf(dv[i], v[j], acc[i][2 + j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      float wp = 0.f;
#pragma unroll
      for (int j = 0; j < kK; ++j) wp = fmaf(__ldg(weight + (c0 + i) * kK + j), acc[i][2 + j], wp);
      const float mu = mean[b * C + c0 + i], rs = rstd[b * C + c0 + i];
      acc[i][1] = rs * (wp - mu * acc[i][0]);
    }
  }
  block_fold<CPT>(red, acc, tx, ty, fy, tcols);
  float* o = acc_out + (long long)b * C * kNAcc;
  for (int i = threadIdx.x; i < C * kNAcc; i += blockDim.x) {
    const int c = i / kNAcc, q = i - c * kNAcc;
    atomicAdd(o + i, red[((c % CPT) * kNAcc + q) * tcols + c / CPT]);
  }
}

// dW[c][j], dgamma[c], dbeta[c] from the per-(b,c) accumulators (see header comment for the algebra).
// One thread per (channel, tap); the tap-0 thread also owns dgamma / dbeta.
__global__ void conv0_bwd_finalize_kernel(const float* __restrict__ acc, const double* __restrict__ stat,
                                          const float* __restrict__ weight, const float* __restrict__ gamma,
                                          const float* __restrict__ mean, const float* __restrict__ rstd, int B, int C,
                                          int T0, float* __restrict__ dW, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta, int accumulate) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * kK) return;
  const int c = idx / kK, j = idx - c * kK;
  double w[kK];
  int ridx[kK];  // packed-upper-triangle index of R[q][j]
#pragma unroll
  for (int q = 0; q < kK; ++q) {
    w[q] = (double)weight[c * kK + q];
    const int lo = q < j ? q : j, hi = q < j ? j : q;
    ridx[q] = lo * kK - lo * (lo - 1) / 2 + (hi - lo);
  }
  const double gm = gamma[c];
  double dw = 0.0, dg = 0.0, db = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* a = acc + ((long long)b * C + c) * kNAcc;
    const double* S = stat + (long long)b * kNStat;
    const double* R = S + kK;
    const double r = rstd[b * C + c], m = mean[b * C + c];
    const double a0 = a[0], a1 = a[1];
    dg += a1;
    db += a0;
    double yx = 0.0;  // sum_t y x_j = sum_q w_q R[q, j]
#pragma unroll
    for (int q = 0; q < kK; ++q) yx += w[q] * R[ridx[q]];
    const double xhat_x = r * (yx - m * S[j]);
    dw += r * gm * ((double)a[2 + j] - a0 / T0 * S[j] - a1 / T0 * xhat_x);
  }
  dW[idx] = (accumulate ? dW[idx] : 0.f) + (float)dw;
  if (j == 0) {
    dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)dg;
    dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)db;
  }
}

// ---------------------------------------------------------------- backward on the tensor pipe
// P_j[b][c] = sum_t dz[b][t][c] x[b][5t + j] and A0[b][c] = sum_t dz[b][t][c] are one contraction over the frames:
// [C x T0] . [T0 x 32] per sample, i.e. a wgrad-shaped fhb_gemm (both operands MN-major, split-K) against this
// im2col of the waveform:  xcol[b][t][0..9] = bf16(x[5t + j]),  [10] = 1,  [16..25] = bf16(x - float(hi)) (the
// low half of a two-term bf16 split: together 16 mantissa bits), everything else 0.
__global__ void __launch_bounds__(256)
conv0_im2col_kernel(const float* __restrict__ wave, long long ld, int T0, __nv_bfloat16* __restrict__ xcol) {
  pdl_sync();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T0) return;
  const float* x = wave + (long long)b * ld + (long long)t * kS;
  float hi[kK], lo[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    const float v = __ldg(x + j);
    hi[j] = __half2float(__float2half_rn(v));
    lo[j] = v - hi[j];
  }
  uint4* o = reinterpret_cast<uint4*>(xcol + ((long long)b * T0 + t) * 32);
  o[0] = make_uint4(pack_f16(hi[0], hi[1]), pack_f16(hi[2], hi[3]), pack_f16(hi[4], hi[5]), pack_f16(hi[6], hi[7]));
  o[1] = make_uint4(pack_f16(hi[8], hi[9]), pack_f16(1.0f, 0.f), 0u, 0u);
  o[2] = make_uint4(pack_f16(lo[0], lo[1]), pack_f16(lo[2], lo[3]), pack_f16(lo[4], lo[5]), pack_f16(lo[6], lo[7]));
  o[3] = make_uint4(pack_f16(lo[8], lo[9]), 0u, 0u, 0u);
}

// dW[c][j], dgamma[c], dbeta[c] from the GEMM accumulators acc32[b][c][32] (layout of xcol's columns); same algebra
// as conv0_bwd_finalize_kernel with A1 = rstd (w . P - mean A0) formed here.
__global__ void conv0_bwd_finalize32_kernel(const float* __restrict__ acc32, const double* __restrict__ stat,
                                            const float* __restrict__ weight, const float* __restrict__ gamma,
                                            const float* __restrict__ mean, const float* __restrict__ rstd, int B, int C,
                                            int T0, float* __restrict__ dW, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int accumulate) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * kK) return;
  const int c = idx / kK, j = idx - c * kK;
  double w[kK];
  int ridx[kK];
#pragma unroll
  for (int q = 0; q < kK; ++q) {
    w[q] = (double)weight[c * kK + q];
    const int lo = q < j ? q : j, hi = q < j ? j : q;
    ridx[q] = lo * kK - lo * (lo - 1) / 2 + (hi - lo);
  }
  const double gm = gamma[c];
  double dw = 0.0, dg = 0.0, db = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* a = acc32 + ((long long)b * C + c) * 32;
    const double* S = stat + (long long)b * kNStat;
    const double* R = S + kK;
    const double r = rstd[b * C + c], m = mean[b * C + c];
    const double a0 = a[10];
    double wp = 0.0, yx = 0.0;
#pragma unroll
    for (int q = 0; q < kK; ++q) {
      wp += w[q] * ((double)a[q] + (double)a[16 + q]);
      yx += w[q] * R[ridx[q]];
    }
    const double a1 = r * (wp - m * a0);
    dg += a1;
    db += a0;
    const double xhat_x = r * (yx - m * S[j]);
    dw += r * gm * ((double)a[j] + (double)a[16 + j] - a0 / T0 * S[j] - a1 / T0 * xhat_x);
  }
  dW[idx] = (accumulate ? dW[idx] : 0.f) + (float)dw;
  if (j == 0) {
    dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)dg;
    dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)db;
  }
}

int check_common(const fhb_conv0_args* a) {
  FHB_ARG_CHECK(a != nullptr, "conv0: null args");
  FHB_ARG_CHECK(a->kernel == kK && a->stride == kS, "conv0: only k=10,s=5 is implemented (got k=%d,s=%d)", a->kernel,
                a->stride);
  FHB_ARG_CHECK(a->B > 0 && a->C > 0 && a->T0 > 0, "conv0: empty problem");
  FHB_ARG_CHECK((long long)(a->T0 - 1) * kS + kK <= a->L, "conv0: T0=%d frames do not fit in L=%d samples", a->T0, a->L);
  FHB_ARG_CHECK(a->C % 8 == 0 && 256 % (a->C / 8) == 0 && a->C <= 2048, "conv0: C=%d must be 8*2^k, <= 2048", a->C);
  FHB_ARG_CHECK(a->wave && a->weight && a->gamma && a->beta && a->stat && a->mean && a->rstd, "conv0: null pointer");
  return 0;
}

}  // namespace

extern "C" int fhb_conv0_gn_gelu_fwd(const fhb_conv0_args* a, fhb_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;
  FHB_ARG_CHECK(a->out != nullptr, "conv0 fwd: null out");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FHB_CUDA_CHECK(cudaMemsetAsync(a->stat, 0, sizeof(double) * kNStat * a->B, s));
  {
    const int fpb = 2048;
    dim3 grid((a->T0 + fpb - 1) / fpb, a->B);
    FHB_CUDA_CHECK(fhb_launch(conv0_stats_kernel, dim3(grid), dim3(256), 0, s, a->wave, a->wave_ld, a->T0, fpb, a->stat));
    FHB_LAUNCH_CHECK();
  }
  {
    dim3 grid((a->C + 127) / 128, a->B);
    FHB_CUDA_CHECK(fhb_launch(conv0_finalize_stats_kernel, dim3(grid), dim3(128), 0, s, a->stat, a->weight, a->C, a->T0, a->eps, a->mean, a->rstd));
    FHB_LAUNCH_CHECK();
  }
  {
    static const bool direct_only = getenv("FHB_CONV0_DIRECT") != nullptr;
    if (!direct_only && a->C % 64 == 0 && a->C <= 512 && 8 % (a->C / 64) == 0) {
      // tensor-pipe path: 1024 frames per block (the 64 B-operand registers of a thread are set up once per block)
      const int frames = 1024;
      dim3 grid((a->T0 + frames - 1) / frames, a->B);
      const size_t smem = sizeof(float) * (frames * kS + 16);
      if (a->gp_out)
        FHB_CUDA_CHECK(fhb_launch(conv0_fwd_mma_kernel<true>, grid, dim3(256), smem, s, a->wave, a->wave_ld, a->T0, a->C, frames,
                                  a->weight, a->gamma, a->beta, a->mean, a->rstd, static_cast<__nv_bfloat16*>(a->out),
                                  static_cast<__nv_bfloat16*>(a->gp_out)));
      else
        FHB_CUDA_CHECK(fhb_launch(conv0_fwd_mma_kernel<false>, grid, dim3(256), smem, s, a->wave, a->wave_ld, a->T0, a->C, frames,
                                  a->weight, a->gamma, a->beta, a->mean, a->rstd, static_cast<__nv_bfloat16*>(a->out),
                                  static_cast<__nv_bfloat16*>(a->gp_out)));
      return 0;
    }
    // 512 frames per block: the 80 weight registers / affine terms of a thread are amortised over
    // 512 / (256 / (C/8)) frames, and the 10 KB waveform slice is staged once
    const int frames = 512;
    dim3 grid((a->T0 + frames - 1) / frames, a->B);
    const size_t smem = sizeof(float) * (frames * kS + (kK - kS));
    FHB_CUDA_CHECK(fhb_launch((conv0_fwd_kernel<8>), dim3(grid), dim3(256), smem, s, a->wave, a->wave_ld, a->T0, a->C, frames, a->weight, a->gamma, a->beta,
                                                 a->mean, a->rstd, static_cast<__nv_bfloat16*>(a->out),
                                                 static_cast<__nv_bfloat16*>(a->gp_out)));
    FHB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int fhb_conv0_gn_gelu_bwd(const fhb_conv0_args* a, fhb_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;
  FHB_ARG_CHECK(a->dy && a->acc && a->dweight && a->dgamma && a->dbeta, "conv0 bwd: null pointer");
  FHB_ARG_CHECK(a->C % 4 == 0 && 256 % (a->C / 4) == 0, "conv0 bwd: C=%d must be 4*2^k, <= 1024", a->C);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FHB_CUDA_CHECK(cudaMemsetAsync(a->acc, 0, sizeof(float) * kNAcc * a->B * a->C, s));
  if (a->dy_is_dz) {
    const int frames = 1024;
    dim3 grid((a->T0 + frames - 1) / frames, a->B);
    const size_t smem = sizeof(float) * (frames * kS + (kK - kS) + a->C * kNAcc);
    FHB_CUDA_CHECK(fhb_launch((conv0_bwd_dz_kernel<4>), dim3(grid), dim3(256), smem, s, a->wave, a->wave_ld, a->T0, a->C, frames, a->weight, a->mean, a->rstd,
                                                   static_cast<const __nv_bfloat16*>(a->dy), a->acc));
    FHB_LAUNCH_CHECK();
    FHB_CUDA_CHECK(fhb_launch(conv0_bwd_finalize_kernel, dim3((a->C * kK + 127) / 128), dim3(128), 0, s, a->acc, a->stat, a->weight, a->gamma, a->mean, a->rstd,
                                                             a->B, a->C, a->T0, a->dweight, a->dgamma, a->dbeta,
                                                             a->accumulate));
    FHB_LAUNCH_CHECK();
    return 0;
  }
  // 2048 frames per block: few blocks per (sample, channel) -> few global atomics per accumulator
  const int frames = 2048;
  dim3 grid((a->T0 + frames - 1) / frames, a->B);
  const size_t smem = sizeof(float) * (frames * kS + (kK - kS) + a->C * kNAcc);
  FHB_CUDA_CHECK(fhb_launch((conv0_bwd_kernel<4>), dim3(grid), dim3(256), smem, s, a->wave, a->wave_ld, a->T0, a->C, frames, a->weight, a->gamma, a->beta,
                                              a->mean, a->rstd, static_cast<const __nv_bfloat16*>(a->dy), a->acc,
                                              a->dy_is_dz));
  FHB_LAUNCH_CHECK();
  FHB_CUDA_CHECK(fhb_launch(conv0_bwd_finalize_kernel, dim3((a->C * kK + 127) / 128), dim3(128), 0, s, a->acc, a->stat, a->weight, a->gamma, a->mean, a->rstd,
                                                           a->B, a->C, a->T0, a->dweight, a->dgamma, a->dbeta,
                                                           a->accumulate));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_conv0_im2col(const float* wave, int64_t wave_ld, int32_t B, int32_t L, int32_t T0, void* xcol,
                                fhb_stream_t stream) {
  FHB_ARG_CHECK(wave && xcol && B > 0 && T0 > 0, "conv0_im2col: bad arguments");
  FHB_ARG_CHECK((long long)(T0 - 1) * kS + kK <= L, "conv0_im2col: T0=%d frames do not fit in L=%d samples", T0, L);
  FHB_CUDA_CHECK(fhb_launch(conv0_im2col_kernel, dim3((T0 + 255) / 256, B), dim3(256), 0, static_cast<cudaStream_t>(stream),
                            wave, (long long)wave_ld, T0, static_cast<__nv_bfloat16*>(xcol)));
  return 0;
}

extern "C" int fhb_conv0_bwd_finalize(const float* acc32, const fhb_conv0_args* a, fhb_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;
  FHB_ARG_CHECK(acc32 && a->dweight && a->dgamma && a->dbeta, "conv0_bwd_finalize: null pointer");
  FHB_CUDA_CHECK(fhb_launch(conv0_bwd_finalize32_kernel, dim3((a->C * kK + 127) / 128), dim3(128), 0,
                            static_cast<cudaStream_t>(stream), acc32, a->stat, a->weight, a->gamma, a->mean, a->rstd, a->B,
                            a->C, a->T0, a->dweight, a->dgamma, a->dbeta, a->accumulate));
  return 0;
}
