// K5 layout/epilogue kernels around the grouped positional convolution, which itself runs as a
// batched fhb_gemm (one batch per (sample, group), K = 128 taps x cp channels, overlapped-row TMA view).
// Reference: modules/module.py:186-200 (weight_norm(dim=2) grouped Conv1d, SamePad, GELU),
// :273-281 (index_put(pad -> 0), x + pos_conv(x), LayerNorm).
#include "fhb_common.cuh"

namespace {

// bf16 vectors of VEC = 2 (one 32-bit word) or 8 (one 16-byte word) elements <-> fp32 registers
template <int VEC>
__device__ __forceinline__ void ldv(const __nv_bfloat16* p, float* f) {
  if constexpr (VEC == 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t a[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_f16(a[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
  } else {
    const float2 t = unpack_f16(*reinterpret_cast<const uint32_t*>(p));
    f[0] = t.x;
    f[1] = t.y;
  }
}
template <int VEC>
__device__ __forceinline__ void stv(__nv_bfloat16* p, const float* f) {
  if constexpr (VEC == 8) {
    *reinterpret_cast<uint4*>(p) =
        make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
  } else {
    *reinterpret_cast<uint32_t*>(p) = pack_f16(f[0], f[1]);
  }
}
// fp16 twins (forward activations are fp16, gradients bf16: fhb_common.cuh)
template <int VEC>
__device__ __forceinline__ void stvh(__nv_bfloat16* p, const float* f) {
  if constexpr (VEC == 8) {
    *reinterpret_cast<uint4*>(p) =
        make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
  } else {
    *reinterpret_cast<uint32_t*>(p) = pack_f16(f[0], f[1]);
  }
}
template <int VEC>
__device__ __forceinline__ void ldf(const float* p, float* f) {
#pragma unroll
  for (int j = 0; j < VEC; j += 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p + j));
    f[j] = t.x;
    f[j + 1] = t.y;
  }
}

// raw (unconverted) vector loads: issuing every load of a row into its own registers before the first use keeps
// them all in flight (the compiler otherwise recycles one destination register and serialises the row on L2 latency)
template <int VEC>
struct RawVec {
  uint32_t w[VEC / 2];
};
template <int VEC>
__device__ __forceinline__ RawVec<VEC> ldraw(const __nv_bfloat16* p) {
  RawVec<VEC> r;
  if constexpr (VEC == 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    r.w[0] = u.x; r.w[1] = u.y; r.w[2] = u.z; r.w[3] = u.w;
  } else {
    r.w[0] = *reinterpret_cast<const uint32_t*>(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void cvtraw(const RawVec<VEC>& r, float* f) {
#pragma unroll
  for (int j = 0; j < VEC / 2; ++j) {
    const float2 t = unpack_f16(r.w[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

template <int VEC>
__device__ __forceinline__ void cvtrawh(const RawVec<VEC>& r, float* f) {
#pragma unroll
  for (int j = 0; j < VEC / 2; ++j) {
    const float2 t = unpack_f16(r.w[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

// xg[b][g][tp][c'] = x[b][tp - pad_l][g*cg + c'] if 0 <= tp - pad_l < valid[b] and c' < cg else 0.
// One thread per 16-byte output chunk (8 channels); the source is read as one 16-byte word (cg % 8 == 0, VEC = 8)
// or as 32-bit words (cg even, VEC = 2: FitHuBERT's 30 channels per group start on 4-byte boundaries only).
template <int VEC>
__global__ void __launch_bounds__(256)
posconv_pack_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ valid, __nv_bfloat16* __restrict__ xg,
                    int T, int C, int G, int cp, int pad_l, int Tp, long long total_chunks) {
  pdl_sync();
  const int cg = C / G, cpv = cp >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cpv) * 8;
    long long r = i / cpv;
    const int tp = r % Tp;
    r /= Tp;
    const int g = r % G;
    const int b = r / G;
    const int t = tp - pad_l;
    const int nv = valid ? min(valid[b], T) : T;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < nv && c0 < cg) {
      const __nv_bfloat16* src = x + ((long long)b * T + t) * C + g * cg + c0;
      if constexpr (VEC == 8) {
        o = __ldg(reinterpret_cast<const uint4*>(src));
      } else if constexpr (VEC == 2) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (c0 + 2 * q < cg) ? __ldg(reinterpret_cast<const uint32_t*>(src + 2 * q)) : 0u;
        o = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t lo = (c0 + 2 * q < cg) ? (uint32_t)reinterpret_cast<const uint16_t*>(src)[2 * q] : 0u;
          const uint32_t hi = (c0 + 2 * q + 1 < cg) ? (uint32_t)reinterpret_cast<const uint16_t*>(src)[2 * q + 1] : 0u;
          w[q] = lo | (hi << 16);
        }
        o = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    reinterpret_cast<uint4*>(xg)[i] = o;
  }
}

// Weight norm (reference modules/module.py:199: nn.utils.weight_norm(dim=2)) in two launches.
//  (1) per-tap sums of squares of v [C*cg][K] (tap fastest): coalesced row reads, one atomic per tap per block
__global__ void __launch_bounds__(256)
posconv_wn_sumsq_kernel(const float* __restrict__ v, long long n_rows, int K, int rows_per_block, float* __restrict__ sumsq) {
  pdl_sync();
  __shared__ float red[256];
  const int lanes = blockDim.x / K;            // row lanes (host guarantees K <= 256 and K divides 256)
  const int j = threadIdx.x % K, rl = threadIdx.x / K;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(n_rows, r0 + rows_per_block);
  float s = 0.f;
#pragma unroll 8
  for (long long r = r0 + rl; r < r1; r += lanes) {
    const float t = __ldg(v + r * K + j);
    s = fmaf(t, t, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (rl == 0) {
    for (int q = 1; q < lanes; ++q) s += red[q * K + j];
    atomicAdd(sumsq + j, s);
  }
}

//  (2) w = g_j * v / ||v_j|| written in GEMM layout (bf16).  Block (r, g, z): z = 0 forward operand
//      W[g][co = r][(j, ci)], z = 1 dgrad operand W[g][ci = r][(K - 1 - j, co)]; the [kk][tap] source tile
//      (kk = the operand's inner channel index) is read with coalesced 4*K-byte rows and transposed through smem.
//  Time blocking (delta > 1): one GEMM row produces `delta` consecutive frames, so the B operand holds delta
//  shifted copies of W:  W'[g][(dl, n)][(j + dl, k)] = W[g][n][(j, k)],  K' = (K + delta) taps, zero elsewhere
//  (the buffer is zeroed once; the nonzero pattern never changes).  N grows from cp (30-48: a sliver of the
//  128 x N UMMA) to delta * cp at +delta/K extra flops.
__global__ void __launch_bounds__(256)
posconv_wn_layout_kernel(const float* __restrict__ v, const float* __restrict__ gain, const float* __restrict__ sumsq,
                         __nv_bfloat16* __restrict__ w_fwd, __nv_bfloat16* __restrict__ w_bwd,
                         float* __restrict__ inv_norm, int C, int G, int K, int cp, int delta) {
  pdl_sync();
  extern __shared__ float tile[];  // [cg][K + 1] then sc[K]
  const int cg = C / G;
  const int r = blockIdx.x, g = blockIdx.y, flip = blockIdx.z;
  __nv_bfloat16* w_out = flip ? w_bwd : w_fwd;
  float* sc = tile + cg * (K + 1);
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    const float inv = rsqrtf(sumsq[j]);
    sc[j] = gain[j] * inv;
    if (r == 0 && g == 0 && flip == 0 && inv_norm) inv_norm[j] = inv;
  }
  if (w_out == nullptr) return;
  for (int i = threadIdx.x; i < cg * K; i += blockDim.x) {
    const int kk = i / K, j = i - kk * K;
    const long long row = flip ? (long long)(g * cg + kk) * cg + r : (long long)(g * cg + r) * cg + kk;
    tile[kk * (K + 1) + j] = __ldg(v + row * K + j);
  }
  __syncthreads();
  const int Kx = K + (delta > 1 ? delta : 0);
  const int half = cg >> 1;  // cg is even (host check): consecutive threads write consecutive bf16 pairs of one tap row
  for (int i = threadIdx.x; i < K * half; i += blockDim.x) {
    const int j = i / half, kk = (i - j * half) * 2;
    const uint32_t pk = pack_f16(sc[j] * tile[kk * (K + 1) + j], sc[j] * tile[(kk + 1) * (K + 1) + j]);
    const int jj = flip ? K - 1 - j : j;
    for (int dl = 0; dl < delta; ++dl)
      *reinterpret_cast<uint32_t*>(w_out + ((long long)((g * delta + dl) * cp + r) * Kx + jj + dl) * cp + kk) = pk;
  }
}

// Warp per (b, t) row:  h = xz + gelu(conv + bias);  y = LN(h).   conv is [B*T][G][cp] (bf16, raw GEMM output).
// Lane owns NV vectors of VEC channels (a vector never straddles a group: cg % VEC == 0); the group offsets are
// computed once per lane, each warp walks `rows_per_warp` rows.
template <int VEC, int NV>
__global__ void __launch_bounds__(256)
posconv_finish_fwd_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ valid,
                          const __nv_bfloat16* __restrict__ conv, const float* __restrict__ bias,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          __nv_bfloat16* __restrict__ h_out, __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out, int B, int T, int C, int G, int cp,
                          float eps, int delta, int rows_per_warp) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long rows = (long long)B * T;
  const int cg = C / G;
  const int R = (T + delta - 1) / delta;
  int coff[NV];  // offset of this lane's i-th vector inside a conv row block [g][dl][cp]; -1: beyond C
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * VEC;
    const int g = c / cg;
    coff[i] = c < C ? g * delta * cp + (c - g * cg) : -1;
  }
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const long long row = warp_global * rows_per_warp + rr;
    if (row >= rows) break;
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    const bool live = valid ? t < valid[b] : true;
    // conv GEMM output layout: [b][r = t / delta][g][dl = t % delta][cp]
    const __nv_bfloat16* cbase = conv + (((long long)b * R + t / delta) * G * delta + (t % delta)) * cp;
    RawVec<VEC> rx[NV], rc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (coff[i] >= 0) {
        const int c = (lane + 32 * i) * VEC;
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) rx[i].w[j] = 0u;
        if (live) rx[i] = ldraw<VEC>(x + row * C + c);
        rc[i] = ldraw<VEC>(cbase + coff[i]);
      }
    }
    float hv[NV][VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (coff[i] >= 0) {
        const int c = (lane + 32 * i) * VEC;
        float xv[VEC], cv[VEC], bv[VEC];
        cvtrawh<VEC>(rx[i], xv);
        cvtrawh<VEC>(rc[i], cv);
        ldf<VEC>(bias + c, bv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          hv[i][j] = xv[j] + gelu_erf(cv[j] + bv[j]);
          s += hv[i][j];
        }
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (coff[i] >= 0) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float dlt = hv[i][j] - mu;
          q = fmaf(dlt, dlt, q);
        }
      }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0 && mean_out) {
      mean_out[row] = mu;
      rstd_out[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (coff[i] >= 0) {
        const int c = (lane + 32 * i) * VEC;
        if (h_out) stvh<VEC>(h_out + row * C + c, hv[i]);
        float gv[VEC], bv[VEC], o[VEC];
        ldf<VEC>(gamma + c, gv);
        ldf<VEC>(beta + c, bv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = (hv[i][j] - mu) * rs * gv[j] + bv[j];
        stvh<VEC>(y + row * C + c, o);
        if (y32) {  // fp32 copy: the residual operand of the first transformer layer's out_proj epilogue
#pragma unroll
          for (int j = 0; j < VEC; j += 2) *reinterpret_cast<float2*>(y32 + row * C + c + j) = make_float2(o[j], o[j + 1]);
        }
      }
    }
  }
}

// Backward of the above.  dh = LNbwd(dy); dconv = dh * gelu'(conv + bias) written group-major, time-padded:
// dcg[b][g][t + pad_l][cc] (pad channels written as 0); dgamma/dbeta/dbias accumulated atomically.
// Column partial sums live in a per-warp shared-memory slab [3][C] (each lane owns its columns: plain
// read-modify-writes, no atomics), which keeps the register budget at two blocks per SM.
template <int VEC, int NV>
__global__ void __launch_bounds__(256, 2)
posconv_finish_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ h,
                          const __nv_bfloat16* __restrict__ conv, const float* __restrict__ bias,
                          const float* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dh,
                          __nv_bfloat16* __restrict__ dcg, float* __restrict__ dgamma, float* __restrict__ dbeta,
                          float* __restrict__ dbias, int B, int T, int C, int G, int cp, int pad_l, int Tp,
                          int rows_per_warp, int delta) {
  pdl_sync();
  extern __shared__ float sred[];  // [warps][3][C]
  const int lane = threadIdx.x & 31;
  const int cg = C / G;
  const int R = (T + delta - 1) / delta;
  const long long rows = (long long)B * T;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float* mine = sred + (threadIdx.x >> 5) * 3 * C;
  for (int i = lane; i < 3 * C; i += 32) mine[i] = 0.f;
  __syncwarp();
  int gcc[NV];  // (group << 16) | channel-in-group of this lane's i-th vector; -1: beyond C
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * VEC;
    const int g = c / cg;
    gcc[i] = c < C ? ((g << 16) | (c - g * cg)) : -1;
  }
  const int npad = G * (cp - cg) / 2;  // zero pad channel PAIRS per row (cp, cg even)
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const long long row = warp_global * rows_per_warp + rr;
    if (row >= rows) break;
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    const __nv_bfloat16* cbase = conv + (((long long)b * R + t / delta) * G * delta + (t % delta)) * cp;
    __nv_bfloat16* dbase = dcg + ((long long)b * G * Tp + t + pad_l) * cp;
    const float mu = mean[row], rs = rstd[row];
    RawVec<VEC> rh[NV], rd[NV], rc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (gcc[i] >= 0) {
        const int c = (lane + 32 * i) * VEC;
        rh[i] = ldraw<VEC>(h + row * C + c);
        rd[i] = ldraw<VEC>(dy + row * C + c);
        rc[i] = ldraw<VEC>(cbase + (gcc[i] >> 16) * delta * cp + (gcc[i] & 0xFFFF));
      }
    }
    float xh[NV][VEC], dv[NV][VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (gcc[i] >= 0) {
        const int c = (lane + 32 * i) * VEC;
        float gv[VEC];
        cvtrawh<VEC>(rh[i], xh[i]);  // saved forward sum h: fp16
        cvtraw<VEC>(rd[i], dv[i]);
        ldf<VEC>(gamma + c, gv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          mine[c + j] = fmaf(dv[i][j], xh[i][j], mine[c + j]);
          mine[C + c + j] += dv[i][j];
          dv[i][j] *= gv[j];  // dxhat
          s1 += dv[i][j];
          s2 = fmaf(dv[i][j], xh[i][j], s2);
        }
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (gcc[i] >= 0) {
        const int c = (lane + 32 * i) * VEC;
        float cv[VEC], bv[VEC], d[VEC], dc[VEC];
        cvtrawh<VEC>(rc[i], cv);  // forward conv output: fp16
        ldf<VEC>(bias + c, bv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          d[j] = rs * (dv[i][j] - s1 - xh[i][j] * s2);
          dc[j] = d[j] * gelu_erf_grad(cv[j] + bv[j]);
          mine[2 * C + c + j] += dc[j];
        }
        stv<VEC>(dh + row * C + c, d);
        stv<VEC>(dbase + (long long)(gcc[i] >> 16) * Tp * cp + (gcc[i] & 0xFFFF), dc);
      }
    }
    // channel padding cg..cp-1 of every group: must read as zero in the dgrad GEMM (0-weight x garbage = NaN)
    for (int i = lane; i < npad; i += 32) {
      const int g = i / ((cp - cg) / 2), cc = cg + 2 * (i % ((cp - cg) / 2));
      *reinterpret_cast<uint32_t*>(dbase + (long long)g * Tp * cp + cc) = 0u;
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sred[w * 3 * C + i];
    float* dst = i < C ? dgamma + i : (i < 2 * C ? dbeta + (i - C) : dbias + (i - 2 * C));
    atomicAdd(dst, t);
  }
}

// dx[b][t][c] = (t < valid[b]) ? dh[b][t][c] + dxc[b][t][g][cc] : 0      (dxc = dgrad GEMM output)
template <int VEC>
__global__ void __launch_bounds__(256)
posconv_unpack_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ dxc,
                          const int* __restrict__ valid, __nv_bfloat16* __restrict__ dx, int T, int C, int G, int cp,
                          long long total_vec, int delta) {
  pdl_sync();
  const int cg = C / G, cv = C / VEC;
  const int R = (T + delta - 1) / delta;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * VEC;
    const long long row = i / cv;
    const int t = row % T;
    const int b = row / T;
    const bool live = valid ? t < valid[b] : true;
    float o[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    if (live) {
      const int g = c / cg, cc = c - g * cg;
      const long long crow = (((long long)b * R + t / delta) * G + g) * delta + (t % delta);
      float a[VEC], d[VEC];
      ldv<VEC>(dh + i * VEC, a);
      ldv<VEC>(dxc + crow * cp + cc, d);
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = a[j] + d[j];
    }
    stv<VEC>(dx + i * VEC, o);
  }
}

// Weight-norm backward.  dwt is the wgrad GEMM output fp32 [G][(j, ci)][co'] (cp-padded), i.e.
// dW[g*cg+co][ci][j] = dwt[g][j*cp + ci][co].  One block per tap j:
//   s_j = sum dW * v ;  dg_j = s_j * inv_j ;  dv = g_j*inv_j * (dW - v * s_j * inv_j^2)
// Time-blocked wgrad (delta > 1): dwt is [G][(j', ci), j' < K + delta][(dl, co)] and
//   dW[.][ci][j] = sum_dl dwt[g][(j + dl) * cp + ci][dl * cp + co]   (the delta shifted partial products).
__device__ __forceinline__ float wn_dw(const float* __restrict__ dwt, int g, int j, int ci, int co, int K, int cp, int delta) {
  if (delta <= 1) return dwt[((long long)g * K * cp + (long long)j * cp + ci) * cp + co];
  const long long rows = (long long)(K + delta) * cp, ldw = (long long)delta * cp;
  float s = 0.f;
  for (int dl = 0; dl < delta; ++dl) s += dwt[((long long)g * rows + (long long)(j + dl) * cp + ci) * ldw + dl * cp + co];
  return s;
}

// One block per (input channel ci, group g): the dW slab [co < cg][j < K] of that pair goes through shared memory, because
// dwt is contiguous along co and v / dv along the tap j (tap-fastest parameter layout) - both sides are then read and
// written in contiguous runs.  PHASE 0: tot[j] += sum_co dW * v (one atomic per tap and block).  PHASE 1 (second launch,
// tot complete): dv, and dg by block (0, 0).  (r05: the one-block-per-tap version walked v at stride K twice: 119 us.)
template <int PHASE>
__global__ void __launch_bounds__(256)
posconv_wn_bwd_kernel(const float* __restrict__ dwt, const float* __restrict__ v, const float* __restrict__ gain,
                      const float* __restrict__ inv_norm, float* __restrict__ dv, float* __restrict__ dg,
                      float* __restrict__ tot, int C, int G, int K, int cp, int accumulate, int delta) {
  extern __shared__ float wn_tile[];  // [cg][K + 1]
  pdl_sync();
  const int ci = blockIdx.x, g = blockIdx.y;
  const int cg = C / G, ld = K + 1, n = cg * K;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {  // co fastest: contiguous runs of dwt
    const int co = idx % cg, j = idx / cg;
    wn_tile[co * ld + j] = wn_dw(dwt, g, j, ci, co, K, cp, delta);
  }
  __syncthreads();
  if (PHASE == 0) {
    // thread = (tap j, a slice of the co's): v read in runs of K floats  (K <= blockDim.x, checked by the host)
    const int slices = blockDim.x / K;
    const int j = threadIdx.x % K, sl = threadIdx.x / K;
    if (sl < slices) {
      float acc = 0.f;
      for (int co = sl; co < cg; co += slices)
        acc += wn_tile[co * ld + j] * __ldg(v + ((long long)(g * cg + co) * cg + ci) * K + j);
      atomicAdd(tot + j, acc);
    }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {  // j fastest: contiguous runs of v / dv
      const int j = idx % K, co = idx / K;
      const long long vi = ((long long)(g * cg + co) * cg + ci) * K + j;
      const float inv = __ldg(inv_norm + j), t = __ldg(tot + j);
      const float val = __ldg(gain + j) * inv * (wn_tile[co * ld + j] - __ldg(v + vi) * t * inv * inv);
      dv[vi] = (accumulate ? dv[vi] : 0.f) + val;
    }
    if (blockIdx.x == 0 && blockIdx.y == 0)
      for (int j = threadIdx.x; j < K; j += blockDim.x) dg[j] = (accumulate ? dg[j] : 0.f) + __ldg(tot + j) * __ldg(inv_norm + j);
  }
}

int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)fhb_num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int fhb_posconv_pack(const void* x, const int32_t* valid, void* xg, int32_t B, int32_t T, int32_t C,
                                int32_t G, int32_t cp, int32_t pad_l, int32_t Tp, fhb_stream_t stream) {
  FHB_ARG_CHECK(x && xg, "posconv_pack: null pointer");
  FHB_ARG_CHECK(C % G == 0 && cp >= C / G && cp % 16 == 0 && Tp >= T + pad_l, "posconv_pack: bad geometry");
  const long long total = (long long)B * G * Tp * (cp / 8);
  const int cg = C / G;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(xg);
  if (cg % 8 == 0 && C % 8 == 0)
    FHB_CUDA_CHECK(fhb_launch(posconv_pack_kernel<8>, dim3(grid_for(total)), dim3(256), 0, s, xp, valid, op, T, C, G, cp, pad_l, Tp, total));
  else if (cg % 2 == 0)
    FHB_CUDA_CHECK(fhb_launch(posconv_pack_kernel<2>, dim3(grid_for(total)), dim3(256), 0, s, xp, valid, op, T, C, G, cp, pad_l, Tp, total));
  else
    FHB_CUDA_CHECK(fhb_launch(posconv_pack_kernel<1>, dim3(grid_for(total)), dim3(256), 0, s, xp, valid, op, T, C, G, cp, pad_l, Tp, total));
  return 0;
}

extern "C" int fhb_posconv_wn_prep(const float* v, const float* g, void* w_fwd, void* w_bwd, float* ws, int32_t C,
                                   int32_t G, int32_t K, int32_t cp, int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(v && g && ws && (w_fwd || w_bwd), "posconv_wn_prep: null pointer");
  FHB_ARG_CHECK(C % G == 0 && cp >= C / G && delta >= 1 && (C / G) % 2 == 0, "posconv_wn_prep: bad geometry");
  FHB_ARG_CHECK(K > 0 && K <= 256 && 256 % K == 0, "posconv_wn_prep: K=%d must divide 256", K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cg = C / G;
  const long long n_rows = (long long)C * cg;
  FHB_CUDA_CHECK(cudaMemsetAsync(ws, 0, K * sizeof(float), s));
  const int rpb = 64;
  FHB_CUDA_CHECK(fhb_launch(posconv_wn_sumsq_kernel, dim3((unsigned)((n_rows + rpb - 1) / rpb)), dim3(256), 0, s, v, n_rows, K, rpb, ws));
  const size_t smem = ((size_t)cg * (K + 1) + K) * sizeof(float);
  FHB_ARG_CHECK(smem <= 48 * 1024, "posconv_wn_prep: cg=%d x K=%d tile does not fit 48 KB of shared memory", cg, K);
  FHB_CUDA_CHECK(fhb_launch(posconv_wn_layout_kernel, dim3(cg, G, w_bwd ? 2 : 1), dim3(256), smem, s, v, g,
                            static_cast<const float*>(ws), static_cast<__nv_bfloat16*>(w_fwd),
                            static_cast<__nv_bfloat16*>(w_bwd), ws + K, C, G, K, cp, delta));
  return 0;
}

template <int VEC>
int launch_finish_fwd(int nv, dim3 grid, cudaStream_t s, const __nv_bfloat16* x, const int32_t* valid, const __nv_bfloat16* conv,
                      const float* bias, const float* gamma, const float* beta, __nv_bfloat16* h_out, __nv_bfloat16* y,
                      float* y32, float* mean, float* rstd, int B, int T, int C, int G, int cp, float eps, int delta, int rpw) {
#define FHB_FF(NV)                                                                                                        \
  FHB_CUDA_CHECK(fhb_launch((posconv_finish_fwd_kernel<VEC, NV>), grid, dim3(256), 0, s, x, valid, conv, bias, gamma, beta, \
                            h_out, y, y32, mean, rstd, B, T, C, G, cp, eps, delta, rpw))
  if constexpr (VEC == 8) {
    if (nv <= 1) FHB_FF(1); else if (nv == 2) FHB_FF(2); else FHB_FF(3);
  } else {
    if (nv <= 4) FHB_FF(4); else if (nv <= 8) FHB_FF(8); else FHB_FF(12);
  }
#undef FHB_FF
  return 0;
}

extern "C" int fhb_posconv_finish_fwd(const void* x, const int32_t* valid, const void* conv, const float* bias,
                                      const float* gamma, const float* beta, void* h_out, void* y, float* y32,
                                      float* mean, float* rstd, int32_t B, int32_t T, int32_t C, int32_t G, int32_t cp,
                                      float eps, int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(x && conv && bias && gamma && beta && y && delta >= 1, "posconv_finish_fwd: null pointer");
  FHB_ARG_CHECK(C <= 768 && C % G == 0 && (C / G) % 2 == 0, "posconv_finish_fwd: C=%d must be <= 768 with an even group width", C);
  const long long rows = (long long)B * T;
  const int rpw = 2;
  const dim3 grid((unsigned)((rows + 8 * rpw - 1) / (8 * rpw)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cg = C / G;
  if (cg % 8 == 0)
    return launch_finish_fwd<8>((C / 8 + 31) / 32, grid, s, static_cast<const __nv_bfloat16*>(x), valid,
                                static_cast<const __nv_bfloat16*>(conv), bias, gamma, beta, static_cast<__nv_bfloat16*>(h_out),
                                static_cast<__nv_bfloat16*>(y), y32, mean, rstd, B, T, C, G, cp, eps, delta, rpw);
  return launch_finish_fwd<2>((C / 2 + 31) / 32, grid, s, static_cast<const __nv_bfloat16*>(x), valid,
                              static_cast<const __nv_bfloat16*>(conv), bias, gamma, beta, static_cast<__nv_bfloat16*>(h_out),
                              static_cast<__nv_bfloat16*>(y), y32, mean, rstd, B, T, C, G, cp, eps, delta, rpw);
}

template <int VEC>
int launch_finish_bwd(int nv, dim3 grid, size_t smem, cudaStream_t s, const __nv_bfloat16* dy, const __nv_bfloat16* h,
                      const __nv_bfloat16* conv, const float* bias, const float* gamma, const float* mean, const float* rstd,
                      __nv_bfloat16* dh, __nv_bfloat16* dcg, float* dgamma, float* dbeta, float* dbias, int B, int T, int C,
                      int G, int cp, int pad_l, int Tp, int rpw, int delta) {
#define FHB_FB(NV)                                                                                                       \
  do {                                                                                                                   \
    FHB_ONCE_PER_DEVICE(FHB_CUDA_CHECK(cudaFuncSetAttribute((posconv_finish_bwd_kernel<VEC, NV>),                        \
        cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * (int)sizeof(float))));                               \
    FHB_CUDA_CHECK(fhb_launch((posconv_finish_bwd_kernel<VEC, NV>), grid, dim3(256), smem, s, dy, h, conv, bias, gamma,   \
                              mean, rstd, dh, dcg, dgamma, dbeta, dbias, B, T, C, G, cp, pad_l, Tp, rpw, delta));         \
  } while (0)
  if constexpr (VEC == 8) {
    if (nv <= 1) FHB_FB(1); else if (nv == 2) FHB_FB(2); else FHB_FB(3);
  } else {
    if (nv <= 4) FHB_FB(4); else if (nv <= 8) FHB_FB(8); else FHB_FB(12);
  }
#undef FHB_FB
  return 0;
}

extern "C" int fhb_posconv_finish_bwd(const void* dy, const void* h, const void* conv, const float* bias,
                                      const float* gamma, const float* mean, const float* rstd, void* dh, void* dcg,
                                      float* dgamma, float* dbeta, float* dbias, int32_t B, int32_t T, int32_t C,
                                      int32_t G, int32_t cp, int32_t pad_l, int32_t Tp, int32_t delta,
                                      fhb_stream_t stream) {
  FHB_ARG_CHECK(dy && h && conv && bias && gamma && mean && rstd && dh && dcg && dgamma && dbeta && dbias && delta >= 1,
                "posconv_finish_bwd: null pointer");
  FHB_ARG_CHECK(C <= 768 && C % G == 0 && (C / G) % 2 == 0 && cp % 2 == 0,
                "posconv_finish_bwd: C=%d must be <= 768 with an even group width", C);
  const long long rows = (long long)B * T;
  // ~4 blocks per SM; every warp walks a contiguous run of rows and carries its column partial sums in registers
  long long blocks = 4LL * fhb_num_sms();
  if (blocks > (rows + 7) / 8) blocks = (rows + 7) / 8;
  const int rpw = (int)((rows + blocks * 8 - 1) / (blocks * 8));
  blocks = (rows + 8LL * rpw - 1) / (8LL * rpw);
  const dim3 grid((unsigned)blocks);
  const size_t smem = 8 * 3 * C * sizeof(float);  // [8 warps][3][C]
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cg = C / G;
  if (cg % 8 == 0)
    return launch_finish_bwd<8>((C / 8 + 31) / 32, grid, smem, s, static_cast<const __nv_bfloat16*>(dy),
                                static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(conv), bias, gamma,
                                mean, rstd, static_cast<__nv_bfloat16*>(dh), static_cast<__nv_bfloat16*>(dcg), dgamma, dbeta,
                                dbias, B, T, C, G, cp, pad_l, Tp, rpw, delta);
  return launch_finish_bwd<2>((C / 2 + 31) / 32, grid, smem, s, static_cast<const __nv_bfloat16*>(dy),
                              static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(conv), bias, gamma,
                              mean, rstd, static_cast<__nv_bfloat16*>(dh), static_cast<__nv_bfloat16*>(dcg), dgamma, dbeta,
                              dbias, B, T, C, G, cp, pad_l, Tp, rpw, delta);
}

extern "C" int fhb_posconv_unpack_bwd(const void* dh, const void* dxc, const int32_t* valid, void* dx, int32_t B,
                                      int32_t T, int32_t C, int32_t G, int32_t cp, int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(dh && dxc && dx && delta >= 1, "posconv_unpack_bwd: null pointer");
  FHB_ARG_CHECK(C % G == 0 && (C / G) % 2 == 0, "posconv_unpack_bwd: the group width must be even");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cg = C / G;
  const __nv_bfloat16 *a = static_cast<const __nv_bfloat16*>(dh), *b = static_cast<const __nv_bfloat16*>(dxc);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(dx);
  if (cg % 8 == 0) {
    const long long total = (long long)B * T * C / 8;
    FHB_CUDA_CHECK(fhb_launch(posconv_unpack_bwd_kernel<8>, dim3(grid_for(total)), dim3(256), 0, s, a, b, valid, o, T, C, G, cp, total, delta));
  } else {
    const long long total = (long long)B * T * C / 2;
    FHB_CUDA_CHECK(fhb_launch(posconv_unpack_bwd_kernel<2>, dim3(grid_for(total)), dim3(256), 0, s, a, b, valid, o, T, C, G, cp, total, delta));
  }
  return 0;
}

extern "C" int fhb_posconv_wn_bwd(const float* dwt, const float* v, const float* g, float* ws, float* dv, float* dg,
                                  int32_t C, int32_t G, int32_t K, int32_t cp, int32_t accumulate, int32_t delta,
                                  fhb_stream_t stream) {
  FHB_ARG_CHECK(dwt && v && g && ws && dv && dg && delta >= 1, "posconv_wn_bwd: null pointer");
  FHB_ARG_CHECK(G > 0 && C % G == 0 && K > 0 && K <= 256, "posconv_wn_bwd: bad geometry (C=%d G=%d K=%d)", C, G, K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cg = C / G;
  const size_t smem = sizeof(float) * (size_t)cg * (K + 1);
  FHB_ARG_CHECK(smem <= 48 * 1024, "posconv_wn_bwd: group width %d x %d taps does not fit 48 KB of shared memory", cg, K);
  // ws is fhb_posconv_wn_prep's workspace: [K, 2K) holds 1 / ||v[:, :, j]||; [0, K) (its sums of squares, dead since
  // the prep) takes the per-tap dot products.  (A stream-ordered cudaMallocAsync scratch cost 8 ms of HOST time per call.)
  float* tot = ws;
  const float* inv_norm = ws + K;
  FHB_CUDA_CHECK(cudaMemsetAsync(tot, 0, sizeof(float) * (size_t)K, s));
  FHB_CUDA_CHECK(fhb_launch(posconv_wn_bwd_kernel<0>, dim3(cg, G), dim3(256), smem, s, dwt, v, g, inv_norm, dv, dg, tot, C, G, K,
                            cp, accumulate, delta));
  FHB_CUDA_CHECK(fhb_launch(posconv_wn_bwd_kernel<1>, dim3(cg, G), dim3(256), smem, s, dwt, v, g, inv_norm, dv, dg, tot, C, G, K,
                            cp, accumulate, delta));
  return 0;
}
