// K5 layout/epilogue kernels around the grouped positional convolution, which itself runs as a
// batched fhb_gemm (one batch per (sample, group), K = 128 taps x cp channels, overlapped-row TMA view).
// Reference: modules/module.py:186-200 (weight_norm(dim=2) grouped Conv1d, SamePad, GELU),
// :273-281 (index_put(pad -> 0), x + pos_conv(x), LayerNorm).
#include "fhb_common.cuh"

namespace {

// xg[b][g][tp][c'] = x[b][tp - pad_l][g*cg + c'] if 0 <= tp - pad_l < valid[b] and c' < cg else 0
__global__ void __launch_bounds__(256)
posconv_pack_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ valid, __nv_bfloat16* __restrict__ xg,
                    int T, int C, int G, int cp, int pad_l, int Tp, long long total) {
  pdl_sync();
  const int cg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % cp;
    long long r = i / cp;
    const int tp = r % Tp;
    r /= Tp;
    const int g = r % G;
    const int b = r / G;
    const int t = tp - pad_l;
    const int nv = valid ? min(valid[b], T) : T;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (c < cg && t >= 0 && t < nv) v = x[((long long)b * T + t) * C + g * cg + c];
    xg[i] = v;
  }
}

// One block per tap j: n_j = ||v[:,:,j]||, w = g_j * v / n_j written in GEMM layout (bf16).
//  flip_transpose = 0: W[g][co][(j, ci)]           (forward B operand, K-major over (j,ci))
//  flip_transpose = 1: W[g][ci][(127 - j, co)]     (dgrad B operand)
// Time blocking (delta > 1): one GEMM row produces `delta` consecutive frames, so the B operand holds delta
// shifted copies of W:  W'[g][(dl, n)][(j + dl, k)] = W[g][n][(j, k)],  K' = (K + delta) taps, zero elsewhere
// (the buffer is zeroed once; the nonzero pattern never changes).  N grows from cp (30-48: a sliver of the
// 128 x N UMMA) to delta * cp at +delta/K extra flops.
__global__ void __launch_bounds__(256)
posconv_wn_prep_kernel(const float* __restrict__ v, const float* __restrict__ gain, __nv_bfloat16* __restrict__ w_out,
                       float* __restrict__ inv_norm, int C, int G, int K, int cp, int flip_transpose, int delta) {
  pdl_sync();
  const int j = blockIdx.x;
  const int cg = C / G;
  const int n = C * cg;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float t = v[(long long)i * K + j];
    s += t * t;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float inv = rsqrtf(tot);
  if (threadIdx.x == 0 && inv_norm) inv_norm[j] = inv;
  const float sc = gain[j] * inv;
  // padded entries are zero
  const int total = G * cp * cp;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int ci = i % cp;
    const int co = (i / cp) % cp;
    const int g = i / (cp * cp);
    float val = 0.f;
    if (ci < cg && co < cg) val = sc * v[((long long)(g * cg + co) * cg + ci) * K + j];
    const int n = flip_transpose ? ci : co, kk = flip_transpose ? co : ci, jj = flip_transpose ? K - 1 - j : j;
    const __nv_bfloat16 bv = __float2bfloat16(val);
    for (int dl = 0; dl < delta; ++dl)
      w_out[((long long)((g * delta + dl) * cp + n) * (K + (delta > 1 ? delta : 0)) + jj + dl) * cp + kk] = bv;
  }
}

// Warp per (b, t) row:  h = xz + gelu(conv + bias);  y = LN(h).   conv is [B*T][G][cp] (bf16, raw GEMM output)
__global__ void __launch_bounds__(256)
posconv_finish_fwd_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ valid,
                          const __nv_bfloat16* __restrict__ conv, const float* __restrict__ bias,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          __nv_bfloat16* __restrict__ h_out, __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, int B, int T, int C, int G, int cp, float eps, int delta) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * T) return;
  const int b = row / T, t = row % T;
  const int cg = C / G;
  const bool live = valid ? t < valid[b] : true;
  // conv GEMM output layout: [b][r = t / delta][g][dl = t % delta][cp]
  const long long crow = ((long long)b * ((T + delta - 1) / delta) + t / delta) * G * delta + (t % delta);
  float hv[24];  // C <= 768
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = lane + 32 * i;
    hv[i] = 0.f;
    if (c < C) {
      const int g = c / cg, cc = c - g * cg;
      const float xv = live ? __bfloat162float(x[row * C + c]) : 0.f;
      const float cv = __bfloat162float(conv[(crow + (long long)g * delta) * cp + cc]) + bias[c];
      hv[i] = xv + gelu_erf(cv);
      s += hv[i];
    }
  }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      const float dlt = hv[i] - mu;
      q += dlt * dlt;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0 && mean_out) {
    mean_out[row] = mu;
    rstd_out[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      if (h_out) h_out[row * C + c] = __float2bfloat16(hv[i]);
      y[row * C + c] = __float2bfloat16((hv[i] - mu) * rs * gamma[c] + beta[c]);
    }
  }
}

// Backward of the above.  dh = LNbwd(dy); dconv = dh * gelu'(conv + bias) written group-major, time-padded:
// dcg[b][g][t + pad_l][cc] (pad channels written as 0); dgamma/dbeta/dbias accumulated atomically.
__global__ void __launch_bounds__(256)
posconv_finish_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ h,
                          const __nv_bfloat16* __restrict__ conv, const float* __restrict__ bias,
                          const float* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dh,
                          __nv_bfloat16* __restrict__ dcg, float* __restrict__ dgamma, float* __restrict__ dbeta,
                          float* __restrict__ dbias, int B, int T, int C, int G, int cp, int pad_l, int Tp,
                          int rows_per_warp, int delta) {
  pdl_sync();
  extern __shared__ float sred[];  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int cg = C / G;
  const long long rows = (long long)B * T;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float pg[24], pb[24], pc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) pg[i] = pb[i] = pc[i] = 0.f;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const long long row = warp_global * rows_per_warp + rr;
    if (row >= rows) break;
    const int b = row / T, t = row % T;
    const long long crow = ((long long)b * ((T + delta - 1) / delta) + t / delta) * G * delta + (t % delta);
    const float mu = mean[row], rs = rstd[row];
    float xh[24], dv[24];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      const int c = lane + 32 * i;
      xh[i] = dv[i] = 0.f;
      if (c < C) {
        xh[i] = (__bfloat162float(h[row * C + c]) - mu) * rs;
        dv[i] = __bfloat162float(dy[row * C + c]);
        const float dxh = dv[i] * gamma[c];
        s1 += dxh;
        s2 += dxh * xh[i];
        pg[i] += dv[i] * xh[i];
        pb[i] += dv[i];
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float d = rs * (dv[i] * gamma[c] - s1 - xh[i] * s2);
        dh[row * C + c] = __float2bfloat16(d);
        const int g = c / cg, cc = c - g * cg;
        const float cv = __bfloat162float(conv[(crow + (long long)g * delta) * cp + cc]) + bias[c];
        const float dc = d * gelu_erf_grad(cv);
        pc[i] += dc;
        dcg[(((long long)b * G + g) * Tp + t + pad_l) * cp + cc] = __float2bfloat16(dc);
      }
    }
    // channel padding cg..cp-1 of every group: must read as zero in the dgrad GEMM (0-weight x garbage = NaN)
    for (int i = lane; i < G * (cp - cg); i += 32) {
      const int g = i / (cp - cg), cc = cg + i % (cp - cg);
      dcg[(((long long)b * G + g) * Tp + t + pad_l) * cp + cc] = __float2bfloat16(0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      atomicAdd(&sred[c], pg[i]);
      atomicAdd(&sred[C + c], pb[i]);
      atomicAdd(&sred[2 * C + c], pc[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, sred[i]);
    atomicAdd(dbeta + i, sred[C + i]);
    atomicAdd(dbias + i, sred[2 * C + i]);
  }
}

// dx[b][t][c] = (t < valid[b]) ? dh[b][t][c] + dxc[b][t][g][cc] : 0      (dxc = dgrad GEMM output)
__global__ void __launch_bounds__(256)
posconv_unpack_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ dxc,
                          const int* __restrict__ valid, __nv_bfloat16* __restrict__ dx, int T, int C, int G, int cp,
                          long long total, int delta) {
  pdl_sync();
  const int cg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long row = i / C;
    const int t = row % T;
    const int b = row / T;
    const bool live = valid ? t < valid[b] : true;
    float v = 0.f;
    if (live) {
      const int g = c / cg, cc = c - g * cg;
      const long long crow = (((long long)b * ((T + delta - 1) / delta) + t / delta) * G + g) * delta + (t % delta);
      v = __bfloat162float(dh[i]) + __bfloat162float(dxc[crow * cp + cc]);
    }
    dx[i] = __float2bfloat16(v);
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

__global__ void __launch_bounds__(256)
posconv_wn_bwd_kernel(const float* __restrict__ dwt, const float* __restrict__ v, const float* __restrict__ gain,
                      const float* __restrict__ inv_norm, float* __restrict__ dv, float* __restrict__ dg, int C, int G,
                      int K, int cp, int accumulate, int delta) {
  pdl_sync();
  const int j = blockIdx.x;
  const int cg = C / G;
  const int n = C * cg;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i % cg, co_full = i / cg;
    const int g = co_full / cg, co = co_full - g * cg;
    const float dw = wn_dw(dwt, g, j, ci, co, K, cp, delta);
    s += dw * v[(long long)i * K + j];
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float inv = inv_norm[j], gj = gain[j];
  if (threadIdx.x == 0) dg[j] = (accumulate ? dg[j] : 0.f) + tot * inv;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i % cg, co_full = i / cg;
    const int g = co_full / cg, co = co_full - g * cg;
    const float dw = wn_dw(dwt, g, j, ci, co, K, cp, delta);
    const long long vi = (long long)i * K + j;
    const float val = gj * inv * (dw - v[vi] * tot * inv * inv);
    dv[vi] = (accumulate ? dv[vi] : 0.f) + val;
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
  const long long total = (long long)B * G * Tp * cp;
  FHB_CUDA_CHECK(fhb_launch(posconv_pack_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), valid, static_cast<__nv_bfloat16*>(xg), T, C, G, cp, pad_l, Tp, total));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_posconv_wn_prep(const float* v, const float* g, void* w_out, float* inv_norm, int32_t C, int32_t G,
                                   int32_t K, int32_t cp, int32_t flip_transpose, int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(v && g && w_out, "posconv_wn_prep: null pointer");
  FHB_ARG_CHECK(C % G == 0 && cp >= C / G && delta >= 1, "posconv_wn_prep: bad geometry");
  FHB_CUDA_CHECK(fhb_launch(posconv_wn_prep_kernel, dim3(K), dim3(256), 0, static_cast<cudaStream_t>(stream), v, g, static_cast<__nv_bfloat16*>(w_out),
                                                                         inv_norm, C, G, K, cp, flip_transpose, delta));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_posconv_finish_fwd(const void* x, const int32_t* valid, const void* conv, const float* bias,
                                      const float* gamma, const float* beta, void* h_out, void* y, float* mean,
                                      float* rstd, int32_t B, int32_t T, int32_t C, int32_t G, int32_t cp, float eps,
                                      int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(x && conv && bias && gamma && beta && y && delta >= 1, "posconv_finish_fwd: null pointer");
  FHB_ARG_CHECK(C <= 768 && C % G == 0, "posconv_finish_fwd: C=%d must be <= 768", C);
  const long long rows = (long long)B * T;
  FHB_CUDA_CHECK(fhb_launch(posconv_finish_fwd_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), valid, static_cast<const __nv_bfloat16*>(conv), bias, gamma, beta,
      static_cast<__nv_bfloat16*>(h_out), static_cast<__nv_bfloat16*>(y), mean, rstd, B, T, C, G, cp, eps, delta));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_posconv_finish_bwd(const void* dy, const void* h, const void* conv, const float* bias,
                                      const float* gamma, const float* mean, const float* rstd, void* dh, void* dcg,
                                      float* dgamma, float* dbeta, float* dbias, int32_t B, int32_t T, int32_t C,
                                      int32_t G, int32_t cp, int32_t pad_l, int32_t Tp, int32_t delta,
                                      fhb_stream_t stream) {
  FHB_ARG_CHECK(dy && h && conv && bias && gamma && mean && rstd && dh && dcg && dgamma && dbeta && dbias && delta >= 1,
                "posconv_finish_bwd: null pointer");
  FHB_ARG_CHECK(C <= 768 && C % G == 0, "posconv_finish_bwd: C=%d must be <= 768", C);
  const long long rows = (long long)B * T;
  const int rpw = 4;
  const long long warps = (rows + rpw - 1) / rpw;
  FHB_CUDA_CHECK(fhb_launch(posconv_finish_bwd_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 3 * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(conv),
      bias, gamma, mean, rstd, static_cast<__nv_bfloat16*>(dh), static_cast<__nv_bfloat16*>(dcg), dgamma, dbeta, dbias, B,
      T, C, G, cp, pad_l, Tp, rpw, delta));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_posconv_unpack_bwd(const void* dh, const void* dxc, const int32_t* valid, void* dx, int32_t B,
                                      int32_t T, int32_t C, int32_t G, int32_t cp, int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(dh && dxc && dx && delta >= 1, "posconv_unpack_bwd: null pointer");
  const long long total = (long long)B * T * C;
  FHB_CUDA_CHECK(fhb_launch(posconv_unpack_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(dh), static_cast<const __nv_bfloat16*>(dxc), valid,
      static_cast<__nv_bfloat16*>(dx), T, C, G, cp, total, delta));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_posconv_wn_bwd(const float* dwt, const float* v, const float* g, const float* inv_norm, float* dv,
                                  float* dg, int32_t C, int32_t G, int32_t K, int32_t cp, int32_t accumulate,
                                  int32_t delta, fhb_stream_t stream) {
  FHB_ARG_CHECK(dwt && v && g && inv_norm && dv && dg && delta >= 1, "posconv_wn_bwd: null pointer");
  FHB_CUDA_CHECK(fhb_launch(posconv_wn_bwd_kernel, dim3(K), dim3(256), 0, static_cast<cudaStream_t>(stream), dwt, v, g, inv_norm, dv, dg, C, G, K, cp,
                                                                        accumulate, delta));
  FHB_LAUNCH_CHECK();
  return 0;
}
