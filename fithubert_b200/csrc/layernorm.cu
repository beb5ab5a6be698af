// K3: warp-per-row LayerNorm forward / backward (bf16 activations, fp32 statistics).
// Reference call sites: modules/model.py:446-447; modules/module.py:251,281,513,518,569,580.
// HBM-bound: one read of x (and dy) and one write per element; rows are 960-1536 bytes, each lane
// moves 16-byte vectors; gamma/beta stay in L1/L2.
#include "fhb_common.cuh"

namespace {

constexpr int kMaxVec = 3;  // C <= 3*32*8 = 768

// forward activations are fp16, gradients bf16 (fhb_common.cuh)
__device__ __forceinline__ void load8h(const __half* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8h(__half* p, const float* f) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
}
__device__ __forceinline__ void cvt8h(const uint4& u, float* f) {
  float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
}

__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8f(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// IN_F32: x is the fp32 pre-LayerNorm sum written by the out_proj / fc2 GEMM epilogue (the residual stream is never
// rounded to bf16 on its way through a layer).  y (bf16) is the next GEMM's A operand; y32 (optional) the fp32 copy the
// next residual add reads.  sub32 / diff_out (optional): diff_out = bf16(x - sub32), i.e. the FFN branch output
// `layer_result` of modules/module.py:577-580 recovered from the sum and its residual operand.
template <bool IN_F32>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const void* __restrict__ x_, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __half* __restrict__ y, float* __restrict__ y32,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ sub32,
                     __half* __restrict__ diff_out, long long rows, int C, float eps) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = C >> 3;
  for (long long row = warp_global; row < rows; row += nwarps) {
    float v[kMaxVec][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        if (IN_F32) load8f(static_cast<const float*>(x_) + row * C + vi * 8, v[i]);
        else load8h(static_cast<const __half*>(x_) + row * C + vi * 8, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mu;
          q += d * d;
        }
      }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0 && mean_out) {
      mean_out[row] = mu;
      rstd_out[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float o[8];
        if (diff_out) {
          float r[8];
          load8f(sub32 + row * C + vi * 8, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = v[i][j] - r[j];
          store8h(diff_out + row * C + vi * 8, o);
        }
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)),
                     g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8)),
                     b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
        store8h(y + row * C + vi * 8, o);
        if (y32) store8f(y32 + row * C + vi * 8, o);
      }
    }
  }
}

// dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)),  dxhat = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; optionally dxsum += sum_rows dx (the bias gradient of
// the linear layer that produced x: saves a separate column-sum pass over dx).
// Column partial sums live in a per-warp shared-memory slab [NS][C] (every lane owns its columns: plain
// read-modify-writes; shared float atomics would be CAS spin loops), folded across the warps at the end: one
// global atomic per column per block.  All 16-byte loads of a row are issued before the first use.
// NV = 16-byte vectors per lane.
__device__ __forceinline__ void cvt8(const uint4& u, float* f) {
  float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// F32: the gradient arrives as dy32 (fp32, optional: the residual stream of the backward pass) and / or dy2 (bf16: a
// projection head's or autograd's contribution), x is the fp32 pre-LayerNorm sum; dx32 (fp32) is the un-masked gradient
// that continues down the residual, dx / dx_drop the bf16 copies the dgrad / wgrad GEMMs consume.
template <int NV, bool DXSUM, bool F32>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const void* __restrict__ dy_, const __nv_bfloat16* __restrict__ dy2,
                     const void* __restrict__ x_, const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dx32,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum,
                     __nv_bfloat16* __restrict__ dx_drop, uint32_t drop_seed, uint32_t drop_thr, float drop_scale,
                     long long rows, int C) {
  pdl_sync();
  extern __shared__ float sred[];  // [warps][NS][C]
  constexpr int NS = DXSUM ? 3 : 2;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = C >> 3;
  float* mine = sred + (threadIdx.x >> 5) * NS * C;
  for (int i = lane; i < NS * C; i += 32) mine[i] = 0.f;
  __syncwarp();
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long row = warp_global; row < rows; row += nwarps) {
    float xh[NV][8], dxh[NV][8];
    uint4 rd2[NV], rr[NV];
    if (F32) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        rd2[i] = rr[i] = zero4;
        if (vi < nvec) {
          load8f(static_cast<const float*>(x_) + row * C + vi * 8, xh[i]);
          if (dy_) load8f(static_cast<const float*>(dy_) + row * C + vi * 8, dxh[i]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) dxh[i][j] = 0.f;
          }
          if (dy2) rd2[i] = *reinterpret_cast<const uint4*>(dy2 + row * C + vi * 8);
        }
      }
    } else {
      uint4 rx[NV], rd[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        rx[i] = rd[i] = rd2[i] = rr[i] = zero4;
        if (vi < nvec) {
          rx[i] = *reinterpret_cast<const uint4*>(static_cast<const __half*>(x_) + row * C + vi * 8);
          rd[i] = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(dy_) + row * C + vi * 8);
          if (dy2) rd2[i] = *reinterpret_cast<const uint4*>(dy2 + row * C + vi * 8);
          if (dres) rr[i] = *reinterpret_cast<const uint4*>(dres + row * C + vi * 8);
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        cvt8h(rx[i], xh[i]);  // the saved forward input: fp16
        cvt8(rd[i], dxh[i]);
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        if (dy2) {  // two gradient streams meet here (layer above + this layer's projection head)
          float t2[8];
          cvt8(rd2[i], t2);
#pragma unroll
          for (int j = 0; j < 8; ++j) dxh[i][j] += t2[j];
        }
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)),
                     g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float* mg = mine + vi;       // slab layout [NS][8][nvec]: lanes hit consecutive banks
        float* mb = mine + C + vi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          const float d = dxh[i][j];
          mg[j * nvec] = fmaf(d, xh[i][j], mg[j * nvec]);
          mb[j * nvec] += d;
          dxh[i][j] = d * g[j];
          s1 += dxh[i][j];
          s2 = fmaf(dxh[i][j], xh[i][j], s2);
        }
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float o[8];
        cvt8(rr[i], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rs * (dxh[i][j] - s1 - xh[i][j] * s2);
        if (dx) store8(dx + row * C + vi * 8, o);
        if (F32 && dx32) store8f(dx32 + row * C + vi * 8, o);
        if (dx_drop) {
          // gradient through nn.Dropout on the branch that produced x (residual + dropout(branch)): the masked
          // copy feeds the branch's dgrad / wgrad / bias gradient, the plain dx continues down the residual
          const uint32_t pair0 = (uint32_t)((row * C + vi * 8) >> 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float m0, m1;
            dropout_pair(drop_seed, pair0 + j, drop_thr, drop_scale, m0, m1);
            o[2 * j] *= m0;
            o[2 * j + 1] *= m1;
          }
          store8(dx_drop + row * C + vi * 8, o);
        }
        if (DXSUM) {
          float* md = mine + 2 * C + vi;
#pragma unroll
          for (int j = 0; j < 8; ++j) md[j * nvec] += o[j];
        }
      }
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < NS * C; i += blockDim.x) {
    const int ns = i / C, c = i - ns * C;
    const int slot = ns * C + (c & 7) * nvec + (c >> 3);
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sred[w * NS * C + slot];
    atomicAdd((ns == 0 ? dgamma : (ns == 1 ? dbeta : dxsum)) + c, t);
  }
}

int ln_grid(long long rows, int rows_per_warp_target) {
  long long blocks = (rows + 8LL * rows_per_warp_target - 1) / (8LL * rows_per_warp_target);
  const long long cap = (long long)fhb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

namespace {
int ln_fwd_launch(bool in_f32, const void* x, const float* gamma, const float* beta, void* y, float* y32, float* mean,
                  float* rstd, const float* sub32, void* diff_out, int64_t rows, int32_t C, float eps, fhb_stream_t stream) {
  FHB_ARG_CHECK(x && gamma && beta && y, "layernorm_fwd: null pointer");
  FHB_ARG_CHECK(rows >= 0 && C > 0 && C % 8 == 0 && C <= kMaxVec * 256, "layernorm_fwd: C=%d must be a multiple of 8, <= %d",
                C, kMaxVec * 256);
  FHB_ARG_CHECK((mean == nullptr) == (rstd == nullptr), "layernorm_fwd: mean and rstd go together");
  FHB_ARG_CHECK((sub32 == nullptr) == (diff_out == nullptr), "layernorm_fwd: sub32 and diff_out go together");
  if (rows == 0) return 0;
  fhb_pdl_hint(rows * C <= 16LL << 20);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_f32)
    FHB_CUDA_CHECK(fhb_launch(layernorm_fwd_kernel<true>, dim3(ln_grid(rows, 1)), dim3(256), 0, s, x, gamma, beta,
                              static_cast<__half*>(y), y32, mean, rstd, sub32, static_cast<__half*>(diff_out),
                              rows, C, eps));
  else
    FHB_CUDA_CHECK(fhb_launch(layernorm_fwd_kernel<false>, dim3(ln_grid(rows, 1)), dim3(256), 0, s, x, gamma, beta,
                              static_cast<__half*>(y), y32, mean, rstd, sub32, static_cast<__half*>(diff_out),
                              rows, C, eps));
  FHB_LAUNCH_CHECK();
  return 0;
}

int ln_bwd_launch(bool f32, const void* dy, const void* dy2, const void* x, const float* gamma, const float* mean,
                  const float* rstd, const void* dres, void* dx, float* dx32, float* dgamma, float* dbeta, float* dxsum,
                  void* dx_drop, uint32_t drop_seed, float drop_p, int64_t rows, int32_t C, fhb_stream_t stream) {
  FHB_ARG_CHECK(x && gamma && mean && rstd && dgamma && dbeta, "layernorm_bwd: null pointer");
  FHB_ARG_CHECK(f32 ? (dy || dy2) && (dx || dx32 || dx_drop) && !dres : (dy && dx && !dx32),
                "layernorm_bwd: missing gradient input / output");
  FHB_ARG_CHECK(!dx_drop || (drop_p >= 0.f && drop_p < 1.f && rows * C < (1LL << 32)), "layernorm_bwd: bad dropout arguments");
  FHB_ARG_CHECK(rows >= 0 && C > 0 && C % 8 == 0 && C <= kMaxVec * 256, "layernorm_bwd: bad C=%d", C);
  if (rows == 0) return 0;
  fhb_pdl_hint(rows * C <= 16LL << 20);
  // 2 blocks per SM (register-limited), every warp streams several rows
  long long blocks = (rows + 15) / 16;
  if (blocks > 2LL * fhb_num_sms()) blocks = 2LL * fhb_num_sms();
  const int nv = (C / 8 + 31) / 32;
  const size_t sm = 8 * (dxsum ? 3 : 2) * C * sizeof(float);  // [8 warps][NS][C]
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define FHB_LN_BWD(NV, DX, F)                                                                                        \
  do {                                                                                                               \
    FHB_ONCE_PER_DEVICE(FHB_CUDA_CHECK(cudaFuncSetAttribute((layernorm_bwd_kernel<NV, DX, F>),                        \
        cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * kMaxVec * 256 * (int)sizeof(float))));                  \
    FHB_CUDA_CHECK(fhb_launch((layernorm_bwd_kernel<NV, DX, F>), dim3((unsigned)blocks), dim3(256), sm, s, dy,       \
      static_cast<const __nv_bfloat16*>(dy2), x, gamma, mean, rstd,                                                  \
      static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dx32, dgamma, dbeta, dxsum,          \
      static_cast<__nv_bfloat16*>(dx_drop), drop_seed, fhb_dropout_thr16(drop_p), fhb_dropout_scale(drop_p), rows, C)); \
  } while (0)
#define FHB_LN_BWD_NV(DX, F) \
  do { if (nv == 1) FHB_LN_BWD(1, DX, F); else if (nv == 2) FHB_LN_BWD(2, DX, F); else FHB_LN_BWD(3, DX, F); } while (0)
  if (f32) {
    if (dxsum) FHB_LN_BWD_NV(true, true); else FHB_LN_BWD_NV(false, true);
  } else {
    if (dxsum) FHB_LN_BWD_NV(true, false); else FHB_LN_BWD_NV(false, false);
  }
#undef FHB_LN_BWD_NV
#undef FHB_LN_BWD
  FHB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

extern "C" int fhb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                 float* rstd, int64_t rows, int32_t C, float eps, fhb_stream_t stream) {
  return ln_fwd_launch(false, x, gamma, beta, y, nullptr, mean, rstd, nullptr, nullptr, rows, C, eps, stream);
}

extern "C" int fhb_layernorm_fwd32(const float* x32, const float* gamma, const float* beta, void* y, float* y32,
                                   float* mean, float* rstd, const float* sub32, void* diff_out, int64_t rows, int32_t C,
                                   float eps, fhb_stream_t stream) {
  return ln_fwd_launch(true, x32, gamma, beta, y, y32, mean, rstd, sub32, diff_out, rows, C, eps, stream);
}

extern "C" int fhb_layernorm_bwd(const void* dy, const void* dy2, const void* x, const float* gamma, const float* mean,
                                 const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 float* dxsum, void* dx_drop, uint32_t drop_seed, float drop_p, int64_t rows, int32_t C,
                                 fhb_stream_t stream) {
  return ln_bwd_launch(false, dy, dy2, x, gamma, mean, rstd, dres, dx, nullptr, dgamma, dbeta, dxsum, dx_drop, drop_seed,
                       drop_p, rows, C, stream);
}

extern "C" int fhb_layernorm_bwd32(const float* dy32, const void* dy2, const float* x32, const float* gamma,
                                   const float* mean, const float* rstd, void* dx, float* dx32, float* dgamma,
                                   float* dbeta, float* dxsum, void* dx_drop, uint32_t drop_seed, float drop_p,
                                   int64_t rows, int32_t C, fhb_stream_t stream) {
  return ln_bwd_launch(true, dy32, dy2, x32, gamma, mean, rstd, nullptr, dx, dx32, dgamma, dbeta, dxsum, dx_drop,
                       drop_seed, drop_p, rows, C, stream);
}
