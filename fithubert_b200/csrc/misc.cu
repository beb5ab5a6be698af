// HBM-bound helpers: fused distillation loss + gradient (K10), multi-tensor AdamW (K11), multi-tensor
// weight preparation (fp32 master -> bf16 GEMM layouts), bias-gradient column sums, mask lengths.
#include <math.h>

#include "fhb_common.cuh"

namespace {

// ------------------------------------------------------------------ K10 loss + gradient
// grid (chunks, n_layers).  pred [L][B][Tp][D], tgt [L][B][Tt][D] (first Tp frames used).
// The grid stride (gridDim.x * 256 vectors) is a multiple of D/8, so a thread always sees the same 8 columns:
// the column sums of the gradient (= bias gradient of the head's final Linear) accumulate in registers.
__global__ void __launch_bounds__(256)
distill_loss_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ tgt,
                    const float* __restrict__ weights, float* __restrict__ layer_loss, __nv_bfloat16* __restrict__ dpred,
                    float* __restrict__ dbias, long long dbias_stride, int B, int Tp, int Tt, int D, int loss_type,
                    float grad_scale) {
  pdl_sync();
  extern __shared__ float csum[];  // [D] (only when dbias)
  const int l = blockIdx.y;
  const float w = weights[l];
  const long long vec_per_row = D >> 3;
  const long long nvec = (long long)B * Tp * vec_per_row;
  const float inv_count = 1.0f / ((float)B * (float)Tp * (float)D);
  const float gs = grad_scale * w * inv_count * (loss_type == 0 ? 2.f : 1.f);
  const __nv_bfloat16* pl = pred + (long long)l * B * Tp * D;
  const __nv_bfloat16* tl = tgt + (long long)l * B * Tt * D;
  __nv_bfloat16* dl = dpred ? dpred + (long long)l * B * Tp * D : nullptr;
  if (dbias) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) csum[i] = 0.f;
    __syncthreads();
  }
  float acc = 0.f;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = i0; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int cv = i - row * vec_per_row;
    const int b = row / Tp, t = row - (long long)b * Tp;
    const uint4 pu = *reinterpret_cast<const uint4*>(pl + row * D + cv * 8);
    const uint4 tu = __ldg(reinterpret_cast<const uint4*>(tl + ((long long)b * Tt + t) * D + cv * 8));
    const uint32_t pa[4] = {pu.x, pu.y, pu.z, pu.w}, ta[4] = {tu.x, tu.y, tu.z, tu.w};
    uint32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 p = unpack_f16(pa[j]), q = unpack_f16(ta[j]);  // forward tensors: fp16; the gradient below: bf16
      const float d0 = p.x - q.x, d1 = p.y - q.y;
      float g0, g1;
      if (loss_type == 0) {
        acc += d0 * d0 + d1 * d1;
        g0 = gs * d0;
        g1 = gs * d1;
      } else {
        acc += fabsf(d0) + fabsf(d1);
        g0 = d0 > 0.f ? gs : (d0 < 0.f ? -gs : 0.f);
        g1 = d1 > 0.f ? gs : (d1 < 0.f ? -gs : 0.f);
      }
      out[j] = pack_f16(g0, g1);
      const float2 gr = unpack_f16(out[j]);  // sum what the downstream GEMMs will read
      cs[2 * j] += gr.x;
      cs[2 * j + 1] += gr.y;
    }
    if (dl) *reinterpret_cast<uint4*>(dl + row * D + cv * 8) = make_uint4(out[0], out[1], out[2], out[3]);
  }
  if (dbias && i0 < nvec) {
    const int cv = (int)(i0 % vec_per_row);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&csum[cv * 8 + j], cs[j]);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(layer_loss + l, s * w * inv_count);
  }
  if (dbias) {
    float* db = dbias + (long long)l * dbias_stride;
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(db + i, csum[i]);
  }
}

// ------------------------------------------------------------------ K10b reconstruction + cosine-similarity loss
// Reference train.py:282-314 with sim_loss_weight > 0 (the `distil_random_layer == 0` branch - the other one
// indexes a 3-D tensor with dim 3 and cannot run): per row r = (layer, b, t) of D features
//   rec = sum_d (p - q)^2 or |p - q|,   c = p.q / (max(|p|, eps) max(|q|, eps)),   sim = -logsigmoid(c)
//   dpred = gr * d rec/dp + gs * (-sigmoid(-c)) * (q / (|p||q|) - c p / |p|^2)
// Warp per row (16-byte vectors, NV per lane), rows dealt round-robin to the warps of blockIdx.y's layer; each
// lane owns fixed columns so the column sums of the gradient (bias gradient of the head's Linear) stay in
// registers until the block folds them through shared memory.
template <int NV>
__global__ void __launch_bounds__(256)
distill_loss_sim_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ tgt,
                        const float* __restrict__ weights, float* __restrict__ rec_loss, float* __restrict__ sim_loss,
                        __nv_bfloat16* __restrict__ dpred, float* __restrict__ dbias, long long dbias_stride, int B,
                        int Tp, int Tt, int D, int loss_type, float rec_grad_scale, float sim_grad_scale) {
  pdl_sync();
  extern __shared__ float csum[];  // [warps][D] (only when dbias)
  const int l = blockIdx.y;
  const float w = weights[l];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nvec = D >> 3;
  const long long rows = (long long)B * Tp;
  const float inv_rows = 1.0f / (float)rows, inv_elems = inv_rows / (float)D;
  const float gr = rec_grad_scale * w * inv_elems * (loss_type == 0 ? 2.f : 1.f);
  const float gs = sim_grad_scale * w * inv_rows;
  const __nv_bfloat16* pl = pred + (long long)l * rows * D;
  const __nv_bfloat16* tl = tgt + (long long)l * B * Tt * D;
  __nv_bfloat16* dl = dpred ? dpred + (long long)l * rows * D : nullptr;
  float cs[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[i][j] = 0.f;
  float acc_rec = 0.f, acc_sim = 0.f;
  for (long long r = (long long)blockIdx.x * nw + warp; r < rows; r += (long long)gridDim.x * nw) {
    const int b = (int)(r / Tp), t = (int)(r - (long long)b * Tp);
    const __nv_bfloat16* pr = pl + r * D;
    const __nv_bfloat16* qr = tl + ((long long)b * Tt + t) * D;
    uint4 pu[NV], qu[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      pu[i] = qu[i] = make_uint4(0u, 0u, 0u, 0u);
      if (vi < nvec) {
        pu[i] = *reinterpret_cast<const uint4*>(pr + vi * 8);
        qu[i] = __ldg(reinterpret_cast<const uint4*>(qr + vi * 8));
      }
    }
    float p[NV][8], q[NV][8];
    float dot = 0.f, pp = 0.f, qq = 0.f, rec = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t pa[4] = {pu[i].x, pu[i].y, pu[i].z, pu[i].w}, qa[4] = {qu[i].x, qu[i].y, qu[i].z, qu[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_f16(pa[j]), c = unpack_f16(qa[j]);
        p[i][2 * j] = a.x; p[i][2 * j + 1] = a.y;
        q[i][2 * j] = c.x; q[i][2 * j + 1] = c.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dot = fmaf(p[i][j], q[i][j], dot);
        pp = fmaf(p[i][j], p[i][j], pp);
        qq = fmaf(q[i][j], q[i][j], qq);
        const float d = p[i][j] - q[i][j];
        rec += loss_type == 0 ? d * d : fabsf(d);
      }
    }
    dot = warp_sum(dot);
    pp = warp_sum(pp);
    qq = warp_sum(qq);
    rec = warp_sum(rec);
    const float eps = 1e-8f;
    const float np = fmaxf(sqrtf(pp), eps), nq = fmaxf(sqrtf(qq), eps);
    const float c = dot / (np * nq);
    // -logsigmoid(c) = softplus(-c), evaluated the overflow-safe way
    const float sim = fmaxf(-c, 0.f) + log1pf(expf(-fabsf(c)));
    const float dsim = -1.0f / (1.0f + expf(c));          // d sim / d c
    const float k_q = gs * dsim / (np * nq);              // coefficient of q
    const float k_p = sqrtf(pp) > eps ? -gs * dsim * c / (np * np) : 0.f;  // coefficient of p (0 in the clamped regime)
    acc_rec += rec;
    acc_sim += sim;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = p[i][2 * j + e], qv = q[i][2 * j + e];
            const float d = pv - qv;
            const float grec = loss_type == 0 ? gr * d : (d > 0.f ? gr : (d < 0.f ? -gr : 0.f));
            g2[e] = grec + k_q * qv + k_p * pv;
          }
          o[j] = pack_f16(g2[0], g2[1]);
          const float2 gb = unpack_f16(o[j]);  // sum what the downstream GEMMs will read
          cs[i][2 * j] += gb.x;
          cs[i][2 * j + 1] += gb.y;
        }
        if (dl) *reinterpret_cast<uint4*>(dl + r * D + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  __shared__ float red[2][8];
  if (lane == 0) {  // acc_* are already warp-uniform (built from warp_sum results)
    red[0][warp] = acc_rec;
    red[1][warp] = acc_sim;
  }
  if (dbias) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) csum[warp * D + vi * 8 + j] = cs[i][j];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c2 = 0.f;
    for (int i = 0; i < nw; ++i) {
      a += red[0][i];
      c2 += red[1][i];
    }
    atomicAdd(rec_loss + l, a * w * inv_elems);
    atomicAdd(sim_loss + l, c2 * w * inv_rows);
  }
  if (dbias) {
    float* db = dbias + (long long)l * dbias_stride;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float t2 = 0.f;
      for (int w2 = 0; w2 < nw; ++w2) t2 += csum[w2 * D + i];
      atomicAdd(db + i, t2);
    }
  }
}

// ------------------------------------------------------------------ K11 AdamW over a tensor table
struct AdamScalars {
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale;
  int mode;
};

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const fhb_adamw_tensor* __restrict__ table, AdamScalars s) {
  pdl_sync();
  const fhb_adamw_tensor e = table[blockIdx.y];
  if (e.g == nullptr) return;
  const long long d1 = e.dim[1], d2 = e.dim[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (long long)gridDim.x * blockDim.x) {
    const long long i2 = i % d2, r = i / d2;
    const long long i1 = r % d1, i0 = r / d1;
    const float g = e.g[i0 * e.gstride[0] + i1 * e.gstride[1] + i2 * e.gstride[2]] * s.grad_scale;
    float p = e.p[i];
    const float m = s.b1 * e.m[i] + (1.f - s.b1) * g;
    const float v = s.b2 * e.v[i] + (1.f - s.b2) * g * g;
    e.m[i] = m;
    e.v[i] = v;
    if (s.mode == 0) {
      // s3prl Lamb(adam=True, correct_bias=True): eps on the un-corrected sqrt(v), decay inside the step
      const float step = s.lr * s.bc2_sqrt / s.bc1;
      p -= step * (m / (sqrtf(v) + s.eps) + s.wd * p);
    } else {
      // torch.optim.AdamW
      p *= 1.f - s.lr * s.wd;
      p -= (s.lr / s.bc1) * m / (sqrtf(v) / s.bc2_sqrt + s.eps);
    }
    e.p[i] = p;
  }
}

// ------------------------------------------------------------------ weight preparation
// dst (bf16 or fp32, contiguous [d0][d1][d2]) = src_fp32[i0*s0 + i1*s1 + i2*s2]
__global__ void __launch_bounds__(256) prep_multi_kernel(const fhb_prep_tensor* __restrict__ table) {
  pdl_sync();
  const fhb_prep_tensor e = table[blockIdx.y];
  const long long d1 = e.dim[1], d2 = e.dim[2];
  const long long n = e.dim[0] * d1 * d2;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
  // fast path (most entries: Linear weights, biases): the source is already laid out like the destination, so the
  // entry is a plain fp32 -> bf16 cast - 16-byte loads, 8-byte stores, no index arithmetic
  const bool contiguous = e.sstride[2] == 1 && (d1 == 1 || e.sstride[1] == d2) && (e.dim[0] == 1 || e.sstride[0] == d1 * d2);
  if (contiguous && !e.dst_is_f32 && (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(e.src) & 15) == 0) &&
      ((reinterpret_cast<uintptr_t>(e.dst) & 7) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(e.src);
    uint2* d2p = reinterpret_cast<uint2*>(e.dst);
    for (long long i = i0; i < (n >> 2); i += step) {
      const float4 v = __ldg(s4 + i);
      d2p[i] = make_uint2(pack_f16(v.x, v.y), pack_f16(v.z, v.w));
    }
    return;
  }
  for (long long i = i0; i < n; i += step) {
    const long long i2 = i % d2, r = i / d2;
    const long long i1 = r % d1, i0b = r / d1;
    const float v = e.src[i0b * e.sstride[0] + i1 * e.sstride[1] + i2 * e.sstride[2]];
    if (e.dst_is_f32)
      static_cast<float*>(e.dst)[i] = e.accumulate ? static_cast<float*>(e.dst)[i] + v : v;
    else
      static_cast<__half*>(e.dst)[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  }
}

// ------------------------------------------------------------------ column sums (bias gradients)
// Block = (64-column slice, row chunk): 256 threads = 32 row lanes x 8 sixteen-byte column chunks, so a warp
// load covers 4 rows x 128 contiguous bytes and a thread carries only 8 accumulators (deep unrolling, many
// loads in flight).  Row lanes are folded through shared memory; one atomic per column per block, and only
// `row chunks` blocks ever hit the same address (a flat row-striped grid made every block hit every column).
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, long long ld, float* __restrict__ out,
              long long x_bstride, long long out_bstride, int rows_per_chunk) {
  pdl_sync();
  __shared__ float red[32][65];
  x += (long long)blockIdx.z * x_bstride;
  out += (long long)blockIdx.z * out_bstride;
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + cl * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < C) {
    const __nv_bfloat16* xp = x + col;
#pragma unroll 8
    for (long long r = r0 + rl; r < r1; r += 32) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xp + r * ld));
      const uint32_t a[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_f16(a[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cl * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    atomicAdd(out + blockIdx.x * 64 + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ projection-head bias gradients
// dz = dpred @ Wlin has no bias term, so the column sums of dz (the gradient of the ConvTranspose1d bias) follow
// from the column sums of dpred the loss kernel already produced:  colsum(dz)[e] = sum_d colsum(dpred)[d] Wlin[d][e]
// - a [1 x D] x [D x E] product per head instead of another pass over the n x B x T' x E tensor dz.
// grid (D / 64, heads), block = E / 2 threads (two columns each); also folds colsum(dpred) into the lin_proj bias
// gradient (each element owned by exactly one block).
__global__ void __launch_bounds__(512)
head_bias_grads_kernel(const float* __restrict__ cs, long long cs_stride, const __nv_bfloat16* __restrict__ wlin,
                       long long wlin_stride, float* __restrict__ dlin_bias, float* __restrict__ dup_bias,
                       long long grad_stride, int D, int E) {
  pdl_sync();
  __shared__ float cs_s[64];
  const int h = blockIdx.y, d0 = blockIdx.x * 64;
  const int nd = min(64, D - d0);
  const float* csh = cs + (long long)h * cs_stride + d0;
  if ((int)threadIdx.x < nd) {
    const float v = csh[threadIdx.x];
    cs_s[threadIdx.x] = v;
    if (dlin_bias) dlin_bias[(long long)h * grad_stride + d0 + threadIdx.x] += v;
  }
  __syncthreads();
  const int e = threadIdx.x * 2;
  if (e >= E) return;
  const __nv_bfloat16* w = wlin + (long long)h * wlin_stride + (long long)d0 * E + e;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
  for (int d = 0; d < nd; ++d) {
    const float2 f = unpack_f16(__ldg(reinterpret_cast<const uint32_t*>(w + (long long)d * E)));
    a0 = fmaf(cs_s[d], f.x, a0);
    a1 = fmaf(cs_s[d], f.y, a1);
  }
  float* o = dup_bias + (long long)h * grad_stride + e;
  atomicAdd(o, a0);
  atomicAdd(o + 1, a1);
}

__global__ void __launch_bounds__(256)
add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y,
                long long nvec) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i], ub = reinterpret_cast<const uint4*>(b)[i];
    const uint32_t aa[4] = {ua.x, ua.y, ua.z, ua.w}, bb[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 p = unpack_f16(aa[j]), q = unpack_f16(bb[j]);
      o[j] = pack_f16(p.x + q.x, p.y + q.y);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[b][i] = dy[b][i] * gelu'(u[b][i]) over B segments of n elements with independent batch strides
__global__ void __launch_bounds__(256)
mul_dgelu_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_bs, const __nv_bfloat16* __restrict__ u, long long u_bs,
                 __nv_bfloat16* __restrict__ out, long long out_bs, long long nvec) {
  pdl_sync();
  const int b = blockIdx.y;
  const uint4* d4 = reinterpret_cast<const uint4*>(dy + b * dy_bs);
  const uint4* u4 = reinterpret_cast<const uint4*>(u + b * u_bs);
  uint4* o4 = reinterpret_cast<uint4*>(out + b * out_bs);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = d4[i], c = u4[i];
    const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, cc[4] = {c.x, c.y, c.z, c.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 p = unpack_f16(aa[j]), q = unpack_f16(cc[j]);
      o[j] = pack_f16(p.x * gelu_erf_grad(q.x), p.y * gelu_erf_grad(q.y));
    }
    o4[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[b][i] = a[b][i] * m[b][i] (m = a saved multiplier, e.g. gelu' from FHB_EPI_AUX_DGELU)
__global__ void __launch_bounds__(256)
mul_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long a_bs, const __nv_bfloat16* __restrict__ m, long long m_bs,
                __nv_bfloat16* __restrict__ out, long long out_bs, long long nvec, float alpha) {
  pdl_sync();
  const int b = blockIdx.y;
  const uint4* a4 = reinterpret_cast<const uint4*>(a + b * a_bs);
  const uint4* m4 = reinterpret_cast<const uint4*>(m + b * m_bs);
  uint4* o4 = reinterpret_cast<uint4*>(out + b * out_bs);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 x = a4[i], c = m4[i];
    const uint32_t aa[4] = {x.x, x.y, x.z, x.w}, cc[4] = {c.x, c.y, c.z, c.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 p = unpack_f16(aa[j]), q = unpack_f16(cc[j]);  // gradient (bf16) x saved gelu' (fp16)
      o[j] = pack_f16(alpha * p.x * q.x, alpha * p.y * q.y);
    }
    o4[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256)
dropout_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, long long nvec, uint32_t seed,
               uint32_t thr, float scale, bool f16) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = reinterpret_cast<const uint4*>(x)[i];
    const uint32_t aa[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float m0, m1;
      dropout_pair(seed, (uint32_t)(i * 4 + j), thr, scale, m0, m1);
      const float2 p = unpack16(aa[j], f16);
      o[j] = pack16(p.x * m0, p.y * m1, f16);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// lengths[b] = number of zero bytes in mask[b][0..L)   (mask: 1 = padding)
__global__ void __launch_bounds__(256) mask_lengths_kernel(const uint8_t* __restrict__ mask, long long L, int* __restrict__ lengths) {
  pdl_sync();
  const uint8_t* row = mask + (long long)blockIdx.x * L;
  int cnt = 0;
  for (long long i = threadIdx.x; i < L; i += blockDim.x) cnt += row[i] ? 0 : 1;
  __shared__ int red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 8; ++i) s += red[i];
    lengths[blockIdx.x] = s;
  }
}

int grid_x(long long work_items, int cap_mult) {
  long long b = (work_items + 255) / 256;
  const long long cap = (long long)fhb_num_sms() * cap_mult;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int fhb_distill_loss_fwd_bwd(const void* pred, const void* tgt, const float* weights, float* layer_loss,
                                        void* dpred, float* dbias, int64_t dbias_layer_stride, int32_t n_layers,
                                        int32_t B, int32_t Tp, int32_t Tt, int32_t D, int32_t loss_type,
                                        float grad_scale, fhb_stream_t stream) {
  FHB_ARG_CHECK(pred && tgt && weights && layer_loss, "distill_loss: null pointer");
  FHB_ARG_CHECK(n_layers > 0 && B > 0 && Tp > 0 && Tt >= Tp && D > 0 && D % 8 == 0,
                "distill_loss: bad shape (layers=%d B=%d Tp=%d Tt=%d D=%d)", n_layers, B, Tp, Tt, D);
  FHB_ARG_CHECK(loss_type == 0 || loss_type == 1, "rec_loss_type must be one of 'l1', 'mse'.");
  FHB_ARG_CHECK(!dbias || dpred, "distill_loss: dbias needs dpred");
  const long long nvec = (long long)B * Tp * (D / 8);
  int gx = grid_x(nvec, 8);
  gx = (gx + n_layers - 1) / n_layers;
  if (gx < 1) gx = 1;
  // make the grid stride a multiple of the row length in vectors: gx * 256 % (D/8) == 0
  long long vpr = D / 8, g = vpr, a = 256;
  while (a) { const long long t2 = g % a; g = a; a = t2; }  // g = gcd(vpr, 256)
  const int unit = (int)(vpr / g);
  gx = (gx + unit - 1) / unit * unit;
  FHB_CUDA_CHECK(fhb_launch(distill_loss_kernel, dim3(dim3(gx, n_layers)), dim3(256), dbias ? D * sizeof(float) : 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(pred), static_cast<const __nv_bfloat16*>(tgt), weights, layer_loss,
      static_cast<__nv_bfloat16*>(dpred), dbias, dbias_layer_stride, B, Tp, Tt, D, loss_type, grad_scale));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_distill_loss_sim_fwd_bwd(const void* pred, const void* tgt, const float* weights, float* rec_layer_loss,
                                            float* sim_layer_loss, void* dpred, float* dbias, int64_t dbias_layer_stride,
                                            int32_t n_layers, int32_t B, int32_t Tp, int32_t Tt, int32_t D,
                                            int32_t loss_type, float rec_grad_scale, float sim_grad_scale,
                                            fhb_stream_t stream) {
  FHB_ARG_CHECK(pred && tgt && weights && rec_layer_loss && sim_layer_loss, "distill_loss_sim: null pointer");
  FHB_ARG_CHECK(n_layers > 0 && B > 0 && Tp > 0 && Tt >= Tp && D > 0 && D % 8 == 0 && D <= 1024,
                "distill_loss_sim: bad shape (layers=%d B=%d Tp=%d Tt=%d D=%d)", n_layers, B, Tp, Tt, D);
  FHB_ARG_CHECK(loss_type == 0 || loss_type == 1, "rec_loss_type must be one of 'l1', 'mse'.");
  FHB_ARG_CHECK(!dbias || dpred, "distill_loss_sim: dbias needs dpred");
  const long long rows = (long long)B * Tp;
  long long gx = (4LL * fhb_num_sms() + n_layers - 1) / n_layers;
  if (gx > (rows + 7) / 8) gx = (rows + 7) / 8;
  if (gx < 1) gx = 1;
  const int nv = (D / 8 + 31) / 32;
  const size_t smem = dbias ? 8 * (size_t)D * sizeof(float) : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define FHB_SIM(NV)                                                                                                    \
  FHB_CUDA_CHECK(fhb_launch(distill_loss_sim_kernel<NV>, dim3((unsigned)gx, n_layers), dim3(256), smem, s,             \
                            static_cast<const __nv_bfloat16*>(pred), static_cast<const __nv_bfloat16*>(tgt), weights, \
                            rec_layer_loss, sim_layer_loss, static_cast<__nv_bfloat16*>(dpred), dbias,                 \
                            (long long)dbias_layer_stride, B, Tp, Tt, D, loss_type, rec_grad_scale, sim_grad_scale))
  if (nv <= 1) FHB_SIM(1); else if (nv == 2) FHB_SIM(2); else if (nv == 3) FHB_SIM(3); else FHB_SIM(4);
#undef FHB_SIM
  return 0;
}

extern "C" int fhb_adamw_multi(const fhb_adamw_tensor* table_dev, int32_t n_tensors, int64_t max_n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int32_t step, int32_t mode,
                               float grad_scale, fhb_stream_t stream) {
  FHB_ARG_CHECK(table_dev && n_tensors > 0 && max_n > 0, "adamw: empty table");
  FHB_ARG_CHECK(step >= 1, "adamw: step is 1-based");
  FHB_ARG_CHECK(mode == 0 || mode == 1, "adamw: mode must be 0 (s3prl) or 1 (torch)");
  AdamScalars s;
  s.lr = lr; s.b1 = beta1; s.b2 = beta2; s.eps = eps; s.wd = weight_decay; s.grad_scale = grad_scale; s.mode = mode;
  s.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  s.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int gx = grid_x(max_n, 2);
  if (gx > 64) gx = 64;
  FHB_CUDA_CHECK(fhb_launch(adamw_multi_kernel, dim3(dim3(gx, n_tensors)), dim3(256), 0, static_cast<cudaStream_t>(stream), table_dev, s));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_prep_multi(const fhb_prep_tensor* table_dev, int32_t n_tensors, int64_t max_n, fhb_stream_t stream) {
  FHB_ARG_CHECK(table_dev && n_tensors > 0 && max_n > 0, "prep_multi: empty table");
  int gx = grid_x(max_n, 2);
  if (gx > 64) gx = 64;
  FHB_CUDA_CHECK(fhb_launch(prep_multi_kernel, dim3(dim3(gx, n_tensors)), dim3(256), 0, static_cast<cudaStream_t>(stream), table_dev));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_colsum_batched(const void* x, int64_t rows, int32_t C, int64_t ld, int64_t x_bstride, float* out,
                                  int64_t out_bstride, int32_t batches, fhb_stream_t stream) {
  FHB_ARG_CHECK(x && out && C > 0 && C % 8 == 0 && ld % 8 == 0 && batches > 0 && x_bstride % 8 == 0,
                "colsum: bad arguments (C=%d)", C);
  if (rows == 0) return 0;
  const int slices = (C + 63) / 64;
  // about 4 blocks per SM in total, at least 64 rows (2 per row lane) per chunk
  long long chunks = (4LL * fhb_num_sms() + (long long)slices * batches - 1) / ((long long)slices * batches);
  const long long max_chunks = (rows + 63) / 64;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int rpc = (int)((rows + chunks - 1) / chunks);
  chunks = (rows + rpc - 1) / rpc;
  fhb_pdl_hint(rows * C * batches <= 32LL << 20);
  FHB_CUDA_CHECK(fhb_launch(colsum_kernel, dim3(slices, (unsigned)chunks, batches), dim3(256), 0,
                            static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), rows, C, ld, out,
                            x_bstride, out_bstride, rpc));
  return 0;
}

extern "C" int fhb_colsum(const void* x, int64_t rows, int32_t C, int64_t ld, float* out, fhb_stream_t stream) {
  return fhb_colsum_batched(x, rows, C, ld, 0, out, 0, 1, stream);
}

extern "C" int fhb_head_bias_grads(const float* colsum_dpred, int64_t cs_stride, const void* wlin, int64_t wlin_stride,
                                   float* dlin_bias, float* dup_bias, int64_t grad_stride, int32_t n_heads, int32_t D,
                                   int32_t E, fhb_stream_t stream) {
  FHB_ARG_CHECK(colsum_dpred && wlin && dup_bias && n_heads > 0 && D > 0, "head_bias_grads: bad arguments");
  FHB_ARG_CHECK(E > 0 && E % 2 == 0 && E <= 1024, "head_bias_grads: E=%d must be even and <= 1024", E);
  FHB_CUDA_CHECK(fhb_launch(head_bias_grads_kernel, dim3((D + 63) / 64, n_heads), dim3((E / 2 + 31) / 32 * 32 < 64 ? 64 : (E / 2 + 31) / 32 * 32), 0,
                            static_cast<cudaStream_t>(stream), colsum_dpred, (long long)cs_stride,
                            static_cast<const __nv_bfloat16*>(wlin), (long long)wlin_stride, dlin_bias, dup_bias,
                            (long long)grad_stride, D, E));
  return 0;
}

extern "C" int fhb_add_bf16(const void* a, const void* b, void* y, int64_t n, fhb_stream_t stream) {
  FHB_ARG_CHECK(a && b && y && n % 8 == 0, "add_bf16: n must be a multiple of 8");
  if (n == 0) return 0;
  FHB_CUDA_CHECK(fhb_launch(add_bf16_kernel, dim3(grid_x(n / 8, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(y), n / 8));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_mul_dgelu(const void* dy, int64_t dy_bstride, const void* u, int64_t u_bstride, void* out,
                             int64_t out_bstride, int32_t B, int64_t n, fhb_stream_t stream) {
  FHB_ARG_CHECK(dy && u && out && B > 0, "mul_dgelu: null pointer");
  FHB_ARG_CHECK(n % 8 == 0 && dy_bstride % 8 == 0 && u_bstride % 8 == 0 && out_bstride % 8 == 0,
                "mul_dgelu: sizes and strides must be multiples of 8 elements");
  if (n == 0) return 0;
  int gx = grid_x(n / 8, 8);
  gx = (gx + B - 1) / B;
  FHB_CUDA_CHECK(fhb_launch(mul_dgelu_kernel, dim3(dim3(gx < 1 ? 1 : gx, B)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(dy), dy_bstride, static_cast<const __nv_bfloat16*>(u), u_bstride,
      static_cast<__nv_bfloat16*>(out), out_bstride, n / 8));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_mul_bf16(const void* a, int64_t a_bstride, const void* m, int64_t m_bstride, void* out,
                            int64_t out_bstride, int32_t B, int64_t n, float alpha, fhb_stream_t stream) {
  FHB_ARG_CHECK(a && m && out && B > 0, "mul_bf16: null pointer");
  FHB_ARG_CHECK(n % 8 == 0 && a_bstride % 8 == 0 && m_bstride % 8 == 0 && out_bstride % 8 == 0,
                "mul_bf16: sizes and strides must be multiples of 8 elements");
  if (n == 0) return 0;
  int gx = grid_x(n / 8, 8);
  gx = (gx + B - 1) / B;
  FHB_CUDA_CHECK(fhb_launch(mul_bf16_kernel, dim3(dim3(gx < 1 ? 1 : gx, B)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(a), a_bstride, static_cast<const __nv_bfloat16*>(m), m_bstride,
      static_cast<__nv_bfloat16*>(out), out_bstride, n / 8, alpha));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_dropout(const void* x, void* y, int64_t n, uint32_t seed, float p, int32_t is_f16,
                           fhb_stream_t stream) {
  FHB_ARG_CHECK(x && y && n % 8 == 0 && n < (1LL << 32), "dropout: n must be a multiple of 8 and < 2^32");
  FHB_ARG_CHECK(p >= 0.f && p < 1.f, "dropout: p=%f must be in [0, 1)", (double)p);
  if (n == 0) return 0;
  fhb_pdl_hint(n <= 16LL << 20);
  FHB_CUDA_CHECK(fhb_launch(dropout_kernel, dim3(grid_x(n / 8, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n / 8, seed, fhb_dropout_thr16(p),
      fhb_dropout_scale(p), is_f16 != 0));
  FHB_LAUNCH_CHECK();
  return 0;
}

extern "C" int fhb_memset2d(void* ptr, int64_t pitch_bytes, int64_t width_bytes, int64_t height, fhb_stream_t stream) {
  FHB_ARG_CHECK(ptr && pitch_bytes >= width_bytes && width_bytes >= 0 && height >= 0, "memset2d: bad arguments");
  if (width_bytes == 0 || height == 0) return 0;
  FHB_CUDA_CHECK(cudaMemset2DAsync(ptr, (size_t)pitch_bytes, 0, (size_t)width_bytes, (size_t)height,
                                   static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int fhb_mask_lengths(const uint8_t* mask, int32_t B, int64_t L, int32_t* lengths, fhb_stream_t stream) {
  FHB_ARG_CHECK(mask && lengths && B > 0 && L > 0, "mask_lengths: bad arguments");
  FHB_CUDA_CHECK(fhb_launch(mask_lengths_kernel, dim3(B), dim3(256), 0, static_cast<cudaStream_t>(stream), mask, L, lengths));
  FHB_LAUNCH_CHECK();
  return 0;
}
