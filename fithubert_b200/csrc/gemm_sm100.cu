// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma
// (UMMA M=128, N<=256, K=16, accumulators in TMEM, double buffered) -> tcgen05.ld -> fused epilogue.
//
// One CTA per SM, 192 threads:
//   warps 0-3  epilogue (warp w owns TMEM lanes 32w..32w+31 = tile rows)
//   warp  4    TMA producer (one elected lane)
//   warp  5    MMA issuer (one elected lane) + TMEM allocation
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static
// round-robin tile schedule (tile = blockIdx.x + i * gridDim.x, n fastest so CTAs of one wave
// share the A tile in L2).
#include <stdarg.h>
#include <string.h>

#include "fhb_common.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kMaxBN = 256;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr uint32_t kABytes = kBM * kBK * 2;     // 16 KiB
constexpr uint32_t kBBytes = kMaxBN * kBK * 2;  // 32 KiB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kThreads = 192;

struct GemmParams {
  int m, n, k, bn;
  int num_m_blk, num_n_blk, num_ob, ob_mod, split_k;
  int kb_per_cb, kb_total, kb_per_split;
  int a_lo_c0, a_hi_c2, a_lo_c2, a_cb_c2;
  int b_lo_c0, b_hi_c2, b_lo_c2, b_cb_c2;
  int a_c1_off, b_c1_off;
  long long d_ld, d_hi_stride, d_lo_stride;
  void* d;
  const float* bias;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* aux_in;
  __nv_bfloat16* aux_out;
  const int* row_valid;
  const __nv_bfloat16* loss_target;
  float* loss_acc;
  float loss_weight, grad_scale;
  int flags;
  uint32_t stage_tx_bytes;
  int total_tiles;
};

struct Tile {
  int ob_hi, ob_lo, m0, n0, kb_begin, kb_end;
};

__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int tile) {
  Tile t;
  const int nb = tile % p.num_n_blk;
  tile /= p.num_n_blk;
  const int mb = tile % p.num_m_blk;
  tile /= p.num_m_blk;
  const int sp = tile % p.split_k;
  const int ob = tile / p.split_k;
  t.ob_hi = ob / p.ob_mod;
  t.ob_lo = ob % p.ob_mod;
  t.m0 = mb * kBM;
  t.n0 = nb * p.bn;
  t.kb_begin = sp * p.kb_per_split;
  t.kb_end = min(t.kb_begin + p.kb_per_split, p.kb_total);
  return t;
}

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1)
fhb_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = bars + 2 * kStages + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const Tile t = decode_tile(p, tile);
        const int a_c2 = t.ob_hi * p.a_hi_c2 + t.ob_lo * p.a_lo_c2;
        const int b_c2 = t.ob_hi * p.b_hi_c2 + t.ob_lo * p.b_lo_c2;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], p.stage_tx_bytes);
          const int cb = kb / p.kb_per_cb;
          const int kr = (kb - cb * p.kb_per_cb) * kBK;
          uint8_t* sa = smem_a + stage * kABytes;
          uint8_t* sb = smem_b + stage * kBBytes;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i)
              tma_load_3d(&tm_a, &full[stage], sa + i * (kBK * 128), t.m0 + i * 64 + t.ob_lo * p.a_lo_c0, kr + p.a_c1_off,
                          a_c2 + cb * p.a_cb_c2);
          } else {
            tma_load_3d(&tm_a, &full[stage], sa, kb * kBK + t.ob_lo * p.a_lo_c0, t.m0, a_c2);
          }
          if (B_MN) {
            const int atoms = (p.bn + 63) >> 6;
            for (int i = 0; i < atoms; ++i)
              tma_load_3d(&tm_b, &full[stage], sb + i * (kBK * 128), t.n0 + i * 64 + t.ob_lo * p.b_lo_c0, kr + p.b_c1_off,
                          b_c2 + cb * p.b_cb_c2);
          } else {
            tma_load_3d(&tm_b, &full[stage], sb, kb * kBK + t.ob_lo * p.b_lo_c0, t.n0, b_c2);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kBM, (uint32_t)p.bn, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart (SBO); MN-major: 64-element atoms kBK*128 B apart (LBO),
      // 8-k-row groups 1024 B apart (SBO).
      const uint32_t a_lbo = A_MN ? kBK * 128 : 0, b_lbo = B_MN ? kBK * 128 : 0;
      const uint32_t a_kstep = A_MN ? 16 * 128 : 32, b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const Tile t = decode_tile(p, tile);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kMaxBN;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t db = umma_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            tc_mma_bf16(tmem_d, da, db, idesc, (kb > t.kb_begin || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&acc_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0-3)
    const int flags = p.flags;
    int it = 0;
    float loss_local = 0.f;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const Tile t = decode_tile(p, tile);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const int row = t.m0 + warp * 32 + lane;
      const bool row_ok = row < p.m;
      bool zero_row = false;
      if ((flags & FHB_EPI_ROWZERO) && row_ok) zero_row = row >= p.row_valid[t.ob_hi];
      const long long off = (long long)t.ob_hi * p.d_hi_stride + (long long)t.ob_lo * p.d_lo_stride +
                            (long long)row * p.d_ld + t.n0;
      const uint32_t taddr = tmem_base + as * kMaxBN + ((uint32_t)(warp * 32) << 16);
      for (int c = 0; c < p.bn; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (flags & FHB_EPI_BIAS) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + t.n0 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(bp + j);
            v[4 * j] += b4.x;
            v[4 * j + 1] += b4.y;
            v[4 * j + 2] += b4.z;
            v[4 * j + 3] += b4.w;
          }
        }
        if (flags & FHB_EPI_STORE_PREACT) {
          uint4* ap = reinterpret_cast<uint4*>(p.aux_out + off + c);
          ap[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          ap[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                             pack_bf16(v[14], v[15]));
        }
        if (flags & FHB_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
        }
        if (flags & FHB_EPI_MUL_DGELU) {
          const uint4* ap = reinterpret_cast<const uint4*>(p.aux_in + off + c);
          const uint4 u0 = __ldg(ap), u1 = __ldg(ap + 1);
          const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = unpack_bf16(uu[j]);
            v[2 * j] *= gelu_erf_grad(f.x);
            v[2 * j + 1] *= gelu_erf_grad(f.y);
          }
        }
        if (flags & FHB_EPI_RESIDUAL) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + c);
          const uint4 u0 = __ldg(rp), u1 = __ldg(rp + 1);
          const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = unpack_bf16(uu[j]);
            v[2 * j] += f.x;
            v[2 * j + 1] += f.y;
          }
        }
        if (flags & FHB_EPI_SQDIFF) {
          const uint4* tp = reinterpret_cast<const uint4*>(p.loss_target + off + c);
          const uint4 u0 = __ldg(tp), u1 = __ldg(tp + 1);
          const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = unpack_bf16(uu[j]);
            const float d0 = v[2 * j] - f.x, d1 = v[2 * j + 1] - f.y;
            loss_local += d0 * d0 + d1 * d1;
            v[2 * j] = d0 * p.grad_scale;
            v[2 * j + 1] = d1 * p.grad_scale;
          }
        }
        if (zero_row) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (flags & FHB_EPI_OUT_F32) {
          float* dp = reinterpret_cast<float*>(p.d) + off + c;
          if (flags & FHB_EPI_ATOMIC_ADD) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dp + j, v[j]);
          } else {
            float4* d4 = reinterpret_cast<float4*>(dp);
#pragma unroll
            for (int j = 0; j < 4; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.d) + off + c);
          dp[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          dp[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                             pack_bf16(v[14], v[15]));
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
    if (flags & FHB_EPI_SQDIFF) {
      loss_local = warp_sum(loss_local);
      if (lane == 0 && loss_local != 0.f) atomicAdd(p.loss_acc, loss_local * p.loss_weight);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* tm, const fhb_tensor3& t, uint32_t box0, uint32_t box1, const char* name) {
  EncodeTiledFn enc = get_encode_fn();
  FHB_ARG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  FHB_ARG_CHECK(t.ptr != nullptr && (reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "gemm: %s pointer must be 16B aligned", name);
  FHB_ARG_CHECK(t.dim[0] > 0 && t.dim[1] > 0 && t.dim[2] > 0, "gemm: %s has an empty dimension", name);
  FHB_ARG_CHECK(t.stride[0] % 8 == 0 && t.stride[0] > 0, "gemm: %s row stride %lld must be a positive multiple of 8 elements",
                name, (long long)t.stride[0]);
  cuuint64_t dims[3] = {(cuuint64_t)t.dim[0], (cuuint64_t)t.dim[1], (cuuint64_t)t.dim[2]};
  int64_t s2 = t.stride[1];
  if (t.dim[2] == 1 && (s2 <= 0 || s2 % 8 != 0)) s2 = t.stride[0] * t.dim[1];
  FHB_ARG_CHECK(s2 % 8 == 0 && s2 > 0, "gemm: %s batch stride %lld must be a positive multiple of 8 elements", name,
                (long long)s2);
  cuuint64_t strides[2] = {(cuuint64_t)t.stride[0] * 2, (cuuint64_t)s2 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(t.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FHB_ARG_CHECK(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %lld,%lld,%lld strides %lld,%lld box %u,%u)",
                name, (int)r, (long long)t.dim[0], (long long)t.dim[1], (long long)t.dim[2], (long long)t.stride[0],
                (long long)s2, box0, box1);
  return 0;
}

int pick_bn(int n) {
  const int parts = (n + kMaxBN - 1) / kMaxBN;
  int bn = (n + parts - 1) / parts;
  bn = (bn + 15) / 16 * 16;
  return bn;
}

template <int A_MN, int B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    FHB_CUDA_CHECK(cudaFuncSetAttribute(fhb_gemm_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int grid = p.total_tiles < fhb_num_sms() ? p.total_tiles : fhb_num_sms();
  fhb_gemm_kernel<A_MN, B_MN><<<grid, kThreads, kSmemBytes, s>>>(ta, tb, p);
  FHB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int fhb_gemm(const fhb_gemm_args* a, fhb_stream_t stream) {
  FHB_ARG_CHECK(a != nullptr, "gemm: null args");
  FHB_ARG_CHECK(a->m > 0 && a->n > 0 && a->k > 0, "gemm: m,n,k must be positive (got %d,%d,%d)", a->m, a->n, a->k);
  FHB_ARG_CHECK(a->n % 16 == 0, "gemm: n=%d must be a multiple of 16", a->n);
  FHB_ARG_CHECK(a->a_major == 0 || a->a_major == 1, "gemm: a_major must be 0 or 1");
  FHB_ARG_CHECK(a->b_major == 0 || a->b_major == 1, "gemm: b_major must be 0 or 1");
  FHB_ARG_CHECK(!(a->a_major == 1 && a->b_major == 0), "gemm: (A MN-major, B K-major) is not instantiated");
  FHB_ARG_CHECK(a->d != nullptr && a->d_ld > 0, "gemm: null output / bad d_ld");
  const int num_ob = a->num_ob > 0 ? a->num_ob : 1;
  const int ob_mod = a->ob_mod > 0 ? a->ob_mod : 1;
  const int num_cb = a->num_cb > 0 ? a->num_cb : 1;
  FHB_ARG_CHECK(num_cb == 1 || a->b_major == 1, "gemm: num_cb > 1 needs MN-major operands");
  const int flags = a->flags;
  FHB_ARG_CHECK(!(flags & FHB_EPI_BIAS) || a->bias, "gemm: FHB_EPI_BIAS without bias");
  FHB_ARG_CHECK(!(flags & FHB_EPI_RESIDUAL) || a->residual, "gemm: FHB_EPI_RESIDUAL without residual");
  FHB_ARG_CHECK(!(flags & FHB_EPI_ROWZERO) || a->row_valid, "gemm: FHB_EPI_ROWZERO without row_valid");
  FHB_ARG_CHECK(!(flags & FHB_EPI_STORE_PREACT) || a->aux_out, "gemm: FHB_EPI_STORE_PREACT without aux_out");
  FHB_ARG_CHECK(!(flags & FHB_EPI_MUL_DGELU) || a->aux_in, "gemm: FHB_EPI_MUL_DGELU without aux_in");
  FHB_ARG_CHECK(!(flags & FHB_EPI_SQDIFF) || (a->loss_target && a->loss_acc), "gemm: FHB_EPI_SQDIFF without target/acc");
  FHB_ARG_CHECK(!(flags & FHB_EPI_ATOMIC_ADD) || (flags & FHB_EPI_OUT_F32), "gemm: atomic accumulate needs fp32 output");
  const int elt = (flags & FHB_EPI_OUT_F32) ? 4 : 2;
  FHB_ARG_CHECK((reinterpret_cast<uintptr_t>(a->d) & 15) == 0 && (a->d_ld * elt) % 16 == 0 &&
                    (a->d_hi_stride * elt) % 16 == 0 && (a->d_lo_stride * elt) % 16 == 0,
                "gemm: output must be 16B aligned in every stride");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m = a->m;
  p.n = a->n;
  p.k = a->k;
  p.bn = pick_bn(a->n);
  p.num_m_blk = (a->m + kBM - 1) / kBM;
  p.num_n_blk = (a->n + p.bn - 1) / p.bn;
  p.num_ob = num_ob;
  p.ob_mod = ob_mod;
  p.kb_per_cb = (a->k + kBK - 1) / kBK;
  p.kb_total = p.kb_per_cb * num_cb;
  int split = a->split_k;
  const int base_tiles = p.num_m_blk * p.num_n_blk * num_ob;
  if (split <= 0) {
    split = 1;
    if (flags & FHB_EPI_ATOMIC_ADD) {  // fill the machine when the output is small (wgrad)
      split = (2 * fhb_num_sms() + base_tiles - 1) / base_tiles;
      if (split > p.kb_total / 4) split = p.kb_total / 4;
      if (split < 1) split = 1;
    }
  }
  FHB_ARG_CHECK(split == 1 || (flags & FHB_EPI_ATOMIC_ADD), "gemm: split_k > 1 needs FHB_EPI_ATOMIC_ADD");
  if (split > p.kb_total) split = p.kb_total;
  p.kb_per_split = (p.kb_total + split - 1) / split;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.total_tiles = base_tiles * p.split_k;
  p.a_lo_c0 = a->a_lo_c0; p.a_hi_c2 = a->a_hi_c2; p.a_lo_c2 = a->a_lo_c2; p.a_cb_c2 = a->a_cb_c2;
  p.b_lo_c0 = a->b_lo_c0; p.b_hi_c2 = a->b_hi_c2; p.b_lo_c2 = a->b_lo_c2; p.b_cb_c2 = a->b_cb_c2;
  p.a_c1_off = a->a_c1_off; p.b_c1_off = a->b_c1_off;
  p.d_ld = a->d_ld; p.d_hi_stride = a->d_hi_stride; p.d_lo_stride = a->d_lo_stride;
  p.d = a->d;
  p.bias = a->bias;
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.aux_in = static_cast<const __nv_bfloat16*>(a->aux_in);
  p.aux_out = static_cast<__nv_bfloat16*>(a->aux_out);
  p.row_valid = a->row_valid;
  p.loss_target = static_cast<const __nv_bfloat16*>(a->loss_target);
  p.loss_acc = a->loss_acc;
  p.loss_weight = a->loss_weight;
  p.grad_scale = a->grad_scale;
  p.flags = flags;

  CUtensorMap ta, tb;
  int rc;
  const int b_atoms = (p.bn + 63) / 64;
  if (a->a_major == 0) {
    if ((rc = make_tmap(&ta, a->a, kBK, kBM, "A")) != 0) return rc;
  } else {
    if ((rc = make_tmap(&ta, a->a, 64, kBK, "A")) != 0) return rc;
  }
  if (a->b_major == 0) {
    if ((rc = make_tmap(&tb, a->b, kBK, (uint32_t)p.bn, "B")) != 0) return rc;
    p.stage_tx_bytes = kABytes + (uint32_t)p.bn * kBK * 2;
  } else {
    if ((rc = make_tmap(&tb, a->b, 64, kBK, "B")) != 0) return rc;
    p.stage_tx_bytes = kABytes + (uint32_t)b_atoms * 64 * kBK * 2;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->a_major == 0 && a->b_major == 0) return launch<0, 0>(ta, tb, p, s);
  if (a->a_major == 0 && a->b_major == 1) return launch<0, 1>(ta, tb, p, s);
  return launch<1, 1>(ta, tb, p, s);
}
