// Shared device/host helpers for libfhb_sm100a.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fhb.h"

// ---------------------------------------------------------------- error plumbing
void fhb_set_error(const char* fmt, ...);
#define FHB_ARG_CHECK(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      fhb_set_error(__VA_ARGS__);                \
      return FHB_ERR_ARG;                        \
    }                                            \
  } while (0)
#define FHB_CUDA_CHECK(expr)                                                       \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      fhb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                              \
    }                                                                              \
  } while (0)
#define FHB_LAUNCH_CHECK() FHB_CUDA_CHECK(cudaGetLastError())

// memoised cuTensorMapEncodeTiled for a 3-D bf16 tensor (dim[0] contiguous, strides in elements), 128B swizzle
int fhb_make_tmap_bf16_3d(CUtensorMap* tm, const void* ptr, const int64_t dim[3], const int64_t stride[2],
                          uint32_t box0, uint32_t box1, const char* name);

int fhb_make_tmap_bf16_4d(CUtensorMap* tm, const void* ptr, const int64_t dim[4], const int64_t stride[3], uint32_t box0,
                          uint32_t box2, const char* name);
int fhb_reserved_sms();  // c_api_common.cu (fhb_set_reserved_sms)
// SM count of the CURRENT device (cached per device: one process may drive several GPUs) minus the reserved ones
static inline int fhb_num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cache[dev & 63];
  if (n == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  const int r = fhb_reserved_sms();
  return n - r > 8 ? n - r : n;
}

// Runs `stmt` the first time this call site is reached on each device (function attributes such as the dynamic
// shared-memory limit are per device, not per process).
#define FHB_ONCE_PER_DEVICE(stmt)                                                   \
  do {                                                                              \
    static unsigned long long _fhb_done = 0ull;                                     \
    int _fhb_dev = 0;                                                               \
    cudaGetDevice(&_fhb_dev);                                                       \
    const unsigned long long _fhb_bit = 1ull << (_fhb_dev & 63);                    \
    if (!(__atomic_load_n(&_fhb_done, __ATOMIC_ACQUIRE) & _fhb_bit)) {              \
      stmt;                                                                         \
      __atomic_fetch_or(&_fhb_done, _fhb_bit, __ATOMIC_RELEASE);                    \
    }                                                                               \
  } while (0)

#ifdef __CUDACC__
// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization may have their CTAs scheduled while the
// previous kernel of the stream is still draining, so the launch latency and the prologue (barrier init, TMEM
// allocation, tensor-map prefetch) overlap the predecessor's tail.  Contract: EVERY thread executes pdl_wait() before
// its first global-memory access and before any exit (it returns once the whole predecessor grid has completed and
// its writes are visible; a no-op for a normally launched kernel); pdl_trigger() right after it lets the successor's
// CTAs queue up behind this grid's last wave.
// Modes (fhb_set_pdl / FHB_PDL): 0 off, 1 every kernel, 2 (default) only launches the host code hints as a few
// microseconds long (the student's 12 448-row GEMMs, LayerNorms, column sums ...).  Interleaved A/B of the full step
// on B200 (profiles/r01y_pdl_ab.log): 23.82 ms off, 23.70 ms all, 23.56 ms small-only - for the long kernels the early
// co-residency costs more than the hidden latency is worth.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}
bool fhb_pdl_enabled();
void fhb_pdl_hint(bool small);
template <typename... KArgs, typename... Args>
static inline cudaError_t fhb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = fhb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ---------------------------------------------------------------- packed fp32x2 math (sm_100: FFMA2 / FADD2 / FMNMX3)
// Blackwell issues two fp32 FMAs / adds per instruction on an aligned register pair and a three-input max; the
// issue-bound inner loops (softmax, GELU epilogues) use them to halve their FFMA / FADD / FMNMX counts.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// ---------------------------------------------------------------- small math
// Exact (erf-based) GELU, written for issue-bound GEMM epilogues (3.5 G evaluations per distillation step on
// 8 epilogue warps per SM):  gelu(x) = max(x, 0) - |x| * T(|x|),  T(a) = 0.5 erfc(a / sqrt2) = 2^t(a), t a
// degree-4 polynomial without constant term (fitted offline with scipy: -ln erfc(z) / z as a cubic in z,
// folded with the 1/sqrt2, log2 e and 0.5 factors).  |gelu error| <= 1.3e-5, |gelu' error| <= 4.3e-5 over all
// x (bf16 resolution near 1 is 4e-3).  gelu: FMNMX, 4 FFMA, MUFU.EX2, FMNMX, FFMA = 8 instructions;
// gelu' (no second MUFU: see gelu_grad_from).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// a = min(|x|, 6 sqrt2) -> 0.5 * erfc(a / sqrt2)
__device__ __forceinline__ float gelu_tail(float a) {
  float t = fmaf(4.389107953e-03f, a, -4.677256044e-02f);
  t = fmaf(t, a, -4.635094653e-01f);
  t = fmaf(t, a, -1.150136897e+00f);
  t = fmaf(t, a, -1.0f);
  return ex2_approx(t);
}
__device__ __forceinline__ float gelu_abs_clamp(float x) { return fminf(fabsf(x), 8.4852814f); }
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = gelu_abs_clamp(x);
  return fmaf(-a, gelu_tail(a), fmaxf(x, 0.f));
}
// gelu'(x) = step(x) + sign(x) * (|x| pdf(|x|) - T(|x|)) with pdf = -T', i.e. the derivative of the SAME approximation the
// forward evaluates: T' = ln2 t'(a) T, so u = T(a) (1 + a ln2 t'(a)) and gelu' = x >= 0 ? 1 - u : u.  A cubic (3 FFMA)
// in place of a second MUFU.EX2: the GELU + gelu' epilogues of the conv stack are bound by their special-function issue.
// |gelu' error| <= 4.3e-5 over all x (fp16 resolution near 1 is 4.9e-4).
__device__ __forceinline__ float gelu_grad_from(float x, float a, float tail) {
  float d = fmaf(1.216919121e-02f, a, -9.726080519e-02f);
  d = fmaf(d, a, -6.425605581e-01f);
  d = fmaf(d, a, -7.972141474e-01f);                              // ln2 t'(a)
  const float u = tail * fmaf(a, d, 1.0f);
  return x >= 0.f ? 1.0f - u : u;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float a = gelu_abs_clamp(x);
  return gelu_grad_from(x, a, gelu_tail(a));
}
// y = gelu(x), g = gelu'(x) sharing the tail evaluation
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& g) {
  const float a = gelu_abs_clamp(x);
  const float tail = gelu_tail(a);
  y = fmaf(-a, tail, fmaxf(x, 0.f));
  g = gelu_grad_from(x, a, tail);
}
// Two GELUs at once: the degree-4 polynomial runs as 4 FFMA2 instead of 8 FFMA (6 instead of 8 issue slots per value)
__device__ __forceinline__ f32x2_t gelu_tail_poly2(f32x2_t a) {
  f32x2_t t = fma2(pack2(4.389107953e-03f, 4.389107953e-03f), a, pack2(-4.677256044e-02f, -4.677256044e-02f));
  t = fma2(t, a, pack2(-4.635094653e-01f, -4.635094653e-01f));
  t = fma2(t, a, pack2(-1.150136897e+00f, -1.150136897e+00f));
  return fma2(t, a, pack2(-1.0f, -1.0f));
}
__device__ __forceinline__ void gelu_erf2(float x0, float x1, float& y0, float& y1) {
  const float a0 = gelu_abs_clamp(x0), a1 = gelu_abs_clamp(x1);
  float t0, t1;
  unpack2(gelu_tail_poly2(pack2(a0, a1)), t0, t1);
  y0 = fmaf(-a0, ex2_approx(t0), fmaxf(x0, 0.f));
  y1 = fmaf(-a1, ex2_approx(t1), fmaxf(x1, 0.f));
}
__device__ __forceinline__ void gelu_erf_both2(float x0, float x1, float& y0, float& y1, float& g0, float& g1) {
  const float a0 = gelu_abs_clamp(x0), a1 = gelu_abs_clamp(x1);
  float t0, t1;
  unpack2(gelu_tail_poly2(pack2(a0, a1)), t0, t1);
  const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
  y0 = fmaf(-a0, e0, fmaxf(x0, 0.f));
  y1 = fmaf(-a1, e1, fmaxf(x1, 0.f));
  // gelu_grad_from on the pair: the cubic ln2 t'(a) and the product with the tail as FFMA2 / FMUL2
  const f32x2_t a2 = pack2(a0, a1);
  f32x2_t d2 = fma2(pack2(1.216919121e-02f, 1.216919121e-02f), a2, pack2(-9.726080519e-02f, -9.726080519e-02f));
  d2 = fma2(d2, a2, pack2(-6.425605581e-01f, -6.425605581e-01f));
  d2 = fma2(d2, a2, pack2(-7.972141474e-01f, -7.972141474e-01f));
  float u0, u1;
  unpack2(mul2(pack2(e0, e1), fma2(a2, d2, pack2(1.0f, 1.0f))), u0, u1);
  g0 = x0 >= 0.f ? 1.0f - u0 : u0;
  g1 = x1 >= 0.f ? 1.0f - u1 : u1;
}
// ---------------------------------------------------------------- dropout (K13)
// Counter-based masks: forward and backward kernels regenerate the same bits from (seed, element index), so no
// mask tensor is ever stored.  One 32-bit hash (murmur3 finaliser) serves the element pair (2i, 2i+1): each
// element keeps iff its 16-bit half >= thr16 = round(p * 65536).  nn.Dropout semantics: kept values are
// scaled by 1 / (1 - p).  (torch's Philox stream cannot be reproduced; parity runs use p = 0)
__device__ __forceinline__ uint32_t fhb_hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t dropout_pair_bits(uint32_t seed, uint32_t pair_idx) {
  return fhb_hash32(pair_idx * 0x9E3779B1u + seed);
}
// multipliers (0 or scale) for elements 2*pair_idx and 2*pair_idx + 1
__device__ __forceinline__ void dropout_pair(uint32_t seed, uint32_t pair_idx, uint32_t thr16, float scale, float& m0,
                                             float& m1) {
  const uint32_t bits = dropout_pair_bits(seed, pair_idx);
  m0 = (bits & 0xFFFFu) >= thr16 ? scale : 0.f;
  m1 = (bits >> 16) >= thr16 ? scale : 0.f;
}
__device__ __forceinline__ float dropout_one(uint32_t seed, uint32_t idx, uint32_t thr16, float scale) {
  const uint32_t bits = dropout_pair_bits(seed, idx >> 1);
  return ((idx & 1u) ? (bits >> 16) : (bits & 0xFFFFu)) >= thr16 ? scale : 0.f;
}
#endif  // __CUDACC__
static inline uint32_t fhb_dropout_thr16(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)t);
}
static inline float fhb_dropout_scale(float p) { return 1.0f / (1.0f - (float)fhb_dropout_thr16(p) / 65536.0f); }
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// ---------------------------------------------------------------- 16-bit storage formats
// Every 16-bit tensor of the library is fp16: weights, activations, saved gelu' multipliers, attention probabilities,
// projections, teacher targets - all bounded - and the gradients, which carry a power-of-two LOSS SCALE (applied by the
// loss kernel, removed by the AdamW kernel) like the reference's own AMP recipe (data/conf/fithubert.yaml: use_fp16).
// fp16's 11-bit significand is 8x finer than bf16's 8 bits at the same tensor-pipe rate and the same bytes.  (tcgen05
// kind::f16 has separate format fields for A and B, but mixing fp16 with bf16 raises an illegal-instruction fault on
// sm_100a, so one format has to serve forward AND backward.)  Conversions saturate instead of producing inf.
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool f16) { return f16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
__device__ __forceinline__ float2 unpack16(uint32_t u, bool f16) { return f16 ? unpack_f16(u) : unpack_bf16(u); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- PTX: mbarrier / TMA / tcgen05
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* desc, uint64_t* bar, void* smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const void* desc, uint64_t* bar, uint32_t smem_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_addr), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100): SWIZZLE_128B, version 1.
//   bits [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor, kind::f16 -> fp32 accumulate.  a_bf16 / b_bf16: operand format (0 = fp16, 1 = bf16).
__host__ __device__ constexpr uint32_t umma_idesc_16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major,
                                                     uint32_t a_bf16, uint32_t b_bf16) {
  return (1u << 4)                 // D format fp32
         | (a_bf16 << 7)           // A format: 0 fp16, 1 bf16
         | (b_bf16 << 10)          // B format
         | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return umma_idesc_16(M, N, a_mn_major, b_mn_major, 1, 1);
}
#endif  // __CUDACC__
