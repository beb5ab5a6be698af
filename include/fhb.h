/* libfhb_sm100a.so — C ABI of the B200-native FitHuBERT distillation hot path.
 *
 * The reference (glory20h/FitHuBERT) has NO operator / plugin / FFI layer: every GPU op is an
 * implicit PyTorch library call.  Each entry point below therefore cites the
 * reference *call site* (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions: plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 * a name ends in _host.  Every function enqueues work on `stream` and returns immediately:
 * 0 = ok, FHB_ERR_ARG (<0) = bad argument, >0 = cudaError_t.  fhb_last_error() gives the text.
 * No function allocates device memory or synchronises.  Activations are bf16 channel-last
 * [batch, time, channels]; statistics, losses, gradients of parameters and optimizer state fp32.
 */
#ifndef FHB_H_
#define FHB_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHB_ERR_ARG (-1)
typedef void* fhb_stream_t; /* cudaStream_t */

const char* fhb_last_error(void);
int fhb_abi_version(void);

/* ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA)
 * D[ob][m][n] = epilogue( sum_{cb,k} A[ob,cb][m][k] * B[ob,cb][n][k] )        bf16 x bf16 -> fp32
 * Replaces every cuBLAS / cuDNN contraction on the path: conv layers 1.. (modules/module.py:46,
 * 72-73 as strided-view GEMMs), post_extract_proj (modules/model.py:480-481),
 * grouped pos-conv (modules/module.py:186-200,276-277), time-reduction conv (:317-321), q/k/v/out
 * projections and fc1/fc2 (:557-579 -> fairseq MultiheadAttention), LayerWiseProjHead (:649-661),
 * and the dgrad / wgrad contractions autograd derives from them.
 *
 * Operands are 3-D strided bf16 tensors (dim[0] contiguous; strides in ELEMENTS, multiples of 8;
 * rows may overlap, i.e. stride[0] < dim[0] is legal and is how k>1 convolutions are expressed).
 *   major 0 (K-major):  dim = {K, rows(M or N), batches}
 *   major 1 (MN-major): dim = {M or N, contraction rows, contraction batches}
 */
typedef struct {
  const void* ptr;
  int64_t dim[3];
  int64_t stride[2];
} fhb_tensor3;

enum {
  FHB_EPI_BIAS = 1,         /* + bias[n] (fp32)                                              */
  FHB_EPI_GELU = 2,         /* exact erf GELU                                                */
  FHB_EPI_RESIDUAL = 4,     /* + residual[m][n] (bf16, laid out like D)                      */
  FHB_EPI_ROWZERO = 8,      /* rows m >= row_valid[ob_hi] are written as 0                   */
  FHB_EPI_STORE_PREACT = 16,/* aux_out[m][n] = value before GELU (bf16, laid out like D)     */
  FHB_EPI_MUL_DGELU = 32,   /* * gelu'(aux_in[m][n])  (backward through a fused GELU)        */
  FHB_EPI_OUT_F32 = 64,     /* D is fp32 (default bf16)                                      */
  FHB_EPI_ATOMIC_ADD = 128, /* D += (fp32 atomics; required when split_k > 1)                */
  FHB_EPI_SQDIFF = 256      /* fused distillation loss, see fhb_gemm_args.loss_*             */
};

typedef struct {
  fhb_tensor3 a, b;
  int32_t a_major, b_major;
  int32_t m, n, k;          /* per (ob, cb) problem; k = contraction length per cb          */
  int32_t num_ob;           /* output batches; ob = ob_hi * ob_mod + ob_lo                   */
  int32_t ob_mod;           /* >= 1                                                          */
  int32_t num_cb;           /* contraction batches (MN-major operands only; else 1)          */
  /* TMA coordinates: c0 += ob_lo*lo_c0 ; c2 = ob_hi*hi_c2 + ob_lo*lo_c2 + cb*cb_c2          */
  int32_t a_lo_c0, a_hi_c2, a_lo_c2, a_cb_c2;
  int32_t b_lo_c0, b_hi_c2, b_lo_c2, b_cb_c2;
  void* d;
  int64_t d_ld, d_hi_stride, d_lo_stride; /* elements */
  int32_t flags;
  int32_t split_k;          /* 0 = choose                                                    */
  const float* bias;
  const void* residual;
  const void* aux_in;
  void* aux_out;
  const int32_t* row_valid; /* [num_ob / ob_mod]                                             */
  /* FHB_EPI_SQDIFF: d_out = scale*(val - target) stored as D; loss_acc[0] += w*sum((val-t)^2) */
  const void* loss_target;  /* bf16, laid out like D                                         */
  float* loss_acc;
  float loss_weight, grad_scale;
} fhb_gemm_args;

int fhb_gemm(const fhb_gemm_args* args, fhb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FHB_H_ */
