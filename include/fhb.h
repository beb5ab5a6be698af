/* libfhb_sm100a.so — C ABI of the B200-native FitHuBERT distillation hot path.
 *
 * The reference (glory20h/FitHuBERT) has NO operator / plugin / FFI layer: every GPU op is an
 * implicit PyTorch library call.  Each entry point below therefore cites the
 * reference *call site* (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions: plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 * a name ends in _host.  Every function enqueues work on `stream` and returns immediately:
 * 0 = ok, FHB_ERR_ARG (<0) = bad argument, >0 = cudaError_t.  fhb_last_error() gives the text.
 * No function allocates device memory or synchronises.  Every 16-bit tensor is fp16, channel-last
 * [batch, time, channels]: weight shadows, activations, saved gelu' multipliers, projections, teacher targets, and the
 * gradients, which carry a power-of-two LOSS SCALE (the caller multiplies it into the grad_scale of the loss kernel
 * and divides it out in the grad_scale of fhb_adamw_multi) - the precision the reference's own AMP recipe trains in
 * (data/conf/fithubert.yaml: use_fp16; Lightning GradScaler).  fp16's significand is 8x finer than bf16's at the
 * same tensor-pipe rate; tcgen05 kind::f16 cannot mix the two formats in one product (illegal instruction on
 * sm_100a), so one format serves forward and backward.  Conversions to fp16 saturate.  The residual stream of the
 * transformer layers and its gradient are fp32.  Statistics, losses, parameter gradients, optimizer state: fp32.
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
/* Programmatic dependent launch.  mode 0 off; 1 every kernel; 2 (default) only kernels the library knows to be a few
 * microseconds long (FHB_PDL=0 / 1 / 2 in the environment select the initial mode).  A kernel launched this way
 * overlaps its prologue with the previous kernel's tail and waits (griddepcontrol.wait) before its first global-memory
 * access.  Returns the previous mode.  Per-kernel event timing switches it off. */
int fhb_set_pdl(int mode);
/* Leave `n` SMs (0 ... 64) out of the grids of the persistent kernels (GEMM, LayerNorm, ...): room for the few CTAs of a
 * gradient all-reduce that overlaps the backward pass (Lightning DDP's overlap, train.py:494).  Returns the previous n. */
int fhb_set_reserved_sms(int n);

/* ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA)
 * D[ob][m][n] = epilogue( sum_{cb,k} A[ob,cb][m][k] * B[ob,cb][n][k] )        bf16 x bf16 -> fp32
 * Replaces every cuBLAS / cuDNN contraction on the path: conv layers 1.. (modules/module.py:46,
 * 72-73 as strided-view GEMMs), post_extract_proj (modules/model.py:480-481),
 * grouped pos-conv (modules/module.py:186-200,276-277), time-reduction conv (:317-321), q/k/v/out
 * projections and fc1/fc2 (:557-579 -> fairseq MultiheadAttention), LayerWiseProjHead (:649-661),
 * and the dgrad / wgrad contractions autograd derives from them.
 *
 * Operands are 3-D strided fp16 tensors (FHB_GEMM_A_BF16 + FHB_GEMM_B_BF16 together select bf16 x bf16; the two
 * formats cannot be mixed; dim[0] contiguous; strides in ELEMENTS, multiples of 8;
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
  FHB_EPI_RESIDUAL = 4,     /* + residual[m][n] (laid out like D; fp16, or see RES_BF16/F32) */
  FHB_EPI_ROWZERO = 8,      /* rows m >= row_valid[ob_hi] are written as 0                   */
  FHB_EPI_STORE_PREACT = 16,/* aux_out[m][n] = value before GELU (D's 16-bit type, like D)   */
  FHB_EPI_MUL_DGELU = 32,   /* * gelu'(aux_in[m][n])  (aux_in fp16: a saved pre-activation)  */
  FHB_EPI_OUT_F32 = 64,     /* D is fp32 (default fp16; FHB_EPI_OUT_BF16: bf16, a gradient)  */
  FHB_EPI_ATOMIC_ADD = 128, /* D += (fp32 atomics; required when split_k > 1)                */
  FHB_EPI_SQDIFF = 256,     /* fused distillation loss, see fhb_gemm_args.loss_*             */
  FHB_EPI_AUX_DGELU = 512,  /* with STORE_PREACT: aux_out = gelu'(value before GELU), ALWAYS fp16 */
  FHB_EPI_MUL_AUX = 1024,   /* * aux_in[m][n] (fp16: a gelu' saved by FHB_EPI_AUX_DGELU)      */
  FHB_EPI_DROPOUT = 2048,   /* nn.Dropout(drop_p) after bias/GELU, before the residual; the    */
                            /* mask of element (ob, m, n) is a hash of (drop_seed, index), see  */
                            /* fhb_dropout; with AUX_DGELU the saved gelu' is masked the same   */
  FHB_EPI_RES_F32 = 4096,   /* the residual is fp32 (the high-precision copy of the residual    */
                            /* stream the LayerNorm kernels keep next to the 16-bit GEMM operand) */
  FHB_GEMM_A_BF16 = 8192,   /* A holds bf16; default fp16 (set both or neither)                */
  FHB_GEMM_B_BF16 = 16384,  /* B holds bf16; default fp16                                      */
  FHB_EPI_OUT_BF16 = 32768, /* D (and a plain STORE_PREACT aux_out) is bf16                    */
  FHB_EPI_RES_BF16 = 65536, /* the residual is bf16                                            */
  FHB_EPI_ALPHA = 131072    /* accumulator x alpha first (keeps sums over many rows, e.g. the  */
                            /* gradient of a folded weight, inside fp16's range)               */
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
  int32_t a_c1_off, b_c1_off; /* added to the contraction-row coordinate of MN-major operands */
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
  int64_t bias_hi_stride;   /* elements: batch ob_hi reads bias + ob_hi * bias_hi_stride (0 = shared bias)  */
  uint32_t drop_seed;       /* FHB_EPI_DROPOUT                                                 */
  float drop_p;
  float alpha;              /* FHB_EPI_ALPHA                                                   */
} fhb_gemm_args;

int fhb_gemm(const fhb_gemm_args* args, fhb_stream_t stream);

/* ------------------------------------------------------------------ conv0 + GroupNorm + GELU (K1)
 * out[b][t][c] = gelu( GN_c( sum_j w[c][j] * wave[b][5t+j] ) ),  GroupNorm with C groups over all T0
 * frames (zero padding included).  Replaces modules/module.py:46,65-71 (layer 0 of
 * ConvFeatureExtractionModel: nn.Conv1d -> Fp32GroupNorm(dim, dim) -> nn.GELU), teacher and student.
 * fwd writes stat/mean/rstd (needed by bwd).  bwd produces dweight/dgamma/dbeta only (the waveform
 * has no gradient, SURVEY App. F backward inventory). */
typedef struct {
  const float* wave;      /* [B][wave_ld] fp32, zero padded                                   */
  int64_t wave_ld;
  int32_t B, L, C, T0, kernel, stride;
  float eps;
  const float* weight;    /* [C][kernel] fp32                                                  */
  const float* gamma;     /* [C]                                                               */
  const float* beta;      /* [C]                                                               */
  double* stat;           /* [B][65] workspace: S_j and packed R_jj' (fwd writes, bwd reads)   */
  float* mean;            /* [B][C]                                                            */
  float* rstd;            /* [B][C]                                                            */
  void* out;              /* fwd: bf16 [B][T0][C]                                              */
  const void* dy;         /* bwd: bf16 [B][T0][C]                                              */
  float* acc;             /* bwd workspace [B][C][12]                                          */
  float* dweight;         /* bwd: [C][kernel]                                                  */
  float* dgamma;          /* bwd: [C]                                                          */
  float* dbeta;           /* bwd: [C]                                                          */
  int32_t accumulate;     /* bwd: add into dweight/dgamma/dbeta instead of overwriting         */
  void* gp_out;           /* fwd, optional: bf16 [B][T0][C] gelu'(GroupNorm output), the backward multiplier */
  int32_t dy_is_dz;       /* bwd: dy was already multiplied by gp_out (e.g. FHB_EPI_MUL_AUX of the next
                             layer's dgrad GEMM): no conv / GroupNorm / gelu' recompute                */
} fhb_conv0_args;
int fhb_conv0_gn_gelu_fwd(const fhb_conv0_args* args, fhb_stream_t stream);
int fhb_conv0_gn_gelu_bwd(const fhb_conv0_args* args, fhb_stream_t stream);
/* Training-path backward of layer 0 on the tensor pipe.  fhb_conv0_im2col writes xcol bf16 [B][T0][32]:
 * columns 0-9 = bf16(x[5t+j]), 10 = 1, 16-25 = the bf16 remainder x - hi, the rest 0.  A wgrad-shaped fhb_gemm
 * (A = dz [B][T0][C] MN-major, B = xcol MN-major, one output block [C][32] fp32 per sample, split-K accumulate) then
 * yields sum_t dz x[5t+j] (hi + lo columns) and sum_t dz (column 10); fhb_conv0_bwd_finalize turns those
 * accumulators acc32 [B][C][32] into dW / dgamma / dbeta (args: weight, gamma, stat, mean, rstd, dweight, dgamma,
 * dbeta, accumulate, B, C, T0 as in fhb_conv0_gn_gelu_bwd; dz = dy * gelu'(z) as saved by the forward). */
int fhb_conv0_im2col(const float* wave, int64_t wave_ld, int32_t B, int32_t L, int32_t T0, void* xcol,
                     fhb_stream_t stream);
int fhb_conv0_bwd_finalize(const float* acc32, const fhb_conv0_args* a, fhb_stream_t stream);

/* ------------------------------------------------------------------ LayerNorm (K3)
 * y = LN(x) * gamma + beta over the last dim C (eps 1e-5), rows = everything else; bf16 in/out, fp32
 * statistics.  Replaces nn.LayerNorm at modules/model.py:446-447 and modules/module.py:251,281,513,
 * 518,569,580.  fwd optionally saves mean/rstd; bwd: dx (bf16, optionally + dres), dgamma/dbeta
 * (fp32, atomically ACCUMULATED: zero them first). */
int fhb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int64_t rows, int32_t C, float eps, fhb_stream_t stream);
int fhb_layernorm_bwd(const void* dy, const void* dy2 /* optional: gradient = dy + dy2 */, const void* x,
                      const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, void* dx_drop,
                      uint32_t drop_seed, float drop_p, int64_t rows, int32_t C, fhb_stream_t stream);
/* The residual stream of the post-LN transformer layers (modules/module.py:557-580: x = LN(x + branch(x))) is carried in
 * fp32 next to the bf16 GEMM operands: the out_proj / fc2 epilogues add an fp32 residual (FHB_EPI_RES_F32) and write the
 * sum in fp32 (FHB_EPI_OUT_F32); these variants read that sum.
 *  fwd32: x32 fp32 [rows][C] -> y (bf16, the next GEMM's A operand), y32 (optional fp32 copy, the next residual operand);
 *         sub32 + diff_out (optional, together): diff_out = bf16(x32 - sub32) = the FFN branch output `layer_result`
 *         the reference returns per layer (modules/module.py:577-580, modules/model.py:493-499).
 *  bwd32: gradient = dy32 (fp32, optional) + dy2 (bf16, optional), at least one; x32 = the saved fp32 sum; outputs
 *         dx32 (optional fp32, continues down the residual), dx (optional bf16), dx_drop (optional bf16, dx * the
 *         dropout mask of (drop_seed, drop_p)); dgamma / dbeta / dxsum ACCUMULATED as in fhb_layernorm_bwd (dxsum sums
 *         dx_drop when given, else dx). */
int fhb_layernorm_fwd32(const float* x32, const float* gamma, const float* beta, void* y, float* y32, float* mean,
                        float* rstd, const float* sub32, void* diff_out, int64_t rows, int32_t C, float eps,
                        fhb_stream_t stream);
int fhb_layernorm_bwd32(const float* dy32, const void* dy2, const float* x32, const float* gamma, const float* mean,
                        const float* rstd, void* dx, float* dx32, float* dgamma, float* dbeta, float* dxsum,
                        void* dx_drop, uint32_t drop_seed, float drop_p, int64_t rows, int32_t C, fhb_stream_t stream);

/* ------------------------------------------------------------------ positional conv (K5) helpers
 * The grouped Conv1d(k=128, pad=64, groups=16) of modules/module.py:186-200,276-278 runs as a batched
 * fhb_gemm over a group-major, time-padded copy of x.  These kernels do the layout work around it:
 *  pack:    xg[b][g][t+pad][cp] = (t < valid[b]) ? x[b][t][g*cg + c] : 0   (zeroes padded frames =
 *           index_put at :273-274; cp = cg rounded up to 16; borders and channel padding zero)
 *  wn_prep: w = g * v / ||v||_(0,1) (weight_norm dim=2, :199) -> bf16 [G][cp_out][k*cp_in]
 *  finish:  h = xz + gelu(conv[..][:cg]) ; y = LN(h)  (SamePad + GELU + residual :276-278, LN :280-281)
 * and their backward counterparts.
 * Time blocking: with delta > 1 one GEMM row produces delta consecutive frames (N = delta * cp instead of a
 * 30-48 column sliver): wn_prep then writes delta shifted weight copies W'[g][(dl, n)][(j + dl, k)] over K + delta
 * taps into a buffer the caller zeroed once, and the GEMM output / dgrad output is laid out
 * [b][ceil(T / delta)][g][delta][cp] (what finish_fwd / finish_bwd / unpack_bwd index when given delta). */
int fhb_posconv_pack(const void* x, const int32_t* valid, void* xg, int32_t B, int32_t T, int32_t C, int32_t G,
                     int32_t cp, int32_t pad_l, int32_t Tp, fhb_stream_t stream);
/* wn_prep writes the forward operand w_fwd = W[g][co][(j, ci)] and (if non-NULL) the dgrad operand
 * w_bwd = W[g][ci][(K - 1 - j, co)] in one call.  ws: 2*K floats of workspace; on return ws[K..2K) holds
 * 1 / ||v[:, :, j]|| (the inv_norm argument of fhb_posconv_wn_bwd).  Group width C / G must be even, K | 256. */
int fhb_posconv_wn_prep(const float* v, const float* g, void* w_fwd, void* w_bwd, float* ws, int32_t C, int32_t G,
                        int32_t K, int32_t cp, int32_t delta, fhb_stream_t stream);
/* y32 (optional): fp32 copy of y, the residual operand of the first transformer layer (see fhb_layernorm_fwd32) */
int fhb_posconv_finish_fwd(const void* x, const int32_t* valid, const void* conv, const float* bias,
                           const float* gamma, const float* beta, void* h_out, void* y, float* y32, float* mean,
                           float* rstd, int32_t B, int32_t T, int32_t C, int32_t G, int32_t cp, float eps, int32_t delta,
                           fhb_stream_t stream);
/* dh = LN-bwd(dy) and dcg[b][g][t+pad_l][cp] = dh * gelu'(conv + bias) (group-major, time-padded: the
 * A operand of the dgrad GEMM and the B operand of the wgrad GEMM) in one pass; dgamma/dbeta/dbias are
 * ACCUMULATED atomically. */
int fhb_posconv_finish_bwd(const void* dy, const void* h, const void* conv, const float* bias, const float* gamma,
                           const float* mean, const float* rstd, void* dh, void* dcg, float* dgamma, float* dbeta,
                           float* dbias, int32_t B, int32_t T, int32_t C, int32_t G, int32_t cp, int32_t pad_l,
                           int32_t Tp, int32_t delta, fhb_stream_t stream);
/* dx[b][t][c] = (t < valid[b]) ? dh[b][t][c] + dxc[b][t][g][c'] : 0   (dxc = dgrad GEMM output) */
int fhb_posconv_unpack_bwd(const void* dh, const void* dxc, const int32_t* valid, void* dx, int32_t B, int32_t T,
                           int32_t C, int32_t G, int32_t cp, int32_t delta, fhb_stream_t stream);
/* dv, dg from dwt (fp32 [G][K*cp][cp], the wgrad GEMM output: dW[g*cg+co][ci][j] = dwt[g][j*cp+ci][co]).
 * delta > 1: dwt is the time-blocked wgrad output [G][(K+delta)*cp][delta*cp] and
 * dW[g*cg+co][ci][j] = sum_dl dwt[g][(j+dl)*cp+ci][dl*cp+co].
 * ws: the 2*K-float workspace fhb_posconv_wn_prep filled: ws[K..2K) = 1 / ||v[:, :, j]|| is read, ws[0..K) is
 * overwritten (scratch for the per-tap dot products sum dW * v). */
int fhb_posconv_wn_bwd(const float* dwt, const float* v, const float* g, float* ws, float* dv, float* dg, int32_t C,
                       int32_t G, int32_t K, int32_t cp, int32_t accumulate, int32_t delta, fhb_stream_t stream);

/* ------------------------------------------------------------------ masked attention (K7)
 * o = softmax(scale * q k^T + keymask) v per (sample, head); replaces fairseq MultiheadAttention's
 * bmm -> masked_fill(-inf) -> fp32 softmax -> bmm chain reached from modules/module.py:558-564.
 * qkv: bf16 [B][T][3*H*d] (q | k | v column blocks, as written by the fused QKV GEMM); keys at
 * t >= valid[b] are masked; padded QUERY rows are still computed.  lse: fp32 [B][H][T].
 * d in {40 (student), 64 (teacher)} run on tcgen05 (forward); any multiple of 8 <= 64 is supported.
 * drop_p > 0: attention dropout (fairseq MultiheadAttention dropout_module on the probabilities) with the
 * counter-based mask of fhb_dropout over index ((b*H + h)*T + q) * 2*ceil(T/2) + k; backward regenerates it. */
int fhb_attn_fwd(const void* qkv, const int32_t* valid, void* out, float* lse, int32_t B, int32_t T, int32_t H,
                 int32_t d, float scale, uint32_t drop_seed, float drop_p, fhb_stream_t stream);
/* delta_ws: fp32 [B][H][T] workspace.  dq_ws: fp32 [B][T][H*d] workspace (optional): with it, d in {40, 64}
 * runs the fused tcgen05 backward (dQ partials of every key tile are TMA-reduce-added into dq_ws, then converted);
 * without it the two-kernel mma.sync backward is used. */
int fhb_attn_bwd(const void* qkv, const int32_t* valid, const void* out, const void* dout, const float* lse,
                 void* dqkv, float* delta_ws, float* dq_ws, int32_t B, int32_t T, int32_t H, int32_t d, float scale,
                 uint32_t drop_seed, float drop_p, fhb_stream_t stream);

/* ------------------------------------------------------------------ attention-map / value-relation distillation
 * The reference's optional recipe (attn_loss_weight / v_rel_loss_weight > 0): train.py:64-77 re-binds every encoder layer's
 * forward to utils/utils.py:190-258 `rtrn_attn_forward`, which asks fairseq's MultiheadAttention for its un-normalised
 * logits (before_softmax=True: bmm(q * scaling, k^T) with -inf at padded keys) and also returns the value relation
 * v_rel = bmm(v * scaling, v^T); train.py:327-368 compares the LAST layer's maps of student and teacher.
 * Only that layer materialises T x T maps here: fp32 [B*H][T][pitch], pitch a multiple of 8 (>= T).
 *
 * fhb_attn_scores: out[b*H+h][i][j] = scale * sum_c a[(b*T+i)*ld + h*d + c] * b[(b*T+j)*ld + h*d + c], -inf where
 *   j >= valid[b] (valid optional).  a / b: fp16 head blocks of the fused [B*T][3*H*d] projection output (q and k for the
 *   logits; v and v, valid = NULL, for the value relation).  d: multiple of 8, <= 64. */
int fhb_attn_scores(const void* a, const void* b, int64_t ld, const int32_t* valid, float* out, int64_t pitch,
                    int32_t B, int32_t T, int32_t H, int32_t d, float scale, fhb_stream_t stream);
/* loss[0] += loss_mult * sum over the B*H*T query rows of
 *   mode 0 (train.py:331-341): sum_{j < min(vs, vt)} (s - t)^2  - keys either side masks are inf / nan in the reference
 *           and dropped from sum and count (the caller passes loss_mult = 1 / (H * T * sum_b min(vs_b, vt_b)));
 *   mode 1 (train.py:342-349,357-364): sum_j p_j (log p_j - log q_j), p = softmax_j(t), q = softmax_j(s) over each side's
 *           un-masked keys, summed over the keys both keep (loss_mult = 1 / (B*H*T): F.kl_div(..).sum(-1).mean()).
 *           A key masked on BOTH sides is nan in the reference (0 * -inf, only inf is patched); here it contributes 0.
 * ds: fp16 [B*H][T][pitch], grad_mult * d(row term)/d(s), every column written (0 at masked keys / row padding);
 * grad_mult carries loss_mult, the loss weight and the loss scale of the fp16 gradients. */
int fhb_attn_map_loss(const float* s, const float* t, int64_t pitch, const int32_t* valid_s, const int32_t* valid_t,
                      void* ds, float* loss, int32_t B, int32_t T, int32_t H, int32_t mode, float loss_mult,
                      float grad_mult, fhb_stream_t stream);
/* Gradient of fhb_attn_scores back into a head block:
 *   trans = 0: out[(b*T+r)*ld_out + h*d + :] (+)= alpha * sum_c g[b*H+h][r][c] * m[(b*T+c)*ld_m + h*d + :]
 *   trans = 1: the same with g[b*H+h][c][r]
 * (dq = dS k, dk = dS^T q, dv = (dR + dR^T) v; alpha = the score scale, divided by any power of two the caller folded
 * into g to centre it in fp16's range).  g: fp16 as written by fhb_attn_map_loss; out: fp16, accumulate != 0 adds to it. */
int fhb_attn_scores_bwd(const void* g, int64_t pitch, const void* m, int64_t ld_m, void* out, int64_t ld_out, int32_t B,
                        int32_t T, int32_t H, int32_t d, float alpha, int32_t trans, int32_t accumulate,
                        fhb_stream_t stream);

/* ------------------------------------------------------------------ distillation loss + gradient (K10)
 * loss_l = w_l * mean_{b,t,d} (pred_l - tgt_l)^2 ; dpred_l = 2 w_l (pred_l - tgt_l) / (B*T'*D)
 * Replaces train.py:250-267 (two torch.stack copies), :282-293 (mse, weighting, mean).  pred: bf16
 * [n_layers][B][Tp][D] (Tp = T' frames); tgt: bf16 [n_layers][B][Tt][D] with Tt >= Tp (narrow, :282).
 * layer_loss (fp32 [n_layers]) is ACCUMULATED: zero it first.  dpred may alias pred.  dbias (optional):
 * dbias[l * dbias_layer_stride + d] += sum_{b,t} dpred_l[b][t][d], the gradient of the bias of the Linear that
 * produced pred_l (LayerWiseProjHead.lin_proj, modules/module.py:661). */
int fhb_distill_loss_fwd_bwd(const void* pred, const void* tgt, const float* weights, float* layer_loss,
                             void* dpred, float* dbias, int64_t dbias_layer_stride, int32_t n_layers, int32_t B,
                             int32_t Tp, int32_t Tt, int32_t D, int32_t loss_type /*0 mse, 1 l1*/, float grad_scale,
                             fhb_stream_t stream);
/* Same with the cosine term of train.py:302-314 (sim_loss_weight > 0; only the `distil_random_layer == 0` branch of the
 * reference can execute): per row of D features sim = -logsigmoid(cos(pred, tgt)), cos with F.cosine_similarity's
 * per-norm eps = 1e-8.  rec_layer_loss[l] += w_l * mean_{b,t,d} rec, sim_layer_loss[l] += w_l * mean_{b,t} sim,
 * dpred = rec_grad_scale * w_l * d mean(rec)/dp + sim_grad_scale * w_l * d mean(sim)/dp.  D <= 1024. */
int fhb_distill_loss_sim_fwd_bwd(const void* pred, const void* tgt, const float* weights, float* rec_layer_loss,
                                 float* sim_layer_loss, void* dpred, float* dbias, int64_t dbias_layer_stride,
                                 int32_t n_layers, int32_t B, int32_t Tp, int32_t Tt, int32_t D, int32_t loss_type,
                                 float rec_grad_scale, float sim_grad_scale, fhb_stream_t stream);

/* ------------------------------------------------------------------ fused AdamW (K11)
 * ONE launch over a device-resident table of tensors.  Replaces s3prl get_optimizer ->
 * Lamb(adam=True, correct_bias=True) reached from train.py:416-420 (mode 0; source
 * not available here: "parity unpinned") or torch.optim.AdamW semantics (mode 1).  Gradients are read
 * through a 3-D stride so they may stay in the layout the wgrad GEMM produced them in
 * (p is contiguous [dim0][dim1][dim2]; g index = i0*gstride0 + i1*gstride1 + i2*gstride2).
 * Entries with g == NULL are skipped (the reference skips parameters whose .grad is None). */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
  int64_t dim[3];
  int64_t gstride[3];
} fhb_adamw_tensor;
int fhb_adamw_multi(const fhb_adamw_tensor* table_dev, int32_t n_tensors, int64_t max_n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, int32_t mode, float grad_scale,
                    fhb_stream_t stream);

/* ------------------------------------------------------------------ weight preparation
 * ONE launch: dst (contiguous [dim0][dim1][dim2], bf16 or fp32) = src_fp32[i0*s0 + i1*s1 + i2*s2].
 * Produces the bf16 GEMM-layout shadows of the fp32 master parameters (conv weights tap-major,
 * ConvTranspose weights as [2*Cout][Cin], fused q|k|v) and, with dst_is_f32, strided fp32 copies
 * (concatenated biases; gradients exported back to parameter layout, optionally accumulated). */
typedef struct {
  const float* src;
  void* dst;
  int64_t dim[3];
  int64_t sstride[3];
  int32_t dst_is_f32;
  int32_t accumulate;
} fhb_prep_tensor;
int fhb_prep_multi(const fhb_prep_tensor* table_dev, int32_t n_tensors, int64_t max_n, fhb_stream_t stream);

/* ------------------------------------------------------------------ small helpers */
/* out[c] += sum_rows x[row][c]   (bias gradients); x bf16 [rows][ld], out fp32 accumulated atomically */
int fhb_colsum(const void* x, int64_t rows, int32_t C, int64_t ld, float* out, fhb_stream_t stream);
/* the same over `batches` matrices: x + b * x_bstride (elements) -> out + b * out_bstride */
int fhb_colsum_batched(const void* x, int64_t rows, int32_t C, int64_t ld, int64_t x_bstride, float* out,
                       int64_t out_bstride, int32_t batches, fhb_stream_t stream);
/* y = a + b (bf16, n % 8 == 0), used where two gradient streams meet */
int fhb_add_bf16(const void* a, const void* b, void* y, int64_t n, fhb_stream_t stream);
/* out = dy * gelu'(u) over B segments of n bf16 elements (independent batch strides); the one place a
 * GELU derivative is not a GEMM epilogue: LayerNorm(512)-backward -> last conv layer (module.py:73) */
int fhb_mul_dgelu(const void* dy, int64_t dy_bstride, const void* u, int64_t u_bstride, void* out,
                  int64_t out_bstride, int32_t B, int64_t n, fhb_stream_t stream);
/* Bias gradients of the n LayerWiseProjHeads (modules/module.py:649-661) from the column sums of dpred alone:
 * dlin_bias[h][d] += cs[h][d] (may be NULL) and dup_bias[h][e] += sum_d cs[h][d] * wlin[h][d][e]  (the column sums
 * of dz = dpred @ Wlin, i.e. the ConvTranspose1d bias gradient, without a pass over dz).  cs: fp32 [n][D] at
 * cs_stride; wlin: bf16 [n][D][E] at wlin_stride; both gradient blocks at grad_stride floats per head. */
int fhb_head_bias_grads(const float* colsum_dpred, int64_t cs_stride, const void* wlin, int64_t wlin_stride,
                        float* dlin_bias, float* dup_bias, int64_t grad_stride, int32_t n_heads, int32_t D, int32_t E,
                        fhb_stream_t stream);
/* out = alpha * a * m elementwise (bf16), same batching; m is a multiplier saved by FHB_EPI_AUX_DGELU, alpha the
 * feature_grad_mult of fairseq's GradMultiply (modules/model.py:428-431; 1 for the FitHuBERT recipe) */
int fhb_mul_bf16(const void* a, int64_t a_bstride, const void* m, int64_t m_bstride, void* out,
                 int64_t out_bstride, int32_t B, int64_t n, float alpha, fhb_stream_t stream);
/* nn.Dropout(p) forward AND backward (the op is its own adjoint): y[i] = x[i] * mask(i) / (1 - p) for a flat
 * tensor of n bf16 elements (y may alias x).  mask(i): element pair (2j, 2j+1) shares the 32-bit murmur3-finalised
 * hash of (j * 0x9E3779B1 + seed); element keeps iff its 16-bit half >= round(p * 65536).  Every fused dropout
 * in this library (GEMM epilogue, LayerNorm backward, attention) generates the same mask for the same index.
 * Replaces nn.Dropout / F.dropout at modules/model.py:489, modules/module.py:294,566,573,578. */
int fhb_dropout(const void* x, void* y, int64_t n, uint32_t seed, float p, int32_t is_f16 /* 1: fp16 (what the
                library uses), 0: bf16 */, fhb_stream_t stream);
/* zero `height` runs of `width_bytes` bytes, `pitch_bytes` apart (halo rows / borders of strided buffers);
 * a cudaMemset2DAsync, no kernel */
int fhb_memset2d(void* ptr, int64_t pitch_bytes, int64_t width_bytes, int64_t height, fhb_stream_t stream);
/* lengths[b] = #(mask[b][:] == 0); mask is the reference's bool padding mask (True = pad),
 * utils/dataset.py:63-74 / fithubert/expert.py:60-63; first step of modules/model.py:453 */
int fhb_mask_lengths(const uint8_t* mask, int32_t B, int64_t L, int32_t* lengths, fhb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FHB_H_ */
