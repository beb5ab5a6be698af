"""Thin tensor -> C-ABI wrappers (no autograd, no fallbacks).  Every function enqueues
hand-written sm_100a kernels from libfhb_sm100a.so on the current torch CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L

f16 = torch.float16     # every 16-bit tensor: weights, activations, saved gelu', projections, targets, and the
bf16 = f16              # gradients (which carry the loss scale, engine.loss_scale_for); the old name stays as an alias


def _cuda(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise L.FhbError("fithubert_b200 kernels need CUDA tensors (no CPU fallback)")
    return t


_GEMM_TIMING = {"on": False, "events": []}
reset_counters = L.reset_counters
launch_count = L.launch_count


def enable_gemm_timing(on: bool) -> None:
    """Bracket every fhb_gemm launch with a CUDA-event pair.  Programmatic dependent launch is switched off
    while timing so that each pair covers the whole kernel (prologue included), not just its un-overlapped part."""
    if on:
        _GEMM_TIMING["pdl"] = L.lib().fhb_set_pdl(0)
    elif "pdl" in _GEMM_TIMING:
        L.lib().fhb_set_pdl(_GEMM_TIMING.pop("pdl"))
    _GEMM_TIMING["on"] = on
    _GEMM_TIMING["events"] = []


def gemm_timing_summary():
    """(total ms, total algorithmic flops, launches) of the fhb_gemm launches recorded since enable."""
    torch.cuda.synchronize()
    ev = _GEMM_TIMING["events"]
    return sum(e[0].elapsed_time(e[1]) for e in ev), float(sum(e[2] for e in ev)), len(ev)


def gemm_timing_groups(classify):
    """{group: (ms, flops, launches)} of the recorded fhb_gemm launches; classify((rows, n, k)) -> group name."""
    torch.cuda.synchronize()
    out = {}
    for e0, e1, fl, shape, _ in _GEMM_TIMING["events"]:
        g = classify(shape)
        ms, f, c = out.get(g, (0.0, 0.0, 0))
        out[g] = (ms + e0.elapsed_time(e1), f + fl, c + 1)
    return out


def gemm_raw(a: L.Tensor3, b: L.Tensor3, d: torch.Tensor, m: int, n: int, k: int, *, a_major=0, b_major=0,
             num_ob=1, ob_mod=1, num_cb=1, a_coord=(0, 0, 0, 0), b_coord=(0, 0, 0, 0),
             d_ld: int, d_hi_stride=0, d_lo_stride=0, flags=0, split_k=0, bias=None, residual=None,
             aux_in=None, aux_out=None, row_valid=None, loss_target=None, loss_acc=None,
             loss_weight=0.0, grad_scale=0.0, d_offset_elems=0, a_c1_off=0, b_c1_off=0, bias_hi_stride=0,
             drop=None, alpha=None) -> None:
    g = L.GemmArgs()
    g.a, g.b = a, b
    g.a_major, g.b_major = a_major, b_major
    g.m, g.n, g.k = m, n, k
    g.num_ob, g.ob_mod, g.num_cb = num_ob, ob_mod, num_cb
    g.a_lo_c0, g.a_hi_c2, g.a_lo_c2, g.a_cb_c2 = a_coord
    g.b_lo_c0, g.b_hi_c2, g.b_lo_c2, g.b_cb_c2 = b_coord
    g.a_c1_off, g.b_c1_off = a_c1_off, b_c1_off
    g.d = d.data_ptr() + d_offset_elems * d.element_size()
    g.d_ld, g.d_hi_stride, g.d_lo_stride = d_ld, d_hi_stride, d_lo_stride
    # operand / output formats travel in the flags: fp16 (forward tensors) unless flagged bf16 (gradients) or fp32
    assert not a.bf16 and not b.bf16, "operands are fp16 (tcgen05 cannot mix fp16 with bf16; gradients carry a loss scale)"
    if d.dtype == torch.float32:
        flags |= L.EPI_OUT_F32
    else:
        assert d.dtype == f16, d.dtype
    assert aux_in is None or aux_in.dtype == f16, "aux_in is a forward quantity (fp16)"
    assert loss_target is None or loss_target.dtype == f16
    if aux_out is not None:
        assert aux_out.dtype == (f16 if (flags & L.EPI_AUX_DGELU) else d.dtype), (aux_out.dtype, d.dtype)
    if alpha is not None and alpha != 1.0:
        flags |= L.EPI_ALPHA
        g.alpha = alpha
    g.flags, g.split_k = flags, split_k
    for name, t in (("bias", bias), ("row_valid", row_valid), ("loss_acc", loss_acc)):
        setattr(g, name, None if t is None else t.data_ptr())
    for name, t in (("residual", residual), ("aux_in", aux_in), ("aux_out", aux_out), ("loss_target", loss_target)):
        # "laid out like D": same element offset as the output
        setattr(g, name, None if t is None else t.data_ptr() + d_offset_elems * t.element_size())
    if residual is not None:  # fp32: the high-precision copy of the residual stream (layernorm_fwd32 / layernorm_bwd32)
        flags |= {torch.float32: L.EPI_RES_F32, f16: 0}[residual.dtype]
    g.flags = flags
    g.loss_weight, g.grad_scale = loss_weight, grad_scale
    g.bias_hi_stride = bias_hi_stride
    if drop is not None and drop[1] > 0.0:  # (seed, p)
        flags |= L.EPI_DROPOUT
        g.flags = flags
        g.drop_seed, g.drop_p = drop[0] & 0xFFFFFFFF, drop[1]
    if _GEMM_TIMING["on"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib().fhb_gemm(C.byref(g), L.stream_ptr()), "fhb_gemm")
        e1.record()
        _GEMM_TIMING["events"].append((e0, e1, 2.0 * m * n * k * max(1, num_ob) * max(1, num_cb),
                                       (m * max(1, num_ob), n, k * max(1, num_cb)),
                                       (a_major, b_major, flags, split_k, num_ob)))
        return
    L.check(L.lib().fhb_gemm(C.byref(g), L.stream_ptr()), "fhb_gemm")


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, gelu=False,
           residual=None, row_valid=None, rows_per_batch=0, preact_out=None, dgelu_out=None, out_dtype=f16,
           out: Optional[torch.Tensor] = None, drop=None) -> torch.Tensor:
    """y[M,N] = epi(x[M,K] @ w[N,K]^T).  x, w fp16 row-major (x may be a strided 2-D view); y fp16 (or fp32).
    preact_out: also store the value before GELU / residual; dgelu_out: store gelu'(that value) instead
    (what the backward epilogue multiplies by: FHB_EPI_MUL_AUX)."""
    _cuda(x)
    M, K = x.shape
    N = w.shape[0]
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=out_dtype)
    flags = 0
    if bias is not None:
        flags |= L.EPI_BIAS
    if gelu:
        flags |= L.EPI_GELU
    if residual is not None:
        flags |= L.EPI_RESIDUAL
    if dgelu_out is not None:
        assert preact_out is None
        flags |= L.EPI_STORE_PREACT | L.EPI_AUX_DGELU
        preact_out = dgelu_out
    elif preact_out is not None:
        flags |= L.EPI_STORE_PREACT
    if row_valid is not None:
        # rows are [batch, rows_per_batch] flattened: run as a batched problem so ob_hi = sample
        flags |= L.EPI_ROWZERO
        nb = M // rows_per_batch
        a3 = L.tensor3(x.view(nb, rows_per_batch, K))
        gemm_raw(a3, L.tensor3(w), y, rows_per_batch, N, K, num_ob=nb, a_coord=(0, 1, 0, 0),
                 d_ld=y.stride(0), d_hi_stride=rows_per_batch * y.stride(0), flags=flags, bias=bias,
                 residual=residual, aux_out=preact_out, row_valid=row_valid, drop=drop)
        return y
    gemm_raw(L.tensor3(x), L.tensor3(w), y, M, N, K, d_ld=y.stride(0), flags=flags, bias=bias,
             residual=residual, aux_out=preact_out, drop=drop)
    return y


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, *, dgelu_of=None, mul_aux=None, residual=None,
                 out: Optional[torch.Tensor] = None, out_dtype=bf16) -> torch.Tensor:
    """dx[M,K] = dy[M,N] @ w[N,K]  (w consumed MN-major: no transposed copy), optionally * gelu'(dgelu_of)
    or * mul_aux (a saved gelu'), + residual (fp16 or fp32); dx fp16 (or fp32)."""
    M, N = dy.shape
    K = w.shape[1]
    dx = out if out is not None else torch.empty(M, K, device=dy.device, dtype=out_dtype)
    assert dgelu_of is None or mul_aux is None
    flags = (L.EPI_MUL_DGELU if dgelu_of is not None else 0) | (L.EPI_MUL_AUX if mul_aux is not None else 0) | \
        (L.EPI_RESIDUAL if residual is not None else 0)
    b3 = L.tensor3(w, data_ptr=w.data_ptr(), dim=(K, N, 1), stride=(w.stride(0), w.stride(0) * N))
    gemm_raw(L.tensor3(dy), b3, dx, M, K, N, b_major=1, d_ld=dx.stride(0), flags=flags,
             aux_in=dgelu_of if dgelu_of is not None else mul_aux, residual=residual)
    return dx


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                 accumulate=False) -> torch.Tensor:
    """dw[N,K] (fp32) = dy[M,N]^T @ x[M,K] ; both operands MN-major, split-K with fp32 atomics."""
    M, N = dy.shape
    K = x.shape[1]
    if out is None:
        out = torch.zeros(N, K, device=dy.device, dtype=torch.float32)
    elif not accumulate:
        out.zero_()
    a3 = L.tensor3(dy, data_ptr=dy.data_ptr(), dim=(N, M, 1), stride=(dy.stride(0), dy.stride(0) * M))
    b3 = L.tensor3(x, data_ptr=x.data_ptr(), dim=(K, M, 1), stride=(x.stride(0), x.stride(0) * M))
    gemm_raw(a3, b3, out, N, K, M, a_major=1, b_major=1, d_ld=out.stride(0), flags=L.EPI_ATOMIC_ADD)
    return out


# ----------------------------------------------------------------------------- non-GEMM kernels
def _f(x):
    return C.c_float(float(x))


def conv0_fwd(wave, weight, gamma, beta, T0, stat, mean, rstd, out, eps=1e-5, gp_out=None):
    assert out.dtype == f16 and (gp_out is None or gp_out.dtype == f16)
    a = L.Conv0Args()
    B, Ld = wave.shape
    a.wave, a.wave_ld = wave.data_ptr(), wave.stride(0)
    a.B, a.L, a.C, a.T0, a.kernel, a.stride, a.eps = B, Ld, weight.shape[0], T0, weight.shape[-1], 5, eps
    a.weight, a.gamma, a.beta = weight.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.stat, a.mean, a.rstd, a.out = stat.data_ptr(), mean.data_ptr(), rstd.data_ptr(), out.data_ptr()
    a.gp_out = None if gp_out is None else gp_out.data_ptr()
    L.check(L.lib().fhb_conv0_gn_gelu_fwd(C.byref(a), L.stream_ptr()), "fhb_conv0_gn_gelu_fwd")


def conv0_bwd(wave, weight, gamma, beta, T0, stat, mean, rstd, dy, acc, dweight, dgamma, dbeta,
              accumulate=True, eps=1e-5, dy_is_dz=False):
    a = L.Conv0Args()
    B, Ld = wave.shape
    a.wave, a.wave_ld = wave.data_ptr(), wave.stride(0)
    a.B, a.L, a.C, a.T0, a.kernel, a.stride, a.eps = B, Ld, weight.shape[0], T0, weight.shape[-1], 5, eps
    a.weight, a.gamma, a.beta = weight.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.stat, a.mean, a.rstd = stat.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.dy, a.acc, a.dy_is_dz = dy.data_ptr(), acc.data_ptr(), int(dy_is_dz)
    a.dweight, a.dgamma, a.dbeta, a.accumulate = dweight.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), int(accumulate)
    L.check(L.lib().fhb_conv0_gn_gelu_bwd(C.byref(a), L.stream_ptr()), "fhb_conv0_gn_gelu_bwd")


def conv0_im2col(wave, T0, xcol):
    """xcol bf16 [B, T0, 32]: two-term bf16 split of the 10 taps of every frame + a ones column (conv0 backward GEMM)."""
    B, Ld = wave.shape
    L.check(L.lib().fhb_conv0_im2col(L.ptr(wave), C.c_int64(wave.stride(0)), B, Ld, T0, L.ptr(xcol), L.stream_ptr()),
            "fhb_conv0_im2col")


def conv0_bwd_finalize(acc32, wave, weight, gamma, beta, T0, stat, mean, rstd, dweight, dgamma, dbeta, accumulate=True,
                       eps=1e-5):
    a = L.Conv0Args()
    B, Ld = wave.shape
    a.wave, a.wave_ld = wave.data_ptr(), wave.stride(0)
    a.B, a.L, a.C, a.T0, a.kernel, a.stride, a.eps = B, Ld, weight.shape[0], T0, weight.shape[-1], 5, eps
    a.weight, a.gamma, a.beta = weight.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.stat, a.mean, a.rstd = stat.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.dweight, a.dgamma, a.dbeta, a.accumulate = dweight.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), int(accumulate)
    L.check(L.lib().fhb_conv0_bwd_finalize(L.ptr(acc32), C.byref(a), L.stream_ptr()), "fhb_conv0_bwd_finalize")


def layernorm_fwd(x, gamma, beta, y, mean=None, rstd=None, eps=1e-5):
    assert x.dtype == f16 and y.dtype == f16
    rows, Cd = x.numel() // x.shape[-1], x.shape[-1]
    L.check(L.lib().fhb_layernorm_fwd(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), L.ptr(mean), L.ptr(rstd),
                                      C.c_int64(rows), Cd, _f(eps), L.stream_ptr()), "fhb_layernorm_fwd")
    return y


def layernorm_fwd32(x32, gamma, beta, y, y32=None, mean=None, rstd=None, sub32=None, diff_out=None, eps=1e-5):
    """LayerNorm of the fp32 pre-LN sum: y bf16 (GEMM operand), y32 the fp32 copy the next residual add reads;
    diff_out = bf16(x32 - sub32) (the FFN branch output the reference returns as `layer_result`)."""
    assert x32.dtype == torch.float32 and (y32 is None or y32.dtype == torch.float32)
    assert y.dtype == f16 and (diff_out is None or diff_out.dtype == f16)
    rows, Cd = x32.numel() // x32.shape[-1], x32.shape[-1]
    L.check(L.lib().fhb_layernorm_fwd32(L.ptr(x32), L.ptr(gamma), L.ptr(beta), L.ptr(y), L.ptr(y32), L.ptr(mean),
                                        L.ptr(rstd), L.ptr(sub32), L.ptr(diff_out), C.c_int64(rows), Cd, _f(eps),
                                        L.stream_ptr()), "fhb_layernorm_fwd32")
    return y


def layernorm_bwd32(dy32, x32, gamma, mean, rstd, dgamma, dbeta, *, dy2=None, dx=None, dx32=None, dxsum=None,
                    dx_drop=None, drop=None):
    """Backward of layernorm_fwd32.  Gradient in = dy32 (fp32, optional) + dy2 (bf16, optional); out = dx32 (fp32,
    the backward residual stream), dx (bf16) and / or dx_drop (bf16, dropout-masked, with drop=(seed, p))."""
    assert x32.dtype == torch.float32 and (dy32 is None or dy32.dtype == torch.float32)
    assert dy2 is None or dy2.dtype == bf16
    assert (dx is None or dx.dtype == bf16) and (dx_drop is None or dx_drop.dtype == bf16)
    rows, Cd = x32.numel() // x32.shape[-1], x32.shape[-1]
    seed, p = (drop[0] & 0xFFFFFFFF, drop[1]) if (drop is not None and dx_drop is not None) else (0, 0.0)
    L.check(L.lib().fhb_layernorm_bwd32(L.ptr(dy32), L.ptr(dy2), L.ptr(x32), L.ptr(gamma), L.ptr(mean), L.ptr(rstd),
                                        L.ptr(dx), L.ptr(dx32), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dxsum), L.ptr(dx_drop),
                                        C.c_uint32(seed), _f(p), C.c_int64(rows), Cd, L.stream_ptr()),
            "fhb_layernorm_bwd32")


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, dres=None, dxsum=None, dy2=None, dx_drop=None,
                  drop=None):
    """dxsum (fp32 [C], accumulated): column sums of dx = bias gradient of the linear layer that produced x.
    dy2: optional second gradient stream, summed with dy on load.  dx_drop + drop=(seed, p): also write
    dx * dropout-mask (the gradient entering a `residual + dropout(branch)` branch; dxsum then sums that)."""
    assert dy.dtype == bf16 and x.dtype == f16 and dx.dtype == bf16, "16-bit tensors are fp16"
    rows, Cd = x.numel() // x.shape[-1], x.shape[-1]
    seed, p = (drop[0] & 0xFFFFFFFF, drop[1]) if (drop is not None and dx_drop is not None) else (0, 0.0)
    L.check(L.lib().fhb_layernorm_bwd(L.ptr(dy), L.ptr(dy2), L.ptr(x), L.ptr(gamma), L.ptr(mean), L.ptr(rstd), L.ptr(dres),
                                      L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dxsum), L.ptr(dx_drop),
                                      C.c_uint32(seed), _f(p), C.c_int64(rows), Cd, L.stream_ptr()), "fhb_layernorm_bwd")
    return dx


def posconv_pack(x, valid, xg, B, T, Cd, G, cp, pad_l, Tp):
    L.check(L.lib().fhb_posconv_pack(L.ptr(x), L.ptr(valid), L.ptr(xg), B, T, Cd, G, cp, pad_l, Tp, L.stream_ptr()),
            "fhb_posconv_pack")


def posconv_wn_prep(v, g, w_fwd, w_bwd, ws, Cd, G, Kt, cp, delta=1):
    """ws: fp32 [2 * Kt] workspace; ws[Kt:] receives 1 / ||v[:, :, j]|| (posconv_wn_bwd's inv_norm)."""
    L.check(L.lib().fhb_posconv_wn_prep(L.ptr(v), L.ptr(g), L.ptr(w_fwd), L.ptr(w_bwd), L.ptr(ws), Cd, G, Kt, cp, delta,
                                        L.stream_ptr()), "fhb_posconv_wn_prep")


def posconv_finish_fwd(x, valid, conv, bias, gamma, beta, h_out, y, mean, rstd, B, T, Cd, G, cp, eps=1e-5, delta=1,
                       y32=None):
    assert x.dtype == f16 and conv.dtype == f16 and y.dtype == f16 and (h_out is None or h_out.dtype == f16)
    L.check(L.lib().fhb_posconv_finish_fwd(L.ptr(x), L.ptr(valid), L.ptr(conv), L.ptr(bias), L.ptr(gamma), L.ptr(beta),
                                           L.ptr(h_out), L.ptr(y), L.ptr(y32), L.ptr(mean), L.ptr(rstd), B, T, Cd, G, cp,
                                           _f(eps), delta, L.stream_ptr()), "fhb_posconv_finish_fwd")


def posconv_finish_bwd(dy, h, conv, bias, gamma, mean, rstd, dh, dcg, dgamma, dbeta, dbias, B, T, Cd, G, cp, pad_l, Tp,
                       delta=1):
    assert dy.dtype == bf16 and h.dtype == f16 and conv.dtype == f16 and dh.dtype == bf16 and dcg.dtype == bf16
    L.check(L.lib().fhb_posconv_finish_bwd(L.ptr(dy), L.ptr(h), L.ptr(conv), L.ptr(bias), L.ptr(gamma), L.ptr(mean),
                                           L.ptr(rstd), L.ptr(dh), L.ptr(dcg), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dbias),
                                           B, T, Cd, G, cp, pad_l, Tp, delta, L.stream_ptr()), "fhb_posconv_finish_bwd")


def posconv_unpack_bwd(dh, dxc, valid, dx, B, T, Cd, G, cp, delta=1):
    L.check(L.lib().fhb_posconv_unpack_bwd(L.ptr(dh), L.ptr(dxc), L.ptr(valid), L.ptr(dx), B, T, Cd, G, cp, delta,
                                           L.stream_ptr()), "fhb_posconv_unpack_bwd")


def posconv_wn_bwd(dwt, v, g, ws, dv, dg, Cd, G, Kt, cp, accumulate=True, delta=1):
    """ws: posconv_wn_prep's fp32 [2 * Kt] workspace (ws[Kt:] = 1 / norm is read, ws[:Kt] is scratch)."""
    assert ws.numel() >= 2 * Kt and ws.dtype == torch.float32
    L.check(L.lib().fhb_posconv_wn_bwd(L.ptr(dwt), L.ptr(v), L.ptr(g), L.ptr(ws), L.ptr(dv), L.ptr(dg), Cd, G, Kt,
                                       cp, int(accumulate), delta, L.stream_ptr()), "fhb_posconv_wn_bwd")


def _drop(drop):
    return (C.c_uint32(drop[0] & 0xFFFFFFFF), _f(drop[1])) if drop is not None else (C.c_uint32(0), _f(0.0))


def attn_fwd(qkv, valid, out, lse, B, T, H, d, scale, drop=None):
    """drop = (seed, p): attention dropout on the probabilities (same pair passed to attn_bwd)."""
    assert qkv.dtype == f16 and out.dtype == f16
    L.check(L.lib().fhb_attn_fwd(L.ptr(qkv), L.ptr(valid), L.ptr(out), L.ptr(lse), B, T, H, d, _f(scale), *_drop(drop),
                                 L.stream_ptr()), "fhb_attn_fwd")


def attn_bwd(qkv, valid, out, dout, lse, dqkv, delta_ws, B, T, H, d, scale, drop=None, dq_ws=None):
    """dq_ws (fp32 [B, T, H*d]): enables the fused tcgen05 backward for d in {40, 64}."""
    assert qkv.dtype == f16 and out.dtype == f16 and dout.dtype == bf16 and dqkv.dtype == bf16
    L.check(L.lib().fhb_attn_bwd(L.ptr(qkv), L.ptr(valid), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(dqkv),
                                 L.ptr(delta_ws), L.ptr(dq_ws), B, T, H, d, _f(scale), *_drop(drop), L.stream_ptr()),
            "fhb_attn_bwd")


def attn_scores(a, b, valid, B, T, H, d, scale, out=None):
    """Un-normalised attention logits / value relation of one layer (reference utils/utils.py:190-232): a, b = two head
    blocks of the fused [B*T, 3*H*d] projection output (strided views q | k | v).  Returns fp32 [B*H, T, pitch]."""
    assert a.dtype == f16 and b.dtype == f16 and a.stride(0) == b.stride(0) and a.stride(1) == 1
    pitch = (T + 7) // 8 * 8
    if out is None:
        out = torch.empty(B * H, T, pitch, device=a.device, dtype=torch.float32)
    L.check(L.lib().fhb_attn_scores(L.ptr(a), L.ptr(b), C.c_int64(a.stride(0)), L.ptr(valid), L.ptr(out), C.c_int64(pitch),
                                    B, T, H, d, _f(scale), L.stream_ptr()), "fhb_attn_scores")
    return out


def attn_map_loss(s, t, valid_s, valid_t, ds, loss, B, T, H, mode, loss_mult, grad_mult):
    """mode 0: mse over the keys neither side masks; 1: KL(softmax(t) || softmax(s)) per row (train.py:327-368).
    loss (fp32 [1]) is accumulated; ds (fp16, laid out like s) receives grad_mult * d(row terms)/d(s)."""
    assert s.dtype == torch.float32 and t.dtype == torch.float32 and ds.dtype == f16 and s.shape == t.shape == ds.shape
    L.check(L.lib().fhb_attn_map_loss(L.ptr(s), L.ptr(t), C.c_int64(s.shape[-1]), L.ptr(valid_s), L.ptr(valid_t), L.ptr(ds),
                                      L.ptr(loss), B, T, H, mode, _f(loss_mult), _f(grad_mult), L.stream_ptr()),
            "fhb_attn_map_loss")


def attn_scores_bwd(g, m, out, B, T, H, d, alpha, trans, accumulate=True):
    """out head block (+)= alpha * (g or g^T) @ m head block: the gradient of attn_scores back into q / k / v."""
    assert g.dtype == f16 and m.dtype == f16 and out.dtype == f16 and m.stride(1) == 1 and out.stride(1) == 1
    L.check(L.lib().fhb_attn_scores_bwd(L.ptr(g), C.c_int64(g.shape[-1]), L.ptr(m), C.c_int64(m.stride(0)), L.ptr(out),
                                        C.c_int64(out.stride(0)), B, T, H, d, _f(alpha), int(trans), int(accumulate),
                                        L.stream_ptr()), "fhb_attn_scores_bwd")


def distill_loss(pred, tgt, weights, layer_loss, dpred, n_layers, B, Tp, Tt, D, loss_type=0, grad_scale=1.0,
                 dbias=None, dbias_layer_stride=0):
    assert pred.dtype == f16 and tgt.dtype == f16 and (dpred is None or dpred.dtype == bf16)
    L.check(L.lib().fhb_distill_loss_fwd_bwd(L.ptr(pred), L.ptr(tgt), L.ptr(weights), L.ptr(layer_loss), L.ptr(dpred),
                                             L.ptr(dbias), C.c_int64(dbias_layer_stride), n_layers, B, Tp, Tt, D,
                                             loss_type, _f(grad_scale), L.stream_ptr()), "fhb_distill_loss_fwd_bwd")


def distill_loss_sim(pred, tgt, weights, rec_layer_loss, sim_layer_loss, dpred, n_layers, B, Tp, Tt, D, loss_type=0,
                     rec_grad_scale=1.0, sim_grad_scale=1.0, dbias=None, dbias_layer_stride=0):
    """Reconstruction (mse / l1) + cosine (-logsigmoid(cos)) hint loss and its gradient in one pass (train.py:282-314)."""
    assert pred.dtype == f16 and tgt.dtype == f16 and (dpred is None or dpred.dtype == bf16)
    L.check(L.lib().fhb_distill_loss_sim_fwd_bwd(L.ptr(pred), L.ptr(tgt), L.ptr(weights), L.ptr(rec_layer_loss),
                                                 L.ptr(sim_layer_loss), L.ptr(dpred), L.ptr(dbias),
                                                 C.c_int64(dbias_layer_stride), n_layers, B, Tp, Tt, D, loss_type,
                                                 _f(rec_grad_scale), _f(sim_grad_scale), L.stream_ptr()),
            "fhb_distill_loss_sim_fwd_bwd")


def adamw_multi(table, n_tensors, max_n, lr, beta1, beta2, eps, wd, step, mode=0, grad_scale=1.0):
    L.check(L.lib().fhb_adamw_multi(L.ptr(table), n_tensors, C.c_int64(max_n), _f(lr), _f(beta1), _f(beta2), _f(eps),
                                    _f(wd), step, mode, _f(grad_scale), L.stream_ptr()), "fhb_adamw_multi")


def prep_multi(table, n_tensors, max_n):
    L.check(L.lib().fhb_prep_multi(L.ptr(table), n_tensors, C.c_int64(max_n), L.stream_ptr()), "fhb_prep_multi")


def colsum(x2d, out):
    assert x2d.dtype == bf16, "16-bit tensors are fp16"
    rows, Cd = x2d.shape
    L.check(L.lib().fhb_colsum(L.ptr(x2d), C.c_int64(rows), Cd, C.c_int64(x2d.stride(0)), L.ptr(out), L.stream_ptr()),
            "fhb_colsum")


def colsum_batched(x3d, out, out_bstride):
    """x3d [n, rows, C] (contiguous), out: fp32 view whose batch b lives at out + b * out_bstride elements."""
    n, rows, Cd = x3d.shape
    assert x3d.dtype == bf16
    L.check(L.lib().fhb_colsum_batched(L.ptr(x3d), C.c_int64(rows), Cd, C.c_int64(x3d.stride(1)),
                                       C.c_int64(x3d.stride(0)), L.ptr(out), C.c_int64(out_bstride), n, L.stream_ptr()),
            "fhb_colsum_batched")


def head_bias_grads(cs, wlin, wlin_stride, dlin_bias, dup_bias, grad_stride, n_heads, D, Ed):
    """cs: fp32 [n, D] column sums of dpred.  dlin_bias (or None) += cs; dup_bias += cs @ Wlin per head."""
    L.check(L.lib().fhb_head_bias_grads(L.ptr(cs), C.c_int64(cs.stride(0)), L.ptr(wlin), C.c_int64(wlin_stride),
                                        L.ptr(dlin_bias), L.ptr(dup_bias), C.c_int64(grad_stride), n_heads, D, Ed,
                                        L.stream_ptr()), "fhb_head_bias_grads")


def add_bf16(a, b, y):
    assert a.dtype == bf16 and b.dtype == bf16 and y.dtype == bf16
    L.check(L.lib().fhb_add_bf16(L.ptr(a), L.ptr(b), L.ptr(y), C.c_int64(a.numel()), L.stream_ptr()), "fhb_add_bf16")
    return y


def mul_dgelu(dy, dy_bs, u, u_bs, out, out_bs, B, n, *, u_off=0, out_off=0):
    L.check(L.lib().fhb_mul_dgelu(L.ptr(dy), C.c_int64(dy_bs), C.c_void_p(u.data_ptr() + 2 * u_off), C.c_int64(u_bs),
                                  C.c_void_p(out.data_ptr() + 2 * out_off), C.c_int64(out_bs), B, C.c_int64(n),
                                  L.stream_ptr()), "fhb_mul_dgelu")


def mul_bf16(a, a_bs, m, m_bs, out, out_bs, B, n, alpha=1.0):
    """out = alpha * a * m (alpha: fairseq GradMultiply's scale, modules/model.py:428-431)."""
    assert a.dtype == bf16 and m.dtype == f16 and out.dtype == bf16, "16-bit tensors are fp16"
    L.check(L.lib().fhb_mul_bf16(L.ptr(a), C.c_int64(a_bs), L.ptr(m), C.c_int64(m_bs), L.ptr(out), C.c_int64(out_bs), B,
                                 C.c_int64(n), _f(alpha), L.stream_ptr()), "fhb_mul_bf16")


def dropout(x, y, seed, p):
    """y = nn.Dropout(p)(x) with the library's counter-based mask (also its own backward); y may alias x."""
    assert x.dtype == f16 and y.dtype == f16
    L.check(L.lib().fhb_dropout(L.ptr(x), L.ptr(y), C.c_int64(x.numel()), C.c_uint32(seed & 0xFFFFFFFF), _f(p),
                                int(x.dtype == f16), L.stream_ptr()), "fhb_dropout")
    return y


def mask_lengths(mask_u8, lengths):
    B, Ld = mask_u8.shape
    L.check(L.lib().fhb_mask_lengths(L.ptr(mask_u8), B, C.c_int64(Ld), L.ptr(lengths), L.stream_ptr()),
            "fhb_mask_lengths")


def zero_rows(buf: torch.Tensor, elem_offset: int, pitch_elems: int, width_elems: int, height: int):
    """Zero `height` runs of `width_elems` elements, `pitch_elems` apart, starting at `elem_offset`."""
    es = buf.element_size()
    L.check(L.lib().fhb_memset2d(C.c_void_p(buf.data_ptr() + elem_offset * es), C.c_int64(pitch_elems * es),
                                 C.c_int64(width_elems * es), C.c_int64(height), L.stream_ptr()), "fhb_memset2d")


def zero_(buf: torch.Tensor):
    n = buf.numel() * buf.element_size()
    L.check(L.lib().fhb_memset2d(L.ptr(buf), C.c_int64(n), C.c_int64(n), C.c_int64(1), L.stream_ptr()), "fhb_memset2d")
    return buf
