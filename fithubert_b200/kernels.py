"""Thin tensor -> C-ABI wrappers (no autograd, no fallbacks).  Every function enqueues
hand-written sm_100a kernels from libfhb_sm100a.so on the current torch CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L

bf16 = torch.bfloat16


def _cuda(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise L.FhbError("fithubert_b200 kernels need CUDA tensors (no CPU fallback)")
    return t


def gemm_raw(a: L.Tensor3, b: L.Tensor3, d: torch.Tensor, m: int, n: int, k: int, *, a_major=0, b_major=0,
             num_ob=1, ob_mod=1, num_cb=1, a_coord=(0, 0, 0, 0), b_coord=(0, 0, 0, 0),
             d_ld: int, d_hi_stride=0, d_lo_stride=0, flags=0, split_k=0, bias=None, residual=None,
             aux_in=None, aux_out=None, row_valid=None, loss_target=None, loss_acc=None,
             loss_weight=0.0, grad_scale=0.0, d_offset_elems=0) -> None:
    g = L.GemmArgs()
    g.a, g.b = a, b
    g.a_major, g.b_major = a_major, b_major
    g.m, g.n, g.k = m, n, k
    g.num_ob, g.ob_mod, g.num_cb = num_ob, ob_mod, num_cb
    g.a_lo_c0, g.a_hi_c2, g.a_lo_c2, g.a_cb_c2 = a_coord
    g.b_lo_c0, g.b_hi_c2, g.b_lo_c2, g.b_cb_c2 = b_coord
    g.d = d.data_ptr() + d_offset_elems * d.element_size()
    g.d_ld, g.d_hi_stride, g.d_lo_stride = d_ld, d_hi_stride, d_lo_stride
    if d.dtype == torch.float32:
        flags |= L.EPI_OUT_F32
    else:
        assert d.dtype == bf16
    g.flags, g.split_k = flags, split_k
    for name, t in (("bias", bias), ("residual", residual), ("aux_in", aux_in), ("aux_out", aux_out),
                    ("row_valid", row_valid), ("loss_target", loss_target), ("loss_acc", loss_acc)):
        setattr(g, name, None if t is None else t.data_ptr())
    g.loss_weight, g.grad_scale = loss_weight, grad_scale
    L.check(L.lib().fhb_gemm(C.byref(g), L.stream_ptr()), "fhb_gemm")


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, gelu=False,
           residual=None, row_valid=None, rows_per_batch=0, preact_out=None, out_dtype=bf16,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[M,N] = epi(x[M,K] @ w[N,K]^T).  x, w bf16 row-major (x may be a strided 2-D view)."""
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
    if preact_out is not None:
        flags |= L.EPI_STORE_PREACT
    if row_valid is not None:
        # rows are [batch, rows_per_batch] flattened: run as a batched problem so ob_hi = sample
        flags |= L.EPI_ROWZERO
        nb = M // rows_per_batch
        a3 = L.tensor3(x.view(nb, rows_per_batch, K))
        gemm_raw(a3, L.tensor3(w), y, rows_per_batch, N, K, num_ob=nb, a_coord=(0, 1, 0, 0),
                 d_ld=y.stride(0), d_hi_stride=rows_per_batch * y.stride(0), flags=flags, bias=bias,
                 residual=residual, aux_out=preact_out, row_valid=row_valid)
        return y
    gemm_raw(L.tensor3(x), L.tensor3(w), y, M, N, K, d_ld=y.stride(0), flags=flags, bias=bias,
             residual=residual, aux_out=preact_out)
    return y


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, *, dgelu_of=None, residual=None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx[M,K] = dy[M,N] @ w[N,K]  (w consumed MN-major: no transposed copy)."""
    M, N = dy.shape
    K = w.shape[1]
    dx = out if out is not None else torch.empty(M, K, device=dy.device, dtype=bf16)
    flags = (L.EPI_MUL_DGELU if dgelu_of is not None else 0) | (L.EPI_RESIDUAL if residual is not None else 0)
    b3 = L.tensor3(data_ptr=w.data_ptr(), dim=(K, N, 1), stride=(w.stride(0), w.stride(0) * N))
    gemm_raw(L.tensor3(dy), b3, dx, M, K, N, b_major=1, d_ld=dx.stride(0), flags=flags, aux_in=dgelu_of,
             residual=residual)
    return dx


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                 accumulate=False) -> torch.Tensor:
    """dw[N,K] (fp32) = dy[M,N]^T @ x[M,K]; both operands MN-major, split-K with fp32 atomics."""
    M, N = dy.shape
    K = x.shape[1]
    if out is None:
        out = torch.zeros(N, K, device=dy.device, dtype=torch.float32)
    elif not accumulate:
        out.zero_()
    a3 = L.tensor3(data_ptr=dy.data_ptr(), dim=(N, M, 1), stride=(dy.stride(0), dy.stride(0) * M))
    b3 = L.tensor3(data_ptr=x.data_ptr(), dim=(K, M, 1), stride=(x.stride(0), x.stride(0) * M))
    gemm_raw(a3, b3, out, N, K, M, a_major=1, b_major=1, d_ld=out.stride(0), flags=L.EPI_ATOMIC_ADD)
    return out
