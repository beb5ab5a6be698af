"""ctypes binding of libfhb_sm100a.so (include/fhb.h).  There is NO fallback: if the
library is missing or a call fails this module raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FHB_LIB: another build of the same library (A/B runs of kernel variants on one box); default: the in-tree build
LIB_PATH = os.environ.get("FHB_LIB") or os.path.join(_HERE, "libfhb_sm100a.so")

EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_ROWZERO = 1, 2, 4, 8
EPI_STORE_PREACT, EPI_MUL_DGELU, EPI_OUT_F32, EPI_ATOMIC_ADD, EPI_SQDIFF = 16, 32, 64, 128, 256
EPI_AUX_DGELU, EPI_MUL_AUX, EPI_DROPOUT, EPI_RES_F32 = 512, 1024, 2048, 4096
GEMM_A_BF16, GEMM_B_BF16, EPI_OUT_BF16, EPI_RES_BF16, EPI_ALPHA = 8192, 16384, 32768, 65536, 131072


class FhbError(RuntimeError):
    pass


class Tensor3(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", C.c_int64 * 3), ("stride", C.c_int64 * 2)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", Tensor3), ("b", Tensor3),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("num_ob", C.c_int32), ("ob_mod", C.c_int32), ("num_cb", C.c_int32),
        ("a_lo_c0", C.c_int32), ("a_hi_c2", C.c_int32), ("a_lo_c2", C.c_int32), ("a_cb_c2", C.c_int32),
        ("b_lo_c0", C.c_int32), ("b_hi_c2", C.c_int32), ("b_lo_c2", C.c_int32), ("b_cb_c2", C.c_int32),
        ("a_c1_off", C.c_int32), ("b_c1_off", C.c_int32),
        ("d", C.c_void_p),
        ("d_ld", C.c_int64), ("d_hi_stride", C.c_int64), ("d_lo_stride", C.c_int64),
        ("flags", C.c_int32), ("split_k", C.c_int32),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("aux_in", C.c_void_p), ("aux_out", C.c_void_p),
        ("row_valid", C.c_void_p),
        ("loss_target", C.c_void_p), ("loss_acc", C.c_void_p),
        ("loss_weight", C.c_float), ("grad_scale", C.c_float),
        ("bias_hi_stride", C.c_int64),
        ("drop_seed", C.c_uint32), ("drop_p", C.c_float), ("alpha", C.c_float),
    ]


class Conv0Args(C.Structure):
    _fields_ = [
        ("wave", C.c_void_p), ("wave_ld", C.c_int64),
        ("B", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("T0", C.c_int32),
        ("kernel", C.c_int32), ("stride", C.c_int32), ("eps", C.c_float),
        ("weight", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("stat", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("out", C.c_void_p), ("dy", C.c_void_p), ("acc", C.c_void_p),
        ("dweight", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("accumulate", C.c_int32),
        ("gp_out", C.c_void_p), ("dy_is_dz", C.c_int32),
    ]


class AdamwTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("n", C.c_int64), ("dim", C.c_int64 * 3), ("gstride", C.c_int64 * 3)]


class PrepTensor(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dim", C.c_int64 * 3), ("sstride", C.c_int64 * 3),
                ("dst_is_f32", C.c_int32), ("accumulate", C.c_int32)]


def table_to_device(entries, device) -> torch.Tensor:
    """Pack a list of ctypes structs into one device-resident byte tensor (a kernel argument table)."""
    arr = (type(entries[0]) * len(entries))(*entries)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FhbError(
                f"{LIB_PATH} not found: build it with `python -m fithubert_b200.build` "
                "(there is no CPU / PyTorch fallback for the FitHuBERT hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.fhb_last_error.restype = C.c_char_p
        _declare(_lib)
    return _lib


def _declare(l):
    for name in EXPORTS:
        fn = getattr(l, name)
        if name not in ("fhb_last_error",):
            fn.restype = C.c_int


# every symbol include/fhb.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "fhb_last_error", "fhb_abi_version", "fhb_set_pdl", "fhb_set_reserved_sms", "fhb_gemm", "fhb_conv0_gn_gelu_fwd", "fhb_conv0_gn_gelu_bwd", "fhb_conv0_im2col", "fhb_conv0_bwd_finalize",
    "fhb_layernorm_fwd", "fhb_layernorm_bwd", "fhb_layernorm_fwd32", "fhb_layernorm_bwd32", "fhb_posconv_pack", "fhb_posconv_wn_prep",
    "fhb_posconv_finish_fwd", "fhb_posconv_finish_bwd", "fhb_posconv_unpack_bwd", "fhb_posconv_wn_bwd",
    "fhb_attn_fwd", "fhb_attn_bwd", "fhb_attn_scores", "fhb_attn_map_loss", "fhb_attn_scores_bwd", "fhb_distill_loss_fwd_bwd", "fhb_distill_loss_sim_fwd_bwd", "fhb_adamw_multi", "fhb_prep_multi",
    "fhb_colsum", "fhb_colsum_batched", "fhb_head_bias_grads", "fhb_add_bf16", "fhb_mul_dgelu", "fhb_mul_bf16", "fhb_dropout", "fhb_mask_lengths", "fhb_memset2d",
]


CALLS: dict = {}
_KERNELS_PER_CALL = {"fhb_conv0_gn_gelu_fwd": 3, "fhb_conv0_gn_gelu_bwd": 2, "fhb_attn_bwd": 3, "fhb_posconv_wn_prep": 2}


def reset_counters() -> None:
    CALLS.clear()


def launch_count() -> int:
    """Kernels of libfhb_sm100a.so launched since reset_counters() (memsets not counted)."""
    return sum(n * _KERNELS_PER_CALL.get(k, 1) for k, n in CALLS.items())


def check(rc: int, what: str) -> None:
    CALLS[what] = CALLS.get(what, 0) + 1
    if rc != 0:
        raise FhbError(f"{what} failed (rc={rc}): {lib().fhb_last_error().decode()}")


# torch.cuda.current_stream() builds a Stream object through three layers of device-index lookups (~5 us; 400 calls per
# distillation step = a fifth of the host time of a step): the raw handle comes straight from the C layer where it exists
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_RAW_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def stream_ptr() -> C.c_void_p:
    if _RAW_STREAM is not None and _RAW_DEVICE is not None:
        return C.c_void_p(_RAW_STREAM(_RAW_DEVICE()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p | None:
    return None if t is None else C.c_void_p(t.data_ptr())


def tensor3(t: torch.Tensor | None = None, *, data_ptr=None, dim=None, stride=None, offset=0, bf16=None) -> Tensor3:
    """Describe a 16-bit GEMM operand.  With `t` given (1-3 D, last dim contiguous) the description is derived;
    otherwise dim/stride (elements, dim[0] contiguous) are taken verbatim, starting `offset` elements into `t`.
    The operand FORMAT travels with the description as `.bf16` (True: bf16, a gradient; False: fp16, a forward tensor):
    taken from t.dtype, or from the `bf16` argument when only a raw pointer is given."""
    r = Tensor3()
    if t is not None:
        assert t.dtype in (torch.bfloat16, torch.float16), t.dtype
        r.bf16 = t.dtype == torch.bfloat16
    else:
        assert bf16 is not None, "a raw-pointer operand needs its format"
        r.bf16 = bool(bf16)
    if t is not None and dim is None:
        assert t.stride(-1) == 1, (t.dtype, t.stride())
        if t.dim() == 2:
            dim = (t.shape[1], t.shape[0], 1)
            stride = (t.stride(0), t.stride(0) * t.shape[0])
        else:
            assert t.dim() == 3
            dim = (t.shape[2], t.shape[1], t.shape[0])
            stride = (t.stride(1), t.stride(0))
        data_ptr = t.data_ptr()
    elif t is not None and data_ptr is None:
        data_ptr = t.data_ptr() + 2 * offset
    r.ptr = data_ptr
    # element-wise stores into the embedded arrays: building (c_int64 * 3)(*dim) objects costs 3x as much, and this
    # runs twice per GEMM launch
    d, st = r.dim, r.stride
    d[0], d[1], d[2] = dim
    st[0], st[1] = stride
    return r
