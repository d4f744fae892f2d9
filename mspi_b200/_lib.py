"""ctypes binding of the C-ABI shared library (include/mspi_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``mspi_b200/lib/libmspi_b200.so``.  There is no fallback: if the library is missing the
import of any compute path raises, and every compute call fails loudly without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmspi_b200.so")
if os.environ.get("MSPI_LIB"):   # tuning aid: a library variant built with other compile-time switches (same ABI)
    LIB_PATH = os.path.abspath(os.environ["MSPI_LIB"])

MSPI_BF16, MSPI_F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID, ACT_SWISH = 0, 1, 2, 3, 4
MAX_TAPS = 64


class MspiError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("a_dtype", C.c_int32),
        ("a_dims", C.c_int32 * 5),
        ("a_strides", C.c_int64 * 5),
        ("box", C.c_int32 * 5),
        ("ntaps", C.c_int32),
        ("tap_off", (C.c_int32 * 4) * MAX_TAPS),
        ("cin_pad", C.c_int32),
        ("cout", C.c_int32),
        ("w_rows", C.c_int32),
        ("bn", C.c_int32),
        ("o_dims", C.c_int32 * 4),
        ("o_strides", C.c_int64 * 4),
        ("r_strides", C.c_int64 * 4),
        ("o_dtype", C.c_int32),
        ("r_dtype", C.c_int32),
        ("act", C.c_int32),
        ("has_residual", C.c_int32),
        ("res_after_act", C.c_int32),
        ("k_row_bytes", C.c_int32),
        ("w_batch_dims", C.c_int32 * 2),
        ("w_strides", C.c_int64 * 3),
    ]


class PatchDesc(C.Structure):
    _fields_ = [
        ("src_layout", C.c_int32),
        ("n", C.c_int32), ("c", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("src_cstride", C.c_int64),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("pt", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32),
        ("ot", C.c_int32), ("oh", C.c_int32), ("ow", C.c_int32),
        ("k_pad", C.c_int32),
    ]


class PoolDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("in_cstride", C.c_int64), ("out_cstride", C.c_int64),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("pt", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32),
        ("ot", C.c_int32), ("oh", C.c_int32), ("ow", C.c_int32),
    ]


class UpDesc(C.Structure):
    _fields_ = [
        ("nt", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("k", C.c_int32),
        ("in_cstride", C.c_int64), ("out_cstride", C.c_int64),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32), ("accumulate", C.c_int32), ("act", C.c_int32),
    ]


class DwDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("ln_eps", C.c_float), ("out_dtype", C.c_int32), ("in_dtype", C.c_int32),
    ]


class Dw3dDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("in_cstride", C.c_int64), ("out_cstride", C.c_int64),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("oh", C.c_int32), ("ow", C.c_int32), ("act", C.c_int32),
    ]


class LnDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("c", C.c_int32),
        ("in_rstride", C.c_int64), ("out_rstride", C.c_int64),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("eps", C.c_float), ("relu", C.c_int32), ("pos_rows", C.c_int32),
        ("rows_per_group", C.c_int64), ("out_gstride", C.c_int64),
    ]


class SgemmDesc(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("batch0", C.c_int32), ("batch1", C.c_int32),
        ("accumulate", C.c_int32), ("alpha", C.c_float),
        ("a_strides", C.c_int64 * 4), ("b_strides", C.c_int64 * 4), ("c_strides", C.c_int64 * 4),
    ]


class PermDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32 * 4), ("src_strides", C.c_int64 * 4), ("dst_strides", C.c_int64 * 4),
        ("dst_dtype", C.c_int32), ("accumulate", C.c_int32),
    ]


_P = C.c_void_p
_I, _L, _F = C.c_int, C.c_int64, C.c_float
_SIGNATURES = {
    "mspi_last_error": (C.c_char_p, []),
    "mspi_version": (C.c_int, []),
    "mspi_arch": (C.c_char_p, []),
    "mspi_launch_count": (C.c_int64, []),
    "mspi_set_pdl": (C.c_int, [C.c_int]),
    "mspi_debug_dw_phase_cycles": (C.c_int, [C.c_void_p, C.c_int]),
    "mspi_debug_gemm_epilogue_cycles": (C.c_int, [C.c_void_p, C.c_int]),
    "mspi_conv_gemm": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "mspi_conv133_small": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_conv_gemm_ln": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, C.c_float, C.c_int, _P, _P]),
    "mspi_conv_wgrad": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "mspi_mlp_fused": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _P]),
    "mspi_mlp_fused_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P,
                                    C.c_float, _P]),
    "mspi_patch_gather": (C.c_int, [C.POINTER(PatchDesc), _P, _P, _P]),
    "mspi_ncdhw_to_ndhwc": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int64, _P]),
    "mspi_clip_to_padded_nhwc4": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_clip_frames_to_padded_nhwc4": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   _P, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_clip_u8_to_padded_nhwc4": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "mspi_gather_rows": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "mspi_ndhwc_to_ncdhw": (C.c_int, [_P, C.c_int, C.c_int64, _P, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_maxpool3d": (C.c_int, [C.POINTER(PoolDesc), _P, _P, _P]),
    "mspi_upsample_bilinear": (C.c_int, [C.POINTER(UpDesc), _P, _P, _P]),
    "mspi_dwconv_ln": (C.c_int, [C.POINTER(DwDesc), _P, _P, _P, _P, _P, _P, _P]),
    "mspi_dwconv3d_bn": (C.c_int, [C.POINTER(Dw3dDesc), _P, _P, _P, _P, _P]),
    "mspi_dwconv3d_bn_mean": (C.c_int, [C.POINTER(Dw3dDesc), _P, _P, _P, _P, _P, _P, _P]),
    "mspi_channel_mean": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, C.c_int64, _P]),
    "mspi_se_gate": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_scale_act": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, C.c_int, C.c_int, _P]),
    "mspi_layernorm": (C.c_int, [C.POINTER(LnDesc), _P, _P, _P, _P, _P, _P]),
    "mspi_attention": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "mspi_softmax_rows": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int64, _P]),
    "mspi_transpose_v": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_sa_gate": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, _P]),
    "mspi_sa_gate_fused": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "mspi_add_bf16": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "mspi_token_mean": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_cast_rows": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_simsiam_loss": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "mspi_logsoftmax2d": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P]),
    "mspi_saliency_metrics": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int64, _P]),
    "mspi_postprocess_maps": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mspi_logspec": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    # training step
    "mspi_bn_train_fwd": (C.c_int, [_P, _L, _P, _L, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P]),
    "mspi_bn_train_bwd": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _L, _L, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mspi_act_fwd": (C.c_int, [_P, _P, _L, _I, _P]),
    "mspi_act_bwd": (C.c_int, [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P, _P]),
    "mspi_maxpool3d_f32": (C.c_int, [C.POINTER(PoolDesc), _P, _P, _P]),
    "mspi_maxpool3d_bwd": (C.c_int, [C.POINTER(PoolDesc), _P, _P, _L, _P, _L, _P]),
    "mspi_upsample_bilinear_bwd": (C.c_int, [C.POINTER(UpDesc), _P, _P, _P, _P]),
    "mspi_layernorm_bwd": (C.c_int, [_P, _L, _P, _L, _L, _L, _P, _P, _F, _P, _L, _L, _I, _I, _P, _P, _P]),
    "mspi_softmax_bwd_rows": (C.c_int, [_P, _P, _L, _I, _L, _F, _P]),
    "mspi_sgemm_strided": (C.c_int, [C.POINTER(SgemmDesc), _P, _P, _P, _P]),
    "mspi_sa_gate_bwd": (C.c_int, [_P, _L, _P, _P, _L, _P, _L, _P, _L, _I, _I, _P]),
    "mspi_conv_c1_fwd": (C.c_int, [_P, _I, _L, _P, _P, _P, _L, _I, _I, _I, _P]),
    "mspi_conv_c1_bwd": (C.c_int, [_P, _L, _P, _P, _P, _L, _P, _P, _L, _I, _I, _I, _I, _P]),
    "mspi_dwconv_wgrad": (C.c_int, [C.POINTER(DwDesc), _P, _P, _P, _P, _P]),
    "mspi_salloss_bwd": (C.c_int, [_P, _P, _P, _F, _P, _P, _P, _I, _L, _F, _P]),
    "mspi_simsiam_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "mspi_token_mean_bwd": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mspi_add_rows": (C.c_int, [_P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _P]),
    "mspi_permute_copy": (C.c_int, [C.POINTER(PermDesc), _P, _P, _P]),
    "mspi_permute_copy_batched": (C.c_int, [_P, _P, _P, _I, _P]),
    "mspi_adamw_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P, _F, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load (once) and return the shared library; raises MspiError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MspiError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "mspi_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mspi_last_error().decode("utf-8", "replace")
        raise MspiError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().mspi_launch_count())
