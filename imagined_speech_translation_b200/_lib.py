"""ctypes binding of libeegx.so (the C ABI declared in include/eegx.h).

There is no fallback: if the shared library is missing or the device is not a
B200 (sm_100), every compute call raises.  Build with
``python -m imagined_speech_translation_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EEGX_LIB") or os.path.join(_PKG, "libeegx.so")     # EEGX_LIB: A/B builds of the same ABI

_lib = None

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_int_p = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/eegx.h declares
# (tests/test_abi.py checks header <-> table <-> .so).
SIGNATURES = {
    "eegx_version": (C.c_int, []),
    "eegx_last_error": (C.c_char_p, []),
    "eegx_device_check": (C.c_int, []),
    "eegx_normalize_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_void_p]),
    "eegx_zscore_time_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "eegx_dsp_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                       c_f32p, C.c_int, C.c_float, C.c_float]),
    "eegx_dsp_plan_destroy": (C.c_int, [C.c_void_p]),
    "eegx_dsp_plan_dims": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "eegx_dsp_plan_kernel": (C.c_int, [C.c_void_p]),
    "eegx_dsp_plan_force_generic": (C.c_int, [C.c_void_p, C.c_int]),
    "eegx_dsp_plan_set_precise": (C.c_int, [C.c_void_p, C.c_int]),
    "eegx_dsp_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_int64, C.c_void_p]),
}


class GemmDesc(C.Structure):
    """Mirror of eegx_gemm_desc (include/eegx.h)."""
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("batch", C.c_int64),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
                ("stride_a", C.c_int64), ("stride_b", C.c_int64), ("stride_d", C.c_int64),
                ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32), ("out_f32", C.c_int32),
                ("epilogue", C.c_int32), ("accumulate", C.c_int32), ("force_block_n", C.c_int32),
                ("alpha", C.c_float), ("reserved", C.c_int32),
                ("groups", C.c_int64), ("stride_a_g", C.c_int64), ("stride_b_g", C.c_int64),
                ("stride_d_g", C.c_int64), ("stride_bias_g", C.c_int64)]


SIGNATURES["eegx_gemm_bf16"] = (C.c_int, [C.POINTER(GemmDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p])


SIGNATURES["eegx_sumsq_workspace_bytes"] = (C.c_size_t, [])
SIGNATURES["eegx_sumsq_f32"] = (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_size_t, C.c_void_p])
SIGNATURES["eegx_adamw_clip_f32"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                               C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p])


class AttnDesc(C.Structure):
    """Mirror of eegx_attn_desc (include/eegx.h)."""
    _fields_ = [("B", C.c_int64), ("H", C.c_int64), ("Sq", C.c_int64), ("Sk", C.c_int64), ("hd", C.c_int64),
                ("q_rs", C.c_int64), ("k_rs", C.c_int64), ("v_rs", C.c_int64), ("o_rs", C.c_int64),
                ("causal", C.c_int32), ("scale", C.c_float)]


_P, _I64, _U32, _F, _I, _SZ = C.c_void_p, C.c_int64, C.c_uint32, C.c_float, C.c_int, C.c_size_t
_RNG = [_P, _U32, _F]            # rng_state, site, p
SIGNATURES.update({
    "eegx_layernorm_fwd_bf16": (_I, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _I] + _RNG + [_P]),
    "eegx_layernorm_bwd_workspace_bytes": (_SZ, [_I64]),
    "eegx_layernorm_bwd_bf16": (_I, [_P] * 9 + [_I, _P, _SZ, _I64, _I64, _I64, _I64, _I] + _RNG + [_P]),
    "eegx_assemble_tokens_fwd_bf16": (_I, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, _P]),
    "eegx_assemble_tokens_bwd_bf16": (_I, [_P, _P, _I64, _I64, _I64, _P]),
    "eegx_add_dropout_fwd_bf16": (_I, [_P, _P, _P, _I64, _F] + _RNG + [_P]),
    "eegx_dropout_scale_bf16": (_I, [_P, _P, _I64, _F] + _RNG + [_P]),
    "eegx_gelu_dropout_fwd_bf16": (_I, [_P, _P, _I64] + _RNG + [_P]),
    "eegx_gelu_dropout_bwd_bf16": (_I, [_P, _P, _P, _I64] + _RNG + [_P]),
    "eegx_glu_fwd_bf16": (_I, [_P, _P, _I64, _I64] + _RNG + [_P]),
    "eegx_glu_bwd_bf16": (_I, [_P, _P, _P, _I64, _I64] + _RNG + [_P]),
    "eegx_colreduce_workspace_bytes": (_SZ, [_I64]),
    "eegx_colsum_bf16": (_I, [_P, _I64, _I64, _I64, _I64, _I64, _P, _I64, _I, _P, _SZ, _P]),
    "eegx_accumulate_partials_f32": (_I, [_P, _I64, _I64, _I64, _P, _I64, _I, _P]),
    "eegx_accumulate_conv_wgrad_f32": (_I, [_P, _I64, _I64, _I64, _I64, _P, _I, _P]),
    "eegx_bn_stats_bf16": (_I, [_P, _I64, _I64, _I64, _I64, _I64, _F, _P, _P, _P, _P, _I64, _F, _P, _SZ, _P]),
    "eegx_bn_act_fwd_bf16": (_I, [_P] * 10 + [_I, _P, _I64, _I64, _I64, _I64, _I64] + _RNG + [_P]),
    "eegx_bn_act_bwd_bf16": (_I, [_P] * 11 + [_I, _I, _P, _P, _P, _P, _SZ, _I64, _I64, _I64, _I64, _I64] + _RNG + [_P]),
    "eegx_dwconv5_fwd_bf16": (_I, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "eegx_dwconv5_bwd_bf16": (_I, [_P, _P, _P, _P, _P, _P, _SZ, _I64, _I64, _I64, _I64, _I64, _P]),
    "eegx_group_mean_bf16": (_I, [_P, _P, _I64, _I64, _I64, _I64, _P]),
    "eegx_group_mean_bwd_bf16": (_I, [_P, _P, _I64, _I64, _I64, _I64, _I, _P]),
    "eegx_se_scale_fwd_bf16": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64] + _RNG + [_P]),
    "eegx_se_scale_bwd_bf16": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64] + _RNG + [_P]),
    "eegx_nct_to_rows_bf16": (_I, [_P, _I64, _P, _I64, _I64, _I64, _I64, _P]),
    "eegx_robust_fit_f32": (_I, [_P, _I64, _I64, _F, _F, _P, _P, _P]),
    "eegx_region_std_f32": (_I, [_P, _I64, _I64, _P, _P]),
    "eegx_augment_f32": (_I, [_P, _P, _I64, _I64, _I64, _P, _P, _P, _P, _U32, _P]),
    "eegx_wake_dense_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I]),
    "eegx_wake_dense_f64": (_I, [_P] * 6 + [_I64, _I64, _I64, _I64, C.c_double, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "eegx_wake_conv2d_f64": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _P, C.c_double, _P, _P, _P]),
    "eegx_wake_maxpool_f64": (_I, [_P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P]),
    "eegx_logsoftmax_topk_f32": (_I, [_P, _I64, _I64, _I64, C.c_int32, _I64, _P, _P, _P]),
    "eegx_ce_fwd_bf16": (_I, [_P, _I64, _P, _I64, _I64, _I64, _P, _P, _P]),
    "eegx_ce_bwd_bf16": (_I, [_P, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "eegx_attn_fwd_bf16": (_I, [C.POINTER(AttnDesc), _P, _P, _P, _P, _P] + _RNG + [_P]),
    "eegx_attn_bwd_bf16": (_I, [C.POINTER(AttnDesc)] + [_P] * 9 + [_I64, _I64, _I64] + _RNG + [_P]),
    "eegx_attn_flash_fwd_bf16": (_I, [C.POINTER(AttnDesc), _P, _P, _P, _P, _P] + _RNG + [_P]),
    "eegx_attn_flash_bwd_bf16": (_I, [C.POINTER(AttnDesc)] + [_P] * 10 + [_I64, _I64, _I64] + _RNG + [_P]),
})


class EegxError(RuntimeError):
    """A libeegx entry point returned a negative status."""


def lib():
    """Load libeegx.so once; raise (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EegxError(
                f"{LIB_PATH} not found: the CUDA library is not built and there is no fallback "
                "path. Run `python -m imagined_speech_translation_b200.build`.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


CALLS = [0]   # number of libeegx compute entry-point calls (each enqueues >= 1 kernel); bench.py reads it


def check(rc: int, what: str = "") -> None:
    CALLS[0] += 1
    if rc != 0:
        msg = lib().eegx_last_error().decode("utf-8", "replace")
        raise EegxError(f"{what or 'libeegx'} failed with status {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as a void*."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
