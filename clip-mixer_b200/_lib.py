"""ctypes binding of libmixerclip.so (C ABI declared in include/mixerclip.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MC_LIB selects another build of the same library (e.g. the -DTM_TRACE debug build of tools/tokenmix_trace.sh)
LIB_PATH = os.environ.get("MC_LIB") or os.path.join(_HERE, "libmixerclip.so")

F32, BF16 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
BIAS_NONE, BIAS_N, BIAS_M = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_GELU_BWD = 0, 1, 2


class MixerClipError(RuntimeError):
    pass


class GemmParams(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("batch", C.c_int64),
        ("A", C.c_void_p), ("a_major", C.c_int32), ("lda", C.c_int64), ("a_batch_stride", C.c_int64),
        ("B", C.c_void_p), ("b_major", C.c_int32), ("ldb", C.c_int64), ("b_batch_stride", C.c_int64),
        ("k_spans_batch", C.c_int32),
        ("C", C.c_void_p), ("c_dtype", C.c_int32), ("ldc", C.c_int64), ("c_batch_stride", C.c_int64),
        ("accumulate", C.c_int32), ("split_k", C.c_int32), ("row_remap", C.c_int32),
        ("bias", C.c_void_p), ("bias_mode", C.c_int32),
        ("zout", C.c_void_p), ("ldz", C.c_int64), ("z_batch_stride", C.c_int64),
        ("zin", C.c_void_p), ("ldzin", C.c_int64), ("zin_batch_stride", C.c_int64),
        ("act", C.c_int32),
        ("R", C.c_void_p), ("ldr", C.c_int64), ("r_batch_stride", C.c_int64),
        ("rowsum_out", C.c_void_p), ("c_transposed", C.c_int32),
        ("A2", C.c_void_p), ("a2_major", C.c_int32), ("lda2", C.c_int64), ("a2_batch_stride", C.c_int64),
        ("B2", C.c_void_p), ("b2_major", C.c_int32), ("ldb2", C.c_int64), ("b2_batch_stride", C.c_int64),
        ("bias2", C.c_void_p),
        ("rowstat_out", C.c_void_p),
    ]


class TokenMixParams(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("P", C.c_int64), ("D", C.c_int64),
        ("u", C.c_void_p),
        ("w1", C.c_void_p), ("ld1", C.c_int64), ("w1t", C.c_void_p), ("ld1t", C.c_int64), ("b1", C.c_void_p),
        ("w2", C.c_void_p), ("ld2", C.c_int64), ("b2", C.c_void_p),
        ("x", C.c_void_p), ("y", C.c_void_p), ("dy", C.c_void_p),
        ("gw1", C.c_void_p), ("ldg1", C.c_int64), ("gw2", C.c_void_p), ("ldg2", C.c_int64), ("gb1", C.c_void_p),
        ("ln_sums", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("u_out", C.c_void_p),
        ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p),
    ]


_I64, _I32, _F, _P = C.c_int64, C.c_int32, C.c_float, C.c_void_p

# name -> argument types (all return int unless noted); mirrors include/mixerclip.h
SIGNATURES = {
    "mc_version": [],
    "mc_device_info": [_P, _P, _P],
    "mc_set_sm_limit": [_I32],
    "mc_gemm_bf16_tc": [C.POINTER(GemmParams), _P],
    "mc_gemm_f32_simt": [C.POINTER(GemmParams), _P],
    "mc_token_mix_supported": [_I64, _I64],
    "mc_token_mix_fwd": [C.POINTER(TokenMixParams), _P],
    "mc_token_mix_dgrad": [C.POINTER(TokenMixParams), _P],
    "mc_token_mix_wgrad": [C.POINTER(TokenMixParams), _P],
    "mc_ln_fwd": [_P, _I64, _P, _P, _I64, _P, _P, _P, _I32, _I64, _P, _P, _I64, _I64, _P],
    "mc_ln_bwd": [_P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P, _I32, _P, _P, _P, _P, _I64, _P, _I64,
                  _I64, _P],
    "mc_colsum": [_P, _I32, _I64, _I64, _I64, _P, _P],
    "mc_rowsum": [_P, _I32, _I64, _I64, _I64, _I64, _P, _P],
    "mc_cast_pad": [_P, _I64, _I64, _I64, _P, _I32, _I64, _P],
    "mc_transpose_bf16": [_P, _I64, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P],
    "mc_im2col": [_P, _I32, _I64, _I64, _I64, _P, _I32, _P],
    "mc_embed_fwd": [_P, _P, _P, _I64, _I64, _I64, _I64, _P],
    "mc_embed_bwd": [_P, _P, _P, _I64, _I64, _I64, _I64, _P],
    "mc_eot_rows": [_P, _P, _I64, _I64, _P],
    "mc_l2norm_fwd": [_P, _P, _P, _I64, _I64, _P],
    "mc_l2norm_bwd": [_P, _P, _P, _P, _P, _I32, _I64, _I64, _P],
    "mc_head_fwd_bwd": [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _P, _P, _P, _P, _P, _I64, _P],
    "mc_sumsq": [_P, _I64, _P, _P],
    "mc_sched_step": [_P, _P, _I64, C.c_double, C.c_double, _I64, C.c_double, C.c_double, C.c_double, _P],
    "mc_adamw": [_P, _P, _P, _P, _P, _P, _I64, _P, _P, _F, _F, _F, _F, _F, _F, _P],
}
SPECIAL_RESTYPE = {"mc_last_error": C.c_char_p, "mc_head_workspace_bytes": C.c_int64}
SPECIAL_ARGTYPES = {"mc_last_error": [], "mc_head_workspace_bytes": [_I64, _I64, _I64]}

_lib = None


def load():
    """Load libmixerclip.so (built in-tree by __graft_entry__.build()); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MixerClipError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name, res in SPECIAL_RESTYPE.items():
        fn = getattr(lib, name)
        fn.argtypes = SPECIAL_ARGTYPES[name]
        fn.restype = res
    _lib = lib
    return lib


def exported_symbols():
    return list(SIGNATURES) + list(SPECIAL_RESTYPE)


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mc_last_error()
        raise MixerClipError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
