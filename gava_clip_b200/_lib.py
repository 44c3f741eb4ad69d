"""ctypes binding of libgava_sm100.so (C ABI in include/gava_sm100.h).

There is no fallback: if the shared object is missing or a symbol is absent this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["GAVA_LIB_PATH"]) if os.environ.get("GAVA_LIB_PATH") else _PKG / "lib" / "libgava_sm100.so"

EPI_BIAS_BF16 = 0
EPI_BIAS_GELU_BF16 = 1
EPI_BIAS_RESID_F32 = 2
EPI_BIAS_F32 = 3
EPI_PATCH_F32 = 4
EPI_DGELU_BF16 = 5
EPI_LNFOLD_BF16 = 7
EPI_BIAS_RESID_STATS_F32 = 8


class GemmEpilogue(C.Structure):
    """mirror of `struct gava_gemm_epilogue`"""
    _fields_ = [
        ("mode", C.c_int32),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p),
        ("ldr", C.c_int64),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("group", C.c_int32),
        ("group_stride", C.c_int32),
        ("row_offset", C.c_int32),
        ("T", C.c_int32),
        ("pos", C.c_void_p),
        ("time_embed", C.c_void_p),
        ("aux", C.c_void_p),
        ("ld_aux", C.c_int64),
        ("aux_out", C.c_void_p),
        ("max_ctas", C.c_int32),
        ("copy_out", C.c_void_p),
        ("ld_copy", C.c_int64),
        ("stats_out", C.c_void_p),
        ("stats_in", C.c_void_p),
        ("ld_stats", C.c_int64),
        ("colsum", C.c_void_p),
        ("n_slots", C.c_int32),
        ("eps", C.c_float),
    ]


_p, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES); must match include/gava_sm100.h
SIGNATURES = {
    "gava_version": [],
    "gava_last_error_string": [],
    "gava_init": [],
    "gava_workspace_bytes": [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32],
    "gava_gemm_bf16_fwd": [_p, _i64, _p, _i64, _i64, _i64, _i64, C.POINTER(GemmEpilogue), _p],
    "gava_gemm_stats_slots": [_i64, _i64],
    "gava_layernorm_fwd": [_p, _i64, _p, _p, _p, _i64, _i32, _i64, _i32, _f32, _p],
    "gava_layernorm2_fwd": [_p, _i64, _p, _p, _p, _i64, _p, _p, _p, _i64, _i64, _i32, _f32, _p],
    "gava_layernorm_resid_fwd": [_p, _i64, _p, _i64, _p, _p, _p, _i64, _i32, _i64, _i32, _f32, _p],
    "gava_patch_gather_fwd": [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_patch_gather_u8_fwd": [_p, _p, _p, C.POINTER(C.c_float), C.POINTER(C.c_float), _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_cls_rows_fwd": [_p, _i32, _p, _p, _p, _i32, _i32, _i32, _p],
    "gava_temporal_attention_fwd": [_p, _p, _i32, _i32, _i32, _p],
    "gava_prompt_rows_fwd": [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _f32, _p],
    "gava_attention_fwd": [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_head_fwd": [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _p],
    "gava_split_bf16_fwd": [_p, _i64, _p, _p, _i64, _i64, _i32, _p],
    "gava_split3_bf16_fwd": [_p, _i64, _p, _i64, _i32, _p],
    "gava_patch_gather_split_fwd": [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_embed_assemble_fwd": [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p],
    "gava_quickgelu_f32_fwd": [_p, _i64, _p],
    "gava_attention_f32_fwd": [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_causal_attention_f32_fwd": [_p, _p, _i32, _i32, _i32, _p],
    "gava_prompt_rows_f32_fwd": [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _f32, _p],
    "gava_layernorm_bwd": [_p, _i64, _i32, _p, _i64, _p, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _i32, _f32, _p],
    "gava_attention_train_fwd": [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_attention_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    "gava_temporal_attention_bwd": [_p, _p, _p, _i32, _i32, _i32, _p],
    "gava_causal_attention_bwd": [_p, _p, _p, _i32, _i32, _i32, _p],
    "gava_prompt_rows_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _f32, _p],
    "gava_group_sum": [_p, _i64, _i32, _p, _i32, _i32, _i32, _i32, _p],
    "gava_transpose_bf16": [_p, _i64, _i32, _p, _i32, _i32, _i32, _p],
    "gava_quickgelu_bf16_fwd": [_p, _p, _i64, _p],
    "gava_preprocess_u8_fwd": [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                               C.POINTER(C.c_float), C.POINTER(C.c_float), _p],
    "gava_adamw_step": [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _p],
}
_RESTYPES = {"gava_last_error_string": C.c_char_p, "gava_workspace_bytes": C.c_int64}

_lib = None


def load() -> C.CDLL:
    """Load the library (building it first if only sources are present and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() and os.environ.get("GAVA_NO_AUTOBUILD") != "1":
        from . import build as _build
        _build.build()
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m gava_clip_b200.build` (no fallback path exists)")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing -> loud
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.gava_version() != 4:
        raise RuntimeError(f"libgava_sm100 ABI version {lib.gava_version()} != 4")
    _lib = lib
    return lib


def last_error() -> str:
    s = load().gava_last_error_string()
    return s.decode() if s else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")
