"""ctypes binding of ``libecoflap_b200.so`` (the C ABI declared in ``include/ecoflap_b200.h``).

There is no fallback: if the shared library is missing (not built) importing this module raises, and
on a box without an sm_100 device every compute entry point returns ``ECF_ERR_NO_DEVICE`` which is
turned into an exception by :func:`check`.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libecoflap_b200.so")

ECF_F32, ECF_F16, ECF_BF16 = 0, 1, 2
OP_SQNORM, OP_ROW_SELECT, OP_LAYER_THRESH, OP_GROUP_REDUCE, OP_HESSIAN, OP_OBS, OP_GLOBAL_SELECT = range(7)
GLOBAL_MAG, GLOBAL_GRAD_MAG_ABS, GLOBAL_GRAD_MAG_SQ, GLOBAL_GRAD_ONLY = range(4)
ERR_INVALID, ERR_WORKSPACE, ERR_CUDA, ERR_NO_DEVICE, ERR_RANGE = -1, -2, -3, -4, -5

EXPORTED = (
    "ecf_version", "ecf_last_error", "ecf_device_sm_count", "ecf_workspace_bytes", "ecf_sqnorm_accum",
    "ecf_sqnorm_batched_workspace_bytes", "ecf_sqnorm_accum_batched",
    "ecf_wanda_row_select_apply", "ecf_wanda_row_select_apply_batched", "ecf_wanda_nm_select_apply", "ecf_wanda_layer_thresh_apply",
    "ecf_layer_thresh_batched_workspace_bytes", "ecf_layer_thresh_flag_offset", "ecf_wanda_layer_thresh_apply_batched", "ecf_group_reduce_chunk_elems",
    "ecf_norm_exchange_staging_bytes", "ecf_norm_exchange_p2p",
    "ecf_group_abs_reduce", "ecf_zo_perturb", "ecf_count_zero", "ecf_hessian_accum", "ecf_obs_prune",
    "ecf_global_chunk_elems", "ecf_global_select", "ecf_global_apply", "ecf_grad_accum", "ecf_global_score_sum",
)


class EcfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ecoflap_b200 error {code}: {msg}")
        self.code = code


class TensorDesc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("numel", C.c_int64), ("dtype", C.c_int32), ("reserved", C.c_int32),
                ("chunk_begin", C.c_int64)]


class GlobalDesc(C.Structure):
    _fields_ = [("W", C.c_void_p), ("G", C.c_void_p), ("numel", C.c_int64), ("dtype", C.c_int32), ("reserved", C.c_int32),
                ("chunk_begin", C.c_int64)]


class SqnormDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("scaler_row", C.c_void_p), ("T", C.c_int64), ("C", C.c_int64), ("ld", C.c_int64),
                ("dtype", C.c_int32), ("rescale", C.c_float), ("inv_n", C.c_float)]


class LayerDesc(C.Structure):
    _fields_ = [("W", C.c_void_p), ("scaler_row", C.c_void_p), ("R", C.c_int64), ("C", C.c_int64), ("ld", C.c_int64),
                ("dtype", C.c_int32), ("kth_index", C.c_int64), ("thres_out", C.c_void_p), ("mask_bits", C.c_void_p),
                ("mask_ld", C.c_int64), ("n_zero", C.c_void_p)]


class RowDesc(C.Structure):
    _fields_ = [("W", C.c_void_p), ("scaler_row", C.c_void_p), ("R", C.c_int64), ("C", C.c_int64), ("ld", C.c_int64),
                ("dtype", C.c_int32), ("k_per_row", C.c_int64), ("mask_bits", C.c_void_p), ("mask_ld", C.c_int64),
                ("n_zero", C.c_void_p)]


LAYER_MAX_BATCH = 8
ROW_MAX_BATCH = 16
SQNORM_MAX_BATCH = 256
SQNORM_MAX_GROUPS = 32


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ecoflap_b200.build` "
            "(ecoflap_b200 has no CPU/PyTorch fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, f64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_size_t
    sig = {
        "ecf_version": (i32, []),
        "ecf_last_error": (C.c_char_p, []),
        "ecf_device_sm_count": (i32, []),
        "ecf_workspace_bytes": (sz, [i32, i64, i64]),
        "ecf_sqnorm_accum": (i32, [vp, i32, i64, i64, i64, vp, f32, f32, vp, sz, vp]),
        "ecf_sqnorm_batched_workspace_bytes": (sz, [C.POINTER(SqnormDesc), i32]),
        "ecf_sqnorm_accum_batched": (i32, [C.POINTER(SqnormDesc), i32, vp, sz, vp]),
        "ecf_wanda_row_select_apply": (i32, [vp, i32, i64, i64, i64, vp, i64, vp, i64, vp, vp, sz, vp]),
        "ecf_wanda_row_select_apply_batched": (i32, [C.POINTER(RowDesc), i32, vp, sz, vp]),
        "ecf_wanda_nm_select_apply": (i32, [vp, i32, i64, i64, i64, vp, i32, i32, vp, i64, vp, vp]),
        "ecf_wanda_layer_thresh_apply": (i32, [vp, i32, i64, i64, i64, vp, i64, vp, vp, i64, vp, vp, sz, vp]),
        "ecf_layer_thresh_batched_workspace_bytes": (sz, [C.POINTER(LayerDesc), i32]),
        "ecf_layer_thresh_flag_offset": (sz, []),
        "ecf_wanda_layer_thresh_apply_batched": (i32, [C.POINTER(LayerDesc), i32, vp, sz, vp]),
        "ecf_norm_exchange_staging_bytes": (sz, [i64]),
        "ecf_norm_exchange_p2p": (i32, [vp, i64, C.POINTER(C.c_void_p), i64, i32, i32, vp]),
        "ecf_group_reduce_chunk_elems": (i64, []),
        "ecf_group_abs_reduce": (i32, [vp, i32, i64, vp, vp, vp, sz, vp]),
        "ecf_zo_perturb": (i32, [vp, i32, i64, vp, f64, f64, vp]),
        "ecf_count_zero": (i32, [vp, i32, i64, vp, vp]),
        "ecf_hessian_accum": (i32, [vp, i32, i64, i64, i64, vp, i64, f32, f32, vp, sz, vp]),
        "ecf_obs_prune": (i32, [vp, i64, i64, i64, vp, i64, C.POINTER(C.c_int64), i32, i32, i32, vp, sz, vp]),
        "ecf_global_chunk_elems": (i64, []),
        "ecf_global_select": (i32, [vp, i32, i64, i32, f64, i32, vp, vp, vp, vp, sz, vp]),
        "ecf_global_apply": (i32, [vp, i32, i64, i32, f64, i32, vp, vp, vp, vp]),
        "ecf_grad_accum": (i32, [vp, vp, i32, i64, i32, vp]),
        "ecf_global_score_sum": (i32, [vp, i32, i64, i32, f64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(code):
    if code != 0:
        raise EcfError(code, lib.ecf_last_error().decode("utf-8", "replace"))
    return code
