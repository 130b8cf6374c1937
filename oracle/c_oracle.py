"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_build/libecf_oracle.so (the plain-C / OpenMP oracle).
Builds it on first use with the Makefile next to this file."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libecf_oracle.so")
DT = {"fp32": 0, "fp16": 1, "bf16": 2}


def build():
    src = os.path.join(HERE, "ecoflap_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        # -march=native would not survive the trip to another CPU: build generic
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", LIB, src, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.ecf_ref_sqnorm_accum.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]
        _lib.ecf_ref_sqnorm_accum.restype = None
        _lib.ecf_ref_wanda_row_prune.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
        _lib.ecf_ref_wanda_row_prune.restype = None
        _lib.ecf_ref_wanda_layer_prune.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
        _lib.ecf_ref_wanda_layer_prune.restype = C.c_float
        _lib.ecf_ref_set_threads.argtypes = [C.c_int]
        _lib.ecf_ref_set_threads.restype = None
        _lib.ecf_ref_max_threads.restype = C.c_int
    return _lib


def set_threads(n):
    """OpenMP threads of the C oracle, set explicitly (an inherited OMP_NUM_THREADS=1 would silently serialise it)."""
    lib().ecf_ref_set_threads(int(n))
    return int(lib().ecf_ref_max_threads())


def to_storage(a_f32, dtype):
    """float32 array (exactly representable values) -> storage array of the given dtype (uint16 patterns)."""
    a = np.ascontiguousarray(a_f32, dtype=np.float32)
    if dtype == "fp32":
        return a.copy()
    if dtype == "fp16":
        return a.astype(np.float16).view(np.uint16).copy()
    return (a.view(np.uint32) >> 16).astype(np.uint16)


def from_storage(s, dtype):
    if dtype == "fp32":
        return s.astype(np.float32)
    if dtype == "fp16":
        return s.view(np.float16).astype(np.float32)
    return (s.astype(np.uint32) << 16).view(np.float32)


def sqnorm_accum(x_store, dtype, scaler_row, n_old, b):
    T, Cc = x_store.reshape(-1, x_store.shape[-1]).shape
    lib().ecf_ref_sqnorm_accum(x_store.ctypes.data, DT[dtype], T, Cc, scaler_row.ctypes.data, n_old, b)


def wanda_row_prune(W_store, dtype, scaler_row, k):
    R, Cc = W_store.shape
    lib().ecf_ref_wanda_row_prune(W_store.ctypes.data, DT[dtype], R, Cc, scaler_row.ctypes.data, int(k))


def wanda_layer_prune(W_store, dtype, scaler_row, idx):
    R, Cc = W_store.shape
    return float(lib().ecf_ref_wanda_layer_prune(W_store.ctypes.data, DT[dtype], R, Cc, scaler_row.ctypes.data, int(idx)))
