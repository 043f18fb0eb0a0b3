"""ctypes binding of the C ABI declared in include/plda_b200.h.

The library is loaded lazily and LOUDLY: if ``plda_b200/lib/libplda_b200.so`` is
missing the product path raises -- there is no CPU fallback (and nothing here
imports ``oracle/``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libplda_b200.so")

F64, F32 = 0, 1
HOST, DEVICE = 0, 1
PREC_BF16X3, PREC_FP64 = 0, 1

E_INVALID, E_CUDA, E_NOTFITTED, E_VALUE, E_INTERNAL = -1, -2, -3, -4, -5

_i64 = C.c_int64
_vp = C.c_void_p
_int = C.c_int

# name -> argtypes (restype is always int unless noted); mirrors include/plda_b200.h
SIGNATURES = {
    "plda_create": [_int, C.POINTER(_vp)],
    "plda_destroy": [_vp],
    "plda_set_precision": [_vp, _int],
    "plda_set_stream": [_vp, _vp],
    "plda_synchronize": [_vp],
    "plda_stream_wait": [_vp, _vp],
    "plda_launch_count": [_vp, C.POINTER(_i64)],
    "plda_profile_gemm": [_vp, _int],
    "plda_profile_collect": [_vp, C.POINTER(C.c_double), C.POINTER(_i64)],
    "plda_set_allreduce": [_vp, _vp, _vp, _vp, _i64],
    "plda_fit": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _int],
    "plda_fit_labels": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _int, _int],
    "plda_fit_timings": [_vp, C.POINTER(C.c_double)],
    "plda_dim": [_vp, C.POINTER(_i64)],
    "plda_get_model": [_vp, _vp, _vp, _vp],
    "plda_set_model": [_vp, _i64, _vp, _vp, _vp],
    "plda_get_covariances": [_vp, _vp, _vp],
    "plda_smooth": [_vp, C.c_double],
    "plda_transform": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i64)],
    "plda_transform_rows": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, C.c_int32, _i64, _vp, _i64, _int, _int],
    "plda_score_pair": [_vp, C.c_uint64, _i64, _vp, _vp, _i64, C.POINTER(C.c_float)],
    "plda_score_grid": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _i64, _int],
    "plda_norm": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _i64, _i64, _i64, _int, _int, _i64, C.c_uint64],
    "plda_norm_rows": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _i64, _i64, _i64, _int, _int, _i64, C.c_uint64,
                       _vp, _vp, _int],
    "plda_norm_selection": [_i64, _i64, C.c_uint64, _vp],
    "plda_score_grid_z": [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _i64, _int, _vp, _vp, _int],
    "plda_score_trials": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _i64, _int, _vp,
                          _int, _vp, _vp, _int, _int],
    "plda_score_hist": [_vp, _vp, _i64, _i64, C.c_int32, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _int, C.c_double,
                        C.c_double, _int, C.c_double, _vp, _vp, _int, _vp, _vp, _vp, _int],
    "plda_znorm_size": [_vp, C.POINTER(_i64)],
    "plda_znorm_get": [_vp, _vp, _vp, _vp, _i64, C.POINTER(_i64)],
    "plda_znorm_clear": [_vp],
    "plda_znorm_set": [_vp, _vp, _vp, _vp, _i64],
    "plda_shard_open": [_vp, _int, _int, _vp, _i64, _vp, C.POINTER(_vp)],
    "plda_shard_connect": [_vp, _int, _vp, _vp],
    "plda_shard_push": [_vp, _vp, _i64, _i64, _int, _int],
    "plda_shard_score": [_vp, _vp, _i64, _i64, _int, _vp, _int, _vp, _i64],
    "plda_shard_step": [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _int, _vp, _i64],
    "plda_shard_open_ragged": [_vp, _int, _int, _vp, _i64, _int, _vp, C.POINTER(_vp)],
    "plda_shard_step_ragged": [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _int, _vp, _int, _vp, _i64],
    "plda_shard_status": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "plda_shard_close": [_vp],
    "plda_dvector_pool": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _i64, _int, _int, _vp, _i64, _int],
    "lda_create": [_int, C.POINTER(_vp)],
    "lda_destroy": [_vp],
    "lda_set_precision": [_vp, _int],
    "lda_launch_count": [_vp, C.POINTER(_i64)],
    "lda_synchronize": [_vp],
    "lda_stream_wait": [_vp, _vp],
    "lda_fit_svd": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _i64],
    "lda_fit_lsqr": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _i64],
    "lda_fit_eigen": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _i64],
    "lda_get_eigenvalues": [_vp, _vp, _i64, C.POINTER(_i64)],
    "lda_class_stats": [_vp, _vp, _i64, _i64, _i64, _int, _int, _vp, C.POINTER(_i64)],
    "lda_get_class_stats": [_vp, _vp, _vp, _vp, _vp],
    "lda_fit_from_stats": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64],
    "lda_get_svd": [_vp, C.POINTER(_i64), _vp, _vp],
    "lda_transform": [_vp, _vp, _i64, _i64, _i64, _int, _int, _i64, _vp, _i64, _int],
    "lda_num_classes": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "lda_get_coef": [_vp, _vp, _vp, _vp],
    "lda_set_coef": [_vp, _i64, _i64, _vp, _vp],
    "lda_predict": [_vp, _vp, _i64, _i64, _i64, _int, _int, _int, _vp, _i64, _int],
    "plda_device_malloc": [_int, C.c_size_t, C.POINTER(_vp)],
    "plda_device_free": [_int, _vp],
    "plda_host_malloc_pinned": [C.c_size_t, C.POINTER(_vp)],
    "plda_host_free_pinned": [_vp],
    "plda_memcpy": [_vp, _vp, C.c_size_t, _int],
    "plda_test_gemm": [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp],
    "plda_debug_counters": [_vp, _vp, _int],
    "plda_test_scatter": [_vp, _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _i64, C.POINTER(_i64)],
    "plda_test_linalg": [_vp, _int, _vp, _i64, _vp, _vp],
}
STRING_FUNCS = ("plda_last_error", "plda_version")

_lib = None
_lock = threading.Lock()


class PldaB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PldaB200Error(
                "plda_b200: %s is missing -- build it with `python -m plda_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _int
        for name in STRING_FUNCS:
            fn = getattr(l, name)
            fn.argtypes = []
            fn.restype = C.c_char_p
        _lib = l
    return _lib


def last_error() -> str:
    return lib().plda_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Map C status codes to the exceptions the reference raises
    (ValueError via PyErr_SetString, src/pldamodule.cpp:56,60,84,130,134)."""
    if status == 0:
        return
    msg = last_error()
    if status in (E_VALUE, E_INVALID):
        raise ValueError(msg)
    if status == E_NOTFITTED:
        raise ValueError(msg)
    raise PldaB200Error("plda_b200 error %d: %s" % (status, msg))


def ptr(a) -> _vp:
    """void* of a numpy array (host) -- the array must stay alive during the call."""
    if a is None:
        return _vp(None)
    return _vp(a.ctypes.data)


def as_matrix(x, name="features"):
    """Validate a host feature matrix like the reference does (float dtype required,
    src/pldamodule.cpp:59-62) and return (array, dtype_code) C-contiguous f64 or f32."""
    x = np.asarray(x)
    if x.dtype.kind != "f":
        raise ValueError("Given Input %s (argument 1) are not floats! Set the dtype to float!" % name)
    if x.ndim != 2:
        raise ValueError("%s must be a 2-D array (n_samples, featdim)" % name)
    if x.dtype == np.float32:
        return np.ascontiguousarray(x), F32
    return np.ascontiguousarray(x, dtype=np.float64), F64


def as_labels(y, n=None, require_unsigned=False):
    """Labels -> uint64 host vector.  The reference insists on an unsigned dtype
    (src/pldamodule.cpp:55-58,133-136) and rejects strings (:128-131); we accept any
    non-negative integer dtype (superset) unless require_unsigned."""
    y = np.asarray(y)
    if y.dtype.kind in ("S", "U", "O"):
        raise ValueError("Labels need to be numpy array of uints, not strings!")
    if y.dtype.kind == "u":
        out = y.astype(np.uint64, copy=False)
    elif y.dtype.kind == "i" and not require_unsigned:
        if y.size and y.min() < 0:
            raise ValueError("Given labels (argument 2) are not an unsigned! Set the dtype to uint!")
        out = y.astype(np.uint64)
    elif y.dtype.kind == "b" and not require_unsigned:
        out = y.astype(np.uint64)
    else:
        raise ValueError("Given labels (argument 2) are not an unsigned! Set the dtype to uint!")
    out = np.ascontiguousarray(out.reshape(-1))
    if n is not None and out.shape[0] != n:
        raise ValueError("labels and features disagree on the number of samples")
    return out
