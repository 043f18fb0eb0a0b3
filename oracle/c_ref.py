"""ctypes wrapper of oracle/plda_ref.c (the C restatement timed as the CPU baseline).
TEST / BASELINE INFRASTRUCTURE -- never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libplda_ref.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "plda_ref.c")):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        l.plda_ref_stats.restype = C.c_void_p
        l.plda_ref_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]
        l.plda_ref_em_iter.argtypes = [C.c_void_p]
        l.plda_ref_get_output.argtypes = [C.c_void_p] * 4
        l.plda_ref_get_covariances.argtypes = [C.c_void_p] * 3
        l.plda_ref_free.argtypes = [C.c_void_p]
        l.plda_ref_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        l.plda_ref_llr.restype = C.c_double
        l.plda_ref_llr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        l.plda_ref_score_grid.restype = C.c_int
        l.plda_ref_score_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_int]
        l.plda_ref_max_threads.restype = C.c_int
        _lib = l
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class RefPlda:
    """stats -> em_iter()* -> output(); mirrors MPlda_fit (src/pldamodule.cpp:42-109)."""

    def __init__(self, x, labels):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        lab = np.ascontiguousarray(labels).astype(np.int64)
        self.k = int(lab.max()) + 1
        self.d = self.x.shape[1]
        self._lab = lab
        self._s = C.c_void_p(lib().plda_ref_stats(_p(self.x), self.x.shape[0], self.d, _p(lab), self.k))

    def em_iter(self):
        lib().plda_ref_em_iter(self._s)

    def output(self):
        mean = np.empty(self.d)
        tr = np.empty((self.d, self.d))
        psi = np.empty(self.d)
        lib().plda_ref_get_output(self._s, _p(mean), _p(tr), _p(psi))
        return mean, tr, psi

    def covariances(self):
        w = np.empty((self.d, self.d))
        b = np.empty((self.d, self.d))
        lib().plda_ref_get_covariances(self._s, _p(w), _p(b))
        return w, b

    def __del__(self):
        if getattr(self, "_s", None):
            lib().plda_ref_free(self._s)
            self._s = None


def transform(mean, tr, psi, x, n):
    out = np.empty(mean.shape[0])
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().plda_ref_transform(_p(mean), _p(tr), _p(psi), mean.shape[0], _p(x), int(n), _p(out))
    return out


def score_grid(psi, enrol, counts, test, threads=0):
    enrol = np.ascontiguousarray(enrol, dtype=np.float64)
    test = np.ascontiguousarray(test, dtype=np.float64)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    out = np.empty((enrol.shape[0], test.shape[0]), dtype=np.float32)
    used = lib().plda_ref_score_grid(_p(psi), psi.shape[0], _p(enrol), _p(counts), enrol.shape[0], _p(test),
                                     test.shape[0], _p(out), int(threads))
    return out, used


def max_threads():
    return lib().plda_ref_max_threads()
