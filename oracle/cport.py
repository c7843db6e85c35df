"""ctypes access to oracle/c/hg_oracle.c (TEST INFRASTRUCTURE): the OpenMP restatement of
``hybrid_ba_gmres_rtp.m`` used as the all-cores CPU baseline by bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libhgoracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        # fork/join heavy code: spinning idle threads starve the workers on shared hosts
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        lib = C.CDLL(_PATH)
        lib.hgo_num_threads.restype = C.c_int
        lib.hgo_hybrid_ba_gmres_rtp.restype = C.c_int
        lib.hgo_hybrid_ba_gmres_rtp.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 8 + \
            [C.c_double, C.c_int, C.c_double] + [C.c_void_p] * 3
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().hgo_num_threads())


def hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam):
    """Same outputs as ``oracle.hybrid_ba_gmres_rtp`` for CSR ``A`` (m x n), ``B`` (n x m)."""
    lib = _load()
    A, B = A.tocsr(), B.tocsr()
    m, n = A.shape
    arrs = [np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64), np.ascontiguousarray(B.indptr, dtype=np.int64),
            np.ascontiguousarray(B.indices, dtype=np.int32), np.ascontiguousarray(B.data, dtype=np.float64),
            np.ascontiguousarray(b, dtype=np.float64), np.ascontiguousarray(x_true, dtype=np.float64)]
    x, err, res = np.zeros(n), np.zeros(maxit), np.zeros(maxit)
    k = lib.hgo_hybrid_ba_gmres_rtp(m, n, *[a.ctypes.data for a in arrs], float(tol), int(maxit), float(lam),
                                    x.ctypes.data, err.ctypes.data, res.ctypes.data)
    return x, err[:k], res[:k], k
