"""ctypes access to oracle/c/hg_oracle.c (TEST INFRASTRUCTURE): the OpenMP restatement of
``hybrid_ba_gmres_rtp.m`` / ``hybrid_ab_gmres_rtp.m`` used as the all-cores CPU baseline by bench.py and
as the checker at BASELINE's full problem sizes."""
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
        lib.hgo_set_num_threads.argtypes = [C.c_int]
        lib.hgo_set_num_threads.restype = None
        lib.hgo_hybrid_rtp.restype = C.c_int
        lib.hgo_hybrid_rtp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 8 + \
            [C.c_double, C.c_int, C.c_double] + [C.c_void_p] * 8
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().hgo_num_threads())


def use_all_cores() -> int:
    """All cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers).  Returns the thread count now in use."""
    lib = _load()
    lib.hgo_set_num_threads(len(os.sched_getaffinity(0)))
    return num_threads()


def _csr_arrays(A, B, b, x_true):
    A, B = A.tocsr(), B.tocsr()
    return [np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64), np.ascontiguousarray(B.indptr, dtype=np.int64),
            np.ascontiguousarray(B.indices, dtype=np.int32), np.ascontiguousarray(B.data, dtype=np.float64),
            np.ascontiguousarray(b, dtype=np.float64), np.ascontiguousarray(x_true, dtype=np.float64)]


def hybrid_rtp(kind, A, B, b, x_true, tol, maxit, lam, orth="mgs", extras=None, want_X=False, solve=True):
    """``kind`` 'ab' / 'ba'.  Same outputs as ``oracle.hybrid_{ab,ba}_gmres_rtp`` for CSR ``A`` (m x n),
    ``B`` (n x m); ``extras`` receives H, beta, t_iter (and X with ``want_X``).  ``solve=False`` runs the
    Arnoldi process only (returns (None, None, None, ksteps))."""
    lib = _load()
    m, n = A.shape
    arrs = _csr_arrays(A, B, b, x_true)
    maxit = int(maxit)
    x, err, res = np.zeros(n), np.zeros(maxit), np.zeros(maxit)
    H = np.zeros((maxit + 1, maxit), order="F")
    beta = np.zeros(1)
    t_iter = np.zeros(maxit)
    X = np.zeros((n, maxit), order="F") if want_X else None
    xv = C.c_int(0)
    k = lib.hgo_hybrid_rtp({"ab": 0, "ba": 1}[kind], {"mgs": 0, "cgs2": 1}[orth], 1 if solve else 0, m, n,
                           *[a.ctypes.data for a in arrs], float(tol), maxit, float(lam), x.ctypes.data,
                           err.ctypes.data, res.ctypes.data, H.ctypes.data, beta.ctypes.data,
                           X.ctypes.data if want_X else None, C.addressof(xv), t_iter.ctypes.data)
    if extras is not None:
        extras.update(H=H, beta=float(beta[0]), t_iter=t_iter[:k])
        if want_X:
            extras["X"] = X[:, :k]
    if not solve:
        return None, None, None, k
    return (x if xv.value else None), err[:k], res[:k], k


def hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam, orth="mgs", extras=None, want_X=False):
    return hybrid_rtp("ba", A, B, b, x_true, tol, maxit, lam, orth, extras, want_X)


def hybrid_ab_gmres_rtp(A, B, b, x_true, tol, maxit, lam, orth="mgs", extras=None, want_X=False):
    return hybrid_rtp("ab", A, B, b, x_true, tol, maxit, lam, orth, extras, want_X)
