"""``compute_gcv_surface`` / ``calculate_gcv_from_H`` of ``plot_gcv_surface.m:58-122`` restated
(TEST INFRASTRUCTURE — PARITY UNPINNED, see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np

from .solvers import EPS, _mldivide_square, _mv


def calculate_gcv_from_H(Hk, tk, lam, problem_size):
    """``plot_gcv_surface.m:105-122``."""
    k = Hk.shape[1]
    yk = _mldivide_square(Hk.T @ Hk + lam * np.eye(k), Hk.T @ tk)  # :108
    residual_norm_sq = np.linalg.norm(tk - Hk @ yk) ** 2  # :110
    s_diag = np.linalg.svd(Hk[:k, :k], compute_uv=False)  # :111
    with np.errstate(divide="ignore", invalid="ignore"):
        trace_val = np.sum(s_diag ** 2 / (s_diag ** 2 + lam))  # :114
        denominator = (problem_size - trace_val) ** 2
        gcv_val = residual_norm_sq / denominator
    if np.isnan(gcv_val) or np.isinf(gcv_val) or denominator < EPS:  # :119
        gcv_val = 1e20
    return float(gcv_val)


def compute_gcv_surface(method_type, A, B, b, n, k_range, lambda_range):
    """``plot_gcv_surface.m:58-102``; ``k_range`` must be ``1..K`` as at the call site (``:16``)."""
    lambda_range = np.asarray(lambda_range, dtype=float)
    K = len(k_range)
    gcv_surface = np.zeros((len(lambda_range), K))
    gcv_path = np.zeros(K)
    if method_type == "ab":
        op = lambda v: _mv(A, _mv(B, v))  # :64
        r0 = np.asarray(b, dtype=float)
        op_size = A.shape[0]
    else:
        op = lambda v: _mv(B, _mv(A, v))  # :68
        r0 = _mv(B, b)
        op_size = A.shape[1]
    Q = np.zeros((op_size, n + 1))
    H = np.zeros((n + 1, n))
    beta = np.linalg.norm(r0)
    Q[:, 0] = r0 / beta
    for k in k_range:
        v = op(Q[:, k - 1])
        for j in range(k):
            H[j, k - 1] = Q[:, j] @ v
            v = v - H[j, k - 1] * Q[:, j]
        H[k, k - 1] = np.linalg.norm(v)
        if H[k, k - 1] < 1e-12:  # :85
            H[k:, k - 1:] = 0
            break
        Q[:, k] = v / H[k, k - 1]
        Hk = H[: k + 1, :k]
        tk = np.zeros(k + 1)
        tk[0] = beta
        vals = np.array([calculate_gcv_from_H(Hk, tk, lam, op_size) for lam in lambda_range])
        gcv_surface[:, k - 1] = vals
        gcv_path[k - 1] = lambda_range[int(np.argmin(vals))]  # :99-100
    return gcv_surface, gcv_path
