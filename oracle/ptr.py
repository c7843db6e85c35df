"""Solve path of the reference's project-then-regularise (PTR) solvers (TEST INFRASTRUCTURE —
PARITY UNPINNED, see oracle/__init__.py).

Restates the Arnoldi + projected solve + iterate + histories of
``ABgmres_hybrid_bounds.m:11-41,83-88``, ``BAgmres_hybrid_bounds.m:11-40,79-84``,
``ABgmres_nonhybrid_bounds.m:12-40,78-80`` and ``BAgmres_nonhybrid_bounds.m:12-39,79-81``.
The filter-factor bound algebra of those files (dense ``eig`` of ``A*B`` / ``B*A``,
``phi``/``dphi`` outputs, the ``DeltaM`` argument) is out of scope
and not restated: only the first four outputs are returned.
"""
from __future__ import annotations

import numpy as np

from .solvers import _mldivide_rect, _mldivide_square, _mv


def _ptr(A, B, b, x_true, tol, maxit, lam, space, hybrid, premultiplied=False, extras=None):
    maxit = int(maxit)
    if space == "m":  # AB: Krylov space of A*B in R^m, x = B*z
        r0 = b - _mv(A, _mv(B, np.zeros(B.shape[1])))  # ABgmres_hybrid_bounds.m:11-12
        op = lambda q: _mv(A, _mv(B, q))  # :25
    else:  # BA: Krylov space of B*A in R^n
        r0 = _mv(B, b - _mv(A, np.zeros(A.shape[1])))  # BAgmres_hybrid_bounds.m:12-13
        if premultiplied:
            M = B @ A  # BAgmres_nonhybrid_bounds.m:4,25 applies the pre-multiplied product
            op = lambda q: _mv(M, q)
        else:
            op = lambda q: _mv(B, _mv(A, q))  # BAgmres_hybrid_bounds.m:25
    beta = np.linalg.norm(r0)
    dim = r0.shape[0]
    Q = np.zeros((dim, maxit + 1))
    H = np.zeros((maxit + 1, maxit))
    Q[:, 0] = r0 / beta
    residual_norm = np.zeros(maxit)
    error_norm = np.zeros(maxit)
    xk = None
    X = np.zeros((A.shape[1], maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        v = op(Q[:, k - 1])
        for j in range(k):  # MGS, e.g. ABgmres_hybrid_bounds.m:26-29
            H[j, k - 1] = Q[:, j] @ v
            v = v - H[j, k - 1] * Q[:, j]
        H[k, k - 1] = np.linalg.norm(v)
        if H[k, k - 1] == 0:  # :31
            break
        Q[:, k] = v / H[k, k - 1]
        Hk = H[: k + 1, :k]
        tk = np.zeros(k + 1)
        tk[0] = beta
        if hybrid:
            yk = _mldivide_square(Hk.T @ Hk + lam * np.eye(k), Hk.T @ tk)  # :36
        else:
            yk = _mldivide_rect(Hk, tk)  # ABgmres_nonhybrid_bounds.m:35
        if space == "m":
            zk = Q[:, :k] @ yk  # :37
            xk = _mv(B, zk)  # :38
        else:
            xk = Q[:, :k] @ yk  # BAgmres_hybrid_bounds.m:37
        residual_norm[k - 1] = np.linalg.norm(b - _mv(A, xk)) / np.linalg.norm(b)  # :40
        error_norm[k - 1] = np.linalg.norm(xk - x_true) / np.linalg.norm(x_true)  # :41
        if X is not None:
            X[:, k - 1] = xk
        if residual_norm[k - 1] <= tol:  # :83
            break
    niters = k
    if extras is not None:
        extras.update(H=H, beta=beta, X=X[:, :niters])
    return xk, error_norm[:niters], residual_norm[:niters], niters


def ABgmres_hybrid_bounds(A, B, b, x_true, tol, maxit, lam, DeltaM=None, extras=None):
    """First four outputs of ``ABgmres_hybrid_bounds.m`` (``DeltaM`` only feeds the bounds)."""
    return _ptr(A, B, b, x_true, tol, maxit, lam, "m", True, extras=extras)


def BAgmres_hybrid_bounds(A, B, b, x_true, tol, maxit, lam, DeltaM=None, extras=None):
    """First four outputs of ``BAgmres_hybrid_bounds.m``."""
    return _ptr(A, B, b, x_true, tol, maxit, lam, "n", True, extras=extras)


def ABgmres_nonhybrid_bounds(A, B, b, x_true, tol, maxit, DeltaM=None, extras=None):
    """First four outputs of ``ABgmres_nonhybrid_bounds.m``."""
    return _ptr(A, B, b, x_true, tol, maxit, 0.0, "m", False, extras=extras)


def BAgmres_nonhybrid_bounds(A, B, b, x_true, tol, maxit, DeltaM=None, extras=None):
    """First four outputs of ``BAgmres_nonhybrid_bounds.m`` (pre-multiplied ``M = B*A``, ``:4,25``)."""
    return _ptr(A, B, b, x_true, tol, maxit, 0.0, "n", False, premultiplied=True, extras=extras)


def hybrid_gmres_gcv(kind, A, B, b, x_true, tol, maxit, lambda_range, extras=None):
    """Hybrid PTR solve with lambda chosen at every iteration — a composition of two
    pieces of reference text: ``lambda_k`` = the grid minimiser of ``calculate_gcv_from_H(Hk, tk, lambda,
    op_size)`` (``plot_gcv_surface.m:92-100,104-122``), then ``yk = (Hk'*Hk + lambda_k*eye(k)) \\ (Hk'*tk)``
    and ``xk`` as at ``ABgmres_hybrid_bounds.m:36-38`` / ``BAgmres_hybrid_bounds.m:36-37``.  MGS Arnoldi,
    ``== 0`` breakdown and stop rule of the ``*_bounds`` files.  Returns (x, err, res, niters, lambda_path)."""
    from .gcv_surface import calculate_gcv_from_H
    maxit = int(maxit)
    lambda_range = np.asarray(lambda_range, dtype=float).ravel()
    if kind == "ab":
        r0, op, op_size = b.copy(), (lambda q: _mv(A, _mv(B, q))), A.shape[0]
    else:
        r0, op, op_size = _mv(B, b), (lambda q: _mv(B, _mv(A, q))), A.shape[1]
    beta = np.linalg.norm(r0)
    Q = np.zeros((r0.shape[0], maxit + 1))
    H = np.zeros((maxit + 1, maxit))
    Q[:, 0] = r0 / beta
    residual_norm, error_norm, path = np.zeros(maxit), np.zeros(maxit), np.zeros(maxit)
    X = np.zeros((A.shape[1], maxit))
    xk, k = None, 0
    for k in range(1, maxit + 1):
        v = op(Q[:, k - 1])
        for j in range(k):
            H[j, k - 1] = Q[:, j] @ v
            v = v - H[j, k - 1] * Q[:, j]
        H[k, k - 1] = np.linalg.norm(v)
        if H[k, k - 1] == 0:
            break
        Q[:, k] = v / H[k, k - 1]
        Hk = H[: k + 1, :k]
        tk = np.zeros(k + 1)
        tk[0] = beta
        vals = np.array([calculate_gcv_from_H(Hk, tk, lam, op_size) for lam in lambda_range])
        lam_k = lambda_range[int(np.argmin(vals))]  # first minimum, as MATLAB's min
        path[k - 1] = lam_k
        yk = _mldivide_square(Hk.T @ Hk + lam_k * np.eye(k), Hk.T @ tk)
        xk = _mv(B, Q[:, :k] @ yk) if kind == "ab" else Q[:, :k] @ yk
        X[:, k - 1] = xk
        residual_norm[k - 1] = np.linalg.norm(b - _mv(A, xk)) / np.linalg.norm(b)
        error_norm[k - 1] = np.linalg.norm(xk - x_true) / np.linalg.norm(x_true)
        if residual_norm[k - 1] <= tol:
            break
    if extras is not None:
        extras.update(H=H, beta=beta, X=X[:, :k])
    return xk, error_norm[:k], residual_norm[:k], k, path[:k]
