"""Line-literal NumPy/SciPy restatement of the reference's ★ solvers.

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (see ``oracle/__init__.py``).

Every function follows one reference file statement by statement (file:line in
each docstring) and keeps its quirks (SURVEY.md §8c "quirks").  ``A`` and ``B``
may be ``scipy.sparse`` matrices or dense ``ndarray`` — both are passed by the
reference's own callers.  Two instrumentation additions that do not change any
returned reference value:

* ``orth``: ``'mgs'`` (the reference, e.g. ``hybrid_ab_gmres_rtp.m:20-23``) or
  ``'cgs2'`` (the north-star device algorithm) — so GPU CGS2 results can be
  compared both with the literal reference arithmetic and with the same
  algorithm in a different summation order;
* ``extras``: pass a dict to receive ``H``, ``beta``, ``Q`` and per-iteration
  iterates ``X`` (columns) for coefficient/iterate parity checks.

MATLAB built-ins are mapped as: ``norm`` → ``numpy.linalg.norm``; square
symmetric ``\\`` → Cholesky (``scipy.linalg.solve(assume_a='pos')``) with LU
fall-back, as MATLAB's ``mldivide`` does; rectangular ``\\`` → QR with column
pivoting (LAPACK ``gelsy``); ``svd`` → ``numpy.linalg.svd``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

EPS = np.finfo(float).eps


def _mv(M, v):
    """MATLAB ``M*v`` for sparse or full ``M``."""
    return np.asarray(M @ v).ravel()


def _rmv(M, u):
    """MATLAB ``M'*u`` (``M.'*u`` for real data)."""
    return np.asarray(M.T @ u).ravel()


def _mldivide_square(M, rhs):
    """MATLAB ``M\\rhs`` for a square, (numerically) symmetric matrix."""
    try:
        return sla.solve(M, rhs, assume_a="pos", check_finite=False)
    except (sla.LinAlgError, ValueError):
        try:
            return sla.solve(M, rhs, check_finite=False)
        except sla.LinAlgError:
            return np.full_like(rhs, np.nan)


def _mldivide_rect(M, rhs):
    """MATLAB ``M\\rhs`` for rectangular ``M`` (QR with column pivoting)."""
    y, *_ = sla.lstsq(M, rhs, lapack_driver="gelsy", check_finite=False)
    return y


def _orthogonalise(Q, v, k, orth):
    """Return (h[0:k], v_orth).  ``k`` = number of basis vectors (1-based k)."""
    h = np.zeros(k)
    if orth == "mgs":
        # hybrid_ab_gmres_rtp.m:20-23
        for j in range(k):
            h[j] = Q[:, j] @ v
            v = v - h[j] * Q[:, j]
    elif orth == "cgs2":
        Qk = Q[:, :k]
        h1 = Qk.T @ v
        v = v - Qk @ h1
        h2 = Qk.T @ v
        v = v - Qk @ h2
        h = h1 + h2
    else:
        raise ValueError("orth must be 'mgs' or 'cgs2'")
    return h, v


def arnoldi(op, r0, kmax, orth="mgs", breakdown=lambda h: h == 0.0):
    """Plain Arnoldi loop shared by the restatements below
    (``hybrid_ab_gmres_rtp.m:9-26`` minus the projected solve).
    Returns (Q, H, beta, ksteps) where ksteps counts completed columns."""
    beta = np.linalg.norm(r0)
    n = r0.shape[0]
    Q = np.zeros((n, kmax + 1))
    H = np.zeros((kmax + 1, kmax))
    Q[:, 0] = r0 / beta
    kdone = 0
    for k in range(1, kmax + 1):
        v = op(Q[:, k - 1])
        h, v = _orthogonalise(Q, v, k, orth)
        H[:k, k - 1] = h
        H[k, k - 1] = np.linalg.norm(v)
        if breakdown(H[k, k - 1]):
            break
        Q[:, k] = v / H[k, k - 1]
        kdone = k
    return Q, H, beta, kdone


def hybrid_ab_gmres_rtp(A, B, b, x_true, tol, maxit, lam, orth="mgs", extras=None):
    """``hybrid_ab_gmres_rtp.m:1-45``.

    Quirks kept: Krylov space is the n-space ``BA + lambda I`` started from
    ``B b`` (``:6-13``); breakdown ``==0`` leaves the loop before x/histories
    (``:25``); ``AQk = A*Qk`` recomputed each iteration (``:31``); true residual
    (``:35``); stop ``<=`` (``:38``).
    """
    n = A.shape[1]
    maxit = int(maxit)
    x0 = np.zeros(n)
    M_reg_op = lambda v: _mv(B, _mv(A, v)) + lam * v  # :6
    d_krylov = _mv(B, b)  # :7
    r0 = d_krylov - M_reg_op(x0)  # :9
    beta = np.linalg.norm(r0)  # :10
    Q = np.zeros((n, maxit + 1))
    H = np.zeros((maxit + 1, maxit))
    Q[:, 0] = r0 / beta
    error_norm = np.zeros(maxit)
    residual_norm = np.zeros(maxit)
    x = None
    X = np.zeros((n, maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        v = M_reg_op(Q[:, k - 1])  # :19
        h, v = _orthogonalise(Q, v, k, orth)  # :20-23
        H[:k, k - 1] = h
        H[k, k - 1] = np.linalg.norm(v)  # :24
        if H[k, k - 1] == 0:  # :25
            break
        Q[:, k] = v / H[k, k - 1]  # :26
        Qk = Q[:, :k]
        AQk = np.asarray(A @ Qk)  # :31
        yk = _mldivide_square(AQk.T @ AQk + lam * np.eye(k), AQk.T @ b)  # :32
        x = Qk @ yk  # :33
        residual_norm[k - 1] = np.linalg.norm(b - _mv(A, x)) / np.linalg.norm(b)  # :35
        error_norm[k - 1] = np.linalg.norm(x - x_true) / np.linalg.norm(x_true)  # :36
        if X is not None:
            X[:, k - 1] = x
        if residual_norm[k - 1] <= tol:  # :38
            break
    niters = k  # :41
    if extras is not None:
        extras.update(H=H, beta=beta, Q=Q, X=X[:, :niters])
    return x, error_norm[:niters], residual_norm[:niters], niters


def hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam, orth="mgs", extras=None):
    """``hybrid_ba_gmres_rtp.m:1-42``.  Same Arnoldi as AB-RTP (``:6-26``);
    projected problem is ``H(1:k+1,1:k) \\ [beta;0]`` (``:28-29``)."""
    n = A.shape[1]
    maxit = int(maxit)
    x = np.zeros(n)  # :4
    M_reg = lambda v: _mv(B, _mv(A, v)) + lam * v  # :6
    d = _mv(B, b)  # :7
    r0 = d - M_reg(x)  # :9
    beta = np.linalg.norm(r0)
    Q = np.zeros((n, maxit + 1))
    H = np.zeros((maxit + 1, maxit))
    Q[:, 0] = r0 / beta
    error_norm = np.zeros(maxit)
    residual_norm = np.zeros(maxit)
    X = np.zeros((n, maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        v = M_reg(Q[:, k - 1])  # :19
        h, v = _orthogonalise(Q, v, k, orth)  # :20-23
        H[:k, k - 1] = h
        H[k, k - 1] = np.linalg.norm(v)  # :24
        if H[k, k - 1] == 0:  # :25
            break
        Q[:, k] = v / H[k, k - 1]  # :26
        Hk = H[: k + 1, :k]  # :28
        rhs = np.zeros(k + 1)
        rhs[0] = beta
        yk = _mldivide_rect(Hk, rhs)  # :29
        x = Q[:, :k] @ yk  # :30
        residual_norm[k - 1] = np.linalg.norm(b - _mv(A, x)) / np.linalg.norm(b)  # :32
        error_norm[k - 1] = np.linalg.norm(x - x_true) / np.linalg.norm(x_true)  # :33
        if X is not None:
            X[:, k - 1] = x
        if residual_norm[k - 1] <= tol:  # :35
            break
    niters = k  # :38
    if extras is not None:
        extras.update(H=H, beta=beta, Q=Q, X=X[:, :niters])
    return x, error_norm[:niters], residual_norm[:niters], niters


def gcv_arnoldi(A, B, b, m, k_gcv, gcv_type, orth="mgs"):
    """The lambda-independent half of ``gcv_function.m`` (``:4-32``): the
    unshifted Arnoldi in m-space ('ab', ``r0=b``) or n-space ('ba', ``r0=B*b``)
    with the ``<1e-12`` breakdown test (``:30``).  Returns (H, beta)."""
    k_gcv = int(k_gcv)
    if gcv_type == "ab":
        r0 = np.asarray(b, dtype=float).copy()  # :5
        op = lambda q: _mv(A, _mv(B, q))  # :20
    else:
        r0 = _mv(B, b)  # :8
        op = lambda q: _mv(B, _mv(A, q))  # :22
    Q, H, beta, _ = arnoldi(op, r0, k_gcv, orth=orth, breakdown=lambda h: h < 1e-12)
    return H, beta


def gcv_from_H(lam, H, beta, trace_m):
    """The projected half of ``gcv_function.m`` (``:33-58``)."""
    k = H.shape[1]  # :33 (always k_gcv, trailing zero columns kept)
    Hk = H[: k + 1, :k]  # :35
    tk = np.zeros(k + 1)
    tk[0] = beta  # :16,36
    yk = _mldivide_square(Hk.T @ Hk + lam * np.eye(k), Hk.T @ tk)  # :38
    residual_norm_sq = np.linalg.norm(tk - Hk @ yk) ** 2  # :40
    s_diag = np.linalg.svd(H[:k, :k], compute_uv=False)  # :42-43
    with np.errstate(divide="ignore", invalid="ignore"):
        trace_val = np.sum(s_diag ** 2 / (s_diag ** 2 + lam))  # :51
        denominator = (trace_m - trace_val) ** 2  # :52
        gcv_val = residual_norm_sq / denominator  # :54
    if np.isnan(gcv_val) or np.isinf(gcv_val) or denominator < EPS:  # :56
        gcv_val = 1e20
    return float(gcv_val)


def gcv_function(lam, A, B, b, m, k_gcv, gcv_type, orth="mgs"):
    """``gcv_val = gcv_function(lambda,A,B,b,m,k_gcv,gcv_type)``
    (``gcv_function.m:1-59``), Arnoldi re-run on every call as the reference
    does."""
    H, beta = gcv_arnoldi(A, B, b, m, k_gcv, gcv_type, orth=orth)
    trace_m = m if gcv_type == "ab" else A.shape[1]  # :46-50
    return gcv_from_H(lam, H, beta, trace_m)


def hybrid_lsqr_solver(A, b, x_true, tol, maxit, lam, extras=None):
    """``hybrid_lsqr_solver.m:1-52`` — LSQR on the explicitly stacked system
    ``[A; sqrt(lambda) I]`` (``:5-6``); true residual (``:43``); strict ``<``
    stop (``:45``)."""
    m, n = A.shape
    maxit = int(maxit)
    if sp.issparse(A):
        A_aug = sp.vstack([A, np.sqrt(lam) * sp.eye(n)]).tocsr()  # :5
    else:
        A_aug = np.vstack([A, np.sqrt(lam) * np.eye(n)])
    b_aug = np.concatenate([b, np.zeros(n)])  # :6
    x = np.zeros(n)
    beta_aug = np.linalg.norm(b_aug)
    u_aug = b_aug / beta_aug
    v_hat = _rmv(A_aug, u_aug)  # :11
    alpha_aug = np.linalg.norm(v_hat)
    v = v_hat / alpha_aug
    w = v.copy()
    phi_bar = beta_aug
    rho_bar = alpha_aug
    error_norm = np.zeros(maxit)
    residual_norm = np.zeros(maxit)
    alphas, betas = [alpha_aug], [beta_aug]
    X = np.zeros((n, maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        u_hat = _mv(A_aug, v) - alpha_aug * u_aug  # :22
        beta_aug = np.linalg.norm(u_hat)
        u_aug = u_hat / beta_aug
        v_hat = _rmv(A_aug, u_aug) - beta_aug * v  # :26
        alpha_aug = np.linalg.norm(v_hat)
        v = v_hat / alpha_aug
        rho = np.sqrt(rho_bar ** 2 + beta_aug ** 2)  # :30
        c = rho_bar / rho
        s = beta_aug / rho
        theta = s * alpha_aug
        rho_bar = -c * alpha_aug
        phi = c * phi_bar
        phi_bar = s * phi_bar
        x = x + (phi / rho) * w  # :39
        w = v - (theta / rho) * w  # :40
        error_norm[k - 1] = np.linalg.norm(x - x_true) / np.linalg.norm(x_true)  # :42
        residual_norm[k - 1] = np.linalg.norm(b - _mv(A, x)) / np.linalg.norm(b)  # :43
        alphas.append(alpha_aug)
        betas.append(beta_aug)
        if X is not None:
            X[:, k - 1] = x
        if residual_norm[k - 1] < tol:  # :45
            break
    niters = k
    if extras is not None:
        extras.update(alpha=np.array(alphas), beta=np.array(betas), X=X[:, :niters])
    return x, error_norm[:niters], residual_norm[:niters], niters


def hybrid_lsmr_solver(A, b, x_true, tol, maxit, lam, extras=None):
    """``hybrid_lsmr_solver.m:1-57``.  Quirks kept: no v-update at
    ``k==maxit`` so ``alpha_k1`` is alpha_k there (``:29-38``); ``e1 e1'`` term
    via ``eye(k,1)*eye(1,k)`` (``:41``); stop ``<=`` (``:50``)."""
    n = A.shape[1]
    maxit = int(maxit)
    x = np.zeros(n)
    u = np.asarray(b, dtype=float).copy()
    beta1 = np.linalg.norm(u)
    u = u / beta1
    V = np.zeros((n, maxit))
    B_k = np.zeros((maxit + 1, maxit))
    v_hat = _rmv(A, u)  # :13
    alpha1 = np.linalg.norm(v_hat)
    v = v_hat / alpha1
    V[:, 0] = v
    error_norm = np.zeros(maxit)
    residual_norm = np.zeros(maxit)
    X = np.zeros((n, maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        B_k[k - 1, k - 1] = alpha1  # :23
        u_hat = _mv(A, v) - alpha1 * u  # :24
        beta_k = np.linalg.norm(u_hat)
        u = u_hat / beta_k
        B_k[k, k - 1] = beta_k  # :27
        if k < maxit:  # :29
            v_hat = _rmv(A, u) - beta_k * v
            alpha_k_plus_1 = np.linalg.norm(v_hat)
            v = v_hat / alpha_k_plus_1
            V[:, k] = v
            alpha1 = alpha_k_plus_1
        Bk = B_k[: k + 1, :k]  # :37
        alpha_k1 = alpha1
        beta_k1 = beta_k
        T = Bk.T @ Bk
        e1 = np.zeros((k, 1))
        e1[0, 0] = 1.0
        LHS = T @ T + (alpha_k1 * beta_k1) ** 2 * (e1 @ e1.T) + lam * np.eye(k)  # :41
        RHS = B_k[0, 0] * beta1 * (T @ e1).ravel()  # :42
        yk = _mldivide_square(LHS, RHS)  # :44
        x = V[:, :k] @ yk  # :45
        error_norm[k - 1] = np.linalg.norm(x - x_true) / np.linalg.norm(x_true)  # :47
        residual_norm[k - 1] = np.linalg.norm(b - _mv(A, x)) / np.linalg.norm(b)  # :48
        if X is not None:
            X[:, k - 1] = x
        if residual_norm[k - 1] <= tol:  # :50
            break
    niters = k
    if extras is not None:
        extras.update(B_k=B_k, beta1=beta1, V=V, X=X[:, :niters])
    return x, error_norm[:niters], residual_norm[:niters], niters


def lsqr_solver(A, b, x_true, tol, maxit, extras=None):
    """``lsqr_solver.m:1-54``.  Quirk kept: residual history is the recurrence
    estimate ``|phi_bar|/norm(b)`` (``:44``) and only the LAST entry is replaced
    by the true residual (``:52``); stop ``<=`` (``:46``)."""
    n = A.shape[1]
    maxit = int(maxit)
    x = np.zeros(n)
    beta = np.linalg.norm(b)
    u = b / beta
    v_hat = _rmv(A, u)  # :10
    alpha = np.linalg.norm(v_hat)
    v = v_hat / alpha
    w = v.copy()
    phi_bar = beta
    rho_bar = alpha
    error_norm = np.zeros(maxit)
    residual_norm = np.zeros(maxit)
    X = np.zeros((n, maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        u_hat = _mv(A, v) - alpha * u  # :22
        beta = np.linalg.norm(u_hat)
        u = u_hat / beta
        v_hat = _rmv(A, u) - beta * v  # :26
        alpha = np.linalg.norm(v_hat)
        v = v_hat / alpha
        rho = np.sqrt(rho_bar ** 2 + beta ** 2)  # :31
        c = rho_bar / rho
        s = beta / rho
        theta = s * alpha
        rho_bar = -c * alpha
        phi = c * phi_bar
        phi_bar = s * phi_bar
        x = x + (phi / rho) * w  # :40
        w = v - (theta / rho) * w  # :41
        error_norm[k - 1] = np.linalg.norm(x - x_true) / np.linalg.norm(x_true)  # :43
        residual_norm[k - 1] = abs(phi_bar) / np.linalg.norm(b)  # :44
        if X is not None:
            X[:, k - 1] = x
        if residual_norm[k - 1] <= tol:  # :46
            break
    niters = k
    error_norm = error_norm[:niters]
    residual_norm = residual_norm[:niters].copy()
    residual_norm[-1] = np.linalg.norm(b - _mv(A, x)) / np.linalg.norm(b)  # :52
    if extras is not None:
        extras.update(X=X[:, :niters])
    return x, error_norm, residual_norm, niters


def _fro(A):
    if sp.issparse(A):
        return float(np.sqrt((A.data ** 2).sum()))
    return float(np.linalg.norm(A, "fro"))


def lsmr_solver(A, b, x_true=None, tol=None, maxit=None, extras=None):
    """``lsmr_solver.m:1-83`` — 5 outputs ``(x, err_hist, res_hist, ar_hist,
    iters)``.  Quirks kept: defaults (``:3-5``), ``>0`` guards (``:12,16,36,40``),
    ``eps`` guards (``:70-71``), NaN ``err_hist`` without ``x_true`` (``:28,72``),
    strict ``<`` stop (``:76``)."""
    if tol is None:
        tol = 1e-6  # :3
    m, n = A.shape
    if maxit is None:
        maxit = min(m, n)  # :5
    maxit = int(maxit)
    x = np.zeros(n)
    u = np.asarray(b, dtype=float).copy()
    beta = np.linalg.norm(u)
    if beta > 0:
        u = u / beta
    v = _rmv(A, u)  # :14
    alpha = np.linalg.norm(v)
    if alpha > 0:
        v = v / alpha
    zetabar = alpha * beta
    alphabar = alpha
    rho = 1.0
    rhobar = 1.0
    cbar = 1.0
    sbar = 0.0
    h = v.copy()
    hbar = np.zeros(n)
    err_hist = np.full(maxit, np.nan)
    res_hist = np.zeros(maxit)
    ar_hist = np.zeros(maxit)
    have_true = x_true is not None and np.size(x_true) > 0
    X = np.zeros((n, maxit)) if extras is not None else None
    k = 0
    for k in range(1, maxit + 1):
        u = _mv(A, v) - alpha * u  # :34
        beta = np.linalg.norm(u)
        if beta > 0:
            u = u / beta
        v = _rmv(A, u) - beta * v  # :38
        alpha = np.linalg.norm(v)
        if alpha > 0:
            v = v / alpha
        alphahat = alphabar  # :42
        rhoold = rho
        rho = np.hypot(alphahat, beta)
        c = alphahat / rho
        s = beta / rho
        thetanew = s * alpha
        alphabar = c * alpha
        rhobarold = rhobar  # :51
        thetabar = sbar * rho
        rhobar = np.hypot(cbar * rho, thetanew)
        cbar = (cbar * rho) / rhobar
        sbar = thetanew / rhobar
        zeta = cbar * zetabar  # :58
        zetabar = -sbar * zetabar
        if k == 1:  # :61
            hbar = h.copy()
        else:
            hbar = h - (thetabar * rho) / (rhoold * rhobarold) * hbar
        x = x + (zeta / (rho * rhobar)) * hbar  # :66
        h = v - (thetanew / rho) * h  # :67
        r = b - _mv(A, x)  # :69
        res_hist[k - 1] = np.linalg.norm(r) / (np.linalg.norm(b) + EPS)  # :70
        ar_hist[k - 1] = np.linalg.norm(_rmv(A, r)) / (_fro(A) * max(np.linalg.norm(r), EPS))  # :71
        if have_true:
            err_hist[k - 1] = np.linalg.norm(x - x_true) / np.linalg.norm(x_true)  # :73
        if X is not None:
            X[:, k - 1] = x
        if res_hist[k - 1] < tol:  # :76
            break
    iters = k
    if extras is not None:
        extras.update(X=X[:, :iters])
    return x, err_hist[:iters], res_hist[:iters], ar_hist[:iters], iters
