"""Oracle test-problem generators (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates ``generate_test_problem.m:1-12``, which dispatches to ``shaw(n)``,
``heat(n)`` and ``deriv2(n)`` of P. C. Hansen's *Regularization Tools*.  That
package is NOT vendored in the reference and no version is pinned (there is no
manifest of any kind), so the published closed-form discretisations are
restated here.  PARITY UNPINNED for these inputs: the solvers
are compared on identical in-memory matrices, never across generators.
"""
from __future__ import annotations

import numpy as np


def deriv2(n: int):
    """Second-derivative Fredholm problem, Hansen's ``deriv2(n)`` example 1.

    Kernel K(s,t) = s(t-1) for s<t, t(s-1) otherwise; Galerkin with box
    functions on [0,1].  Call site: ``generate_test_problem.m:8``.
    """
    h = 1.0 / n
    h2 = h * h
    h32 = h * np.sqrt(h)
    A = np.zeros((n, n))
    for i in range(1, n + 1):
        A[i - 1, i - 1] = h2 * ((i * i - i + 0.25) * h - (i - 2.0 / 3.0))
        for j in range(1, i):
            A[i - 1, j - 1] = h2 * (j - 0.5) * ((i - 0.5) * h - 1.0)
    A = A + np.tril(A, -1).T
    i = np.arange(1, n + 1, dtype=float)
    b = h32 * (i - 0.5) * ((i * i + (i - 1.0) ** 2) * h2 / 2.0 - 1.0) / 6.0
    x = h32 * (i - 0.5)
    return A, b, x


def shaw(n: int):
    """One-dimensional image restoration model, Hansen's ``shaw(n)``.

    Call site: ``generate_test_problem.m:4``.  ``n`` must be even.
    """
    if n % 2:
        raise ValueError("The order n must be even")
    h = np.pi / n
    theta = -np.pi / 2 + (np.arange(n) + 0.5) * h
    co = np.cos(theta)
    psi = np.pi * np.sin(theta)
    ss = psi[:, None] + psi[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        sinc = np.where(ss == 0.0, 1.0, np.sin(ss) / ss)
    # the anti-diagonal has psi_i + psi_j == 0 analytically; force the limit there
    idx = np.arange(n)
    sinc[idx, n - 1 - idx] = 1.0
    A = h * ((co[:, None] + co[None, :]) * sinc) ** 2
    A = 0.5 * (A + A.T)
    x = 2.0 * np.exp(-6.0 * (theta - 0.8) ** 2) + np.exp(-2.0 * (theta + 0.5) ** 2)
    b = A @ x
    return A, b, x


def heat(n: int, kappa: float = 1.0):
    """Inverse heat equation, Hansen's ``heat(n)`` with kappa=1.

    Call site: ``generate_test_problem.m:6``.  ``n`` must be even.
    """
    if n % 2:
        raise ValueError("The order n must be even")
    h = 1.0 / n
    t = h / 2 + h * np.arange(n)
    c = h / (2.0 * kappa * np.sqrt(np.pi))
    d = 1.0 / (4.0 * kappa ** 2)
    k = c * t ** (-1.5) * np.exp(-d / t)
    A = np.zeros((n, n))
    for j in range(n):
        A[j:, j] = k[: n - j]
    x = np.zeros(n)
    for i in range(1, n // 2 + 1):
        ti = i * 20.0 / n
        if ti < 2:
            x[i - 1] = 0.75 * ti * ti / 4.0
        elif ti < 3:
            x[i - 1] = 0.75 + (ti - 2.0) * (3.0 - ti)
        else:
            x[i - 1] = 0.75 * np.exp(-(ti - 3.0) * 2.0)
    b = A @ x
    return A, b, x


def generate_test_problem(name: str, n: int):
    """``[A, b_exact, x_true] = generate_test_problem(name, n)``
    (``generate_test_problem.m:1-12``)."""
    key = name.lower()
    if key == "shaw":
        return shaw(n)
    if key == "heat":
        return heat(n)
    if key == "deriv2":
        return deriv2(n)
    raise ValueError("Unknown problem name. Use shaw, heat, or deriv2.")


def add_noise(b_exact: np.ndarray, level: float, seed: int = 0) -> np.ndarray:
    """``b + level*norm(b)*noise/norm(noise)`` (``run_equivalence_plots.m:6-8``,
    ``run_2D_phantom.m:17-20``).  MATLAB's ``rng(0); randn`` stream is not
    reproducible without MATLAB; ``numpy.random.default_rng(seed)`` is used."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal(b_exact.shape)
    return b_exact + level * np.linalg.norm(b_exact) * noise / np.linalg.norm(noise)
