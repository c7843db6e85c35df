"""CPU tests pinning the oracle: closed-form properties of the restated
generators, the relations asserted by the reference's figure
titles (run_equivalence_plots.m:33,44), and MATLAB-fminbnd behaviour.  The
reference ships no golden vectors, so this is all the pinning that exists
("parity unpinned", oracle/__init__.py)."""
import numpy as np
import pytest
import scipy.optimize as so
import scipy.sparse as sp

import oracle
from oracle import ct
from oracle.generators import add_noise


def test_deriv2_properties():
    A, b, x = oracle.generate_test_problem("deriv2", 32)
    assert np.allclose(A, A.T)
    assert abs(np.linalg.cond(A) - 1.244e3) / 1.244e3 < 1e-3
    assert np.linalg.norm(A @ x - b) < 1e-15
    # discretised Green's function of -d2/dx2 is negative definite
    assert np.all(np.linalg.eigvalsh(A) < 0)


def test_shaw_properties():
    A, b, x = oracle.generate_test_problem("shaw", 32)
    assert np.array_equal(A, A.T)
    assert abs(np.linalg.norm(x) - 5.647) < 1e-3
    assert np.linalg.cond(A) > 1e15
    assert np.allclose(A @ x, b)
    with pytest.raises(ValueError):
        oracle.shaw(31)


def test_heat_properties():
    A, b, x = oracle.generate_test_problem("heat", 32)
    assert np.allclose(A, np.tril(A))
    for d in range(1, 5):  # Toeplitz
        assert np.allclose(np.diag(A, -d), A[d, 0])
    assert np.all(x[16:] == 0)
    assert np.allclose(A @ x, b)


def test_generate_test_problem_dispatch():
    with pytest.raises(ValueError, match="Unknown problem name"):
        oracle.generate_test_problem("phillips", 32)
    A, _, _ = oracle.generate_test_problem("SHAW", 8)  # lower() as in generate_test_problem.m:2
    assert A.shape == (8, 8)


def _problem(name="deriv2"):
    A, b_exact, x_true = oracle.generate_test_problem(name, 32)
    return A, A.T.copy(), add_noise(b_exact, 1e-2, 0), x_true


def test_title_claim_ba_gmres_equals_lsmr():
    """run_equivalence_plots.m:33 'BA-GMRES vs. LSMR Solution (≡)': BA-RTP with lambda=0 is plain
    BA-GMRES; with B=A' its iterates equal LSMR's (first iterations, before Arnoldi breaks down)."""
    A, B, b, x_true = _problem()
    e1, e2 = {}, {}
    oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, 5, 0.0, extras=e1)
    oracle.lsmr_solver(A, b, x_true, 0.0, 5, extras=e2)
    for k, tol in ((0, 1e-13), (2, 1e-11), (4, 1e-6)):
        d = np.linalg.norm(e1["X"][:, k] - e2["X"][:, k]) / np.linalg.norm(e2["X"][:, k])
        assert d < tol, (k, d)


def test_title_claim_ab_gmres_equals_lsqr():
    """run_equivalence_plots.m:44 'AB-GMRES vs. LSQR Solution (≡)' via AB-RTP with lambda=0."""
    A, B, b, x_true = _problem()
    e1, e2 = {}, {}
    oracle.hybrid_ab_gmres_rtp(A, B, b, x_true, 0.0, 5, 0.0, extras=e1)
    oracle.lsqr_solver(A, b, x_true, 0.0, 5, extras=e2)
    for k, tol in ((0, 1e-13), (2, 1e-11), (4, 1e-6)):
        d = np.linalg.norm(e1["X"][:, k] - e2["X"][:, k]) / np.linalg.norm(e2["X"][:, k])
        assert d < tol, (k, d)


def test_ab_rtp_equals_hybrid_lsqr_for_matched_B():
    """Same Krylov space and same Tikhonov projection."""
    A, B, b, x_true = _problem()
    e1, e2 = {}, {}
    oracle.hybrid_ab_gmres_rtp(A, B, b, x_true, 0.0, 5, 1e-3, extras=e1)
    oracle.hybrid_lsqr_solver(A, b, x_true, 0.0, 5, 1e-3, extras=e2)
    for k, tol in ((0, 1e-13), (2, 1e-10), (4, 1e-6)):
        d = np.linalg.norm(e1["X"][:, k] - e2["X"][:, k]) / np.linalg.norm(e2["X"][:, k])
        assert d < tol, (k, d)


def test_mgs_and_cgs2_agree_on_ct():
    """North star mandates CGS2, the reference is MGS: iterates agree far below 1e-8."""
    A, B, b, x_true = ct.make_ct_problem(24, 36, "parallel", "perturbed", mismatch=1e-2)
    e1, e2 = {}, {}
    r1 = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 40, 1e-2, orth="mgs", extras=e1)
    r2 = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 40, 1e-2, orth="cgs2", extras=e2)
    assert r1[3] == r2[3]
    assert np.max(np.abs(r1[2] - r2[2]) / r1[2]) < 1e-10
    assert np.linalg.norm(e1["X"] - e2["X"]) / np.linalg.norm(e1["X"]) < 1e-10


def test_solver_quirks():
    A, B, b, x_true = _problem()
    # lsqr_solver.m:44,52 — only the last residual entry is the true residual
    x, err, res, it = oracle.lsqr_solver(A, b, x_true, 1e-6, 8)
    assert abs(res[-1] - np.linalg.norm(b - A @ x) / np.linalg.norm(b)) < 1e-15
    # lsmr_solver.m: five outputs, NaN error history without x_true, default maxit
    out = oracle.lsmr_solver(A, b)
    assert len(out) == 5 and np.all(np.isnan(out[1])) and out[4] <= 32
    # histories are truncated to niters (hybrid_ab_gmres_rtp.m:41-43)
    x, err, res, it = oracle.hybrid_ab_gmres_rtp(A, B, b, x_true, 0.5, 32, 1e-3)
    assert len(err) == len(res) == it and res[-1] <= 0.5
    # breakdown at k=1 leaves AB's x unassigned (:25): zero rhs direction
    Z = np.zeros((4, 4))
    Z[0, 0] = 1.0
    bb = np.array([1.0, 0, 0, 0])
    x, err, res, it = oracle.hybrid_ab_gmres_rtp(Z, Z.T, bb, bb, 1e-6, 3, 0.0)
    assert it == 1 and x is None and res[0] == 0.0


def test_gcv_function_sentinel_and_types():
    A, B, b, x_true = _problem("shaw")
    for t in ("ab", "ba"):
        v = oracle.gcv_function(1e-4, A, B, b, 32, 20, t)
        assert np.isfinite(v) and v > 0
    # denominator < eps -> 1e20 (gcv_function.m:56-58): trace_m equal to the trace term
    H = np.zeros((3, 2))
    H[0, 0] = H[1, 1] = 1.0
    from oracle.solvers import gcv_from_H
    assert gcv_from_H(0.0, H, 1.0, 2.0) == 1e20


def test_fminbnd_matches_scipy_and_brackets():
    f = lambda x: (x - 0.3) ** 2 + 0.1 * np.sin(20 * x)
    xf, fv, flag, cnt = oracle.fminbnd(f, 0.0, 1.0, 1e-8)
    xs, fs, ierr, ns = so.fminbound(f, 0.0, 1.0, xtol=1e-8, full_output=True)
    assert abs(xf - xs) < 1e-7 and cnt == ns and flag == 1
    tr = []
    oracle.fminbnd(f, 0.0, 1.0, 1e-4, trace=tr)
    assert abs(tr[0] - (3 - np.sqrt(5)) / 2) < 1e-15  # first golden-section point


def test_ct_generator_properties():
    N = 24
    A = ct.projector(N, np.arange(0, 180, 5.0))
    assert A.shape == (36 * int(round(np.sqrt(2) * N)), N * N)
    assert A.data.min() > 0 and A.data.max() <= np.sqrt(2) + 1e-12
    # a view's rays tile the image: summed intersection lengths ~ area
    p = int(round(np.sqrt(2) * N))
    per_view = np.asarray(A.sum(axis=1)).ravel().reshape(36, p).sum(axis=1)
    assert np.all(np.abs(per_view - N * N) / (N * N) < 0.02)
    # axis-aligned view: every ray crosses exactly N unit pixels
    A0 = ct.projector(N, np.array([0.0]), N, "parallel")
    assert np.allclose(np.asarray(A0.sum(axis=1)).ravel(), N)
    # pixel-driven B is close to, but not equal to, A'
    Bp = ct.backprojector_pixel_driven(N, np.arange(0, 180, 5.0))
    u = A @ ct.shepp_logan(N)
    rel = np.linalg.norm(A.T @ u - Bp @ u) / np.linalg.norm(A.T @ u)
    assert 1e-4 < rel < 0.1
    Bq = ct.backprojector_perturbed(A, 1e-2)
    assert abs(sp.linalg.norm(Bq - A.T) - 1e-2) < 1e-12


def test_all_four_title_claims_with_the_ptr_solvers():
    """run_equivalence_plots.m:33,44,55,66 and run_ptr_rtp_comparison.m:29 with the reference's
    own solver pairs (SURVEY App. A magnitudes)."""
    from oracle import ptr
    A, B, b, x_true = _problem()

    def X(f, *a):
        e = {}
        f(*a, extras=e)
        return e["X"]

    rel = lambda P, Q, k: np.linalg.norm(P[:, k] - Q[:, k]) / np.linalg.norm(Q[:, k])
    ba, lsmr = X(ptr.BAgmres_nonhybrid_bounds, A, B, b, x_true, 0.0, 8), X(oracle.lsmr_solver, A, b, x_true, 0.0, 8)
    ab, lsqr = X(ptr.ABgmres_nonhybrid_bounds, A, B, b, x_true, 0.0, 8), X(oracle.lsqr_solver, A, b, x_true, 0.0, 8)
    assert rel(ba, lsmr, 0) < 1e-15 and rel(ba, lsmr, 2) < 1e-12 and rel(ba, lsmr, 4) < 1e-6
    assert rel(ab, lsqr, 0) < 1e-15 and rel(ab, lsqr, 2) < 1e-12 and rel(ab, lsqr, 4) < 1e-6
    hba = X(ptr.BAgmres_hybrid_bounds, A, B, b, x_true, 0.0, 8, 1e-3)
    hlsmr = X(oracle.hybrid_lsmr_solver, A, b, x_true, 0.0, 8, 1e-3)
    assert rel(hba, hlsmr, 0) < 1e-15 and max(rel(hba, hlsmr, k) for k in range(5)) < 1e-5  # GKB breaks down at k~6
    hab = X(ptr.ABgmres_hybrid_bounds, A, B, b, x_true, 0.0, 8, 1e-3)
    hlsqr = X(oracle.hybrid_lsqr_solver, A, b, x_true, 0.0, 8, 1e-3)
    assert min(rel(hab, hlsqr, k) for k in range(8)) > 1e-2  # "(≠)"
    rtp = X(oracle.hybrid_ba_gmres_rtp, A, B, b, x_true, 0.0, 8, 1e-3)
    assert min(rel(hba, rtp, k) for k in range(8)) > 1e-2  # PTR ≠ RTP
