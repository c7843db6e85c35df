"""GPU: every hot-path entry point against the committed golden vectors, with inputs
handed over the way MATLAB would (CSC / full column-major)."""
import numpy as np
import pytest

from tests.golden_util import NAMES, load, strict_iters

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _rel_cols(X, G, k):
    return max(np.linalg.norm(X[:, i] - G[:, i]) / np.linalg.norm(G[:, i]) for i in range(k))


@pytest.mark.parametrize("name", NAMES)
def test_solvers_against_golden(hg, ctx, name):
    A, B, g = load(name)
    b, x_true = g["b"], g["x_true"]
    tol, maxit, lam = float(g["tol"]), int(g["maxit"]), float(g["lam"])
    ks = strict_iters(name)
    for key, f in (("ab_rtp", hg.hybrid_ab_gmres_rtp), ("ba_rtp", hg.hybrid_ba_gmres_rtp)):
        ex = {}
        x, err, res, it = f(A, B, b, x_true, tol, maxit, lam, ctx=ctx, extras=ex)
        assert it == int(g[key + "_it"])
        k = min(it, ks)
        assert np.max(np.abs(res[:k] - g[key + "_res"][:k]) / g[key + "_res"][:k]) < TOL
        assert np.max(np.abs(err[:k] - g[key + "_err"][:k]) / g[key + "_err"][:k]) < TOL
        assert _rel_cols(ex["X"], g[key + "_X"], k) < TOL
        assert abs(ex["beta"] - float(g[key + "_beta"])) / float(g[key + "_beta"]) < 1e-13
        if name != "deriv2_n32":  # CT problems: the whole history holds
            assert np.max(np.abs(res - g[key + "_res"]) / g[key + "_res"]) < TOL
            assert np.linalg.norm(x - g[key + "_x"]) / np.linalg.norm(g[key + "_x"]) < TOL
    for key, f in (("hybrid_lsqr", hg.hybrid_lsqr_solver), ("hybrid_lsmr", hg.hybrid_lsmr_solver)):
        ex = {}
        x, err, res, it = f(A, b, x_true, tol, maxit, lam, ctx=ctx, extras=ex)
        assert it == int(g[key + "_it"])
        assert _rel_cols(ex["X"], g[key + "_X"], ks) < TOL
        assert np.max(np.abs(res[:ks] - g[key + "_res"][:ks]) / g[key + "_res"][:ks]) < TOL
    ex = {}
    x, err, res, it = hg.lsqr_solver(A, b, x_true, tol, maxit, ctx=ctx, extras=ex)
    assert it == int(g["lsqr_it"]) and _rel_cols(ex["X"], g["lsqr_X"], ks) < TOL
    ex = {}
    x, err, res, ar, it = hg.lsmr_solver(A, b, x_true, tol, maxit, ctx=ctx, extras=ex)
    assert it == int(g["lsmr_it"]) and _rel_cols(ex["X"], g["lsmr_X"], ks) < TOL
    assert np.max(np.abs(ar[:ks] - g["lsmr_ar"][:ks]) / g["lsmr_ar"][:ks]) < 1e-7


@pytest.mark.parametrize("name", ["ct16_perturbed", "ct20_fan_pixel"])
@pytest.mark.parametrize("gcv_type", ["ab", "ba"])
def test_gcv_against_golden(hg, ctx, name, gcv_type):
    A, B, g = load(name)
    b = g["b"]
    k_gcv = int(g["k_gcv"])
    prob = hg.gcv_prepare(A, B, b, A.shape[0], k_gcv, gcv_type, ctx=ctx)
    H, beta = prob.get(k_gcv)
    G = g[f"gcv_{gcv_type}_H"]
    assert abs(beta - float(g[f"gcv_{gcv_type}_beta"])) / beta < 1e-13
    for k in range(k_gcv):
        assert np.linalg.norm(H[: k + 2, k] - G[: k + 2, k]) / np.linalg.norm(G[: k + 2, k]) < TOL
    vals = np.array([prob.eval(l) for l in g["gcv_lams"]])
    assert np.max(np.abs(vals - g[f"gcv_{gcv_type}_vals"]) / g[f"gcv_{gcv_type}_vals"]) < 1e-7


# ------------------------------------------------------------------------------------------
# against fixtures produced by executing the reference's own .m files (make_reference_golden.py)
# ------------------------------------------------------------------------------------------
from tests.golden_util import GOLDEN_DIR, REF_NAMES, load_ref, ref_strict_iters  # noqa: E402


def _relmax(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(b)))


@pytest.mark.parametrize("name", REF_NAMES)
def test_device_against_executed_reference(hg, ctx, name):
    """Every §8a entry point vs the outputs of the untouched reference source on the same inputs:
    same stopping iteration, histories / Hessenberg columns / iterates within 1e-8 over the
    iterations where the reference's own arithmetic is reproducible (all 25 on the CT cases)."""
    A, B, b, x_true, tol, maxit, lam, k_gcv, r = load_ref(name)
    ct = name.startswith("ref_ct")
    ks = ref_strict_iters(name)
    for fn in ("hybrid_ab_gmres_rtp", "hybrid_ba_gmres_rtp"):
        ex = {}
        x, err, res, it = getattr(hg, fn)(A, B, b, x_true, tol, maxit, lam, ctx=ctx, extras=ex)
        assert it == int(r[fn + "_it"]) and len(res) == len(r[fn + "_res"])
        k = min(it, ks)
        assert _relmax(res[:k], r[fn + "_res"][:k]) < TOL, fn
        assert _relmax(err[:k], r[fn + "_err"][:k]) < TOL, fn
        assert abs(ex["beta"] - float(r[fn + "_beta"])) < 1e-13 * ex["beta"]
        Hr = r[fn + "_H"]
        for j in range(k):
            assert np.linalg.norm(ex["H"][: j + 2, j] - Hr[: j + 2, j]) / np.linalg.norm(Hr[: j + 2, j]) < TOL, (fn, j)
        assert _rel_cols(ex["X"], r[fn + "_X"], k) < (TOL if ct else 1e-6)
        if ct:
            assert np.linalg.norm(x - r[fn + "_x"]) / np.linalg.norm(r[fn + "_x"]) < TOL
    for fn in ("ABgmres_hybrid_bounds", "BAgmres_hybrid_bounds", "ABgmres_nonhybrid_bounds", "BAgmres_nonhybrid_bounds"):
        args = (A, B, b, x_true, tol, maxit) + ((lam,) if "_hybrid" in fn else ())
        x, err, res, it = getattr(hg, fn)(*args, ctx=ctx)
        assert it == int(r[fn + "_it"])
        k = min(it, ks if ct else 3)
        assert _relmax(res[:k], r[fn + "_res"][:k]) < TOL and _relmax(err[:k], r[fn + "_err"][:k]) < TOL, fn
    kg = min(ks, 8)  # Golub-Kahan without reorthogonalisation: see test_gkb_solvers_vs_oracle
    for fn in ("hybrid_lsqr_solver", "hybrid_lsmr_solver"):
        ex = {}
        x, err, res, it = getattr(hg, fn)(A, b, x_true, tol, maxit, lam, ctx=ctx, extras=ex)
        assert it == int(r[fn + "_it"])
        assert _relmax(res[:kg], r[fn + "_res"][:kg]) < TOL and _rel_cols(ex["X"], r[fn + "_X"], kg) < (TOL if ct else 1e-6)
    x, err, res, it = hg.lsqr_solver(A, b, x_true, tol, maxit, ctx=ctx)
    assert it == int(r["lsqr_solver_it"]) and _relmax(res[:kg], r["lsqr_solver_res"][:kg]) < TOL
    x, err, res, ar, it = hg.lsmr_solver(A, b, x_true, tol, maxit, ctx=ctx)
    assert it == int(r["lsmr_solver_it"]) and _relmax(res[:kg], r["lsmr_solver_res"][:kg]) < TOL
    assert _relmax(ar[:kg], r["lsmr_solver_ar"][:kg]) < 1e-7
    if ct:
        x, err, res, ar, it = hg.lsmr_solver(A, b, ctx=ctx)  # lsmr_solver.m:3-5 defaults, NaN error history
        assert it == int(r["lsmr_defaults_it"]) and np.all(np.isnan(err))
    for t in ("ab", "ba"):
        prob = hg.gcv_prepare(A, B, b, A.shape[0], k_gcv, t, ctx=ctx)
        H, beta = prob.get(k_gcv)
        Hr = r[f"gcv_{t}_H"]
        assert abs(beta - float(r[f"gcv_{t}_beta"])) < 1e-13 * beta
        for j in range(min(ks, k_gcv)):
            assert np.linalg.norm(H[: j + 2, j] - Hr[: j + 2, j]) / np.linalg.norm(Hr[: j + 2, j]) < TOL
        if ct:
            vals = np.array([prob.eval(l) for l in r["gcv_lams"]])
            assert _relmax(vals, r[f"gcv_{t}_vals"]) < 1e-7
            lam_d, fval, cnt, _ = prob.fminbnd(1e-9, 1e-1, 1e-8)  # analyze_regularization.m:37-46
            lam_r = float(r[f"gcv_{t}_fminbnd_lambda"])
            assert abs(lam_d - lam_r) <= 1e-3 * lam_r
            assert abs(prob.eval(lam_d) - prob.eval(lam_r)) <= 1e-8 * prob.eval(lam_r)
            # plot_gcv_surface.m:58-102 from a device Arnoldi of the surface's depth
            K = int(r["surface_k"])
            ps = hg.gcv_prepare(A, B, b, A.shape[0], K, t, ctx=ctx)
            surf, path = ps.surface(r["surface_lams"], K)
            assert np.max(np.abs(surf - r[f"surface_{t}"]) / np.abs(r[f"surface_{t}"])) < 1e-7
            assert np.array_equal(path, r[f"surface_{t}_path"])


def test_breakdown_epilogue_matches_executed_reference(hg, ctx):
    """H(2,1) == 0 at k = 1 (A = B = I, b = e1): hybrid_ab_gmres_rtp.m:25,41-43 leaves x unassigned,
    hybrid_ba_gmres_rtp.m:25,38-40 returns its zeros; both report niters = 1 and the untouched zero
    history entries — SURVEY §8a row a8, against what the reference source does."""
    import os
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_breakdown.npz")))
    n = int(r["n"])
    I, e1, xt = np.eye(n), np.eye(n)[:, 0].copy(), np.ones(n)
    ex = {}
    x, err, res, it = hg.hybrid_ab_gmres_rtp(I, I, e1, xt, 1e-6, 4, 1e-2, ctx=ctx, extras=ex)
    assert not bool(r["ab_x_assigned"]) and x is None
    assert it == int(r["ab_it"]) == 1
    assert np.array_equal(res, r["ab_res"]) and np.array_equal(err, r["ab_err"])
    assert ex["H"][1, 0] == 0.0 and abs(ex["H"][0, 0] - r["ab_H"][0, 0]) < 1e-15
    x, err, res, it = hg.hybrid_ba_gmres_rtp(I, I, e1, xt, 1e-6, 4, 1e-2, ctx=ctx)
    assert it == int(r["ba_it"]) == 1 and np.array_equal(x, r["ba_x"])
    assert np.array_equal(res, r["ba_res"]) and np.array_equal(err, r["ba_err"])
    import scipy.sparse as sp
    x, err, res, it = hg.hybrid_ab_gmres_rtp(sp.identity(n, format="csc"), sp.identity(n, format="csr"), e1, xt,
                                             1e-6, 4, 1e-2, ctx=ctx)
    assert x is None and it == 1 and res[0] == 0.0


@pytest.mark.parametrize("name", ["ref_ct16_perturbed", "ref_ct20_fan_pixel"])
@pytest.mark.parametrize("t", ["ab", "ba"])
def test_lambda_k_hybrid_solve_against_executed_reference(hg, ctx, name, t):
    """hg_gmres_ptr_gcv (lambda chosen at every iteration from the growing H): the
    lambda_k path equals the reference's compute_gcv_surface path exactly, the iterates equal what the
    reference's PTR solver gives with that lambda_k."""
    A, B, b, x_true, tol, maxit, lam, k_gcv, r = load_ref(name)
    K = int(r["surface_k"])
    ex = {}
    x, err, res, it, path = hg.hybrid_gmres_gcv(t, A, B, b, x_true, 0.0, K, r["surface_lams"], ctx=ctx, extras=ex)
    assert it == K and np.array_equal(path, r[f"surface_{t}_path"])
    assert _rel_cols(ex["X"], r[f"lamk_{t}_X"], K) < TOL
    assert _relmax(err, r[f"lamk_{t}_err"]) < TOL and _relmax(res, r[f"lamk_{t}_res"]) < TOL
    assert np.linalg.norm(x - r[f"lamk_{t}_X"][:, -1]) <= TOL * np.linalg.norm(x)


def test_lambda_k_hybrid_solve_on_ct_problem(hg, ctx, ct48_unmatched):
    """the same mode on a larger unmatched CT problem against the oracle restatement (oracle/ptr.py)"""
    import oracle
    A, B, b, x_true = ct48_unmatched
    lams = np.logspace(-8, -1, 40)
    for t in ("ab", "ba"):
        x, err, res, it, path = hg.hybrid_gmres_gcv(t, A, B, b, x_true, 1e-6, 25, lams, ctx=ctx)
        xo, erro, reso, ito, patho = oracle.hybrid_gmres_gcv(t, A, B, b, x_true, 1e-6, 25, lams)
        assert it == ito and np.array_equal(path, patho)
        assert _relmax(err, erro) < 1e-7 and _relmax(res, reso) < 1e-7
