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
