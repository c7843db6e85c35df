"""CPU: the oracle reproduces the committed golden fixtures (guards the oracle against
drift; the fixtures are what the GPU tests and oracle/replay.m compare against)."""
import numpy as np
import pytest

import oracle
from tests.golden_util import NAMES, load


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    A, B, g = load(name)
    if not isinstance(A, np.ndarray):
        A, B = A.tocsr(), B.tocsr()
    b, x_true = g["b"], g["x_true"]
    tol, maxit, lam = float(g["tol"]), int(g["maxit"]), float(g["lam"])
    for key, f in (("ab_rtp", oracle.hybrid_ab_gmres_rtp), ("ba_rtp", oracle.hybrid_ba_gmres_rtp)):
        x, err, res, it = f(A, B, b, x_true, tol, maxit, lam)
        assert it == int(g[key + "_it"])
        k = min(it, 5)
        assert np.allclose(res[:k], g[key + "_res"][:k], rtol=1e-9)
        assert np.allclose(err[:k], g[key + "_err"][:k], rtol=1e-9)
    x, err, res, it = oracle.hybrid_lsqr_solver(A, b, x_true, tol, maxit, lam)
    assert it == int(g["hybrid_lsqr_it"]) and np.allclose(res[:5], g["hybrid_lsqr_res"][:5], rtol=1e-9)
    x, err, res, ar, it = oracle.lsmr_solver(A, b, x_true, tol, maxit)
    assert it == int(g["lsmr_it"]) and np.allclose(ar[:5], g["lsmr_ar"][:5], rtol=1e-8)
    for t in ("ab", "ba"):
        vals = [oracle.gcv_function(l, A, B, b, A.shape[0], int(g["k_gcv"]), t) for l in g["gcv_lams"]]
        assert np.allclose(vals, g[f"gcv_{t}_vals"], rtol=1e-6)


def test_golden_inputs_are_matlab_shaped():
    """CSC with sorted row indices and int64 index arrays — what a MEX gateway receives."""
    A, B, g = load("ct16_perturbed")
    assert A.format == "csc" and A.has_sorted_indices
    assert g["A_jc"].dtype == np.int64 and g["A_ir"].dtype == np.int64
    assert A.shape == (B.shape[1], B.shape[0])
