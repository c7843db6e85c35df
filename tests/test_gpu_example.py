"""the run_2D_phantom-shaped driver (examples/run_2D_phantom.py) on a small phantom —
its reconstructions and the mismatch sweep `B = A' + c E` (run_2D_phantom.m:79-102) against the oracle's
PTR / RTP solvers on the same host matrices."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_example():
    spec = importlib.util.spec_from_file_location("run_2D_phantom", os.path.join(ROOT, "examples", "run_2D_phantom.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_run_2d_phantom_driver_matches_oracle(hg, ctx):
    import oracle
    ex = _load_example()
    N, nv, maxit, lam, tol = 24, 60, 20, 1e-2, 1e-6
    levels = np.array([1e-4, 1e-2, 1.0])
    out = ex.run(N, nv, 0.25, maxit, lam, tol, levels, ctx=ctx, verbose=False)
    A, At, b, x_true = out["A"], out["At"], out["b"], out["x_true"]
    assert out["sinogram"].shape == (int(round(np.sqrt(2.0) * N)), nv)  # reshape(b, num_detectors, num_angles), :25
    assert abs(At - A.T).max() == 0  # the device transposition gives exactly A'
    ref = {"non-hybrid AB-GMRES": lambda B: oracle.ABgmres_nonhybrid_bounds(A, B, b, x_true, tol, maxit),
           "non-hybrid BA-GMRES": lambda B: oracle.BAgmres_nonhybrid_bounds(A, B, b, x_true, tol, maxit),
           "hybrid AB-GMRES (PTR)": lambda B: oracle.ABgmres_hybrid_bounds(A, B, b, x_true, tol, maxit, lam),
           "hybrid BA-GMRES (PTR)": lambda B: oracle.BAgmres_hybrid_bounds(A, B, b, x_true, tol, maxit, lam),
           "hybrid AB-GMRES (RTP)": lambda B: oracle.hybrid_ab_gmres_rtp(A, B, b, x_true, tol, maxit, lam),
           "hybrid BA-GMRES (RTP)": lambda B: oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam)}
    for name, (err, it) in out["recon"].items():
        xo, erro, reso, ito = ref[name](At)
        assert it == ito, name
        assert np.max(np.abs(err - erro) / erro) < 1e-7, name
    # the sweep: final errors of the four methods and the GCV-selected lambda per mismatch level
    names = list(ref)[:4]
    for i, c in enumerate(levels):
        B = sp.csr_matrix((At.data + out["E"][i], At.indices, At.indptr), shape=At.shape)
        for j, name in enumerate(names):
            erro = ref[name](B)[1]
            assert abs(out["sweep"][i, j] - erro[-1]) <= 1e-7 * erro[-1], (c, name)
        lam_o, fval_o, *_ = oracle.fminbnd(lambda l: oracle.gcv_function(l, A, B, b, A.shape[0], 20, "ba"), 1e-9, 1e-1, 1e-8)
        # flat objective: same minimum value, minimiser within the optimiser's resolution of it
        assert abs(out["sweep"][i, 5] - fval_o) <= 1e-8 * fval_o
        assert abs(out["sweep"][i, 4] - lam_o) <= 1e-3 * lam_o + 1e-8
