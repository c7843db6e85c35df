"""GPU parity of the project-then-regularise solvers against the
oracle restatement of the *_bounds.m solve paths, plus the reference's figure-title claims
(run_equivalence_plots.m:33,44,55,66; run_ptr_rtp_comparison.m:29,39) on the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-8
NAMES = ["ABgmres_hybrid_bounds", "BAgmres_hybrid_bounds", "ABgmres_nonhybrid_bounds", "BAgmres_nonhybrid_bounds"]


def _rel(Xd, Xo):
    k = min(Xd.shape[1], Xo.shape[1])
    return np.array([np.linalg.norm(Xd[:, i] - Xo[:, i]) / np.linalg.norm(Xo[:, i]) for i in range(k)])


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("problem", ["ct64", "ct48_unmatched"])
def test_ptr_vs_oracle(hg, ctx, request, name, problem):
    import oracle
    A, B, b, x_true = request.getfixturevalue(problem)
    args = (A, B, b, x_true, 1e-6, 40) + ((1e-2,) if "_hybrid" in name else ())
    ed, eo = {}, {}
    x, err, res, it = getattr(hg, name)(*args, ctx=ctx, extras=ed)
    xo, erro, reso, ito = getattr(oracle, name)(*args, extras=eo)
    assert it == ito
    assert np.max(np.abs(res - reso) / reso) < TOL
    assert np.max(np.abs(err - erro) / erro) < TOL
    assert np.max(_rel(ed["X"], eo["X"])) < TOL
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < TOL


def test_ptr_stop_rule_and_gcv_pipeline(hg, ctx, ct64):
    """plot_error_vs_mismatch_norm.m:46-57: lambda from fminbnd(gcv_function), then the PTR hybrid
    solve with that lambda; the stopping iteration must match the oracle's."""
    import oracle
    A, B, b, x_true = ct64
    lam, fval, cnt = hg.fminbnd_gcv(A, B, b, A.shape[0], 20, "ba", 1e-9, 1e-1, 1e-8, ctx=ctx)
    for tol in (0.1, 0.03):
        x, err, res, it = hg.BAgmres_hybrid_bounds(A, B, b, x_true, tol, 40, lam, ctx=ctx)
        xo, erro, reso, ito = oracle.BAgmres_hybrid_bounds(A, B, b, x_true, tol, 40, lam)
        assert it == ito and len(res) == it


def test_reference_title_claims_on_device(hg, ctx):
    """deriv2 n=32, B=A', 1% noise, lambda=1e-3 (run_equivalence_plots.m:3-11): BA-GMRES == LSMR,
    AB-GMRES == LSQR, hybrid BA == hybrid LSMR (k=1 exactly, later ~1e-6), hybrid AB != hybrid
    LSQR, PTR != RTP — with every solver running on the GPU."""
    import oracle
    from oracle.generators import add_noise
    A, b_exact, x_true = oracle.generate_test_problem("deriv2", 32)
    B = A.T.copy()
    b = add_noise(b_exact, 1e-2, 0)

    def X(f, *a):
        e = {}
        f(*a, ctx=ctx, extras=e)
        return e["X"]

    ba, lsmr = X(hg.BAgmres_nonhybrid_bounds, A, B, b, x_true, 0.0, 5), X(hg.lsmr_solver, A, b, x_true, 0.0, 5)
    ab, lsqr = X(hg.ABgmres_nonhybrid_bounds, A, B, b, x_true, 0.0, 5), X(hg.lsqr_solver, A, b, x_true, 0.0, 5)
    hba = X(hg.BAgmres_hybrid_bounds, A, B, b, x_true, 0.0, 5, 1e-3)
    hlsmr = X(hg.hybrid_lsmr_solver, A, b, x_true, 0.0, 5, 1e-3)
    hab = X(hg.ABgmres_hybrid_bounds, A, B, b, x_true, 0.0, 5, 1e-3)
    hlsqr = X(hg.hybrid_lsqr_solver, A, b, x_true, 0.0, 5, 1e-3)
    rtp = X(hg.hybrid_ba_gmres_rtp, A, B, b, x_true, 0.0, 5, 1e-3)
    assert _rel(ba, lsmr)[0] < 1e-13 and _rel(ba, lsmr)[2] < 1e-10          # (≡)
    assert _rel(ab, lsqr)[0] < 1e-13 and _rel(ab, lsqr)[2] < 1e-10          # (≡)
    assert _rel(hba, hlsmr)[0] < 1e-13 and np.max(_rel(hba, hlsmr)) < 1e-4  # (≡)
    assert np.min(_rel(hab, hlsqr)) > 1e-2                                  # (≠)
    assert np.min(_rel(hba, rtp)) > 1e-2                                    # PTR ≠ RTP
