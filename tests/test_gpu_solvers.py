"""GPU parity of the whole solvers (called through the C ABI with the
reference's own argument lists) against the literal oracle restatement.

Tolerance: BASELINE.json's north star — per-iteration residual norms, Arnoldi
coefficients and iterates within 1e-8 relative over the first 50 iterations;
stopping iteration identical.  Arnoldi coefficients are compared norm-wise per
column of H (element-wise is meaningless: with matched B the upper part of H is
rounding noise, SURVEY.md §7 hard part 1).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-8   # north-star tolerance vs the literal (MGS) reference restatement
TIGHT = 1e-10  # same algorithm (CGS2), different summation order


def _colwise(Hd, Ho, kmax):
    out = []
    for k in range(kmax):
        out.append(np.linalg.norm(Hd[: k + 2, k] - Ho[: k + 2, k]) / np.linalg.norm(Ho[: k + 2, k]))
    return np.array(out)


def _iter_rel(Xd, Xo):
    k = min(Xd.shape[1], Xo.shape[1])
    return np.array([np.linalg.norm(Xd[:, i] - Xo[:, i]) / np.linalg.norm(Xo[:, i]) for i in range(k)])


@pytest.mark.parametrize("solver", ["ba", "ab"])
@pytest.mark.parametrize("problem", ["ct64", "ct48_unmatched"])
def test_rtp_vs_oracle(hg, ctx, request, solver, problem):
    import oracle
    A, B, b, x_true = request.getfixturevalue(problem)
    maxit, lam, tol = 50, 1e-2, 1e-6
    f_dev = hg.hybrid_ba_gmres_rtp if solver == "ba" else hg.hybrid_ab_gmres_rtp
    f_orc = oracle.hybrid_ba_gmres_rtp if solver == "ba" else oracle.hybrid_ab_gmres_rtp
    ex_d = {}
    x, err, res, it = f_dev(A, B, b, x_true, tol, maxit, lam, ctx=ctx, extras=ex_d)
    for orth, bound in (("mgs", TOL), ("cgs2", TIGHT)):
        ex_o = {}
        xo, erro, reso, ito = f_orc(A, B, b, x_true, tol, maxit, lam, orth=orth, extras=ex_o)
        assert it == ito
        assert np.max(np.abs(res - reso) / reso) < bound
        assert np.max(np.abs(err - erro) / erro) < bound
        assert np.max(_iter_rel(ex_d["X"], ex_o["X"])) < bound
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < bound
        assert abs(ex_d["beta"] - ex_o["beta"]) / ex_o["beta"] < 1e-13
        assert np.max(_colwise(ex_d["H"], ex_o["H"], it)) < bound


@pytest.mark.parametrize("solver", ["ba", "ab"])
def test_rtp_literal_residual_mode(hg, ctx, ct64, solver):
    """residual_mode=1 recomputes b - A*x with an SpMV exactly as the reference does."""
    import oracle
    A, B, b, x_true = ct64
    f_dev = hg.hybrid_ba_gmres_rtp if solver == "ba" else hg.hybrid_ab_gmres_rtp
    f_orc = oracle.hybrid_ba_gmres_rtp if solver == "ba" else oracle.hybrid_ab_gmres_rtp
    x, err, res, it = f_dev(A, B, b, x_true, 1e-6, 20, 1e-2, ctx=ctx, residual_mode=1)
    xo, erro, reso, ito = f_orc(A, B, b, x_true, 1e-6, 20, 1e-2)
    assert it == ito
    assert np.max(np.abs(res - reso) / reso) < TOL
    assert np.max(np.abs(err - erro) / erro) < TOL


def test_rtp_stops_on_tolerance(hg, ctx, ct64):
    """Stopping iteration must match exactly (<= tol, hybrid_ba_gmres_rtp.m:35)."""
    import oracle
    A, B, b, x_true = ct64
    for tol in (0.2, 0.05, 0.02):
        for f_dev, f_orc in ((hg.hybrid_ba_gmres_rtp, oracle.hybrid_ba_gmres_rtp),
                             (hg.hybrid_ab_gmres_rtp, oracle.hybrid_ab_gmres_rtp)):
            x, err, res, it = f_dev(A, B, b, x_true, tol, 40, 1e-2, ctx=ctx)
            xo, erro, reso, ito = f_orc(A, B, b, x_true, tol, 40, 1e-2)
            assert it == ito and len(res) == it
            assert res[-1] <= tol or it == 40


def test_rtp_on_reference_dense_inputs(hg, ctx):
    """run_ptr_rtp_comparison.m:4-19: deriv2 n=32, full A, B=A', lambda=1e-3, maxit=n.
    Arnoldi breaks down numerically on this problem: only BA-RTP
    iterates are reproducible beyond k~6, AB-RTP is checked on the first 5."""
    import oracle
    from oracle.generators import add_noise
    A, b_exact, x_true = oracle.generate_test_problem("deriv2", 32)
    B = A.T.copy()
    b = add_noise(b_exact, 1e-2, 0)
    ex_d, ex_o = {}, {}
    x, err, res, it = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 32, 1e-3, ctx=ctx, extras=ex_d)
    xo, erro, reso, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 32, 1e-3, extras=ex_o)
    assert it == ito
    assert np.max(_iter_rel(ex_d["X"], ex_o["X"])) < 1e-7
    assert np.max(np.abs(err - erro) / erro) < 1e-7
    ex_d, ex_o = {}, {}
    x, err, res, it = hg.hybrid_ab_gmres_rtp(A, B, b, x_true, 1e-6, 5, 1e-3, ctx=ctx, extras=ex_d)
    xo, erro, reso, ito = oracle.hybrid_ab_gmres_rtp(A, B, b, x_true, 1e-6, 5, 1e-3, extras=ex_o)
    assert it == ito
    assert np.max(_iter_rel(ex_d["X"], ex_o["X"])) < 1e-6


def test_arnoldi_handle_matches_solver(hg, ctx, ct64):
    """hg_arnoldi_* (the bench path, no host sync per step) gives the same H as the solver."""
    A, B, b, x_true = ct64
    dA, dB = hg.DeviceMatrix.from_any(A, ctx), hg.DeviceMatrix.from_any(B, ctx)
    ar = hg.Arnoldi(dA, dB, "n", 30)
    ar.set_rhs(b)
    ar.reset(1e-2)
    ar.steps(30)
    H, beta, k = ar.get()
    ex = {}
    hg.hybrid_ba_gmres_rtp(dA, dB, b, x_true, 0.0, 30, 1e-2, ctx=ctx, extras=ex)
    assert k == 30
    assert np.array_equal(H, ex["H"]) and beta == ex["beta"]  # deterministic reductions
    # orthonormal basis
    Q = np.stack([ar.q(j) for j in range(31)], axis=1)
    assert np.linalg.norm(Q.T @ Q - np.eye(31)) < 1e-13
    # rerun is bit-identical
    ar.reset(1e-2)
    ar.steps(30)
    H2, beta2, _ = ar.get()
    assert np.array_equal(H, H2) and beta == beta2


@pytest.mark.parametrize("gcv_type", ["ab", "ba"])
def test_gcv_function_on_ct(hg, ctx, ct64, gcv_type):
    """gcv_function.m on the config-1 CT problem: the device Arnoldi (run once, memoised) and
    the host projected part reproduce the oracle's GCV values.  On this problem the GCV curve
    is flat to 9 digits over [1e-9,1e-1], so the minimiser itself is decided by rounding noise
    (the oracle disagrees with itself under a summation-order change); the parity statement is
    therefore on the objective: both minimisers reach the same GCV value."""
    import oracle
    from oracle.solvers import gcv_arnoldi
    A, B, b, x_true = ct64
    m = A.shape[0]
    prob = hg.gcv_prepare(A, B, b, m, 20, gcv_type, ctx=ctx)
    Hd, betad = prob.get(20)
    Ho, betao = gcv_arnoldi(A, B, b, m, 20, gcv_type)
    assert abs(betad - betao) / betao < 1e-13
    assert np.max(_colwise(Hd, Ho, 20)) < TIGHT
    for lam in (1e-9, 1e-6, 1e-3, 1e-1):
        v_dev = hg.gcv_function(lam, A, B, b, m, 20, gcv_type, ctx=ctx)
        v_orc = oracle.gcv_function(lam, A, B, b, m, 20, gcv_type)
        assert abs(v_dev - v_orc) / v_orc < 1e-10
    lam_o, f_o, flag, cnt_o = oracle.fminbnd(lambda l: oracle.gcv_function(l, A, B, b, m, 20, gcv_type),
                                             1e-9, 1e-1, 1e-8)
    lam_d, f_d, cnt_d, tr_d = prob.fminbnd(1e-9, 1e-1, 1e-8)
    assert abs(f_d - f_o) / f_o < 1e-10
    assert abs(lam_d - lam_o) / lam_o < 1e-2
    # the selected lambda drives the same stopping iteration in the solver
    x, err, res, it = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.05, 40, lam_d, ctx=ctx)
    xo, erro, reso, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.05, 40, lam_o)
    assert it == ito


@pytest.mark.parametrize("name,gcv_type", [("shaw", "ab"), ("shaw", "ba"), ("heat", "ab"), ("deriv2", "ba")])
def test_gcv_fminbnd_on_reference_call_sites(hg, ctx, name, gcv_type):
    """analyze_regularization.m:35-46 / plot_error_vs_mismatch_norm.m:46-50: fminbnd over
    gcv_function on [1e-9,1e-1], TolX=1e-8, k_gcv=20, n=32 dense problems.  The selected
    lambda must agree within the optimiser's own TolX and the evaluation count must match."""
    import oracle
    from oracle.generators import add_noise
    A, b_exact, x_true = oracle.generate_test_problem(name, 32)
    B = A.T.copy()
    b = add_noise(b_exact, 1e-2, 0)
    lam_o, f_o, flag, cnt_o = oracle.fminbnd(lambda l: oracle.gcv_function(l, A, B, b, 32, 20, gcv_type),
                                             1e-9, 1e-1, 1e-8)
    lam_d, f_d, cnt_d = hg.fminbnd_gcv(A, B, b, 32, 20, gcv_type, 1e-9, 1e-1, 1e-8, ctx=ctx)
    assert abs(lam_d - lam_o) <= 1e-8, (lam_d, lam_o)
    assert abs(cnt_d - cnt_o) <= 1
    # memoised objective == fminbnd's objective
    assert hg.gcv_function(lam_d, A, B, b, 32, 20, gcv_type, ctx=ctx) == f_d


def test_gcv_early_breakdown_keeps_zero_columns(hg, ctx):
    """gcv_function.m:30,33: after `H(k+1,k) < 1e-12` the trailing zero columns stay."""
    import oracle
    from oracle.generators import add_noise
    A, b_exact, x_true = oracle.generate_test_problem("shaw", 32)
    B = A.T.copy()
    b = add_noise(b_exact, 1e-2, 0)
    for t in ("ab", "ba"):
        for lam in (1e-6, 1e-3):
            v_dev = hg.gcv_function(lam, A, B, b, 32, 20, t, ctx=ctx)
            v_orc = oracle.gcv_function(lam, A, B, b, 32, 20, t)
            assert np.isfinite(v_dev)
            assert abs(v_dev - v_orc) / v_orc < 1e-5  # Arnoldi has broken down: loose


@pytest.mark.parametrize("name", ["hybrid_lsqr_solver", "hybrid_lsmr_solver", "lsqr_solver", "lsmr_solver"])
def test_gkb_solvers_vs_oracle(hg, ctx, ct48_unmatched, name):
    """Golub-Kahan solvers vs the oracle.  Without reorthogonalisation the GKB recurrences
    amplify rounding differences once Ritz values converge: the oracle run on an input
    perturbed at the 1e-16 level departs from itself by 1e-8 at k~12 and 1e-2 by k~15 on this
    problem (see DESIGN.md "Parity").  So: strict 1e-8 for as many iterations as the oracle's own
    sensitivity permits (10-11 here, computed below and asserted), a graded bound of 5 x that sensitivity on
    the well-determined stretch (2e-14 at k <= 5), and a bound scaled by the sensitivity afterwards."""
    import oracle
    A, B, b, x_true = ct48_unmatched
    maxit, lam, tol = 40, 1e-2, 1e-6
    f_dev, f_orc = getattr(hg, name), getattr(oracle, name)

    def run(f, rhs, **kw):
        ex = {}
        if name.startswith("hybrid"):
            out = f(A, rhs, x_true, tol, maxit, lam, extras=ex, **kw)
        else:
            out = f(A, rhs, x_true, tol, maxit, extras=ex, **kw)
        return out, ex["X"]

    out_d, Xd = run(f_dev, b, ctx=ctx)
    out_o, Xo = run(f_orc, b)
    assert out_d[-1] == out_o[-1]  # iterations
    sens = np.zeros(Xo.shape[1])
    for seed in (11, 12, 13):  # the oracle against itself, b perturbed in the last bit
        b_pert = b * (1.0 + 2.2e-16 * np.random.default_rng(seed).standard_normal(b.shape))
        sens = np.maximum(sens, np.maximum.accumulate(_iter_rel(run(f_orc, b_pert)[1], Xo)))
    diff = _iter_rel(Xd, Xo)
    # strict bar for as long as the problem determines the iterates to it: the leading iterations on which the
    # oracle's own sensitivity stays below the bar (10 for hybrid LSQR, 11 for the other three; the device
    # differs from the oracle by 2e-10 / 2e-10 / 4e-10 / 1e-10 at the last of them)
    k_strict = int(np.argmax(sens >= TOL)) if np.any(sens >= TOL) else len(sens)
    assert k_strict >= 10, (name, k_strict)
    assert np.max(diff[:k_strict]) < TOL, (name, k_strict, diff[:k_strict])
    # graded: while the iterates are well determined the device is within 5 x the oracle's self-sensitivity
    # (measured <= 1.1 x), floored at rounding level; through the blow-up (k >= 14) within 1e3 x
    m = min(len(diff), len(sens))
    well = sens[:m] < 1e-9
    assert np.all(diff[:m][well] <= np.maximum(2e-14, 5 * sens[:m])[well]), (name, diff[:14], sens[:14])
    assert np.max(diff[:5]) < 2e-14
    assert np.all(diff <= np.maximum(TOL, 1e3 * sens[: len(diff)]))
    for hd, ho in zip(out_d[1:-1], out_o[1:-1]):  # histories
        assert hd.shape == ho.shape
        assert np.max(np.abs(hd[:k_strict] - ho[:k_strict]) / np.abs(ho[:k_strict])) < TOL
        rel = np.abs(hd - ho) / np.abs(ho)
        assert np.all(rel <= np.maximum(TOL, 1e3 * sens[: len(rel)]))


@pytest.mark.parametrize("name", ["hybrid_lsqr_solver", "hybrid_lsmr_solver", "lsmr_solver"])
def test_gkb_residual_from_relation_matches_literal(hg, ctx, ct48_unmatched, name):
    """Default: `b - A*x` (hybrid_lsqr_solver.m:43, hybrid_lsmr_solver.m:48, lsmr_solver.m:69) comes from the
    Golub-Kahan relation (`A v_k = alpha_k u_k + beta_{k+1} u_{k+1}`: vector recurrences for LSQR / LSMR,
    `U_{k+1}(beta1 e1 - B_k y)` for hybrid LSMR) instead of one more product with A per iteration; `gkb_residual`
    = 1 forms it literally.  The iterates do not depend on it (bit-identical); the residual histories (and
    `norm(A'*r)` of lsmr_solver, a product with the recurrence's r) agree to rounding."""
    A, B, b, x_true = ct48_unmatched
    args = (1e-6, 40) if name == "lsmr_solver" else (1e-6, 40, 1e-2)
    out = {}
    try:
        for mode in (0, 1):
            hg.set_option("gkb_residual", mode)
            ex = {}
            out[mode] = (getattr(hg, name)(A, b, x_true, *args, ctx=ctx, extras=ex), ex["X"])
    finally:
        hg.set_option("gkb_residual", 0)
    (r0, X0), (r1, X1) = out[0], out[1]
    assert r0[-1] == r1[-1] == 40
    assert np.array_equal(X0, X1) and np.array_equal(r0[0], r1[0]) and np.array_equal(r0[1], r1[1])
    assert np.max(np.abs(r0[2] - r1[2]) / r1[2]) < 1e-11
    if name == "lsmr_solver":
        assert np.max(np.abs(r0[3] - r1[3]) / r1[3]) < 1e-9
    # and it stops where the literal form stops
    tol = float(r1[2][24]) * (1 + 1e-9)
    try:
        its = []
        for mode in (0, 1):
            hg.set_option("gkb_residual", mode)
            its.append(getattr(hg, name)(A, b, x_true, tol, *args[1:], ctx=ctx)[-1])
    finally:
        hg.set_option("gkb_residual", 0)
    assert its[0] == its[1] <= 25


def test_lsmr_defaults_and_missing_x_true(hg, ctx):
    """lsmr_solver.m:3-5 defaults; err_hist is NaN without x_true (:28,72-74)."""
    import oracle
    A, b_exact, x_true = oracle.generate_test_problem("heat", 32)
    x, err, res, ar, it = hg.lsmr_solver(A, b_exact, ctx=ctx)
    xo, erro, reso, aro, ito = oracle.lsmr_solver(A, b_exact)
    assert it == ito
    assert np.all(np.isnan(err))
    k = min(8, it)
    assert np.max(np.abs(res[:k] - reso[:k]) / reso[:k]) < 1e-6


def test_gkb_with_preuploaded_transpose(hg, ctx, ct48_unmatched):
    A, B, b, x_true = ct48_unmatched
    dA = hg.DeviceMatrix.from_any(A, ctx)
    dAt = dA.transpose()
    r1 = hg.lsqr_solver(dA, b, x_true, 1e-6, 10, ctx=ctx, At=dAt)
    r2 = hg.lsqr_solver(A, b, x_true, 1e-6, 10, ctx=ctx)
    assert np.array_equal(r1[0], r2[0])


def test_cross_solver_identity(hg, ctx, ct64):
    """AB-RTP with matched B equals hybrid LSQR (same Krylov space and Tikhonov
    projection) on the first iterations."""
    A, B, b, x_true = ct64
    e1, e2 = {}, {}
    hg.hybrid_ab_gmres_rtp(A, B, b, x_true, 0.0, 5, 1e-2, ctx=ctx, extras=e1)
    hg.hybrid_lsqr_solver(A, b, x_true, 0.0, 5, 1e-2, ctx=ctx, extras=e2)
    assert np.max(_iter_rel(e1["X"], e2["X"])) < 1e-8


def test_cgs_fused_option_matches_default(hg, ctx, ct64):
    """The one-pass CGS2 middle stage (update + second-pass dot products from a shared-memory tile,
    `cgs_fused` = 2) gives the same Arnoldi factorisation as the separate kernels (`cgs_fused` = 0).  The
    whole-step kernel is switched off so that the middle stage under test is the one that runs; the removed
    L2-re-read variant (1) is refused."""
    A, B, b, x_true = ct64
    out = {}
    try:
        hg.set_option("cgs_step_max_n", 0)
        for fused in (0, 2):
            hg.set_option("cgs_fused", fused)
            ex = {}
            r = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 45, 1e-2, ctx=ctx, extras=ex)
            out[fused] = (r, ex)
        with pytest.raises(Exception):
            hg.set_option("cgs_fused", 1)
    finally:
        hg.set_option("cgs_fused", 2)
        hg.set_option("cgs_step_max_n", 140000)
    (r0, e0), (r1, e1) = out[0], out[2]
    assert r0[3] == r1[3]
    assert np.max(_colwise(e1["H"], e0["H"], r0[3])) < 1e-11
    assert np.max(np.abs(r0[2] - r1[2]) / r0[2]) < 1e-11
    assert np.linalg.norm(r0[0] - r1[0]) / np.linalg.norm(r0[0]) < 1e-11


@pytest.mark.parametrize("solver", ["hybrid_ab_gmres_rtp", "hybrid_ba_gmres_rtp"])
def test_nspace_permutation_is_a_similarity(hg, ctx, ct48_unmatched, solver):
    """Running the n-space in 4x4-tile pixel order (hg_matrix_permute) changes only the summation
    order: same H, beta, histories and iterate as the natural order."""
    from hybrid_gmres_b200.ct import tile_permutation
    A, B, b, x_true = ct48_unmatched
    f = getattr(hg, solver)
    e0, e1 = {}, {}
    x0, err0, res0, k0 = f(A, B, b, x_true, 0.0, 30, 1e-2, ctx=ctx, extras=e0)
    x1, err1, res1, k1 = f(A, B, b, x_true, 0.0, 30, 1e-2, ctx=ctx, extras=e1, nperm=tile_permutation(48, 4))
    assert k0 == k1 == 30
    assert abs(e0["beta"] - e1["beta"]) <= 1e-13 * abs(e0["beta"])
    for j in range(30):
        assert np.linalg.norm(e0["H"][:, j] - e1["H"][:, j]) <= 1e-10 * np.linalg.norm(e0["H"][:, j])
    assert np.max(np.abs(res0 - res1) / res0) < 1e-10
    assert np.max(np.abs(err0 - err1) / err0) < 1e-10
    assert np.linalg.norm(x0 - x1) <= 1e-10 * np.linalg.norm(x0)
    assert np.linalg.norm(e0["X"] - e1["X"]) <= 1e-10 * np.linalg.norm(e0["X"])


def test_cgs_staged_fused_stage_keeps_the_arnoldi_factorisation(hg, ctx):
    """cgs_fused = 2 (default): `w1 = w0 - V h1` and `V' w1` in one pass over the basis, tiles staged in
    shared memory by cp.async (csrc/cgs_staged.cu).  n = 65536 rows >= 256 per SM so the kernel is used, and
    k runs through its three tile shapes (k <= 40, <= 88, <= 208).  On this problem H itself is ill
    determined beyond k ~ 40 (switching the SpMV kernel alone moves column 44 by 3e-2), so the comparison
    with the separate kernels is on the first 38 columns and the full run is judged by what CGS2 must
    deliver: an orthonormal basis and the Arnoldi relation (B A + lambda I) Q_k = Q_{k+1} H, plus
    bit-identical reruns."""
    import scipy.sparse as sp
    from hybrid_gmres_b200.ct import ct_backprojector, ct_projector, shepp_logan
    N, K, lam = 256, 100, 1e-2
    angles = np.arange(180) * 2.0
    dA = ct_projector(N, angles, None, "fan", ctx=ctx)
    dB = ct_backprojector(N, angles, None, "fan", ctx=ctx)
    b = dA.matvec(shepp_logan(N))
    out = {}
    Q = None
    try:
        for fused in (0, 2, 2):
            hg.set_option("cgs_fused", fused)
            ar = hg.Arnoldi(dA, dB, "n", K)
            ar.set_rhs(b)
            ar.reset(lam)
            ar.steps(K)
            out.setdefault(fused, []).append(ar.get())
            if fused == 2 and Q is None:
                Q = np.column_stack([ar.q(j) for j in range(K + 1)])
            ar.close()
    finally:
        hg.set_option("cgs_fused", 2)
    (H0, beta0, k0), (H2, beta2, k2), (H2b, _, _) = out[0][0], out[2][0], out[2][1]
    assert k0 == k2 == K and beta0 == beta2
    assert np.array_equal(H2, H2b)  # deterministic
    for j in range(38):
        assert np.linalg.norm(H2[:j + 2, j] - H0[:j + 2, j]) <= 1e-9 * np.linalg.norm(H0[:j + 2, j]), j
    assert np.max(np.abs(Q.T @ Q - np.eye(K + 1))) < 1e-12
    A = sp.csr_matrix(tuple(reversed(dA.download())), shape=dA.shape)
    B = sp.csr_matrix(tuple(reversed(dB.download())), shape=dB.shape)
    MQ = B @ (A @ Q[:, :K]) + lam * Q[:, :K]
    R = MQ - Q @ H2
    assert np.max(np.linalg.norm(R, axis=0) / np.linalg.norm(MQ, axis=0)) < 1e-12


@pytest.mark.parametrize("b_kind", ["pixel", "perturbed"])
def test_config2_256_ba_rtp_vs_oracle(hg, ctx, b_kind):
    """BASELINE configs[1] at its real size — 256^2 phantom, hybrid BA-GMRES with an unmatched
    back-projector (pixel-driven, and one level of the mismatch sweep B = A' + c E,
    run_2D_phantom.m:79-89) — through the default device path (sliced 16-bit SpMV where it applies,
    n-space in 4x4 tiles, one-pass staged CGS2 middle stage: n = 65536 rows is inside its range)
    against the literal MGS oracle: residual / error histories and every iterate within 1e-8 over the
    first 50 iterations, identical stopping iteration.

    Where the problem itself is ill conditioned the bar is the oracle's own sensitivity: on the
    pixel-driven problem Ritz values converge around k = 23-27 and k = 43-47, where the oracle's residual
    history moves by 4e-5 / 5e-4 when b is perturbed by 1e-15 or MGS is replaced by CGS2, and re-converges
    to 1e-14 / 1e-9 in between and after (every device variant shows the same profile); the bound per
    iteration is max(1e-8, 50 x that measured sensitivity, taken over a window of +-2 iterations)."""
    import oracle
    from oracle import ct
    from hybrid_gmres_b200.ct import tile_permutation
    A, B, b, x_true = ct.make_ct_problem(256, 180, "parallel", b_kind, noise=0.01, mismatch=1e-2)
    maxit, lam, tol = 50, 1e-2, 1e-6
    ex_d, ex_o, ex_p = {}, {}, {}
    x, err, res, it = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam, ctx=ctx, extras=ex_d,
                                             nperm=tile_permutation(256, 4))
    xo, erro, reso, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam, extras=ex_o)
    b_pert = b * (1.0 + 1e-15 * np.random.default_rng(1).standard_normal(b.shape))
    xp, errp, resp, itp = oracle.hybrid_ba_gmres_rtp(A, B, b_pert, x_true, tol, maxit, lam, extras=ex_p)
    assert it == ito == itp
    # the device orthogonalises with CGS2, the oracle (as the reference) with MGS: both the response to a
    # 1e-15 perturbation of b and the MGS-vs-CGS2 difference of the oracle measure how well the histories
    # are determined at iteration k; the sensitive stretch may sit an iteration earlier or later
    xc, errc, resc, itc = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam, orth="cgs2", extras=(ex_c := {}))

    def window(v):
        w = np.copy(v)
        for sh in (1, 2):
            w[sh:] = np.maximum(w[sh:], v[:-sh])
            w[:-sh] = np.maximum(w[:-sh], v[sh:])
        return w

    sens_res = window(np.maximum(np.abs(resp - reso) / reso, np.abs(resc - reso) / reso))
    sens_x = window(np.maximum(_iter_rel(ex_p["X"], ex_o["X"]), _iter_rel(ex_c["X"], ex_o["X"])))
    d_res = np.abs(res - reso) / reso
    d_err = np.abs(err - erro) / erro
    d_x = _iter_rel(ex_d["X"], ex_o["X"])
    for name, d, sens in (("residual", d_res, sens_res), ("iterate", d_x, sens_x), ("error", d_err, sens_x)):
        ratio = d / np.maximum(TOL, 50 * sens)
        k = int(np.argmax(ratio))
        assert ratio[k] <= 1.0, (name, k + 1, d[k], sens[k], [f"{v:.1e}" for v in d[38:]], [f"{v:.1e}" for v in sens[38:]])
    assert np.max(d_res[:20]) < TOL and np.max(d_x[:20]) < TOL  # before the first sensitive stretch: the plain bar
    assert abs(ex_d["beta"] - ex_o["beta"]) / ex_o["beta"] < 1e-13

    # Same algorithm on both sides: the device against the oracle's CGS2 variant, per iteration, with the
    # tightest bound the problem allows — 20 x the CGS2 oracle's own response to the 1e-15 perturbation of b
    # (window of +-2 iterations), floored at the rounding level of the quantity.  Measured on B200: the
    # device/oracle difference is at most 5 x that sensitivity at every k (pixel: residual 7e-16 at k=10,
    # 3e-13 at k=19, 7e-4 at k=26, 2e-14 at k=31; perturbed: <= 2e-10 everywhere), i.e. the device is as close
    # to the oracle as the oracle is to itself.
    xq, errq, resq, itq = oracle.hybrid_ba_gmres_rtp(A, B, b_pert, x_true, tol, maxit, lam, orth="cgs2", extras=(ex_q := {}))
    assert it == itc == itq
    sens_c = {
        "residual": window(np.abs(resq - resc) / resc),
        "iterate": window(_iter_rel(ex_q["X"], ex_c["X"])),
        "hessenberg": window(_colwise(ex_q["H"], ex_c["H"], it)),
    }
    diff_c = {
        "residual": np.abs(res - resc) / resc,
        "iterate": _iter_rel(ex_d["X"], ex_c["X"]),
        "hessenberg": _colwise(ex_d["H"], ex_c["H"], it),
    }
    floor = {"residual": 5e-14, "iterate": 1e-12, "hessenberg": 5e-14}
    for name in diff_c:
        ratio = diff_c[name] / np.maximum(floor[name], 20 * sens_c[name])
        k = int(np.argmax(ratio))
        assert ratio[k] <= 1.0, (name, "cgs2", k + 1, diff_c[name][k], sens_c[name][k])
    assert np.max(diff_c["residual"][:14]) < 1e-13 and np.max(diff_c["hessenberg"][:14]) < 1e-13
    assert np.max(diff_c["iterate"][:16]) < 1e-12


def test_rtp_error_history_modes_agree(hg, ctx, ct48_unmatched):
    """error_mode 2 (||x_k - x_true|| from the orthonormal basis, x formed once; the default for n >= 200000) against error_mode 1
    (x_k and the difference formed at every iteration, hybrid_ba_gmres_rtp.m:30,33 literally) and the oracle."""
    import oracle
    A, B, b, x_true = ct48_unmatched
    for f, fo in ((hg.hybrid_ba_gmres_rtp, oracle.hybrid_ba_gmres_rtp), (hg.hybrid_ab_gmres_rtp, oracle.hybrid_ab_gmres_rtp)):
        x0, e0, r0, it0 = f(A, B, b, x_true, 1e-6, 40, 1e-2, ctx=ctx, error_mode=2)
        x1, e1, r1, it1 = f(A, B, b, x_true, 1e-6, 40, 1e-2, ctx=ctx, error_mode=1)
        xo, eo, ro, ito = fo(A, B, b, x_true, 1e-6, 40, 1e-2)
        assert it0 == it1 == ito
        assert np.max(np.abs(e0 - e1) / e1) < 1e-10 and np.array_equal(r0, r1)
        assert np.max(np.abs(e0 - eo) / eo) < 1e-8
        assert np.linalg.norm(x0 - x1) <= 1e-13 * np.linalg.norm(x1)
        # early stop: the iterate returned is the one of the stopping iteration
        xs, es, rs, its = f(A, B, b, x_true, float(r1[7]) * 1.0000001, 40, 1e-2, ctx=ctx, error_mode=2)
        xso, eso, rso, itso = fo(A, B, b, x_true, float(r1[7]) * 1.0000001, 40, 1e-2)
        assert its == itso == 8 and np.linalg.norm(xs - xso) <= 1e-8 * np.linalg.norm(xso)


def test_rtp_error_history_small_errors_fall_back_to_explicit(hg, ctx):
    """A consistent, well conditioned problem converges to a relative error far below 1 %: there the
    algebraic formula would cancel, so those iterations form x_k explicitly — the history must follow the
    oracle down to 1e-9 relative error with the usual 1e-8 agreement."""
    import oracle
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    m, n = 600, 400
    A = sp.random(m, n, 0.05, random_state=7, format="csr") + sp.vstack([sp.identity(n) * 3.0, sp.csr_matrix((m - n, n))])
    A = A.tocsr()
    B = A.T.tocsr()
    x_true = rng.standard_normal(n)
    b = A @ x_true
    x, err, res, it = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, 60, 1e-12, ctx=ctx, error_mode=2)
    xo, erro, reso, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, 60, 1e-12)
    assert it == ito and erro[-1] < 1e-6  # the regime the guard exists for
    ok = erro > 1e-9
    assert np.max(np.abs(err[ok] - erro[ok]) / erro[ok]) < 1e-6
    big = erro > 1e-6
    assert np.max(np.abs(err[big] - erro[big]) / erro[big]) < 1e-8
    assert np.linalg.norm(x - xo) <= 1e-8 * np.linalg.norm(xo)
