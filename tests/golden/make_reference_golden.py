"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN .m SOURCE.

    python tests/golden/make_reference_golden.py          (needs /root/reference; build container only)

MATLAB / Octave are not in the image, so the untouched files under /root/reference are run by
``oracle/mlab.py`` — an interpreter for the MATLAB subset they use — with NumPy/SciPy standing in for
MATLAB's built-ins only (``*``, ``norm``, ``\\``, ``svd`` ...).  Control flow, indexing, stop rules,
breakdown handling, which variables exist at exit: all of that is the reference text itself.  The
outputs below therefore pin ``oracle/solvers.py`` (the hand restatement) and, through the GPU tests,
the CUDA path, to the reference — up to rounding inside the built-ins.

Per case the file holds, for every hot-path function of SURVEY.md §8a, the function's outputs plus
the internals read out of the reference's own workspace at exit (``H``, ``beta``/``beta1``, ``Q``) and
the iterate after every iteration (a hook on the assignments to ``x``).  Inputs live in the older
fixtures (CT cases) or inside the file (dense n = 32 cases, which are the reference's call-site inputs:
``run_ptr_rtp_comparison.m:4-13``, ``run_equivalence_plots.m:3-11``, ``plot_gcv_surface.m:5-17``).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import generators, mlab  # noqa: E402
from tests.golden_util import load  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HG_REFERENCE_DIR", "/root/reference")

RTP = ("hybrid_ab_gmres_rtp", "hybrid_ba_gmres_rtp")
PTR_H = ("ABgmres_hybrid_bounds", "BAgmres_hybrid_bounds")
PTR_N = ("ABgmres_nonhybrid_bounds", "BAgmres_nonhybrid_bounds")
GKB_H = ("hybrid_lsqr_solver", "hybrid_lsmr_solver")
GCV_LAMS = np.logspace(-9, -1, 9)


def S(v):
    return np.array([[float(v)]])


def col(v):
    return np.asarray(v, dtype=float).reshape(-1, 1)


def toolbox():
    """Hansen's Regularization Tools generators are third-party and not vendored; the
    closed-form restatements of oracle/generators.py stand in for them."""
    def wrap(fn):
        def f(args, nargout):
            A, b, x = fn(int(mlab.scalar(args[0])))
            return A, col(b), col(x)
        return f
    return {"shaw": wrap(generators.shaw), "heat": wrap(generators.heat), "deriv2": wrap(generators.deriv2)}


class Recorder:
    """collects the value of `x` after each assignment inside the named functions"""

    def __init__(self, names):
        self.X = {n: [] for n in names}
        self.hooks = {(n, "x"): (lambda v, ws, n=n: self.X[n].append(np.asarray(v, dtype=float).reshape(-1).copy()))
                      for n in names}

    def take(self, name, drop_first):
        xs = self.X[name][drop_first:]
        self.X[name] = []
        return np.stack(xs, axis=1) if xs else np.zeros((0, 0))


def run_case(A, B, b, x_true, tol, maxit, lam, k_gcv, surface_k=None):
    """every hot-path function of the reference on one input set -> dict of arrays"""
    names = RTP + GKB_H + ("lsqr_solver", "lsmr_solver")
    rec = Recorder(names)
    ip = mlab.Interp([REF], extra=toolbox(), hooks=rec.hooks)
    b, x_true = col(b), col(x_true)
    m, n = A.shape
    out = {}

    def put(key, x, err, res, it):
        out[key + "_x"] = np.asarray(x, dtype=float).reshape(-1)
        out[key + "_err"] = np.asarray(err, dtype=float).reshape(-1)
        out[key + "_res"] = np.asarray(res, dtype=float).reshape(-1)
        out[key + "_it"] = int(mlab.scalar(it))

    for fn in RTP:
        x, err, res, it = ip.call(fn, [A, B, b, x_true, S(tol), S(maxit), S(lam)], 4)
        put(fn, x, err, res, it)
        ws = ip.last_ws[fn]
        out[fn + "_H"] = np.asarray(ws["H"], dtype=float)
        out[fn + "_beta"] = float(mlab.scalar(ws["beta"]))
        out[fn + "_Q"] = np.asarray(ws["Q"], dtype=float)
        # hybrid_ba_gmres_rtp.m:4 assigns x = zeros(n,1) before the loop: drop that one
        out[fn + "_X"] = rec.take(fn, 1 if fn == "hybrid_ba_gmres_rtp" else 0)
    DM = np.zeros((m, m))
    DN = np.zeros((n, n))
    for fn in PTR_H:
        x, err, res, it = ip.call(fn, [A, B, b, x_true, S(tol), S(maxit), S(lam), DM if fn[0] == "A" else DN], 4)
        put(fn, x, err, res, it)
        out[fn + "_H"] = np.asarray(ip.last_ws[fn]["H"], dtype=float)
    for fn in PTR_N:
        x, err, res, it = ip.call(fn, [A, B, b, x_true, S(tol), S(maxit), DM if fn[0] == "A" else DN], 4)
        put(fn, x, err, res, it)
    for fn in GKB_H:
        x, err, res, it = ip.call(fn, [A, b, x_true, S(tol), S(maxit), S(lam)], 4)
        put(fn, x, err, res, it)
        out[fn + "_X"] = rec.take(fn, 1)  # both start with x = zeros(n,1)
    out["hybrid_lsmr_solver_Bk"] = np.asarray(ip.last_ws["hybrid_lsmr_solver"]["B_k"], dtype=float)
    x, err, res, it = ip.call("lsqr_solver", [A, b, x_true, S(tol), S(maxit)], 4)
    put("lsqr_solver", x, err, res, it)
    out["lsqr_solver_X"] = rec.take("lsqr_solver", 1)
    x, err, res, ar, it = ip.call("lsmr_solver", [A, b, x_true, S(tol), S(maxit)], 5)
    put("lsmr_solver", x, err, res, it)
    out["lsmr_solver_ar"] = np.asarray(ar, dtype=float).reshape(-1)
    out["lsmr_solver_X"] = rec.take("lsmr_solver", 1)
    # optional arguments: lsmr_solver(A,b) -> tol 1e-6, maxit min(m,n), NaN error history (:3-5,28,72-74)
    x, err, res, ar, it = ip.call("lsmr_solver", [A, b], 5)
    put("lsmr_defaults", x, err, res, it)
    rec.take("lsmr_solver", 0)
    # gcv_function on a lambda grid, both types, and the fminbnd call of analyze_regularization.m:35-46
    for t in ("ab", "ba"):
        out[f"gcv_{t}_vals"] = np.array([float(mlab.scalar(
            ip.call("gcv_function", [S(l), A, B, b, S(m), S(k_gcv), t], 1)[0])) for l in GCV_LAMS])
        out[f"gcv_{t}_H"] = np.asarray(ip.last_ws["gcv_function"]["H"], dtype=float)
        out[f"gcv_{t}_beta"] = float(mlab.scalar(ip.last_ws["gcv_function"]["beta"]))
    ws = dict(A=A, B_pert=B, b=b, err_norms_ab=np.ones((1, 2)), err_norms_ba=np.ones((1, 2)),
              lambda_range=np.ones((1, 2)))
    ip.run_lines(os.path.join(REF, "analyze_regularization.m"), 35, 52, ws)
    out["gcv_ab_fminbnd_lambda"] = float(mlab.scalar(ws["lambda_gcv_ab"]))
    out["gcv_ba_fminbnd_lambda"] = float(mlab.scalar(ws["lambda_gcv_ba"]))
    out["gcv_lams"] = GCV_LAMS
    # plot_gcv_surface.m:58-102 (local function compute_gcv_surface): GCV surface and lambda_k path
    if surface_k:
        f = ip.local_function("plot_gcv_surface", "compute_gcv_surface")
        lam_grid = np.logspace(-8, -1, 30).reshape(1, -1)
        kr = np.arange(1, surface_k + 1, dtype=float).reshape(1, -1)
        for t in ("ab", "ba"):
            surf, path = ip.call(f, [t, A, B, b, S(surface_k), kr, lam_grid], 2)
            out[f"surface_{t}"] = np.asarray(surf, dtype=float)
            out[f"surface_{t}_path"] = np.asarray(path, dtype=float).reshape(-1)
            # per-iteration lambda_k hybrid solve: the iterate the reference's own PTR
            # solver returns after k iterations with lambda = lambda_k of the path above
            fn = ("AB" if t == "ab" else "BA") + "gmres_hybrid_bounds"
            Xk, ek, rk = [], [], []
            for kk in range(1, surface_k + 1):
                lam_k = float(out[f"surface_{t}_path"][kk - 1])
                xk, e_, r_, it_ = ip.call(fn, [A, B, b, x_true, S(0.0), S(kk), S(lam_k), DM if t == "ab" else DN], 4)
                assert int(mlab.scalar(it_)) == kk
                Xk.append(np.asarray(xk, dtype=float).reshape(-1))
                ek.append(float(np.asarray(e_).reshape(-1)[-1]))
                rk.append(float(np.asarray(r_).reshape(-1)[-1]))
            out[f"lamk_{t}_X"] = np.stack(Xk, axis=1)
            out[f"lamk_{t}_err"] = np.array(ek)
            out[f"lamk_{t}_res"] = np.array(rk)
        out["surface_lams"] = lam_grid.reshape(-1)
        out["surface_k"] = surface_k
    out["ref_files_executed"] = np.array(sorted({os.path.basename(f) for _, f in ip.calls}))
    return out


def breakdown_case():
    """A = B = I, b = e1: H(2,1) == 0 at k = 1 (hybrid_ab_gmres_rtp.m:25).  AB leaves `x` unassigned
    (MATLAB: 'Output argument "x" not assigned'), BA returns its initial zeros; both return
    niters = 1 and the untouched zero first history entries (:41-43 / :38-40)."""
    n = 6
    I = np.eye(n)
    e1 = np.zeros((n, 1))
    e1[0] = 1.0
    xt = np.ones((n, 1))
    ip = mlab.Interp([REF])
    out = {"n": n}
    try:
        ip.call("hybrid_ab_gmres_rtp", [I, I, e1, xt, S(1e-6), S(4), S(1e-2)], 4)
        out["ab_x_assigned"] = True
    except mlab.MlabError as e:
        assert "not assigned" in str(e), e
        out["ab_x_assigned"] = False
    ws = ip.last_ws["hybrid_ab_gmres_rtp"]
    assert ("x" in ws) == out["ab_x_assigned"]
    out["ab_it"] = int(mlab.scalar(ws["niters"]))
    out["ab_res"] = np.asarray(ws["residual_norm"]).reshape(-1)
    out["ab_err"] = np.asarray(ws["error_norm"]).reshape(-1)
    out["ab_H"] = np.asarray(ws["H"])
    x, err, res, it = ip.call("hybrid_ba_gmres_rtp", [I, I, e1, xt, S(1e-6), S(4), S(1e-2)], 4)
    out.update(ba_x=np.asarray(x).reshape(-1), ba_err=np.asarray(err).reshape(-1), ba_res=np.asarray(res).reshape(-1),
               ba_it=int(mlab.scalar(it)))
    return out


def dense_case(name):
    """the reference's own call-site inputs: n = 32, full A, 1 % noise, lambda = 1e-3, maxit = n,
    tol = 1e-6 (run_ptr_rtp_comparison.m:4-13); B = A' + 1e-4 E (plot_gcv_surface.m:14-15)."""
    ip = mlab.Interp([REF], extra=toolbox())
    A, b_exact, x_true = ip.call("generate_test_problem", [name, S(32)], 3)  # generate_test_problem.m:1-12
    rng = np.random.default_rng(0)  # MATLAB's rng(0) randn stream cannot be reproduced without MATLAB
    noise = rng.standard_normal(b_exact.shape)
    b = b_exact + 1e-2 * np.linalg.norm(b_exact) * noise / np.linalg.norm(noise)
    B = A.T + 1e-4 * rng.standard_normal(A.T.shape)
    return np.asarray(A), np.asarray(B), b.reshape(-1), x_true.reshape(-1)


def generate():
    files = {}
    for name in ("ct16_perturbed", "ct20_fan_pixel"):
        A, B, g = load(name)
        d = run_case(A, B, g["b"], g["x_true"], float(g["tol"]), int(g["maxit"]), float(g["lam"]), int(g["k_gcv"]),
                     surface_k=12)
        d["inputs_from"] = name
        files["ref_" + name] = d
    for name in ("deriv2", "shaw", "heat"):
        A, B, b, x_true = dense_case(name)
        d = run_case(A, B, b, x_true, 1e-6, 12, 1e-3, 20, surface_k=12)
        d.update(A_dense=A, B_dense=B, b=b, x_true=x_true, tol=1e-6, maxit=12, lam=1e-3, k_gcv=20)
        files[f"ref_{name}_n32"] = d
    files["ref_breakdown"] = breakdown_case()
    return files


def main():
    for name, d in generate().items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, os.path.getsize(path) // 1024, "KiB", "files:", list(d.get("ref_files_executed", [])))


if __name__ == "__main__":
    main()
