"""Generates tests/golden/*.npz from the oracle (the reference cannot run here: no MATLAB /
Octave — SURVEY.md §8c — so these fixtures pin the ORACLE, and through it the CUDA path;
they are not outputs of the reference itself: "parity unpinned").

    python tests/golden/make_golden.py

Each file holds the inputs in MATLAB-replayable form (CSC arrays of A and B, b, x_true,
scalars) and the oracle outputs of every hot-path function, so oracle/replay.m can run the
untouched reference on the same inputs under MATLAB/Octave and compare."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import ct  # noqa: E402
from oracle.generators import add_noise  # noqa: E402
from oracle.solvers import gcv_arnoldi  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pack_matrix(prefix, M, out):
    if sp.issparse(M):
        C = M.tocsc()
        C.sort_indices()
        out[prefix + "_jc"] = C.indptr.astype(np.int64)
        out[prefix + "_ir"] = C.indices.astype(np.int64)
        out[prefix + "_pr"] = C.data
        out[prefix + "_shape"] = np.array(M.shape, dtype=np.int64)
    else:
        out[prefix + "_dense"] = np.asarray(M)


def run_all(A, B, b, x_true, tol, maxit, lam, k_gcv):
    out = {}
    for name, f in (("ab_rtp", oracle.hybrid_ab_gmres_rtp), ("ba_rtp", oracle.hybrid_ba_gmres_rtp)):
        ex = {}
        x, err, res, it = f(A, B, b, x_true, tol, maxit, lam, extras=ex)
        out.update({f"{name}_x": x, f"{name}_err": err, f"{name}_res": res, f"{name}_it": it,
                    f"{name}_H": ex["H"], f"{name}_beta": ex["beta"], f"{name}_X": ex["X"]})
    for name, f in (("hybrid_lsqr", oracle.hybrid_lsqr_solver), ("hybrid_lsmr", oracle.hybrid_lsmr_solver)):
        ex = {}
        x, err, res, it = f(A, b, x_true, tol, maxit, lam, extras=ex)
        out.update({f"{name}_x": x, f"{name}_err": err, f"{name}_res": res, f"{name}_it": it, f"{name}_X": ex["X"]})
    ex = {}
    x, err, res, it = oracle.lsqr_solver(A, b, x_true, tol, maxit, extras=ex)
    out.update(lsqr_x=x, lsqr_err=err, lsqr_res=res, lsqr_it=it, lsqr_X=ex["X"])
    ex = {}
    x, err, res, ar, it = oracle.lsmr_solver(A, b, x_true, tol, maxit, extras=ex)
    out.update(lsmr_x=x, lsmr_err=err, lsmr_res=res, lsmr_ar=ar, lsmr_it=it, lsmr_X=ex["X"])
    lams = np.logspace(-9, -1, 9)
    for t in ("ab", "ba"):
        H, beta = gcv_arnoldi(A, B, b, A.shape[0], k_gcv, t)
        out[f"gcv_{t}_H"] = H
        out[f"gcv_{t}_beta"] = beta
        out[f"gcv_{t}_vals"] = np.array([oracle.gcv_function(l, A, B, b, A.shape[0], k_gcv, t) for l in lams])
        lam_opt, fval, flag, cnt = oracle.fminbnd(lambda l: oracle.gcv_function(l, A, B, b, A.shape[0], k_gcv, t),
                                                  1e-9, 1e-1, 1e-8)
        out[f"gcv_{t}_fminbnd"] = np.array([lam_opt, fval, cnt])
    out["gcv_lams"] = lams
    return out


def main():
    # 1) the reference's own input shape: deriv2 n=32, full A, B=A', 1% noise, lambda=1e-3
    #    (run_ptr_rtp_comparison.m:4-13, run_equivalence_plots.m:3-11)
    A, b_exact, x_true = oracle.generate_test_problem("deriv2", 32)
    B = A.T.copy()
    b = add_noise(b_exact, 1e-2, 0)
    d = dict(b=b, x_true=x_true, tol=1e-6, maxit=12, lam=1e-3, k_gcv=20)
    pack_matrix("A", A, d)
    pack_matrix("B", B, d)
    d.update(run_all(A, B, b, x_true, 1e-6, 12, 1e-3, 20))
    np.savez_compressed(os.path.join(HERE, "deriv2_n32.npz"), **d)
    # 2) small CT problem with an unmatched (perturbed) back-projector, config-2 shape
    A, B, b, x_true = ct.make_ct_problem(16, 24, "parallel", "perturbed", noise=0.01, mismatch=1e-2)
    d = dict(b=b, x_true=x_true, tol=1e-6, maxit=25, lam=1e-2, k_gcv=20)
    pack_matrix("A", A, d)
    pack_matrix("B", B, d)
    d.update(run_all(A, B, b, x_true, 1e-6, 25, 1e-2, 20))
    np.savez_compressed(os.path.join(HERE, "ct16_perturbed.npz"), **d)
    # 3) fan beam with the structurally unmatched pixel-driven B, config-4 shape in miniature
    A, B, b, x_true = ct.make_ct_problem(20, 30, "fan", "pixel", noise=0.01)
    d = dict(b=b, x_true=x_true, tol=1e-6, maxit=25, lam=1e-2, k_gcv=20)
    pack_matrix("A", A, d)
    pack_matrix("B", B, d)
    d.update(run_all(A, B, b, x_true, 1e-6, 25, 1e-2, 20))
    np.savez_compressed(os.path.join(HERE, "ct20_fan_pixel.npz"), **d)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
