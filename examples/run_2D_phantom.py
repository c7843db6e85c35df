#!/usr/bin/env python
"""run_2D_phantom-shaped driver on the accelerated path.

Follows the flow of the reference's only CT script, ``run_2D_phantom.m``: fan-beam
("fancurved", ``:12``) Shepp-Logan problem, noise ``rng(0)``-style (``:17-20``), sinogram
``reshape(b, num_detectors, num_angles)`` (``:25-26``), the four reconstructions, and the
mismatch sweep ``B = A' + c*E`` with ``c in logspace(-4,0,10)`` (``:79-102``).  Differences,
all forced by the reference's un-vendored / built-in dependencies: the system matrices come
from the device generators instead of ``PRtomo_mismatched``; the solvers are the reference's
own hot-path solvers (``*_bounds`` PTR solve paths, ``hybrid_*_rtp``) instead of MATLAB's
built-in ``gmres``/``lsqr`` on explicit products; ``E`` lives on the sparsity pattern of ``A'``.
Prints tables instead of drawing figures.

    python examples/run_2D_phantom.py [N]        (default N = 64; needs a B200)
"""
import math
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_gmres_b200 as hg  # noqa: E402
from hybrid_gmres_b200.ct import shepp_logan  # noqa: E402


def run(N=64, n_views=180, noise_lvl=0.25, maxit=80, lam=1e-2, tol=1e-6, levels=None, ctx=None, verbose=True,
        k_gcv=20):
    """The script's computation as a function (tests/test_gpu_example.py compares the returned numbers with
    the oracle): returns a dict with the problem in host form, the four reconstructions' error histories
    and the mismatch-sweep table."""
    say = print if verbose else (lambda *a, **k: None)
    ctx = ctx or hg.default_context()
    levels = np.logspace(-4, 0, 10) if levels is None else np.asarray(levels, dtype=float)  # :79
    angles = np.arange(n_views) * (360.0 / n_views)
    p = int(round(math.sqrt(2.0) * N))
    dA = hg.ct_projector(N, angles, p, "fan", ctx=ctx)
    x_true = shepp_logan(N)
    b_exact = dA.matvec(x_true)
    rng = np.random.default_rng(0)
    e = rng.standard_normal(b_exact.shape)
    b_noise = b_exact + e / np.linalg.norm(e) * noise_lvl * np.linalg.norm(b_exact)  # :18-20
    sino = b_noise.reshape(p, n_views, order="F")  # :25-26
    say(f"N={N}: A {dA.shape}, nnz {dA.nnz}; sinogram {sino.shape}, noise {noise_lvl:.0%}")

    dAt = dA.transpose()  # matched back-projector B = A'
    indptr, indices, data = dAt.download()
    out = {"A": sp.csr_matrix(tuple(reversed(dA.download())), shape=dA.shape),
           "At": sp.csr_matrix((data, indices, indptr), shape=dAt.shape), "b": b_noise, "x_true": x_true,
           "sinogram": sino, "recon": {}, "sweep": [], "levels": levels, "E": []}
    say("\nReconstructions with matched B = A' (final relative error, iterations):")
    for name, f, args in (("non-hybrid AB-GMRES", hg.ABgmres_nonhybrid_bounds, ()),
                          ("non-hybrid BA-GMRES", hg.BAgmres_nonhybrid_bounds, ()),
                          ("hybrid AB-GMRES (PTR)", hg.ABgmres_hybrid_bounds, (lam,)),
                          ("hybrid BA-GMRES (PTR)", hg.BAgmres_hybrid_bounds, (lam,)),
                          ("hybrid AB-GMRES (RTP)", hg.hybrid_ab_gmres_rtp, (lam,)),
                          ("hybrid BA-GMRES (RTP)", hg.hybrid_ba_gmres_rtp, (lam,))):
        x, err, res, it = f(dA, dAt, b_noise, x_true, tol, maxit, *args, ctx=ctx)
        out["recon"][name] = (err, it)
        say(f"  {name:24s} err {err[-1]:.4f}  min err {err.min():.4f} at k={int(err.argmin()) + 1:3d}  iters {it}")

    say("\nRobustness to mismatch, B = A' + c*E (final relative error):")
    say("      c        nonhy-AB  nonhy-BA  hybrid-AB  hybrid-BA  GCV lambda(ba)")
    for c in levels:
        E = rng.standard_normal(data.shape[0])
        E = E / np.linalg.norm(E) * c  # :88 (Frobenius norm c, on the pattern of A')
        dB = hg.DeviceMatrix.from_csr(indptr, indices, data + E, dAt.shape, ctx)
        row = []
        for f, args in ((hg.ABgmres_nonhybrid_bounds, ()), (hg.BAgmres_nonhybrid_bounds, ()),
                        (hg.ABgmres_hybrid_bounds, (lam,)), (hg.BAgmres_hybrid_bounds, (lam,))):
            x, err, res, it = f(dA, dB, b_noise, x_true, tol, maxit, *args, ctx=ctx)
            row.append(err[-1])  # :97-100
        lam_gcv, fval, _ = hg.fminbnd_gcv(dA, dB, b_noise, dA.shape[0], k_gcv, "ba", 1e-9, 1e-1, 1e-8, ctx=ctx)
        say(f"  {c:9.2e}   {row[0]:8.4f}  {row[1]:8.4f}  {row[2]:9.4f}  {row[3]:9.4f}  {lam_gcv:.3e}")
        out["sweep"].append(row + [lam_gcv, fval])
        out["E"].append(E)
        dB.close()
    out["sweep"] = np.array(out["sweep"])
    return out


def main():
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 64)


if __name__ == "__main__":
    main()
