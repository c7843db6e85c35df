#!/usr/bin/env python
"""Throughput of every solver on the hot path on device-resident
matrices of BASELINE.json configs 2-4: iterations/s (host wall clock around the public call, results
read back) and, from a second instrumented call, the per-kernel-class device time and achieved
algorithmic GB/s.  Writes one JSON object per line to stdout (kept under profiles/).

    python tools/measure_paths.py > profiles/r01_paths.jsonl        (needs a B200)
"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_gmres_b200 as hg  # noqa: E402
from hybrid_gmres_b200.ct import ct_backprojector, ct_projector, shepp_logan, tile_permutation  # noqa: E402

LAM = 1e-2


def problem(ctx, N, geometry, nviews=180, tile=True):
    angles = np.arange(nviews) * ((360.0 if geometry == "fan" else 180.0) / nviews)
    p = int(round(math.sqrt(2.0) * N))
    A = ct_projector(N, angles, p, geometry, ctx=ctx)
    B = ct_backprojector(N, angles, p, geometry, ctx=ctx)
    x_true = shepp_logan(N)
    b = A.matvec(x_true)
    rng = np.random.default_rng(0)
    e = rng.standard_normal(b.shape[0])
    b = b + 0.01 * np.linalg.norm(b) * e / np.linalg.norm(e)
    if tile:
        q = tile_permutation(N, 4)
        A2, B2 = A.permute(None, q, sort=False), B.permute(q, None)
        A.close(), B.close()
        A, B, x_true = A2, B2, np.ascontiguousarray(x_true[q])
    return A, B, b, x_true


def run(ctx, name, config, fn, iters_of):
    fn()  # warm-up: lazy SpMV forms, buffer cache
    ctx.sync()
    t0 = time.perf_counter()
    out = fn()
    ctx.sync()
    dt = time.perf_counter() - t0
    ctx.timing_enable(True)
    ctx.timing_reset()
    fn()
    tim = ctx.timing()
    ctx.timing_enable(False)
    it = iters_of(out)
    classes = {k: {"ms": round(v[0], 3), "launches": v[1], "GBps": round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None}
               for k, v in tim.items() if v[1]}
    dev_ms = sum(v[0] for v in tim.values())
    print(json.dumps({"path": name, "config": config, "iterations": it, "seconds": round(dt, 4),
                      "iters_per_s": round(it / dt, 1), "device_ms_instrumented": round(dev_ms, 2),
                      "algorithmic_GBps_over_device_time": round(sum(v[2] for v in tim.values()) / (dev_ms * 1e-3) / 1e9, 1),
                      "per_class": classes}), flush=True)


def main():
    ctx = hg.Context(0)
    # config 2: 256^2, hybrid BA-GMRES with unmatched B, 100 iterations
    A, B, b, xt = problem(ctx, 256, "parallel")
    cfg = "configs[1]: 256^2 parallel beam, pixel-driven (unmatched) B, 100 iterations"
    run(ctx, "hybrid_ba_gmres_rtp", cfg, lambda: hg.hybrid_ba_gmres_rtp(A, B, b, xt, 0.0, 100, LAM, ctx=ctx), lambda o: o[3])
    run(ctx, "hybrid_ab_gmres_rtp", cfg, lambda: hg.hybrid_ab_gmres_rtp(A, B, b, xt, 0.0, 100, LAM, ctx=ctx), lambda o: o[3])
    run(ctx, "BAgmres_hybrid_bounds (PTR solve path)", cfg,
        lambda: hg.BAgmres_hybrid_bounds(A, B, b, xt, 0.0, 100, LAM, ctx=ctx), lambda o: o[3])
    run(ctx, "ABgmres_hybrid_bounds (PTR solve path)", cfg,
        lambda: hg.ABgmres_hybrid_bounds(A, B, b, xt, 0.0, 100, LAM, ctx=ctx), lambda o: o[3])
    A.close(), B.close()
    # config 3: 512^2, LSQR / LSMR family vs AB-GMRES (run_equivalence_plots shape), 100 iterations
    A, B, b, xt = problem(ctx, 512, "parallel")
    At = A.transpose()
    cfg = "configs[2]: 512^2 parallel beam, matched A' for the Golub-Kahan solvers, 100 iterations"
    for nm in ("hybrid_lsqr_solver", "hybrid_lsmr_solver"):
        f = getattr(hg, nm)
        run(ctx, nm, cfg, lambda f=f: f(A, b, xt, 0.0, 100, LAM, ctx=ctx, At=At), lambda o: o[3])
    run(ctx, "lsqr_solver", cfg, lambda: hg.lsqr_solver(A, b, xt, 0.0, 100, ctx=ctx, At=At), lambda o: o[3])
    run(ctx, "lsmr_solver", cfg, lambda: hg.lsmr_solver(A, b, xt, 0.0, 100, ctx=ctx, At=At), lambda o: o[4])
    run(ctx, "hybrid_ab_gmres_rtp (B = A')", cfg, lambda: hg.hybrid_ab_gmres_rtp(A, At, b, xt, 0.0, 100, LAM, ctx=ctx),
        lambda o: o[3])
    print(json.dumps({"note": "SpMV forms", "A": f"{A.spmv_form}/idx{A.spmv_index_bits}",
                      "At": f"{At.spmv_form}/idx{At.spmv_index_bits}"}), flush=True)
    A.close(), B.close(), At.close()
    # config 4: 1024^2 fan beam: gcv_function through fminbnd (k_gcv = 20; analyze_regularization.m:35-46)
    A, B, b, xt = problem(ctx, 1024, "fan")
    cfg = "configs[3]: 1024^2 fan beam, pixel-driven B; fminbnd(gcv_function) on [1e-9,1e-1], TolX 1e-8, k_gcv 20"
    for t in ("ab", "ba"):
        run(ctx, f"fminbnd(gcv_function '{t}')", cfg,
            lambda t=t: hg.fminbnd_gcv(A, B, b, A.shape[0], 20, t, 1e-9, 1e-1, 1e-8, ctx=ctx), lambda o: o[2])


if __name__ == "__main__":
    main()
