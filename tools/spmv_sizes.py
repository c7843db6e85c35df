#!/usr/bin/env python
"""SpMV kernel time per launch for the matrix shapes of BASELINE configs 2-4 and of one rank's shard at
8 GPUs (A_p: m/8 rows, B^p: n rows x m/8 columns), in the form each matrix runs with by default.

    python tools/spmv_sizes.py [N ...]        (needs a B200; HG_TPR / HG_SPMV / HG_IDX16 select variants)
"""
import math
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_gmres_b200 as hg  # noqa: E402
from hybrid_gmres_b200.ct import tile_permutation  # noqa: E402

REPS = int(os.environ.get("REPS", "40"))


def measure(ctx, name, M, x):
    M.matvec(x)
    ctx.timing_enable(True)
    ctx.timing_reset()
    for _ in range(REPS):
        M.matvec(x)
    t = ctx.timing()["spmv"]
    ctx.timing_enable(False)
    ms = t[0] / t[1]
    print(f"{name:28s} {M.shape[0]:8d} x {M.shape[1]:8d} nnz {M.nnz:11d} {M.spmv_form:7s} idx{M.spmv_index_bits}: "
          f"{ms * 1e3:7.1f} us  {12 * M.nnz / ms / 1e6:6.0f} GB/s @12B/nnz  {t[2] / t[1] / ms / 1e6:6.0f} GB/s stored", flush=True)


def main():
    ctx = hg.Context(0)
    rng = np.random.default_rng(0)
    for N in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
        nv = 180
        angles = np.arange(nv) * 2.0
        p = int(round(math.sqrt(2.0) * N))
        q = tile_permutation(N, 4)
        for P in [int(v) for v in os.environ.get("SHARDS", "1,8").split(",")]:
            mine = np.arange(0, nv, P)
            A0 = hg.ct_projector(N, angles[mine], p, "fan", ctx=ctx)
            B0 = hg.ct_backprojector(N, angles[mine], p, "fan", ctx=ctx)
            A, B = A0.permute(None, q, sort=False), B0.permute(q, None)
            A0.close(), B0.close()
            measure(ctx, f"N={N} P={P} A", A, rng.standard_normal(A.shape[1]))
            measure(ctx, f"N={N} P={P} B", B, rng.standard_normal(B.shape[1]))
            A.close(), B.close()


if __name__ == "__main__":
    main()
