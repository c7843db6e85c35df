#!/usr/bin/env python
"""Whole-step CGS2 kernel (csrc/cgs2_step.cu) vs the separate kernels: Arnoldi cycle time and per-class device
time for several vector lengths on one GPU.   python tools/cgs_step_compare.py [N ...]"""
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_gmres_b200 as hg  # noqa: E402
from hybrid_gmres_b200.ct import shepp_logan, tile_permutation  # noqa: E402
import torch  # noqa: E402

K = int(os.environ.get("K", "200"))


def main():
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx = hg.Context(0, stream=st.cuda_stream)
    for N in [int(a) for a in sys.argv[1:]] or [256, 362, 512, 724, 1024]:
        nv = 180
        angles = np.arange(nv) * 2.0
        p = int(round(math.sqrt(2.0) * N))
        A0 = hg.ct_projector(N, angles, p, "fan", ctx=ctx)
        B0 = hg.ct_backprojector(N, angles, p, "fan", ctx=ctx)
        if N % 4 == 0:
            q = tile_permutation(N, 4)
            A, B = A0.permute(None, q, sort=False), B0.permute(q, None)
            A0.close(), B0.close()
        else:
            A, B = A0, B0
        b = A.matvec(np.ones(A.shape[1]))
        for mode in (0, 4000000):
            hg.set_option("cgs_step_max_n", mode)
            ar = hg.Arnoldi(A, B, "n", K)
            ar.set_rhs(b)
            for _ in range(2):
                ar.reset(1e-2)
                ar.steps(K)
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ar.reset(1e-2)
                ar.steps(K)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            ctx.timing_enable(True)
            ctx.timing_reset()
            ar.reset(1e-2)
            ar.steps(K)
            tim = ctx.timing()
            ctx.timing_enable(False)
            ar.close()
            cls = {k: round(v[0], 2) for k, v in tim.items() if v[1]}
            non_spmv = sum(v for k, v in cls.items() if k != "spmv")
            print(json.dumps({"N": N, "n": A.shape[1], "K": K, "whole_step_kernel": bool(mode), "cycle_ms": round(ms, 2),
                              "us_per_step": round(1e3 * ms / K, 1), "class_ms": cls, "non_spmv_ms": round(non_spmv, 2)}), flush=True)
        hg.set_option("cgs_step_max_n", 140000)
        A.close(), B.close()


if __name__ == "__main__":
    main()
