# per-matrix SpMV kernel time on the headline workload: 16- vs 32-bit column indices
import sys, math, numpy as np
sys.path.insert(0, '.')
import hybrid_gmres_b200 as hg
from hybrid_gmres_b200.ct import tile_permutation
N = 1024
ctx = hg.Context(0)
angles = np.arange(180) * 2.0
p = int(round(math.sqrt(2.0) * N))
q = tile_permutation(N, 4)
A0 = hg.ct_projector(N, angles, p, "fan", ctx=ctx)
B0 = hg.ct_backprojector(N, angles, p, "fan", ctx=ctx)
rng = np.random.default_rng(0)
xa, xb = rng.standard_normal(A0.shape[1]), rng.standard_normal(B0.shape[1])
for idx16 in (1, 0, 1, 0):
    hg.set_option("spmv_idx16", idx16)
    for name, M0, x, perm in (("A", A0, xa, (None, q)), ("B", B0, xb, (q, None))):
        M = M0.permute(*perm, sort=False) if name == "A" else M0.permute(*perm)
        M.matvec(x)
        ctx.timing_enable(True); ctx.timing_reset()
        for _ in range(30): M.matvec(x)
        t = ctx.timing()["spmv"]; ctx.timing_enable(False)
        ms = t[0] / t[1]
        print(f"{name} {M.spmv_form:7s} idx{M.spmv_index_bits}: {ms*1e3:7.1f} us/launch, {M.nnz/ms/1e6:6.1f} G nnz/s, "
              f"{t[2]/t[1]/ms/1e6:6.0f} GB/s of its own bytes, {12*M.nnz/ms/1e6:6.0f} GB/s at 12 B/nnz", flush=True)
        M.close()
