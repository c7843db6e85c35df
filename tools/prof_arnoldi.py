# profiling driver: build the headline workload (tile4 n-space order) and run a few Arnoldi steps
import sys, math
import numpy as np
sys.path.insert(0, '.')
import hybrid_gmres_b200 as hg
from hybrid_gmres_b200.ct import tile_permutation
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = hg.Context(0)
angles = np.arange(180) * 2.0
p = int(round(math.sqrt(2.0) * N))
dA0 = hg.ct_projector(N, angles, p, "fan", ctx=ctx)
dB0 = hg.ct_backprojector(N, angles, p, "fan", ctx=ctx)
b = dA0.matvec(hg.ct.shepp_logan(N))
q = tile_permutation(N, 4)
dA, dB = dA0.permute(None, q, sort=False), dB0.permute(q, None)  # as bench.py builds them
dA0.close(); dB0.close()
print(dA.spmv_form, dB.spmv_form)
ar = hg.Arnoldi(dA, dB, "n", steps)
ar.set_rhs(b)
ar.reset(1e-2)
ar.steps(steps)
H, beta, k = ar.get()
print("ok", k, beta, H[1, 0])
