import sys, numpy as np
sys.path.insert(0, '.')
import hybrid_gmres_b200 as hg
from hybrid_gmres_b200.ct import ct_backprojector, ct_projector, shepp_logan
ctx = hg.Context(0)
N, K = 256, 100
angles = np.arange(180) * 2.0
dA = ct_projector(N, angles, None, "fan", ctx=ctx)
dB = ct_backprojector(N, angles, None, "fan", ctx=ctx)
b = dA.matvec(shepp_logan(N))
def run(fused, spmv_mode=0):
    hg.set_option("cgs_fused", fused); hg.set_option("spmv_mode", spmv_mode)
    A2 = dA.permute(None, None); B2 = dB.permute(None, None)   # fresh matrices (SpMV form chosen under this mode)
    ar = hg.Arnoldi(A2, B2, "n", K); ar.set_rhs(b); ar.reset(1e-2); ar.steps(K)
    H, beta, k = ar.get(); ar.close(); A2.close(); B2.close()
    return H
H0 = run(0)
for name, H in (("fused staged (2)", run(2)), ("separate, CSR-only SpMV", run(0, 1))):
    e = [np.linalg.norm(H[:j+2, j]-H0[:j+2, j])/np.linalg.norm(H0[:j+2, j]) for j in range(K)]
    print(name, " ".join(f"{j}:{e[j]:.1e}" for j in (5, 20, 30, 38, 40, 42, 44, 50, 60, 80, 99)))
hg.set_option("cgs_fused", 2); hg.set_option("spmv_mode", 0)
