"""Multi-GPU worker (torchrun, one process per GPU): sharded device Arnoldi vs the oracle.
Also exercises the sharded device CT generators against the host shards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hybrid_gmres_b200 as hg  # noqa: E402
import oracle  # noqa: E402
from hybrid_gmres_b200 import sharding  # noqa: E402
from hybrid_gmres_b200.distributed import Communicator, ShardedArnoldi  # noqa: E402
from oracle import ct  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, P = dist.get_rank(), dist.get_world_size()
ctx = hg.Context(local_rank)
comm = Communicator(ctx)

N, nv, K, lam = 40, 60, 30, 1e-2
A, B, b, x_true = ct.make_ct_problem(N, nv, "fan", "pixel")
n = A.shape[1]
A_p, B_p, (lo, hi) = sharding.shard_host_matrices(A, B, P, rank)
dA, dB = hg.DeviceMatrix.from_any(A_p, ctx), hg.DeviceMatrix.from_any(B_p, ctx)
op = lambda v: np.asarray(B @ (A @ v)).ravel() + lam * v
Qo, Ho, betao, _ = oracle.arnoldi(op, np.asarray(B @ b).ravel(), K, orth="cgs2")
H_by_transport = {}
worst = 0.0
# dist_transport 1: NCCL collectives between the kernels; 2: NVLink peer memory inside the kernels
# (required, no fall-back); then 0 = auto for the rest of the worker
for mode in (1, 2, 0):
    hg.set_option("dist_transport", mode)
    ar = ShardedArnoldi(comm, dA, dB, K)
    assert comm.transport.startswith("nccl" if mode == 1 else "peer"), (mode, comm.transport)
    for rep in range(2):  # a second cycle re-uses the workspace, flags and inbox slots
        ar.set_rhs(b[lo:hi])
        ar.reset(lam)
        ar.steps(K // 2)
        ar.steps(K - K // 2)
        H, beta, k = ar.get()
        assert k == K
        assert abs(beta - betao) / betao < 1e-13, (beta, betao)
        w = max(np.linalg.norm(H[:j + 2, j] - Ho[:j + 2, j]) / np.linalg.norm(Ho[:j + 2, j]) for j in range(K))
        assert w < 1e-10, (mode, w)
        worst = max(worst, w)
        if rep == 0:
            H_by_transport[mode] = H.copy()
        else:
            assert np.array_equal(H, H_by_transport[mode])  # deterministic
    q, r0 = ar.q_slice(K)
    seg = Qo[r0:min(r0 + ar.n_p, n), K]
    assert np.linalg.norm(q[:seg.shape[0]] - seg) < 1e-9 and np.all(q[seg.shape[0]:] == 0)
    # all ranks hold bit-identical H (replicated host projected problem relies on it)
    t = torch.from_numpy(H.copy()).cuda()
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    assert torch.equal(t, t0), mode
    ar.close()
# a deeper Krylov space on a larger slice: every tile shape of the whole-step kernel (k <= 16 / 40 / 96 / 208,
# csrc/cgs2_step.cu) with its in-kernel collectives, against the separate-kernel peer path and the oracle
N2, nv2, K2 = 64, 48, 110
A2, B2, b2, xt2 = ct.make_ct_problem(N2, nv2, "fan", "pixel")
A2_p, B2_p, (lo2, hi2) = sharding.shard_host_matrices(A2, B2, P, rank)
dA2, dB2 = hg.DeviceMatrix.from_any(A2_p, ctx), hg.DeviceMatrix.from_any(B2_p, ctx)
op2 = lambda v: np.asarray(B2 @ (A2 @ v)).ravel() + lam * v
Qo2, Ho2, betao2, _ = oracle.arnoldi(op2, np.asarray(B2 @ b2).ravel(), K2, orth="cgs2")
hg.set_option("dist_transport", 2)
H2 = {}
for max_n in (4000000, 0, 4000000):
    hg.set_option("cgs_step_max_n_dist", max_n)
    ar2 = ShardedArnoldi(comm, dA2, dB2, K2)
    ar2.set_rhs(b2[lo2:hi2])
    ar2.reset(lam)
    l0 = ctx.launch_count
    ar2.steps(K2)
    Hk, betak, kk = ar2.get()
    launches = ctx.launch_count - l0
    assert kk == K2 and abs(betak - betao2) / betao2 < 1e-13
    if max_n in H2:
        assert np.array_equal(Hk, H2[max_n])  # deterministic
    else:
        H2[max_n] = Hk.copy()
        H2[("launches", max_n)] = launches
    for j in range(20):  # the well-determined columns against the oracle (beyond ~30 the problem itself is not)
        assert np.linalg.norm(Hk[:j + 2, j] - Ho2[:j + 2, j]) <= 1e-10 * np.linalg.norm(Ho2[:j + 2, j]), (max_n, j)
    Qs = np.column_stack([ar2.q_slice(j)[0] for j in range(K2 + 1)])
    G2 = torch.from_numpy(Qs.T @ Qs).cuda()
    dist.all_reduce(G2)
    assert float((G2.cpu() - torch.eye(K2 + 1, dtype=torch.float64)).abs().max()) < 1e-12, max_n
    t = torch.from_numpy(Hk.copy()).cuda()
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    assert torch.equal(t, t0), max_n  # bit-identical H on every rank
    ar2.close()
hg.set_option("cgs_step_max_n_dist", 0)
hg.set_option("dist_transport", 0)
assert H2[("launches", 4000000)] <= 4 * K2 + 8 < H2[("launches", 0)], H2[("launches", 4000000)]
for j in range(20):
    assert np.linalg.norm(H2[4000000][:j + 2, j] - H2[0][:j + 2, j]) <= 1e-10 * np.linalg.norm(H2[0][:j + 2, j]), j
dA2.close()
dB2.close()

dH = np.linalg.norm(H_by_transport[1] - H_by_transport[2]) / np.linalg.norm(H_by_transport[1])
assert dH < 1e-10, dH  # different summation orders (whole-step kernel vs separate kernels + NCCL)
assert np.array_equal(H_by_transport[0], H_by_transport[2])
print("TRANSPORT", rank, comm.transport, "|H_nccl - H_peer|/|H| =", dH, flush=True)

# sharded device generators == host shards
angles = np.arange(nv) * (360.0 / nv)
p = int(round(np.sqrt(2.0) * N))
import ctypes as C  # noqa: E402
from hybrid_gmres_b200.ct import ct_backprojector_cols, ct_projector_rows  # noqa: E402
gA = ct_projector_rows(N, angles, p, "fan", lo, hi, ctx=ctx)
ip, ix, dv = gA.download()
assert np.array_equal(ip, A_p.indptr) and np.array_equal(ix, A_p.indices) and np.array_equal(dv, A_p.data)
gB = ct_backprojector_cols(N, angles, p, "fan", lo, hi, ctx=ctx)
ip, ix, dv = gB.download()
Bs = B_p.copy()
Bs.sort_indices()
assert np.array_equal(ip, Bs.indptr) and np.array_equal(ix, Bs.indices) and np.allclose(dv, Bs.data, rtol=1e-12, atol=1e-12)
# sharded gcv_function (m-space 'ab' and n-space 'ba') vs the oracle
from oracle.solvers import gcv_arnoldi  # noqa: E402
from hybrid_gmres_b200 import distributed as hgd0  # noqa: E402
for t in ("ab", "ba"):
    prob = hgd0.gcv_prepare(comm, dA, dB, b[lo:hi], A.shape[0], 20, t)
    Hg, bg = prob.get(20)
    Hgo, bgo = gcv_arnoldi(A, B, b, A.shape[0], 20, t, orth="cgs2")
    assert abs(bg - bgo) / bgo < 1e-13
    wg = max(np.linalg.norm(Hg[:j + 2, j] - Hgo[:j + 2, j]) / np.linalg.norm(Hgo[:j + 2, j]) for j in range(20))
    assert wg < 1e-9, (t, wg)
    for l in (1e-6, 1e-3):
        v, vo = prob.eval(l), oracle.gcv_function(l, A, B, b, A.shape[0], 20, t)
        assert abs(v - vo) / vo < 1e-8, (t, l, v, vo)

# sharded whole solvers vs the oracle
from hybrid_gmres_b200 import distributed as hgd  # noqa: E402
for f_dev, f_orc in ((hgd.hybrid_ba_gmres_rtp, oracle.hybrid_ba_gmres_rtp),
                     (hgd.hybrid_ab_gmres_rtp, oracle.hybrid_ab_gmres_rtp)):
    for tol in (1e-6, 0.05):
        xs, errs, ress, its = f_dev(comm, dA, dB, b[lo:hi], x_true, tol, K, lam)
        xo, erro, reso, ito = f_orc(A, B, b, x_true, tol, K, lam)
        assert its == ito, (its, ito)
        assert np.max(np.abs(ress - reso) / reso) < 1e-8
        assert np.max(np.abs(errs - erro) / erro) < 1e-8
        assert np.linalg.norm(xs - xo) / np.linalg.norm(xo) < 1e-8
# sharded Golub-Kahan solvers vs the oracle (first 8 iterations: see tests/test_gpu_solvers.py)
for name in ("hybrid_lsqr_solver", "hybrid_lsmr_solver", "lsqr_solver", "lsmr_solver"):
    f_dev, f_orc = getattr(hgd, name), getattr(oracle, name)
    if name.startswith("hybrid"):
        out_d = f_dev(comm, dA, b[lo:hi], x_true, 1e-6, 8, lam)
        out_o = f_orc(A, b, x_true, 1e-6, 8, lam)
    else:
        out_d = f_dev(comm, dA, b[lo:hi], x_true, 1e-6, 8)
        out_o = f_orc(A, b, x_true, 1e-6, 8)
    assert out_d[-1] == out_o[-1], name
    for hd, ho in zip(out_d[1:-1], out_o[1:-1]):
        assert np.max(np.abs(hd - ho) / np.abs(ho)) < 1e-8, (name, np.max(np.abs(hd - ho) / np.abs(ho)))
    assert np.linalg.norm(out_d[0] - out_o[0]) / np.linalg.norm(out_o[0]) < 1e-8, name
print("NCCL_OK", rank, worst, flush=True)
ar.close()
comm.close()
dist.destroy_process_group()
