"""Worker of tests/test_sharding.py::test_gloo_world_size_2 (launched by torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid_gmres_b200 import sharding  # noqa: E402
from oracle import ct  # noqa: E402
from oracle.sharded import sharded_arnoldi_local, sharded_arnoldi_rank  # noqa: E402

dist.init_process_group("gloo")
rank, P = dist.get_rank(), dist.get_world_size()
A, B, b, x_true = ct.make_ct_problem(16, 24, "parallel", "perturbed")
n = A.shape[1]
n_p = sharding.slice_len(n, P)
A_p, B_p, (lo, hi) = sharding.shard_host_matrices(A, B, P, rank)
H, beta, Q = sharded_arnoldi_rank(A_p, B_p, b[lo:hi], n, n_p, 1e-2, 12, dist, torch)
blocks = [sharding.shard_host_matrices(A, B, P, r) for r in range(P)]
Hl, betal, Ql = sharded_arnoldi_local([x[0] for x in blocks], [x[1] for x in blocks],
                                      [b[x[2][0]:x[2][1]] for x in blocks], n, n_p, 1e-2, 12)
assert abs(beta - betal) / betal < 1e-13
assert np.linalg.norm(H - Hl) / np.linalg.norm(Hl) < 1e-11
assert np.linalg.norm(Q - Ql[rank]) < 1e-10
print("GLOO_OK", rank, flush=True)
dist.destroy_process_group()
