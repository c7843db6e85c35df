"""CPU tests of the multi-GPU host logic: the partitioner, the sharded oracle against the
unsharded one, and a real world_size-2 ``gloo`` run of the per-rank algorithm."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from hybrid_gmres_b200 import sharding
from oracle import ct
from oracle.sharded import sharded_arnoldi_local

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_balanced_row_blocks():
    A, B, b, x_true = ct.make_ct_problem(24, 36, "fan", "pixel")
    for P in (1, 2, 3, 8):
        bounds = sharding.balanced_row_blocks(A.indptr, P)
        assert bounds[0] == 0 and bounds[-1] == A.shape[0] and np.all(np.diff(bounds) >= 0)
        nnz = np.diff(A.indptr[bounds])
        assert nnz.sum() == A.nnz
        assert nnz.max() - nnz.min() <= 2 * np.diff(A.indptr).max()  # within a row of perfect balance
    # degenerate inputs: more ranks than rows, empty matrix
    assert list(sharding.balanced_row_blocks(np.array([0, 3, 3, 5]), 8))[-1] == 3
    assert list(sharding.balanced_row_blocks(np.array([0, 0, 0]), 2)) == [0, 0, 2]
    assert list(sharding.uniform_row_blocks(10, 4)) == [0, 2, 5, 7, 10]
    assert sharding.slice_len(100, 8) == 32 and sharding.slice_len(1048576, 8) == 131072
    assert sharding.slice_len(1, 1) == 32


def test_shards_reassemble():
    A, B, b, x_true = ct.make_ct_problem(16, 24, "parallel", "perturbed")
    P = 3
    u = np.random.default_rng(0).standard_normal(A.shape[1])
    parts, back = [], np.zeros(A.shape[1])
    for r in range(P):
        A_p, B_p, (lo, hi) = sharding.shard_host_matrices(A, B, P, r)
        assert A_p.shape == (hi - lo, A.shape[1]) and B_p.shape == (A.shape[1], hi - lo)
        parts.append(A_p @ u)
        back += B_p @ (A_p @ u)
    assert np.allclose(np.concatenate(parts), A @ u, rtol=1e-14)
    assert np.allclose(back, B @ (A @ u), rtol=1e-12)


@pytest.mark.parametrize("P", [1, 2, 4])
def test_sharded_oracle_equals_unsharded(P):
    A, B, b, x_true = ct.make_ct_problem(20, 30, "fan", "pixel")
    n = A.shape[1]
    n_p = sharding.slice_len(n, P)
    blocks = [sharding.shard_host_matrices(A, B, P, r) for r in range(P)]
    H, beta, Q = sharded_arnoldi_local([x[0] for x in blocks], [x[1] for x in blocks],
                                       [b[x[2][0]:x[2][1]] for x in blocks], n, n_p, 1e-2, 15)
    op = lambda v: np.asarray(B @ (A @ v)).ravel() + 1e-2 * v
    Qo, Ho, betao, _ = oracle.arnoldi(op, np.asarray(B @ b).ravel(), 15, orth="cgs2")
    assert abs(beta - betao) / betao < 1e-13
    for k in range(15):
        assert np.linalg.norm(H[:k + 2, k] - Ho[:k + 2, k]) / np.linalg.norm(Ho[:k + 2, k]) < 1e-10
    Qfull = np.concatenate(Q, axis=0)[:n]
    assert np.linalg.norm(Qfull - Qo) < 1e-9


def test_gloo_world_size_2():
    """Two real processes, gloo backend: the per-rank algorithm with torch.distributed
    collectives reproduces the single-process sharded oracle."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script],
                       capture_output=True, text=True, env=env, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("GLOO_OK") == 2
