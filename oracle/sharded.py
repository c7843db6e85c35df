"""P-way sharded restatement of the n-space Arnoldi (TEST INFRASTRUCTURE — see
oracle/__init__.py): the same arithmetic as ``oracle.solvers.arnoldi`` with the data
partitioned as SURVEY.md §8(e) prescribes and every exchange made explicit through a small
collective interface, so the multi-GPU semantics can be pinned on CPUs (``gloo``)."""
from __future__ import annotations

import numpy as np


class LocalCollectives:
    """All ranks in one process: ``parts`` are lists indexed by rank."""

    @staticmethod
    def reduce_scatter(parts, n_p):
        total = np.sum(parts, axis=0)
        return [total[p * n_p:(p + 1) * n_p].copy() for p in range(len(parts))]

    @staticmethod
    def all_reduce(parts):
        total = np.sum(parts, axis=0)
        return [total.copy() for _ in parts]

    @staticmethod
    def all_gather(parts):
        full = np.concatenate(parts)
        return [full.copy() for _ in parts]


def sharded_arnoldi_local(A_blocks, B_blocks, b_blocks, n, n_p, shift, kmax):
    """Run the sharded CGS2 Arnoldi for all ranks inside one process.
    Returns ``(H, beta, Q_slices)``; rank p's basis slice is ``Q_slices[p]``."""
    P = len(A_blocks)
    n_pad = n_p * P
    col = LocalCollectives
    pad = lambda v: np.concatenate([v, np.zeros(n_pad - n)])
    w = col.reduce_scatter([pad(np.asarray(B_blocks[p] @ b_blocks[p]).ravel()) for p in range(P)], n_p)
    s = col.all_reduce([np.array([w[p] @ w[p]]) for p in range(P)])[0][0]
    beta = np.sqrt(s)
    Q = [np.zeros((n_p, kmax + 1)) for _ in range(P)]
    for p in range(P):
        Q[p][:, 0] = w[p] / beta
    q_full = col.all_gather([Q[p][:, 0] for p in range(P)])[0]
    H = np.zeros((kmax + 1, kmax))
    for k in range(1, kmax + 1):
        u = [np.asarray(A_blocks[p] @ q_full[:n]).ravel() for p in range(P)]
        w = col.reduce_scatter([pad(np.asarray(B_blocks[p] @ u[p]).ravel()) for p in range(P)], n_p)
        w = [w[p] + shift * Q[p][:, k - 1] for p in range(P)]
        h1 = col.all_reduce([Q[p][:, :k].T @ w[p] for p in range(P)])[0]
        w = [w[p] - Q[p][:, :k] @ h1 for p in range(P)]
        h2 = col.all_reduce([Q[p][:, :k].T @ w[p] for p in range(P)])[0]
        w = [w[p] - Q[p][:, :k] @ h2 for p in range(P)]
        H[:k, k - 1] = h1 + h2
        s = col.all_reduce([np.array([w[p] @ w[p]]) for p in range(P)])[0][0]
        H[k, k - 1] = np.sqrt(s)
        for p in range(P):
            Q[p][:, k] = w[p] / H[k, k - 1]
        q_full = col.all_gather([Q[p][:, k] for p in range(P)])[0]
    return H, beta, Q


def sharded_arnoldi_rank(A_p, B_p, b_p, n, n_p, shift, kmax, dist, torch):
    """One rank of the same algorithm with real collectives (``torch.distributed``,
    any backend): what each GPU process does, in NumPy."""
    P = dist.get_world_size()
    n_pad = n_p * P

    def reduce_scatter(v):
        t = torch.from_numpy(np.concatenate([v, np.zeros(n_pad - v.shape[0])]))
        dist.all_reduce(t)  # gloo has no reduce_scatter for CPU tensors: all-reduce + slice
        r = dist.get_rank()
        return t.numpy()[r * n_p:(r + 1) * n_p].copy()

    def all_reduce(v):
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    def all_gather(v):
        out = [torch.zeros(n_p, dtype=torch.float64) for _ in range(P)]
        dist.all_gather(out, torch.from_numpy(np.ascontiguousarray(v)))
        return np.concatenate([o.numpy() for o in out])

    w = reduce_scatter(np.asarray(B_p @ b_p).ravel())
    beta = float(np.sqrt(all_reduce(np.array([w @ w]))[0]))
    Q = np.zeros((n_p, kmax + 1))
    Q[:, 0] = w / beta
    q_full = all_gather(Q[:, 0])
    H = np.zeros((kmax + 1, kmax))
    for k in range(1, kmax + 1):
        u = np.asarray(A_p @ q_full[:n]).ravel()
        w = reduce_scatter(np.asarray(B_p @ u).ravel()) + shift * Q[:, k - 1]
        h1 = all_reduce(Q[:, :k].T @ w)
        w = w - Q[:, :k] @ h1
        h2 = all_reduce(Q[:, :k].T @ w)
        w = w - Q[:, :k] @ h2
        H[:k, k - 1] = h1 + h2
        H[k, k - 1] = np.sqrt(all_reduce(np.array([w @ w]))[0])
        Q[:, k] = w / H[k, k - 1]
        q_full = all_gather(Q[:, k])
    return H, beta, Q
