"""One-process-per-GPU sharded Arnoldi (``hg_darnoldi_*``).  ``torch.distributed`` is used
only to broadcast the NCCL unique id; the collectives between the kernels (reduce-scatter,
all-reduce, all-gather) are issued by ``libhgmres.so`` on its own stream."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from ._lib import check
from .api import Context, DeviceMatrix, _ptr, _vec
from .sharding import slice_len


class Communicator:
    """``hg_comm``: an NCCL communicator over all ranks of the default process group."""

    def __init__(self, ctx: Context, rank: int | None = None, world: int | None = None):
        import torch
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank()
        if world is None:
            world = dist.get_world_size()
        self.ctx, self.rank, self.world = ctx, rank, world
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            check(ctx._lib.hg_comm_unique_id(_ptr(uid)))
        if world > 1:
            t = torch.from_numpy(uid)
            if dist.get_backend() == "nccl":
                t = t.cuda()
                dist.broadcast(t, src=0)
                uid = t.cpu().numpy()
            else:
                dist.broadcast(t, src=0)
        h = C.c_void_p()
        check(ctx._lib.hg_comm_init(ctx._h, world, rank, _ptr(np.ascontiguousarray(uid)), C.byref(h)))
        self._h = h

    @property
    def transport(self) -> str:
        """``"peer: ..."`` (collectives inside the kernels over NVLink peer memory) or ``"nccl: ..."`` —
        decided collectively when a sharded Arnoldi is created (``hg_comm_transport``)."""
        t = C.c_int()
        why = C.create_string_buffer(256)
        check(self.ctx._lib.hg_comm_transport(self._h, C.byref(t), why, 256))
        return ("nccl: " if t.value == 0 else "peer: ") + why.value.decode()

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.hg_comm_destroy(self._h)
            self._h = None


class ShardedArnoldi:
    """CGS2 Arnoldi on ``B*A + shift*I`` with ``A`` row-sharded and ``B`` column-sharded."""

    def __init__(self, comm: Communicator, A_p: DeviceMatrix, B_p: DeviceMatrix, kmax: int):
        self.comm, self.ctx = comm, comm.ctx
        self.A_p, self.B_p, self.kmax = A_p, B_p, int(kmax)
        self.n = A_p.shape[1]
        self.n_p = slice_len(self.n, comm.world)
        h = C.c_void_p()
        check(self.ctx._lib.hg_darnoldi_create(self.ctx._h, comm._h, A_p._h, B_p._h, self.kmax, C.byref(h)))
        self._h = h

    def set_rhs(self, b_p):
        b_p = _vec(b_p, self.A_p.shape[0], "b_p")
        check(self.ctx._lib.hg_darnoldi_set_rhs(self._h, _ptr(b_p)))

    def reset(self, shift: float):
        check(self.ctx._lib.hg_darnoldi_reset(self._h, float(shift)))

    def steps(self, n: int):
        check(self.ctx._lib.hg_darnoldi_steps(self._h, int(n)))

    def get(self):
        H = np.zeros((self.kmax + 1, self.kmax), order="F")
        beta, k = C.c_double(), C.c_int()
        check(self.ctx._lib.hg_darnoldi_get(self._h, _ptr(H), self.kmax + 1, C.byref(beta), C.byref(k)))
        return H, beta.value, k.value

    def q_slice(self, j: int):
        q = np.empty(self.n_p)
        r0, nr = C.c_int64(), C.c_int64()
        check(self.ctx._lib.hg_darnoldi_get_q(self._h, int(j), _ptr(q), C.byref(r0), C.byref(nr)))
        return q, int(r0.value)

    def step_bytes(self, k: int) -> float:
        out = C.c_double()
        check(self.ctx._lib.hg_darnoldi_step_bytes(self._h, int(k), C.byref(out)))
        return out.value

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.hg_darnoldi_destroy(self._h)
            self._h = None


def _dist_rtp(kind, comm, A_p, B_p, b_p, x_true, tol, maxit, lam, extras, nperm=None, cache=True, stats=None):
    from ._lib import HgExtras, c_double_p
    from .api import _resident
    ctx = comm.ctx
    maxit = int(maxit)
    n = A_p.shape[1]
    b_p = _vec(b_p, A_p.shape[0], "b_p")
    x_true = _vec(x_true, n, "x_true")
    if nperm is not None:  # n-space order of the solve (api._rtp): A_p(:,q), B^p(q,:), x_true(q)
        nperm = np.ascontiguousarray(nperm, dtype=np.int32)
        x_true = np.ascontiguousarray(x_true[nperm])
    dA, ownA = _resident(ctx, A_p, "A", nperm, cache, stats)
    dB, ownB = _resident(ctx, B_p, "B", nperm, cache, stats)
    x, err, res = np.zeros(n), np.zeros(maxit), np.zeros(maxit)
    niters, x_valid = C.c_int(), C.c_int()
    ex, bufs = None, None
    if extras is not None:
        bufs = {"H": np.zeros((maxit + 1, maxit), order="F"), "beta": np.zeros(1)}
        ex = HgExtras()
        ex.H = bufs["H"].ctypes.data_as(c_double_p)
        ex.beta = bufs["beta"].ctypes.data_as(c_double_p)
    try:
        t0 = time.perf_counter()
        check(ctx._lib.hg_dist_hybrid_rtp(kind, ctx._h, comm._h, dA._h, dB._h, _ptr(b_p), _ptr(x_true), float(tol),
                                          maxit, float(lam), _ptr(x), _ptr(err), _ptr(res), C.byref(niters),
                                          C.byref(x_valid), C.byref(ex) if ex else None))
        if stats is not None:
            stats["solve_ms"] = 1e3 * (time.perf_counter() - t0)
    finally:
        if ownA:
            dA.close()
        if ownB:
            dB.close()
    k = niters.value
    if nperm is not None:
        xu = np.empty_like(x)
        xu[nperm] = x
        x = xu
    if extras is not None:
        extras.update(H=bufs["H"], beta=float(bufs["beta"][0]))
    return (x if x_valid.value else None), err[:k], res[:k], k


def hybrid_ab_gmres_rtp(comm, A_p, B_p, b_p, x_true, tol, maxit, lam, *, extras=None, nperm=None, cache=True,
                        stats=None):
    """Sharded ``hybrid_ab_gmres_rtp.m``: ``A_p``/``B_p`` are this rank's shards (device matrices, or
    host matrices kept resident on the device between calls), ``b_p`` its slice of ``b``; returns the
    reference's four outputs on every rank.  ``nperm``: n-space order on the device, as in the single-GPU
    solver."""
    return _dist_rtp(0, comm, A_p, B_p, b_p, x_true, tol, maxit, lam, extras, nperm, cache, stats)


def hybrid_ba_gmres_rtp(comm, A_p, B_p, b_p, x_true, tol, maxit, lam, *, extras=None, nperm=None, cache=True,
                        stats=None):
    """Sharded ``hybrid_ba_gmres_rtp.m``."""
    return _dist_rtp(1, comm, A_p, B_p, b_p, x_true, tol, maxit, lam, extras, nperm, cache, stats)


def gcv_prepare(comm, A_p, B_p, b_p, m, k_gcv, gcv_type):
    """Sharded ``gcv_function`` Arnoldi (``hg_dist_gcv_prepare``): returns a
    :class:`hybrid_gmres_b200.GcvProblem` whose ``eval`` / ``fminbnd`` run on the host of
    every rank with identical results."""
    from .api import GcvProblem
    ctx = comm.ctx
    b_p = _vec(b_p, A_p.shape[0], "b_p")
    h = C.c_void_p()
    check(ctx._lib.hg_dist_gcv_prepare(ctx._h, comm._h, A_p._h, B_p._h, _ptr(b_p), int(m), int(k_gcv),
                                       {"ab": 0, "ba": 1}[gcv_type], C.byref(h)))
    return GcvProblem(h, ctx._lib)


def _dist_gkb(which, comm, A_p, b_p, x_true, tol, maxit, lam, At_p):
    ctx = comm.ctx
    n = A_p.shape[1]
    maxit = int(maxit)
    b_p = _vec(b_p, A_p.shape[0], "b_p")
    xt = _vec(x_true, n, "x_true") if x_true is not None and np.size(x_true) > 0 else None
    x, err, res, ar = np.zeros(n), np.zeros(maxit), np.zeros(maxit), np.zeros(maxit)
    niters = C.c_int()
    check(ctx._lib.hg_dist_gkb_solver(which, ctx._h, comm._h, A_p._h, At_p._h if At_p is not None else None,
                                      _ptr(b_p), _ptr(xt), float(tol), maxit, float(lam or 0.0), _ptr(x), _ptr(err),
                                      _ptr(res), _ptr(ar), C.byref(niters), None))
    k = niters.value
    if which == 3:
        return x, err[:k], res[:k], ar[:k], k
    return x, err[:k], res[:k], k


def hybrid_lsqr_solver(comm, A_p, b_p, x_true, tol, maxit, lam, *, At_p=None):
    """Sharded ``hybrid_lsqr_solver.m`` (``A_p``: this rank's row block of ``A``)."""
    return _dist_gkb(0, comm, A_p, b_p, x_true, tol, maxit, lam, At_p)


def hybrid_lsmr_solver(comm, A_p, b_p, x_true, tol, maxit, lam, *, At_p=None):
    """Sharded ``hybrid_lsmr_solver.m``."""
    return _dist_gkb(1, comm, A_p, b_p, x_true, tol, maxit, lam, At_p)


def lsqr_solver(comm, A_p, b_p, x_true, tol, maxit, *, At_p=None):
    """Sharded ``lsqr_solver.m``."""
    return _dist_gkb(2, comm, A_p, b_p, x_true, tol, maxit, None, At_p)


def lsmr_solver(comm, A_p, b_p, x_true=None, tol=1e-6, maxit=None, *, At_p=None):
    """Sharded ``lsmr_solver.m`` (five outputs)."""
    if maxit is None:  # lsmr_solver.m:5 — min(m, n) with the GLOBAL row count m = sum of the shards' rows
        m = A_p.shape[0]
        try:
            import torch
            import torch.distributed as dist
            if dist.is_initialized() and comm.world > 1:
                t = torch.tensor([float(m)], dtype=torch.float64,
                                 device="cuda" if dist.get_backend() == "nccl" else "cpu")
                dist.all_reduce(t)
                m = int(t.item())
        except ImportError:
            pass
        maxit = min(m, A_p.shape[1])
    return _dist_gkb(3, comm, A_p, b_p, x_true, tol, maxit, None, At_p)
