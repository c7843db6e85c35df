"""Synthetic CT test problems (TEST INFRASTRUCTURE — see oracle/__init__.py).

The reference's only CT script obtains its matrices from ``PRtomo_mismatched``
(``run_2D_phantom.m:12-15``), a privately modified IR-Tools generator that is
NOT in the reference tree and not public.  SURVEY.md §8(d) therefore defines
the synthetic inputs: a line-intersection projector ``A`` (rows angle-major /
detector-minor to match ``reshape(b, p, n_angles)`` at ``run_2D_phantom.m:25``,
image vectorised column-major as ``reshape(x, N, N)`` at ``:57``), and three
kinds of back-projector ``B``.  PARITY UNPINNED against ``PRtomo_mismatched``.

The ray tracer uses only IEEE add/sub/mul/div/floor/min/max on per-view and
per-ray trig tables computed on the host, in a fixed operation order with no
fused multiply-add — the device generator (``hybrid_gmres_b200/csrc/hg_ct.cu``)
repeats exactly this arithmetic, so the two are compared bit for bit in
``tests/test_ct_generator.py``.

Geometry (pixel size 1, image centred on the origin, ``half = N/2``):

* pixel ``(ix, iy)`` covers ``[ix-half, ix+1-half] x [iy-half, iy+1-half]``;
  its vector index is ``(N-1-iy) + N*ix`` (row ``N-1-iy`` from the top,
  column ``ix``, column-major);
* parallel beam: view angle ``th``; ray ``i`` of ``p`` has detector offset
  ``s_i = i - (p-1)/2``, origin ``(cos(th)*s_i, sin(th)*s_i)``, direction
  ``(-sin th, cos th)``;
* fan beam ("fancurved", ``run_2D_phantom.m:12``): source at
  ``R*(cos th, sin th)``, ``R = 2N``; ray ``i`` leaves the source at fan angle
  ``g_i = -gmax + i*dg`` from the central ray, ``gmax = asin(sqrt(2)/2*N/R)``,
  ``dg = 2*gmax/(p-1)``; direction ``-(c*cg - s*sg, s*cg + c*sg)``.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------
# ray tables (shared verbatim with the device generator via the C-ABI)
# ----------------------------------------------------------------------------

def fan_gmax(N: int, R: float) -> float:
    return math.asin(math.sqrt(2.0) / 2.0 * N / R)


def ray_tables(N: int, angles_deg, p: int, geometry: str = "parallel", R: float | None = None):
    """Return ``(cos_th, sin_th, ray_a, ray_b)``.

    parallel: ``ray_a = s_i`` (detector offsets), ``ray_b`` unused (zeros).
    fan:      ``ray_a = cos(g_i)``, ``ray_b = sin(g_i)``.
    """
    th = np.deg2rad(np.asarray(angles_deg, dtype=float))
    cos_th = np.cos(th)
    sin_th = np.sin(th)
    i = np.arange(p, dtype=float)
    if geometry == "parallel":
        ray_a = i - (p - 1) / 2.0
        ray_b = np.zeros(p)
    elif geometry == "fan":
        if R is None:
            R = 2.0 * N
        gmax = fan_gmax(N, R)
        dg = 2.0 * gmax / (p - 1)
        g = -gmax + i * dg
        ray_a = np.cos(g)
        ray_b = np.sin(g)
    else:
        raise ValueError("geometry must be 'parallel' or 'fan'")
    return cos_th, sin_th, ray_a, ray_b


def _rays(N, cos_th, sin_th, ray_a, ray_b, geometry, R):
    """Origins and directions of all rays, view-major (row = view*p + i)."""
    c = cos_th[:, None]
    s = sin_th[:, None]
    if geometry == "parallel":
        a = ray_a[None, :]
        ox = c * a
        oy = s * a
        dx = np.broadcast_to(-s, ox.shape).copy()
        dy = np.broadcast_to(c, ox.shape).copy()
    else:
        cg = ray_a[None, :]
        sg = ray_b[None, :]
        ox = np.broadcast_to(R * c, (c.shape[0], cg.shape[1])).copy()
        oy = np.broadcast_to(R * s, ox.shape).copy()
        dx = -(c * cg - s * sg)
        dy = -(s * cg + c * sg)
    return ox.ravel(), oy.ravel(), dx.ravel(), dy.ravel()


def _slab(o, d, half):
    """Entry/exit parameters of one axis slab; handles d == 0 explicitly."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t1 = (-half - o) * inv
        t2 = (half - o) * inv
    tlo = np.minimum(t1, t2)
    thi = np.maximum(t1, t2)
    zero = d == 0.0
    inside = (o >= -half) & (o < half)
    tlo = np.where(zero, np.where(inside, -np.inf, np.inf), tlo)
    thi = np.where(zero, np.inf, thi)
    return inv, tlo, thi


def projector(N: int, angles_deg, p: int | None = None, geometry: str = "parallel",
              R: float | None = None) -> sp.csr_matrix:
    """Line-intersection system matrix ``A`` (m = n_views*p rows, n = N*N
    columns), CSR with the entries of each row in traversal order."""
    if p is None:
        p = int(round(math.sqrt(2.0) * N))
    if R is None:
        R = 2.0 * N
    cos_th, sin_th, ray_a, ray_b = ray_tables(N, angles_deg, p, geometry, R)
    ox, oy, dx, dy = _rays(N, cos_th, sin_th, ray_a, ray_b, geometry, float(R))
    M = ox.shape[0]
    half = N / 2.0
    invdx, txlo, txhi = _slab(ox, dx, half)
    invdy, tylo, tyhi = _slab(oy, dy, half)
    tmin = np.maximum(txlo, tylo)
    tmax = np.minimum(txhi, tyhi)
    active = tmax > tmin
    tsafe = np.where(active, tmin, 0.0)
    ex = ox + tsafe * dx
    ey = oy + tsafe * dy
    ix = np.clip(np.floor(ex + half), 0, N - 1)
    iy = np.clip(np.floor(ey + half), 0, N - 1)
    offx = np.where(dx > 0, 1.0, 0.0)
    offy = np.where(dy > 0, 1.0, 0.0)
    sgx = np.where(dx > 0, 1.0, -1.0)
    sgy = np.where(dy > 0, 1.0, -1.0)
    zx = dx == 0.0
    zy = dy == 0.0
    t = tsafe.copy()
    rows_l, cols_l, vals_l = [], [], []
    ray_id = np.arange(M)
    for _ in range(2 * N + 2):
        if not active.any():
            break
        with np.errstate(invalid="ignore"):
            tmx = np.where(zx, np.inf, ((ix + offx) - half - ox) * invdx)
            tmy = np.where(zy, np.inf, ((iy + offy) - half - oy) * invdy)
        tn = np.minimum(tmx, tmy)
        ln = tn - t
        emit = active & (ln > 0)
        if emit.any():
            rows_l.append(ray_id[emit])
            cols_l.append(((N - 1) - iy[emit]) + N * ix[emit])
            vals_l.append(ln[emit])
        stepx = tmx <= tmy
        stepy = tmy <= tmx
        ix = np.where(active & stepx, ix + sgx, ix)
        iy = np.where(active & stepy, iy + sgy, iy)
        t = np.where(active, tn, t)
        active = active & (ix >= 0) & (ix < N) & (iy >= 0) & (iy < N)
    if rows_l:
        rows = np.concatenate(rows_l)
        cols = np.concatenate(cols_l).astype(np.int64)
        vals = np.concatenate(vals_l)
    else:
        rows = np.zeros(0, dtype=np.int64)
        cols = np.zeros(0, dtype=np.int64)
        vals = np.zeros(0)
    order = np.argsort(rows, kind="stable")  # keeps traversal order within a row
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=indptr[1:])
    A = sp.csr_matrix((vals, cols.astype(np.int32), indptr), shape=(M, N * N))
    return A


def backprojector_pixel_driven(N: int, angles_deg, p: int | None = None,
                               geometry: str = "parallel", R: float | None = None) -> sp.csr_matrix:
    """Structurally unmatched back-projector ``B`` (n x m): pixel-driven with
    linear interpolation between the two nearest detector bins (SURVEY.md
    §8(d) "cfg 4-5").  Entries of a row are ordered by view, then bin."""
    if p is None:
        p = int(round(math.sqrt(2.0) * N))
    if R is None:
        R = 2.0 * N
    cos_th, sin_th, _, _ = ray_tables(N, angles_deg, p, geometry, R)
    nv = cos_th.shape[0]
    half = N / 2.0
    ixs, iys = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float), indexing="ij")
    ixs = ixs.ravel()
    iys = iys.ravel()
    xc = (ixs + 0.5) - half
    yc = (iys + 0.5) - half
    prow = (((N - 1) - iys) + N * ixs).astype(np.int64)
    rows_l, cols_l, vals_l = [], [], []
    if geometry == "fan":
        gmax = fan_gmax(N, float(R))
        dg = 2.0 * gmax / (p - 1)
    for v in range(nv):
        c = cos_th[v]
        s = sin_th[v]
        if geometry == "parallel":
            f = (xc * c + yc * s) + (p - 1) / 2.0
            scale = np.ones_like(f)
        else:
            rx = xc - R * c
            ry = yc - R * s
            ecx, ecy = -c, -s
            cr = ecx * ry - ecy * rx
            dt = ecx * rx + ecy * ry
            g = np.arctan2(cr, dt)
            f = (g + gmax) / dg
            scale = 1.0 / (np.sqrt(rx * rx + ry * ry) * dg)
        i0 = np.floor(f)
        w1 = f - i0
        w0 = 1.0 - w1
        ok0 = (i0 >= 0) & (i0 < p)
        ok1 = (i0 + 1 >= 0) & (i0 + 1 < p)
        rows_l += [prow[ok0], prow[ok1]]
        cols_l += [(v * p + i0[ok0]).astype(np.int64), (v * p + i0[ok1] + 1).astype(np.int64)]
        vals_l += [(w0 * scale)[ok0], (w1 * scale)[ok1]]
    rows = np.concatenate(rows_l)
    cols = np.concatenate(cols_l)
    vals = np.concatenate(vals_l)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(N * N + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=N * N), out=indptr[1:])
    return sp.csr_matrix((vals, cols.astype(np.int32), indptr), shape=(N * N, nv * p))


def backprojector_perturbed(A: sp.csr_matrix, c: float, seed: int = 0) -> sp.csr_matrix:
    """``B = A' + c*E`` with ``E ~ N(0,1)`` on the sparsity pattern of ``A'``,
    ``||E||_F = 1`` (``run_2D_phantom.m:79,87-89`` uses a dense ``E``;
    pattern-restricted here — documented deviation, SURVEY.md §7 hard part 8)."""
    At = A.T.tocsr()
    At.sort_indices()
    rng = np.random.default_rng(seed)
    E = rng.standard_normal(At.data.shape[0])
    E /= np.linalg.norm(E)
    return sp.csr_matrix((At.data + c * E, At.indices.copy(), At.indptr.copy()), shape=At.shape)


def shepp_logan(N: int) -> np.ndarray:
    """Modified Shepp-Logan phantom, N x N, returned column-major vectorised
    (``x_true(:)``; ``reshape(x, N, N)`` at ``run_2D_phantom.m:57``)."""
    #        A     a      b      x0     y0    phi
    ell = [(1.0, .69, .92, 0.0, 0.0, 0.0),
           (-.8, .6624, .8740, 0.0, -.0184, 0.0),
           (-.2, .1100, .3100, .22, 0.0, -18.0),
           (-.2, .1600, .4100, -.22, 0.0, 18.0),
           (.1, .2100, .2500, 0.0, .35, 0.0),
           (.1, .0460, .0460, 0.0, .1, 0.0),
           (.1, .0460, .0460, 0.0, -.1, 0.0),
           (.1, .0460, .0230, -.08, -.605, 0.0),
           (.1, .0230, .0230, 0.0, -.606, 0.0),
           (.1, .0230, .0460, .06, -.605, 0.0)]
    xs = ((np.arange(N) + 0.5) - N / 2.0) / (N / 2.0)
    X, Y = np.meshgrid(xs, -xs)  # row 0 is the top (y = +1)
    img = np.zeros((N, N))
    for Aa, a, b, x0, y0, phi in ell:
        ph = math.radians(phi)
        xr = (X - x0) * math.cos(ph) + (Y - y0) * math.sin(ph)
        yr = -(X - x0) * math.sin(ph) + (Y - y0) * math.cos(ph)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += Aa
    return img.ravel(order="F")


def make_ct_problem(N: int, n_views: int = 180, geometry: str = "parallel", b_kind: str = "matched",
                    noise: float = 0.01, mismatch: float = 1e-2, seed: int = 0):
    """Build ``(A, B, b_noisy, x_true)`` for one of BASELINE.json's config
    shapes.  ``b_kind``: 'matched' (B = A'), 'perturbed' (A' + c E) or
    'pixel' (pixel-driven interpolating back-projector)."""
    if geometry == "parallel":
        angles = np.arange(n_views) * (180.0 / n_views)
    else:
        angles = np.arange(n_views) * (360.0 / n_views)
    p = int(round(math.sqrt(2.0) * N))
    A = projector(N, angles, p, geometry)
    if b_kind == "matched":
        B = A.T.tocsr()
        B.sort_indices()
    elif b_kind == "perturbed":
        B = backprojector_perturbed(A, mismatch, seed + 1)
    elif b_kind == "pixel":
        B = backprojector_pixel_driven(N, angles, p, geometry)
    else:
        raise ValueError("b_kind")
    x_true = shepp_logan(N)
    b_exact = np.asarray(A @ x_true).ravel()
    rng = np.random.default_rng(seed)
    e = rng.standard_normal(b_exact.shape)
    b = b_exact + noise * np.linalg.norm(b_exact) * e / np.linalg.norm(e)  # run_2D_phantom.m:18-19
    return A, B, b, x_true
