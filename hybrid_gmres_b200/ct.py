"""Device generation of the synthetic CT matrices (``hg_ct_projector`` /
``hg_ct_backprojector``): SURVEY.md §8(d) "Synthetic inputs".  Only the small
per-view / per-ray trig tables are computed on the host; the matrices are
built in HBM and never cross PCIe."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._lib import check
from .api import Context, DeviceMatrix, _ptr, default_context


def ray_tables(N, angles_deg, p, geometry="parallel", R=None):
    """Per-view ``cos/sin`` and per-ray tables handed to the device generator.
    parallel: ``ray_a`` = detector offsets ``i-(p-1)/2``; fan: ``ray_a``/``ray_b``
    = ``cos``/``sin`` of the fan angle, ``gmax = asin(sqrt(2)/2*N/R)``."""
    th = np.deg2rad(np.asarray(angles_deg, dtype=float))
    cos_th, sin_th = np.cos(th), np.sin(th)
    i = np.arange(p, dtype=float)
    if geometry == "parallel":
        return cos_th, sin_th, i - (p - 1) / 2.0, np.zeros(p)
    if geometry != "fan":
        raise ValueError("geometry must be 'parallel' or 'fan'")
    if R is None:
        R = 2.0 * N
    gmax = math.asin(math.sqrt(2.0) / 2.0 * N / R)
    dg = 2.0 * gmax / (p - 1)
    g = -gmax + i * dg
    return cos_th, sin_th, np.cos(g), np.sin(g)


def _geom(geometry):
    return {"parallel": 0, "fan": 1}[geometry]


def ct_projector(N, angles_deg, p=None, geometry="parallel", R=None, ctx: Context | None = None) -> DeviceMatrix:
    ctx = ctx or default_context()
    if p is None:
        p = int(round(math.sqrt(2.0) * N))
    if R is None:
        R = 2.0 * N
    c, s, a, b = (np.ascontiguousarray(t) for t in ray_tables(N, angles_deg, p, geometry, R))
    h = C.c_void_p()
    check(ctx._lib.hg_ct_projector(ctx._h, int(N), int(c.shape[0]), int(p), _geom(geometry), float(R),
                                   _ptr(c), _ptr(s), _ptr(a), _ptr(b), C.byref(h)))
    return DeviceMatrix(h, ctx)


def ct_backprojector(N, angles_deg, p=None, geometry="parallel", R=None, ctx: Context | None = None) -> DeviceMatrix:
    ctx = ctx or default_context()
    if p is None:
        p = int(round(math.sqrt(2.0) * N))
    if R is None:
        R = 2.0 * N
    th = np.deg2rad(np.asarray(angles_deg, dtype=float))
    c, s = np.ascontiguousarray(np.cos(th)), np.ascontiguousarray(np.sin(th))
    h = C.c_void_p()
    check(ctx._lib.hg_ct_backprojector(ctx._h, int(N), int(c.shape[0]), int(p), _geom(geometry), float(R),
                                       _ptr(c), _ptr(s), C.byref(h)))
    return DeviceMatrix(h, ctx)


def shepp_logan(N: int) -> np.ndarray:
    """Modified Shepp-Logan phantom, N x N, column-major vectorised (``x_true(:)``,
    ``reshape(x, N, N)`` at ``run_2D_phantom.m:57``) — the synthetic ``x_true`` of the bench."""
    ell = [(1.0, .69, .92, 0.0, 0.0, 0.0), (-.8, .6624, .8740, 0.0, -.0184, 0.0),
           (-.2, .1100, .3100, .22, 0.0, -18.0), (-.2, .1600, .4100, -.22, 0.0, 18.0),
           (.1, .2100, .2500, 0.0, .35, 0.0), (.1, .0460, .0460, 0.0, .1, 0.0),
           (.1, .0460, .0460, 0.0, -.1, 0.0), (.1, .0460, .0230, -.08, -.605, 0.0),
           (.1, .0230, .0230, 0.0, -.606, 0.0), (.1, .0230, .0460, .06, -.605, 0.0)]
    xs = ((np.arange(N) + 0.5) - N / 2.0) / (N / 2.0)
    X, Y = np.meshgrid(xs, -xs)
    img = np.zeros((N, N))
    for amp, a, b, x0, y0, phi in ell:
        ph = math.radians(phi)
        xr = (X - x0) * math.cos(ph) + (Y - y0) * math.sin(ph)
        yr = -(X - x0) * math.sin(ph) + (Y - y0) * math.cos(ph)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += amp
    return img.ravel(order="F")


def ct_projector_rows(N, angles_deg, p, geometry, row_lo, row_hi, R=None, ctx: Context | None = None) -> DeviceMatrix:
    """Rows ``[row_lo,row_hi)`` (rays, view-major) of the projector — one rank's ``A_p``."""
    ctx = ctx or default_context()
    if R is None:
        R = 2.0 * N
    c, s, a, b = (np.ascontiguousarray(t) for t in ray_tables(N, angles_deg, p, geometry, R))
    h = C.c_void_p()
    check(ctx._lib.hg_ct_projector_rows(ctx._h, int(N), int(c.shape[0]), int(p), _geom(geometry), float(R),
                                        _ptr(c), _ptr(s), _ptr(a), _ptr(b), int(row_lo), int(row_hi), C.byref(h)))
    return DeviceMatrix(h, ctx)


def ct_backprojector_cols(N, angles_deg, p, geometry, col_lo, col_hi, R=None, ctx: Context | None = None) -> DeviceMatrix:
    """Sinogram columns ``[col_lo,col_hi)`` of the back-projector (local column indices) —
    one rank's ``B^p``."""
    ctx = ctx or default_context()
    if R is None:
        R = 2.0 * N
    th = np.deg2rad(np.asarray(angles_deg, dtype=float))
    c, s = np.ascontiguousarray(np.cos(th)), np.ascontiguousarray(np.sin(th))
    h = C.c_void_p()
    check(ctx._lib.hg_ct_backprojector_cols(ctx._h, int(N), int(c.shape[0]), int(p), _geom(geometry), float(R),
                                            _ptr(c), _ptr(s), int(col_lo), int(col_hi), C.byref(h)))
    return DeviceMatrix(h, ctx)


def tile_permutation(N: int, tile: int = 4) -> np.ndarray:
    """Pixel order that keeps ``tile x tile`` blocks of the N x N image contiguous: ``perm[new] = old``
    with ``old`` the column-major pixel index of ``x_true(:)``.  A 4 x 4 tile of doubles is one 128-byte
    line, so a ray of the projector touches ~5 pixels per gathered line instead of 1-2 when it runs
    across the image columns (the L1 wavefront count of the SpMV gather, profiles/r01_spmv_variants.md).
    Use as ``A.permute(None, perm)``, ``B.permute(perm, None)``, ``x_true[perm]``; ``x[perm] = x_tiled``."""
    if N % tile:
        raise ValueError("tile_permutation: N must be a multiple of the tile size")
    nt = N // tile
    tr, tc, ir, ic = np.meshgrid(np.arange(nt), np.arange(nt), np.arange(tile), np.arange(tile), indexing="ij")
    # new index = ((tile_col * nt + tile_row) * tile + in_col) * tile + in_row  (column-major at both levels)
    new = ((tc * nt + tr) * tile + ic) * tile + ir
    old = (tc * tile + ic) * N + (tr * tile + ir)
    perm = np.empty(N * N, dtype=np.int32)
    perm[new.ravel()] = old.ravel()
    return perm
