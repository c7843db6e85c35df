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
