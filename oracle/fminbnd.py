"""MATLAB ``fminbnd`` restated (TEST INFRASTRUCTURE — see oracle/__init__.py).

The reference drives ``gcv_function`` through MATLAB's built-in ``fminbnd``
(``analyze_regularization.m:37-46``: bounds ``[1e-9, 1e-1]``,
``optimset('TolX',1e-8)``; also ``plot_error_vs_mismatch_norm.m:46-50``,
``plot_error_vs_noise_level.m:38-44``).  ``fminbnd`` is the golden-section /
parabolic-interpolation method of Forsythe, Malcolm & Moler (Brent's
``localmin``); this restates that published algorithm with MATLAB's constants
(``sqrt(eps)``, ``TolX/3``).  PARITY UNPINNED against MATLAB itself.
"""
from __future__ import annotations

import math

_EPS = 2.220446049250313e-16


def _sign(x: float) -> float:
    return float(x > 0) - float(x < 0)


def fminbnd(fun, ax, bx, tolx=1e-4, max_fun_evals=500, max_iter=500, trace=None):
    """Return ``(xf, fval, exitflag, funccount)``.  ``trace`` (a list) receives
    every evaluated abscissa in order — used to check that two GCV objectives
    produce the *identical* evaluation sequence."""
    seps = math.sqrt(_EPS)
    c = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = float(ax), float(bx)
    v = a + c * (b - a)
    w = v
    xf = v
    d = 0.0
    e = 0.0
    x = xf
    fx = fun(x)
    if trace is not None:
        trace.append(x)
    funccount = 1
    iters = 0
    fv = fx
    fw = fx
    xm = 0.5 * (a + b)
    tol1 = seps * abs(xf) + tolx / 3.0
    tol2 = 2.0 * tol1
    exitflag = 1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        gs = True
        if abs(e) > tol1:
            gs = False
            r = (xf - w) * (fx - fv)
            q = (xf - v) * (fx - fw)
            p = (xf - v) * q - (xf - w) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = d
            if (abs(p) < abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                d = p / q
                x = xf + d
                if ((x - a) < tol2) or ((b - x) < tol2):
                    si = _sign(xm - xf) + ((xm - xf) == 0)
                    d = tol1 * si
            else:
                gs = True
        if gs:
            e = (a - xf) if xf >= xm else (b - xf)
            d = c * e
        si = _sign(d) + (d == 0)
        x = xf + si * max(abs(d), tol1)
        fu = fun(x)
        if trace is not None:
            trace.append(x)
        funccount += 1
        iters += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            v, fv = w, fw
            w, fw = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if (fu <= fw) or (w == xf):
                v, fv = w, fw
                w, fw = x, fu
            elif (fu <= fv) or (v == xf) or (v == w):
                v, fv = x, fu
        xm = 0.5 * (a + b)
        tol1 = seps * abs(xf) + tolx / 3.0
        tol2 = 2.0 * tol1
        if funccount >= max_fun_evals or iters >= max_iter:
            exitflag = 0
            break
    return xf, fx, exitflag, funccount
