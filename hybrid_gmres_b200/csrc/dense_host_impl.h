// Template part of dense_host.h (fminbnd).
#pragma once
#include <cmath>

namespace hgd {

template <class F>
FminResult fminbnd(F&& fun, double ax, double bx, double tolx, int maxfun, int maxiter,
                   double* trace, int trace_cap) {
    const double eps = 2.220446049250313e-16;
    const double seps = std::sqrt(eps);
    const double c = 0.5 * (3.0 - std::sqrt(5.0));
    auto sign = [](double v) { return (double)((v > 0) - (v < 0)); };
    int ntrace = 0;
    auto eval = [&](double xx) {
        if (trace && ntrace < trace_cap) trace[ntrace] = xx;
        ++ntrace;
        return fun(xx);
    };
    double a = ax, b = bx;
    double v = a + c * (b - a);
    double w = v, xf = v;
    double d = 0.0, e = 0.0;
    double x = xf;
    double fx = eval(x);
    int funccount = 1, iter = 0;
    double fv = fx, fw = fx;
    double xm = 0.5 * (a + b);
    double tol1 = seps * std::fabs(xf) + tolx / 3.0;
    double tol2 = 2.0 * tol1;
    int exitflag = 1;
    while (std::fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        bool gs = true;
        if (std::fabs(e) > tol1) {
            gs = false;
            double r = (xf - w) * (fx - fv);
            double q = (xf - v) * (fx - fw);
            double p = (xf - v) * q - (xf - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = std::fabs(q);
            r = e;
            e = d;
            if ((std::fabs(p) < std::fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                d = p / q;
                x = xf + d;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
                    const double si = sign(xm - xf) + ((xm - xf) == 0 ? 1.0 : 0.0);
                    d = tol1 * si;
                }
            } else {
                gs = true;
            }
        }
        if (gs) {
            e = (xf >= xm) ? a - xf : b - xf;
            d = c * e;
        }
        const double si = sign(d) + (d == 0 ? 1.0 : 0.0);
        x = xf + si * std::fmax(std::fabs(d), tol1);
        const double fu = eval(x);
        ++funccount;
        ++iter;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            v = w; fv = fw;
            w = xf; fw = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fw) || (w == xf)) {
                v = w; fv = fw;
                w = x; fw = fu;
            } else if ((fu <= fv) || (v == xf) || (v == w)) {
                v = x; fv = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = seps * std::fabs(xf) + tolx / 3.0;
        tol2 = 2.0 * tol1;
        if (funccount >= maxfun || iter >= maxiter) {
            exitflag = 0;
            break;
        }
    }
    return FminResult{xf, fx, exitflag, funccount};
}

}  // namespace hgd
