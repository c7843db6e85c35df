// Host-side projected-problem arithmetic (see dense_host.h).
#include "dense_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace hgd {

// ---------------------------------------------------------------------------
void HessenbergLS::reset(int kmax, double beta) {
    kmax_ = kmax;
    k_ = 0;
    R_.assign((size_t)kmax * kmax, 0.0);
    cs_.assign(kmax, 0.0);
    sn_.assign(kmax, 0.0);
    g_.assign(kmax + 1, 0.0);
    g_[0] = beta;
}

void HessenbergLS::add_column(const double* col) {
    const int k = k_;
    std::vector<double> t(col, col + k + 2);
    for (int i = 0; i < k; ++i) {
        const double a = cs_[i] * t[i] + sn_[i] * t[i + 1];
        t[i + 1] = -sn_[i] * t[i] + cs_[i] * t[i + 1];
        t[i] = a;
    }
    const double r = std::hypot(t[k], t[k + 1]);
    double c = 1.0, s = 0.0;
    if (r != 0.0) {
        c = t[k] / r;
        s = t[k + 1] / r;
    }
    cs_[k] = c;
    sn_[k] = s;
    for (int i = 0; i < k; ++i) R_[(size_t)k * kmax_ + i] = t[i];
    R_[(size_t)k * kmax_ + k] = r;
    g_[k + 1] = -s * g_[k];
    g_[k] = c * g_[k];
    k_ = k + 1;
}

void HessenbergLS::solve(double* y) const {
    const int k = k_;
    for (int i = k - 1; i >= 0; --i) {
        double acc = g_[i];
        for (int j = i + 1; j < k; ++j) acc -= R_[(size_t)j * kmax_ + i] * y[j];
        y[i] = acc / R_[(size_t)i * kmax_ + i];
    }
}

// ---------------------------------------------------------------------------
void BorderedCholesky::reset(int kmax, double lambda) {
    kmax_ = kmax;
    k_ = 0;
    lambda_ = lambda;
    L_.assign((size_t)kmax * kmax, 0.0);
}

bool BorderedCholesky::add_row(const double* g) {
    const int k = k_;
    double* Lk = &L_[(size_t)k * kmax_];
    for (int j = 0; j < k; ++j) {
        const double* Lj = &L_[(size_t)j * kmax_];
        double acc = g[j];
        for (int t = 0; t < j; ++t) acc -= Lk[t] * Lj[t];
        Lk[j] = acc / Lj[j];
    }
    double d = g[k] + lambda_;
    for (int t = 0; t < k; ++t) d -= Lk[t] * Lk[t];
    if (!(d > 0.0)) return false;
    Lk[k] = std::sqrt(d);
    k_ = k + 1;
    return true;
}

void BorderedCholesky::solve(const double* rhs, double* y) const {
    const int k = k_;
    std::vector<double> z(k);
    for (int i = 0; i < k; ++i) {
        const double* Li = &L_[(size_t)i * kmax_];
        double acc = rhs[i];
        for (int t = 0; t < i; ++t) acc -= Li[t] * z[t];
        z[i] = acc / Li[i];
    }
    for (int i = k - 1; i >= 0; --i) {
        double acc = z[i];
        for (int t = i + 1; t < k; ++t) acc -= L_[(size_t)t * kmax_ + i] * y[t];
        y[i] = acc / L_[(size_t)i * kmax_ + i];
    }
}

// ---------------------------------------------------------------------------
static bool cholesky_solve(int n, const double* M, int ld, const double* rhs, double* y) {
    // row-oriented Cholesky on a copy (lower triangle), skipping structural zeros
    std::vector<double> L((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) {
        double* Li = &L[(size_t)i * n];
        for (int j = 0; j < i; ++j) {
            const double* Lj = &L[(size_t)j * n];
            double acc = M[(size_t)j * ld + i];
            for (int t = 0; t < j; ++t) acc -= Li[t] * Lj[t];
            Li[j] = acc / Lj[j];
        }
        double d = M[(size_t)i * ld + i];
        for (int t = 0; t < i; ++t) d -= Li[t] * Li[t];
        if (!(d > 0.0)) return false;
        Li[i] = std::sqrt(d);
    }
    std::vector<double> z(n);
    for (int i = 0; i < n; ++i) {
        double acc = rhs[i];
        for (int t = 0; t < i; ++t) acc -= L[(size_t)i * n + t] * z[t];
        z[i] = acc / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double acc = z[i];
        for (int t = i + 1; t < n; ++t) acc -= L[(size_t)t * n + i] * y[t];
        y[i] = acc / L[(size_t)i * n + i];
    }
    return true;
}

static bool lu_solve(int n, double* M, int ld, const double* rhs, double* y) {
    std::vector<int> piv(n);
    std::vector<double> b(rhs, rhs + n);
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double mx = std::fabs(M[(size_t)k * ld + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = std::fabs(M[(size_t)k * ld + i]);
            if (v > mx) {
                mx = v;
                p = i;
            }
        }
        piv[k] = p;
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(M[(size_t)j * ld + k], M[(size_t)j * ld + p]);
            std::swap(b[k], b[p]);
        }
        const double d = M[(size_t)k * ld + k];
        if (d == 0.0) {
            ok = false;
            continue;
        }
        for (int i = k + 1; i < n; ++i) {
            const double l = M[(size_t)k * ld + i] / d;
            M[(size_t)k * ld + i] = l;
            if (l != 0.0) {
                for (int j = k + 1; j < n; ++j) M[(size_t)j * ld + i] -= l * M[(size_t)j * ld + k];
                b[i] -= l * b[k];
            }
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double acc = b[i];
        for (int j = i + 1; j < n; ++j) acc -= M[(size_t)j * ld + i] * y[j];
        y[i] = acc / M[(size_t)i * ld + i];
    }
    return ok;
}

bool solve_square(int n, double* M, int ld, const double* rhs, double* y) {
    bool sym = true;
    for (int j = 0; j < n && sym; ++j) {
        if (!(M[(size_t)j * ld + j] > 0.0)) sym = false;
        for (int i = j + 1; i < n && sym; ++i)
            if (M[(size_t)j * ld + i] != M[(size_t)i * ld + j]) sym = false;
    }
    if (sym && cholesky_solve(n, M, ld, rhs, y)) return true;
    return lu_solve(n, M, ld, rhs, y);
}

// ---------------------------------------------------------------------------
void singular_values(int n, double* M, int ld, double* s) {
    const double eps = 2.220446049250313e-16;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double* ap = M + (size_t)p * ld;
                double* aq = M + (size_t)q * ld;
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < n; ++i) {
                    alpha += ap[i] * ap[i];
                    beta += aq[i] * aq[i];
                    gamma += ap[i] * aq[i];
                }
                if (gamma == 0.0 || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t);
                const double sn = c * t;
                for (int i = 0; i < n; ++i) {
                    const double x = ap[i], yv = aq[i];
                    ap[i] = c * x - sn * yv;
                    aq[i] = sn * x + c * yv;
                }
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < n; ++j) {
        double acc = 0;
        for (int i = 0; i < n; ++i) acc += M[(size_t)j * ld + i] * M[(size_t)j * ld + i];
        s[j] = std::sqrt(acc);
    }
    std::sort(s, s + n, [](double a, double b) { return a > b; });
}

// ---------------------------------------------------------------------------
double gcv_value(double lambda, const double* H, int ldh, int k, double beta, double trace_m,
                 const double* sv) {
    const double eps = 2.220446049250313e-16;
    // gcv_function.m:38  yk = (Hk'*Hk + lambda*eye(k)) \ (Hk'*tk)
    std::vector<double> N((size_t)k * k), rhs(k), y(k, 0.0);
    for (int j = 0; j < k; ++j) {
        for (int i = 0; i <= j; ++i) {
            double acc = 0.0;
            for (int t = 0; t <= k; ++t) acc += H[(size_t)i * ldh + t] * H[(size_t)j * ldh + t];
            N[(size_t)j * k + i] = acc;
            N[(size_t)i * k + j] = acc;
        }
        N[(size_t)j * k + j] += lambda;
        rhs[j] = H[(size_t)j * ldh + 0] * beta;
    }
    solve_square(k, N.data(), k, rhs.data(), y.data());
    // :40 residual_norm_sq = norm(tk - Hk*yk)^2
    double ss = 0.0;
    for (int t = 0; t <= k; ++t) {
        double r = (t == 0) ? beta : 0.0;
        for (int j = 0; j < k; ++j) r -= H[(size_t)j * ldh + t] * y[j];
        ss += r * r;
    }
    const double nr = std::sqrt(ss);
    const double residual_norm_sq = nr * nr;
    // :51-52
    double trace_val = 0.0;
    for (int i = 0; i < k; ++i) trace_val += (sv[i] * sv[i]) / (sv[i] * sv[i] + lambda);
    const double denominator = (trace_m - trace_val) * (trace_m - trace_val);
    double gcv = residual_norm_sq / denominator;
    if (std::isnan(gcv) || std::isinf(gcv) || denominator < eps) gcv = 1e20;  // :56-58
    return gcv;
}

}  // namespace hgd
