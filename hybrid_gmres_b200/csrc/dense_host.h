// Small dense host-side linear algebra for the projected problems.  By the
// north star the (k+1) x k Hessenberg / bidiagonal problem stays on the host:
// least squares, SPD solves, singular values, GCV value and fminbnd.
#pragma once
#include <cstdint>
#include <vector>

namespace hgd {

// Incremental Givens QR of an upper-Hessenberg matrix for
//   y_k = argmin || beta e1 - H(1:k+1,1:k) y ||      (hybrid_ba_gmres_rtp.m:28-29)
class HessenbergLS {
public:
    void reset(int kmax, double beta);
    // col has k+1 entries: H(1:k+1, k) for the next column k (1-based count = ncols()+1)
    void add_column(const double* col);
    int ncols() const { return k_; }
    // solves the k x k triangular system; y has k entries
    void solve(double* y) const;
    // |last rhs entry| = Krylov residual norm of the projected problem
    double projected_residual() const { return k_ < (int)g_.size() ? (g_[k_] < 0 ? -g_[k_] : g_[k_]) : 0.0; }

private:
    int kmax_ = 0, k_ = 0;
    std::vector<double> R_;  // kmax x kmax column-major upper triangular
    std::vector<double> cs_, sn_, g_;
};

// Row-by-row (Cholesky-Banachiewicz) factorisation of a growing SPD matrix
//   G_k + lambda I,  G_k = (A Q_k)'(A Q_k)              (hybrid_ab_gmres_rtp.m:32)
// Adding a bordering row reproduces, operation for operation, what a from-scratch
// row-oriented Cholesky of the k x k matrix computes.
class BorderedCholesky {
public:
    void reset(int kmax, double lambda);
    // g has k entries: G(k, 1:k) (new row incl. diagonal, WITHOUT lambda). Returns false
    // when the pivot is not positive (caller falls back to LU on the full matrix).
    bool add_row(const double* g);
    int size() const { return k_; }
    void solve(const double* rhs, double* y) const;  // L L' y = rhs

private:
    int kmax_ = 0, k_ = 0;
    double lambda_ = 0.0;
    std::vector<double> L_;  // kmax x kmax row-major lower triangular
};

// y = M \ rhs for square M (n x n column-major, ld): Cholesky when M is symmetric
// with positive pivots, otherwise LU with partial pivoting — MATLAB mldivide's
// dispatch for a full square matrix.  M is overwritten. Returns false if singular.
bool solve_square(int n, double* M, int ld, const double* rhs, double* y);

// singular values (descending) of the n x n column-major matrix M (overwritten);
// one-sided Jacobi.
void singular_values(int n, double* M, int ld, double* s);

// gcv_function.m:33-58 given the Arnoldi output; H is (k+1) x k column-major.
// sv2 = squared singular values of H(1:k,1:k) (lambda independent, cached).
double gcv_value(double lambda, const double* H, int ldh, int k, double beta, double trace_m,
                 const double* sv);

// MATLAB fminbnd restated (Forsythe-Malcolm-Moler / Brent localmin).
struct FminResult {
    double x, fval;
    int exitflag, funccount;
};
template <class F>
FminResult fminbnd(F&& fun, double ax, double bx, double tolx, int maxfun, int maxiter,
                   double* trace, int trace_cap);

}  // namespace hgd

#include "dense_host_impl.h"
