/*
 * hg_oracle.c — multi-threaded C restatement of hybrid_ba_gmres_rtp.m / hybrid_ab_gmres_rtp.m
 * (TEST INFRASTRUCTURE).
 *
 * Two jobs: (i) the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) on every host
 * core; (ii) the checker at BASELINE's full sizes (1024^2, 2048^2 shards), where the NumPy oracle
 * is too slow: tests/test_gpu_fullsize.py and bench.py's `parity` field compare the CUDA path
 * with this file.  It is the same literal algorithm as oracle/solvers.py (which is pinned to the
 * executed reference source, oracle/__init__.py) with OpenMP-parallel sparse mat-vecs and BLAS-1
 * loops; tests/test_oracle_c.py holds it to oracle/solvers.py.  Never imported by the product.
 *
 * Besides the reference's MGS sweep (`orth` 0) it offers CGS2 (`orth` 1, the north star's
 * orthogonalisation) so the device algorithm can be compared like for like as well, and it can
 * return H, beta and the iterate after every iteration.  hybrid_ab_gmres_rtp.m:31 recomputes
 * A*Q(:,1:k) every iteration; the columns are deterministic, so caching them (done here) gives
 * bit-identical numbers in O(K nnz) instead of O(K^2 nnz).
 *
 * Statement map (hybrid_ba_gmres_rtp.m; the AB file differs in :28-33 only):
 *   :6  M_reg(v) = B*(A*v) + lambda*v      -> op()
 *   :7-13 d = B*b; r0 = d - M_reg(0); beta; Q(:,1)
 *   :19-26 MGS Arnoldi step, breakdown ==0
 *   :28-30 yk = H(1:k+1,1:k) \ [beta;0] (Householder QR least squares), x = Q(:,1:k)*yk
 *   :32-35 true residual, error, stop <=
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t rows, cols;
    const int64_t* ptr;
    const int32_t* idx;
    const double* val;
} csr_t;

int hgo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void spmv(const csr_t* M, const double* x, double* y) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < M->rows; ++r) {
        double s = 0.0;
        for (int64_t i = M->ptr[r]; i < M->ptr[r + 1]; ++i) s += M->val[i] * x[M->idx[i]];
        y[r] = s;
    }
}

static double dot(const double* a, const double* b, int64_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static void axpy(double a, const double* x, double* y, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}

/* least squares min || rhs - H y ||, H (k+1) x k column-major with leading dim ld, by Householder QR */
static void lstsq_hess(const double* H, int ld, int k, double beta, double* y) {
    int rows = k + 1;
    double* R = (double*)malloc(sizeof(double) * rows * k);
    double* g = (double*)calloc(rows, sizeof(double));
    for (int j = 0; j < k; ++j) memcpy(R + (size_t)j * rows, H + (size_t)j * ld, sizeof(double) * rows);
    g[0] = beta;
    for (int j = 0; j < k; ++j) {
        double nrm = 0.0;
        for (int i = j; i < rows; ++i) nrm += R[(size_t)j * rows + i] * R[(size_t)j * rows + i];
        nrm = sqrt(nrm);
        if (nrm == 0.0) continue;
        double alpha = R[(size_t)j * rows + j] > 0 ? -nrm : nrm;
        double v0 = R[(size_t)j * rows + j] - alpha;
        double vnorm2 = v0 * v0;
        for (int i = j + 1; i < rows; ++i) vnorm2 += R[(size_t)j * rows + i] * R[(size_t)j * rows + i];
        if (vnorm2 == 0.0) continue;
        /* apply I - 2 v v'/(v'v) to the remaining columns and to g */
        for (int c = j + 1; c < k; ++c) {
            double s = v0 * R[(size_t)c * rows + j];
            for (int i = j + 1; i < rows; ++i) s += R[(size_t)j * rows + i] * R[(size_t)c * rows + i];
            s = 2.0 * s / vnorm2;
            R[(size_t)c * rows + j] -= s * v0;
            for (int i = j + 1; i < rows; ++i) R[(size_t)c * rows + i] -= s * R[(size_t)j * rows + i];
        }
        double s = v0 * g[j];
        for (int i = j + 1; i < rows; ++i) s += R[(size_t)j * rows + i] * g[i];
        s = 2.0 * s / vnorm2;
        g[j] -= s * v0;
        for (int i = j + 1; i < rows; ++i) g[i] -= s * R[(size_t)j * rows + i];
        R[(size_t)j * rows + j] = alpha;
    }
    for (int i = k - 1; i >= 0; --i) {
        double acc = g[i];
        for (int j = i + 1; j < k; ++j) acc -= R[(size_t)j * rows + i] * y[j];
        y[i] = acc / R[(size_t)i * rows + i];
    }
    free(R);
    free(g);
}

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

void hgo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* H(1:k,k) and the orthogonalised v.  orth 0: hybrid_ab_gmres_rtp.m:20-23 (MGS);
 * orth 1: h1 = Q'v, v -= Q h1, h2 = Q'v, v -= Q h2, h = h1 + h2 (CGS2) */
static void orthogonalise(int orth, const double* Q, int64_t n, int k, double* v, double* hcol, double* tmp) {
    if (orth == 0) {
        for (int j = 0; j < k; ++j) {
            const double* qj = Q + (size_t)j * n;
            const double h = dot(qj, v, n);
            hcol[j] = h;
            axpy(-h, qj, v, n);
        }
        return;
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < k; ++j) tmp[j] = dot(Q + (size_t)j * n, v, n);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j < k; ++j) s += Q[(size_t)j * n + i] * tmp[j];
            v[i] -= s;
        }
        for (int j = 0; j < k; ++j) hcol[j] = pass == 0 ? tmp[j] : hcol[j] + tmp[j];
    }
}

/* y = M \ rhs for the symmetric k x k M (column-major, ld k): Cholesky when every pivot is
 * positive, else LU with partial pivoting — MATLAB's mldivide order (hybrid_ab_gmres_rtp.m:32) */
static void solve_sym(int k, const double* M, const double* rhs, double* y) {
    double* L = (double*)malloc(sizeof(double) * k * k);
    memcpy(L, M, sizeof(double) * k * k);
    int spd = 1;
    for (int j = 0; j < k && spd; ++j) {
        double d = L[(size_t)j * k + j];
        for (int t = 0; t < j; ++t) d -= L[(size_t)t * k + j] * L[(size_t)t * k + j];
        if (!(d > 0.0)) { spd = 0; break; }
        d = sqrt(d);
        L[(size_t)j * k + j] = d;
        for (int i = j + 1; i < k; ++i) {
            double s = L[(size_t)j * k + i];
            for (int t = 0; t < j; ++t) s -= L[(size_t)t * k + i] * L[(size_t)t * k + j];
            L[(size_t)j * k + i] = s / d;
        }
    }
    if (spd) {  /* L (lower, stored in columns) L' y = rhs */
        for (int i = 0; i < k; ++i) {
            double s = rhs[i];
            for (int t = 0; t < i; ++t) s -= L[(size_t)t * k + i] * y[t];
            y[i] = s / L[(size_t)i * k + i];
        }
        for (int i = k - 1; i >= 0; --i) {
            double s = y[i];
            for (int t = i + 1; t < k; ++t) s -= L[(size_t)i * k + t] * y[t];
            y[i] = s / L[(size_t)i * k + i];
        }
    } else {
        memcpy(L, M, sizeof(double) * k * k);
        int* piv = (int*)malloc(sizeof(int) * k);
        for (int i = 0; i < k; ++i) { piv[i] = i; y[i] = rhs[i]; }
        for (int j = 0; j < k; ++j) {
            int p = j;
            for (int i = j + 1; i < k; ++i)
                if (fabs(L[(size_t)j * k + i]) > fabs(L[(size_t)j * k + p])) p = i;
            if (p != j) {
                for (int c = 0; c < k; ++c) { double t = L[(size_t)c * k + j]; L[(size_t)c * k + j] = L[(size_t)c * k + p]; L[(size_t)c * k + p] = t; }
                double t = y[j]; y[j] = y[p]; y[p] = t;
            }
            const double d = L[(size_t)j * k + j];
            for (int i = j + 1; i < k; ++i) {
                const double f = L[(size_t)j * k + i] / d;
                L[(size_t)j * k + i] = f;
                for (int c = j + 1; c < k; ++c) L[(size_t)c * k + i] -= f * L[(size_t)c * k + j];
                y[i] -= f * y[j];
            }
        }
        for (int i = k - 1; i >= 0; --i) {
            double s = y[i];
            for (int c = i + 1; c < k; ++c) s -= L[(size_t)c * k + i] * y[c];
            y[i] = s / L[(size_t)i * k + i];
        }
        free(piv);
    }
    free(L);
}

/* hybrid_ab_gmres_rtp.m (kind 0) / hybrid_ba_gmres_rtp.m (kind 1).  Returns niters.
 * Optional outputs (NULL to skip): H (maxit+1) x maxit column-major, beta, X_hist n x maxit
 * (iterate after every iteration), x_valid (0 when the reference leaves x unassigned, AB :25 at
 * k = 1), t_iter[k-1] = seconds from loop entry to the end of iteration k.
 * solve 0: Arnoldi only — operator + orthogonalisation + normalisation (:19-26), no projected
 * solve / iterate / histories (the quantity bench.py's `value` measures on the device). */
int hgo_hybrid_rtp(int kind, int orth, int solve, int64_t m, int64_t n, const int64_t* Ap, const int32_t* Ai,
                   const double* Ax, const int64_t* Bp, const int32_t* Bi, const double* Bx, const double* b,
                   const double* x_true, double tol, int maxit, double lambda, double* x, double* error_norm,
                   double* residual_norm, double* H_out, double* beta_out, double* X_hist, int* x_valid,
                   double* t_iter) {
    csr_t A = {m, n, Ap, Ai, Ax}, B = {n, m, Bp, Bi, Bx};
    const int ldh = maxit + 1;
    double* Q = (double*)calloc((size_t)n * (maxit + 1), sizeof(double));
    double* H = (double*)calloc((size_t)ldh * maxit, sizeof(double));
    double* W = (kind == 0 && solve) ? (double*)calloc((size_t)m * maxit, sizeof(double)) : NULL; /* A*Q(:,1:k) */
    double* G = (kind == 0 && solve) ? (double*)calloc((size_t)maxit * maxit, sizeof(double)) : NULL;
    double* Gk = (kind == 0 && solve) ? (double*)malloc(sizeof(double) * maxit * maxit) : NULL;
    double* gb = (double*)calloc(maxit, sizeof(double));
    double* u = (double*)malloc(sizeof(double) * m);
    double* v = (double*)malloc(sizeof(double) * n);
    double* r = (double*)malloc(sizeof(double) * m);
    double* y = (double*)malloc(sizeof(double) * maxit);
    double* tmp = (double*)malloc(sizeof(double) * (maxit + 1));
    int have_x = 0;
    if (x && kind == 1) { memset(x, 0, sizeof(double) * n); have_x = 1; }  /* BA :4 */
    spmv(&B, b, v);                                        /* :7, r0 = d - M_reg(0) = d */
    const double beta = sqrt(dot(v, v, n));                /* :10 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) Q[i] = v[i] / beta;    /* :13 */
    double nb = 0.0, nxt = 0.0;
    if (solve) {
        nb = sqrt(dot(b, b, m));
        nxt = sqrt(dot(x_true, x_true, n));
        for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    }
    const double t0 = now_s();
    int k;
    for (k = 1; k <= maxit; ++k) {
        const double* q = Q + (size_t)(k - 1) * n;
        double* hcol = H + (size_t)(k - 1) * ldh;
        spmv(&A, q, u);                                    /* :19 */
        spmv(&B, u, v);
        axpy(lambda, q, v, n);
        orthogonalise(orth, Q, n, k, v, hcol, tmp);        /* :20-23 */
        const double hk = sqrt(dot(v, v, n));              /* :24 */
        hcol[k] = hk;
        if (hk == 0.0) break;                              /* :25 */
        double* qn = Q + (size_t)k * n;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) qn[i] = v[i] / hk; /* :26 */
        if (solve) {
            if (kind == 1) {
                lstsq_hess(H, ldh, k, beta, y);            /* BA :28-29 */
            } else {
                /* AB :31-32  AQk = A*Qk (column k is u; the earlier ones are unchanged);
                 * yk = (AQk'*AQk + lambda*eye(k)) \ (AQk'*b) */
                double* wk = W + (size_t)(k - 1) * m;
                memcpy(wk, u, sizeof(double) * m);
                for (int j = 0; j < k; ++j) {
                    const double g = dot(W + (size_t)j * m, wk, m);
                    G[(size_t)(k - 1) * maxit + j] = g;
                    G[(size_t)j * maxit + (k - 1)] = g;
                }
                gb[k - 1] = dot(wk, b, m);
                for (int j = 0; j < k; ++j)
                    for (int i = 0; i < k; ++i) Gk[(size_t)j * k + i] = G[(size_t)j * maxit + i] + (i == j ? lambda : 0.0);
                solve_sym(k, Gk, gb, y);
            }
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) {              /* x = Qk*yk  (BA :30 / AB :33) */
                double s = 0.0;
                for (int j = 0; j < k; ++j) s += Q[(size_t)j * n + i] * y[j];
                x[i] = s;
            }
            have_x = 1;
            if (X_hist) memcpy(X_hist + (size_t)(k - 1) * n, x, sizeof(double) * n);
            spmv(&A, x, r);                                /* BA :32 / AB :35 */
            double rs = 0.0, es = 0.0;
#pragma omp parallel for reduction(+ : rs) schedule(static)
            for (int64_t i = 0; i < m; ++i) rs += (b[i] - r[i]) * (b[i] - r[i]);
#pragma omp parallel for reduction(+ : es) schedule(static)
            for (int64_t i = 0; i < n; ++i) es += (x[i] - x_true[i]) * (x[i] - x_true[i]);
            residual_norm[k - 1] = sqrt(rs) / nb;
            error_norm[k - 1] = sqrt(es) / nxt;            /* :33 / :36 */
        }
        if (t_iter) t_iter[k - 1] = now_s() - t0;
        if (solve && residual_norm[k - 1] <= tol) break;   /* :35 / :38 */
    }
    if (k > maxit) k = maxit;
    if (H_out) memcpy(H_out, H, sizeof(double) * ldh * maxit);
    if (beta_out) *beta_out = beta;
    if (x_valid) *x_valid = have_x;
    free(Q); free(H); free(W); free(G); free(Gk); free(gb); free(u); free(v); free(r); free(y); free(tmp);
    return k;
}

/* kept for callers of the round-1 entry point */
int hgo_hybrid_ba_gmres_rtp(int64_t m, int64_t n, const int64_t* Ap, const int32_t* Ai, const double* Ax,
                            const int64_t* Bp, const int32_t* Bi, const double* Bx, const double* b,
                            const double* x_true, double tol, int maxit, double lambda, double* x,
                            double* error_norm, double* residual_norm) {
    return hgo_hybrid_rtp(1, 0, 1, m, n, Ap, Ai, Ax, Bp, Bi, Bx, b, x_true, tol, maxit, lambda, x, error_norm,
                          residual_norm, NULL, NULL, NULL, NULL, NULL);
}
