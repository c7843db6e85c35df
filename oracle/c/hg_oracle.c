/*
 * hg_oracle.c — multi-threaded C restatement of hybrid_ba_gmres_rtp.m (TEST INFRASTRUCTURE).
 *
 * PARITY UNPINNED (see oracle/__init__.py): the reference cannot be executed here.  This file
 * exists so that the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) uses every
 * host core: it is the same literal algorithm as oracle/solvers.py:hybrid_ba_gmres_rtp with
 * OpenMP-parallel sparse mat-vecs and BLAS-1 loops.  It is checked against the NumPy oracle in
 * tests/test_oracle_c.py and is never imported by the product.
 *
 * Statement map (hybrid_ba_gmres_rtp.m):
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

/* returns niters; x (n), error_norm / residual_norm (maxit, first niters valid) */
int hgo_hybrid_ba_gmres_rtp(int64_t m, int64_t n, const int64_t* Ap, const int32_t* Ai, const double* Ax,
                            const int64_t* Bp, const int32_t* Bi, const double* Bx, const double* b,
                            const double* x_true, double tol, int maxit, double lambda, double* x,
                            double* error_norm, double* residual_norm) {
    csr_t A = {m, n, Ap, Ai, Ax}, B = {n, m, Bp, Bi, Bx};
    const int ldh = maxit + 1;
    double* Q = (double*)calloc((size_t)n * (maxit + 1), sizeof(double));
    double* H = (double*)calloc((size_t)ldh * maxit, sizeof(double));
    double* u = (double*)malloc(sizeof(double) * m);
    double* v = (double*)malloc(sizeof(double) * n);
    double* r = (double*)malloc(sizeof(double) * m);
    double* y = (double*)malloc(sizeof(double) * maxit);
    memset(x, 0, sizeof(double) * n);                      /* :4 */
    spmv(&B, b, v);                                        /* :7, r0 = d - M_reg(0) = d */
    const double beta = sqrt(dot(v, v, n));                /* :10 */
    for (int64_t i = 0; i < n; ++i) Q[i] = v[i] / beta;    /* :13 */
    const double nb = sqrt(dot(b, b, m)), nxt = sqrt(dot(x_true, x_true, n));
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    int k;
    for (k = 1; k <= maxit; ++k) {
        const double* q = Q + (size_t)(k - 1) * n;
        spmv(&A, q, u);                                    /* :19 */
        spmv(&B, u, v);
        axpy(lambda, q, v, n);
        for (int j = 0; j < k; ++j) {                      /* :20-23 MGS */
            const double* qj = Q + (size_t)j * n;
            const double h = dot(qj, v, n);
            H[(size_t)(k - 1) * ldh + j] = h;
            axpy(-h, qj, v, n);
        }
        const double hk = sqrt(dot(v, v, n));              /* :24 */
        H[(size_t)(k - 1) * ldh + k] = hk;
        if (hk == 0.0) break;                              /* :25 */
        double* qn = Q + (size_t)k * n;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) qn[i] = v[i] / hk; /* :26 */
        lstsq_hess(H, ldh, k, beta, y);                    /* :28-29 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {                  /* :30 */
            double s = 0.0;
            for (int j = 0; j < k; ++j) s += Q[(size_t)j * n + i] * y[j];
            x[i] = s;
        }
        spmv(&A, x, r);                                    /* :32 */
        double rs = 0.0, es = 0.0;
#pragma omp parallel for reduction(+ : rs) schedule(static)
        for (int64_t i = 0; i < m; ++i) rs += (b[i] - r[i]) * (b[i] - r[i]);
#pragma omp parallel for reduction(+ : es) schedule(static)
        for (int64_t i = 0; i < n; ++i) es += (x[i] - x_true[i]) * (x[i] - x_true[i]);
        residual_norm[k - 1] = sqrt(rs) / nb;
        error_norm[k - 1] = sqrt(es) / nxt;                /* :33 */
        if (residual_norm[k - 1] <= tol) break;            /* :35 */
    }
    if (k > maxit) k = maxit;
    free(Q); free(H); free(u); free(v); free(r); free(y);
    return k;
}
