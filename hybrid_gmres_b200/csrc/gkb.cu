// Golub-Kahan bidiagonalisation solvers on the device:
//   hybrid_lsqr_solver.m, hybrid_lsmr_solver.m, lsqr_solver.m, lsmr_solver.m
// A' is an explicit CSR matrix (MATLAB's CSC of A already is CSR of A'), so the
// transposed product is the same row-per-warp SpMV as the forward one — no
// atomics.  Scalars (Givens / LSMR recurrences) stay on the host.
#include <algorithm>

#include "common.cuh"
#include "dense_host.h"

namespace {

struct DBuf {
    double* p = nullptr;
    ~DBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t n) {
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double));
        if (e != cudaSuccess) {
            hg_set_error("device allocation of %zu doubles failed: %s", n, cudaGetErrorString(e));
            return HG_ERR_NOMEM;
        }
        return HG_OK;
    }
};

struct MatHolder {
    hg_matrix* m = nullptr;
    ~MatHolder() { hg_matrix_destroy(m); }
};

// Shared plumbing for the four solvers.
struct Gkb {
    hg_ctx* ctx;
    const hg_matrix* A;
    const hg_matrix* At;
    MatHolder own_at;
    int64_t m, n;
    DBuf b, xt, stat;
    double norm_b = 0, norm_xt = 0;
    bool have_xt = false;
    double* hs;  // pinned scalars
    double* ds;  // device scalars

    int init(hg_ctx* c, const hg_matrix* A_, const hg_matrix* At_, const double* hb, const double* hxt) {
        ctx = c;
        A = A_;
        m = A->rows;
        n = A->cols;
        hs = ctx->h_scalars;
        ds = ctx->d_scalars;
        if (At_) {
            HG_REQUIRE(At_->rows == n && At_->cols == m, "GKB: At must be the n x m transpose of A");
            At = At_;
        } else {
            HG_TRY(hg_transpose_device(ctx, A, &own_at.m));
            At = own_at.m;
        }
        HG_TRY(b.alloc((size_t)m));
        HG_TRY(xt.alloc((size_t)n));
        HG_TRY(stat.alloc((size_t)(m + n) / 4 + 4096));
        HG_CUDA(cudaMemcpyAsync(b.p, hb, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        double t = 0;
        HG_TRY(hg_norm2_sync(ctx, b.p, m, &t));
        norm_b = std::sqrt(t);
        if (hxt) {
            have_xt = true;
            HG_CUDA(cudaMemcpyAsync(xt.p, hxt, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
            HG_TRY(hg_norm2_sync(ctx, xt.p, n, &t));
            norm_xt = std::sqrt(t);
        }
        return HG_OK;
    }
    // sqrt(sum of np partials at stat.p) -> device slot + host value (synchronises)
    int norm_from_stat(int np, int slot, double* host) {
        HG_TRY(hg_k_reduce(ctx, stat.p, np, 1, ds + slot, false, nullptr, true));
        HG_CUDA(cudaMemcpyAsync(hs + slot, ds + slot, 8, cudaMemcpyDeviceToHost, ctx->stream));
        HG_CUDA(cudaStreamSynchronize(ctx->stream));
        *host = hs[slot];
        return HG_OK;
    }
    // ||b - A x|| / ||b||  (true residual, e.g. hybrid_lsqr_solver.m:43); optional r store
    int residual(const double* x, double* r_out, double* rn) {
        hg_spmv_epilogue ep;
        ep.alpha = -1.0;
        ep.z1 = b.p;
        ep.g1 = 1.0;
        ep.stat = stat.p;
        int np = 0;
        HG_TRY(hg_k_spmv(ctx, A, x, r_out, ep, &np));
        return norm_from_stat(np, 10, rn);
    }
};

inline void swap_ptr(double*& a, double*& b) {
    double* t = a;
    a = b;
    b = t;
}

}  // namespace

// ---------------------------------------------------------------------------
// hybrid_lsqr_solver.m:1-52 — stacked operator [A; sqrt(lambda) I] never formed
// ---------------------------------------------------------------------------
extern "C" int hg_hybrid_lsqr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At,
                                     const double* b, const double* x_true, double tol, int maxit,
                                     double lambda, double* x, double* error_norm,
                                     double* residual_norm, int* niters, hg_extras* extras) {
    HG_REQUIRE(ctx && A && b && x_true && x && error_norm && residual_norm && niters,
               "hg_hybrid_lsqr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_hybrid_lsqr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    Gkb g;
    HG_TRY(g.init(ctx, A, At, b, x_true));
    const int64_t m = g.m, n = g.n;
    const double sl = std::sqrt(lambda);  // :5
    DBuf bu1, bu2, bt1, bt2, bv, bt3, bw, bx;
    HG_TRY(bu1.alloc(m)); HG_TRY(bu2.alloc(n)); HG_TRY(bt1.alloc(m)); HG_TRY(bt2.alloc(n));
    HG_TRY(bv.alloc(n)); HG_TRY(bt3.alloc(n)); HG_TRY(bw.alloc(n)); HG_TRY(bx.alloc(n));
    double *u1 = bu1.p, *u2 = bu2.p, *t1 = bt1.p, *t2 = bt2.p, *v = bv.p, *t3 = bt3.p, *w = bw.p, *dx = bx.p;
    cudaStream_t st = ctx->stream;
    HG_CUDA(cudaMemsetAsync(dx, 0, (size_t)n * 8, st));
    HG_CUDA(cudaMemsetAsync(u2, 0, (size_t)n * 8, st));
    // beta = norm(b_aug) ; u = b_aug/beta                       (:9-10)
    double beta_aug = 0, alpha_aug = 0;
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, g.b.p, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta_aug));
    HG_CUDA(cudaMemcpyAsync(u1, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_scale_div(ctx, u1, m, g.ds + 5));
    // v_hat = A_aug'*u ; alpha ; v                                (:11-13)
    {
        hg_spmv_epilogue ep;
        ep.z1 = u2;
        ep.g1 = sl;
        ep.stat = g.stat.p;
        HG_TRY(hg_k_spmv(ctx, g.At, u1, v, ep, &np));
        HG_TRY(g.norm_from_stat(np, 6, &alpha_aug));
        HG_TRY(hg_k_scale_div(ctx, v, n, g.ds + 6));
    }
    HG_CUDA(cudaMemcpyAsync(w, v, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));  // :14
    double phi_bar = beta_aug, rho_bar = alpha_aug;
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    if (extras && extras->aux) {
        extras->aux[0] = alpha_aug;
        extras->aux[maxit + 1] = beta_aug;
    }
    int k;
    for (k = 1; k <= maxit; ++k) {
        // u_hat = A_aug*v - alpha*u                               (:22)
        int np1 = 0, np2 = 0;
        {
            hg_spmv_epilogue ep;
            ep.z1 = u1;
            ep.g1 = -alpha_aug;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, A, v, t1, ep, &np1));
        }
        HG_TRY(hg_k_axpby(ctx, n, sl, v, -alpha_aug, u2, t2, nullptr, g.stat.p + np1, &np2));
        HG_TRY(g.norm_from_stat(np1 + np2, 5, &beta_aug));  // :23
        HG_TRY(hg_k_scale_div(ctx, t1, m, g.ds + 5));       // :24
        HG_TRY(hg_k_scale_div(ctx, t2, n, g.ds + 5));
        swap_ptr(u1, t1);
        swap_ptr(u2, t2);
        // v_hat = A_aug'*u - beta*v                               (:26)
        {
            hg_spmv_epilogue ep;
            ep.z1 = u2;
            ep.g1 = sl;
            ep.z2 = v;
            ep.g2 = -beta_aug;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, g.At, u1, t3, ep, &np));
        }
        HG_TRY(g.norm_from_stat(np, 6, &alpha_aug));  // :27
        HG_TRY(hg_k_scale_div(ctx, t3, n, g.ds + 6));  // :28
        swap_ptr(v, t3);
        // Givens                                                 (:30-37)
        const double rho = std::sqrt(rho_bar * rho_bar + beta_aug * beta_aug);
        const double c = rho_bar / rho;
        const double s = beta_aug / rho;
        const double theta = s * alpha_aug;
        rho_bar = -c * alpha_aug;
        const double phi = c * phi_bar;
        phi_bar = s * phi_bar;
        // x, w updates + error                                   (:39-42)
        HG_TRY(hg_k_lsqr_update(ctx, n, dx, w, v, phi / rho, theta / rho, g.xt.p, g.stat.p, &np));
        double en = 0, rn = 0;
        HG_TRY(g.norm_from_stat(np, 7, &en));
        HG_TRY(g.residual(dx, nullptr, &rn));  // :43
        error_norm[k - 1] = en / g.norm_xt;
        residual_norm[k - 1] = rn / g.norm_b;
        if (extras && extras->aux) {
            extras->aux[k] = alpha_aug;
            extras->aux[maxit + 1 + k] = beta_aug;
        }
        if (extras && extras->X_hist)
            HG_CUDA(cudaMemcpy(extras->X_hist + (size_t)(k - 1) * n, dx, (size_t)n * 8, cudaMemcpyDeviceToHost));
        if (residual_norm[k - 1] < tol) break;  // :45 strict
    }
    if (k > maxit) k = maxit;
    *niters = k;
    HG_CUDA(cudaMemcpyAsync(x, dx, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaStreamSynchronize(st));
    return HG_OK;
}

// ---------------------------------------------------------------------------
// hybrid_lsmr_solver.m:1-57
// ---------------------------------------------------------------------------
extern "C" int hg_hybrid_lsmr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At,
                                     const double* b, const double* x_true, double tol, int maxit,
                                     double lambda, double* x, double* error_norm,
                                     double* residual_norm, int* niters, hg_extras* extras) {
    HG_REQUIRE(ctx && A && b && x_true && x && error_norm && residual_norm && niters,
               "hg_hybrid_lsmr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_hybrid_lsmr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    Gkb g;
    HG_TRY(g.init(ctx, A, At, b, x_true));
    const int64_t m = g.m, n = g.n;
    const int64_t ldv = (n + 31) / 32 * 32;
    DBuf bu, bt, bV, bx, by;
    HG_TRY(bu.alloc(m)); HG_TRY(bt.alloc(m)); HG_TRY(bV.alloc((size_t)ldv * maxit));
    HG_TRY(bx.alloc(n)); HG_TRY(by.alloc(maxit));
    double *u = bu.p, *t = bt.p, *V = bV.p, *dx = bx.p;
    cudaStream_t st = ctx->stream;
    HG_CUDA(cudaMemsetAsync(dx, 0, (size_t)n * 8, st));
    std::vector<double> Bk((size_t)(maxit + 1) * maxit, 0.0);  // (maxit+1) x maxit col-major
    const int ldb = maxit + 1;
    double beta1 = 0, alpha1 = 0;
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, g.b.p, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta1));  // :7
    HG_CUDA(cudaMemcpyAsync(u, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_scale_div(ctx, u, m, g.ds + 5));  // :8
    {
        hg_spmv_epilogue ep;
        ep.stat = g.stat.p;
        HG_TRY(hg_k_spmv(ctx, g.At, u, V, ep, &np));  // :13
        HG_TRY(g.norm_from_stat(np, 6, &alpha1));
        HG_TRY(hg_k_scale_div(ctx, V, n, g.ds + 6));  // :15-16
    }
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    std::vector<double> T, LHS, RHS, y;
    int k;
    for (k = 1; k <= maxit; ++k) {
        double* v = V + (size_t)(k - 1) * ldv;
        Bk[(size_t)(k - 1) * ldb + (k - 1)] = alpha1;  // :23
        double beta_k = 0;
        {
            hg_spmv_epilogue ep;
            ep.z1 = u;
            ep.g1 = -alpha1;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, A, v, t, ep, &np));  // :24
            HG_TRY(g.norm_from_stat(np, 5, &beta_k));
            HG_TRY(hg_k_scale_div(ctx, t, m, g.ds + 5));  // :26
            swap_ptr(u, t);
        }
        Bk[(size_t)(k - 1) * ldb + k] = beta_k;  // :27
        if (k < maxit) {                          // :29-35
            double* vn = V + (size_t)k * ldv;
            hg_spmv_epilogue ep;
            ep.z1 = v;
            ep.g1 = -beta_k;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, g.At, u, vn, ep, &np));
            double a_next = 0;
            HG_TRY(g.norm_from_stat(np, 6, &a_next));
            HG_TRY(hg_k_scale_div(ctx, vn, n, g.ds + 6));
            alpha1 = a_next;
        }
        // projected problem on the host                           (:37-44)
        const double alpha_k1 = alpha1, beta_k1 = beta_k;
        T.assign((size_t)k * k, 0.0);
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < k; ++i) {
                double acc = 0.0;
                for (int r = 0; r <= k; ++r) acc += Bk[(size_t)i * ldb + r] * Bk[(size_t)j * ldb + r];
                T[(size_t)j * k + i] = acc;
            }
        LHS.assign((size_t)k * k, 0.0);
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < k; ++i) {
                double acc = 0.0;
                for (int r = 0; r < k; ++r) acc += T[(size_t)r * k + i] * T[(size_t)j * k + r];
                LHS[(size_t)j * k + i] = acc;
            }
        LHS[0] += (alpha_k1 * beta_k1) * (alpha_k1 * beta_k1);
        for (int i = 0; i < k; ++i) LHS[(size_t)i * k + i] += lambda;
        RHS.assign(k, 0.0);
        for (int i = 0; i < k; ++i) RHS[i] = Bk[0] * beta1 * T[i];  // B_k(1,1)*beta1*(T*e1)
        y.assign(k, 0.0);
        hgd::solve_square(k, LHS.data(), k, RHS.data(), y.data());
        HG_CUDA(cudaMemcpyAsync(by.p, y.data(), (size_t)k * 8, cudaMemcpyHostToDevice, st));
        HG_CUDA(cudaStreamSynchronize(st));  // y is pageable host memory
        // x = V(:,1:k)*yk + error                                 (:45,47)
        HG_TRY(hg_k_lincomb(ctx, V, ldv, n, k, by.p, 1.0, nullptr, dx, g.xt.p, g.stat.p, &np));
        double en = 0, rn = 0;
        HG_TRY(g.norm_from_stat(np, 7, &en));
        HG_TRY(g.residual(dx, nullptr, &rn));  // :48
        error_norm[k - 1] = en / g.norm_xt;
        residual_norm[k - 1] = rn / g.norm_b;
        if (extras && extras->X_hist)
            HG_CUDA(cudaMemcpy(extras->X_hist + (size_t)(k - 1) * n, dx, (size_t)n * 8, cudaMemcpyDeviceToHost));
        if (residual_norm[k - 1] <= tol) break;  // :50
    }
    if (k > maxit) k = maxit;
    *niters = k;
    if (extras && extras->aux)
        for (int j = 0; j < maxit; ++j) {
            extras->aux[j] = Bk[(size_t)j * ldb + j];
            extras->aux[maxit + 1 + j] = Bk[(size_t)j * ldb + j + 1];
        }
    HG_CUDA(cudaMemcpyAsync(x, dx, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaStreamSynchronize(st));
    return HG_OK;
}

// ---------------------------------------------------------------------------
// lsqr_solver.m:1-54
// ---------------------------------------------------------------------------
extern "C" int hg_lsqr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                              const double* x_true, double tol, int maxit, double* x,
                              double* error_norm, double* residual_norm, int* niters,
                              hg_extras* extras) {
    HG_REQUIRE(ctx && A && b && x_true && x && error_norm && residual_norm && niters,
               "hg_lsqr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_lsqr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    Gkb g;
    HG_TRY(g.init(ctx, A, At, b, x_true));
    const int64_t m = g.m, n = g.n;
    DBuf bu, bt, bv, bt3, bw, bx;
    HG_TRY(bu.alloc(m)); HG_TRY(bt.alloc(m)); HG_TRY(bv.alloc(n)); HG_TRY(bt3.alloc(n));
    HG_TRY(bw.alloc(n)); HG_TRY(bx.alloc(n));
    double *u = bu.p, *t = bt.p, *v = bv.p, *t3 = bt3.p, *w = bw.p, *dx = bx.p;
    cudaStream_t st = ctx->stream;
    HG_CUDA(cudaMemsetAsync(dx, 0, (size_t)n * 8, st));
    double beta = 0, alpha = 0;
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, g.b.p, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta));  // :7
    HG_CUDA(cudaMemcpyAsync(u, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_scale_div(ctx, u, m, g.ds + 5));  // :8
    {
        hg_spmv_epilogue ep;
        ep.stat = g.stat.p;
        HG_TRY(hg_k_spmv(ctx, g.At, u, v, ep, &np));  // :10
        HG_TRY(g.norm_from_stat(np, 6, &alpha));
        HG_TRY(hg_k_scale_div(ctx, v, n, g.ds + 6));  // :12
    }
    HG_CUDA(cudaMemcpyAsync(w, v, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    double phi_bar = beta, rho_bar = alpha;
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    int k;
    for (k = 1; k <= maxit; ++k) {
        {
            hg_spmv_epilogue ep;
            ep.z1 = u;
            ep.g1 = -alpha;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, A, v, t, ep, &np));  // :22
            HG_TRY(g.norm_from_stat(np, 5, &beta));
            HG_TRY(hg_k_scale_div(ctx, t, m, g.ds + 5));
            swap_ptr(u, t);
        }
        {
            hg_spmv_epilogue ep;
            ep.z1 = v;
            ep.g1 = -beta;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, g.At, u, t3, ep, &np));  // :26
            HG_TRY(g.norm_from_stat(np, 6, &alpha));
            HG_TRY(hg_k_scale_div(ctx, t3, n, g.ds + 6));
            swap_ptr(v, t3);
        }
        const double rho = std::sqrt(rho_bar * rho_bar + beta * beta);  // :31
        const double c = rho_bar / rho;
        const double s = beta / rho;
        const double theta = s * alpha;
        rho_bar = -c * alpha;
        const double phi = c * phi_bar;
        phi_bar = s * phi_bar;
        HG_TRY(hg_k_lsqr_update(ctx, n, dx, w, v, phi / rho, theta / rho, g.xt.p, g.stat.p, &np));  // :40-41
        double en = 0;
        HG_TRY(g.norm_from_stat(np, 7, &en));
        error_norm[k - 1] = en / g.norm_xt;                  // :43
        residual_norm[k - 1] = std::fabs(phi_bar) / g.norm_b;  // :44
        if (extras && extras->X_hist)
            HG_CUDA(cudaMemcpy(extras->X_hist + (size_t)(k - 1) * n, dx, (size_t)n * 8, cudaMemcpyDeviceToHost));
        if (residual_norm[k - 1] <= tol) break;  // :46
    }
    if (k > maxit) k = maxit;
    *niters = k;
    double rn = 0;
    HG_TRY(g.residual(dx, nullptr, &rn));
    residual_norm[k - 1] = rn / g.norm_b;  // :52
    HG_CUDA(cudaMemcpyAsync(x, dx, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaStreamSynchronize(st));
    return HG_OK;
}

// ---------------------------------------------------------------------------
// lsmr_solver.m:1-83
// ---------------------------------------------------------------------------
extern "C" int hg_lsmr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                              const double* x_true, double tol, int maxit, double* x,
                              double* err_hist, double* res_hist, double* ar_hist, int* iters,
                              hg_extras* extras) {
    HG_REQUIRE(ctx && A && b && x && err_hist && res_hist && ar_hist && iters,
               "hg_lsmr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_lsmr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    const double eps = 2.220446049250313e-16;
    Gkb g;
    HG_TRY(g.init(ctx, A, At, b, x_true));
    const int64_t m = g.m, n = g.n;
    DBuf bu, bt, bv, bt3, bh, bhb, bx, br, bar;
    HG_TRY(bu.alloc(m)); HG_TRY(bt.alloc(m)); HG_TRY(bv.alloc(n)); HG_TRY(bt3.alloc(n));
    HG_TRY(bh.alloc(n)); HG_TRY(bhb.alloc(n)); HG_TRY(bx.alloc(n)); HG_TRY(br.alloc(m));
    double *u = bu.p, *t = bt.p, *v = bv.p, *t3 = bt3.p, *h = bh.p, *hbar = bhb.p, *dx = bx.p, *r = br.p;
    cudaStream_t st = ctx->stream;
    HG_CUDA(cudaMemsetAsync(dx, 0, (size_t)n * 8, st));
    HG_CUDA(cudaMemsetAsync(hbar, 0, (size_t)n * 8, st));
    // norm(A,'fro'): recomputed every iteration by the reference (:71), constant here
    double fro2 = 0;
    HG_TRY(hg_norm2_sync(ctx, A->vals, A->nnz, &fro2));
    const double normA = std::sqrt(fro2);
    double beta = 0, alpha = 0;
    int np = 0;
    HG_CUDA(cudaMemcpyAsync(u, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_sumsq(ctx, u, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta));                          // :11
    if (beta > 0) HG_TRY(hg_k_scale_div(ctx, u, m, g.ds + 5));       // :12
    {
        hg_spmv_epilogue ep;
        ep.stat = g.stat.p;
        HG_TRY(hg_k_spmv(ctx, g.At, u, v, ep, &np));                 // :14
        HG_TRY(g.norm_from_stat(np, 6, &alpha));
        if (alpha > 0) HG_TRY(hg_k_scale_div(ctx, v, n, g.ds + 6));  // :16
    }
    double zetabar = alpha * beta, alphabar = alpha, rho = 1, rhobar = 1, cbar = 1, sbar = 0;  // :19-23
    HG_CUDA(cudaMemcpyAsync(h, v, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));              // :25
    const double nan = std::nan("");
    for (int i = 0; i < maxit; ++i) {
        err_hist[i] = nan;
        res_hist[i] = 0.0;
        ar_hist[i] = 0.0;
    }
    int k;
    for (k = 1; k <= maxit; ++k) {
        {
            hg_spmv_epilogue ep;
            ep.z1 = u;
            ep.g1 = -alpha;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, A, v, t, ep, &np));  // :34
            HG_TRY(g.norm_from_stat(np, 5, &beta));
            if (beta > 0) HG_TRY(hg_k_scale_div(ctx, t, m, g.ds + 5));  // :36
            swap_ptr(u, t);
        }
        {
            hg_spmv_epilogue ep;
            ep.z1 = v;
            ep.g1 = -beta;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, g.At, u, t3, ep, &np));  // :38
            HG_TRY(g.norm_from_stat(np, 6, &alpha));
            if (alpha > 0) HG_TRY(hg_k_scale_div(ctx, t3, n, g.ds + 6));  // :40
            swap_ptr(v, t3);
        }
        const double alphahat = alphabar;  // :42-49
        const double rhoold = rho;
        rho = std::hypot(alphahat, beta);
        const double c = alphahat / rho;
        const double s = beta / rho;
        const double thetanew = s * alpha;
        alphabar = c * alpha;
        const double rhobarold = rhobar;  // :51-55
        const double thetabar = sbar * rho;
        rhobar = std::hypot(cbar * rho, thetanew);
        cbar = (cbar * rho) / rhobar;
        sbar = thetanew / rhobar;
        const double zeta = cbar * zetabar;  // :58-59
        zetabar = -sbar * zetabar;
        const double c0 = (k == 1) ? 0.0 : (thetabar * rho) / (rhoold * rhobarold);  // :64
        HG_TRY(hg_k_lsmr_update(ctx, n, dx, h, hbar, v, k == 1 ? 1 : 0, c0, zeta / (rho * rhobar),
                                thetanew / rho, g.have_xt ? g.xt.p : nullptr, g.stat.p, &np));  // :61-67
        double en = 0, rn = 0, arn = 0;
        HG_TRY(g.norm_from_stat(np, 7, &en));
        HG_TRY(g.residual(dx, r, &rn));  // :69
        {
            hg_spmv_epilogue ep;
            ep.stat = g.stat.p;
            HG_TRY(hg_k_spmv(ctx, g.At, r, nullptr, ep, &np));  // norm(A.'*r)  :71
            HG_TRY(g.norm_from_stat(np, 8, &arn));
        }
        res_hist[k - 1] = rn / (g.norm_b + eps);                      // :70
        ar_hist[k - 1] = arn / (normA * std::max(rn, eps));           // :71
        if (g.have_xt) err_hist[k - 1] = en / g.norm_xt;              // :72-74
        if (extras && extras->X_hist)
            HG_CUDA(cudaMemcpy(extras->X_hist + (size_t)(k - 1) * n, dx, (size_t)n * 8, cudaMemcpyDeviceToHost));
        if (res_hist[k - 1] < tol) break;  // :76
    }
    if (k > maxit) k = maxit;
    *iters = k;
    HG_CUDA(cudaMemcpyAsync(x, dx, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaStreamSynchronize(st));
    return HG_OK;
}
