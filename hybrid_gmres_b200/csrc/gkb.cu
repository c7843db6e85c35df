// Golub-Kahan bidiagonalisation solvers on the device:
//   hybrid_lsqr_solver.m, hybrid_lsmr_solver.m, lsqr_solver.m, lsmr_solver.m
// A' is an explicit CSR matrix (MATLAB's CSC of A already is CSR of A'), so the
// transposed product is the same row-per-warp SpMV as the forward one — no
// atomics.  Scalars (Givens / LSMR recurrences) stay on the host.
//
// One implementation serves one GPU and P GPUs (GKB row): with a
// communicator, A is this rank's detector-row block A_p (m_p x n), A' its transpose
// (n x m_p); m-vectors (u, b, r) are local blocks, n-vectors (v, w, x, h) are equal row
// slices.  A_p*v needs the all-gathered v; A_p'*u_p is a partial n-vector that is
// reduce-scattered; norms are all-reduced.  Without a communicator every collective
// disappears and the epilogues stay fused in the SpMV exactly as before.
#include <algorithm>

#include "common.cuh"
#include "dense_host.h"

namespace {

struct DBuf {
    double* p = nullptr;
    ~DBuf() { hg_dfree(p); }
    int alloc(size_t n, cudaStream_t st = nullptr, bool zero = false) {
        cudaError_t e = hg_dmalloc_cur((void**)&p, std::max<size_t>(n, 1) * sizeof(double));
        if (e != cudaSuccess) {
            hg_set_error("device allocation of %zu doubles failed: %s", n, cudaGetErrorString(e));
            return HG_ERR_NOMEM;
        }
        if (zero) cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(double), st);
        return HG_OK;
    }
};

struct MatHolder {
    hg_matrix* m = nullptr;
    ~MatHolder() { hg_matrix_destroy(m); }
};

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// Shared plumbing for the four solvers.
struct Gkb {
    hg_ctx* ctx = nullptr;
    hg_comm* comm = nullptr;  // nullptr: single GPU
    const hg_matrix* A = nullptr;
    const hg_matrix* At = nullptr;
    MatHolder own_at;
    int64_t m = 0;      // local rows of A (m_p)
    int64_t n = 0;      // global n
    int64_t nl = 0;     // length of the n-vector slices held here (n, or n_p zero padded)
    int64_t n_pad = 0;  // nl * P
    int64_t row0 = 0;   // first global row of this rank's slice
    int64_t nvalid = 0; // rows of the slice that exist (rest is padding)
    DBuf b, xt, stat, vfull, part, tmp;
    double norm_b = 0, norm_xt = 0;
    bool have_xt = false;
    double* hs = nullptr;  // pinned scalars
    double* ds = nullptr;  // device scalars
    cudaStream_t st = nullptr;

    int init(hg_ctx* c, hg_comm* cm, const hg_matrix* A_, const hg_matrix* At_, const double* hb,
             const double* hxt) {
        ctx = c;
        comm = cm;
        st = ctx->stream;
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
        if (comm) {
            const int P = hg_comm_size(comm);
            nl = round_up((n + P - 1) / P, 32);
            n_pad = nl * P;
            row0 = (int64_t)hg_comm_rank(comm) * nl;
            nvalid = std::max<int64_t>(0, std::min(n, row0 + nl) - row0);
            HG_TRY(vfull.alloc((size_t)n_pad, st, true));
            HG_TRY(part.alloc((size_t)n_pad, st, true));
            HG_TRY(tmp.alloc((size_t)nl, st, true));
        } else {
            nl = n_pad = nvalid = n;
            row0 = 0;
        }
        HG_TRY(b.alloc((size_t)m));
        HG_TRY(xt.alloc((size_t)nl, st, true));
        HG_TRY(stat.alloc((size_t)(m + n_pad) / 4 + 4096));
        if (m) HG_CUDA(cudaMemcpyAsync(b.p, hb, (size_t)m * 8, cudaMemcpyHostToDevice, st));
        int np = 0;
        HG_TRY(hg_k_sumsq(ctx, b.p, m, stat.p, &np));
        HG_TRY(norm_from_stat(np, 11, &norm_b));
        if (hxt) {
            have_xt = true;
            if (nvalid) HG_CUDA(cudaMemcpyAsync(xt.p, hxt + row0, (size_t)nvalid * 8, cudaMemcpyHostToDevice, st));
            HG_TRY(hg_k_sumsq(ctx, xt.p, nl, stat.p, &np));
            HG_TRY(norm_from_stat(np, 12, &norm_xt));
        }
        return HG_OK;
    }
    // global sqrt(sum of np local partials at stat.p) -> device slot + host value (synchronises)
    int norm_from_stat(int np, int slot, double* host) {
        if (comm) {
            HG_TRY(hg_k_reduce(ctx, stat.p, np, 1, ds + 32, false, nullptr, false));
            HG_TRY(hg_comm_allreduce(comm, ds + 32, 1, st));
            HG_TRY(hg_k_reduce(ctx, ds + 32, 1, 1, ds + slot, false, nullptr, true));
        } else {
            HG_TRY(hg_k_reduce(ctx, stat.p, np, 1, ds + slot, false, nullptr, true));
        }
        HG_CUDA(cudaMemcpyAsync(hs + slot, ds + slot, 8, cudaMemcpyDeviceToHost, st));
        HG_CUDA(cudaStreamSynchronize(st));
        *host = hs[slot];
        return HG_OK;
    }
    // ---- lagged histories ---------------------------------------------------------------------------
    // The error / residual norms of iteration k are queued without a host synchronisation (device slots
    // 20..22 -> pinned ring) and read after the first synchronisation of iteration k+1, which follows them in
    // the stream: an iteration synchronises twice (beta, alpha) instead of four or five times.  The stop rule
    // of iteration k is therefore applied one half-step later; x still holds x_k at that point.
    int enqueue_norm(int np, int dslot) {  // sqrt of the global sum of np partials at stat.p -> ds[dslot]
        if (comm) {
            HG_TRY(hg_k_reduce(ctx, stat.p, np, 1, ds + 33, false, nullptr, false));
            HG_TRY(hg_comm_allreduce(comm, ds + 33, 1, st));
            return hg_k_reduce(ctx, ds + 33, 1, 1, ds + dslot, false, nullptr, true);
        }
        return hg_k_reduce(ctx, stat.p, np, 1, ds + dslot, false, nullptr, true);
    }
    int residual_enqueue(const double* x, double* r_out, int dslot) {  // ||b - A x|| -> ds[dslot]
        int np = 0;
        HG_TRY(apply_A(x, r_out, -1.0, b.p, 1.0, nullptr, true, 0, &np));
        return enqueue_norm(np, dslot);
    }
    int hist_copy(int k) {  // ds[20..23) -> ring slot of iteration k
        HG_CUDA(cudaMemcpyAsync(hs + 40 + (k & 1) * 4, ds + 20, 24, cudaMemcpyDeviceToHost, st));
        return HG_OK;
    }
    const double* hist(int k) const { return hs + 40 + (k & 1) * 4; }  // {err, res, ar} of iteration k

    // out(m) = alpha*A*v + g1*z1, v an n-slice; optional stat partials at stat.p (+offset)
    int apply_A(const double* v, double* out, double alpha, const double* z1, double g1, const double* ref,
                bool want_stat, int stat_off, int* np) {
        const double* vin = v;
        if (comm) {
            HG_TRY(hg_comm_allgather(comm, v, vfull.p, (size_t)nl, st));
            vin = vfull.p;
        }
        hg_spmv_epilogue ep;
        ep.alpha = alpha;
        ep.z1 = z1;
        ep.g1 = g1;
        ep.ref = ref;
        ep.stat = want_stat ? stat.p + stat_off : nullptr;
        return hg_k_spmv(ctx, A, vin, out, ep, np);
    }
    // out(n-slice) = A'*u + g1*z1 + g2*z2 (z1, z2 n-slices); stat partials of out^2 at stat.p
    int apply_At(const double* u, double* out, const double* z1, double g1, const double* z2, double g2,
                 bool want_stat, int* np) {
        if (!comm) {
            hg_spmv_epilogue ep;
            ep.z1 = z1;
            ep.g1 = g1;
            ep.z2 = z2;
            ep.g2 = g2;
            ep.stat = want_stat ? stat.p : nullptr;
            return hg_k_spmv(ctx, At, u, out, ep, np);
        }
        hg_spmv_epilogue ep;
        HG_TRY(hg_k_spmv(ctx, At, u, part.p, ep, nullptr));  // rows >= n of `part` stay zero
        double* rs = (z1 || z2 || out == nullptr) ? tmp.p : out;
        HG_TRY(hg_comm_reduce_scatter(comm, part.p, rs, (size_t)nl, st));
        if (z1 && z2) {
            HG_TRY(hg_k_axpby(ctx, nl, 1.0, rs, g1, z1, rs, nullptr, nullptr, nullptr));
            return hg_k_axpby(ctx, nl, 1.0, rs, g2, z2, out, nullptr, want_stat ? stat.p : nullptr, np);
        }
        if (z1) return hg_k_axpby(ctx, nl, 1.0, rs, g1, z1, out, nullptr, want_stat ? stat.p : nullptr, np);
        if (want_stat) return hg_k_sumsq(ctx, rs, nl, stat.p, np);
        if (np) *np = 0;
        return HG_OK;
    }
    // ||b - A x|| (global); optional local residual block r_out
    int residual(const double* x, double* r_out, double* rn) {
        int np = 0;
        HG_TRY(apply_A(x, r_out, -1.0, b.p, 1.0, nullptr, true, 0, &np));
        return norm_from_stat(np, 10, rn);
    }
    // full n-vector to the host (all ranks)
    int fetch_x(const double* xs, double* host) {
        if (comm) {
            HG_TRY(hg_comm_allgather(comm, xs, vfull.p, (size_t)nl, st));
            HG_CUDA(cudaMemcpyAsync(host, vfull.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        } else {
            HG_CUDA(cudaMemcpyAsync(host, xs, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        }
        HG_CUDA(cudaStreamSynchronize(st));
        return HG_OK;
    }
};

inline void swap_ptr(double*& a, double*& b) {
    double* t = a;
    a = b;
    b = t;
}

// ---------------------------------------------------------------------------
// hybrid_lsqr_solver.m:1-52 — stacked operator [A; sqrt(lambda) I] never formed
// ---------------------------------------------------------------------------
int hybrid_lsqr_impl(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A, const hg_matrix* At, const double* b,
                     const double* x_true, double tol, int maxit, double lambda, double* x,
                     double* error_norm, double* residual_norm, int* niters, hg_extras* extras) {
    HG_REQUIRE(ctx && A && x_true && x && error_norm && residual_norm && niters,
               "hg_hybrid_lsqr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_hybrid_lsqr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    Gkb g;
    HG_TRY(g.init(ctx, comm, A, At, b, x_true));
    const int64_t m = g.m, nl = g.nl, n = g.n;
    cudaStream_t st = ctx->stream;
    const double sl = std::sqrt(lambda);  // :5
    DBuf bu1, bu2, bt1, bt2, bv, bt3, bw, bx;
    HG_TRY(bu1.alloc(m)); HG_TRY(bu2.alloc(nl, st, true)); HG_TRY(bt1.alloc(m)); HG_TRY(bt2.alloc(nl, st, true));
    HG_TRY(bv.alloc(nl, st, true)); HG_TRY(bt3.alloc(nl, st, true)); HG_TRY(bw.alloc(nl)); HG_TRY(bx.alloc(nl, st, true));
    double *u1 = bu1.p, *u2 = bu2.p, *t1 = bt1.p, *t2 = bt2.p, *v = bv.p, *t3 = bt3.p, *w = bw.p, *dx = bx.p;
    // beta = norm(b_aug) ; u = b_aug/beta                       (:9-10)
    double beta_aug = 0, alpha_aug = 0;
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, g.b.p, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta_aug));
    if (m) HG_CUDA(cudaMemcpyAsync(u1, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_scale_div(ctx, u1, m, g.ds + 5));
    // v_hat = A_aug'*u ; alpha ; v                                (:11-13)
    HG_TRY(g.apply_At(u1, v, u2, sl, nullptr, 0.0, true, &np));
    HG_TRY(g.norm_from_stat(np, 6, &alpha_aug));
    HG_TRY(hg_k_scale_div(ctx, v, nl, g.ds + 6));
    HG_CUDA(cudaMemcpyAsync(w, v, (size_t)nl * 8, cudaMemcpyDeviceToDevice, st));  // :14
    double phi_bar = beta_aug, rho_bar = alpha_aug;
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    if (extras && extras->aux) {
        extras->aux[0] = alpha_aug;
        extras->aux[maxit + 1] = beta_aug;
    }
    // norm(b - A*x) (:43) without a third product per iteration: r_k and d_k = A*w_k follow from the Golub-Kahan
    // relation A*v_k = alpha_k*u_k + beta_{k+1}*u_{k+1} (hg_k_gkb_resid); gkb_residual = 1 forms b - A*x literally
    const bool literal_res = hg_gkb_residual_mode() == 1;
    DBuf bd, br;
    if (!literal_res) {
        HG_TRY(bd.alloc(m));
        HG_TRY(br.alloc(m));
        if (m) HG_CUDA(cudaMemcpyAsync(br.p, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));  // r_0 = b
    }
    double c_prev = 0.0;
    int k, pending = 0;
    bool stopped = false;
    for (k = 1; k <= maxit; ++k) {
        const double alpha_k = alpha_aug;
        // u_hat = A_aug*v - alpha*u                               (:22)
        int np1 = 0, np2 = 0;
        HG_TRY(g.apply_A(v, t1, 1.0, u1, -alpha_aug, nullptr, true, 0, &np1));
        HG_TRY(hg_k_axpby(ctx, nl, sl, v, -alpha_aug, u2, t2, nullptr, g.stat.p + np1, &np2));
        HG_TRY(g.norm_from_stat(np1 + np2, 5, &beta_aug));  // :23
        if (k > 1) {  // histories of iteration k-1 have landed: stop rule of that iteration (:45, strict)
            error_norm[k - 2] = g.hist(k - 1)[0] / g.norm_xt;
            residual_norm[k - 2] = g.hist(k - 1)[1] / g.norm_b;
            pending = 0;
            if (residual_norm[k - 2] < tol) {
                k = k - 1;
                stopped = true;
                break;
            }
        }
        HG_TRY(hg_k_scale_div(ctx, t1, m, g.ds + 5));       // :24
        HG_TRY(hg_k_scale_div(ctx, t2, nl, g.ds + 5));
        swap_ptr(u1, t1);
        swap_ptr(u2, t2);
        // v_hat = A_aug'*u - beta*v                               (:26)
        HG_TRY(g.apply_At(u1, t3, u2, sl, v, -beta_aug, true, &np));
        HG_TRY(g.norm_from_stat(np, 6, &alpha_aug));   // :27
        HG_TRY(hg_k_scale_div(ctx, t3, nl, g.ds + 6));  // :28
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
        HG_TRY(hg_k_lsqr_update(ctx, nl, dx, w, v, phi / rho, theta / rho, g.xt.p, g.stat.p, &np));
        HG_TRY(g.enqueue_norm(np, 20));               // :42
        if (literal_res) {
            HG_TRY(g.residual_enqueue(dx, nullptr, 21));  // :43
        } else {  // t1 holds u_k (top block), u1 holds u_{k+1}
            HG_TRY(hg_k_gkb_resid(ctx, m, t1, alpha_k, u1, beta_aug, bd.p, c_prev, nullptr, 0.0, k == 1, br.p, phi / rho,
                                  g.stat.p, &np));
            HG_TRY(g.enqueue_norm(np, 21));
            c_prev = theta / rho;
        }
        HG_TRY(g.hist_copy(k));
        pending = k;
        if (extras && extras->aux) {
            extras->aux[k] = alpha_aug;
            extras->aux[maxit + 1 + k] = beta_aug;
        }
        if (extras && extras->X_hist) HG_TRY(g.fetch_x(dx, extras->X_hist + (size_t)(k - 1) * n));
    }
    if (!stopped) k = maxit;
    if (pending) {  // the last iteration's histories
        HG_CUDA(cudaStreamSynchronize(st));
        error_norm[pending - 1] = g.hist(pending)[0] / g.norm_xt;
        residual_norm[pending - 1] = g.hist(pending)[1] / g.norm_b;
    }
    *niters = k;
    return g.fetch_x(dx, x);
}

// ---------------------------------------------------------------------------
// hybrid_lsmr_solver.m:1-57
// ---------------------------------------------------------------------------
int hybrid_lsmr_impl(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A, const hg_matrix* At, const double* b,
                     const double* x_true, double tol, int maxit, double lambda, double* x,
                     double* error_norm, double* residual_norm, int* niters, hg_extras* extras) {
    HG_REQUIRE(ctx && A && x_true && x && error_norm && residual_norm && niters,
               "hg_hybrid_lsmr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_hybrid_lsmr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    Gkb g;
    HG_TRY(g.init(ctx, comm, A, At, b, x_true));
    const int64_t m = g.m, nl = g.nl, n = g.n;
    cudaStream_t st = ctx->stream;
    const int64_t ldv = round_up(std::max<int64_t>(nl, 1), 32);
    // norm(b - A*x) (:48) without a third product per iteration: A*V_k = U_{k+1}*B_k (Golub-Kahan relation), so
    // b - A*V_k*y = U_{k+1}*(beta1*e1 - B_k*y): the u vectors are kept (m x (maxit+1)) and the residual is one
    // combination of k+1 columns; gkb_residual = 1 forms b - A*x literally
    const bool literal_res = hg_gkb_residual_mode() == 1;
    const int64_t ldu = round_up(std::max<int64_t>(m, 1), 32);
    DBuf bu, bt, bV, bx, by, bU, bz;
    if (literal_res) {
        HG_TRY(bu.alloc(m));
        HG_TRY(bt.alloc(m));
    } else {
        HG_TRY(bU.alloc((size_t)ldu * (maxit + 1), st, true));
        HG_TRY(bz.alloc(maxit + 1));
    }
    HG_TRY(bV.alloc((size_t)ldv * maxit, st, true));
    HG_TRY(bx.alloc(nl, st, true)); HG_TRY(by.alloc(maxit));
    double *u = literal_res ? bu.p : bU.p, *t = literal_res ? bt.p : bU.p + ldu, *V = bV.p, *dx = bx.p;
    std::vector<double> Bk((size_t)(maxit + 1) * maxit, 0.0);  // (maxit+1) x maxit col-major
    const int ldb = maxit + 1;
    double beta1 = 0, alpha1 = 0;
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, g.b.p, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta1));  // :7
    if (m) HG_CUDA(cudaMemcpyAsync(u, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_scale_div(ctx, u, m, g.ds + 5));  // :8
    HG_TRY(g.apply_At(u, V, nullptr, 0.0, nullptr, 0.0, true, &np));  // :13
    HG_TRY(g.norm_from_stat(np, 6, &alpha1));
    HG_TRY(hg_k_scale_div(ctx, V, nl, g.ds + 6));  // :15-16
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    std::vector<double> T, LHS, RHS;
    struct PinY {
        double* p = nullptr;
        ~PinY() { hg_hfree(p); }
    } hy;
    if (hg_hmalloc_cur((void**)&hy.p, (size_t)2 * (2 * maxit + 1) * sizeof(double)) != cudaSuccess) {
        hg_set_error("hybrid_lsmr: pinned allocation failed");
        return HG_ERR_NOMEM;
    }
    int k, pending = 0;
    bool stopped = false;
    for (k = 1; k <= maxit; ++k) {
        double* v = V + (size_t)(k - 1) * ldv;
        Bk[(size_t)(k - 1) * ldb + (k - 1)] = alpha1;  // :23
        double beta_k = 0;
        HG_TRY(g.apply_A(v, t, 1.0, u, -alpha1, nullptr, true, 0, &np));  // :24
        HG_TRY(g.norm_from_stat(np, 5, &beta_k));
        if (k > 1) {  // histories of iteration k-1 have landed: its stop rule (:50)
            error_norm[k - 2] = g.hist(k - 1)[0] / g.norm_xt;
            residual_norm[k - 2] = g.hist(k - 1)[1] / g.norm_b;
            pending = 0;
            if (residual_norm[k - 2] <= tol) {
                Bk[(size_t)(k - 1) * ldb + (k - 1)] = 0.0;  // entries the reference never reached
                k = k - 1;
                stopped = true;
                break;
            }
        }
        HG_TRY(hg_k_scale_div(ctx, t, m, g.ds + 5));  // :26
        if (literal_res) {
            swap_ptr(u, t);
        } else {  // U(:,k+1) stays where it is; the next u_hat goes to the next column
            u = t;
            t = u + ldu;
        }
        Bk[(size_t)(k - 1) * ldb + k] = beta_k;  // :27
        if (k < maxit) {                          // :29-35
            double* vn = V + (size_t)k * ldv;
            HG_TRY(g.apply_At(u, vn, v, -beta_k, nullptr, 0.0, true, &np));
            double a_next = 0;
            HG_TRY(g.norm_from_stat(np, 6, &a_next));
            HG_TRY(hg_k_scale_div(ctx, vn, nl, g.ds + 6));
            alpha1 = a_next;
        }
        // projected problem on the host                           (:37-44)
        const double alpha_k1 = alpha1, beta_k1 = beta_k;
        // T = Bk'*Bk and T^2 (:41-42).  Bk is bidiagonal, so T is tridiagonal and T^2 pentadiagonal: only the
        // non-zero terms of the reference's dense products are formed, in the same (increasing) order of the
        // summation index — the skipped terms are exact zeros, so the values are those of the dense products.
        T.assign((size_t)k * k, 0.0);
        for (int j = 0; j < k; ++j)
            for (int i = std::max(0, j - 1); i <= std::min(k - 1, j + 1); ++i) {
                double acc = 0.0;
                for (int r = std::max(i, j); r <= std::min(i, j) + 1; ++r)  // rows where both columns are non-zero
                    acc += Bk[(size_t)i * ldb + r] * Bk[(size_t)j * ldb + r];
                T[(size_t)j * k + i] = acc;
            }
        LHS.assign((size_t)k * k, 0.0);
        for (int j = 0; j < k; ++j)
            for (int i = std::max(0, j - 2); i <= std::min(k - 1, j + 2); ++i) {
                double acc = 0.0;
                for (int r = std::max(0, std::max(i, j) - 1); r <= std::min(k - 1, std::min(i, j) + 1); ++r)
                    acc += T[(size_t)r * k + i] * T[(size_t)j * k + r];
                LHS[(size_t)j * k + i] = acc;
            }
        LHS[0] += (alpha_k1 * beta_k1) * (alpha_k1 * beta_k1);
        for (int i = 0; i < k; ++i) LHS[(size_t)i * k + i] += lambda;
        RHS.assign(k, 0.0);
        for (int i = 0; i < k; ++i) RHS[i] = Bk[0] * beta1 * T[i];  // B_k(1,1)*beta1*(T*e1)
        double* yk = hy.p + (size_t)(k & 1) * (2 * maxit + 1);  // pinned, two slots: no wait for the upload
        hgd::solve_square(k, LHS.data(), k, RHS.data(), yk);
        HG_CUDA(cudaMemcpyAsync(by.p, yk, (size_t)k * 8, cudaMemcpyHostToDevice, st));
        // x = V(:,1:k)*yk + error                                 (:45,47)
        HG_TRY(hg_k_lincomb(ctx, V, ldv, nl, k, by.p, 1.0, nullptr, dx, g.xt.p, g.stat.p, &np));
        HG_TRY(g.enqueue_norm(np, 20));
        if (literal_res) {
            HG_TRY(g.residual_enqueue(dx, nullptr, 21));  // :48
        } else {
            // z = beta1*e1 - B_k*y (B_k lower bidiagonal, (k+1) x k); residual = || U(:,1:k+1) * z ||
            double* zk = yk + maxit;
            for (int i = 0; i <= k; ++i) {
                double acc = 0.0;
                if (i >= 1) acc += Bk[(size_t)(i - 1) * ldb + i] * yk[i - 1];  // beta_{i+1} y_i
                if (i < k) acc += Bk[(size_t)i * ldb + i] * yk[i];             // alpha_{i+1} y_{i+1}
                zk[i] = (i == 0 ? beta1 : 0.0) - acc;
            }
            HG_CUDA(cudaMemcpyAsync(bz.p, zk, (size_t)(k + 1) * 8, cudaMemcpyHostToDevice, st));
            HG_TRY(hg_k_lincomb(ctx, bU.p, ldu, m, k + 1, bz.p, 1.0, nullptr, nullptr, nullptr, g.stat.p, &np));
            HG_TRY(g.enqueue_norm(np, 21));
        }
        HG_TRY(g.hist_copy(k));
        pending = k;
        if (extras && extras->X_hist) HG_TRY(g.fetch_x(dx, extras->X_hist + (size_t)(k - 1) * n));
    }
    if (!stopped) k = maxit;
    if (pending) {
        HG_CUDA(cudaStreamSynchronize(st));
        error_norm[pending - 1] = g.hist(pending)[0] / g.norm_xt;
        residual_norm[pending - 1] = g.hist(pending)[1] / g.norm_b;
    }
    *niters = k;
    if (extras && extras->aux)
        for (int j = 0; j < maxit; ++j) {
            extras->aux[j] = Bk[(size_t)j * ldb + j];
            extras->aux[maxit + 1 + j] = Bk[(size_t)j * ldb + j + 1];
        }
    return g.fetch_x(dx, x);
}

// ---------------------------------------------------------------------------
// lsqr_solver.m:1-54
// ---------------------------------------------------------------------------
int lsqr_impl(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A, const hg_matrix* At, const double* b,
              const double* x_true, double tol, int maxit, double* x, double* error_norm,
              double* residual_norm, int* niters, hg_extras* extras) {
    HG_REQUIRE(ctx && A && x_true && x && error_norm && residual_norm && niters, "hg_lsqr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_lsqr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    Gkb g;
    HG_TRY(g.init(ctx, comm, A, At, b, x_true));
    const int64_t m = g.m, nl = g.nl, n = g.n;
    cudaStream_t st = ctx->stream;
    DBuf bu, bt, bv, bt3, bw, bx;
    HG_TRY(bu.alloc(m)); HG_TRY(bt.alloc(m)); HG_TRY(bv.alloc(nl, st, true)); HG_TRY(bt3.alloc(nl, st, true));
    HG_TRY(bw.alloc(nl)); HG_TRY(bx.alloc(nl, st, true));
    double *u = bu.p, *t = bt.p, *v = bv.p, *t3 = bt3.p, *w = bw.p, *dx = bx.p;
    double beta = 0, alpha = 0;
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, g.b.p, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta));  // :7
    if (m) HG_CUDA(cudaMemcpyAsync(u, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_scale_div(ctx, u, m, g.ds + 5));  // :8
    HG_TRY(g.apply_At(u, v, nullptr, 0.0, nullptr, 0.0, true, &np));  // :10
    HG_TRY(g.norm_from_stat(np, 6, &alpha));
    HG_TRY(hg_k_scale_div(ctx, v, nl, g.ds + 6));  // :12
    HG_CUDA(cudaMemcpyAsync(w, v, (size_t)nl * 8, cudaMemcpyDeviceToDevice, st));
    double phi_bar = beta, rho_bar = alpha;
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    int k, pending = 0;
    for (k = 1; k <= maxit; ++k) {
        HG_TRY(g.apply_A(v, t, 1.0, u, -alpha, nullptr, true, 0, &np));  // :22
        HG_TRY(g.norm_from_stat(np, 5, &beta));
        if (pending) {  // error norm of the previous iteration
            error_norm[pending - 1] = g.hist(pending)[0] / g.norm_xt;
            pending = 0;
        }
        HG_TRY(hg_k_scale_div(ctx, t, m, g.ds + 5));
        swap_ptr(u, t);
        HG_TRY(g.apply_At(u, t3, v, -beta, nullptr, 0.0, true, &np));  // :26
        HG_TRY(g.norm_from_stat(np, 6, &alpha));
        HG_TRY(hg_k_scale_div(ctx, t3, nl, g.ds + 6));
        swap_ptr(v, t3);
        const double rho = std::sqrt(rho_bar * rho_bar + beta * beta);  // :31
        const double c = rho_bar / rho;
        const double s = beta / rho;
        const double theta = s * alpha;
        rho_bar = -c * alpha;
        const double phi = c * phi_bar;
        phi_bar = s * phi_bar;
        HG_TRY(hg_k_lsqr_update(ctx, nl, dx, w, v, phi / rho, theta / rho, g.xt.p, g.stat.p, &np));  // :40-41
        HG_TRY(g.enqueue_norm(np, 20));  // :43 — read after the next synchronisation
        HG_TRY(g.hist_copy(k));
        pending = k;
        residual_norm[k - 1] = std::fabs(phi_bar) / g.norm_b;  // :44
        if (extras && extras->X_hist) HG_TRY(g.fetch_x(dx, extras->X_hist + (size_t)(k - 1) * n));
        if (residual_norm[k - 1] <= tol) break;  // :46
    }
    if (k > maxit) k = maxit;
    *niters = k;
    if (pending) {
        HG_CUDA(cudaStreamSynchronize(st));
        error_norm[pending - 1] = g.hist(pending)[0] / g.norm_xt;
    }
    double rn = 0;
    HG_TRY(g.residual(dx, nullptr, &rn));
    residual_norm[k - 1] = rn / g.norm_b;  // :52
    return g.fetch_x(dx, x);
}

// ---------------------------------------------------------------------------
// lsmr_solver.m:1-83
// ---------------------------------------------------------------------------
int lsmr_impl(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A, const hg_matrix* At, const double* b,
              const double* x_true, double tol, int maxit, double* x, double* err_hist, double* res_hist,
              double* ar_hist, int* iters, hg_extras* extras) {
    HG_REQUIRE(ctx && A && x && err_hist && res_hist && ar_hist && iters, "hg_lsmr_solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "hg_lsmr_solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    const double eps = 2.220446049250313e-16;
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    Gkb g;
    HG_TRY(g.init(ctx, comm, A, At, b, x_true));
    const int64_t m = g.m, nl = g.nl, n = g.n;
    cudaStream_t st = ctx->stream;
    DBuf bu, bt, bv, bt3, bh, bhb, bx, br, bah, bahb;
    HG_TRY(bu.alloc(m)); HG_TRY(bt.alloc(m)); HG_TRY(bv.alloc(nl, st, true)); HG_TRY(bt3.alloc(nl, st, true));
    HG_TRY(bh.alloc(nl)); HG_TRY(bhb.alloc(nl, st, true)); HG_TRY(bx.alloc(nl, st, true)); HG_TRY(br.alloc(m));
    double *u = bu.p, *t = bt.p, *v = bv.p, *t3 = bt3.p, *h = bh.p, *hbar = bhb.p, *dx = bx.p, *r = br.p;
    // r = b - A*x (:69) from the Golub-Kahan relation (A*h_k and A*hbar_k by recurrence, hg_k_gkb_resid) instead of
    // a product with A; the product A'*r of :71 stays.  gkb_residual = 1 forms b - A*x literally.
    const bool literal_res = hg_gkb_residual_mode() == 1;
    if (!literal_res) {
        HG_TRY(bah.alloc(m));
        HG_TRY(bahb.alloc(m));
        if (m) HG_CUDA(cudaMemcpyAsync(r, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));  // r_0 = b
    }
    double c_prev = 0.0;
    // norm(A,'fro'): recomputed every iteration by the reference (:71), constant here
    int np = 0;
    double normA = 0;
    HG_TRY(hg_k_sumsq(ctx, A->vals, A->nnz, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 9, &normA));
    double beta = 0, alpha = 0;
    if (m) HG_CUDA(cudaMemcpyAsync(u, g.b.p, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_sumsq(ctx, u, m, g.stat.p, &np));
    HG_TRY(g.norm_from_stat(np, 5, &beta));                         // :11
    if (beta > 0) HG_TRY(hg_k_scale_div(ctx, u, m, g.ds + 5));      // :12
    HG_TRY(g.apply_At(u, v, nullptr, 0.0, nullptr, 0.0, true, &np));  // :14
    HG_TRY(g.norm_from_stat(np, 6, &alpha));
    if (alpha > 0) HG_TRY(hg_k_scale_div(ctx, v, nl, g.ds + 6));    // :16
    double zetabar = alpha * beta, alphabar = alpha, rho = 1, rhobar = 1, cbar = 1, sbar = 0;  // :19-23
    HG_CUDA(cudaMemcpyAsync(h, v, (size_t)nl * 8, cudaMemcpyDeviceToDevice, st));              // :25
    const double nan = std::nan("");
    for (int i = 0; i < maxit; ++i) {
        err_hist[i] = nan;
        res_hist[i] = 0.0;
        ar_hist[i] = 0.0;
    }
    int k, pending = 0;
    bool stopped = false;
    auto consume = [&](int j) {  // histories of iteration j from the ring (:70-74)
        const double en = g.hist(j)[0], rn = g.hist(j)[1], arn = g.hist(j)[2];
        res_hist[j - 1] = rn / (g.norm_b + eps);
        ar_hist[j - 1] = arn / (normA * std::max(rn, eps));
        if (g.have_xt) err_hist[j - 1] = en / g.norm_xt;
    };
    for (k = 1; k <= maxit; ++k) {
        const double alpha_k = alpha;
        HG_TRY(g.apply_A(v, t, 1.0, u, -alpha, nullptr, true, 0, &np));  // :34
        HG_TRY(g.norm_from_stat(np, 5, &beta));
        if (k > 1) {  // histories of iteration k-1 have landed: its stop rule (:76, strict)
            consume(k - 1);
            pending = 0;
            if (res_hist[k - 2] < tol) {
                k = k - 1;
                stopped = true;
                break;
            }
        }
        if (beta > 0) HG_TRY(hg_k_scale_div(ctx, t, m, g.ds + 5));  // :36
        swap_ptr(u, t);
        HG_TRY(g.apply_At(u, t3, v, -beta, nullptr, 0.0, true, &np));  // :38
        HG_TRY(g.norm_from_stat(np, 6, &alpha));
        if (alpha > 0) HG_TRY(hg_k_scale_div(ctx, t3, nl, g.ds + 6));  // :40
        swap_ptr(v, t3);
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
        HG_TRY(hg_k_lsmr_update(ctx, nl, dx, h, hbar, v, k == 1 ? 1 : 0, c0, zeta / (rho * rhobar),
                                thetanew / rho, g.have_xt ? g.xt.p : nullptr, g.stat.p, &np));  // :61-67
        HG_TRY(g.enqueue_norm(np, 20));
        if (literal_res) {
            HG_TRY(g.residual_enqueue(dx, r, 21));  // :69
        } else {  // t holds u_k, u holds u_{k+1} (unscaled, i.e. zero, when beta == 0)
            HG_TRY(hg_k_gkb_resid(ctx, m, t, alpha_k, u, beta, bah.p, c_prev, bahb.p, c0, k == 1, r,
                                  zeta / (rho * rhobar), g.stat.p, &np));
            HG_TRY(g.enqueue_norm(np, 21));
            c_prev = thetanew / rho;
        }
        HG_TRY(g.apply_At(r, nullptr, nullptr, 0.0, nullptr, 0.0, true, &np));  // norm(A.'*r)  :71
        HG_TRY(g.enqueue_norm(np, 22));
        HG_TRY(g.hist_copy(k));
        pending = k;
        if (extras && extras->X_hist) HG_TRY(g.fetch_x(dx, extras->X_hist + (size_t)(k - 1) * n));
    }
    if (!stopped) k = maxit;
    if (pending) {
        HG_CUDA(cudaStreamSynchronize(st));
        consume(pending);
    }
    *iters = k;
    return g.fetch_x(dx, x);
}

}  // namespace

// Residual histories of the hybrid GKB solvers: 0 (default) from the Golub-Kahan relation (no third product with A per
// iteration), 1 literal b - A*x by SpMV.  Option "gkb_residual" / env HG_GKB_RESIDUAL.
static int g_gkb_residual = -1;
int hg_gkb_residual_mode() {
    if (g_gkb_residual < 0) {
        const char* e = getenv("HG_GKB_RESIDUAL");
        g_gkb_residual = (e && e[0] == '1') ? 1 : 0;
    }
    return g_gkb_residual;
}
void hg_gkb_residual_mode_set(int v) { g_gkb_residual = v == 1 ? 1 : 0; }

extern "C" int hg_hybrid_lsqr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                                     const double* x_true, double tol, int maxit, double lambda, double* x,
                                     double* error_norm, double* residual_norm, int* niters,
                                     hg_extras* extras) {
    HG_REQUIRE(b, "hg_hybrid_lsqr_solver: NULL argument");
    return hybrid_lsqr_impl(ctx, nullptr, A, At, b, x_true, tol, maxit, lambda, x, error_norm, residual_norm,
                            niters, extras);
}

extern "C" int hg_hybrid_lsmr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                                     const double* x_true, double tol, int maxit, double lambda, double* x,
                                     double* error_norm, double* residual_norm, int* niters,
                                     hg_extras* extras) {
    HG_REQUIRE(b, "hg_hybrid_lsmr_solver: NULL argument");
    return hybrid_lsmr_impl(ctx, nullptr, A, At, b, x_true, tol, maxit, lambda, x, error_norm, residual_norm,
                            niters, extras);
}

extern "C" int hg_lsqr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                              const double* x_true, double tol, int maxit, double* x, double* error_norm,
                              double* residual_norm, int* niters, hg_extras* extras) {
    HG_REQUIRE(b, "hg_lsqr_solver: NULL argument");
    return lsqr_impl(ctx, nullptr, A, At, b, x_true, tol, maxit, x, error_norm, residual_norm, niters, extras);
}

extern "C" int hg_lsmr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                              const double* x_true, double tol, int maxit, double* x, double* err_hist,
                              double* res_hist, double* ar_hist, int* iters, hg_extras* extras) {
    HG_REQUIRE(b, "hg_lsmr_solver: NULL argument");
    return lsmr_impl(ctx, nullptr, A, At, b, x_true, tol, maxit, x, err_hist, res_hist, ar_hist, iters, extras);
}

// Sharded GKB solvers (GKB row).  which: 0 hybrid_lsqr, 1 hybrid_lsmr, 2 lsqr,
// 3 lsmr.  A_p: this rank's row block of A; At_p: its transpose or NULL; b_p: its m_p entries
// of b; x_true / x: full n-vectors (x_true may be NULL for lsmr); ar_hist only for lsmr.
extern "C" int hg_dist_gkb_solver(int which, hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p,
                                  const hg_matrix* At_p, const double* b_p, const double* x_true, double tol,
                                  int maxit, double lambda, double* x, double* error_norm,
                                  double* residual_norm, double* ar_hist, int* niters, hg_extras* extras) {
    HG_REQUIRE(comm, "hg_dist_gkb_solver: NULL communicator");
    HG_REQUIRE(A_p && (b_p || A_p->rows == 0), "hg_dist_gkb_solver: NULL argument");
    switch (which) {
        case 0:
            return hybrid_lsqr_impl(ctx, comm, A_p, At_p, b_p, x_true, tol, maxit, lambda, x, error_norm,
                                    residual_norm, niters, extras);
        case 1:
            return hybrid_lsmr_impl(ctx, comm, A_p, At_p, b_p, x_true, tol, maxit, lambda, x, error_norm,
                                    residual_norm, niters, extras);
        case 2:
            return lsqr_impl(ctx, comm, A_p, At_p, b_p, x_true, tol, maxit, x, error_norm, residual_norm, niters,
                             extras);
        case 3:
            HG_REQUIRE(ar_hist, "hg_dist_gkb_solver: ar_hist is required for lsmr");
            return lsmr_impl(ctx, comm, A_p, At_p, b_p, x_true, tol, maxit, x, error_norm, residual_norm, ar_hist,
                             niters, extras);
        default:
            hg_set_error("hg_dist_gkb_solver: unknown solver %d", which);
            return HG_ERR_INVALID;
    }
}
