// Device Arnoldi (CGS2) and the solvers built on it:
//   hybrid_ab_gmres_rtp.m, hybrid_ba_gmres_rtp.m, gcv_function.m
#include <algorithm>
#include <chrono>
#include <memory>

#include "common.cuh"
#include "dense_host.h"

int hg_multidot_nslabs(const hg_ctx* ctx, int64_t n);

// wall-clock breakdown of the last whole-solver call on this thread (hg_last_solve_stats)
thread_local double g_solve_stats[8] = {0};

extern "C" int hg_last_solve_stats(double* out, int n) {
    HG_REQUIRE(out && n >= 0, "hg_last_solve_stats: bad argument");
    for (int i = 0; i < n; ++i) out[i] = i < 8 ? g_solve_stats[i] : 0.0;
    return HG_OK;
}

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

struct hg_arnoldi {
    hg_ctx* ctx = nullptr;
    const hg_matrix* A = nullptr;
    const hg_matrix* B = nullptr;
    const hg_matrix* M1 = nullptr;  // applied first
    const hg_matrix* M2 = nullptr;  // applied second; basis lives in rows(M2)
    int space = HG_SPACE_N;
    int kmax = 0;
    int64_t nq = 0, nt = 0, ldq = 0, ldt = 0, mb = 0;
    bool store_t = false;
    double* Q = nullptr;   // nq x (kmax+1)
    double* T = nullptr;   // n-space: m x (kmax+1), column 0 = b, column k = A*q_k
                           // m-space: one scratch column of length n
    double* d_b = nullptr; // m-space: b (length m); n-space: alias of T column 0
    double* w0 = nullptr;
    double* w1 = nullptr;
    double* d_H = nullptr;  // (kmax+1) x kmax, ld = kmax+1
    double* d_hcur = nullptr;
    double* d_beta = nullptr;
    double* h_H = nullptr;  // pinned mirror
    double* h_beta = nullptr;
    double* partials = nullptr;  // multidot partials
    double* stat = nullptr;      // norm partials
    double shift = 0.0;
    int k = 0;
    bool have_rhs = false, started = false;
    int ldh() const { return kmax + 1; }
};

extern "C" int hg_arnoldi_destroy(hg_arnoldi* a) {
    if (!a) return HG_OK;
    cudaStreamSynchronize(a->ctx->stream);
    hg_dfree(a->Q);
    hg_dfree(a->T);
    if (a->space == HG_SPACE_M) hg_dfree(a->d_b);
    hg_dfree(a->w0);
    hg_dfree(a->w1);
    hg_dfree(a->d_H);
    hg_dfree(a->d_hcur);
    hg_dfree(a->d_beta);
    hg_dfree(a->partials);
    hg_dfree(a->stat);
    hg_hfree(a->h_H);
    hg_hfree(a->h_beta);
    delete a;
    return HG_OK;
}

static int arnoldi_create(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, int space, int kmax, bool store_t_m,
                          hg_arnoldi** out);

extern "C" int hg_arnoldi_create(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, int space,
                                 int kmax, hg_arnoldi** out) {
    return arnoldi_create(ctx, A, B, space, kmax, false, out);
}

// store_t_m: in m-space also keep the columns B*q_k (n x (kmax+1)), so x = B*(Q y) is a combination of cached
// columns instead of a product with B (PTR AB solvers)
static int arnoldi_create(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, int space, int kmax, bool store_t_m,
                          hg_arnoldi** out) {
    HG_REQUIRE(ctx && A && B && out, "hg_arnoldi_create: NULL argument");
    HG_REQUIRE(space == HG_SPACE_N || space == HG_SPACE_M, "hg_arnoldi_create: bad space");
    HG_REQUIRE(kmax >= 1, "hg_arnoldi_create: kmax must be >= 1");
    HG_REQUIRE(A->rows == B->cols && A->cols == B->rows,
               "hg_arnoldi_create: B must be n x m for A m x n (A %lld x %lld, B %lld x %lld)",
               (long long)A->rows, (long long)A->cols, (long long)B->rows, (long long)B->cols);
    HG_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    hg_arnoldi* a = new (std::nothrow) hg_arnoldi();
    if (!a) {
        hg_set_error("hg_arnoldi_create: out of host memory");
        return HG_ERR_NOMEM;
    }
    a->ctx = ctx;
    a->A = A;
    a->B = B;
    a->space = space;
    a->kmax = kmax;
    a->mb = A->rows;
    if (space == HG_SPACE_N) {
        a->M1 = A;
        a->M2 = B;
        a->store_t = true;
    } else {
        a->M1 = B;
        a->M2 = A;
        a->store_t = store_t_m;
    }
    a->nq = a->M2->rows;
    a->nt = a->M1->rows;
    a->ldq = round_up(std::max<int64_t>(a->nq, 1), 32);
    a->ldt = round_up(std::max<int64_t>(a->nt, 1), 32);
    const int nslabs = hg_multidot_nslabs(ctx, std::max(a->nq, a->nt));
    const size_t npart = (size_t)(kmax + 2) * (size_t)(std::max(nslabs, ctx->sm_count) + 1);
    const size_t nstat = hg_stat_capacity(ctx, std::max(a->nq, a->nt));
    cudaError_t e = cudaSuccess;
    auto alloc = [&](double** p, size_t n) {
        if (e == cudaSuccess) e = hg_dmalloc(ctx, p, std::max<size_t>(n, 1) * sizeof(double));
    };
    alloc(&a->Q, (size_t)a->ldq * (kmax + 1));
    alloc(&a->T, a->store_t ? (size_t)a->ldt * (kmax + 1) : (size_t)a->ldt);
    if (space == HG_SPACE_M) alloc(&a->d_b, (size_t)a->mb);
    else a->d_b = a->T;
    alloc(&a->w0, (size_t)a->ldq);
    alloc(&a->w1, (size_t)a->ldq);
    alloc(&a->d_H, (size_t)a->ldh() * kmax);
    alloc(&a->d_hcur, (size_t)kmax + 1);
    alloc(&a->d_beta, 8);
    alloc(&a->partials, npart);
    alloc(&a->stat, nstat);
    if (e == cudaSuccess) e = hg_hmalloc(ctx, &a->h_H, (size_t)a->ldh() * kmax * sizeof(double));
    if (e == cudaSuccess) e = hg_hmalloc(ctx, &a->h_beta, 8 * sizeof(double));
    if (e != cudaSuccess) {
        hg_set_error("hg_arnoldi_create: allocation failed (basis %lld x %d): %s",
                     (long long)a->nq, kmax + 1, cudaGetErrorString(e));
        hg_arnoldi_destroy(a);
        return HG_ERR_NOMEM;
    }
    // padding rows of Q/T are never read by the kernels (all loops are bounded by n)
    *out = a;
    return HG_OK;
}

// device-resident rhs (no host copy): used by the solvers and the bench
int hg_arnoldi_set_rhs_device(hg_arnoldi* a, const double* d_b) {
    HG_CUDA(cudaMemcpyAsync(a->d_b, d_b, (size_t)a->mb * 8, cudaMemcpyDeviceToDevice, a->ctx->stream));
    a->have_rhs = true;
    return HG_OK;
}

extern "C" int hg_arnoldi_set_rhs(hg_arnoldi* a, const double* b) {
    HG_REQUIRE(a && b, "hg_arnoldi_set_rhs: NULL argument");
    HG_CUDA(cudaSetDevice(a->ctx->device));
    HG_CUDA(cudaMemcpyAsync(a->d_b, b, (size_t)a->mb * 8, cudaMemcpyHostToDevice, a->ctx->stream));
    HG_CUDA(cudaStreamSynchronize(a->ctx->stream));  // caller may free b
    a->have_rhs = true;
    return HG_OK;
}

extern "C" int hg_arnoldi_reset(hg_arnoldi* a, double shift) {
    HG_REQUIRE(a, "hg_arnoldi_reset: NULL");
    if (!a->have_rhs) {
        hg_set_error("hg_arnoldi_reset: right-hand side not set");
        return HG_ERR_STATE;
    }
    hg_ctx* ctx = a->ctx;
    HG_CUDA(cudaSetDevice(ctx->device));
    a->shift = shift;
    a->k = 0;
    // reuse of the pinned mirror: make sure earlier async copies have landed
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    memset(a->h_H, 0, (size_t)a->ldh() * a->kmax * sizeof(double));
    HG_CUDA(cudaMemsetAsync(a->d_H, 0, (size_t)a->ldh() * a->kmax * sizeof(double), ctx->stream));
    double* q0 = a->Q;
    if (a->space == HG_SPACE_N) {
        // r0 = B*b - M(0) = B*b      (hybrid_ba_gmres_rtp.m:7-9; gcv_function.m:8)
        hg_spmv_epilogue ep;
        HG_TRY(hg_k_spmv(ctx, a->B, a->d_b, q0, ep, nullptr));
    } else {
        // r0 = b                      (gcv_function.m:5)
        HG_CUDA(cudaMemcpyAsync(q0, a->d_b, (size_t)a->nq * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, q0, a->nq, a->stat, &np));
    HG_TRY(hg_k_reduce(ctx, a->stat, np, 1, a->d_beta, false, nullptr, true));  // beta = norm(r0)
    HG_TRY(hg_k_scale_div(ctx, q0, a->nq, a->d_beta));                          // Q(:,1) = r0/beta
    HG_CUDA(cudaMemcpyAsync(a->h_beta, a->d_beta, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    a->started = true;
    return HG_OK;
}

// one CGS2 Arnoldi step; kk is the 1-based column being produced
static int arnoldi_step(hg_arnoldi* a, int kk) {
    hg_ctx* ctx = a->ctx;
    const double* q = a->Q + (size_t)(kk - 1) * a->ldq;
    double* tcol = a->store_t ? a->T + (size_t)kk * a->ldt : a->T;
    double* qnext = a->Q + (size_t)kk * a->ldq;
    double* Hcol = a->d_H + (size_t)(kk - 1) * a->ldh();
    // v = M2*(M1*q) + shift*q          (hybrid_ab_gmres_rtp.m:6,19; gcv_function.m:20,22)
    {
        hg_spmv_epilogue ep;
        HG_TRY(hg_k_spmv(ctx, a->M1, q, tcol, ep, nullptr));
    }
    {
        hg_spmv_epilogue ep;
        if (a->shift != 0.0) {
            ep.z1 = q;
            ep.g1 = a->shift;
        }
        HG_TRY(hg_k_spmv(ctx, a->M2, tcol, a->w0, ep, nullptr));
    }
    // CGS2: h1 = Q_k' v ; v -= Q_k h1 ; h2 = Q_k' v ; v -= Q_k h2 ; H(1:k,k) = h1 + h2
    // (the reference's MGS sweep, hybrid_ab_gmres_rtp.m:20-23, in its two-pass
    //  classical form mandated by the north star)
    if (hg_cgs2_step_eligible(ctx, a->nq, kk)) {
        // small / medium vectors: the whole orthogonalisation, norm and normalisation in one persistent kernel
        HG_TRY(hg_k_cgs2_step(ctx, a->Q, a->ldq, a->nq, kk, a->w0, a->w1, qnext, Hcol, a->d_hcur, a->partials));
        HG_CUDA(cudaMemcpyAsync(a->h_H + (size_t)(kk - 1) * a->ldh(), Hcol, (size_t)(kk + 1) * 8,
                                cudaMemcpyDeviceToHost, ctx->stream));
        return HG_OK;
    }
    int ns = 0, np = 0;
    HG_TRY(hg_k_multidot(ctx, a->Q, a->ldq, a->nq, kk, a->w0, a->partials, &ns));
    HG_TRY(hg_k_reduce(ctx, a->partials, ns, kk, Hcol, false, a->d_hcur, false));
    if (hg_cgs_fused_mode() == 2 && hg_cgs_staged_nparts(ctx, a->nq, kk) > 0) {
        // v -= Q h1 and h2 = Q' v in one pass over Q: the tile is staged in shared memory by cp.async
        HG_TRY(hg_k_cgs_mid_staged(ctx, a->Q, a->ldq, a->nq, kk, a->d_hcur, a->w0, a->w1, a->partials, &ns));
    } else {
        HG_TRY(hg_k_lincomb_push(ctx, a->Q, a->ldq, a->nq, kk, a->d_hcur, -1.0, a->w0, a->w1, nullptr,
                                 nullptr, nullptr, nullptr));
        HG_TRY(hg_k_multidot(ctx, a->Q, a->ldq, a->nq, kk, a->w1, a->partials, &ns));
    }
    HG_TRY(hg_k_reduce(ctx, a->partials, ns, kk, Hcol, true, a->d_hcur, false));
    HG_TRY(hg_k_lincomb_push(ctx, a->Q, a->ldq, a->nq, kk, a->d_hcur, -1.0, a->w1, qnext, nullptr,
                             a->stat, &np, nullptr));
    // H(k+1,k) = norm(v) ; Q(:,k+1) = v / H(k+1,k)       (:24,26)
    HG_TRY(hg_k_reduce(ctx, a->stat, np, 1, Hcol + kk, false, nullptr, true));
    HG_TRY(hg_k_scale_div(ctx, qnext, a->nq, Hcol + kk));
    HG_CUDA(cudaMemcpyAsync(a->h_H + (size_t)(kk - 1) * a->ldh(), Hcol, (size_t)(kk + 1) * 8,
                            cudaMemcpyDeviceToHost, ctx->stream));
    return HG_OK;
}

extern "C" int hg_arnoldi_steps(hg_arnoldi* a, int nsteps) {
    HG_REQUIRE(a, "hg_arnoldi_steps: NULL");
    if (!a->started) {
        hg_set_error("hg_arnoldi_steps: call hg_arnoldi_reset first");
        return HG_ERR_STATE;
    }
    HG_REQUIRE(nsteps >= 0 && a->k + nsteps <= a->kmax, "hg_arnoldi_steps: %d steps from k=%d exceeds kmax=%d",
               nsteps, a->k, a->kmax);
    HG_CUDA(cudaSetDevice(a->ctx->device));
    for (int i = 0; i < nsteps; ++i) {
        HG_TRY(arnoldi_step(a, a->k + 1));
        a->k += 1;
    }
    return HG_OK;
}

extern "C" int hg_arnoldi_get(hg_arnoldi* a, double* H, int ldh, double* beta, int* ksteps) {
    HG_REQUIRE(a, "hg_arnoldi_get: NULL");
    HG_CUDA(cudaStreamSynchronize(a->ctx->stream));
    if (H) {
        HG_REQUIRE(ldh >= a->ldh(), "hg_arnoldi_get: ldh too small");
        for (int j = 0; j < a->kmax; ++j)
            memcpy(H + (size_t)j * ldh, a->h_H + (size_t)j * a->ldh(), (size_t)a->ldh() * 8);
    }
    if (beta) *beta = a->h_beta[0];
    if (ksteps) *ksteps = a->k;
    return HG_OK;
}

extern "C" int hg_arnoldi_get_q(hg_arnoldi* a, int j, double* q) {
    HG_REQUIRE(a && q, "hg_arnoldi_get_q: NULL");
    HG_REQUIRE(j >= 0 && j <= a->kmax, "hg_arnoldi_get_q: column out of range");
    HG_CUDA(cudaMemcpyAsync(q, a->Q + (size_t)j * a->ldq, (size_t)a->nq * 8, cudaMemcpyDeviceToHost,
                            a->ctx->stream));
    HG_CUDA(cudaStreamSynchronize(a->ctx->stream));
    return HG_OK;
}

extern "C" int hg_arnoldi_step_bytes(hg_arnoldi* a, int k, double* bytes) {
    HG_REQUIRE(a && bytes, "hg_arnoldi_step_bytes: NULL");
    // S(k) = matrix streams + 16 m + 88 n + CGS2(k) (n-space; m-space swaps the vector
    // terms).  Matrix streams: (8 + iw) bytes per entry + pointers in the form each matrix runs with
    // (iw = 4, or 2 + base words for the sliced form with 16-bit offsets).  CGS2(k) = 32 k n with
    // separate update / multi-dot kernels, 24 k n when the middle stage is the one-pass staged kernel.
    const double nq = (double)a->nq, nt = (double)a->nt;
    const bool one_pass = hg_cgs2_step_eligible(a->ctx, a->nq, k) ||
                          (hg_cgs_fused_mode() == 2 && hg_cgs_staged_nparts(a->ctx, a->nq, k) > 0);
    *bytes = hg_spmv_stream_bytes(a->A) + hg_spmv_stream_bytes(a->B) + 16.0 * nt + 88.0 * nq +
             (one_pass ? 24.0 : 32.0) * (double)k * nq;
    return HG_OK;
}

// ===========================================================================
// RTP solvers
// ===========================================================================
namespace {

struct DBuf {
    double* p = nullptr;
    ~DBuf() { hg_dfree(p); }
    int alloc(size_t n) {
        cudaError_t e = hg_dmalloc_cur((void**)&p, std::max<size_t>(n, 1) * sizeof(double));
        if (e != cudaSuccess) {
            hg_set_error("device allocation of %zu doubles failed: %s", n, cudaGetErrorString(e));
            return HG_ERR_NOMEM;
        }
        return HG_OK;
    }
};

struct PinBuf {
    double* p = nullptr;
    ~PinBuf() { hg_hfree(p); }
    int alloc(size_t n) {
        cudaError_t e = hg_hmalloc_cur((void**)&p, std::max<size_t>(n, 1) * sizeof(double));
        if (e != cudaSuccess) {
            hg_set_error("pinned allocation of %zu doubles failed: %s", n, cudaGetErrorString(e));
            return HG_ERR_NOMEM;
        }
        return HG_OK;
    }
};

struct EventRing {
    std::vector<cudaEvent_t> e;
    ~EventRing() {
        for (cudaEvent_t x : e) cudaEventDestroy(x);
    }
    int create(int n) {
        for (int i = 0; i < n; ++i) {
            cudaEvent_t x;
            HG_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
            e.push_back(x);
        }
        return HG_OK;
    }
};

struct ArnoldiHolder {
    hg_arnoldi* a = nullptr;
    ~ArnoldiHolder() { hg_arnoldi_destroy(a); }
};

enum RtpKind { RTP_AB, RTP_BA };

double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int rtp_solver(RtpKind kind, hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, const double* b,
               const double* x_true, double tol, int maxit, double lambda, double* x,
               double* error_norm, double* residual_norm, int* niters, int* x_valid,
               const hg_solver_opts* opts, hg_extras* extras) {
    HG_REQUIRE(ctx && A && B && b && x_true && x && error_norm && residual_norm && niters,
               "rtp solver: NULL argument");
    HG_REQUIRE(maxit >= 1, "rtp solver: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    const int residual_mode = opts ? opts->residual_mode : 0;
    // error_mode 2: ||x_k - x_true|| from the orthonormal basis, x formed once at the end (see below);
    // 1: form x_k and the difference explicitly at every iteration (hybrid_ba_gmres_rtp.m:30,33 literally).
    // The literal residual mode and a request for every iterate (extras->X_hist) need x_k anyway.
    static const int env_error_mode = [] {
        const char* e = getenv("HG_ERROR_MODE");
        return e ? atoi(e) : -1;
    }();
    int error_mode = opts ? opts->error_mode : 0;
    if (env_error_mode >= 0) error_mode = env_error_mode;
    const int64_t n = A->cols, m = A->rows;
    // auto: the algebraic form pays once x = Q_k y_k (8 k n bytes) costs more than the two small launches the
    // extra dot product needs — measured: a gain at 1024^2 (loop 338 -> 319 ms), a loss at 256^2 (5 950 -> 5 380 it/s)
    if (error_mode == 0) error_mode = n >= 200000 ? 2 : 1;
    if (residual_mode != 0 || (extras && extras->X_hist)) error_mode = 1;
    error_mode = error_mode == 2 ? 0 : 1;  // below: 0 algebraic, 1 explicit
    const bool trace = getenv("HG_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_host = 0.0, t_wait = 0.0;
    ArnoldiHolder holder;
    HG_TRY(hg_arnoldi_create(ctx, A, B, HG_SPACE_N, maxit, &holder.a));
    hg_arnoldi* a = holder.a;
    const double t_created = now();
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    // The loop is software pipelined (see below): iterates alternate between two buffers, and every
    // host-visible result of an iteration has RING slots so that a slot is rewritten only after the
    // host has consumed it.
    constexpr int RING = 4;
    DBuf d_x[2], d_xt, d_y, d_g, d_c, stat_e, stat_r;
    PinBuf h_y, h_g, h_s, h_c;
    EventRing ev;
    HG_TRY(ev.create(RING));
    HG_TRY(d_x[0].alloc((size_t)n));
    HG_TRY(d_x[1].alloc((size_t)n));
    HG_TRY(d_xt.alloc((size_t)n));
    HG_TRY(d_y.alloc((size_t)maxit + 1));
    HG_TRY(d_g.alloc((size_t)maxit + 2));
    HG_TRY(d_c.alloc((size_t)maxit + 2));
    HG_TRY(stat_e.alloc(hg_stat_capacity(ctx, std::max(n, m))));
    HG_TRY(stat_r.alloc(hg_stat_capacity(ctx, std::max(n, m))));
    HG_TRY(h_c.alloc((size_t)maxit + 2));
    HG_TRY(h_y.alloc((size_t)RING * (maxit + 1)));
    HG_TRY(h_g.alloc((size_t)RING * (maxit + 2)));
    HG_TRY(h_s.alloc((size_t)RING * 2));
    const double t_bufs = now();
    HG_CUDA(cudaMemsetAsync(d_x[0].p, 0, (size_t)n * 8, ctx->stream));
    HG_CUDA(cudaMemsetAsync(d_x[1].p, 0, (size_t)n * 8, ctx->stream));
    HG_CUDA(cudaMemcpyAsync(d_xt.p, x_true, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    HG_TRY(hg_arnoldi_set_rhs(a, b));
    double nb2 = 0, nx2 = 0;
    HG_TRY(hg_norm2_sync(ctx, a->d_b, m, &nb2));
    HG_TRY(hg_norm2_sync(ctx, d_xt.p, n, &nx2));
    const double norm_b = std::sqrt(nb2), norm_xt = std::sqrt(nx2);
    const double t_norms = now();
    HG_TRY(hg_arnoldi_reset(a, lambda));
    // c_j = q_j' x_true, one dot product per basis vector: with an orthonormal basis (CGS2: 1e-15)
    //   ||Q_k y - x_true||^2 = ||y||^2 - 2 y'c + ||x_true||^2
    // gives the error history in O(n) per iteration instead of the O(k n) product x = Q_k y_k
    auto queue_c = [&](int j) -> int {  // j: 0-based basis column
        int ns = 0;
        HG_TRY(hg_k_multidot(ctx, a->Q + (size_t)j * a->ldq, a->ldq, n, 1, d_xt.p, a->partials, &ns));
        HG_TRY(hg_k_reduce(ctx, a->partials, ns, 1, d_c.p + j, false, nullptr, false));
        HG_CUDA(cudaMemcpyAsync(h_c.p + j, d_c.p + j, 8, cudaMemcpyDeviceToHost, ctx->stream));
        return HG_OK;
    };
    if (error_mode == 0) HG_TRY(queue_c(0));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    const double beta = a->h_beta[0];

    const double t_setup = now();
    if (trace)
        fprintf(stderr, "[hg trace] rtp setup: basis alloc %.1f ms, buffers %.1f ms, rhs + norms %.1f ms, r0 (incl. lazy SpMV forms) %.1f ms\n",
                t_created - t_begin, t_bufs - t_created, t_norms - t_bufs, t_setup - t_norms);
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    hgd::HessenbergLS ls;
    hgd::BorderedCholesky chol;
    std::vector<double> Gfull, rhs;
    bool chol_ok = true;
    if (kind == RTP_BA) ls.reset(maxit, beta);
    else {
        chol.reset(maxit, lambda);
        Gfull.assign((size_t)maxit * maxit, 0.0);
        rhs.assign(maxit, 0.0);
    }
    const int ldh = a->ldh();

    // ---- software-pipelined loop --------------------------------------------------------------
    // Arnoldi step k+1 does not depend on y_k, so it is queued BEFORE the host looks at step k:
    //   stream:  step 1 | step 2 | iterate 1 | step 3 | iterate 2 | ...
    //   host  :  waits for "iterate k-1 done" (which implies step k done), solves the projected
    //            problem of step k while the device runs step k+1, queues iterate k behind it.
    // The reference's control flow is kept exactly: the stop rule of iteration k-1 (:38 / :35) is
    // applied before anything of iteration k is looked at, the `== 0` breakdown of step k (:25)
    // leaves before x / histories of iteration k exist; at most one queued step + one iterate are
    // discarded when the loop ends early.
    int enq = 0;         // Arnoldi steps queued so far
    int last_x = 0;      // last iteration whose iterate exists (0: none; BA then returns its zeros, :4)
    int x_formed = 0;    // last iteration whose iterate was actually written to d_x[k & 1]
    auto queue_step = [&](int kk) -> int {
        HG_TRY(hg_arnoldi_steps(a, 1));
        if (kind == RTP_AB) {
            // Gram column of W = A*Q_k against [b, W]: one stream over the cached columns
            // replaces `AQk = A*Qk; AQk'*AQk; AQk'*b` (hybrid_ab_gmres_rtp.m:31-32)
            int ns = 0;
            HG_TRY(hg_k_multidot(ctx, a->T, a->ldt, m, kk + 1, a->T + (size_t)kk * a->ldt, a->partials, &ns));
            HG_TRY(hg_k_reduce(ctx, a->partials, ns, kk + 1, d_g.p, false, nullptr, false));
            HG_CUDA(cudaMemcpyAsync(h_g.p + (size_t)(kk % RING) * (maxit + 2), d_g.p, (size_t)(kk + 1) * 8,
                                    cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (error_mode == 0 && kk < maxit) HG_TRY(queue_c(kk));  // q_{kk+1}' x_true (needed from iteration kk+1 on)
        if (kk == 1) HG_CUDA(cudaEventRecord(ev.e[0], ctx->stream));
        return HG_OK;
    };
    std::vector<double> err_alg(maxit, 0.0);  // error_mode 0: error norms formed on the host
    std::vector<char> err_explicit(maxit, 0);
    auto finish_iterate = [&](int j, bool* stop) -> int {  // histories of iteration j, stop rule
        const double tw = now();
        HG_CUDA(cudaEventSynchronize(ev.e[j % RING]));
        t_wait += now() - tw;
        const double* hs = h_s.p + (size_t)(j % RING) * 2;
        residual_norm[j - 1] = hs[1] / norm_b;
        error_norm[j - 1] = (error_mode == 0 && !err_explicit[j - 1]) ? err_alg[j - 1] / norm_xt : hs[0] / norm_xt;
        last_x = j;
        *stop = residual_norm[j - 1] <= tol;  // :38 / :35
        return HG_OK;
    };
    int k = 0;
    bool ended_early = false;
    for (k = 1; k <= maxit; ++k) {
        while (enq < std::min(k + 1, maxit)) {
            ++enq;
            HG_TRY(queue_step(enq));
        }
        if (k == 1) {
            const double tw = now();
            HG_CUDA(cudaEventSynchronize(ev.e[0]));
            t_wait += now() - tw;
        } else {
            bool stop = false;
            HG_TRY(finish_iterate(k - 1, &stop));
            if (stop) {
                k = k - 1;
                ended_early = true;
                break;
            }
        }
        const double th = now();
        const double* hcol = a->h_H + (size_t)(k - 1) * ldh;
        if (hcol[k] == 0.0) {  // :25 — leaves before x / histories are touched
            ended_early = true;
            break;
        }
        double* yk = h_y.p + (size_t)(k % RING) * (maxit + 1);
        if (kind == RTP_BA) {
            ls.add_column(hcol);  // yk = H(1:k+1,1:k) \ [beta;0]   (hybrid_ba_gmres_rtp.m:28-29)
            ls.solve(yk);
        } else {
            const double* g = h_g.p + (size_t)(k % RING) * (maxit + 2);
            rhs[k - 1] = g[0];
            for (int j = 0; j < k; ++j) {
                Gfull[(size_t)(k - 1) * maxit + j] = g[1 + j];
                Gfull[(size_t)j * maxit + (k - 1)] = g[1 + j];
            }
            if (chol_ok) chol_ok = chol.add_row(g + 1);
            if (chol_ok) {
                chol.solve(rhs.data(), yk);
            } else {  // mldivide's non-SPD path
                std::vector<double> M((size_t)k * k);
                for (int j = 0; j < k; ++j)
                    for (int i2 = 0; i2 < k; ++i2)
                        M[(size_t)j * k + i2] = Gfull[(size_t)j * maxit + i2] + (i2 == j ? lambda : 0.0);
                hgd::solve_square(k, M.data(), k, rhs.data(), yk);
            }
        }
        t_host += now() - th;
        HG_CUDA(cudaMemcpyAsync(d_y.p, yk, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
        bool form_x = error_mode != 0;
        if (error_mode == 0) {
            // ||x_k - x_true||^2 = ||y||^2 - 2 y'c + ||x_true||^2 in extended precision; when the error is small
            // against ||x_true|| the subtraction cancels, so below 1 % relative error the iterate and the
            // difference are formed explicitly for this iteration (rounding of the formula: ~1e-15 ||x_true||^2)
            long double yy = 0.0L, yc = 0.0L;
            for (int j = 0; j < k; ++j) {
                yy += (long double)yk[j] * yk[j];
                yc += (long double)yk[j] * h_c.p[j];
            }
            const long double e2 = yy - 2.0L * yc + (long double)nx2;
            if (e2 > 1e-4L * (long double)nx2) {
                err_alg[k - 1] = (double)sqrtl(e2);
            } else {
                err_explicit[k - 1] = 1;
                form_x = true;
            }
            if (form_x) x_formed = k;
        }
        // x = Q(:,1:k)*yk fused with ||x - x_true||^2          (:33,36 / :30,33)
        double* xk = d_x[k & 1].p;
        if (residual_mode == 0) {
            // ... and ||b - A*x|| with A*x = (A*Q_k) yk from the cached columns, both norms finished by the
            // last block of the same launch
            // (error_mode 0: only the residual job; x is formed once, after the loop)
            HG_TRY(hg_k_iterate(ctx, a->Q, a->ldq, form_x ? n : 0, a->T + a->ldt, a->ldt, m, k, d_y.p, a->d_b, xk, d_xt.p,
                                stat_e.p, reinterpret_cast<unsigned int*>(ctx->d_scalars + 40), ctx->d_scalars + 1));
        } else {
            int np_e = 0, np_r = 0;
            HG_TRY(hg_k_lincomb(ctx, a->Q, a->ldq, n, k, d_y.p, 1.0, nullptr, xk, d_xt.p, stat_e.p, &np_e));
            hg_spmv_epilogue ep;  // literal: norm(b - A*x)       (:35 / :32)
            ep.alpha = -1.0;
            ep.z1 = a->d_b;
            ep.g1 = 1.0;
            ep.stat = stat_r.p;
            HG_TRY(hg_k_spmv(ctx, A, xk, nullptr, ep, &np_r));
            HG_TRY(hg_k_reduce(ctx, stat_e.p, np_e, 1, ctx->d_scalars + 1, false, nullptr, true));
            HG_TRY(hg_k_reduce(ctx, stat_r.p, np_r, 1, ctx->d_scalars + 2, false, nullptr, true));
        }
        HG_CUDA(cudaMemcpyAsync(h_s.p + (size_t)(k % RING) * 2, ctx->d_scalars + 1, 16, cudaMemcpyDeviceToHost,
                                ctx->stream));
        if (extras && extras->X_hist)
            HG_CUDA(cudaMemcpyAsync(extras->X_hist + (size_t)(k - 1) * n, xk, (size_t)n * 8,
                                    cudaMemcpyDeviceToHost, ctx->stream));
        HG_CUDA(cudaEventRecord(ev.e[k % RING], ctx->stream));
    }
    if (!ended_early) {  // MATLAB leaves k at its last value
        k = maxit;
        bool stop = false;
        HG_TRY(finish_iterate(maxit, &stop));
    }
    *niters = k;
    const bool have_x = (kind == RTP_BA) || last_x > 0;  // BA initialises x = zeros (hybrid_ba_gmres_rtp.m:4)
    const double t_loop_end = now();
    if (error_mode == 0 && last_x > 0 && x_formed != last_x) {
        // x = Q(:,1:k)*yk for the iteration the reference stops at, formed once     (:33 / :30)
        const double* yl = h_y.p + (size_t)(last_x % RING) * (maxit + 1);
        HG_CUDA(cudaMemcpyAsync(d_y.p, yl, (size_t)last_x * 8, cudaMemcpyHostToDevice, ctx->stream));
        HG_TRY(hg_k_lincomb(ctx, a->Q, a->ldq, n, last_x, d_y.p, 1.0, nullptr, d_x[last_x & 1].p, nullptr, nullptr, nullptr));
    }
    HG_CUDA(cudaMemcpyAsync(x, d_x[last_x & 1].p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));  // also drains a discarded speculative step
    g_solve_stats[0] = t_setup - t_begin;
    g_solve_stats[1] = t_loop_end - t_setup;
    g_solve_stats[2] = t_host;
    g_solve_stats[3] = t_wait;
    g_solve_stats[4] = now() - t_loop_end;
    g_solve_stats[5] = (double)k;
    if (trace)
        fprintf(stderr, "[hg trace] rtp %s: setup %.1f ms, loop %.1f ms (host solve %.1f, gpu wait %.1f), k=%d\n",
                kind == RTP_AB ? "AB" : "BA", t_setup - t_begin, now() - t_setup, t_host, t_wait, k);
    if (x_valid) *x_valid = have_x ? 1 : 0;
    if (extras) {
        if (extras->beta) *extras->beta = beta;
        if (extras->H) {  // columns the reference never reached stay zero (a discarded speculative step wrote one)
            memset(extras->H, 0, (size_t)ldh * maxit * 8);
            memcpy(extras->H, a->h_H, (size_t)ldh * k * 8);
        }
        if (extras->X_hist && last_x < maxit)  // the discarded speculative iterate, if any
            memset(extras->X_hist + (size_t)last_x * n, 0, (size_t)(maxit - last_x) * n * 8);
    }
    return HG_OK;
}

}  // namespace

extern "C" int hg_hybrid_ab_gmres_rtp(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B,
                                      const double* b, const double* x_true, double tol, int maxit,
                                      double lambda, double* x, double* error_norm,
                                      double* residual_norm, int* niters, int* x_valid,
                                      const hg_solver_opts* opts, hg_extras* extras) {
    const double t0 = wall_ms();
    const int st = rtp_solver(RTP_AB, ctx, A, B, b, x_true, tol, maxit, lambda, x, error_norm, residual_norm,
                              niters, x_valid, opts, extras);
    if (getenv("HG_TRACE")) fprintf(stderr, "[hg trace] rtp AB: whole call %.1f ms (incl. release of device buffers)\n", wall_ms() - t0);
    return st;
}

extern "C" int hg_hybrid_ba_gmres_rtp(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B,
                                      const double* b, const double* x_true, double tol, int maxit,
                                      double lambda, double* x, double* error_norm,
                                      double* residual_norm, int* niters, int* x_valid,
                                      const hg_solver_opts* opts, hg_extras* extras) {
    const double t0 = wall_ms();
    const int st = rtp_solver(RTP_BA, ctx, A, B, b, x_true, tol, maxit, lambda, x, error_norm, residual_norm,
                              niters, x_valid, opts, extras);
    if (getenv("HG_TRACE")) fprintf(stderr, "[hg trace] rtp BA: whole call %.1f ms (incl. release of device buffers)\n", wall_ms() - t0);
    return st;
}

// ===========================================================================
// gcv_function
// ===========================================================================
struct hg_gcv {
    int k = 0;
    double beta = 0.0;
    double trace_m = 0.0;
    std::vector<double> H;   // (k+1) x k
    std::vector<double> sv;  // singular values of H(1:k,1:k)
};

static void gcv_fill_sv(hg_gcv* g);

extern "C" int hg_gcv_prepare(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, const double* b,
                              int64_t m, int k_gcv, int gcv_type, hg_gcv** out) {
    HG_REQUIRE(ctx && A && B && b && out, "hg_gcv_prepare: NULL argument");
    HG_REQUIRE(gcv_type == 0 || gcv_type == 1, "hg_gcv_prepare: gcv_type must be 0 ('ab') or 1 ('ba')");
    HG_REQUIRE(k_gcv >= 1, "hg_gcv_prepare: k_gcv must be >= 1");
    HG_REQUIRE(gcv_type == 1 || m == A->rows,
               "hg_gcv_prepare: m (%lld) must equal size(A,1) (%lld) for 'ab' (gcv_function.m:6,13)",
               (long long)m, (long long)A->rows);
    *out = nullptr;
    ArnoldiHolder holder;
    HG_TRY(hg_arnoldi_create(ctx, A, B, gcv_type == 0 ? HG_SPACE_M : HG_SPACE_N, k_gcv, &holder.a));
    hg_arnoldi* a = holder.a;
    HG_TRY(hg_arnoldi_set_rhs(a, b));
    HG_TRY(hg_arnoldi_reset(a, 0.0));  // unshifted operator (gcv_function.m:20,22)
    for (int k = 1; k <= k_gcv; ++k) {
        HG_TRY(hg_arnoldi_steps(a, 1));
        HG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (a->h_H[(size_t)(k - 1) * a->ldh() + k] < 1e-12) break;  // :30
    }
    hg_gcv* g = new (std::nothrow) hg_gcv();
    if (!g) {
        hg_set_error("hg_gcv_prepare: out of host memory");
        return HG_ERR_NOMEM;
    }
    g->k = k_gcv;  // :33 k = size(H,2): trailing zero columns are kept after an early break
    g->beta = a->h_beta[0];
    g->trace_m = gcv_type == 0 ? (double)m : (double)A->cols;  // :46-50
    g->H.assign(a->h_H, a->h_H + (size_t)(k_gcv + 1) * k_gcv);
    gcv_fill_sv(g);  // :42
    *out = g;
    return HG_OK;
}

static void gcv_fill_sv(hg_gcv* g) {
    const int k = g->k;
    std::vector<double> sq((size_t)k * k);
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < k; ++i) sq[(size_t)j * k + i] = g->H[(size_t)j * (k + 1) + i];
    g->sv.resize(k);
    hgd::singular_values(k, sq.data(), k, g->sv.data());  // gcv_function.m:42
}

extern "C" int hg_gcv_from_H(const double* H, int ldh, int k, double beta, double trace_m,
                             hg_gcv** out) {
    HG_REQUIRE(H && out, "hg_gcv_from_H: NULL argument");
    HG_REQUIRE(k >= 1 && ldh >= k + 1, "hg_gcv_from_H: bad shape");
    hg_gcv* g = new (std::nothrow) hg_gcv();
    if (!g) {
        hg_set_error("hg_gcv_from_H: out of host memory");
        return HG_ERR_NOMEM;
    }
    g->k = k;
    g->beta = beta;
    g->trace_m = trace_m;
    g->H.resize((size_t)(k + 1) * k);
    for (int j = 0; j < k; ++j) memcpy(&g->H[(size_t)j * (k + 1)], H + (size_t)j * ldh, (size_t)(k + 1) * 8);
    gcv_fill_sv(g);
    *out = g;
    return HG_OK;
}

extern "C" int hg_host_hessenberg_ls(const double* H, int ldh, int k, double beta, double* y) {
    HG_REQUIRE(H && y && k >= 1 && ldh >= k + 1, "hg_host_hessenberg_ls: bad argument");
    hgd::HessenbergLS ls;
    ls.reset(k, beta);
    for (int j = 0; j < k; ++j) ls.add_column(H + (size_t)j * ldh);
    ls.solve(y);
    return HG_OK;
}

extern "C" int hg_host_solve_square(int n, const double* M, int ld, const double* rhs, double* y) {
    HG_REQUIRE(M && rhs && y && n >= 1 && ld >= n, "hg_host_solve_square: bad argument");
    std::vector<double> W((size_t)n * n);
    for (int j = 0; j < n; ++j) memcpy(&W[(size_t)j * n], M + (size_t)j * ld, (size_t)n * 8);
    hgd::solve_square(n, W.data(), n, rhs, y);
    return HG_OK;
}

extern "C" int hg_host_singular_values(int n, const double* M, int ld, double* s) {
    HG_REQUIRE(M && s && n >= 1 && ld >= n, "hg_host_singular_values: bad argument");
    std::vector<double> W((size_t)n * n);
    for (int j = 0; j < n; ++j) memcpy(&W[(size_t)j * n], M + (size_t)j * ld, (size_t)n * 8);
    hgd::singular_values(n, W.data(), n, s);
    return HG_OK;
}

extern "C" int hg_gcv_eval(const hg_gcv* g, double lambda, double* gcv_val) {
    HG_REQUIRE(g && gcv_val, "hg_gcv_eval: NULL argument");
    *gcv_val = hgd::gcv_value(lambda, g->H.data(), g->k + 1, g->k, g->beta, g->trace_m, g->sv.data());
    return HG_OK;
}

extern "C" int hg_gcv_get(const hg_gcv* g, double* H, double* beta) {
    HG_REQUIRE(g, "hg_gcv_get: NULL argument");
    if (H) memcpy(H, g->H.data(), g->H.size() * 8);
    if (beta) *beta = g->beta;
    return HG_OK;
}

extern "C" int hg_gcv_fminbnd(const hg_gcv* g, double lo, double hi, double tolx, double* lambda,
                              double* fval, int* funccount, double* trace, int trace_cap) {
    HG_REQUIRE(g && lambda, "hg_gcv_fminbnd: NULL argument");
    auto f = [&](double l) {
        return hgd::gcv_value(l, g->H.data(), g->k + 1, g->k, g->beta, g->trace_m, g->sv.data());
    };
    hgd::FminResult r = hgd::fminbnd(f, lo, hi, tolx, 500, 500, trace, trace_cap);
    *lambda = r.x;
    if (fval) *fval = r.fval;
    if (funccount) *funccount = r.funccount;
    return HG_OK;
}

extern "C" int hg_gcv_destroy(hg_gcv* g) {
    delete g;
    return HG_OK;
}

// ===========================================================================
// Project-then-regularise (PTR) solvers — the solve path of ABgmres_hybrid_bounds.m:11-41,
// BAgmres_hybrid_bounds.m:11-40 and their non-hybrid siblings.  These
// are what every GCV-driven script calls after fminbnd (plot_error_vs_mismatch_norm.m:53-57,
// analyze_regularization.m:106-107).  Unshifted Arnoldi in m-space (AB: x = B*z) or n-space
// (BA); projected problem: Tikhonov on H (hybrid) or plain least squares.  The filter-factor
// bound outputs (phi, dphi; dense eig of A*B / B*A) are out of scope.
// ===========================================================================
// lambdas / nl / lambda_path non-null: the regularisation parameter is chosen at every iteration as the
// grid minimiser of GCV(lambda, H_k) (compute_gcv_surface + calculate_gcv_from_H, plot_gcv_surface.m:58-122)
//
// Like the RTP solvers the loop is software pipelined (Arnoldi step k+1 is queued before the host looks at step
// k) and no product with A or B is repeated: the Arnoldi process already formed A*q_j (BA) / B*q_j (AB), so
//   BA:  x_k = Q_k y_k,            b - A x_k = b - (A Q_k) y_k            (cached columns T = A Q)
//   AB:  x_k = B (Q_k y_k) = (B Q_k) y_k  (cached columns T = B Q),
//        b - A x_k = b - (A B Q_k) y_k = b - Q_{k+1} (H_k y_k)            (the Arnoldi relation, exact to the
//                                                                          rounding of the CGS2 step: 1e-15)
// i.e. 2 products per iteration instead of 3 (BA) or 4 (AB), and one launch for iterate + both history norms.
static int ptr_solver(hg_ctx* ctx, int kind, int hybrid, const hg_matrix* A, const hg_matrix* B,
                      const double* b, const double* x_true, double tol, int maxit, double lambda,
                      const double* lambdas, int nl, double* lambda_path,
                      double* x, double* error_norm, double* residual_norm, int* niters, int* x_valid,
                      hg_extras* extras) {
    HG_REQUIRE(ctx && A && B && b && x_true && x && error_norm && residual_norm && niters,
               "hg_gmres_ptr: NULL argument");
    HG_REQUIRE(kind == 0 || kind == 1, "hg_gmres_ptr: kind must be 0 (AB) or 1 (BA)");
    HG_REQUIRE(maxit >= 1, "hg_gmres_ptr: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = A->cols, m = A->rows;
    ArnoldiHolder holder;
    HG_TRY(arnoldi_create(ctx, A, B, kind == 0 ? HG_SPACE_M : HG_SPACE_N, maxit, true, &holder.a));
    hg_arnoldi* a = holder.a;
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    constexpr int RING = 4;
    DBuf d_x[2], d_xt, d_y, d_hy, stat_e;
    PinBuf h_y, h_hy, h_s;
    EventRing ev;
    HG_TRY(ev.create(RING));
    HG_TRY(d_x[0].alloc((size_t)n));
    HG_TRY(d_x[1].alloc((size_t)n));
    HG_TRY(d_xt.alloc((size_t)n));
    HG_TRY(d_y.alloc((size_t)maxit + 1));
    HG_TRY(d_hy.alloc((size_t)maxit + 2));
    HG_TRY(stat_e.alloc(hg_stat_capacity(ctx, std::max(n, m))));
    HG_TRY(h_y.alloc((size_t)RING * (maxit + 1)));
    HG_TRY(h_hy.alloc((size_t)RING * (maxit + 2)));
    HG_TRY(h_s.alloc((size_t)RING * 2));
    HG_CUDA(cudaMemsetAsync(d_x[0].p, 0, (size_t)n * 8, ctx->stream));
    HG_CUDA(cudaMemsetAsync(d_x[1].p, 0, (size_t)n * 8, ctx->stream));
    HG_CUDA(cudaMemcpyAsync(d_xt.p, x_true, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    HG_TRY(hg_arnoldi_set_rhs(a, b));
    double nb2 = 0, nx2 = 0;
    HG_TRY(hg_norm2_sync(ctx, a->d_b, m, &nb2));
    HG_TRY(hg_norm2_sync(ctx, d_xt.p, n, &nx2));
    const double norm_b = std::sqrt(nb2), norm_xt = std::sqrt(nx2);
    HG_TRY(hg_arnoldi_reset(a, 0.0));  // r0 = b (AB, :11-12) or B*b (BA, :12-13); unshifted operator (:25)
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    const double beta = a->h_beta[0];
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    hgd::HessenbergLS ls;
    hgd::BorderedCholesky chol;
    std::vector<double> rhs(maxit, 0.0), grow(maxit + 1, 0.0);
    bool chol_ok = true;
    if (hybrid) chol.reset(maxit, lambda);
    else ls.reset(maxit, beta);
    const int ldh = a->ldh();
    const double* H = a->h_H;
    int enq = 0, last_x = 0;
    auto finish_iterate = [&](int j, bool* stop) -> int {
        HG_CUDA(cudaEventSynchronize(ev.e[j % RING]));
        const double* hs = h_s.p + (size_t)(j % RING) * 2;
        error_norm[j - 1] = hs[0] / norm_xt;     // :41
        residual_norm[j - 1] = hs[1] / norm_b;   // :40
        last_x = j;
        *stop = residual_norm[j - 1] <= tol;     // :83
        return HG_OK;
    };
    int k;
    bool ended_early = false;
    for (k = 1; k <= maxit; ++k) {
        while (enq < std::min(k + 1, maxit)) {
            ++enq;
            HG_TRY(hg_arnoldi_steps(a, 1));
            if (enq == 1) HG_CUDA(cudaEventRecord(ev.e[0], ctx->stream));
        }
        if (k == 1) {
            HG_CUDA(cudaEventSynchronize(ev.e[0]));
        } else {
            bool stop = false;
            HG_TRY(finish_iterate(k - 1, &stop));
            if (stop) {
                k = k - 1;
                ended_early = true;
                break;
            }
        }
        const double* hcol = H + (size_t)(k - 1) * ldh;
        if (hcol[k] == 0.0) {  // :31
            ended_early = true;
            break;
        }
        double* yk = h_y.p + (size_t)(k % RING) * (maxit + 1);
        if (hybrid && nl > 0) {
            // lambda_k = first grid minimiser of GCV(lambda, H_k)        (plot_gcv_surface.m:92-100, :104-122)
            std::vector<double> sq((size_t)k * k), sv(k);
            for (int j = 0; j < k; ++j)
                for (int i2 = 0; i2 < k; ++i2) sq[(size_t)j * k + i2] = H[(size_t)j * ldh + i2];
            hgd::singular_values(k, sq.data(), k, sv.data());  // :111
            const double trace_m = kind == 0 ? (double)m : (double)n;  // op_size, :66,70
            int best = 0;
            double vbest = 0.0;
            for (int i2 = 0; i2 < nl; ++i2) {
                const double v = hgd::gcv_value(lambdas[i2], H, ldh, k, beta, trace_m, sv.data());
                if (i2 == 0 || v < vbest) {
                    vbest = v;
                    best = i2;
                }
            }
            const double lam_k = lambdas[best];
            lambda_path[k - 1] = lam_k;
            // yk = (Hk'*Hk + lambda_k*I) \ (Hk'*tk), from scratch: lambda changes with k
            std::vector<double> M((size_t)k * k);
            for (int j = 0; j < k; ++j)
                for (int i2 = 0; i2 < k; ++i2) {
                    double acc = 0.0;
                    for (int t = 0; t <= k; ++t) acc += H[(size_t)i2 * ldh + t] * H[(size_t)j * ldh + t];
                    M[(size_t)j * k + i2] = acc + (i2 == j ? lam_k : 0.0);
                }
            for (int j = 0; j < k; ++j) rhs[j] = beta * H[(size_t)j * ldh];
            hgd::solve_square(k, M.data(), k, rhs.data(), yk);
        } else if (hybrid) {
            // yk = (Hk'*Hk + lambda*I) \ (Hk'*tk): Hk'*Hk grows by bordering (column k adds row k+1,
            // which is zero in every earlier column), so the row Cholesky is continued    (:34-36)
            for (int j = 0; j < k; ++j) {
                double acc = 0.0;
                const double* hj = H + (size_t)j * ldh;
                for (int t = 0; t <= std::min(j, k - 1) + 1; ++t) acc += hj[t] * hcol[t];
                grow[j] = acc;
            }
            rhs[k - 1] = beta * hcol[0];
            if (chol_ok) chol_ok = chol.add_row(grow.data());
            if (chol_ok) {
                chol.solve(rhs.data(), yk);
            } else {
                std::vector<double> M((size_t)k * k);
                for (int j = 0; j < k; ++j)
                    for (int i2 = 0; i2 < k; ++i2) {
                        double acc = 0.0;
                        for (int t = 0; t <= k; ++t) acc += H[(size_t)i2 * ldh + t] * H[(size_t)j * ldh + t];
                        M[(size_t)j * k + i2] = acc + (i2 == j ? lambda : 0.0);
                    }
                hgd::solve_square(k, M.data(), k, rhs.data(), yk);
            }
        } else {
            ls.add_column(hcol);  // yk = Hk \ [beta;0]   (ABgmres_nonhybrid_bounds.m:34-35)
            ls.solve(yk);
        }
        HG_CUDA(cudaMemcpyAsync(d_y.p, yk, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
        double* xk = d_x[k & 1].p;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(ctx->d_scalars + 40);
        if (kind == 1) {
            // xk = Q(:,1:k)*yk (BA :37); norm(b - A*xk) (:40) from the cached columns A*q_j
            HG_TRY(hg_k_iterate2(ctx, a->Q, a->ldq, n, k, d_y.p, xk, d_xt.p, a->T + a->ldt, a->ldt, m, k, d_y.p, a->d_b,
                                 stat_e.p, ticket, ctx->d_scalars + 1));
        } else {
            // xk = B*(Q(:,1:k)*yk) (AB :37-38) from the cached columns B*q_j; b - A*xk = b - Q(:,1:k+1)*(Hk*yk)
            double* hy = h_hy.p + (size_t)(k % RING) * (maxit + 2);
            for (int t = 0; t <= k; ++t) {
                double acc = 0.0;
                for (int j = std::max(0, t - 1); j < k; ++j) acc += H[(size_t)j * ldh + t] * yk[j];  // Hk is upper Hessenberg
                hy[t] = acc;
            }
            HG_CUDA(cudaMemcpyAsync(d_hy.p, hy, (size_t)(k + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            HG_TRY(hg_k_iterate2(ctx, a->T + a->ldt, a->ldt, n, k, d_y.p, xk, d_xt.p, a->Q, a->ldq, m, k + 1, d_hy.p,
                                 a->d_b, stat_e.p, ticket, ctx->d_scalars + 1));
        }
        HG_CUDA(cudaMemcpyAsync(h_s.p + (size_t)(k % RING) * 2, ctx->d_scalars + 1, 16, cudaMemcpyDeviceToHost,
                                ctx->stream));
        if (extras && extras->X_hist)
            HG_CUDA(cudaMemcpyAsync(extras->X_hist + (size_t)(k - 1) * n, xk, (size_t)n * 8,
                                    cudaMemcpyDeviceToHost, ctx->stream));
        HG_CUDA(cudaEventRecord(ev.e[k % RING], ctx->stream));
    }
    if (!ended_early) {
        k = maxit;
        bool stop = false;
        HG_TRY(finish_iterate(maxit, &stop));
    }
    *niters = k;
    HG_CUDA(cudaMemcpyAsync(x, d_x[last_x & 1].p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));  // also drains a discarded speculative step
    if (x_valid) *x_valid = last_x > 0 ? 1 : 0;  // `x = xk` (:86) is undefined after a breakdown at k = 1
    if (extras) {
        if (extras->beta) *extras->beta = beta;
        if (extras->H) {  // columns the reference never reached stay zero
            memset(extras->H, 0, (size_t)ldh * maxit * 8);
            memcpy(extras->H, a->h_H, (size_t)ldh * k * 8);
        }
        if (extras->X_hist && last_x < maxit)
            memset(extras->X_hist + (size_t)last_x * n, 0, (size_t)(maxit - last_x) * n * 8);
    }
    if (lambda_path)
        for (int j = k; j < maxit; ++j) lambda_path[j] = 0.0;  // a discarded speculative iteration chose one more
    return HG_OK;
}

extern "C" int hg_gmres_ptr(hg_ctx* ctx, int kind, int hybrid, const hg_matrix* A, const hg_matrix* B,
                            const double* b, const double* x_true, double tol, int maxit, double lambda,
                            double* x, double* error_norm, double* residual_norm, int* niters, int* x_valid,
                            hg_extras* extras) {
    return ptr_solver(ctx, kind, hybrid, A, B, b, x_true, tol, maxit, lambda, nullptr, 0, nullptr, x, error_norm,
                      residual_norm, niters, x_valid, extras);
}

extern "C" int hg_gmres_ptr_gcv(hg_ctx* ctx, int kind, const hg_matrix* A, const hg_matrix* B, const double* b,
                                const double* x_true, double tol, int maxit, const double* lambdas, int nl,
                                double* x, double* error_norm, double* residual_norm, double* lambda_path,
                                int* niters, int* x_valid, hg_extras* extras) {
    HG_REQUIRE(lambdas && nl >= 1 && lambda_path, "hg_gmres_ptr_gcv: a lambda grid and lambda_path are required");
    for (int i = 0; i < maxit; ++i) lambda_path[i] = 0.0;
    return ptr_solver(ctx, kind, 1, A, B, b, x_true, tol, maxit, 0.0, lambdas, nl, lambda_path, x, error_norm,
                      residual_norm, niters, x_valid, extras);
}

// plot_gcv_surface.m:58-102 (compute_gcv_surface) given the Arnoldi factorisation already held by
// the handle: GCV over a (lambda, k) grid and the per-iteration grid minimiser lambda_k — the
// "Arnoldi once, GCV for many (lambda,k)" form.  surface is nl x k
// column-major (zero in the columns from a `H(k+1,k) < 1e-12` breakdown on, as the reference
// leaves them, :85), path has k entries.
extern "C" int hg_gcv_surface(const hg_gcv* g, const double* lambdas, int nl, double* surface, double* path) {
    HG_REQUIRE(g && lambdas && surface && path && nl >= 1, "hg_gcv_surface: bad argument");
    const int K = g->k, ldh = g->k + 1;
    for (size_t i = 0; i < (size_t)nl * K; ++i) surface[i] = 0.0;
    for (int k = 0; k < K; ++k) path[k] = 0.0;
    std::vector<double> sq, sv;
    for (int k = 1; k <= K; ++k) {
        if (g->H[(size_t)(k - 1) * ldh + k] < 1e-12) break;  // :85
        sq.assign((size_t)k * k, 0.0);
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < k; ++i) sq[(size_t)j * k + i] = g->H[(size_t)j * ldh + i];
        sv.assign(k, 0.0);
        hgd::singular_values(k, sq.data(), k, sv.data());  // :111
        int best = 0;
        for (int i = 0; i < nl; ++i) {
            const double v = hgd::gcv_value(lambdas[i], g->H.data(), ldh, k, g->beta, g->trace_m, sv.data());
            surface[(size_t)(k - 1) * nl + i] = v;
            if (v < surface[(size_t)(k - 1) * nl + best]) best = i;  // first minimum, as MATLAB's min (:99)
        }
        path[k - 1] = lambdas[best];
    }
    return HG_OK;
}
