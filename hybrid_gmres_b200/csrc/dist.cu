// Multi-GPU (one process per GPU) sharded Arnoldi — SURVEY.md §8(e).
//
// Rank p owns a contiguous, nnz-balanced detector-row block A_p (m_p x n) of A, the
// matching column block B^p (n x m_p) of B, the row slice [p*n_p, (p+1)*n_p) of every
// Krylov vector and the m_p-slice of T = [b, A q_1, ...].  Per step:
//     u_p = A_p q            local SpMV (q replicated)
//     w   = sum_p B^p u_p    local SpMV -> ncclReduceScatter of the n-vector
//     w  += shift * q        on the slice
//     CGS2 on slices         two ncclAllReduce of k coefficients, one of the norm
//     q_{k+1}                ncclAllGather of the normalised slice
// NCCL is resolved at run time with dlopen (the single-GPU library has no NCCL
// dependency); the communicator is built from a unique id that the Python side
// broadcasts with torch.distributed (plumbing only).
#include <dlfcn.h>

#include <algorithm>

#include "dist_internal.h"

namespace {

typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclChar = 0, ncclDouble = 8 };  // ncclInt8, ncclFloat64
enum { ncclSum = 0, ncclMin = 3 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} g_nccl;

int load_nccl() {
    if (g_nccl.lib) return HG_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        hg_set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
        return HG_ERR_NCCL;
    }
#define HG_SYM(field, name)                                                    \
    do {                                                                       \
        *(void**)(&g_nccl.field) = dlsym(lib, name);                           \
        if (!g_nccl.field) {                                                   \
            hg_set_error("NCCL: symbol %s not found", name);                   \
            return HG_ERR_NCCL;                                                \
        }                                                                      \
    } while (0)
    HG_SYM(GetUniqueId, "ncclGetUniqueId");
    HG_SYM(CommInitRank, "ncclCommInitRank");
    HG_SYM(CommDestroy, "ncclCommDestroy");
    HG_SYM(AllReduce, "ncclAllReduce");
    HG_SYM(ReduceScatter, "ncclReduceScatter");
    HG_SYM(AllGather, "ncclAllGather");
    HG_SYM(GetErrorString, "ncclGetErrorString");
#undef HG_SYM
    g_nccl.lib = lib;
    return HG_OK;
}

#define HG_NCCL(call)                                                                          \
    do {                                                                                       \
        int _r = (call);                                                                       \
        if (_r != ncclSuccess) {                                                               \
            hg_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
            return HG_ERR_NCCL;                                                                \
        }                                                                                      \
    } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace

int hg_nccl_allgather_bytes(hg_comm* c, const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t st) {
    HG_NCCL(g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclChar, c->comm, st));
    return HG_OK;
}
int hg_nccl_allreduce_min(hg_comm* c, double* d_buf, cudaStream_t st) {
    HG_NCCL(g_nccl.AllReduce(d_buf, d_buf, 1, ncclDouble, ncclMin, c->comm, st));
    return HG_OK;
}
int hg_nccl_barrier(hg_comm* c, cudaStream_t st) {
    HG_NCCL(g_nccl.AllReduce(c->ctx->d_scalars + 60, c->ctx->d_scalars + 60, 1, ncclDouble, ncclSum, c->comm, st));
    HG_CUDA(cudaStreamSynchronize(st));
    return HG_OK;
}

extern "C" int hg_comm_unique_id(void* out128) {
    HG_REQUIRE(out128, "hg_comm_unique_id: NULL");
    HG_TRY(load_nccl());
    ncclUniqueId id;
    HG_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return HG_OK;
}

extern "C" int hg_comm_init(hg_ctx* ctx, int nranks, int rank, const void* id128, hg_comm** out) {
    HG_REQUIRE(ctx && id128 && out, "hg_comm_init: NULL argument");
    HG_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "hg_comm_init: bad rank %d of %d", rank, nranks);
    HG_TRY(load_nccl());
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_comm* c = new (std::nothrow) hg_comm();
    if (!c) {
        hg_set_error("hg_comm_init: out of host memory");
        return HG_ERR_NOMEM;
    }
    c->rank = rank;
    c->nranks = nranks;
    c->ctx = ctx;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        hg_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return HG_ERR_NCCL;
    }
    *out = c;
    return HG_OK;
}

extern "C" int hg_comm_destroy(hg_comm* c) {
    if (!c) return HG_OK;
    hg_peer_destroy(c);  // collective when a peer workspace exists
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return HG_OK;
}

/* 0: NCCL collectives between the kernels, 1: NVLink peer memory inside the kernels */
extern "C" int hg_comm_transport(hg_comm* c, int* transport, char* why, int why_len) {
    HG_REQUIRE(c && transport, "hg_comm_transport: NULL argument");
    *transport = c->transport;
    if (why && why_len > 0) snprintf(why, (size_t)why_len, "%s", c->why);
    return HG_OK;
}

// ===========================================================================
struct hg_darnoldi {
    hg_ctx* ctx = nullptr;
    hg_comm* comm = nullptr;
    const hg_matrix* A = nullptr;  // m_p x n
    const hg_matrix* B = nullptr;  // n x m_p
    int kmax = 0, k = 0;
    int64_t n = 0, n_p = 0, n_pad = 0, m_p = 0, ldq = 0, ldt = 0;
    double* Q = nullptr;       // n_p x (kmax+1) row slice of the basis (zero padded)
    double* T = nullptr;       // m_p x (kmax+1): column 0 = b_p, column k = A_p q_k
    double* q_full = nullptr;  // n_pad, replicated current basis vector
    double* w_part = nullptr;  // n_pad, partial B^p u_p
    double *w0 = nullptr, *w1 = nullptr;  // n_p
    double *d_H = nullptr, *d_hcur = nullptr, *d_s = nullptr;
    double *h_H = nullptr, *h_beta = nullptr;
    double *partials = nullptr, *stat = nullptr;
    double shift = 0.0;
    bool have_rhs = false, started = false;
    bool peer = false;  // NVLink peer-memory transport (dist_peer.cu): q_full / w_part live in the workspace
    int qbuf = 0;       // which half of the double-buffered replicated q holds the current vector
    int64_t row0 = 0;   // first global row of this rank's slice
    int ldh() const { return kmax + 1; }
};

int hg_multidot_nslabs(const hg_ctx* ctx, int64_t n);

extern "C" int hg_darnoldi_destroy(hg_darnoldi* a) {
    if (!a) return HG_OK;
    cudaStreamSynchronize(a->ctx->stream);
    if (a->peer) hg_peer_release(a->comm, a);
    else { hg_dfree(a->q_full); hg_dfree(a->w_part); }
    hg_dfree(a->Q); hg_dfree(a->T);
    hg_dfree(a->w0); hg_dfree(a->w1); hg_dfree(a->d_H); hg_dfree(a->d_hcur); hg_dfree(a->d_s);
    hg_dfree(a->partials); hg_dfree(a->stat);
    hg_hfree(a->h_H);
    hg_hfree(a->h_beta);
    delete a;
    return HG_OK;
}

extern "C" int hg_darnoldi_create(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* B_p,
                                  int kmax, hg_darnoldi** out) {
    HG_REQUIRE(ctx && comm && A_p && B_p && out, "hg_darnoldi_create: NULL argument");
    HG_REQUIRE(kmax >= 1, "hg_darnoldi_create: kmax must be >= 1");
    HG_REQUIRE(A_p->cols == B_p->rows && A_p->rows == B_p->cols,
               "hg_darnoldi_create: shard shapes do not match (A_p %lld x %lld, B^p %lld x %lld)",
               (long long)A_p->rows, (long long)A_p->cols, (long long)B_p->rows, (long long)B_p->cols);
    HG_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    hg_darnoldi* a = new (std::nothrow) hg_darnoldi();
    if (!a) {
        hg_set_error("hg_darnoldi_create: out of host memory");
        return HG_ERR_NOMEM;
    }
    a->ctx = ctx;
    a->comm = comm;
    a->A = A_p;
    a->B = B_p;
    a->kmax = kmax;
    a->n = A_p->cols;
    a->m_p = A_p->rows;
    const int P = comm->nranks;
    a->n_p = round_up((a->n + P - 1) / P, 32);  // equal, 256-byte aligned slices (NCCL needs equal counts)
    a->n_pad = a->n_p * P;
    a->ldq = a->n_p;
    a->ldt = round_up(std::max<int64_t>(a->m_p, 1), 32);
    const int nslabs = hg_multidot_nslabs(ctx, std::max(a->n_p, a->m_p));
    cudaError_t e = cudaSuccess;
    auto alloc = [&](double** p, size_t cnt) {
        if (e == cudaSuccess) e = hg_dmalloc(ctx, p, std::max<size_t>(cnt, 1) * sizeof(double));
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, std::max<size_t>(cnt, 1) * sizeof(double), ctx->stream);
    };
    a->row0 = (int64_t)comm->rank * a->n_p;
    a->peer = hg_peer_acquire(comm, a->n_pad, kmax, a);  // collective; every rank takes the same decision
    if (!a->peer && hg_dist_transport_wanted() == 2) {
        hg_set_error("hg_darnoldi_create: peer-memory transport requested (HG_DIST=peer) but unavailable: %s", comm->why);
        hg_darnoldi_destroy(a);
        return HG_ERR_STATE;
    }
    alloc(&a->Q, (size_t)a->ldq * (kmax + 1));
    alloc(&a->T, (size_t)a->ldt * (kmax + 1));
    if (!a->peer) {
        alloc(&a->q_full, (size_t)a->n_pad);
        alloc(&a->w_part, (size_t)a->n_pad);
    }
    alloc(&a->w0, (size_t)a->n_p);
    alloc(&a->w1, (size_t)a->n_p);
    alloc(&a->d_H, (size_t)a->ldh() * kmax);
    alloc(&a->d_hcur, (size_t)kmax + 1);
    alloc(&a->d_s, 8);
    alloc(&a->partials, (size_t)(kmax + 2) * (size_t)(std::max(nslabs, ctx->sm_count) + 1));
    alloc(&a->stat, (size_t)std::max(a->n_pad, a->m_p) / 8 + 2048);
    if (e == cudaSuccess) e = hg_hmalloc(ctx, &a->h_H, (size_t)a->ldh() * kmax * sizeof(double));
    if (e == cudaSuccess) e = hg_hmalloc(ctx, &a->h_beta, 8 * sizeof(double));
    if (e != cudaSuccess) {
        hg_set_error("hg_darnoldi_create: allocation failed: %s", cudaGetErrorString(e));
        hg_darnoldi_destroy(a);
        return HG_ERR_NOMEM;
    }
    *out = a;
    return HG_OK;
}

extern "C" int hg_darnoldi_set_rhs(hg_darnoldi* a, const double* b_p) {
    HG_REQUIRE(a && (b_p || a->m_p == 0), "hg_darnoldi_set_rhs: NULL argument");
    HG_CUDA(cudaSetDevice(a->ctx->device));
    if (a->m_p)
        HG_CUDA(cudaMemcpyAsync(a->T, b_p, (size_t)a->m_p * 8, cudaMemcpyHostToDevice, a->ctx->stream));
    HG_CUDA(cudaStreamSynchronize(a->ctx->stream));
    a->have_rhs = true;
    return HG_OK;
}

// global norm of a slice vector: d_s[0] <- sqrt(allreduce(sum v^2)); then v /= d_s[0] and all-gather
static int normalise_and_gather(hg_darnoldi* a, double* v_slice, double* norm_out_dev) {
    hg_ctx* ctx = a->ctx;
    cudaStream_t st = ctx->stream;
    HG_NCCL(g_nccl.AllReduce(a->d_s, a->d_s, 1, ncclDouble, ncclSum, a->comm->comm, st));
    HG_TRY(hg_k_reduce(ctx, a->d_s, 1, 1, norm_out_dev, false, nullptr, true));  // sqrt of the global sum
    HG_TRY(hg_k_scale_div(ctx, v_slice, a->n_p, norm_out_dev));
    HG_NCCL(g_nccl.AllGather(v_slice, a->q_full, (size_t)a->n_p, ncclDouble, a->comm->comm, st));
    return HG_OK;
}

extern "C" int hg_darnoldi_reset(hg_darnoldi* a, double shift) {
    HG_REQUIRE(a, "hg_darnoldi_reset: NULL");
    if (!a->have_rhs) {
        hg_set_error("hg_darnoldi_reset: right-hand side not set");
        return HG_ERR_STATE;
    }
    hg_ctx* ctx = a->ctx;
    cudaStream_t st = ctx->stream;
    HG_CUDA(cudaSetDevice(ctx->device));
    a->shift = shift;
    a->k = 0;
    HG_CUDA(cudaStreamSynchronize(st));
    memset(a->h_H, 0, (size_t)a->ldh() * a->kmax * sizeof(double));
    HG_CUDA(cudaMemsetAsync(a->d_H, 0, (size_t)a->ldh() * a->kmax * sizeof(double), st));
    // r0 = B*b = sum_p B^p b_p  (hybrid_ba_gmres_rtp.m:7-9)
    hg_spmv_epilogue ep;
    if (a->peer) {
        hg_comm* c = a->comm;
        int ns = 0, npp = 0;
        a->qbuf = 0;
        HG_TRY(hg_k_spmv(ctx, a->B, a->T, hg_peer_ypart(c), ep, nullptr));
        // r0 slice (pulled) goes to every rank's replicated vector un-normalised; the norm
        // all-reduce is also the barrier for those stores; then everybody scales locally
        HG_TRY(hg_k_pull_multidot(c, a->row0, nullptr, 0.0, a->Q, a->Q, a->ldq, a->n_p, 0, a->partials, &ns, true));
        hg_out_list push;
        hg_peer_push_list(c, a->qbuf, a->row0, &push);
        HG_TRY(hg_k_lincomb_push(ctx, a->Q, a->ldq, a->n_p, 0, a->d_hcur, 1.0, a->Q, nullptr, nullptr, a->stat, &npp, &push));
        HG_TRY(hg_k_reduce_allreduce(c, a->stat, npp, 1, a->d_s + 1, nullptr, false, true, true));  // beta
        HG_TRY(hg_k_scale2(c, hg_peer_qfull(c, a->qbuf), a->n_pad, a->Q, a->n_p, a->d_s + 1));
        HG_CUDA(cudaMemcpyAsync(a->h_beta, a->d_s + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        a->started = true;
        return HG_OK;
    }
    HG_TRY(hg_k_spmv(ctx, a->B, a->T, a->w_part, ep, nullptr));
    HG_NCCL(g_nccl.ReduceScatter(a->w_part, a->Q, (size_t)a->n_p, ncclDouble, ncclSum, a->comm->comm, st));
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, a->Q, a->n_p, a->stat, &np));
    HG_TRY(hg_k_reduce(ctx, a->stat, np, 1, a->d_s, false, nullptr, false));
    HG_TRY(normalise_and_gather(a, a->Q, a->d_s + 1));  // beta in d_s[1]
    HG_CUDA(cudaMemcpyAsync(a->h_beta, a->d_s + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
    a->started = true;
    return HG_OK;
}

static int darnoldi_step(hg_darnoldi* a, int kk) {
    hg_ctx* ctx = a->ctx;
    cudaStream_t st = ctx->stream;
    const double* q_slice = a->Q + (size_t)(kk - 1) * a->ldq;
    double* tcol = a->T + (size_t)kk * a->ldt;
    double* qnext = a->Q + (size_t)kk * a->ldq;
    double* Hcol = a->d_H + (size_t)(kk - 1) * a->ldh();
    hg_spmv_epilogue ep;
    if (a->peer) {
        // the collectives run inside the kernels, over NVLink peer memory (dist_peer.cu)
        hg_comm* c = a->comm;
        int ns = 0, np = 0;
        HG_TRY(hg_k_spmv(ctx, a->A, hg_peer_qfull(c, a->qbuf), tcol, ep, nullptr));  // u_p = A_p q
        HG_TRY(hg_k_spmv(ctx, a->B, tcol, hg_peer_ypart(c), ep, nullptr));           // partial B^p u_p
        if (hg_cgs2_step_eligible_dist(ctx, a->n_p, kk)) {
            HG_TRY(hg_k_peer_signal(c, HG_FLAG_Y));
            // slices of a few hundred thousand rows: reduce-scatter, CGS2 with its three all-reduces, all-gather
            // and normalisation in ONE persistent kernel (cgs2_step.cu) — 4 launches per step instead of 10
            a->qbuf ^= 1;
            HG_TRY(hg_k_cgs2_step_peer(c, a->Q, a->ldq, a->n_p, kk, a->w0, a->w1, qnext, Hcol, a->d_hcur, a->partials,
                                       a->row0, q_slice, a->shift, a->qbuf));
            HG_CUDA(cudaMemcpyAsync(a->h_H + (size_t)(kk - 1) * a->ldh(), Hcol, (size_t)(kk + 1) * 8,
                                    cudaMemcpyDeviceToHost, st));
            return HG_OK;
        }
        // "B^p u_p complete" to every peer, then w0 = sum_p (B^p u_p)[slice] + shift*q[slice], fused with h1 = Q_k' w0
        HG_TRY(hg_k_pull_multidot(c, a->row0, q_slice, a->shift, a->w0, a->Q, a->ldq, a->n_p, kk, a->partials, &ns, true));
        HG_TRY(hg_k_reduce_allreduce(c, a->partials, ns, kk, a->d_hcur, Hcol, false, false));
        if (hg_cgs_fused_mode() == 2 && hg_cgs_staged_nparts(ctx, a->n_p, kk) > 0) {
            HG_TRY(hg_k_cgs_mid_staged(ctx, a->Q, a->ldq, a->n_p, kk, a->d_hcur, a->w0, a->w1, a->partials, &ns));
        } else {
            HG_TRY(hg_k_lincomb_push(ctx, a->Q, a->ldq, a->n_p, kk, a->d_hcur, -1.0, a->w0, a->w1, nullptr, nullptr, nullptr,
                                     nullptr));
            HG_TRY(hg_k_multidot(ctx, a->Q, a->ldq, a->n_p, kk, a->w1, a->partials, &ns));
        }
        HG_TRY(hg_k_reduce_allreduce(c, a->partials, ns, kk, a->d_hcur, Hcol, true, false));  // H(1:k,k) = h1 + h2
        // v = w1 - Q h2: rows go to the local basis AND to every rank's replicated vector (all-gather
        // by the producing kernel); the norm all-reduce is the barrier for them
        a->qbuf ^= 1;
        hg_out_list push;
        hg_peer_push_list(c, a->qbuf, a->row0, &push);
        HG_TRY(hg_k_lincomb_push(ctx, a->Q, a->ldq, a->n_p, kk, a->d_hcur, -1.0, a->w1, qnext, nullptr, a->stat, &np, &push));
        HG_TRY(hg_k_reduce_allreduce(c, a->stat, np, 1, Hcol + kk, nullptr, false, true, true));  // H(k+1,k) = norm(v)
        HG_TRY(hg_k_scale2(c, hg_peer_qfull(c, a->qbuf), a->n_pad, qnext, a->n_p, Hcol + kk));  // q_{k+1} = v / H(k+1,k)
        HG_CUDA(cudaMemcpyAsync(a->h_H + (size_t)(kk - 1) * a->ldh(), Hcol, (size_t)(kk + 1) * 8,
                                cudaMemcpyDeviceToHost, st));
        return HG_OK;
    }
    HG_TRY(hg_k_spmv(ctx, a->A, a->q_full, tcol, ep, nullptr));       // u_p = A_p q
    HG_TRY(hg_k_spmv(ctx, a->B, tcol, a->w_part, ep, nullptr));       // partial B^p u_p
    HG_NCCL(g_nccl.ReduceScatter(a->w_part, a->w1, (size_t)a->n_p, ncclDouble, ncclSum, a->comm->comm, st));
    // w0 = w + shift*q on the slice
    HG_TRY(hg_k_axpby(ctx, a->n_p, 1.0, a->w1, a->shift, q_slice, a->w0, nullptr, nullptr, nullptr));
    int ns = 0, np = 0;
    // CGS2 pass 1
    HG_TRY(hg_k_multidot(ctx, a->Q, a->ldq, a->n_p, kk, a->w0, a->partials, &ns));
    HG_TRY(hg_k_reduce(ctx, a->partials, ns, kk, a->d_hcur, false, nullptr, false));
    HG_NCCL(g_nccl.AllReduce(a->d_hcur, a->d_hcur, (size_t)kk, ncclDouble, ncclSum, a->comm->comm, st));
    HG_CUDA(cudaMemcpyAsync(Hcol, a->d_hcur, (size_t)kk * 8, cudaMemcpyDeviceToDevice, st));
    HG_TRY(hg_k_lincomb(ctx, a->Q, a->ldq, a->n_p, kk, a->d_hcur, -1.0, a->w0, a->w1, nullptr, nullptr, nullptr));
    // pass 2
    HG_TRY(hg_k_multidot(ctx, a->Q, a->ldq, a->n_p, kk, a->w1, a->partials, &ns));
    HG_TRY(hg_k_reduce(ctx, a->partials, ns, kk, a->d_hcur, false, nullptr, false));
    HG_NCCL(g_nccl.AllReduce(a->d_hcur, a->d_hcur, (size_t)kk, ncclDouble, ncclSum, a->comm->comm, st));
    HG_TRY(hg_k_axpby(ctx, kk, 1.0, Hcol, 1.0, a->d_hcur, Hcol, nullptr, nullptr, nullptr));  // H(1:k,k) = h1 + h2
    HG_TRY(hg_k_lincomb(ctx, a->Q, a->ldq, a->n_p, kk, a->d_hcur, -1.0, a->w1, qnext, nullptr, a->stat, &np));
    HG_TRY(hg_k_reduce(ctx, a->stat, np, 1, a->d_s, false, nullptr, false));
    HG_TRY(normalise_and_gather(a, qnext, Hcol + kk));
    HG_CUDA(cudaMemcpyAsync(a->h_H + (size_t)(kk - 1) * a->ldh(), Hcol, (size_t)(kk + 1) * 8,
                            cudaMemcpyDeviceToHost, st));
    return HG_OK;
}

extern "C" int hg_darnoldi_steps(hg_darnoldi* a, int nsteps) {
    HG_REQUIRE(a, "hg_darnoldi_steps: NULL");
    if (!a->started) {
        hg_set_error("hg_darnoldi_steps: call hg_darnoldi_reset first");
        return HG_ERR_STATE;
    }
    HG_REQUIRE(nsteps >= 0 && a->k + nsteps <= a->kmax, "hg_darnoldi_steps: exceeds kmax");
    HG_CUDA(cudaSetDevice(a->ctx->device));
    for (int i = 0; i < nsteps; ++i) {
        HG_TRY(darnoldi_step(a, a->k + 1));
        a->k += 1;
    }
    return HG_OK;
}

extern "C" int hg_darnoldi_get(hg_darnoldi* a, double* H, int ldh, double* beta, int* ksteps) {
    HG_REQUIRE(a, "hg_darnoldi_get: NULL");
    HG_CUDA(cudaStreamSynchronize(a->ctx->stream));
    if (a->peer) HG_TRY(hg_peer_check(a->comm));
    if (H) {
        HG_REQUIRE(ldh >= a->ldh(), "hg_darnoldi_get: ldh too small");
        for (int j = 0; j < a->kmax; ++j)
            memcpy(H + (size_t)j * ldh, a->h_H + (size_t)j * a->ldh(), (size_t)a->ldh() * 8);
    }
    if (beta) *beta = a->h_beta[0];
    if (ksteps) *ksteps = a->k;
    return HG_OK;
}

// local slice (n_p entries, zero padded) of basis vector j; *row0 receives its first global row
extern "C" int hg_darnoldi_get_q(hg_darnoldi* a, int j, double* q_slice, int64_t* row0, int64_t* nrows) {
    HG_REQUIRE(a && q_slice, "hg_darnoldi_get_q: NULL");
    HG_REQUIRE(j >= 0 && j <= a->kmax, "hg_darnoldi_get_q: column out of range");
    HG_CUDA(cudaMemcpyAsync(q_slice, a->Q + (size_t)j * a->ldq, (size_t)a->n_p * 8, cudaMemcpyDeviceToHost,
                            a->ctx->stream));
    HG_CUDA(cudaStreamSynchronize(a->ctx->stream));
    if (row0) *row0 = (int64_t)a->comm->rank * a->n_p;
    if (nrows) *nrows = a->n_p;
    return HG_OK;
}

extern "C" int hg_darnoldi_step_bytes(hg_darnoldi* a, int k, double* bytes) {
    HG_REQUIRE(a && bytes, "hg_darnoldi_step_bytes: NULL");
    // this rank's share of S(k) (hg_arnoldi_step_bytes) plus the replicated q / partial w vectors
    const double np = (double)a->n_p, mp = (double)a->m_p, n = (double)a->n_pad;
    const bool one_pass = a->peer && (hg_cgs2_step_eligible_dist(a->ctx, a->n_p, k) ||
                                      (hg_cgs_fused_mode() == 2 && hg_cgs_staged_nparts(a->ctx, a->n_p, k) > 0));
    *bytes = hg_spmv_stream_bytes(a->A) + hg_spmv_stream_bytes(a->B) + 16.0 * mp + 16.0 * n + 72.0 * np +
             (one_pass ? 24.0 : 32.0) * (double)k * np;
    return HG_OK;
}

// ===========================================================================
// Sharded RTP solvers: hybrid_ab_gmres_rtp.m / hybrid_ba_gmres_rtp.m on row-sharded A,
// column-sharded B.  The projected problem is solved redundantly on every rank from the
// all-reduced (bit-identical) H / Gram data, so all ranks take the same decisions.
// ===========================================================================
#include "dense_host.h"

namespace {
struct DBufD {
    double* p = nullptr;
    ~DBufD() { hg_dfree(p); }
    int alloc(size_t n) {
        if (hg_dmalloc_cur((void**)&p, std::max<size_t>(n, 1) * sizeof(double)) != cudaSuccess) {
            hg_set_error("device allocation of %zu doubles failed", n);
            return HG_ERR_NOMEM;
        }
        return HG_OK;
    }
};
struct PinD {
    double* p = nullptr;
    ~PinD() { hg_hfree(p); }
    int alloc(size_t n) {
        if (hg_hmalloc_cur((void**)&p, std::max<size_t>(n, 1) * sizeof(double)) != cudaSuccess) {
            hg_set_error("pinned allocation of %zu doubles failed", n);
            return HG_ERR_NOMEM;
        }
        return HG_OK;
    }
};
struct DHolder {
    hg_darnoldi* a = nullptr;
    ~DHolder() { hg_darnoldi_destroy(a); }
};
}  // namespace

// kind 0: AB-RTP, 1: BA-RTP.  b_p: this rank's m_p entries of b; x_true, x: full n-vectors
// (replicated on the host of every rank).
extern "C" int hg_dist_hybrid_rtp(int kind, hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p,
                                  const hg_matrix* B_p, const double* b_p, const double* x_true, double tol,
                                  int maxit, double lambda, double* x, double* error_norm,
                                  double* residual_norm, int* niters, int* x_valid, hg_extras* extras) {
    HG_REQUIRE(ctx && comm && A_p && B_p && x_true && x && error_norm && residual_norm && niters,
               "hg_dist_hybrid_rtp: NULL argument");
    HG_REQUIRE(kind == 0 || kind == 1, "hg_dist_hybrid_rtp: kind must be 0 (AB) or 1 (BA)");
    HG_REQUIRE(maxit >= 1, "hg_dist_hybrid_rtp: maxit must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DHolder holder;
    HG_TRY(hg_darnoldi_create(ctx, comm, A_p, B_p, maxit, &holder.a));
    hg_darnoldi* a = holder.a;
    const int64_t n = a->n, n_p = a->n_p, m_p = a->m_p;
    const int64_t row0 = (int64_t)comm->rank * n_p;
    const int64_t nloc = std::max<int64_t>(0, std::min(n, row0 + n_p) - row0);
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    // software-pipelined like the single-GPU solver (arnoldi.cu: rtp_solver): step k+1 is queued before the
    // host looks at step k; every rank takes the same decisions (the histories are bit-identical on all
    // ranks), so all ranks queue the same kernel sequence.
    constexpr int RING = 4;
    DBufD d_x[2], d_xt, d_y, d_g, stat_e, stat_r, d_xfull;
    PinD h_y, h_g, h_s;
    std::vector<cudaEvent_t> ev(RING, nullptr);
    struct EvGuard {
        std::vector<cudaEvent_t>& e;
        ~EvGuard() {
            for (cudaEvent_t x : e)
                if (x) cudaEventDestroy(x);
        }
    } ev_guard{ev};
    for (int i = 0; i < RING; ++i) HG_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    HG_TRY(d_x[0].alloc((size_t)n_p)); HG_TRY(d_x[1].alloc((size_t)n_p)); HG_TRY(d_xt.alloc((size_t)n_p));
    HG_TRY(d_y.alloc((size_t)maxit + 1));
    HG_TRY(d_g.alloc((size_t)maxit + 2)); HG_TRY(stat_e.alloc(hg_stat_capacity(ctx, std::max(n_p, m_p)) + 1024));
    HG_TRY(stat_r.alloc(hg_stat_capacity(ctx, std::max(n_p, m_p)) + 1024)); HG_TRY(d_xfull.alloc((size_t)a->n_pad));
    HG_TRY(h_y.alloc((size_t)RING * (maxit + 1))); HG_TRY(h_g.alloc((size_t)RING * (maxit + 2)));
    HG_TRY(h_s.alloc((size_t)RING * 2 + 4));
    double* h_norms = h_s.p + (size_t)RING * 2;
    HG_CUDA(cudaMemsetAsync(d_x[0].p, 0, (size_t)n_p * 8, st));
    HG_CUDA(cudaMemsetAsync(d_x[1].p, 0, (size_t)n_p * 8, st));
    HG_CUDA(cudaMemsetAsync(d_xt.p, 0, (size_t)n_p * 8, st));
    if (nloc > 0)
        HG_CUDA(cudaMemcpyAsync(d_xt.p, x_true + row0, (size_t)nloc * 8, cudaMemcpyHostToDevice, st));
    HG_TRY(hg_darnoldi_set_rhs(a, b_p));
    // global ||b||, ||x_true||
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, a->T, m_p, stat_e.p, &np));
    HG_TRY(hg_k_reduce(ctx, stat_e.p, np, 1, ctx->d_scalars + 3, false, nullptr, false));
    HG_TRY(hg_k_sumsq(ctx, d_xt.p, n_p, stat_r.p, &np));
    HG_TRY(hg_k_reduce(ctx, stat_r.p, np, 1, ctx->d_scalars + 4, false, nullptr, false));
    HG_NCCL(g_nccl.AllReduce(ctx->d_scalars + 3, ctx->d_scalars + 3, 2, ncclDouble, ncclSum, comm->comm, st));
    HG_CUDA(cudaMemcpyAsync(h_norms, ctx->d_scalars + 3, 16, cudaMemcpyDeviceToHost, st));
    HG_TRY(hg_darnoldi_reset(a, lambda));
    HG_CUDA(cudaStreamSynchronize(st));
    const double norm_b = std::sqrt(h_norms[0]), norm_xt = std::sqrt(h_norms[1]);
    const double beta = a->h_beta[0];
    for (int i = 0; i < maxit; ++i) error_norm[i] = residual_norm[i] = 0.0;
    hgd::HessenbergLS ls;
    hgd::BorderedCholesky chol;
    std::vector<double> Gfull, rhs;
    bool chol_ok = true;
    if (kind == 1) ls.reset(maxit, beta);
    else {
        chol.reset(maxit, lambda);
        Gfull.assign((size_t)maxit * maxit, 0.0);
        rhs.assign(maxit, 0.0);
    }
    const int ldh = a->ldh();
    int enq = 0, last_x = 0;
    auto queue_step = [&](int kk) -> int {
        HG_TRY(hg_darnoldi_steps(a, 1));
        if (kind == 0) {
            int ns = 0;
            HG_TRY(hg_k_multidot(ctx, a->T, a->ldt, m_p, kk + 1, a->T + (size_t)kk * a->ldt, a->partials, &ns));
            if (a->peer) {
                HG_TRY(hg_k_reduce_allreduce(comm, a->partials, ns, kk + 1, d_g.p, nullptr, false, false));
            } else {
                HG_TRY(hg_k_reduce(ctx, a->partials, ns, kk + 1, d_g.p, false, nullptr, false));
                HG_NCCL(g_nccl.AllReduce(d_g.p, d_g.p, (size_t)(kk + 1), ncclDouble, ncclSum, comm->comm, st));
            }
            HG_CUDA(cudaMemcpyAsync(h_g.p + (size_t)(kk % RING) * (maxit + 2), d_g.p, (size_t)(kk + 1) * 8,
                                    cudaMemcpyDeviceToHost, st));
        }
        if (kk == 1) HG_CUDA(cudaEventRecord(ev[0], st));
        return HG_OK;
    };
    auto finish_iterate = [&](int j, bool* stop) -> int {
        HG_CUDA(cudaEventSynchronize(ev[j % RING]));
        if (a->peer) HG_TRY(hg_peer_check(comm));  // a lost rank ends the solve here, not after maxit time-outs
        const double* hs = h_s.p + (size_t)(j % RING) * 2;
        error_norm[j - 1] = std::sqrt(hs[0]) / norm_xt;
        residual_norm[j - 1] = std::sqrt(hs[1]) / norm_b;
        last_x = j;
        *stop = residual_norm[j - 1] <= tol;
        return HG_OK;
    };
    int k;
    bool ended_early = false;
    for (k = 1; k <= maxit; ++k) {
        while (enq < std::min(k + 1, maxit)) {
            ++enq;
            HG_TRY(queue_step(enq));
        }
        if (k == 1) {
            HG_CUDA(cudaEventSynchronize(ev[0]));
        } else {
            bool stop = false;
            HG_TRY(finish_iterate(k - 1, &stop));
            if (stop) {
                k = k - 1;
                ended_early = true;
                break;
            }
        }
        const double* hcol = a->h_H + (size_t)(k - 1) * ldh;
        if (hcol[k] == 0.0) {
            ended_early = true;
            break;
        }
        double* yk = h_y.p + (size_t)(k % RING) * (maxit + 1);
        if (kind == 1) {
            ls.add_column(hcol);
            ls.solve(yk);
        } else {
            const double* g = h_g.p + (size_t)(k % RING) * (maxit + 2);
            rhs[k - 1] = g[0];
            for (int j = 0; j < k; ++j) {
                Gfull[(size_t)(k - 1) * maxit + j] = g[1 + j];
                Gfull[(size_t)j * maxit + (k - 1)] = g[1 + j];
            }
            if (chol_ok) chol_ok = chol.add_row(g + 1);
            if (chol_ok) chol.solve(rhs.data(), yk);
            else {
                std::vector<double> M((size_t)k * k);
                for (int j = 0; j < k; ++j)
                    for (int i2 = 0; i2 < k; ++i2)
                        M[(size_t)j * k + i2] = Gfull[(size_t)j * maxit + i2] + (i2 == j ? lambda : 0.0);
                hgd::solve_square(k, M.data(), k, rhs.data(), yk);
            }
        }
        HG_CUDA(cudaMemcpyAsync(d_y.p, yk, (size_t)k * 8, cudaMemcpyHostToDevice, st));
        double* xk = d_x[k & 1].p;
        int np_e = 0, np_r = 0;
        HG_TRY(hg_k_lincomb(ctx, a->Q, a->ldq, n_p, k, d_y.p, 1.0, nullptr, xk, d_xt.p, stat_e.p, &np_e));
        HG_TRY(hg_k_lincomb(ctx, a->T + a->ldt, a->ldt, m_p, k, d_y.p, -1.0, a->T, nullptr, nullptr, stat_r.p, &np_r));
        if (a->peer) {
            // error and residual sums of squares in one exchange
            HG_TRY(hg_k_reduce_allreduce(comm, stat_e.p, np_e, 2, ctx->d_scalars + 1, nullptr, false, false, false,
                                         stat_r.p, np_r));
        } else {
            HG_TRY(hg_k_reduce(ctx, stat_e.p, np_e, 1, ctx->d_scalars + 1, false, nullptr, false));
            HG_TRY(hg_k_reduce(ctx, stat_r.p, np_r, 1, ctx->d_scalars + 2, false, nullptr, false));
            HG_NCCL(g_nccl.AllReduce(ctx->d_scalars + 1, ctx->d_scalars + 1, 2, ncclDouble, ncclSum, comm->comm, st));
        }
        HG_CUDA(cudaMemcpyAsync(h_s.p + (size_t)(k % RING) * 2, ctx->d_scalars + 1, 16, cudaMemcpyDeviceToHost, st));
        HG_CUDA(cudaEventRecord(ev[k % RING], st));
    }
    if (!ended_early) {
        k = maxit;
        bool stop = false;
        HG_TRY(finish_iterate(maxit, &stop));
    }
    *niters = k;
    const bool have_x = (kind == 1) || last_x > 0;
    HG_NCCL(g_nccl.AllGather(d_x[last_x & 1].p, d_xfull.p, (size_t)n_p, ncclDouble, comm->comm, st));
    HG_CUDA(cudaMemcpyAsync(x, d_xfull.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaStreamSynchronize(st));
    if (a->peer) HG_TRY(hg_peer_check(comm));
    if (x_valid) *x_valid = have_x ? 1 : 0;
    if (extras) {
        if (extras->beta) *extras->beta = beta;
        if (extras->H) {
            memset(extras->H, 0, (size_t)ldh * maxit * 8);
            memcpy(extras->H, a->h_H, (size_t)ldh * k * 8);
        }
    }
    return HG_OK;
}

// ===========================================================================
// Sharded gcv_function (gcv_function.m:4-32 on the device, :33-58 on the host).
//   'ba' (n-space): the sharded Arnoldi above with shift 0 and the <1e-12 breakdown test.
//   'ab' (m-space): Q is sharded by the detector-row blocks m_p;
//        z = B q = sum_p B^p q_p  -> ncclAllReduce of the n-vector -> w_p = A_p z (local);
//        CGS2 on the m_p slices with all-reduced coefficients.  No all-gather is needed.
// ===========================================================================
extern "C" int hg_gcv_from_H(const double* H, int ldh, int k, double beta, double trace_m, hg_gcv** out);

static int dist_gcv_ab(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* B_p,
                       const double* b_p, int k_gcv, std::vector<double>& H, double* beta_out) {
    cudaStream_t st = ctx->stream;
    const int64_t m_p = A_p->rows, n = A_p->cols;
    const int64_t ldq = round_up(std::max<int64_t>(m_p, 1), 32);
    const int ldh = k_gcv + 1;
    const int nslabs = hg_multidot_nslabs(ctx, std::max<int64_t>(m_p, 1));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    DBufD Q, z, w0, w1, dH, hcur, ds, partials, stat;
    PinD hH;
    HG_TRY(Q.alloc((size_t)ldq * (k_gcv + 1))); HG_TRY(z.alloc((size_t)n)); HG_TRY(w0.alloc((size_t)ldq));
    HG_TRY(w1.alloc((size_t)ldq)); HG_TRY(dH.alloc((size_t)ldh * k_gcv)); HG_TRY(hcur.alloc((size_t)k_gcv + 1));
    HG_TRY(ds.alloc(8)); HG_TRY(partials.alloc((size_t)(k_gcv + 2) * (nslabs + 1)));
    HG_TRY(stat.alloc((size_t)std::max(m_p, n) / 8 + 2048)); HG_TRY(hH.alloc((size_t)ldh * k_gcv + 8));
    HG_CUDA(cudaMemsetAsync(Q.p, 0, (size_t)ldq * (k_gcv + 1) * 8, st));
    HG_CUDA(cudaMemsetAsync(dH.p, 0, (size_t)ldh * k_gcv * 8, st));
    memset(hH.p, 0, ((size_t)ldh * k_gcv + 8) * 8);
    // r0 = b; beta = norm(r0); Q(:,1) = r0/beta            (gcv_function.m:5,12,15)
    if (m_p) HG_CUDA(cudaMemcpyAsync(Q.p, b_p, (size_t)m_p * 8, cudaMemcpyHostToDevice, st));
    int np = 0;
    HG_TRY(hg_k_sumsq(ctx, Q.p, m_p, stat.p, &np));
    HG_TRY(hg_k_reduce(ctx, stat.p, np, 1, ds.p, false, nullptr, false));
    HG_NCCL(g_nccl.AllReduce(ds.p, ds.p, 1, ncclDouble, ncclSum, comm->comm, st));
    HG_TRY(hg_k_reduce(ctx, ds.p, 1, 1, ds.p + 1, false, nullptr, true));
    HG_TRY(hg_k_scale_div(ctx, Q.p, m_p, ds.p + 1));
    HG_CUDA(cudaMemcpyAsync(hH.p + (size_t)ldh * k_gcv, ds.p + 1, 8, cudaMemcpyDeviceToHost, st));
    for (int kk = 1; kk <= k_gcv; ++kk) {
        const double* q = Q.p + (size_t)(kk - 1) * ldq;
        double* qn = Q.p + (size_t)kk * ldq;
        double* Hcol = dH.p + (size_t)(kk - 1) * ldh;
        hg_spmv_epilogue ep;
        HG_TRY(hg_k_spmv(ctx, B_p, q, z.p, ep, nullptr));                     // partial B^p q_p
        HG_NCCL(g_nccl.AllReduce(z.p, z.p, (size_t)n, ncclDouble, ncclSum, comm->comm, st));
        HG_TRY(hg_k_spmv(ctx, A_p, z.p, w0.p, ep, nullptr));                  // v = A*(B*q)     (:20)
        int ns = 0;
        HG_TRY(hg_k_multidot(ctx, Q.p, ldq, m_p, kk, w0.p, partials.p, &ns));
        HG_TRY(hg_k_reduce(ctx, partials.p, ns, kk, hcur.p, false, nullptr, false));
        HG_NCCL(g_nccl.AllReduce(hcur.p, hcur.p, (size_t)kk, ncclDouble, ncclSum, comm->comm, st));
        HG_CUDA(cudaMemcpyAsync(Hcol, hcur.p, (size_t)kk * 8, cudaMemcpyDeviceToDevice, st));
        HG_TRY(hg_k_lincomb(ctx, Q.p, ldq, m_p, kk, hcur.p, -1.0, w0.p, w1.p, nullptr, nullptr, nullptr));
        HG_TRY(hg_k_multidot(ctx, Q.p, ldq, m_p, kk, w1.p, partials.p, &ns));
        HG_TRY(hg_k_reduce(ctx, partials.p, ns, kk, hcur.p, false, nullptr, false));
        HG_NCCL(g_nccl.AllReduce(hcur.p, hcur.p, (size_t)kk, ncclDouble, ncclSum, comm->comm, st));
        HG_TRY(hg_k_axpby(ctx, kk, 1.0, Hcol, 1.0, hcur.p, Hcol, nullptr, nullptr, nullptr));
        HG_TRY(hg_k_lincomb(ctx, Q.p, ldq, m_p, kk, hcur.p, -1.0, w1.p, qn, nullptr, stat.p, &np));
        HG_TRY(hg_k_reduce(ctx, stat.p, np, 1, ds.p, false, nullptr, false));
        HG_NCCL(g_nccl.AllReduce(ds.p, ds.p, 1, ncclDouble, ncclSum, comm->comm, st));
        HG_TRY(hg_k_reduce(ctx, ds.p, 1, 1, Hcol + kk, false, nullptr, true));     // H(k+1,k)    (:29)
        HG_TRY(hg_k_scale_div(ctx, qn, m_p, Hcol + kk));
        HG_CUDA(cudaMemcpyAsync(hH.p + (size_t)(kk - 1) * ldh, Hcol, (size_t)(kk + 1) * 8,
                                cudaMemcpyDeviceToHost, st));
        HG_CUDA(cudaStreamSynchronize(st));
        if (hH.p[(size_t)(kk - 1) * ldh + kk] < 1e-12) break;                 // :30 (same on all ranks)
    }
    HG_CUDA(cudaStreamSynchronize(st));
    H.assign(hH.p, hH.p + (size_t)ldh * k_gcv);
    *beta_out = hH.p[(size_t)ldh * k_gcv];
    return HG_OK;
}

extern "C" int hg_dist_gcv_prepare(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* B_p,
                                   const double* b_p, int64_t m, int k_gcv, int gcv_type, hg_gcv** out) {
    HG_REQUIRE(ctx && comm && A_p && B_p && out, "hg_dist_gcv_prepare: NULL argument");
    HG_REQUIRE(gcv_type == 0 || gcv_type == 1, "hg_dist_gcv_prepare: gcv_type must be 0 ('ab') or 1 ('ba')");
    HG_REQUIRE(k_gcv >= 1, "hg_dist_gcv_prepare: k_gcv must be >= 1");
    HG_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    std::vector<double> H;
    double beta = 0.0;
    if (gcv_type == 0) {
        HG_TRY(dist_gcv_ab(ctx, comm, A_p, B_p, b_p, k_gcv, H, &beta));
    } else {
        DHolder holder;
        HG_TRY(hg_darnoldi_create(ctx, comm, A_p, B_p, k_gcv, &holder.a));
        hg_darnoldi* a = holder.a;
        HG_TRY(hg_darnoldi_set_rhs(a, b_p));
        HG_TRY(hg_darnoldi_reset(a, 0.0));
        for (int k = 1; k <= k_gcv; ++k) {
            HG_TRY(hg_darnoldi_steps(a, 1));
            HG_CUDA(cudaStreamSynchronize(ctx->stream));
            if (a->h_H[(size_t)(k - 1) * a->ldh() + k] < 1e-12) break;
        }
        H.assign(a->h_H, a->h_H + (size_t)(k_gcv + 1) * k_gcv);
        beta = a->h_beta[0];
    }
    const double trace_m = gcv_type == 0 ? (double)m : (double)A_p->cols;  // gcv_function.m:46-50
    return hg_gcv_from_H(H.data(), k_gcv + 1, k_gcv, beta, trace_m, out);
}

// ---- thin collective wrappers for the other translation units (gkb.cu) -------------------
int hg_comm_rank(const hg_comm* c) { return c->rank; }
int hg_comm_size(const hg_comm* c) { return c->nranks; }
int hg_comm_allreduce(hg_comm* c, double* buf, size_t count, cudaStream_t st) {
    HG_NCCL(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, c->comm, st));
    return HG_OK;
}
int hg_comm_reduce_scatter(hg_comm* c, const double* send, double* recv, size_t recvcount, cudaStream_t st) {
    HG_NCCL(g_nccl.ReduceScatter(send, recv, recvcount, ncclDouble, ncclSum, c->comm, st));
    return HG_OK;
}
int hg_comm_allgather(hg_comm* c, const double* send, double* recv, size_t sendcount, cudaStream_t st) {
    HG_NCCL(g_nccl.AllGather(send, recv, sendcount, ncclDouble, c->comm, st));
    return HG_OK;
}
