// Internal declarations shared by the libhgmres translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "hgmres.h"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
void hg_set_error(const char* fmt, ...);

#define HG_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            hg_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,       \
                         __LINE__, cudaGetErrorString(_e));                                  \
            return HG_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define HG_TRY(call)                   \
    do {                               \
        int _s = (call);               \
        if (_s != HG_OK) return _s;    \
    } while (0)

#define HG_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            hg_set_error(__VA_ARGS__);   \
            return HG_ERR_INVALID;       \
        }                                \
    } while (0)

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct hg_timed_launch {
    int klass;
    cudaEvent_t e0, e1;
};

struct hg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    // timing
    bool timing = false;
    std::vector<hg_timed_launch> pending;
    std::vector<cudaEvent_t> event_pool;
    double t_ms[HG_K_NCLASSES] = {0};
    uint64_t t_count[HG_K_NCLASSES] = {0};
    double t_bytes[HG_K_NCLASSES] = {0};
    // scratch for deterministic two-stage reductions
    double* d_partials = nullptr;  // device
    size_t partials_cap = 0;       // in doubles
    double* d_scalars = nullptr;   // 64 device doubles
    double* h_scalars = nullptr;   // 64 pinned doubles
    // device-buffer cache (hg_dmalloc / hg_dfree): released blocks >= 1 MB wait here for the next
    // request of (nearly) the same size, so repeated solver calls neither cudaMalloc nor cudaFree
    std::multimap<size_t, void*> pool_free;
    size_t pool_cached = 0;
    std::multimap<size_t, void*> hpool_free;  // pinned host blocks (hg_hmalloc / hg_hfree)
};

// Stream-ordered reuse: a cached block is only handed to requests of the context that released it,
// and every consumer of a context runs on its one stream.
cudaError_t hg_dmalloc_bytes(hg_ctx* ctx, void** p, size_t bytes);
template <class T>
inline cudaError_t hg_dmalloc(hg_ctx* ctx, T** p, size_t bytes) {
    return hg_dmalloc_bytes(ctx, reinterpret_cast<void**>(p), bytes);
}
void hg_dfree(void* p);           // nullptr ok; blocks from hg_dmalloc go back to their context's cache
// pinned host memory with the same caching (a solver call in steady state makes no CUDA allocation call:
// cudaFree / cudaFreeHost were measured to stall for 100-400 ms now and then on a box with tens of GB mapped)
cudaError_t hg_hmalloc_bytes(hg_ctx* ctx, void** p, size_t bytes);
template <class T>
inline cudaError_t hg_hmalloc(hg_ctx* ctx, T** p, size_t bytes) {
    return hg_hmalloc_bytes(ctx, reinterpret_cast<void**>(p), bytes);
}
void hg_hfree(void* p);
// the context whose cache serves the RAII buffers (DBuf / PinBuf) of the solver running on this thread
struct hg_alloc_scope {
    hg_ctx* prev;
    explicit hg_alloc_scope(hg_ctx* c);
    ~hg_alloc_scope();
};
cudaError_t hg_dmalloc_cur(void** p, size_t bytes);  // current context's cache, else plain cudaMalloc
cudaError_t hg_hmalloc_cur(void** p, size_t bytes);  // current context's cache, else plain cudaMallocHost
void hg_pool_trim(hg_ctx* ctx);   // cudaFree everything cached

struct hg_matrix {
    hg_ctx* ctx = nullptr;
    int64_t rows = 0, cols = 0, nnz = 0;
    int64_t* rowptr = nullptr;  // device, rows+1
    int32_t* colind = nullptr;  // device, nnz
    double* vals = nullptr;     // device, nnz
    int tpr = 32;               // threads per row chosen at upload
    // streaming-SpMV work partition (lazily built cache, see spmv_stream.cu)
    int64_t* unit_row = nullptr;  // device, n_units+1 row boundaries, nnz balanced
    int n_units = 0;
    // 32-row sliced copy for the row-per-lane SpMV (lazily built cache, see spmv_sell.cu)
    int sell_state = 0;           // 0 not examined, 1 built, -1 not eligible
    int64_t sell_slices = 0, sell_entries = 0;
    int64_t* sell_ptr = nullptr;  // device, sell_slices+1 entry offsets (multiples of 32)
    int32_t* sell_col = nullptr;  // device, sell_entries, slice-column-major (freed when sell_col16 exists)
    double* sell_val = nullptr;
    // 16-bit column offsets (spmv_idx16.cu): col = base[group] + col16[entry]
    uint8_t* sell_col8 = nullptr;    // sell_entries; byte offsets [batch of 128][lane][4] from a base per slice column (sell_base: 4 per batch)
    uint16_t* sell_col16 = nullptr;  // sell_entries; group = 128 consecutive entries (4 columns of a slice)
    int32_t* sell_base = nullptr;    // sell_entries / 128
    int csr16_state = 0;             // 0 not examined, 1 built, -1 not eligible
    uint16_t* csr_col16 = nullptr;   // nnz, same indexing as colind; group = 32 consecutive entries of a row
    int32_t* csr_base = nullptr;     // csr_groups
    int64_t* csr_gptr = nullptr;     // rows+1: first group of each row
    int64_t csr_groups = 0;
    // row-group interleaved copy for the gather-bound row-per-warp matrices (lazily built, spmv_group.cu)
    int grp_state = 0;            // 0 not examined, 1 built, -1 not eligible
    int grp_G = 0;                // rows per group
    int64_t grp_groups = 0, grp_entries = 0;
    int64_t* grp_ptr = nullptr;   // grp_groups+1 entry offsets (multiples of 32)
    int32_t* grp_col = nullptr;   // grp_entries
    double* grp_val = nullptr;
    int32_t* grp_col0 = nullptr;  // 16-bit form: 4 column checkpoints of each lane (grp_groups * 4 * 32) ...
    int16_t* grp_d16 = nullptr;   // ... and per-lane column differences round to round (grp_entries); grp_col unused
};

// colind / vals are allocated with this many zero entries of tail padding so the
// 16-byte granular bulk copies of the streaming SpMV may over-read safely
constexpr int kNnzPad = 16;

#include <mutex>
// serialises the lazily built per-matrix SpMV forms (sliced copy, 16-bit companions, streaming work table):
// two host threads using one hg_matrix through different contexts build them once
std::mutex& hg_matrix_form_mutex();

int hg_ensure_partials(hg_ctx* ctx, size_t ndoubles);
// doubles a per-block `stat` partial buffer needs for a vector / SpMV of `rows` rows: the block kernels
// write at most rows/8 + 1024 partials, the streaming SpMV one per work unit (sm_count * 8)
inline size_t hg_stat_capacity(const hg_ctx* ctx, int64_t rows) {
    const size_t a = (size_t)(rows > 0 ? rows : 0) / 8 + 1024, b = (size_t)ctx->sm_count * 16 + 64;
    return a > b ? a : b;
}
int hg_matrix_alloc(hg_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz, hg_matrix** out);
void hg_matrix_pick_tpr(hg_matrix* m);

// RAII-less launch bracket: counts the launch and, when timing is on, records
// CUDA events on the launch stream around it.
struct hg_launch_scope {
    hg_ctx* ctx;
    int klass;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    hg_launch_scope(hg_ctx* c, int k, double bytes);
    ~hg_launch_scope();
};

// ---------------------------------------------------------------------------
// device kernels' host wrappers (kernels.cu).  All pointers are device pointers.
// ---------------------------------------------------------------------------

// y[r] = alpha * (M x)[r] + g1 * z1[r] + g2 * z2[r]; optional stat:
// partials <- per-block sum of (y[r] - ref[r])^2 (ref may be null -> y[r]^2);
// returns the number of partials written in *nparts (0 if stat == nullptr).
// If y == nullptr the result is not stored (stat only).
struct hg_spmv_epilogue {
    double alpha = 1.0;
    const double* z1 = nullptr;
    double g1 = 0.0;
    const double* z2 = nullptr;
    double g2 = 0.0;
    const double* ref = nullptr;
    double* stat = nullptr;
};
int hg_k_spmv(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
              const hg_spmv_epilogue& ep, int* nparts);

double hg_spmv_stream_bytes(const hg_matrix* m);  // (8 + index bytes) per entry + pointers, as stored

// partials[j*nslabs + slab] = sum over slab rows V[r,j]*w[r];  returns nslabs
int hg_k_multidot(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* w,
                  double* partials, int* nslabs);

// out[j] (+)= sum_i partials[j*np + i]  for j<k; if out2 != null also out2[j] = that sum
// (not accumulated); if do_sqrt, out[j] = sqrt(sum) (no accumulate).
int hg_k_reduce(hg_ctx* ctx, const double* partials, int np, int k, double* out, bool accumulate,
                double* out2, bool do_sqrt);

// out[r] = (z ? z[r] : 0) + s * sum_j V[r,j]*c[j];  c is a DEVICE array of k
// doubles.  Optional stat partials of (out[r]-ref[r])^2 (ref null -> out^2).
// out may be null (stat only).
int hg_k_lincomb(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* c,
                 double s, const double* z, double* out, const double* ref, double* stat,
                 int* nparts);

// iterate + both histories of a hybrid iteration in one launch (kernels.cu: iterate_kernel)
int hg_k_iterate(hg_ctx* ctx, const double* Q, int64_t ldq, int64_t n, const double* T, int64_t ldt, int64_t m,
                 int k, const double* y, const double* b, double* x, const double* x_true, double* stat,
                 unsigned int* ticket, double* out2);

int hg_k_iterate2(hg_ctx* ctx, const double* V0, int64_t ld0, int64_t n0, int k0, const double* c0, double* x,
                  const double* x_true, const double* V1, int64_t ld1, int64_t n1, int k1, const double* c1,
                  const double* b, double* stat, unsigned int* ticket, double* out2);

// same, with the result rows also stored to up to 16 further destinations (peer-GPU copies of the
// vector: the all-gather of the sharded Arnoldi is done by the producing kernel, dist_peer.cu)
struct hg_out_list {
    double* p[16];
    int n = 0;
};
int hg_k_lincomb_push(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* c,
                      double s, const double* z, double* out, const double* ref, double* stat,
                      int* nparts, const hg_out_list* extra);

// v[i] = v[i] / *d_div   (division, as `v / H(k+1,k)` in the reference)
int hg_k_scale_div(hg_ctx* ctx, double* v, int64_t n, const double* d_div);
// out = a*x + b*y (y may be null); optional stat of (out-ref)^2
int hg_k_axpby(hg_ctx* ctx, int64_t n, double a, const double* x, double b, const double* y,
               double* out, const double* ref, double* stat, int* nparts);
// LSQR: x += c1*w ; w = v - c2*w ; stat of (x - ref)^2   (hybrid_lsqr_solver.m:39-42)
// d = a*u_old + b*u_new - (first ? 0 : cprev*d);  [d2 = d - (first ? 0 : c0*d2);]  r -= step*(d2 ? d2 : d);
// stat: partials of r.^2 (LSQR / LSMR residual from the Golub-Kahan relation, gkb.cu)
int hg_k_gkb_resid(hg_ctx* ctx, int64_t n, const double* u_old, double a, const double* u_new, double b, double* d,
                   double cprev, double* d2, double c0, bool first, double* r, double step, double* stat,
                   int* nparts);
int hg_gkb_residual_mode();  // 0 (default): residuals of the GKB solvers from the Golub-Kahan relation; 1: b - A*x by SpMV
void hg_gkb_residual_mode_set(int v);
int hg_k_lsqr_update(hg_ctx* ctx, int64_t n, double* x, double* w, const double* v, double c1,
                     double c2, const double* ref, double* stat, int* nparts);
// LSMR: hbar = h - c0*hbar (or hbar = h when first); x += c1*hbar; h = v - c2*h
// (lsmr_solver.m:61-67)
int hg_k_lsmr_update(hg_ctx* ctx, int64_t n, double* x, double* h, double* hbar, const double* v,
                     int first, double c0, double c1, double c2, const double* ref, double* stat,
                     int* nparts);
// partials of sum x[i]^2 over a plain array (norms, Frobenius norm of vals)
int hg_k_sumsq(hg_ctx* ctx, const double* x, int64_t n, double* stat, int* nparts);

// convenience: full norm^2 of a device vector into a host double (synchronises)
int hg_norm2_sync(hg_ctx* ctx, const double* x, int64_t n, double* out);
// reduce `np` partials at ctx->d_partials into d_scalars[slot] (optionally sqrt)
int hg_reduce_to_scalar(hg_ctx* ctx, int np, int slot, bool do_sqrt);

// shared-memory-staged (cp.async) fused CGS2 middle stage (cgs_staged.cu): w1 = w0 - V h and partials = V^T w1 in one pass
// over V.  hg_cgs_staged_nparts: partials per column the kernel writes, 0 when (n, k) is out of range.
int hg_cgs_staged_nparts(const hg_ctx* ctx, int64_t n, int k);
int hg_k_cgs_mid_staged(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* h,
                     const double* w0, double* w1, double* partials, int* nparts);

// whole CGS2 step in one persistent cooperative kernel (cgs2_step.cu): small / medium Krylov vectors
bool hg_cgs2_step_eligible(const hg_ctx* ctx, int64_t n, int k);
bool hg_cgs2_step_eligible_dist(const hg_ctx* ctx, int64_t n_p, int k);
int64_t hg_cgs2_step_max_n();
void hg_cgs2_step_max_n_set(int v);
void hg_cgs2_step_max_n_dist_set(int v);
int hg_k_cgs2_step(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* w0, double* w1,
                   double* qnext, double* Hcol, double* hcur, double* partials);

// options (spmv_stream.cu)
int hg_cgs_fused_mode();  // 0 separate update / multi-dot kernels, 2 (default) shared-memory-staged one-pass kernel

// streaming SpMV (spmv_stream.cu)
bool hg_spmv_stream_eligible(const hg_matrix* m);
int hg_k_spmv_stream(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                     const hg_spmv_epilogue& ep, int* nparts);

// row-per-lane SpMV over 32-row slices (spmv_sell.cu)
int hg_spmv_mode();
bool hg_sell_ready(hg_ctx* ctx, const hg_matrix* m);
int hg_k_spmv_sell(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                   const hg_spmv_epilogue& ep, double bytes, int* nparts);

// 16-bit column offsets (spmv_idx16.cu)
bool hg_idx16_enabled();      // sliced form streams 16-bit column offsets (default)
bool hg_idx16_csr_enabled();  // row-per-warp CSR kernel too (opt-in: option spmv_idx16 = 2)
bool hg_idx8_wanted(const hg_matrix* m);
void hg_idx8_set(int v);
void hg_sell_compress(hg_ctx* ctx, hg_matrix* m);          // after the sliced copy is built
bool hg_csr16_ready(hg_ctx* ctx, const hg_matrix* m);      // lazily builds the CSR companion arrays
int hg_k_spmv_sell16(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                     const hg_spmv_epilogue& ep, double bytes, int* nparts);
int hg_k_spmv_csr16(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                    const hg_spmv_epilogue& ep, double bytes, int* nparts);
void hg_idx16_free(hg_matrix* m);

// row-group interleaved form (spmv_group.cu): option "spmv_group" / env HG_SPMV_GROUP
#ifndef HG_SPMV_GROUP_DEFAULT
#define HG_SPMV_GROUP_DEFAULT 4
#endif
int hg_spmv_group();
void hg_spmv_group_set(int v);
bool hg_group_ready(hg_ctx* ctx, const hg_matrix* m);
void hg_group_free(hg_matrix* m);
int hg_spmv_group16();
void hg_spmv_group16_set(int v);
int hg_spmv_group_split();
void hg_spmv_group_split_set(int v);
int64_t hg_spmv_group_min_rows();
void hg_spmv_group_min_rows_set(int v);
int hg_k_spmv_group(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y, const hg_spmv_epilogue& ep,
                    double bytes, int* nparts);

// transposition (matrix.cu)
int hg_transpose_device(hg_ctx* ctx, const hg_matrix* m, hg_matrix** out);

// collectives of the multi-GPU path (dist.cu); comm == nullptr means single GPU
int hg_comm_rank(const hg_comm* c);
int hg_comm_size(const hg_comm* c);
int hg_comm_allreduce(hg_comm* c, double* buf, size_t count, cudaStream_t st);
int hg_comm_reduce_scatter(hg_comm* c, const double* send, double* recv, size_t recvcount, cudaStream_t st);
int hg_comm_allgather(hg_comm* c, const double* send, double* recv, size_t sendcount, cudaStream_t st);
