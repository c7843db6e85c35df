/*
 * hgmres.h — C ABI of libhgmres.so, the B200-native (sm_100a) implementation of
 * the Arnoldi / Golub-Kahan hot path of luisayang-malaxiangguo/Hybrid-GMRES.
 *
 * The reference has no FFI layer: its drop-in boundary is the MATLAB function
 * signature.  Each "whole solver" entry point below therefore
 * carries exactly the arguments and outputs of one reference function, as raw
 * pointers + sizes; a MEX gateway or the ctypes binding in
 * hybrid_gmres_b200/_lib.py unpacks its arrays and calls it.
 *
 *   reference function (file:line)                         entry point
 *   hybrid_ab_gmres_rtp.m:1                                hg_hybrid_ab_gmres_rtp
 *   hybrid_ba_gmres_rtp.m:1                                hg_hybrid_ba_gmres_rtp
 *   gcv_function.m:1   (Arnoldi :4-32 / projected :33-58)  hg_gcv_prepare / hg_gcv_eval
 *   MATLAB fminbnd over gcv_function
 *     (analyze_regularization.m:37-46)                     hg_gcv_fminbnd
 *   hybrid_lsqr_solver.m:1                                 hg_hybrid_lsqr_solver
 *   hybrid_lsmr_solver.m:1                                 hg_hybrid_lsmr_solver
 *   lsqr_solver.m:1                                        hg_lsqr_solver
 *   lsmr_solver.m:1                                        hg_lsmr_solver
 *
 * Conventions
 *   - every function returns an hg_status; hg_last_error() gives the message of
 *     the calling thread's last failure.  No C++ exception crosses this ABI.
 *   - all floating point data is IEEE double; dense arrays are column-major
 *     (MATLAB layout).  "host" pointers are ordinary (or pinned) CPU memory,
 *     never written unless documented as outputs.
 *   - matrices live on the device as CSR (int64 row pointers, int32 column
 *     indices, double values).  B is always a separately stored matrix — there
 *     is no atomics-based transposed product anywhere.
 *   - numerical breakdowns are not errors; they follow the reference's own
 *     semantics (e.g. hybrid_ab_gmres_rtp.m:25,41-43).
 *   - there is no CPU fallback: every entry point that computes fails with
 *     HG_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef HGMRES_H
#define HGMRES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hg_status {
    HG_OK = 0,
    HG_ERR_INVALID = 1, /* bad argument                                    */
    HG_ERR_CUDA = 2,    /* CUDA runtime failure / no usable device         */
    HG_ERR_NOMEM = 3,   /* host or device allocation failed                */
    HG_ERR_NCCL = 4,    /* NCCL failure (multi-GPU path)                   */
    HG_ERR_STATE = 5    /* call sequence error (e.g. rhs not set)          */
} hg_status;

typedef struct hg_ctx hg_ctx;         /* one device + one stream + scratch   */
typedef struct hg_matrix hg_matrix;   /* device-resident CSR matrix          */
typedef struct hg_arnoldi hg_arnoldi; /* device Krylov basis + Hessenberg    */
typedef struct hg_gcv hg_gcv;         /* memoised gcv_function Arnoldi       */

/* ---- library ------------------------------------------------------------ */
const char* hg_last_error(void);
int hg_version(void);

/* Tuning knobs.  "spmv_mode": 0 auto (default; also env HG_SPMV=v1|v2|v3): row-per-lane kernel over
 * 32-row slices when that traversal gathers fewer 128-byte lines of x per entry than the row-per-warp
 * one does (sampled at first use; padding <= 10 %), else the row-per-warp CSR kernel; 1 force the
 * row-per-warp CSR SpMV, 2 force the TMA-staged streaming SpMV, 3 force the sliced form.
 * "cgs_fused" / env HG_CGS_FUSED: the first CGS2 update and the second-pass dot products in one pass
 * over the basis (it crosses HBM three times per step instead of four).  2 (default): tiles staged in
 * shared memory by cp.async, persistent CTAs (csrc/cgs_staged.cu; 599 -> 634 it/s on the headline
 * workload); 0: separate update and multi-dot kernels.  (Two measured-slower experiments of round 1 — an
 * L2-re-read fused kernel and backwards-walking update sweeps — were removed in round 2; their numbers are
 * in profiles/r01_cgs_fusion.md.)
 * "cgs_step_max_n" / env HG_CGS_STEP_MAX_N (default 140000; 0 disables): Krylov vectors up to this length
 * run the whole CGS2 step (orthogonalisation, norm, normalisation) in ONE persistent cooperative kernel
 * (csrc/cgs2_step.cu) instead of seven launches.  "cgs_step_max_n_dist" / HG_CGS_STEP_MAX_N_DIST (default 0 =
 * off: measured slower than the separate kernels): the same for a rank's slice on several GPUs, where the
 * kernel also does the step's collectives.
 * "spmv_group" / env HG_SPMV_GROUP (default 4; 0 off, 2, 4, 8): long-row matrices that run the row-per-warp
 * kernel (ray-driven projectors) get a copy in which G adjacent rows are interleaved per warp, so a warp-wide
 * gather serves G adjacent rays from the same cache lines (csrc/spmv_group.cu: L1 wavefronts 66 % -> 48 %).
 * "spmv_group16" / HG_SPMV_GROUP16 (default 1): that copy stores its columns as 16-bit per-lane differences (10
 * instead of 12 bytes per non-zero; automatic 32-bit fall-back).  "spmv_group_split" / HG_SPMV_GROUP_SPLIT (0 = by
 * size; 1, 2, 4): warps sharing one group.  "spmv_group_min_rows" / HG_SPMV_GROUP_MIN_ROWS (default 16384): smallest
 * matrix that takes the form by default.
 * "spmv_idx16" / env HG_IDX16 (default 1; 0 off; 2 also the row-per-warp kernel): 16-bit column offsets from a base
 * per 128 entries in the sliced form.  "spmv_idx8" / env HG_IDX8 (-1 default: matrices of >= 131 072 rows; 1 always;
 * 0 never): byte offsets from a base per slice column (9.1 bytes per non-zero) when every slice column spans < 256.
 * All index widths give bit-identical products.
 * "gkb_residual" / env HG_GKB_RESIDUAL (default 0): the hybrid LSQR / hybrid LSMR / LSMR solvers take `b - A*x` of
 * their residual histories from the Golub-Kahan relation (vector recurrences, no extra product with A); 1 forms it by
 * SpMV as the reference does.  The iterates do not depend on it.
 * "dist_transport": see hg_comm_transport. */
int hg_set_option(const char* name, int value);

/* ---- context ------------------------------------------------------------ */
/* `stream` is a cudaStream_t to launch on (e.g. torch's current stream), or
 * NULL to let the library create its own non-blocking stream. */
int hg_ctx_create(int device, void* stream, hg_ctx** out);
int hg_ctx_destroy(hg_ctx* ctx);
int hg_ctx_sync(hg_ctx* ctx);
/* Matrices and Krylov workspaces released by this context are cached (up to HG_POOL_GB, default 64)
 * for the next request of the same size, so repeated solver calls do not cudaMalloc / cudaFree;
 * hg_ctx_trim returns the cache to the driver. */
int hg_ctx_trim(hg_ctx* ctx);
/* number of hand-written kernels launched by this context so far */
int hg_ctx_launch_count(hg_ctx* ctx, uint64_t* out);

/* Per-kernel-class device timing (CUDA events recorded on the launch stream
 * around every launch of that class while enabled). */
typedef enum hg_kernel_class {
    HG_K_SPMV = 0,     /* CSR SpMV (all epilogues)                           */
    HG_K_MULTIDOT = 1, /* h = V^T w                                          */
    HG_K_LINCOMB = 2,  /* out = z + s * V c  (CGS update, iterate, residual) */
    HG_K_VECTOR = 3,   /* scale / axpby / fused LSQR-LSMR vector updates     */
    HG_K_REDUCE = 4,   /* second-stage deterministic reductions              */
    HG_K_SETUP = 5,    /* transpose / conversion / generators                */
    HG_K_COMM = 6,     /* flag barriers of the NVLink peer-memory transport  */
    HG_K_NCLASSES = 7
} hg_kernel_class;
int hg_ctx_timing_enable(hg_ctx* ctx, int on);
/* Synchronises, folds pending events, returns totals since the last reset:
 * summed device milliseconds, launches, and algorithmic bytes. */
int hg_ctx_timing_get(hg_ctx* ctx, int kernel_class, double* ms, uint64_t* launches, double* bytes);
int hg_ctx_timing_reset(hg_ctx* ctx);

/* Page-lock / unlock a caller-owned host range so uploads run at PCIe rate. */
int hg_host_register(void* ptr, size_t bytes);
int hg_host_unregister(void* ptr);

/* ---- matrices (SURVEY.md §8b "Input types") ------------------------------ */
/* CSR upload.  ptr_bits is 32 or 64 (width of rowptr entries, signed). */
int hg_matrix_from_csr(hg_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz, const void* rowptr,
                       int ptr_bits, const int32_t* colind, const double* vals, hg_matrix** out);
/* MATLAB sparse (CSC: Jc cols+1 entries, Ir nnz entries, Pr values).  idx_bits
 * is 64 for mwIndex arrays, 32 for scipy.  The CSC arrays are uploaded as the
 * CSR of the transpose and transposed on the device (deterministic). */
int hg_matrix_from_csc(hg_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz, const void* jc,
                       const void* ir, int idx_bits, const double* pr, hg_matrix** out);
/* Full column-major matrix (every reference script with n=32 passes full A,B). */
int hg_matrix_from_dense(hg_ctx* ctx, int64_t rows, int64_t cols, const double* a, int64_t lda,
                         hg_matrix** out);
/* out = M^T as a new CSR matrix, column indices sorted within each row. */
int hg_matrix_transpose(hg_ctx* ctx, const hg_matrix* m, hg_matrix** out);
/* out = M(rowperm, colperm) in MATLAB's gather convention: out row i is M row rowperm[i], out
 * column j is M column colperm[j] (0-based host arrays; either may be NULL = identity).  Columns are
 * re-sorted within each row unless flags has HG_PERMUTE_KEEP_ENTRY_ORDER (entries then keep their
 * order within the row and only get new column labels — enough for products, and what a projector
 * wants: the entries of a ray stay in traversal order).  Used to run the n-space of a solve in a cache-friendly pixel order
 * (A(:,q), B(q,:), x_true(q); the iterate comes back as x(q) = x_q) — an orthogonal similarity of the
 * operator B*A + lambda*I of hybrid_ab_gmres_rtp.m:6, so H, beta and the histories are unchanged up
 * to summation order. */
#define HG_PERMUTE_KEEP_ENTRY_ORDER 1
int hg_matrix_permute(hg_ctx* ctx, const hg_matrix* m, const int32_t* rowperm, const int32_t* colperm,
                      int flags, hg_matrix** out);
/* Which SpMV kernel this matrix runs with: low 4 bits 0 CSR row-per-thread-group, 1 row-per-lane over
 * 32-row slices (built lazily when rows of a slice have near-equal length), 2 TMA-staged streaming, 3 the
 * row-group interleaved form (csrc/spmv_group.cu; option "spmv_group" / env HG_SPMV_GROUP = rows per group);
 * bit 4 (value 16) set when the column indices are streamed as 16-bit values — offsets from per-group bases
 * (csrc/spmv_idx16.cu: 10 instead of 12 bytes per entry; option "spmv_idx16" / env HG_IDX16: 0 off,
 * 1 (default) for the sliced form, 2 also for the row-per-warp CSR kernel) or per-lane differences in the
 * row-group form ("spmv_group16"); bit 5 (value 32) set when the sliced form streams byte offsets from a base per
 * slice column (9.1 bytes per entry; "spmv_idx8"). */
int hg_matrix_spmv_form(hg_ctx* ctx, const hg_matrix* m, int* form);
int hg_matrix_info(const hg_matrix* m, int64_t* rows, int64_t* cols, int64_t* nnz);
int hg_matrix_download_csr(hg_ctx* ctx, const hg_matrix* m, int64_t* rowptr, int32_t* colind,
                           double* vals);
int hg_matrix_destroy(hg_matrix* m);

/* ---- device CT generators (synthetic inputs) -------- */
/* geometry: 0 parallel, 1 fan.  Tables as produced by oracle/ct.py:ray_tables. */
int hg_ct_projector(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                    const double* cos_th, const double* sin_th, const double* ray_a,
                    const double* ray_b, hg_matrix** out);
int hg_ct_backprojector(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                        const double* cos_th, const double* sin_th, hg_matrix** out);

/* Shards for the multi-GPU path: rays [row_lo,row_hi) of the projector (rows of A), and the
 * sinogram columns [col_lo,col_hi) of the back-projector (column block of B, local indices). */
int hg_ct_projector_rows(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                         const double* cos_th, const double* sin_th, const double* ray_a,
                         const double* ray_b, int64_t row_lo, int64_t row_hi, hg_matrix** out);
int hg_ct_backprojector_cols(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                             const double* cos_th, const double* sin_th, int64_t col_lo, int64_t col_hi,
                             hg_matrix** out);

/* ---- building blocks on host vectors (tests, setup) ---------------------- */
/* y = M x */
int hg_spmv(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y);
/* h = V^T w and out = z + s*V*c on a column-major host V (n x k, ld) — exercise
 * the CGS2 kernels in isolation. */
int hg_multidot(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, const double* w,
                double* h);
int hg_lincomb(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, const double* c,
               double s, const double* z, double* out, double* out_norm2);
/* CGS2 middle stage: w1 = w0 - V h and d = V' w1 (h: k coefficients).  fused != 0: the one-pass kernel
 * that stages the basis tile in shared memory (csrc/cgs_staged.cu; 24 <= k <= 208 and n >= 256 rows per
 * SM, else HG_ERR_INVALID); 0: the separate update + multi-dot kernels. */
int hg_cgs_mid(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, const double* h,
               const double* w0, int fused, double* w1, double* d);

/* One whole CGS2 step (two-pass classical Gram-Schmidt, norm, normalisation) of w0 against the k columns of
 * V in the single persistent cooperative kernel of csrc/cgs2_step.cu: hcol[0..k) = V'w0 + V'(w0 - V V'w0),
 * hcol[k] = ||v||, q = v / ||v||.  1 <= k <= 208, else HG_ERR_INVALID.  (The solvers use this kernel for
 * vectors up to option "cgs_step_max_n"; longer ones run the separate streaming kernels.) */
int hg_cgs2_step(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, const double* w0, double* hcol,
                 double* q);

/* ---- Arnoldi on device-resident data (a1-a4, a9 in SURVEY.md §8a) -------- */
typedef enum hg_space {
    HG_SPACE_N = 0, /* operator B*(A*q) + shift*q, start B*b  (hybrid_*_rtp.m:6-13,
                       gcv_function.m:8,22)                                        */
    HG_SPACE_M = 1  /* operator A*(B*q) + shift*q, start b    (gcv_function.m:5,20) */
} hg_space;
int hg_arnoldi_create(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, int space, int kmax,
                      hg_arnoldi** out);
int hg_arnoldi_destroy(hg_arnoldi* a);
/* upload b (m doubles) and keep it device resident */
int hg_arnoldi_set_rhs(hg_arnoldi* a, const double* b);
/* r0, beta, Q(:,1); k <- 0.  Enqueued, no host sync. */
int hg_arnoldi_reset(hg_arnoldi* a, double shift);
/* enqueue `nsteps` CGS2 Arnoldi steps without any host synchronisation */
int hg_arnoldi_steps(hg_arnoldi* a, int nsteps);
/* wait; copy H ((kmax+1) x kmax, column-major, ldh >= kmax+1), beta, steps done */
int hg_arnoldi_get(hg_arnoldi* a, double* H, int ldh, double* beta, int* ksteps);
/* copy basis vector j (0-based) to the host: length rows(Q) */
int hg_arnoldi_get_q(hg_arnoldi* a, int j, double* q);
/* algorithmic bytes of Arnoldi step k (1-based), SURVEY.md §8d S(k) */
int hg_arnoldi_step_bytes(hg_arnoldi* a, int k, double* bytes);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ---------------- */
typedef struct hg_comm hg_comm;
typedef struct hg_darnoldi hg_darnoldi;
/* rank 0 creates a 128-byte NCCL unique id; the caller broadcasts it (torch.distributed) */
int hg_comm_unique_id(void* out128);
int hg_comm_init(hg_ctx* ctx, int nranks, int rank, const void* id128, hg_comm** out);
int hg_comm_destroy(hg_comm* c);
/* How the sharded Arnoldi step moves data between ranks.  1: NVLink peer memory — the
 * reduce-scatter of B^p u_p is pulled by the CTAs of the first CGS2 multi-dot, the coefficient
 * all-reduces are a P2P inbox exchange inside the second-stage reduction, the all-gather of
 * q_{k+1} is pushed by the normalisation kernel (csrc/dist_peer.cu; workspace mapped with cudaIpc).
 * 0: ncclReduceScatter / ncclAllReduce / ncclAllGather calls between the kernels.  The choice is
 * made collectively when a sharded Arnoldi is created: option "dist_transport" / env HG_DIST =
 * auto (0, default: peer memory when every rank can map every other rank), nccl (1), peer (2:
 * fail instead of falling back).  `why` (optional) receives a one-line description. */
int hg_comm_transport(hg_comm* c, int* transport, char* why, int why_len);
/* Sharded n-space Arnoldi.  A_p: this rank's detector-row block of A (m_p x n); B_p: the
 * matching column block of B (n x m_p, local column indices).  Krylov vectors are sharded in
 * equal row slices of n_p = roundup32(ceil(n/P)) entries (zero padded). */
int hg_darnoldi_create(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* B_p, int kmax,
                       hg_darnoldi** out);
int hg_darnoldi_destroy(hg_darnoldi* a);
int hg_darnoldi_set_rhs(hg_darnoldi* a, const double* b_p); /* this rank's m_p entries of b */
int hg_darnoldi_reset(hg_darnoldi* a, double shift);
int hg_darnoldi_steps(hg_darnoldi* a, int nsteps);
int hg_darnoldi_get(hg_darnoldi* a, double* H, int ldh, double* beta, int* ksteps);
int hg_darnoldi_get_q(hg_darnoldi* a, int j, double* q_slice, int64_t* row0, int64_t* nrows);
int hg_darnoldi_step_bytes(hg_darnoldi* a, int k, double* bytes);

/* ---- whole solvers: the reference signatures ----------------------------- */
typedef struct hg_solver_opts {
    int residual_mode; /* 0: r = b - W*y from the cached A*Q columns (default)
                          1: literal r = b - A*x SpMV (hybrid_ab_gmres_rtp.m:35) */
    int error_mode;    /* 2: ||x_k - x_true||^2 = ||y_k||^2 - 2 y_k'c + ||x_true||^2 with c = Q_k'x_true (one dot
                          product per new basis vector; exact for an orthonormal basis, which CGS2 delivers to
                          1e-15), x_k formed explicitly only when that error is below 1 % of ||x_true||
                          (cancellation) and once for the returned iterate;
                          1: x_k = Q_k y_k and the difference formed at every iteration (:33,36 literally);
                          0 (default): 2 for n >= 200000, where the product costs more than the extra dot, else 1 */
    int reserved[6];
} hg_solver_opts;

/* Optional extra outputs for parity tests (any pointer may be NULL). */
typedef struct hg_extras {
    double* H;      /* (maxit+1) x maxit column-major Hessenberg                */
    double* beta;   /* norm of r0                                               */
    double* X_hist; /* n x maxit column-major: iterate after every iteration    */
    double* aux;    /* solver specific: GKB alphas/betas (2*(maxit+1))          */
} hg_extras;

/* [x,error_norm,residual_norm,niters] = hybrid_ab_gmres_rtp(A,B,b,x_true,tol,maxit,lambda)
 * x: n, error_norm/residual_norm: maxit (first niters entries valid).
 * x_valid (may be NULL) is 0 only when the reference would leave x unassigned
 * (breakdown at k=1, hybrid_ab_gmres_rtp.m:25). */
int hg_hybrid_ab_gmres_rtp(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, const double* b,
                           const double* x_true, double tol, int maxit, double lambda, double* x,
                           double* error_norm, double* residual_norm, int* niters, int* x_valid,
                           const hg_solver_opts* opts, hg_extras* extras);
int hg_hybrid_ba_gmres_rtp(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, const double* b,
                           const double* x_true, double tol, int maxit, double lambda, double* x,
                           double* error_norm, double* residual_norm, int* niters, int* x_valid,
                           const hg_solver_opts* opts, hg_extras* extras);

/* Wall-clock breakdown (milliseconds) of the last hybrid_*_gmres_rtp call made on this thread:
 * [0] setup (workspace, rhs upload, norms, r0), [1] the iteration loop, of which [2] host projected
 * solves and [3] host waiting for the device, [4] download of x, [5] iterations.  bench.py reports it
 * as e2e.breakdown_ms. */
int hg_last_solve_stats(double* out, int n);

/* Project-then-regularise solvers — the solve path (first four outputs) of
 *   kind 0, hybrid 1: ABgmres_hybrid_bounds.m:11-41      kind 1, hybrid 1: BAgmres_hybrid_bounds.m:11-40
 *   kind 0, hybrid 0: ABgmres_nonhybrid_bounds.m:12-40   kind 1, hybrid 0: BAgmres_nonhybrid_bounds.m:12-39
 * Unshifted Arnoldi in m-space (AB, x = B*z) or n-space (BA); y_k from Tikhonov on H_k (hybrid)
 * or H_k \ beta*e1.  The filter-factor bound outputs (phi, dphi, DeltaM) are out of scope.
 * x_valid is 0 when `x = xk` would be undefined in the reference (breakdown at k = 1). */
int hg_gmres_ptr(hg_ctx* ctx, int kind, int hybrid, const hg_matrix* A, const hg_matrix* B, const double* b,
                 const double* x_true, double tol, int maxit, double lambda, double* x, double* error_norm,
                 double* residual_norm, int* niters, int* x_valid, hg_extras* extras);

/* The same hybrid PTR solve with the regularisation parameter chosen AT EVERY ITERATION from the growing
 * Hessenberg matrix: lambda_k = the first minimiser over `lambdas[0..nl)` of
 * GCV(lambda, H_k) exactly as compute_gcv_surface / calculate_gcv_from_H evaluate it
 * (plot_gcv_surface.m:58-122), then y_k = (H_k'H_k + lambda_k I) \ (H_k' beta e1) and x_k as in
 * AB/BAgmres_hybrid_bounds.m:34-38.  lambda_path (maxit entries) receives lambda_1..lambda_niters.
 * No device work beyond hg_gmres_ptr's: the choice is made on the host from H. */
int hg_gmres_ptr_gcv(hg_ctx* ctx, int kind, const hg_matrix* A, const hg_matrix* B, const double* b,
                     const double* x_true, double tol, int maxit, const double* lambdas, int nl, double* x,
                     double* error_norm, double* residual_norm, double* lambda_path, int* niters, int* x_valid,
                     hg_extras* extras);

/* gcv_function(lambda,A,B,b,m,k_gcv,gcv_type): the lambda-independent Arnoldi
 * (gcv_function.m:4-32) runs ONCE in hg_gcv_prepare; hg_gcv_eval is the
 * projected part (:33-58) on the host.  gcv_type: 0 'ab', 1 'ba'. */
int hg_gcv_prepare(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* B, const double* b,
                   int64_t m, int k_gcv, int gcv_type, hg_gcv** out);
int hg_gcv_eval(const hg_gcv* g, double lambda, double* gcv_val);
/* copies H ((k_gcv+1) x k_gcv col-major) and beta */
int hg_gcv_get(const hg_gcv* g, double* H, double* beta);
/* MATLAB fminbnd(@(l) gcv_function(l,...), lo, hi, optimset('TolX',tolx)) */
int hg_gcv_fminbnd(const hg_gcv* g, double lo, double hi, double tolx, double* lambda,
                   double* fval, int* funccount, double* trace, int trace_cap);
int hg_gcv_destroy(hg_gcv* g);
/* plot_gcv_surface.m:58-102 (compute_gcv_surface) from the factorisation held by the handle:
 * surface (nl x k_gcv, column-major) = GCV(lambda_i, k) with H(1:k+1,1:k); path[k-1] = the grid
 * minimiser lambda_k.  Columns from a `H(k+1,k) < 1e-12` breakdown on stay zero (:85). */
int hg_gcv_surface(const hg_gcv* g, const double* lambdas, int nl, double* surface, double* path);
/* Host-only constructor from an existing Arnoldi factorisation (no device work):
 * H is (k+1) x k column-major with leading dimension ldh, trace_m is m ('ab') or
 * n ('ba') (gcv_function.m:46-50).  This is plot_gcv_surface.m:58-122's "Arnoldi
 * once, GCV for many lambda" form. */
int hg_gcv_from_H(const double* H, int ldh, int k, double beta, double trace_m, hg_gcv** out);

/* ---- host-side projected problems (north star: these stay on the CPU) ---- */
/* y = argmin || beta*e1 - H(1:k+1,1:k) y ||  by Givens QR (hybrid_ba_gmres_rtp.m:28-29) */
int hg_host_hessenberg_ls(const double* H, int ldh, int k, double beta, double* y);
/* y = M \ rhs, square column-major M (not modified): Cholesky if symmetric with
 * positive pivots, else LU with partial pivoting (hybrid_ab_gmres_rtp.m:32,
 * gcv_function.m:38, hybrid_lsmr_solver.m:44).  Returns HG_OK also when singular. */
int hg_host_solve_square(int n, const double* M, int ld, const double* rhs, double* y);
/* singular values, descending (gcv_function.m:42) */
int hg_host_singular_values(int n, const double* M, int ld, double* s);

/* At may be NULL: the library then builds A^T on the device for the call. */
int hg_hybrid_lsqr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                          const double* x_true, double tol, int maxit, double lambda, double* x,
                          double* error_norm, double* residual_norm, int* niters,
                          hg_extras* extras);
int hg_hybrid_lsmr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                          const double* x_true, double tol, int maxit, double lambda, double* x,
                          double* error_norm, double* residual_norm, int* niters,
                          hg_extras* extras);
int hg_lsqr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                   const double* x_true, double tol, int maxit, double* x, double* error_norm,
                   double* residual_norm, int* niters, hg_extras* extras);
/* x_true may be NULL (err_hist is then NaN, lsmr_solver.m:28,72-74) */
int hg_lsmr_solver(hg_ctx* ctx, const hg_matrix* A, const hg_matrix* At, const double* b,
                   const double* x_true, double tol, int maxit, double* x, double* err_hist,
                   double* res_hist, double* ar_hist, int* iters, hg_extras* extras);


/* ---- multi-GPU whole solvers ---------------------------------------------- */
/* Sharded hybrid_ab_gmres_rtp (kind 0) / hybrid_ba_gmres_rtp (kind 1): same arguments and
 * outputs as the single-GPU entry points, with A, B given as this rank's shards and b as this
 * rank's m_p entries; x_true and x are full n-vectors, histories are identical on all ranks. */
int hg_dist_hybrid_rtp(int kind, hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* B_p,
                       const double* b_p, const double* x_true, double tol, int maxit, double lambda,
                       double* x, double* error_norm, double* residual_norm, int* niters, int* x_valid,
                       hg_extras* extras);
/* Sharded gcv_function Arnoldi: 'ab' runs in m-space on the detector-row blocks (all-reduce of
 * the n-vector B*q), 'ba' in n-space; the returned handle is evaluated with hg_gcv_eval /
 * hg_gcv_fminbnd on every rank (identical H on all ranks). b_p: this rank's m_p entries. */
int hg_dist_gcv_prepare(hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* B_p,
                        const double* b_p, int64_t m, int k_gcv, int gcv_type, hg_gcv** out);
/* Sharded Golub-Kahan solvers.  which: 0 hybrid_lsqr_solver, 1 hybrid_lsmr_solver, 2 lsqr_solver,
 * 3 lsmr_solver.  A_p: this rank's detector-row block of A (m_p x n); At_p: its transpose or NULL
 * (built on the device); b_p: its m_p entries of b; x_true, x: full n-vectors (x_true may be NULL
 * for lsmr_solver); lambda is ignored by 2 and 3; ar_hist is required by 3 only. */
int hg_dist_gkb_solver(int which, hg_ctx* ctx, hg_comm* comm, const hg_matrix* A_p, const hg_matrix* At_p,
                       const double* b_p, const double* x_true, double tol, int maxit, double lambda,
                       double* x, double* error_norm, double* residual_norm, double* ar_hist, int* niters,
                       hg_extras* extras);

#ifdef __cplusplus
}
#endif
#endif /* HGMRES_H */
