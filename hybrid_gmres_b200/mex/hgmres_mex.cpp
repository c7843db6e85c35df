// MEX gateway for libhgmres — the drop-in boundary of SURVEY.md §8b.
//
// One source, seven names: build once and install the binary as
//   hybrid_ab_gmres_rtp.mexa64  hybrid_ba_gmres_rtp.mexa64  gcv_function.mexa64
//   hybrid_lsqr_solver.mexa64   hybrid_lsmr_solver.mexa64   lsqr_solver.mexa64
//   lsmr_solver.mexa64
// next to the reference's .m files (a MEX file shadows the same-named .m).  The gateway
// dispatches on mexFunctionName(), unpacks the mxArrays (read only) and calls the C ABI of
// include/hgmres.h; all arithmetic happens in libhgmres.so on the B200.
//
//   mex -I<repo>/include hgmres_mex.cpp -L<repo>/hybrid_gmres_b200 -lhgmres -output hybrid_ba_gmres_rtp
//
// Ownership and errors:
//   * prhs are never written.  Outputs are mxCreate* arrays owned by MATLAB.
//   * Device matrices live in a small cache that survives across calls (mexLock + mexAtExit), keyed
//     on what identifies the caller's array: data pointers, dimensions, nnz and a 64-bit content hash
//     (complete up to 1 MB per array, strided sample + head + tail above).  The ~30 gcv_function calls
//     of one fminbnd, a lambda sweep, or four solvers called on the same A upload it once; a new array
//     that MATLAB happens to place at a recycled address hashes differently and is a miss.
//   * mexErrMsgIdAndTxt longjmps out of the gateway: C++ destructors do not run.  So nothing with a
//     destructor owns a resource here — temporaries are mxCalloc'ed (MATLAB frees those itself on an
//     error exit), device matrices belong to the cache, and every failing C-ABI call goes through
//     fail(), which first drops what the current call pinned and only then raises.
//
// UNVERIFIED under MATLAB / Octave (neither exists in the build container); compiled against
// mex/stub/mex.h and executed against the mock MATLAB API of tests/mex_mock (tests/test_gpu_mex.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "hgmres.h"
#include "mex.h"

namespace {

hg_ctx* g_ctx = nullptr;

// ---- content hash ---------------------------------------------------------------------------
uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) {
        h ^= c[i];
        h *= 1099511628211ull;
    }
    return h;
}
uint64_t hash_array(const void* p, size_t bytes) {
    uint64_t h = 1469598103934665603ull ^ (uint64_t)bytes;
    if (!p || bytes == 0) return h;
    const unsigned char* c = static_cast<const unsigned char*>(p);
    constexpr size_t kFull = 1u << 20, kChunks = 4096, kChunk = 256;
    if (bytes <= kFull) return fnv1a(c, bytes, h);
    h = fnv1a(c, 65536, h);
    h = fnv1a(c + bytes - 65536, 65536, h);
    const size_t step = (bytes - kChunk) / kChunks;
    for (size_t i = 0; i < kChunks; ++i) h = fnv1a(c + i * step, kChunk, h);
    return h;
}

// ---- resident matrices ----------------------------------------------------------------------
struct Entry {
    hg_matrix* m = nullptr;   // the matrix as passed
    hg_matrix* mt = nullptr;  // its transpose (Golub-Kahan solvers), built on first use
    const void *pr = nullptr, *ir = nullptr, *jc = nullptr;
    size_t rows = 0, cols = 0, nnz = 0;
    bool sparse = false;
    uint64_t hash = 0;
    uint64_t stamp = 0;
    int pinned = 0;  // in use by the running call: not evictable
};
constexpr int kCacheEntries = 6;
Entry g_cache[kCacheEntries];
uint64_t g_stamp = 0;
int g_hits = 0, g_misses = 0, g_gcv_hits = 0, g_gcv_misses = 0;  // observable through hgmres_mex_cache_stats

struct Gcv {  // memoised lambda-independent part of gcv_function (gcv_function.m:4-32)
    hg_gcv* g = nullptr;
    uint64_t ha = 0, hb = 0, hrhs = 0;
    size_t nb = 0;
    int k = 0, type = -1;
    double m = 0;
} g_gcv;

void release(Entry& e) {
    if (e.m) hg_matrix_destroy(e.m);
    if (e.mt) hg_matrix_destroy(e.mt);
    e = Entry();
}
void unpin_all() {
    for (Entry& e : g_cache) e.pinned = 0;
}
void at_exit() {
    if (g_gcv.g) hg_gcv_destroy(g_gcv.g);
    g_gcv = Gcv();
    for (Entry& e : g_cache) release(e);
    if (g_ctx) hg_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

// release what this call holds, THEN raise (mexErrMsgIdAndTxt does not return)
void fail(const char* what) {
    unpin_all();
    mexErrMsgIdAndTxt("hgmres:error", "%s: %s", what, hg_last_error());
}
void usage(const char* id, const char* msg) {
    unpin_all();
    mexErrMsgIdAndTxt(id, "%s", msg);
}

hg_ctx* ctx() {
    if (!g_ctx) {
        if (hg_ctx_create(0, nullptr, &g_ctx) != HG_OK) fail("hg_ctx_create");
        mexAtExit(at_exit);
        mexLock();
    }
    return g_ctx;
}

// Device-resident version of a MATLAB matrix (sparse CSC or full); pinned until the call returns.
Entry* resident(const mxArray* a, const char* name) {
    if (!mxIsDouble(a) || mxIsComplex(a)) usage("hgmres:type", "matrix arguments must be real double");
    const size_t rows = mxGetM(a), cols = mxGetN(a);
    const bool sparse = mxIsSparse(a);
    const void* pr = mxGetPr(a);
    const void* ir = sparse ? (const void*)mxGetIr(a) : nullptr;
    const mwIndex* jc = sparse ? mxGetJc(a) : nullptr;
    const size_t nnz = sparse ? (size_t)jc[cols] : rows * cols;
    uint64_t h = hash_array(pr, nnz * sizeof(double));
    if (sparse) {
        h ^= hash_array(ir, nnz * sizeof(mwIndex)) * 3u;
        h ^= hash_array(jc, (cols + 1) * sizeof(mwIndex)) * 5u;
    }
    for (Entry& e : g_cache)
        if (e.m && e.pr == pr && e.ir == ir && e.jc == (const void*)jc && e.rows == rows && e.cols == cols &&
            e.nnz == nnz && e.sparse == sparse && e.hash == h) {
            e.stamp = ++g_stamp;
            e.pinned = 1;
            ++g_hits;
            return &e;
        }
    ++g_misses;
    Entry* slot = nullptr;  // an empty slot, else the least recently used one that this call does not hold
    for (Entry& e : g_cache) {
        if (e.pinned) continue;
        if (!e.m) {
            slot = &e;
            break;
        }
        if (!slot || e.stamp < slot->stamp) slot = &e;
    }
    if (!slot) usage("hgmres:cache", "matrix cache exhausted within one call");
    release(*slot);
    int st;
    if (sparse)
        st = hg_matrix_from_csc(ctx(), (int64_t)rows, (int64_t)cols, (int64_t)nnz, jc, ir, 8 * (int)sizeof(mwIndex),
                                (const double*)pr, &slot->m);
    else
        st = hg_matrix_from_dense(ctx(), (int64_t)rows, (int64_t)cols, (const double*)pr, (int64_t)rows, &slot->m);
    if (st != HG_OK) {
        slot->m = nullptr;
        fail(name);
    }
    slot->pr = pr;
    slot->ir = ir;
    slot->jc = jc;
    slot->rows = rows;
    slot->cols = cols;
    slot->nnz = nnz;
    slot->sparse = sparse;
    slot->hash = h;
    slot->stamp = ++g_stamp;
    slot->pinned = 1;
    return slot;
}

const double* vec(const mxArray* a, size_t n, const char* name) {
    if (!mxIsDouble(a) || mxIsSparse(a) || mxGetNumberOfElements(a) != n) {
        unpin_all();
        mexErrMsgIdAndTxt("hgmres:size", "%s must be a full double vector of length %d", name, (int)n);
    }
    return mxGetPr(a);
}

mxArray* column(const double* src, int n) {
    mxArray* out = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    if (n > 0) memcpy(mxGetPr(out), src, (size_t)n * sizeof(double));
    return out;
}

double* scratch(int n) { return static_cast<double*>(mxCalloc((size_t)(n > 0 ? n : 1), sizeof(double))); }

// [x,error_norm,residual_norm,niters] = hybrid_{ab,ba}_gmres_rtp(A,B,b,x_true,tol,maxit,lambda)
void rtp(bool ab, int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 7) usage("hgmres:nargin", "expected (A,B,b,x_true,tol,maxit,lambda)");
    Entry* A = resident(prhs[0], "A");
    Entry* B = resident(prhs[1], "B");
    const size_t m = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    const double* b = vec(prhs[2], m, "b");
    const double* xt = vec(prhs[3], n, "x_true");
    const double tol = mxGetScalar(prhs[4]), lambda = mxGetScalar(prhs[6]);
    const int maxit = (int)mxGetScalar(prhs[5]);
    double* err = scratch(maxit);
    double* res = scratch(maxit);
    mxArray* x = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    int niters = 0, x_valid = 0;
    const int st = (ab ? hg_hybrid_ab_gmres_rtp : hg_hybrid_ba_gmres_rtp)(
        ctx(), A->m, B->m, b, xt, tol, maxit, lambda, mxGetPr(x), err, res, &niters, &x_valid, nullptr, nullptr);
    if (st != HG_OK) fail(ab ? "hybrid_ab_gmres_rtp" : "hybrid_ba_gmres_rtp");
    unpin_all();
    if (!x_valid)  // hybrid_ab_gmres_rtp.m:25 at k = 1: MATLAB raises the same error for the .m file
        mexErrMsgIdAndTxt("hgmres:unassigned", "Output argument \"x\" (and maybe others) not assigned during call.");
    plhs[0] = x;
    if (nlhs > 1) plhs[1] = column(err, niters);
    if (nlhs > 2) plhs[2] = column(res, niters);
    if (nlhs > 3) plhs[3] = mxCreateDoubleScalar((double)niters);
}

// gcv_val = gcv_function(lambda,A,B,b,m,k_gcv,gcv_type)
void gcv(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    if (nrhs != 7) usage("hgmres:nargin", "expected (lambda,A,B,b,m,k_gcv,gcv_type)");
    const double lambda = mxGetScalar(prhs[0]);
    const double m = mxGetScalar(prhs[4]);
    const int k = (int)mxGetScalar(prhs[5]);
    char type[8] = {0};
    mxGetString(prhs[6], type, sizeof(type));
    const int t = strcmp(type, "ab") == 0 ? 0 : 1;  // gcv_function.m:4,7: anything else is 'ba'
    Entry* A = resident(prhs[1], "A");
    Entry* B = resident(prhs[2], "B");
    const size_t nb = mxGetM(prhs[1]);
    const double* b = vec(prhs[3], nb, "b");
    const uint64_t hrhs = hash_array(b, nb * sizeof(double));
    // the memo is keyed on CONTENT (the matrix hashes of the cache entries and a hash of b), so the
    // B_pert / b_noise rebuilt at identical sizes — possibly at recycled addresses — by the loops of
    // plot_error_vs_mismatch_norm.m:30-49 and plot_error_vs_noise_level.m:28-43 never hit a stale entry
    if (!(g_gcv.g && g_gcv.ha == A->hash && g_gcv.hb == B->hash && g_gcv.hrhs == hrhs && g_gcv.nb == nb &&
          g_gcv.k == k && g_gcv.type == t && g_gcv.m == m)) {
        ++g_gcv_misses;
        if (g_gcv.g) hg_gcv_destroy(g_gcv.g);
        g_gcv = Gcv();
        if (hg_gcv_prepare(ctx(), A->m, B->m, b, (int64_t)m, k, t, &g_gcv.g) != HG_OK) {
            g_gcv.g = nullptr;
            fail("gcv_function");
        }
        g_gcv.ha = A->hash;
        g_gcv.hb = B->hash;
        g_gcv.hrhs = hrhs;
        g_gcv.nb = nb;
        g_gcv.k = k;
        g_gcv.type = t;
        g_gcv.m = m;
    }
    else ++g_gcv_hits;
    double val = 0.0;
    if (hg_gcv_eval(g_gcv.g, lambda, &val) != HG_OK) fail("gcv_function");
    unpin_all();
    plhs[0] = mxCreateDoubleScalar(val);
}

// [x,error_norm,residual_norm,niters] = {hybrid_lsqr,hybrid_lsmr}_solver(A,b,x_true,tol,maxit,lambda)
//                                     = lsqr_solver(A,b,x_true,tol,maxit)
// [x,err,res,ar,iters]                = lsmr_solver(A,b,x_true,tol,maxit)   (tol, maxit optional)
void gkb(const std::string& name, int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const bool hybrid = name.rfind("hybrid_", 0) == 0;
    const bool lsmr5 = name == "lsmr_solver";
    if (nrhs < (lsmr5 ? 2 : (hybrid ? 6 : 5))) usage("hgmres:nargin", "too few input arguments");
    Entry* A = resident(prhs[0], "A");
    if (!A->mt && hg_matrix_transpose(ctx(), A->m, &A->mt) != HG_OK) {  // A' stays resident with A
        A->mt = nullptr;
        fail("transpose");
    }
    const size_t m = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    const double* b = vec(prhs[1], m, "b");
    const double* xt = nullptr;
    if (nrhs >= 3 && !mxIsEmpty(prhs[2])) xt = vec(prhs[2], n, "x_true");
    double tol = 1e-6;                             // lsmr_solver.m:3
    int maxit = (int)(m < n ? m : n);              // lsmr_solver.m:5
    if (nrhs >= 4 && !mxIsEmpty(prhs[3])) tol = mxGetScalar(prhs[3]);
    if (nrhs >= 5 && !mxIsEmpty(prhs[4])) maxit = (int)mxGetScalar(prhs[4]);
    const double lambda = hybrid ? mxGetScalar(prhs[5]) : 0.0;
    if (!lsmr5 && !xt) usage("hgmres:nargin", "x_true is required");
    double* err = scratch(maxit);
    double* res = scratch(maxit);
    double* ar = scratch(maxit);
    mxArray* x = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    int it = 0, st;
    if (name == "hybrid_lsqr_solver")
        st = hg_hybrid_lsqr_solver(ctx(), A->m, A->mt, b, xt, tol, maxit, lambda, mxGetPr(x), err, res, &it, nullptr);
    else if (name == "hybrid_lsmr_solver")
        st = hg_hybrid_lsmr_solver(ctx(), A->m, A->mt, b, xt, tol, maxit, lambda, mxGetPr(x), err, res, &it, nullptr);
    else if (name == "lsqr_solver")
        st = hg_lsqr_solver(ctx(), A->m, A->mt, b, xt, tol, maxit, mxGetPr(x), err, res, &it, nullptr);
    else
        st = hg_lsmr_solver(ctx(), A->m, A->mt, b, xt, tol, maxit, mxGetPr(x), err, res, ar, &it, nullptr);
    if (st != HG_OK) fail(name.c_str());
    unpin_all();
    plhs[0] = x;
    if (nlhs > 1) plhs[1] = column(err, it);
    if (nlhs > 2) plhs[2] = column(res, it);
    if (lsmr5) {
        if (nlhs > 3) plhs[3] = column(ar, it);
        if (nlhs > 4) plhs[4] = mxCreateDoubleScalar((double)it);
    } else if (nlhs > 3) {
        plhs[3] = mxCreateDoubleScalar((double)it);
    }
}

}  // namespace

// matrix-cache and gcv-memo counters since load: {matrix hits, matrix misses, gcv hits, gcv misses}
extern "C" void hgmres_mex_cache_stats(int out[4]) {
    out[0] = g_hits;
    out[1] = g_misses;
    out[2] = g_gcv_hits;
    out[3] = g_gcv_misses;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fname = mexFunctionName();
    const std::string name = fname ? fname : "";
    if (name == "hybrid_ab_gmres_rtp") rtp(true, nlhs, plhs, nrhs, prhs);
    else if (name == "hybrid_ba_gmres_rtp") rtp(false, nlhs, plhs, nrhs, prhs);
    else if (name == "gcv_function") gcv(nlhs, plhs, nrhs, prhs);
    else if (name == "hybrid_lsqr_solver" || name == "hybrid_lsmr_solver" || name == "lsqr_solver" ||
             name == "lsmr_solver") gkb(name, nlhs, plhs, nrhs, prhs);
    else mexErrMsgIdAndTxt("hgmres:name", "hgmres_mex installed under an unknown name '%s'", fname ? fname : "");
}
