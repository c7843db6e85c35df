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
// UNVERIFIED: neither MATLAB nor Octave (mex.h, mkoctfile) exists in the build container;
// this file is syntax-checked against mex/stub/mex.h only (tests/test_host_logic.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "hgmres.h"
#include "mex.h"

namespace {

hg_ctx* g_ctx = nullptr;

// gcv_function is called ~30 times by one fminbnd with the same (A,B,b,m,k_gcv,type): the
// lambda-independent device Arnoldi is memoised on the data pointers and sizes.
struct GcvCache {
    const void *a = nullptr, *b = nullptr, *rhs = nullptr;
    size_t nnz_a = 0, nnz_b = 0;
    int k = 0, type = -1;
    double m = 0;
    hg_gcv* g = nullptr;
} g_gcv;

void at_exit() {
    if (g_gcv.g) hg_gcv_destroy(g_gcv.g);
    g_gcv.g = nullptr;
    if (g_ctx) hg_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

void fail(const char* what) { mexErrMsgIdAndTxt("hgmres:error", "%s: %s", what, hg_last_error()); }

hg_ctx* ctx() {
    if (!g_ctx) {
        if (hg_ctx_create(0, nullptr, &g_ctx) != HG_OK) fail("hg_ctx_create");
        mexAtExit(at_exit);
        mexLock();
    }
    return g_ctx;
}

struct Mat {  // uploads a MATLAB matrix (sparse CSC or full) for the duration of one call
    hg_matrix* m = nullptr;
    ~Mat() { hg_matrix_destroy(m); }
    void upload(const mxArray* a, const char* name) {
        if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("hgmres:type", "%s must be real double", name);
        const int64_t rows = (int64_t)mxGetM(a), cols = (int64_t)mxGetN(a);
        int st;
        if (mxIsSparse(a)) {
            const mwIndex* jc = mxGetJc(a);
            st = hg_matrix_from_csc(ctx(), rows, cols, (int64_t)jc[cols], jc, mxGetIr(a), 8 * (int)sizeof(mwIndex),
                                    mxGetPr(a), &m);
        } else {
            st = hg_matrix_from_dense(ctx(), rows, cols, mxGetPr(a), rows, &m);
        }
        if (st != HG_OK) fail(name);
    }
};

const double* vec(const mxArray* a, size_t n, const char* name) {
    if (!mxIsDouble(a) || mxIsSparse(a) || mxGetNumberOfElements(a) != n)
        mexErrMsgIdAndTxt("hgmres:size", "%s must be a full double vector of length %d", name, (int)n);
    return mxGetPr(a);
}

mxArray* column(const double* src, int n) {
    mxArray* out = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    if (n > 0) memcpy(mxGetPr(out), src, (size_t)n * sizeof(double));
    return out;
}

// [x,error_norm,residual_norm,niters] = hybrid_{ab,ba}_gmres_rtp(A,B,b,x_true,tol,maxit,lambda)
void rtp(bool ab, int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 7) mexErrMsgIdAndTxt("hgmres:nargin", "expected (A,B,b,x_true,tol,maxit,lambda)");
    Mat A, B;
    A.upload(prhs[0], "A");
    B.upload(prhs[1], "B");
    const size_t m = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    const double* b = vec(prhs[2], m, "b");
    const double* xt = vec(prhs[3], n, "x_true");
    const double tol = mxGetScalar(prhs[4]), lambda = mxGetScalar(prhs[6]);
    const int maxit = (int)mxGetScalar(prhs[5]);
    std::vector<double> err((size_t)maxit), res((size_t)maxit);
    mxArray* x = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    int niters = 0, x_valid = 0;
    const int st = (ab ? hg_hybrid_ab_gmres_rtp : hg_hybrid_ba_gmres_rtp)(
        ctx(), A.m, B.m, b, xt, tol, maxit, lambda, mxGetPr(x), err.data(), res.data(), &niters, &x_valid, nullptr,
        nullptr);
    if (st != HG_OK) fail(ab ? "hybrid_ab_gmres_rtp" : "hybrid_ba_gmres_rtp");
    if (!x_valid) mexErrMsgIdAndTxt("hgmres:unassigned", "Output argument \"x\" not assigned (breakdown at k=1).");
    plhs[0] = x;
    if (nlhs > 1) plhs[1] = column(err.data(), niters);
    if (nlhs > 2) plhs[2] = column(res.data(), niters);
    if (nlhs > 3) plhs[3] = mxCreateDoubleScalar((double)niters);
}

// gcv_val = gcv_function(lambda,A,B,b,m,k_gcv,gcv_type)
void gcv(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    if (nrhs != 7) mexErrMsgIdAndTxt("hgmres:nargin", "expected (lambda,A,B,b,m,k_gcv,gcv_type)");
    const double lambda = mxGetScalar(prhs[0]);
    const double m = mxGetScalar(prhs[4]);
    const int k = (int)mxGetScalar(prhs[5]);
    char type[8] = {0};
    mxGetString(prhs[6], type, sizeof(type));
    const int t = strcmp(type, "ab") == 0 ? 0 : 1;  // gcv_function.m:4,7: anything else is 'ba'
    const void* pa = mxGetPr(prhs[1]);
    const void* pb = mxGetPr(prhs[2]);
    const void* pr = mxGetPr(prhs[3]);
    const size_t za = mxGetNumberOfElements(prhs[1]), zb = mxGetNumberOfElements(prhs[2]);
    if (!(g_gcv.g && g_gcv.a == pa && g_gcv.b == pb && g_gcv.rhs == pr && g_gcv.nnz_a == za && g_gcv.nnz_b == zb &&
          g_gcv.k == k && g_gcv.type == t && g_gcv.m == m)) {
        if (g_gcv.g) hg_gcv_destroy(g_gcv.g);
        g_gcv = GcvCache();
        Mat A, B;
        A.upload(prhs[1], "A");
        B.upload(prhs[2], "B");
        const double* b = vec(prhs[3], mxGetM(prhs[1]), "b");
        if (hg_gcv_prepare(ctx(), A.m, B.m, b, (int64_t)m, k, t, &g_gcv.g) != HG_OK) fail("gcv_function");
        g_gcv.a = pa; g_gcv.b = pb; g_gcv.rhs = pr; g_gcv.nnz_a = za; g_gcv.nnz_b = zb;
        g_gcv.k = k; g_gcv.type = t; g_gcv.m = m;
    }
    double val = 0.0;
    if (hg_gcv_eval(g_gcv.g, lambda, &val) != HG_OK) fail("gcv_function");
    plhs[0] = mxCreateDoubleScalar(val);
}

// [x,error_norm,residual_norm,niters] = {hybrid_lsqr,hybrid_lsmr}_solver(A,b,x_true,tol,maxit,lambda)
//                                     = lsqr_solver(A,b,x_true,tol,maxit)
// [x,err,res,ar,iters]                = lsmr_solver(A,b,x_true,tol,maxit)   (tol, maxit optional)
void gkb(const std::string& name, int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const bool hybrid = name.rfind("hybrid_", 0) == 0;
    const bool lsmr5 = name == "lsmr_solver";
    if (nrhs < (lsmr5 ? 2 : (hybrid ? 6 : 5))) mexErrMsgIdAndTxt("hgmres:nargin", "too few input arguments");
    // MATLAB's CSC of A is the CSR of A': upload it as At directly and derive A on the device.
    Mat A, At;
    A.upload(prhs[0], "A");
    if (hg_matrix_transpose(ctx(), A.m, &At.m) != HG_OK) fail("transpose");
    const size_t m = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    const double* b = vec(prhs[1], m, "b");
    const double* xt = nullptr;
    if (nrhs >= 3 && !mxIsEmpty(prhs[2])) xt = vec(prhs[2], n, "x_true");
    double tol = 1e-6;                             // lsmr_solver.m:3
    int maxit = (int)(m < n ? m : n);              // lsmr_solver.m:5
    if (nrhs >= 4 && !mxIsEmpty(prhs[3])) tol = mxGetScalar(prhs[3]);
    if (nrhs >= 5 && !mxIsEmpty(prhs[4])) maxit = (int)mxGetScalar(prhs[4]);
    const double lambda = hybrid ? mxGetScalar(prhs[5]) : 0.0;
    std::vector<double> err((size_t)maxit), res((size_t)maxit), ar((size_t)maxit);
    mxArray* x = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    int it = 0, st;
    if (name == "hybrid_lsqr_solver")
        st = hg_hybrid_lsqr_solver(ctx(), A.m, At.m, b, xt, tol, maxit, lambda, mxGetPr(x), err.data(), res.data(), &it, nullptr);
    else if (name == "hybrid_lsmr_solver")
        st = hg_hybrid_lsmr_solver(ctx(), A.m, At.m, b, xt, tol, maxit, lambda, mxGetPr(x), err.data(), res.data(), &it, nullptr);
    else if (name == "lsqr_solver")
        st = hg_lsqr_solver(ctx(), A.m, At.m, b, xt, tol, maxit, mxGetPr(x), err.data(), res.data(), &it, nullptr);
    else
        st = hg_lsmr_solver(ctx(), A.m, At.m, b, xt, tol, maxit, mxGetPr(x), err.data(), res.data(), ar.data(), &it, nullptr);
    if (st != HG_OK) fail(name.c_str());
    plhs[0] = x;
    if (nlhs > 1) plhs[1] = column(err.data(), it);
    if (nlhs > 2) plhs[2] = column(res.data(), it);
    if (lsmr5) {
        if (nlhs > 3) plhs[3] = column(ar.data(), it);
        if (nlhs > 4) plhs[4] = mxCreateDoubleScalar((double)it);
    } else if (nlhs > 3) {
        plhs[3] = mxCreateDoubleScalar((double)it);
    }
}

}  // namespace

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const std::string name = mexFunctionName();
    if (name == "hybrid_ab_gmres_rtp") rtp(true, nlhs, plhs, nrhs, prhs);
    else if (name == "hybrid_ba_gmres_rtp") rtp(false, nlhs, plhs, nrhs, prhs);
    else if (name == "gcv_function") gcv(nlhs, plhs, nrhs, prhs);
    else if (name == "hybrid_lsqr_solver" || name == "hybrid_lsmr_solver" || name == "lsqr_solver" ||
             name == "lsmr_solver") gkb(name, nlhs, plhs, nrhs, prhs);
    else mexErrMsgIdAndTxt("hgmres:name", "hgmres_mex installed under an unknown name '%s'", name.c_str());
}
