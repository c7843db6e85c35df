// Mock of the MATLAB MEX C API (the subset hybrid_gmres_b200/mex/stub/mex.h declares), so that the gateway
// hybrid_gmres_b200/mex/hgmres_mex.cpp can be EXECUTED without MATLAB (tests/test_gpu_mex.py):
//   * mxArray is a small struct (full / sparse CSC with mwIndex = size_t / char);
//   * mexErrMsgIdAndTxt longjmps back to the caller, exactly like MATLAB: no C++ destructor of the gateway
//     runs — the property the gateway's ownership rules are written for;
//   * mxCalloc blocks and mxCreate* arrays made during a failing call are reclaimed by the "interpreter".
// Test infrastructure only.
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "mex.h"

struct mxArray_tag {
    bool sparse = false, is_char = false;
    size_t m = 0, n = 0;
    double* pr = nullptr;
    mwIndex *ir = nullptr, *jc = nullptr;
    std::string str;
    bool owns = true;
};

namespace {
jmp_buf g_jmp;
bool g_in_call = false;
char g_err[1024];
std::string g_fname;
void (*g_at_exit)(void) = nullptr;
int g_locked = 0;
std::vector<void*> g_scratch;       // mxCalloc blocks of the running call
std::vector<mxArray*> g_created;    // arrays created during the running call
}  // namespace

extern "C" {
bool mxIsSparse(const mxArray* a) { return a->sparse; }
bool mxIsDouble(const mxArray* a) { return !a->is_char; }
bool mxIsComplex(const mxArray*) { return false; }
bool mxIsChar(const mxArray* a) { return a->is_char; }
bool mxIsEmpty(const mxArray* a) { return a->m == 0 || a->n == 0; }
size_t mxGetM(const mxArray* a) { return a->m; }
size_t mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
double* mxGetPr(const mxArray* a) { return a->pr; }
mwIndex* mxGetIr(const mxArray* a) { return a->ir; }
mwIndex* mxGetJc(const mxArray* a) { return a->jc; }
double mxGetScalar(const mxArray* a) { return a->pr ? a->pr[0] : 0.0; }
int mxGetString(const mxArray* a, char* buf, mwSize len) {
    if (!a->is_char || len == 0) return 1;
    strncpy(buf, a->str.c_str(), len - 1);
    buf[len - 1] = 0;
    return a->str.size() >= len;
}
void* mxCalloc(size_t n, size_t sz) {
    void* p = calloc(n ? n : 1, sz ? sz : 1);
    g_scratch.push_back(p);
    return p;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) {
    mxArray* a = new mxArray_tag();
    a->m = m;
    a->n = n;
    a->pr = static_cast<double*>(calloc(m * n ? m * n : 1, sizeof(double)));
    g_created.push_back(a);
    return a;
}
mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->pr[0] = v;
    return a;
}
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char msg[900];
    vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    snprintf(g_err, sizeof(g_err), "%s|%s", id, msg);
    if (!g_in_call) abort();
    longjmp(g_jmp, 1);  // like MATLAB: unwinds WITHOUT running C++ destructors
}
const char* mexFunctionName(void) { return g_fname.c_str(); }
int mexAtExit(void (*f)(void)) {
    g_at_exit = f;
    return 0;
}
void mexLock(void) { ++g_locked; }

// ---- the "interpreter" side, driven from Python -------------------------------------------------
void mock_free(mxArray* a) {
    if (!a) return;
    if (a->owns) {
        free(a->pr);
        free(a->ir);
        free(a->jc);
    }
    delete a;
}
mxArray* mock_new_full(size_t m, size_t n, const double* data) {
    mxArray* a = new mxArray_tag();
    a->m = m;
    a->n = n;
    a->pr = static_cast<double*>(malloc((m * n ? m * n : 1) * sizeof(double)));
    if (data && m * n) memcpy(a->pr, data, m * n * sizeof(double));
    return a;
}
// copies the CSC arrays (jc: n+1, ir/pr: nnz) into "MATLAB-owned" memory, mwIndex = 64-bit unsigned
mxArray* mock_new_sparse(size_t m, size_t n, const long long* jc, const long long* ir, const double* pr) {
    mxArray* a = new mxArray_tag();
    a->sparse = true;
    a->m = m;
    a->n = n;
    const size_t nnz = (size_t)jc[n];
    a->jc = static_cast<mwIndex*>(malloc((n + 1) * sizeof(mwIndex)));
    a->ir = static_cast<mwIndex*>(malloc((nnz ? nnz : 1) * sizeof(mwIndex)));
    a->pr = static_cast<double*>(malloc((nnz ? nnz : 1) * sizeof(double)));
    for (size_t i = 0; i <= n; ++i) a->jc[i] = (mwIndex)jc[i];
    for (size_t i = 0; i < nnz; ++i) a->ir[i] = (mwIndex)ir[i];
    if (nnz) memcpy(a->pr, pr, nnz * sizeof(double));
    return a;
}
mxArray* mock_new_string(const char* s) {
    mxArray* a = new mxArray_tag();
    a->is_char = true;
    a->str = s;
    a->m = 1;
    a->n = a->str.size();
    return a;
}
void mock_shape(const mxArray* a, size_t* m, size_t* n) {
    *m = a->m;
    *n = a->n;
}
double* mock_data(const mxArray* a) { return a->pr; }
const char* mock_last_error(void) { return g_err; }
int mock_locked(void) { return g_locked; }

// Calls mexFunction under the name `fname`.  Returns 0, or 1 when the gateway raised a MATLAB error
// (plhs entries are then NULL and everything the call allocated has been reclaimed).
int mock_call(const char* fname, int nlhs, mxArray** plhs, int nrhs, mxArray** prhs) {
    g_fname = fname;
    g_err[0] = 0;
    g_scratch.clear();
    g_created.clear();
    for (int i = 0; i < (nlhs > 0 ? nlhs : 1); ++i) plhs[i] = nullptr;
    g_in_call = true;
    int rc = 0;
    if (setjmp(g_jmp) == 0) {
        mexFunction(nlhs, plhs, nrhs, const_cast<const mxArray**>(prhs));
    } else {
        rc = 1;
        for (mxArray* a : g_created) mock_free(a);
        for (int i = 0; i < (nlhs > 0 ? nlhs : 1); ++i) plhs[i] = nullptr;
    }
    g_in_call = false;
    for (void* p : g_scratch) free(p);
    g_scratch.clear();
    g_created.clear();
    return rc;
}
void mock_unload(void) {
    if (g_at_exit) g_at_exit();
    g_at_exit = nullptr;
}
}
