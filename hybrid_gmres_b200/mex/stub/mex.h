/* Minimal declarations of the MEX C API used by hgmres_mex.cpp.  The real mex.h ships with MATLAB /
 * Octave, neither of which exists in the build container; this header serves the syntax check
 * (`g++ -fsyntax-only -Istub`) and the mock MATLAB API the gateway is EXECUTED against in
 * tests/test_gpu_mex.py (tests/mex_mock/mex_mock.cpp implements these functions). */
#ifndef HG_STUB_MEX_H
#define HG_STUB_MEX_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
#ifdef __cplusplus
extern "C" {
#endif
bool mxIsSparse(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsChar(const mxArray*);
bool mxIsEmpty(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
double* mxGetPr(const mxArray*);
mwIndex* mxGetIr(const mxArray*);
mwIndex* mxGetJc(const mxArray*);
double mxGetScalar(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
void* mxCalloc(size_t, size_t);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
const char* mexFunctionName(void);
int mexAtExit(void (*)(void));
void mexLock(void);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
