/* TEST SCAFFOLDING: a minimal stand-in for MATLAB's mex.h (MATLAB / Octave are not in the image), just enough to
 * COMPILE the reference's unmodified gateways mex/mexGPisMap3.cpp and mex/mexGPisMap.cpp against this repo's drop-in
 * class headers (include/gpismap/) — tests/test_mex_gateways.py. Nothing here executes. */
#ifndef GPIS_TEST_MEX_STUB_H
#define GPIS_TEST_MEX_STUB_H
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>   /* the real mex.h pulls the C library headers in (matrix.h) */
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
extern "C" {
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID classid, mxComplexity flag);
mxClassID mxGetClassID(const mxArray* pa);
void* mxGetData(const mxArray* pa);
double* mxGetPr(const mxArray* pa);
const mwSize* mxGetDimensions(const mxArray* pa);
mwSize mxGetNumberOfDimensions(const mxArray* pa);
size_t mxGetNumberOfElements(const mxArray* pa);
int mxGetString(const mxArray* pa, char* buf, mwSize buflen);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
}
#endif
