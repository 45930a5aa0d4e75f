/* mex.h -- hand-written STUB of the subset of MATLAB's MEX C API used by the gateways in this directory.
 * MATLAB is not installed in the build container, so the gateways can only be compile-checked; on a MATLAB box the
 * real <mex.h> shadows this file (mex puts its own include directory first).  Declarations only. */
#ifndef QGMAP_STUB_MEX_H
#define QGMAP_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef unsigned char mxLogical;
typedef enum { mxUNKNOWN_CLASS = 0, mxLOGICAL_CLASS = 3, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxUINT8_CLASS = 9,
               mxUINT64_CLASS = 15, mxSTRUCT_CLASS = 2 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...);
int mexPrintf(const char *fmt, ...);
int mexAtExit(void (*fn)(void));
void mexLock(void);
void mexUnlock(void);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateLogicalArray(mwSize ndim, const mwSize *dims);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names);
void mxSetField(mxArray *s, mwIndex i, const char *name, mxArray *v);
mxArray *mxGetField(const mxArray *s, mwIndex i, const char *name);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
double mxGetScalar(const mxArray *a);
mxLogical *mxGetLogicals(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsComplex(const mxArray *a);
int mxIsLogical(const mxArray *a);
int mxIsStruct(const mxArray *a);
int mxIsChar(const mxArray *a);
int mxIsEmpty(const mxArray *a);
int mxIsUint8(const mxArray *a);
char *mxArrayToString(const mxArray *a);
void mxFree(void *p);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
