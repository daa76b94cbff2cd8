/* Minimal stand-in for MATLAB's mex.h so that the reference's computecontrast5.cpp
 * (root of /root/reference) compiles verbatim with g++.  TEST INFRASTRUCTURE ONLY. */
#ifndef CDS_ORACLE_MEX_SHIM_H
#define CDS_ORACLE_MEX_SHIM_H
#include <stddef.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>
typedef struct mxArray_tag { size_t M, N; double* pr; } mxArray;
enum mxComplexity { mxREAL = 0 };
static inline size_t mxGetM(const mxArray* a) { return a->M; }
static inline size_t mxGetN(const mxArray* a) { return a->N; }
static inline double* mxGetPr(const mxArray* a) { return a->pr; }
static inline mxArray* mxCreateDoubleMatrix(size_t m, size_t n, int) {
  mxArray* a = (mxArray*)malloc(sizeof(mxArray));
  a->M = m; a->N = n; a->pr = (double*)calloc(m * n ? m * n : 1, sizeof(double));
  return a;
}
static inline void mxDestroyArray(mxArray* a) { if (a) { free(a->pr); free(a); } }
/* mexErrMsgIdAndTxt aborts the mex call in MATLAB; the stand-in records the message (tests read it back). */
#include <stdarg.h>
#include <stdio.h>
static char g_mex_shim_error[512];
static inline void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  int n = snprintf(g_mex_shim_error, sizeof(g_mex_shim_error), "%s: ", id);
  vsnprintf(g_mex_shim_error + n, sizeof(g_mex_shim_error) - n, fmt, ap);
  va_end(ap);
}
#endif
