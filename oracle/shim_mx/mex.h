/* Stand-in for MATLAB's mex.h / matrix.h, large enough for libsvm-3.17/matlab/svmpredict.c and svm_model_matlab.c to compile
 * verbatim (and for bindings/libsvmpredict_cds.cpp to be exercised without MATLAB): dense and sparse double matrices, 1x1 struct
 * arrays, char row vectors.  C and C++.  TEST INFRASTRUCTURE ONLY. */
#ifndef CDS_ORACLE_MEX_STRUCT_SHIM_H
#define CDS_ORACLE_MEX_STRUCT_SHIM_H
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef size_t mwIndex;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef struct mxArray_tag {
  size_t M, N;
  int cls;                       /* 0 double, 1 struct, 2 char */
  double* pr;                    /* dense: M*N values, column-major; sparse: nzmax values */
  int sparse; size_t nzmax; mwIndex* ir; mwIndex* jc;
  int nfields; char** names; struct mxArray_tag** fields;
  char* str;
} mxArray;

static inline mxArray* mx_new_(size_t m, size_t n, int cls) {
  mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
  a->M = m; a->N = n; a->cls = cls;
  return a;
}
static inline size_t mxGetM(const mxArray* a) { return a->M; }
static inline size_t mxGetN(const mxArray* a) { return a->N; }
static inline double* mxGetPr(const mxArray* a) { return a->pr; }
static inline int mxIsDouble(const mxArray* a) { return a->cls == 0; }
static inline int mxIsStruct(const mxArray* a) { return a->cls == 1; }
static inline int mxIsChar(const mxArray* a) { return a->cls == 2; }
static inline int mxIsSparse(const mxArray* a) { return a->sparse; }
static inline int mxIsEmpty(const mxArray* a) { return a->M == 0 || a->N == 0; }
static inline mwIndex* mxGetIr(const mxArray* a) { return a->ir; }
static inline mwIndex* mxGetJc(const mxArray* a) { return a->jc; }
static inline size_t mxGetNzmax(const mxArray* a) { return a->nzmax; }
static inline void* mxMalloc(size_t n) { return malloc(n ? n : 1); }
static inline void* mxCalloc(size_t n, size_t s) { return calloc(n ? n : 1, s ? s : 1); }
static inline void mxFree(void* p) { free(p); }
static inline mxArray* mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c) {
  mxArray* a = mx_new_(m, n, 0); (void)c;
  a->pr = (double*)calloc(m * n ? m * n : 1, sizeof(double));
  return a;
}
static inline mxArray* mxCreateSparse(size_t m, size_t n, size_t nzmax, mxComplexity c) {
  mxArray* a = mx_new_(m, n, 0); (void)c;
  a->sparse = 1; a->nzmax = nzmax ? nzmax : 1;
  a->pr = (double*)calloc(a->nzmax, sizeof(double));
  a->ir = (mwIndex*)calloc(a->nzmax, sizeof(mwIndex));
  a->jc = (mwIndex*)calloc(n + 1, sizeof(mwIndex));
  return a;
}
static inline mxArray* mxCreateStructMatrix(size_t m, size_t n, int nfields, const char** names) {
  mxArray* a = mx_new_(m, n, 1);
  a->nfields = nfields;
  a->names = (char**)calloc(nfields ? nfields : 1, sizeof(char*));
  a->fields = (mxArray**)calloc(nfields ? nfields : 1, sizeof(mxArray*));
  for (int i = 0; i < nfields; i++) { a->names[i] = (char*)malloc(strlen(names[i]) + 1); strcpy(a->names[i], names[i]); }
  return a;
}
static inline mxArray* mxCreateString(const char* s) {
  mxArray* a = mx_new_(1, strlen(s), 2);
  a->str = (char*)malloc(strlen(s) + 1); strcpy(a->str, s);
  return a;
}
static inline int mxGetNumberOfFields(const mxArray* a) { return a->nfields; }
static inline mxArray* mxGetFieldByNumber(const mxArray* a, mwIndex index, int field) { (void)index; return (field >= 0 && field < a->nfields) ? a->fields[field] : NULL; }
static inline mxArray* mxGetField(const mxArray* a, mwIndex index, const char* name) {
  (void)index;
  for (int i = 0; i < a->nfields; i++) if (!strcmp(a->names[i], name)) return a->fields[i];
  return NULL;
}
static inline void mxSetField(mxArray* a, mwIndex index, const char* name, mxArray* v) {
  (void)index;
  for (int i = 0; i < a->nfields; i++) if (!strcmp(a->names[i], name)) { a->fields[i] = v; return; }
}
static inline int mxGetString(const mxArray* a, char* buf, mwSize len) {
  if (a->cls != 2 || !a->str || strlen(a->str) + 1 > len) return 1;
  strcpy(buf, a->str);
  return 0;
}
static inline void mxDestroyArray(mxArray* a) {
  if (!a) return;
  for (int i = 0; i < a->nfields; i++) { mxDestroyArray(a->fields[i]); free(a->names[i]); }
  free(a->names); free(a->fields); free(a->pr); free(a->ir); free(a->jc); free(a->str); free(a);
}
static inline mxArray* mxDuplicateArray(const mxArray* s) {
  if (!s) return NULL;
  mxArray* a = mx_new_(s->M, s->N, s->cls);
  if (s->sparse) {
    a->sparse = 1; a->nzmax = s->nzmax;
    a->pr = (double*)malloc(s->nzmax * sizeof(double)); memcpy(a->pr, s->pr, s->nzmax * sizeof(double));
    a->ir = (mwIndex*)malloc(s->nzmax * sizeof(mwIndex)); memcpy(a->ir, s->ir, s->nzmax * sizeof(mwIndex));
    a->jc = (mwIndex*)malloc((s->N + 1) * sizeof(mwIndex)); memcpy(a->jc, s->jc, (s->N + 1) * sizeof(mwIndex));
  } else if (s->pr) {
    const size_t n = s->M * s->N ? s->M * s->N : 1;
    a->pr = (double*)malloc(n * sizeof(double)); memcpy(a->pr, s->pr, n * sizeof(double));
  }
  if (s->str) { a->str = (char*)malloc(strlen(s->str) + 1); strcpy(a->str, s->str); }
  if (s->nfields) {
    a->nfields = s->nfields;
    a->names = (char**)calloc(s->nfields, sizeof(char*)); a->fields = (mxArray**)calloc(s->nfields, sizeof(mxArray*));
    for (int i = 0; i < s->nfields; i++) { a->names[i] = (char*)malloc(strlen(s->names[i]) + 1); strcpy(a->names[i], s->names[i]); a->fields[i] = mxDuplicateArray(s->fields[i]); }
  }
  return a;
}
/* what MATLAB prints / raises is recorded; tests read it back through mx_shim_last_message() */
#ifdef MX_SHIM_DEFINE_MESSAGE            /* one definition per shared object: the harness */
char g_mx_shim_message[1024];
#else
extern char g_mx_shim_message[1024];
#endif
static inline int mexPrintf(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  const size_t n = strlen(g_mx_shim_message);
  vsnprintf(g_mx_shim_message + n, sizeof(g_mx_shim_message) - n, fmt, ap);
  va_end(ap);
  return 0;
}
static inline void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  int n = snprintf(g_mx_shim_message, sizeof(g_mx_shim_message), "%s: ", id);
  vsnprintf(g_mx_shim_message + n, sizeof(g_mx_shim_message) - n, fmt, ap);
  va_end(ap);
}
/* the two MATLAB functions libsvm's interface calls back: transpose and full (svmpredict.c:97-118, svm_model_matlab.c:192,333) */
static inline int mexCallMATLAB(int nlhs, mxArray** plhs, int nrhs, mxArray** prhs, const char* fn) {
  if (nlhs != 1 || nrhs != 1 || !prhs[0] || prhs[0]->cls != 0) return 1;
  const mxArray* a = prhs[0];
  if (!strcmp(fn, "transpose") && a->sparse) {
    const size_t nnz = a->jc[a->N];
    mxArray* t = mxCreateSparse(a->N, a->M, nnz, mxREAL);
    size_t* cnt = (size_t*)calloc(a->M + 1, sizeof(size_t));
    for (size_t k = 0; k < nnz; k++) cnt[a->ir[k] + 1]++;
    for (size_t r = 0; r < a->M; r++) cnt[r + 1] += cnt[r];
    memcpy(t->jc, cnt, (a->M + 1) * sizeof(size_t));
    for (size_t c = 0; c < a->N; c++)
      for (size_t k = a->jc[c]; k < a->jc[c + 1]; k++) { const size_t p = cnt[a->ir[k]]++; t->ir[p] = c; t->pr[p] = a->pr[k]; }
    free(cnt);
    plhs[0] = t;
    return 0;
  }
  if (!strcmp(fn, "transpose")) {
    mxArray* t = mxCreateDoubleMatrix(a->N, a->M, mxREAL);
    for (size_t c = 0; c < a->N; c++) for (size_t r = 0; r < a->M; r++) t->pr[r * a->N + c] = a->pr[c * a->M + r];
    plhs[0] = t;
    return 0;
  }
  if (!strcmp(fn, "full")) {
    mxArray* t = mxCreateDoubleMatrix(a->M, a->N, mxREAL);
    if (a->sparse) { for (size_t c = 0; c < a->N; c++) for (size_t k = a->jc[c]; k < a->jc[c + 1]; k++) t->pr[c * a->M + a->ir[k]] = a->pr[k]; }
    else memcpy(t->pr, a->pr, a->M * a->N * sizeof(double));
    plhs[0] = t;
    return 0;
  }
  return 1;
}
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
