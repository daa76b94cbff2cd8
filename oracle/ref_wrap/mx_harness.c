/* Calls a libsvmpredict-style mexFunction (the reference's svmpredict.c or bindings/libsvmpredict_cds.cpp, whichever this file is
 * linked with) on plain arrays: builds the label vector, the dense instance matrix and the model struct with its 11 fields in the
 * order of svm_model_matlab.c:17-29 (sparse SVs), calls mexFunction, copies the three outputs.  TEST INFRASTRUCTURE ONLY. */
#define MX_SHIM_DEFINE_MESSAGE
#include "mex.h"

static const char* kFields[11] = {"Parameters", "nr_class", "totalSV", "rho", "Label", "sv_indices", "ProbA", "ProbB", "nSV", "sv_coef", "SVs"};

static mxArray* scalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); mxGetPr(a)[0] = v; return a; }

/* SV row-major nsv x F; X column-major N x F.  Returns 0, or -1 when the call produced empty outputs (the reference's fake_answer). */
int mx_libsvmpredict(const double* labels, const double* X, int N, int F, int nsv, const double* SV, const double* coef, double rho,
                     double gamma, int label0, int label1, int nsv0, const char* options, double* out_label, double* out_acc,
                     double* out_dec, char* message, int message_len) {
  mxArray* prhs[4]; mxArray* plhs[3] = {NULL, NULL, NULL};
  g_mx_shim_message[0] = 0;
  prhs[0] = mxCreateDoubleMatrix(N, 1, mxREAL); memcpy(mxGetPr(prhs[0]), labels, sizeof(double) * N);
  prhs[1] = mxCreateDoubleMatrix(N, F, mxREAL); memcpy(mxGetPr(prhs[1]), X, sizeof(double) * N * F);
  mxArray* m = mxCreateStructMatrix(1, 1, 11, kFields);
  mxArray* par = mxCreateDoubleMatrix(5, 1, mxREAL);
  mxGetPr(par)[0] = 0; mxGetPr(par)[1] = 2; mxGetPr(par)[2] = 3; mxGetPr(par)[3] = gamma; mxGetPr(par)[4] = 0;
  mxSetField(m, 0, "Parameters", par);
  mxSetField(m, 0, "nr_class", scalar(2)); mxSetField(m, 0, "totalSV", scalar(nsv)); mxSetField(m, 0, "rho", scalar(rho));
  mxArray* lab = mxCreateDoubleMatrix(2, 1, mxREAL); mxGetPr(lab)[0] = label0; mxGetPr(lab)[1] = label1; mxSetField(m, 0, "Label", lab);
  mxArray* idx = mxCreateDoubleMatrix(nsv, 1, mxREAL); for (int i = 0; i < nsv; i++) mxGetPr(idx)[i] = i + 1; mxSetField(m, 0, "sv_indices", idx);
  mxSetField(m, 0, "ProbA", mxCreateDoubleMatrix(0, 0, mxREAL)); mxSetField(m, 0, "ProbB", mxCreateDoubleMatrix(0, 0, mxREAL));
  mxArray* ns = mxCreateDoubleMatrix(2, 1, mxREAL); mxGetPr(ns)[0] = nsv0; mxGetPr(ns)[1] = nsv - nsv0; mxSetField(m, 0, "nSV", ns);
  mxArray* sc = mxCreateDoubleMatrix(nsv, 1, mxREAL); memcpy(mxGetPr(sc), coef, sizeof(double) * nsv); mxSetField(m, 0, "sv_coef", sc);
  size_t nnz = 0;
  for (long i = 0; i < (long)nsv * F; i++) nnz += SV[i] != 0;
  mxArray* sv = mxCreateSparse(nsv, F, nnz, mxREAL);
  size_t k = 0;
  for (int d = 0; d < F; d++) {
    mxGetJc(sv)[d] = k;
    for (int i = 0; i < nsv; i++) if (SV[(size_t)i * F + d] != 0) { mxGetIr(sv)[k] = i; mxGetPr(sv)[k] = SV[(size_t)i * F + d]; k++; }
  }
  mxGetJc(sv)[F] = k;
  mxSetField(m, 0, "SVs", sv);
  prhs[2] = m;
  int nrhs = 3;
  if (options && options[0]) { prhs[3] = mxCreateString(options); nrhs = 4; }
  mexFunction(3, plhs, nrhs, (const mxArray**)prhs);
  int rc = 0;
  if (!plhs[0] || mxIsEmpty(plhs[0]) || !plhs[2] || mxIsEmpty(plhs[2])) rc = -1;
  else {
    memcpy(out_label, mxGetPr(plhs[0]), sizeof(double) * N);
    if (plhs[1] && !mxIsEmpty(plhs[1])) memcpy(out_acc, mxGetPr(plhs[1]), sizeof(double) * 3);
    memcpy(out_dec, mxGetPr(plhs[2]), sizeof(double) * N);
  }
  if (message && message_len > 0) { strncpy(message, g_mx_shim_message, message_len - 1); message[message_len - 1] = 0; }
  for (int i = 0; i < nrhs; i++) mxDestroyArray(prhs[i]);
  for (int i = 0; i < 3; i++) mxDestroyArray(plhs[i]);
  return rc;
}
