// Calls the mexFunction of bindings/computecontrast5_cds.cpp the way MATLAB would (through the mex.h stand-in).
// TEST INFRASTRUCTURE ONLY.
#include "../../bindings/computecontrast5_cds.cpp"
#include <string.h>
extern "C" const char* mexcds_last_error() { return g_mex_shim_error; }
extern "C" int mexcds_computecontrast5(const double* pad, int m1, int n1, const double* W, int wrows, int len, int win, int m, int n,
                                       double* out) {
  g_mex_shim_error[0] = 0;
  mxArray a0 = {(size_t)m1, (size_t)n1, (double*)pad};
  mxArray a1 = {(size_t)wrows, (size_t)wrows, (double*)W};
  double dl = len, dw = win, dm = m, dn = n;
  mxArray a2 = {1, 1, &dl}, a3 = {1, 1, &dw}, a4 = {1, 1, &dm}, a5 = {1, 1, &dn};
  const mxArray* prhs[6] = {&a0, &a1, &a2, &a3, &a4, &a5};
  mxArray* plhs[1] = {0};
  mexFunction(1, plhs, 6, prhs);
  if (g_mex_shim_error[0]) { if (plhs[0]) mxDestroyArray(plhs[0]); return -1; }
  memcpy(out, plhs[0]->pr, sizeof(double) * (size_t)m * n);
  mxDestroyArray(plhs[0]);
  return 0;
}
