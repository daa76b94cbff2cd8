// C-callable wrapper around the reference's verbatim mexFunction of computecontrast5.cpp:52.
// TEST INFRASTRUCTURE ONLY.
#include "computecontrast5.cpp"
extern "C" void ref_computecontrast5(const double* pad, int m1, int n1, const double* W, int len,
                                     int win, int m, int n, double* out) {
  mxArray a0 = {(size_t)m1, (size_t)n1, (double*)pad};
  mxArray a1 = {(size_t)win, (size_t)win, (double*)W};
  double dl = len, dw = win, dm = m, dn = n;
  mxArray a2 = {1, 1, &dl}, a3 = {1, 1, &dw}, a4 = {1, 1, &dm}, a5 = {1, 1, &dn};
  const mxArray* prhs[6] = {&a0, &a1, &a2, &a3, &a4, &a5};
  mxArray* plhs[1] = {0};
  mexFunction(1, plhs, 6, prhs);
  memcpy(out, plhs[0]->pr, sizeof(double) * (size_t)m * n);
  mxDestroyArray(plhs[0]);
}
