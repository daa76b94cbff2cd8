// computecontrast5_cds.cpp — drop-in for the reference's computecontrast5 mex (computecontrast5.cpp:52-96,
// shipped as computecontrast5.mexw64): same mexFunction signature and argument order
//     contrastmap = computecontrast5(subsaliencymap3, W, len, WINDOW_SIZE, m, n)      (Sudoku_contrastcpp5.m:37)
// so Sudoku_contrastcpp5.m / cu_contrast5.m run unmodified against it.  The arithmetic runs on the B200 through
// cds_contrast5 (include/cds.h).  Unlike the reference (no argument checking at all) bad arguments raise a MATLAB error.
//
//   mex -I<repo>/include computecontrast5_cds.cpp -L<repo>/compressd_domain_saliencyprediction_b200 -lcds_b200 -output computecontrast5
//
// tests/test_gpu_bindings.py compiles this file against a minimal mex.h stand-in and calls mexFunction the way MATLAB would.
#include "mex.h"
#include "cds.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 6 || nlhs > 1) { mexErrMsgIdAndTxt("cds:computecontrast5:args", "usage: out = computecontrast5(pad, W, len, winsize, m, n)"); return; }
  const double* pad = mxGetPr(prhs[0]);
  const int m1 = (int)mxGetM(prhs[0]), n1 = (int)mxGetN(prhs[0]);       // computecontrast5.cpp:64-65
  const double* W = mxGetPr(prhs[1]);                                     // :61
  const int len = (int)*mxGetPr(prhs[2]), win = (int)*mxGetPr(prhs[3]);  // :66-67
  const int m = (int)*mxGetPr(prhs[4]), n = (int)*mxGetPr(prhs[5]);      // :62-63
  if ((int)mxGetM(prhs[1]) != win || (int)mxGetN(prhs[1]) != win) { mexErrMsgIdAndTxt("cds:computecontrast5:args", "W must be winsize x winsize"); return; }
  plhs[0] = mxCreateDoubleMatrix(m > 0 ? m : 0, n > 0 ? n : 0, mxREAL);  // :73
  if (cds_contrast5(pad, m1, n1, W, len, win, m, n, mxGetPr(plhs[0])) != CDS_OK)
    mexErrMsgIdAndTxt("cds:computecontrast5:device", "%s", cds_last_error());
}
