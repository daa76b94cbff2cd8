// libsvmpredict_cds.cpp — drop-in for libsvm-3.17's MATLAB predictor (matlab/svmpredict.c:272 mexFunction), computing the decision
// values on the B200 through libcds_b200.so.
//
//   [predicted_label, accuracy, decision_values] = libsvmpredict(testing_label_vector, testing_instance_matrix, model [, 'options'])
//
// main.m:297 calls it once per picture with 40 000 x 9 instances.  Build with MATLAB:
//   mex -I<repo>/include libsvmpredict_cds.cpp -L<libdir> -lcds_b200 -output libsvmpredict
// Supported: the configuration the reference uses — C-SVC, RBF kernel, two classes, decision values (no '-b 1').  The model struct is
// read field by field in the fixed order of svm_model_matlab.c:17-29, as matlab_matrix_to_model (:211-375) does.  Error style follows
// svmpredict.c: a message through mexPrintf and empty outputs (fake_answer, :36-47).  The model is parsed on every call like in the
// reference (:335,361); cache the cds_svm_model by mxArray pointer when calling in a loop.
#include <string.h>
#include <vector>
#include "mex.h"
#include "cds.h"

static void fake_answer(int nlhs, mxArray* plhs[]) {
  for (int i = 0; i < nlhs; i++) plhs[i] = mxCreateDoubleMatrix(0, 0, mxREAL);
}

static std::vector<double> dense_row_major(const mxArray* a) {       // l x F, sparse (CSC) or dense column-major -> row-major
  const size_t l = mxGetM(a), F = mxGetN(a);
  std::vector<double> out(l * F, 0.0);
  const double* pr = mxGetPr(a);
  if (mxIsSparse(a)) {
    const mwIndex* ir = mxGetIr(a); const mwIndex* jc = mxGetJc(a);
    for (size_t d = 0; d < F; d++) for (mwIndex k = jc[d]; k < jc[d + 1]; k++) out[(size_t)ir[k] * F + d] = pr[k];
  } else {
    for (size_t d = 0; d < F; d++) for (size_t i = 0; i < l; i++) out[i * F + d] = pr[d * l + i];
  }
  return out;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nlhs == 2 || nlhs > 3 || nrhs > 4 || nrhs < 3) { mexPrintf("Usage: [predicted_label, accuracy, decision_values] = libsvmpredict(testing_label_vector, testing_instance_matrix, model, 'libsvm_options')\n"); fake_answer(nlhs, plhs); return; }
  if (!mxIsDouble(prhs[0]) || !mxIsDouble(prhs[1])) { mexPrintf("Error: label vector and instance matrix must be double\n"); fake_answer(nlhs, plhs); return; }
  if (!mxIsStruct(prhs[2])) { mexPrintf("model file should be a struct array\n"); fake_answer(nlhs, plhs); return; }
  if (nrhs == 4) {
    char opt[256];
    if (mxGetString(prhs[3], opt, sizeof(opt)) == 0 && strstr(opt, "-b 1")) { mexPrintf("Error: probability estimates (-b 1) are not supported by the B200 path\n"); fake_answer(nlhs, plhs); return; }
  }
  const mxArray* m = prhs[2];
  if (mxGetNumberOfFields(m) != 11) { mexPrintf("Error: can't read model: number of return field is not correct.\n"); fake_answer(nlhs, plhs); return; }
  // Parameters, nr_class, totalSV, rho, Label, sv_indices, ProbA, ProbB, nSV, sv_coef, SVs
  const double* par = mxGetPr(mxGetFieldByNumber(m, 0, 0));
  const int nr_class = (int)mxGetPr(mxGetFieldByNumber(m, 0, 1))[0];
  if ((int)par[0] != 0 || (int)par[1] != 2 || nr_class != 2) { mexPrintf("Error: the B200 path supports 2-class C-SVC with an RBF kernel only\n"); fake_answer(nlhs, plhs); return; }
  const mxArray* lab = mxGetFieldByNumber(m, 0, 4);
  const mxArray* coef = mxGetFieldByNumber(m, 0, 9);
  const mxArray* SVs = mxGetFieldByNumber(m, 0, 10);
  if (mxIsEmpty(lab) || mxIsEmpty(coef) || mxIsEmpty(SVs) || mxGetM(coef) != mxGetM(SVs)) { mexPrintf("Error: can't read model: Label / sv_coef / SVs are inconsistent\n"); fake_answer(nlhs, plhs); return; }
  const int N = (int)mxGetM(prhs[1]), F = (int)mxGetN(prhs[1]);
  if ((int)mxGetM(prhs[0]) != N) { mexPrintf("Length of label vector does not match # of instances.\n"); fake_answer(nlhs, plhs); return; }
  if ((int)mxGetN(prhs[0]) != 1) { mexPrintf("label (1st argument) should be a vector (# of column is 1).\n"); fake_answer(nlhs, plhs); return; }
  if ((int)mxGetN(SVs) != F) { mexPrintf("Error: the model has %d features, the instance matrix %d\n", (int)mxGetN(SVs), F); fake_answer(nlhs, plhs); return; }
  const std::vector<double> sv = dense_row_major(SVs);
  cds_svm_model* mdl = NULL;
  if (cds_svm_model_create((int)mxGetM(SVs), F, sv.data(), mxGetPr(coef), mxGetPr(mxGetFieldByNumber(m, 0, 3))[0], par[3],
                           (int)mxGetPr(lab)[0], (int)mxGetPr(lab)[1], &mdl) != CDS_OK) { mexPrintf("Error: %s\n", cds_last_error()); fake_answer(nlhs, plhs); return; }
  std::vector<double> X;                                           // dense column-major N x F
  const double* Xp = mxGetPr(prhs[1]);
  if (mxIsSparse(prhs[1])) {
    X.assign((size_t)N * F, 0.0);
    const mwIndex* ir = mxGetIr(prhs[1]); const mwIndex* jc = mxGetJc(prhs[1]);
    for (int d = 0; d < F; d++) for (mwIndex k = jc[d]; k < jc[d + 1]; k++) X[(size_t)d * N + ir[k]] = Xp[k];
    Xp = X.data();
  }
  mxArray* out_label = mxCreateDoubleMatrix(N, 1, mxREAL);
  mxArray* out_dec = mxCreateDoubleMatrix(N, 1, mxREAL);
  const int rc = cds_svm_rbf_predict(mdl, Xp, N, F, mxGetPr(out_dec), mxGetPr(out_label));
  cds_svm_model_free(mdl);
  if (rc != CDS_OK) { mexPrintf("Error: %s\n", cds_last_error()); mxDestroyArray(out_label); mxDestroyArray(out_dec); fake_answer(nlhs, plhs); return; }
  // accuracy, mean squared error, squared correlation coefficient (svmpredict.c:205-238)
  int correct = 0; double error = 0, sump = 0, sumt = 0, sumpp = 0, sumtt = 0, sumpt = 0;
  const double* t = mxGetPr(prhs[0]); const double* p = mxGetPr(out_label);
  for (int i = 0; i < N; i++) {
    if (p[i] == t[i]) ++correct;
    error += (p[i] - t[i]) * (p[i] - t[i]);
    sump += p[i]; sumt += t[i]; sumpp += p[i] * p[i]; sumtt += t[i] * t[i]; sumpt += p[i] * t[i];
  }
  const double total = N;
  mxArray* acc = mxCreateDoubleMatrix(3, 1, mxREAL);
  mxGetPr(acc)[0] = (double)correct / total * 100;
  mxGetPr(acc)[1] = error / total;
  mxGetPr(acc)[2] = ((total * sumpt - sump * sumt) * (total * sumpt - sump * sumt)) / ((total * sumpp - sump * sump) * (total * sumtt - sumt * sumt));
  plhs[0] = out_label;
  if (nlhs == 3) { plhs[1] = acc; plhs[2] = out_dec; } else { mxDestroyArray(acc); mxDestroyArray(out_dec); }
}
