// C-callable wrapper around the reference's vendored libsvm-3.17 (svm.cpp:2501 svm_predict_values,
// svm.cpp:316 Kernel::k_function).  The model is assembled field by field the way
// matlab/svm_model_matlab.c:211-375 does for a 2-class C-SVC/RBF model.  TEST INFRASTRUCTURE ONLY.
#include "svm.h"
#include <stdlib.h>
#include <string.h>
extern "C" {
// SV: nsv x nfeat row-major dense (zeros are dropped like a MATLAB sparse matrix would).
void* ref_svm_build(int nsv, int nfeat, const double* SV, const double* coef, double rho, double gamma,
                    int label0, int label1, int nsv0) {
  svm_model* m = (svm_model*)calloc(1, sizeof(svm_model));
  m->param.svm_type = C_SVC; m->param.kernel_type = RBF; m->param.gamma = gamma;
  m->param.degree = 3; m->param.coef0 = 0;
  m->nr_class = 2; m->l = nsv;
  m->rho = (double*)malloc(sizeof(double)); m->rho[0] = rho;
  m->label = (int*)malloc(2 * sizeof(int)); m->label[0] = label0; m->label[1] = label1;
  m->nSV = (int*)malloc(2 * sizeof(int)); m->nSV[0] = nsv0; m->nSV[1] = nsv - nsv0;
  m->sv_coef = (double**)malloc(sizeof(double*));
  m->sv_coef[0] = (double*)malloc(sizeof(double) * nsv);
  memcpy(m->sv_coef[0], coef, sizeof(double) * nsv);
  m->SV = (svm_node**)malloc(sizeof(svm_node*) * nsv);
  for (int i = 0; i < nsv; i++) {
    svm_node* row = (svm_node*)malloc(sizeof(svm_node) * (nfeat + 1));
    int k = 0;
    for (int j = 0; j < nfeat; j++) {
      double v = SV[(size_t)i * nfeat + j];
      if (v != 0) { row[k].index = j + 1; row[k].value = v; k++; }
    }
    row[k].index = -1; m->SV[i] = row;
  }
  m->free_sv = 1;
  return m;
}
// X: col-major N x F (MATLAB layout, matlab/svmpredict.c:165-170). dec/label: N.
void ref_svm_predict(void* model, const double* X, int N, int F, double* dec, double* label) {
  svm_model* m = (svm_model*)model;
  svm_node* x = (svm_node*)malloc(sizeof(svm_node) * (F + 1));
  for (int n = 0; n < N; n++) {
    for (int j = 0; j < F; j++) { x[j].index = j + 1; x[j].value = X[(size_t)j * N + n]; }
    x[F].index = -1;
    double d = 0;
    label[n] = svm_predict_values(m, x, &d);
    dec[n] = d;
  }
  free(x);
}
void ref_svm_free(void* model) { svm_model* m = (svm_model*)model; svm_free_and_destroy_model(&m); }
int ref_svm_save(void* model, const char* path) { return svm_save_model(path, (svm_model*)model); }
void* ref_svm_load(const char* path) { return svm_load_model(path); }
int ref_svm_nsv(void* model) { return ((svm_model*)model)->l; }
static void ref_svm_quiet(const char*) {}
// svm-train -s 0 -t 2 -c C -g gamma on dense row-major X (n x F), labels y (svm-train.c defaults otherwise)
void* ref_svm_train(const double* X, const double* y, int n, int F, double C, double gamma) {
  svm_set_print_string_function(ref_svm_quiet);
  svm_problem* prob = (svm_problem*)calloc(1, sizeof(svm_problem));
  prob->l = n; prob->y = (double*)malloc(sizeof(double) * n); prob->x = (svm_node**)malloc(sizeof(svm_node*) * n);
  for (int i = 0; i < n; i++) {
    prob->y[i] = y[i];
    svm_node* row = (svm_node*)malloc(sizeof(svm_node) * (F + 1)); int k = 0;
    for (int j = 0; j < F; j++) { double v = X[(size_t)i * F + j]; if (v != 0) { row[k].index = j + 1; row[k].value = v; k++; } }
    row[k].index = -1; prob->x[i] = row;
  }
  svm_parameter param; memset(&param, 0, sizeof(param));
  param.svm_type = C_SVC; param.kernel_type = RBF; param.degree = 3; param.gamma = gamma; param.coef0 = 0;
  param.nu = 0.5; param.cache_size = 100; param.C = C; param.eps = 1e-3; param.p = 0.1; param.shrinking = 1; param.probability = 0;
  return svm_train(prob, &param);   // the model references prob->x: prob is intentionally kept alive
}
// export a 2-class model to dense arrays
void ref_svm_export(void* model, int F, double* SV, double* coef, double* rho, double* gamma, int* labels) {
  svm_model* m = (svm_model*)model;
  for (int i = 0; i < m->l; i++) {
    for (int j = 0; j < F; j++) SV[(size_t)i * F + j] = 0;
    for (const svm_node* p = m->SV[i]; p->index != -1; ++p) SV[(size_t)i * F + p->index - 1] = p->value;
    coef[i] = m->sv_coef[0][i];
  }
  *rho = m->rho[0]; *gamma = m->param.gamma; labels[0] = m->label[0]; labels[1] = m->label[1];
}
}
