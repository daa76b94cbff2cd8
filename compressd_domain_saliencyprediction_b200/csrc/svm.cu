// svm.cu — stage (d): libsvm C-SVC / RBF decision values for every block against every SV.
//
// Replaces libsvm-3.17 svm_predict_values (svm.cpp:2501-2575, C-SVC branch with nr_class == 2) and
// Kernel::k_function RBF (svm.cpp:325-365) as driven by matlab/svmpredict.c:165-204:
//     dec(x) = sum_i coef_i * exp(-gamma * ||x - sv_i||^2) - rho ,  label = dec > 0 ? Label(1) : Label(2)
//
// B200 design (FP32 CUDA cores + MUFU.EX2): with g2 = gamma*log2(e), x' = sqrt(2 g2) x, s' = sqrt(2 g2) s,
//     -g2 ||x-s||^2 = x'.s' - ||x'||^2/2 - ||s'||^2/2
// so one pair costs 1 FADD + 9 FFMA + 1 MUFU.EX2 + 1 FFMA.  Each thread keeps IPT instances in
// registers; SVs stream through shared memory in 48-byte records read as three broadcast LDS.128.
#include "cds_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

struct cds_svm_model {
  int nsv = 0, nfeat = 0;
  double gamma = 0, rho = 0;
  int label0 = 1, label1 = -1;
  int nsv_pad = 0;               // multiple of SV_CHUNK
  float* d_rec = nullptr;        // F == 9 fast path: [nsv_pad][12] = 9 scaled feats, -||s'||^2/2, coef, 0
  float* d_sv = nullptr;         // generic path: [nsv][nfeat] unscaled
  float* d_coef = nullptr;       // generic path: [nsv]
};

namespace cds {

constexpr int SV_CHUNK = 256;
constexpr int SVM_IPT = 4;
constexpr int SVM_THREADS = 128;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// X row-major [N][9] (unscaled); rec as described above.  dec[N] = sum - rho.
__global__ void __launch_bounds__(SVM_THREADS)
svm_rbf9_kernel(const float* __restrict__ X, long N, const float4* __restrict__ rec, int nsv_pad,
                float xscale, float rho, float* __restrict__ dec) {
  __shared__ float4 sbuf[2][SV_CHUNK * 3];
  const long base = ((long)blockIdx.x * SVM_THREADS + threadIdx.x) * SVM_IPT;
  float x[SVM_IPT][9], na[SVM_IPT], acc[SVM_IPT];
#pragma unroll
  for (int i = 0; i < SVM_IPT; i++) {
    const long n = base + i;
    float s2 = 0.f;
#pragma unroll
    for (int d = 0; d < 9; d++) {
      const float v = (n < N) ? __ldg(X + n * 9 + d) * xscale : 0.f;
      x[i][d] = v; s2 = fmaf(v, v, s2);
    }
    na[i] = -0.5f * s2; acc[i] = 0.f;
  }
  const int nchunks = nsv_pad / SV_CHUNK;
  for (int t = threadIdx.x; t < SV_CHUNK * 3; t += SVM_THREADS) sbuf[0][t] = __ldg(rec + t);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ch++) {
    const int cur = ch & 1;
    if (ch + 1 < nchunks)
      for (int t = threadIdx.x; t < SV_CHUNK * 3; t += SVM_THREADS)
        sbuf[cur ^ 1][t] = __ldg(rec + (size_t)(ch + 1) * SV_CHUNK * 3 + t);
    const float4* sb = sbuf[cur];
#pragma unroll 2
    for (int s = 0; s < SV_CHUNK; s++) {
      const float4 a = sb[s * 3 + 0], b = sb[s * 3 + 1], c = sb[s * 3 + 2];   // c = {f8, -|s'|^2/2, coef, 0}
#pragma unroll
      for (int i = 0; i < SVM_IPT; i++) {
        float e = c.y + na[i];
        e = fmaf(x[i][0], a.x, e); e = fmaf(x[i][1], a.y, e); e = fmaf(x[i][2], a.z, e);
        e = fmaf(x[i][3], a.w, e); e = fmaf(x[i][4], b.x, e); e = fmaf(x[i][5], b.y, e);
        e = fmaf(x[i][6], b.z, e); e = fmaf(x[i][7], b.w, e); e = fmaf(x[i][8], c.x, e);
        acc[i] = fmaf(c.z, ex2_approx(e), acc[i]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < SVM_IPT; i++)
    if (base + i < N) dec[base + i] = acc[i] - rho;
}

// Generic feature count (<= 64): difference form in libsvm's own order, one instance per thread.
__global__ void __launch_bounds__(128)
svm_rbf_generic_kernel(const float* __restrict__ X, long N, int F, const float* __restrict__ SV,
                       const float* __restrict__ coef, int nsv, float g2, float rho, float* __restrict__ dec) {
  extern __shared__ float ssv[];   // [chunk][F+1]
  const int CH = 64;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float xr[64];
  for (int d = 0; d < F; d++) xr[d] = (n < N) ? X[n * F + d] : 0.f;
  float acc = 0.f;
  for (int s0 = 0; s0 < nsv; s0 += CH) {
    const int cn = min(CH, nsv - s0);
    __syncthreads();
    for (int t = threadIdx.x; t < cn * (F + 1); t += blockDim.x) {
      const int s = t / (F + 1), d = t - s * (F + 1);
      ssv[t] = d < F ? SV[(size_t)(s0 + s) * F + d] : coef[s0 + s];
    }
    __syncthreads();
    for (int s = 0; s < cn; s++) {
      const float* sv = ssv + s * (F + 1);
      float d2 = 0.f;
      for (int d = 0; d < F; d++) { const float t = xr[d] - sv[d]; d2 = fmaf(t, t, d2); }
      acc = fmaf(sv[F], ex2_approx(-g2 * d2), acc);
    }
  }
  if (n < N) dec[n] = acc - rho;
}

// column-major double N x F  ->  row-major float [N][F]
__global__ void colmajor_d2f_kernel(const double* __restrict__ in, float* __restrict__ out, long N, int F) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < N * F) { const long n = i / F; const int d = (int)(i - n * F); out[i] = (float)in[(size_t)d * N + n]; }
}
__global__ void dec_f2d_label_kernel(const float* __restrict__ dec, double* __restrict__ out, double* __restrict__ label,
                                     long N, double l0, double l1) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < N) { const float d = dec[i]; out[i] = (double)d; if (label) label[i] = d > 0.f ? l0 : l1; }
}

int svm_predict_launch(const cds_svm_model* m, const float* X, long N, float* dec, cudaStream_t st) {
  const double g2 = m->gamma * 1.4426950408889634;   // gamma * log2(e)
  if (m->nfeat == 9) {
    const long per_block = (long)SVM_THREADS * SVM_IPT;
    const unsigned grid = (unsigned)((N + per_block - 1) / per_block);
    svm_rbf9_kernel<<<grid, SVM_THREADS, 0, st>>>(X, N, reinterpret_cast<const float4*>(m->d_rec), m->nsv_pad,
                                                 (float)sqrt(2 * g2), (float)m->rho, dec);
  } else {
    const unsigned grid = (unsigned)((N + 127) / 128);
    svm_rbf_generic_kernel<<<grid, 128, 64 * (m->nfeat + 1) * 4, st>>>(X, N, m->nfeat, m->d_sv, m->d_coef, m->nsv,
                                                                      (float)g2, (float)m->rho, dec);
  }
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

static std::mutex g_svm_mu;
static DevBuf g_xd, g_xf, g_dec, g_decd, g_lab;

static int model_upload(cds_svm_model* m, const double* SV, const double* coef) {
  const int F = m->nfeat, nsv = m->nsv;
  if (F == 9) {
    m->nsv_pad = ceil_div(nsv, SV_CHUNK) * SV_CHUNK;
    std::vector<float> rec((size_t)m->nsv_pad * 12, 0.f);
    const double sc = sqrt(2 * m->gamma * 1.4426950408889634);
    for (int i = 0; i < nsv; i++) {
      double s2 = 0;
      for (int d = 0; d < 9; d++) {
        const float v = (float)(SV[(size_t)i * 9 + d] * sc);
        rec[(size_t)i * 12 + d] = v; s2 += (double)v * v;
      }
      rec[(size_t)i * 12 + 9] = (float)(-0.5 * s2);
      rec[(size_t)i * 12 + 10] = (float)coef[i];
    }
    // padding records: coef 0 and a hugely negative exponent bias -> contribute exactly 0
    for (int i = nsv; i < m->nsv_pad; i++) rec[(size_t)i * 12 + 9] = -1e30f;
    CDS_CUDA(cudaMalloc(&m->d_rec, rec.size() * 4));
    CDS_CUDA(cudaMemcpy(m->d_rec, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<float> sv((size_t)nsv * F), cf(nsv);
    for (size_t i = 0; i < sv.size(); i++) sv[i] = (float)SV[i];
    for (int i = 0; i < nsv; i++) cf[i] = (float)coef[i];
    CDS_CUDA(cudaMalloc(&m->d_sv, sv.size() * 4));
    CDS_CUDA(cudaMalloc(&m->d_coef, cf.size() * 4));
    CDS_CUDA(cudaMemcpy(m->d_sv, sv.data(), sv.size() * 4, cudaMemcpyHostToDevice));
    CDS_CUDA(cudaMemcpy(m->d_coef, cf.data(), cf.size() * 4, cudaMemcpyHostToDevice));
  }
  return CDS_OK;
}

}  // namespace cds

using namespace cds;

extern "C" int cds_svm_model_create(int nsv, int nfeat, const double* SV, const double* coef, double rho,
                                    double gamma, int label0, int label1, cds_svm_model** out) {
  if (int e = ensure_device()) return e;
  if (!SV || !coef || !out || nsv < 1 || nfeat < 1 || nfeat > 64 || !(gamma > 0)) {
    set_error("cds_svm_model_create: bad arguments (1 <= nfeat <= 64, gamma > 0)"); return CDS_EINVAL;
  }
  cds_svm_model* m = new cds_svm_model();
  m->nsv = nsv; m->nfeat = nfeat; m->gamma = gamma; m->rho = rho; m->label0 = label0; m->label1 = label1;
  if (int e = model_upload(m, SV, coef)) { cds_svm_model_free(m); return e; }
  *out = m;
  return CDS_OK;
}

// libsvm text model (svm.cpp:2756 svm_save_model / :2641 svm_load_model)
extern "C" int cds_svm_model_load(const char* path, cds_svm_model** out) {
  if (!path || !out) { set_error("cds_svm_model_load: null argument"); return CDS_EINVAL; }
  FILE* f = fopen(path, "r");
  if (!f) { set_error("cds_svm_model_load: cannot open %s", path); return CDS_EINVAL; }
  char key[128];
  int nr_class = 0, total_sv = 0, label[2] = {1, -1};
  double gamma = 0, rho = 0;
  bool ok_type = false, ok_kernel = false, in_sv = false;
  while (!in_sv && fscanf(f, "%127s", key) == 1) {
    if (!strcmp(key, "svm_type")) { char v[64]; if (fscanf(f, "%63s", v) != 1) break; ok_type = !strcmp(v, "c_svc"); }
    else if (!strcmp(key, "kernel_type")) { char v[64]; if (fscanf(f, "%63s", v) != 1) break; ok_kernel = !strcmp(v, "rbf"); }
    else if (!strcmp(key, "gamma")) { if (fscanf(f, "%lf", &gamma) != 1) break; }
    else if (!strcmp(key, "nr_class")) { if (fscanf(f, "%d", &nr_class) != 1) break; }
    else if (!strcmp(key, "total_sv")) { if (fscanf(f, "%d", &total_sv) != 1) break; }
    else if (!strcmp(key, "rho")) { if (fscanf(f, "%lf", &rho) != 1) break; }
    else if (!strcmp(key, "label")) { if (fscanf(f, "%d %d", &label[0], &label[1]) != 2) break; }
    else if (!strcmp(key, "nr_sv")) { int a, b; if (fscanf(f, "%d %d", &a, &b) != 2) break; }
    else if (!strcmp(key, "probA") || !strcmp(key, "probB")) { double a; if (fscanf(f, "%lf", &a) != 1) break; }
    else if (!strcmp(key, "SV")) { in_sv = true; }
    else { break; }
  }
  if (!in_sv || !ok_type || !ok_kernel || nr_class != 2 || total_sv < 1) {
    fclose(f); set_error("cds_svm_model_load: %s is not a 2-class c_svc/rbf libsvm model", path); return CDS_EFORMAT;
  }
  int c; while ((c = fgetc(f)) != EOF && c != '\n') {}
  std::vector<double> coef(total_sv);
  std::vector<std::vector<std::pair<int, double>>> rows(total_sv);
  int maxidx = 0;
  std::vector<char> line(1 << 16);
  for (int i = 0; i < total_sv; i++) {
    if (!fgets(line.data(), (int)line.size(), f)) { fclose(f); set_error("cds_svm_model_load: truncated SV section"); return CDS_EFORMAT; }
    char* p = line.data(); char* end;
    coef[i] = strtod(p, &end); p = end;
    while (true) {
      long idx = strtol(p, &end, 10);
      if (end == p || *end != ':') break;
      p = end + 1;
      double v = strtod(p, &end); p = end;
      rows[i].push_back({(int)idx, v});
      if (idx > maxidx) maxidx = (int)idx;
    }
  }
  fclose(f);
  if (maxidx < 1 || maxidx > 64) { set_error("cds_svm_model_load: feature index %d out of range", maxidx); return CDS_EFORMAT; }
  std::vector<double> SV((size_t)total_sv * maxidx, 0.0);
  for (int i = 0; i < total_sv; i++) for (auto& kv : rows[i]) SV[(size_t)i * maxidx + kv.first - 1] = kv.second;
  return cds_svm_model_create(total_sv, maxidx, SV.data(), coef.data(), rho, gamma, label[0], label[1], out);
}

extern "C" void cds_svm_model_free(cds_svm_model* m) {
  if (!m) return;
  if (m->d_rec) cudaFree(m->d_rec);
  if (m->d_sv) cudaFree(m->d_sv);
  if (m->d_coef) cudaFree(m->d_coef);
  delete m;
}
extern "C" int cds_svm_model_nsv(const cds_svm_model* m) { return m ? m->nsv : 0; }
extern "C" int cds_svm_model_nfeat(const cds_svm_model* m) { return m ? m->nfeat : 0; }

extern "C" int cds_svm_rbf_predict_dev(const cds_svm_model* m, const float* X, long N, float* dec, void* stream) {
  if (int e = ensure_device()) return e;
  if (!m || !X || !dec || N < 1) { set_error("cds_svm_rbf_predict_dev: bad arguments"); return CDS_EINVAL; }
  return svm_predict_launch(m, X, N, dec, (cudaStream_t)stream);
}

extern "C" int cds_svm_rbf_predict(const cds_svm_model* m, const double* X, int N, int F, double* dec, double* label) {
  if (int e = ensure_device()) return e;
  if (!m || !X || !dec || N < 1) { set_error("cds_svm_rbf_predict: bad arguments"); return CDS_EINVAL; }
  if (F != m->nfeat) { set_error("cds_svm_rbf_predict: X has %d features, model has %d", F, m->nfeat); return CDS_EINVAL; }
  std::lock_guard<std::mutex> lk(g_svm_mu);
  const size_t n = (size_t)N * F;
  if (int e = g_xd.reserve(n * 8)) return e;
  if (int e = g_xf.reserve(n * 4)) return e;
  if (int e = g_dec.reserve((size_t)N * 4)) return e;
  if (int e = g_decd.reserve((size_t)N * 8)) return e;
  if (int e = g_lab.reserve((size_t)N * 8)) return e;
  cudaStream_t st = 0;
  CDS_CUDA(cudaMemcpyAsync(g_xd.p, X, n * 8, cudaMemcpyHostToDevice, st));
  colmajor_d2f_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g_xd.as<double>(), g_xf.as<float>(), N, F);
  CDS_LAUNCH_CHECK();
  if (int e = svm_predict_launch(m, g_xf.as<float>(), N, g_dec.as<float>(), st)) return e;
  dec_f2d_label_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(g_dec.as<float>(), g_decd.as<double>(),
                                                                    label ? g_lab.as<double>() : nullptr, N,
                                                                    (double)m->label0, (double)m->label1);
  CDS_LAUNCH_CHECK();
  CDS_CUDA(cudaMemcpyAsync(dec, g_decd.p, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
  if (label) CDS_CUDA(cudaMemcpyAsync(label, g_lab.p, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  return CDS_OK;
}
