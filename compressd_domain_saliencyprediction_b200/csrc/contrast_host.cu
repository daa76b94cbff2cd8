// contrast_host.cu — host-buffer drop-ins for Sudoku_contrastcpp5.m and cu_contrast5.m built from the
// same device kernels the clip pipeline uses (block mean, +1e-8, zero pad, pm_norm, contrast, /max,
// replicate).  Double in / double out like the MATLAB functions; arithmetic is FP32 on the device.
#include "maps.cuh"
#include <math.h>
#include <mutex>
#include <vector>

namespace cds {
int contrast_sep_launch(const float* pad, float* out, unsigned* mx, int nmaps, int R, int C, int win,
                        const float* gr, const float* gc, float* dW_scratch, cudaStream_t st, int ctas_per_sm, unsigned* sched);
int contrast_generic_launch(const float* pad, const float* dW, float* out, unsigned* mx, int nmaps,
                            int R, int C, int win, int len, cudaStream_t st);
void gaussian_factor(int win, double delta, float* g);
static std::mutex g_ch_mu;
static DevBuf g_src, g_sub, g_pad, g_raw, g_out, g_keys, g_W;

static int contrast_host(const double* map, int rows, int cols, int block_cus, int win, double delta, double* out) {
  if (int e = ensure_device()) return e;
  if (!map || !out || rows < 1 || cols < 1 || block_cus < 1 || win < 1 || win > 64 || !(delta > 0)) {
    set_error("contrast: bad arguments (window 1..64, delta > 0)"); return CDS_EINVAL;
  }
  std::lock_guard<std::mutex> lk(g_ch_mu);
  const int srows = ceil_div(rows, block_cus), scols = ceil_div(cols, block_cus), len = win / 2;
  const size_t n = (size_t)rows * cols, ns = (size_t)srows * scols, np = (size_t)(srows + 2 * len) * (scols + 2 * len);
  if (int e = g_src.reserve(n * 4)) return e;
  if (int e = g_sub.reserve(ns * 4)) return e;
  if (int e = g_pad.reserve(np * 4)) return e;
  if (int e = g_raw.reserve(ns * 4)) return e;
  if (int e = g_out.reserve(n * 4)) return e;
  if (int e = g_keys.reserve(16)) return e;
  if (int e = g_W.reserve(64 * 64 * 4)) return e;
  std::vector<float> hf(n);
  for (size_t i = 0; i < n; i++) hf[i] = (float)map[i];
  cudaStream_t st = 0;
  const unsigned init[4] = {0xffffffffu, 0u, 0u, 0u};     // min key, max key, contrast max bits
  CDS_CUDA(cudaMemcpyAsync(g_keys.p, init, 16, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemcpyAsync(g_src.p, hf.data(), n * 4, cudaMemcpyHostToDevice, st));
  if (int e = contrast_sub_launch(g_src.as<float>(), (long)n, 0, 0, 1, rows, cols, block_cus, g_sub.as<float>(), g_keys.as<unsigned>(), st)) return e;
  if (int e = contrast_pad_launch(g_sub.as<float>(), g_keys.as<unsigned>(), 1, srows, scols, len, g_pad.as<float>(), st)) return e;
  float g[64];
  gaussian_factor(win, delta, g);
  CDS_CUDA(cudaStreamSynchronize(st));
  if (int e = contrast_sep_launch(g_pad.as<float>(), g_raw.as<float>(), g_keys.as<unsigned>() + 2, 1, srows, scols, win, g, g, g_W.as<float>(), st, 0, nullptr)) return e;
  if (int e = contrast_post_launch(g_raw.as<float>(), g_keys.as<unsigned>() + 2, 1, srows, scols, block_cus, rows, cols, g_out.as<float>(), 0, st)) return e;
  unsigned hk[4];
  CDS_CUDA(cudaMemcpyAsync(hf.data(), g_out.p, n * 4, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaMemcpyAsync(hk, g_keys.p, 16, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  if (hk[2] == 0) { for (size_t i = 0; i < n; i++) out[i] = NAN; }    // 0/0 in the reference (Sudoku_contrastcpp5.m:50)
  else for (size_t i = 0; i < n; i++) out[i] = (double)hf[i];
  return CDS_OK;
}
}  // namespace cds

static int matlab_round_i(double x) { return (int)(x < 0 ? -floor(-x + 0.5) : floor(x + 0.5)); }

extern "C" int cds_sudoku_contrast(const double* map, int rows, int cols, double delta, double cusize,
                                   double patchsize, double* out) {
  if (!(cusize > 0)) { cds::set_error("cds_sudoku_contrast: cusize must be positive"); return CDS_EINVAL; }
  return cds::contrast_host(map, rows, cols, 1, matlab_round_i(patchsize / cusize), delta, out);   // Sudoku_contrastcpp5.m:6
}

extern "C" int cds_cu_contrast5(const double* map, int rows, int cols, int cusize, double delta,
                                double patchsize, double* out) {
  if (cusize < 1) { cds::set_error("cds_cu_contrast5: cusize must be >= 1"); return CDS_EINVAL; }
  return cds::contrast_host(map, rows, cols, cusize, matlab_round_i(patchsize / cusize), delta, out);   // cu_contrast5.m:34,59
}
