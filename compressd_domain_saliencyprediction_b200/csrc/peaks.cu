// peaks.cu — micro-kernels that measure the FP32-pipe and MUFU.EX2 issue ceilings of the device the
// library runs on.  MEASURED_PEAKS.json only has HBM and bf16-tensor numbers; the contrast and SVM
// kernels are FP32/MUFU bound, so bench.py measures their roofline denominators live with these.
#include "cds_common.cuh"

namespace cds {

// 16 independent chains of the contrast kernel's own instruction pair (FADD ; FFMA |.|), no memory.
__global__ void __launch_bounds__(256) fp32_pair_peak_kernel(float* out, float seed, int iters, float wgt) {
  float a[16], c[16];
#pragma unroll
  for (int k = 0; k < 16; k++) { a[k] = 0.f; c[k] = seed * (k + 1) + threadIdx.x * 1e-6f; }
  float p = seed * 0.37f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int k = 0; k < 16; k++) a[k] = fmaf(fabsf(p - c[k]), wgt, a[k]);   // weight as a constant-bank/uniform operand, like the real kernel
      p += 1e-7f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float seed, int iters) {
  float a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = seed * (k + 1) + threadIdx.x * 1e-6f;
  const float m = 0.9999f + seed * 1e-9f, b = seed * 1e-3f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < 16; k++) a[k] = fmaf(a[k], m, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, float seed, int iters) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = -seed * (k + 1) - threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[k])); a[k] = -y; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 123.456f) out[0] = s;
}

}  // namespace cds

using namespace cds;

// Returns lane-instructions per second: fp32_pair (FADD+FFMA mix, the contrast pair), ffma (pure FFMA),
// mufu (MUFU.EX2).  Each is the best of `reps` timed launches that fill every SM with 8 CTAs of 256 threads.
extern "C" int cds_measure_peaks(double* fp32_pair_ips, double* ffma_ips, double* mufu_ips) {
  if (int e = ensure_device()) return e;
  int dev = 0, sms = 0;
  CDS_CUDA(cudaGetDevice(&dev));
  CDS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* d = nullptr;
  CDS_CUDA(cudaMalloc(&d, 256));
  cudaEvent_t e0, e1;
  CDS_CUDA(cudaEventCreate(&e0)); CDS_CUDA(cudaEventCreate(&e1));
  const int grid = sms * 8, iters = 2000, reps = 5;
  double best[3] = {0, 0, 0};
  for (int which = 0; which < 3; which++) {
    for (int r = 0; r < reps + 1; r++) {
      CDS_CUDA(cudaEventRecord(e0));
      if (which == 0) fp32_pair_peak_kernel<<<grid, 256>>>(d, 0.5f, iters, 0.999f);
      else if (which == 1) ffma_peak_kernel<<<grid, 256>>>(d, 0.5f, iters);
      else mufu_peak_kernel<<<grid, 256>>>(d, 0.5f, iters);
      CDS_LAUNCH_CHECK();
      CDS_CUDA(cudaEventRecord(e1));
      CDS_CUDA(cudaEventSynchronize(e1));
      float ms = 0; CDS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r == 0) continue;   // warm-up
      const double per_thread = which == 0 ? (double)iters * 8 * 16 * 2 : which == 1 ? (double)iters * 16 * 16 : (double)iters * 16 * 8;
      const double ips = per_thread * grid * 256.0 / (ms * 1e-3);
      if (ips > best[which]) best[which] = ips;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  if (fp32_pair_ips) *fp32_pair_ips = best[0];
  if (ffma_ips) *ffma_ips = best[1];
  if (mufu_ips) *mufu_ips = best[2];
  return CDS_OK;
}
