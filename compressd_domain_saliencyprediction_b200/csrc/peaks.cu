// peaks.cu — micro-kernels that measure the FP32-pipe and MUFU.EX2 issue ceilings of the device the
// library runs on.  MEASURED_PEAKS.json only has HBM and bf16-tensor numbers; the contrast and SVM
// kernels are FP32/MUFU bound, so bench.py measures their roofline denominators live with these.
#include "cds_common.cuh"
#include "tc_common.cuh"

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

// The same two opcodes with NO dependence between them: 8 FADD chains and 8 FFMA chains, interleaved.  If this reaches the pure-FFMA
// rate, the gap between the pair mix above and pure FFMA comes from the FADD -> FFMA dependence (every FFMA of the contrast pair
// waits for its own FADD), not from mixing the opcodes on the FMA pipe.
__global__ void __launch_bounds__(256) fp32_indep_mix_peak_kernel(float* out, float seed, int iters) {
  float a[8], d[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = seed * (k + 1) + threadIdx.x * 1e-6f; d[k] = seed * (k + 2); }
  const float m = 0.9999f + seed * 1e-9f, b = seed * 1e-3f, c = seed * 1e-7f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) { d[k] = d[k] + c; a[k] = fmaf(a[k], m, b); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k] + d[k];
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


// ---- instruction-mix probes (tools/mix_probe.py): how well FFMA / FFMA2 co-issue with MUFU.EX2 and LDS ----
// which = 0: 11 FFMA : 1 MUFU.EX2 (the SVM pair), 8 independent chains
//         1: pure FFMA2 (fma.rn.f32x2), 8 chains of 2 lanes
//         2: 5 FFMA2 + 1 FFMA : 1 MUFU per pair-of-pairs equivalent (packed SVM inner product)
//         3: 11 FFMA : 1 MUFU : 0.75 LDS.128 (adds the broadcast shared-memory loads)
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(128) mix_probe_kernel(float* out, float seed, int iters, int which) {
  __shared__ float4 sb[256];
  for (int i = threadIdx.x; i < 256; i += 128) sb[i] = make_float4(seed * i, seed, -seed, 0.25f * seed);
  __syncthreads();
  float x[9], acc[8], e[8];
#pragma unroll
  for (int d = 0; d < 9; d++) x[d] = seed * (d + 1) * 1e-3f + threadIdx.x * 1e-7f;
#pragma unroll
  for (int k = 0; k < 8; k++) { acc[k] = 0.f; e[k] = -seed * k; }
  if (which == 0 || which == 3) {
    for (int it = 0; it < iters; it++) {
      float4 a = make_float4(seed, seed * 0.5f, seed * 0.25f, seed * 0.125f), b = a, c = a;
      if (which == 3) { a = sb[(it * 3) & 255]; b = sb[(it * 3 + 1) & 255]; c = sb[(it * 3 + 2) & 255]; }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float t = c.y + e[k];
        t = fmaf(x[0], a.x, t); t = fmaf(x[1], a.y, t); t = fmaf(x[2], a.z, t); t = fmaf(x[3], a.w, t);
        t = fmaf(x[4], b.x, t); t = fmaf(x[5], b.y, t); t = fmaf(x[6], b.z, t); t = fmaf(x[7], b.w, t); t = fmaf(x[8], c.x, t);
        acc[k] = fmaf(c.z, ex2f(t), acc[k]);
        if (which == 3 && (k & 3) == 3) { a = sb[(it * 3 + k) & 255]; b = sb[(it * 3 + k + 1) & 255]; c = sb[(it * 3 + k + 2) & 255]; }
      }
    }
  } else if (which == 1) {
    unsigned long long A[8];
#pragma unroll
    for (int k = 0; k < 8; k++) A[k] = pk(seed * k, seed * (k + 1));
    const unsigned long long m = pk(0.9999f, 0.9998f), b = pk(seed * 1e-3f, seed * 2e-3f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 12; r++)
#pragma unroll
        for (int k = 0; k < 8; k++) A[k] = ffma2(A[k], m, b);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = __uint_as_float((unsigned)A[k]) + __uint_as_float((unsigned)(A[k] >> 32));
  } else {
    // packed SVM pair: two instances (lo, hi) against one SV: e2 = (na + c.y) ; 9 FFMA2 ; 2 MUFU ; 1 FFMA2 (coef)
    unsigned long long X[9], E[4], ACC[4];
#pragma unroll
    for (int d = 0; d < 9; d++) X[d] = pk(x[d], x[d] * 1.01f);
#pragma unroll
    for (int k = 0; k < 4; k++) { E[k] = pk(e[k], e[k + 4]); ACC[k] = pk(0.f, 0.f); }
    for (int it = 0; it < iters; it++) {
      const float s = seed + it * 1e-9f;
      const unsigned long long a0 = pk(s, s), a1 = pk(s * 0.5f, s * 0.5f), a2 = pk(s * 0.25f, s * 0.25f), cz = pk(0.5f, 0.5f);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        unsigned long long t = E[k];
        t = ffma2(X[0], a0, t); t = ffma2(X[1], a1, t); t = ffma2(X[2], a2, t); t = ffma2(X[3], a0, t); t = ffma2(X[4], a1, t);
        t = ffma2(X[5], a2, t); t = ffma2(X[6], a0, t); t = ffma2(X[7], a1, t); t = ffma2(X[8], a2, t);
        const float lo = ex2f(__uint_as_float((unsigned)t)), hi = ex2f(__uint_as_float((unsigned)(t >> 32)));
        ACC[k] = ffma2(cz, pk(lo, hi), ACC[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) { acc[k] = __uint_as_float((unsigned)ACC[k]); acc[k + 4] = __uint_as_float((unsigned)(ACC[k] >> 32)); }
  }
  float sres = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) sres += acc[k];
  if (sres == 123.456f) out[0] = sres;
}

}  // namespace cds

using namespace cds;

// Returns lane-instructions per second: fp32_pair (FADD+FFMA mix, the contrast pair), ffma (pure FFMA),
// mufu (MUFU.EX2).  Each is the best of `reps` timed launches that fill every SM with 8 CTAs of 256 threads.
extern "C" int cds_measure_peaks_ex(double* vals, int n);
extern "C" int cds_measure_peaks(double* fp32_pair_ips, double* ffma_ips, double* mufu_ips) {
  double v[4];
  if (int e = cds_measure_peaks_ex(v, 4)) return e;
  if (fp32_pair_ips) *fp32_pair_ips = v[0];
  if (ffma_ips) *ffma_ips = v[1];
  if (mufu_ips) *mufu_ips = v[2];
  return CDS_OK;
}

// vals[0..n) = {FADD -> FFMA pair mix, pure FFMA, MUFU.EX2, FADD + FFMA without dependence} in lane-instructions per second (n <= 4)
extern "C" int cds_measure_peaks_ex(double* vals, int n) {
  if (int e = ensure_device()) return e;
  if (!vals || n < 1 || n > 4) { set_error("cds_measure_peaks_ex: bad arguments"); return CDS_EINVAL; }
  int dev = 0, sms = 0;
  CDS_CUDA(cudaGetDevice(&dev));
  CDS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* d = nullptr;
  CDS_CUDA(cudaMalloc(&d, 256));
  cudaEvent_t e0, e1;
  CDS_CUDA(cudaEventCreate(&e0)); CDS_CUDA(cudaEventCreate(&e1));
  const int grid = sms * 8, iters = 2000, reps = 5;
  double best[4] = {0, 0, 0, 0};
  for (int which = 0; which < n; which++) {
    for (int r = 0; r < reps + 1; r++) {
      CDS_CUDA(cudaEventRecord(e0));
      if (which == 0) fp32_pair_peak_kernel<<<grid, 256>>>(d, 0.5f, iters, 0.999f);
      else if (which == 1) ffma_peak_kernel<<<grid, 256>>>(d, 0.5f, iters);
      else if (which == 2) mufu_peak_kernel<<<grid, 256>>>(d, 0.5f, iters);
      else fp32_indep_mix_peak_kernel<<<grid, 256>>>(d, 0.5f, iters);
      CDS_LAUNCH_CHECK();
      CDS_CUDA(cudaEventRecord(e1));
      CDS_CUDA(cudaEventSynchronize(e1));
      float ms = 0; CDS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r == 0) continue;   // warm-up
      // pair kernel: 16 x (FADD + FFMA) + the one FADD that advances p, per inner round
      const double per_thread = which == 0 ? (double)iters * 8 * (16 * 2 + 1) : which == 1 ? (double)iters * 16 * 16
                              : which == 2 ? (double)iters * 16 * 8 : (double)iters * 16 * 8 * 2;
      const double ips = per_thread * grid * 256.0 / (ms * 1e-3);
      if (ips > best[which]) best[which] = ips;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  for (int i = 0; i < n; i++) vals[i] = best[i];
  return CDS_OK;
}

// Probe: returns FP32 FMA lane-operations per second (an FFMA2 counts as 2) and MUFU per second for mix `which`
// (see mix_probe_kernel).  ctas_per_sm CTAs of 128 threads per SM.
// Asynchronous spinner for co-residency experiments (tools/corun_probe.py): which = 0 MUFU.EX2 only, 1 FFMA only, 2 FFMA2 only,
// 3 MUFU.EX2 + one FADD each (the SVM epilogue's instruction mix), 4 the same with one packed FADD2 per two MUFU; `ctas_per_sm` CTAs of `threads` threads on every SM, on `stream`.
__global__ void mufu_fadd_spin_kernel(float* out, float seed, int iters) {
  float a[8], acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = -seed * (k + 1) - threadIdx.x * 1e-6f; acc[k] = 0.f; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[k])); acc[k] += y; a[k] = -y; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k] + acc[k];
  if (s == 123.456f) out[0] = s;
}
__global__ void mufu_fadd2_spin_kernel(float* out, float seed, int iters) {
  float a[8]; unsigned long long acc2[4];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = -seed * (k + 1) - threadIdx.x * 1e-6f;
#pragma unroll
  for (int k = 0; k < 4; k++) acc2[k] = 0ull;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
      float y[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[k]) : "f"(a[k])); a[k] = -y[k]; }
#pragma unroll
      for (int k = 0; k < 4; k++) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc2[k]) : "l"(pk(y[2 * k], y[2 * k + 1])));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) s += __uint_as_float((unsigned)acc2[k]) + __uint_as_float((unsigned)(acc2[k] >> 32)) + a[2 * k] + a[2 * k + 1];
  if (s == 123.456f) out[0] = s;
}
extern "C" int cds_probe_spin(int which, int ctas_per_sm, int threads, int iters, void* stream) {
  if (int e = ensure_device()) return e;
  if (threads < 32 || threads > 256 || ctas_per_sm < 1) { set_error("cds_probe_spin: bad arguments"); return CDS_EINVAL; }
  static float* d = nullptr;
  if (!d) CDS_CUDA(cudaMalloc(&d, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = device_sm_count() * ctas_per_sm;
  // the shared-memory carve-out of an SM cannot change while a CTA is resident: ask for the one the contrast kernel needs
  static bool carve = false;
  if (!carve) {
    carve = true;
    CDS_CUDA(cudaFuncSetAttribute(mufu_peak_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CDS_CUDA(cudaFuncSetAttribute(ffma_peak_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CDS_CUDA(cudaFuncSetAttribute(mix_probe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CDS_CUDA(cudaFuncSetAttribute(mufu_fadd_spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CDS_CUDA(cudaFuncSetAttribute(mufu_fadd2_spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  if (which == 0) mufu_peak_kernel<<<grid, threads, 0, st>>>(d, 0.5f, iters);
  else if (which == 1) ffma_peak_kernel<<<grid, threads, 0, st>>>(d, 0.5f, iters);
  else if (which == 2) mix_probe_kernel<<<grid, 128, 0, st>>>(d, 0.5f, iters, 1);
  else if (which == 4) mufu_fadd2_spin_kernel<<<grid, threads, 0, st>>>(d, 0.5f, iters);
  else mufu_fadd_spin_kernel<<<grid, threads, 0, st>>>(d, 0.5f, iters);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

extern "C" int cds_probe_mix(int which, int ctas_per_sm, double* fma_lane_ops_per_s, double* mufu_per_s) {
  if (int e = ensure_device()) return e;
  int dev = 0, sms = 0;
  CDS_CUDA(cudaGetDevice(&dev));
  CDS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* d = nullptr;
  CDS_CUDA(cudaMalloc(&d, 256));
  cudaEvent_t e0, e1;
  CDS_CUDA(cudaEventCreate(&e0)); CDS_CUDA(cudaEventCreate(&e1));
  const int grid = sms * ctas_per_sm, iters = 4000;
  double best = 1e30;
  for (int r = 0; r < 4; r++) {
    CDS_CUDA(cudaEventRecord(e0));
    mix_probe_kernel<<<grid, 128>>>(d, 0.5f, iters, which);
    CDS_LAUNCH_CHECK();
    CDS_CUDA(cudaEventRecord(e1));
    CDS_CUDA(cudaEventSynchronize(e1));
    float ms = 0; CDS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  const double thr = (double)grid * 128;
  double fma = 0, mufu = 0;
  if (which == 0 || which == 3) { fma = (double)iters * 8 * 10; mufu = (double)iters * 8; }     // + 1 FADD per pair not counted
  else if (which == 1) { fma = (double)iters * 12 * 8 * 2; }
  else { fma = (double)iters * 4 * 10 * 2; mufu = (double)iters * 8; }
  if (fma_lane_ops_per_s) *fma_lane_ops_per_s = fma * thr / (best * 1e-3);
  if (mufu_per_s) *mufu_per_s = mufu * thr / (best * 1e-3);
  return CDS_OK;
}

// ---- UMMA layout probe (tools/umma_probe.py): one tcgen05.mma.kind::tf32 128 x 256 x 8 with caller-supplied shared-memory
// images and descriptor strides, so operand layouts can be checked against numpy without recompiling ----
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const float* __restrict__ a_img, int a_bytes, const float* __restrict__ b_img, int b_bytes, unsigned a_lbo, unsigned a_sbo,
                  unsigned b_lbo, unsigned b_sbo, unsigned idesc, float* __restrict__ D) {
  using namespace cds::tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem; uint8_t* sB = smem + 16384;
  for (int i = threadIdx.x; i < a_bytes / 4; i += 128) reinterpret_cast<float*>(sA)[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes / 4; i += 128) reinterpret_cast<float*>(sB)[i] = b_img[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x == 0) {
    mma_tf32(tb, smem_desc(smem_u32(sA), a_lbo, a_sbo), smem_desc(smem_u32(sB), b_lbo, b_sbo), idesc, 0u);
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < 256; c += 32) {
    uint32_t v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++) D[(size_t)(warp * 32 + lane) * 256 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory"); }
}

// a_img / b_img: host float arrays (<= 16 KB each) copied verbatim into shared memory; D: host [128][256].
extern "C" int cds_probe_umma(const float* a_img, int a_bytes, const float* b_img, int b_bytes, unsigned a_lbo, unsigned a_sbo,
                              unsigned b_lbo, unsigned b_sbo, int a_mn_major, int b_mn_major, float* D) {
  if (int e = ensure_device()) return e;
  if (a_bytes > 16384 || b_bytes > 16384) { set_error("cds_probe_umma: images are limited to 16 KB"); return CDS_EINVAL; }
  float *da = nullptr, *db = nullptr, *dd = nullptr;
  CDS_CUDA(cudaMalloc(&da, 16384)); CDS_CUDA(cudaMalloc(&db, 16384)); CDS_CUDA(cudaMalloc(&dd, 128 * 256 * 4));
  CDS_CUDA(cudaMemcpy(da, a_img, a_bytes, cudaMemcpyHostToDevice)); CDS_CUDA(cudaMemcpy(db, b_img, b_bytes, cudaMemcpyHostToDevice));
  umma_probe_kernel<<<1, 128, 32768>>>(da, a_bytes, db, b_bytes, a_lbo, a_sbo, b_lbo, b_sbo, cds::tc::idesc_tf32(128, 256, a_mn_major, b_mn_major), dd);
  CDS_LAUNCH_CHECK();
  CDS_CUDA(cudaMemcpy(D, dd, 128 * 256 * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return CDS_OK;
}
