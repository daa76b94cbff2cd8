// maps.cu — stage (e): normalisation, smoothing, temporal difference, replicate+Gaussian, resize.
//
// Replaces main.m:144-257 and 275-305 with upsample.m, fourX2.m, pm_norm.m, mysterythrehold.m,
// immove.m, getdistMatrix.m and MATLAB's imfilter/fspecial/imresize as used there.  The MATLAB code
// materialises nine H x W double maps per frame; here
//   * fourX2 + imfilter(gaussian) is evaluated as a polyphase filter on the h/4 x w/4 map (each
//     full-res sample is a <=k/4+1 tap combination of cells), horizontal pass to an (h/4 x W)
//     intermediate, vertical pass straight into min/max/sum reductions (pm_norm needs only those);
//   * imresize(pm_norm(Y),[200 200]) is affine in imresize(Y) (weights sum to 1), and
//     imresize(imfilter(fourX2(X))) is one composite banded operator applied to X, so the six
//     non-temporal features never exist at full resolution;
//   * the three temporal features need a clamp at full resolution (mysterythrehold, main.m:255-257)
//     and are the only maps stored at H x W;
//   * every global reduction (pm_norm min/max, mysterythrehold sums) is block partials + a
//     deterministic finalize, so results do not depend on scheduling.
// All of these passes are bandwidth/latency class (HBM / L2), not FLOP class.
#include "maps.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <map>
#include <mutex>

namespace cds {

// =================================================================================================
// host-side operator construction (double precision, once per context)
// =================================================================================================
static double matlab_round(double x) { return x < 0 ? -floor(-x + 0.5) : floor(x + 0.5); }

// fspecial('gaussian', k, sigma) = exp(-(x^2 + y^2) / (2 sigma^2)) with entries below eps * max zeroed, then normalised.  At the path's
// k = round(h / 7.5), sigma = round(h / 30) the corner entry is exp(-4) of the centre, so the truncation never triggers and the kernel is
// the outer product of this normalised 1-D factor (tests/test_matlab_semantics_cpu.py checks both statements for every BASELINE height).
static std::vector<double> gauss1d(int k, double sigma) {
  std::vector<double> g(k);
  const double siz = (k - 1) / 2.0; double s = 0;
  for (int i = 0; i < k; i++) { const double x = i - siz; g[i] = exp(-(x * x) / (2 * sigma * sigma)); s += g[i]; }
  for (int i = 0; i < k; i++) g[i] /= s;
  return g;
}

void build_gauss_polyphase(int k, double sigma, Polyphase* out) {
  const std::vector<double> g = gauss1d(k, sigma);
  const int c0 = (k + 1) / 2 - 1;                         // imfilter origin: floor((k+1)/2), 1-based
  auto fdiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
  const int dmin = fdiv(0 - c0, 4), dmax = fdiv(3 + k - 1 - c0, 4);
  out->dmin = dmin; out->nt = dmax - dmin + 1;
  std::vector<double> acc((size_t)out->nt * 4, 0.0);
  for (int ph = 0; ph < 4; ph++)
    for (int u = 0; u < k; u++) { const int d = fdiv(ph + u - c0, 4); acc[(size_t)(d - dmin) * 4 + ph] += g[u]; }
  out->g.resize(acc.size());
  for (size_t i = 0; i < acc.size(); i++) out->g[i] = (float)acc[i];
}

std::vector<double> dense_gauss_rep4(int n_cells, int k, double sigma) {
  const std::vector<double> g = gauss1d(k, sigma);
  const int c0 = (k + 1) / 2 - 1, n = 4 * n_cells;
  std::vector<double> A((size_t)n * n_cells, 0.0);
  for (int y = 0; y < n; y++)
    for (int u = 0; u < k; u++) { const int s = y + u - c0; if (s >= 0 && s < n) A[(size_t)y * n_cells + s / 4] += g[u]; }
  return A;
}

std::vector<double> dense_gauss(int n, int k, double sigma) {
  const std::vector<double> g = gauss1d(k, sigma);
  const int c0 = (k + 1) / 2 - 1;
  std::vector<double> A((size_t)n * n, 0.0);
  for (int y = 0; y < n; y++)
    for (int u = 0; u < k; u++) { const int s = y + u - c0; if (s >= 0 && s < n) A[(size_t)y * n + s] += g[u]; }
  return A;
}

// MATLAB imresize 'bilinear' contributions(): triangle kernel, widened by 1/scale when shrinking,
// weights normalised to 1, symmetric boundary indices aux = [1:n, n:-1:1].
std::vector<double> dense_imresize(int n_in, int n_out) {
  const double scale = (double)n_out / n_in;
  double kw = 2.0; const bool aa = scale < 1;
  if (aa) kw /= scale;
  const int P = (int)ceil(kw) + 2;
  std::vector<double> A((size_t)n_out * n_in, 0.0);
  std::vector<double> wt(P);
  for (int x = 1; x <= n_out; x++) {
    const double u = x / scale + 0.5 * (1 - 1 / scale);
    const int left = (int)floor(u - kw / 2);
    double s = 0;
    for (int p = 0; p < P; p++) {
      double t = u - (left + p);
      if (aa) t *= scale;
      double v = (t >= -1 && t < 0) ? t + 1 : (t >= 0 && t <= 1) ? 1 - t : 0;
      if (aa) v *= scale;
      wt[p] = v; s += v;
    }
    for (int p = 0; p < P; p++) {
      const int mm = 2 * n_in;
      const int q = (((left + p - 1) % mm) + mm) % mm;
      const int idx = q < n_in ? q : mm - 1 - q;
      A[(size_t)(x - 1) * n_in + idx] += wt[p] / s;
    }
  }
  return A;
}

std::vector<double> dense_mul(const std::vector<double>& A, int ar, int ac, const std::vector<double>& B, int bc) {
  std::vector<double> C((size_t)ar * bc, 0.0);
  for (int i = 0; i < ar; i++)
    for (int k = 0; k < ac; k++) {
      const double a = A[(size_t)i * ac + k];
      if (a == 0) continue;
      const double* b = &B[(size_t)k * bc]; double* c = &C[(size_t)i * bc];
      for (int j = 0; j < bc; j++) c[j] += a * b[j];
    }
  return C;
}

void banded_from_dense(const std::vector<double>& A, int n_out, int n_in, Banded* out) {
  out->n_out = n_out; out->n_in = n_in; out->start.assign(n_out, 0);
  int nt = 1;
  std::vector<int> lo(n_out, 0), hi(n_out, 0);
  for (int i = 0; i < n_out; i++) {
    int l = -1, h = -1;
    for (int j = 0; j < n_in; j++) if (A[(size_t)i * n_in + j] != 0) { if (l < 0) l = j; h = j; }
    if (l < 0) { l = 0; h = 0; }
    lo[i] = l; hi[i] = h; if (h - l + 1 > nt) nt = h - l + 1;
  }
  out->nt = nt; out->w.assign((size_t)n_out * nt, 0.f);
  for (int i = 0; i < n_out; i++) {
    out->start[i] = lo[i];
    for (int j = lo[i]; j <= hi[i]; j++) out->w[(size_t)i * nt + (j - lo[i])] = (float)A[(size_t)i * n_in + j];
  }
  out->span4 = 0;
  for (int i = 0; i < n_out; i++) {
    if (i + 1 < n_out && out->start[i + 1] < out->start[i]) { out->span4 = -1; break; }
    out->span4 = std::max(out->span4, out->start[std::min(i + 3, n_out - 1)] - out->start[i]);
  }

}

int Banded::upload() {
  CDS_CUDA(cudaMalloc(&d_w, w.size() * 4));
  CDS_CUDA(cudaMalloc(&d_start, start.size() * 4));
  CDS_CUDA(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CDS_CUDA(cudaMemcpy(d_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> wT(w.size());
  for (int i = 0; i < n_out; i++) for (int t = 0; t < nt; t++) wT[(size_t)t * n_out + i] = w[(size_t)i * nt + t];
  CDS_CUDA(cudaMalloc(&d_wT, wT.size() * 4));
  CDS_CUDA(cudaMemcpy(d_wT, wT.data(), wT.size() * 4, cudaMemcpyHostToDevice));
  return CDS_OK;
}
void Banded::release() {
  if (d_w) cudaFree(d_w); if (d_start) cudaFree(d_start); if (d_wT) cudaFree(d_wT);
  d_w = nullptr; d_start = nullptr; d_wT = nullptr;
}
int Polyphase::upload() {
  CDS_CUDA(cudaMalloc(&d_g, g.size() * 4));
  CDS_CUDA(cudaMemcpy(d_g, g.data(), g.size() * 4, cudaMemcpyHostToDevice));
  return CDS_OK;
}
void Polyphase::release() { if (d_g) cudaFree(d_g); if (d_aimg) cudaFree(d_aimg); d_g = nullptr; d_aimg = nullptr; }

// =================================================================================================
// block reductions (deterministic: fixed tree)
// =================================================================================================
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T* red, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  T r = red[0];
  for (int i = 1; i < nw; i++) r = op(r, red[i]);
  return r;
}
struct OpMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinD { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct OpMaxD { __device__ double operator()(double a, double b) const { return fmax(a, b); } };
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpAddI { __device__ int operator()(int a, int b) const { return a + b; } };

// =================================================================================================
// smoothing: causal mean over the last PS pictures (main.m:144-165)
// =================================================================================================
// A thread owns one cell for a group of SG consecutive pictures: their windows are almost the same ring pictures, so each ring value
// is loaded once and added into every picture of the group whose window holds it (uniform predicates).  Every picture still adds its
// own values in ascending picture order -- the reference's loop order; the two FP64 sums must match the oracle to the bit -- while the
// loads drop from SG x PS to PS + SG - 1 per cell and map (one picture per thread read 3.5 GB per batch out of the L2 at 1080p / 50 fps).
// Measured per 64 pictures: one picture per thread 0.27 ms; groups of 8 (103 registers, a quarter of the warps resident: latency
// bound) 0.30 ms; groups of 4 (64 registers) 0.17 ms.
constexpr int SG = 4;
__global__ void __launch_bounds__(128)
smooth_kernel(int n4, int PS, int RL, int first_poc, const double* __restrict__ Vx, const double* __restrict__ Vy,
              const float* __restrict__ mvn, const float* __restrict__ bitn, const float* __restrict__ depn,
              float4* __restrict__ sm4 /*ring [RL][n4]: {mvx, mvy, bit, depth} smoothed*/,
              float* __restrict__ mv_s, double* __restrict__ cxy, int B) {
  const int o = blockIdx.x * 128 + threadIdx.x, f0 = blockIdx.y * SG;
  if (o >= n4) return;
  const int ng = min(SG, B - f0), poc0 = first_poc + f0;
  float a[SG], d[SG], e[SG];
  double b[SG], c[SG];                                             // mv_x / mv_y stay double: see camera_motion_kernel<double>
#pragma unroll
  for (int g = 0; g < SG; g++) { a[g] = 0.f; d[g] = 0.f; e[g] = 0.f; b[g] = 0.0; c[g] = 0.0; }
  const int q_lo = max(0, poc0 - PS + 1), q_hi = poc0 + ng - 1;
#pragma unroll 6
  for (int q = q_lo; q <= q_hi; q++) {                            // ascending, like the reference loop
    const size_t s = (size_t)(q % RL) * n4 + o;
    const float va = __ldg(mvn + s), vd = __ldg(bitn + s), ve = __ldg(depn + s);
    const double vb = __ldg(Vx + s), vc = __ldg(Vy + s);
#pragma unroll
    for (int g = 0; g < SG; g++) {
      const int poc = poc0 + g;                                    // window of picture g: [poc - min(PS, poc + 1) + 1, poc]
      if (g < ng && q <= poc && q > poc - PS) { a[g] += va; b[g] += vb; c[g] += vc; d[g] += vd; e[g] += ve; }
    }
  }
#pragma unroll
  for (int g = 0; g < SG; g++) {
    if (g < ng) {
      const int poc = poc0 + g, f = f0 + g;
      const int cnt = min(PS, poc + 1);
      const float fc = (float)cnt;
      const double bm = b[g] / (double)cnt, cm = c[g] / (double)cnt;
      const size_t so = (size_t)(poc % RL) * n4 + o;
      mv_s[(size_t)f * n4 + o] = a[g] / fc;
      sm4[so] = make_float4((float)bm, (float)cm, d[g] / fc, e[g] / fc);
      if (cxy) { cxy[(size_t)f * n4 + o] = bm; cxy[((size_t)B + f) * n4 + o] = cm; }     // [mvx: B maps | mvy: B maps]
    }
  }
}

int smooth_launch(int n4, int PS, int ring_len, int first_poc, int B, const double* Vx, const double* Vy, const float* mvn,
                  const float* bitn, const float* depn, float4* sm4, float* mv_s,
                  double* cxy, cudaStream_t st) {
  dim3 grid(ceil_div(n4, 128), ceil_div(B, SG));
  smooth_kernel<<<grid, 128, 0, st>>>(n4, PS, ring_len, first_poc, Vx, Vy, mvn, bitn, depn, sm4, mv_s, cxy, B);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// =================================================================================================
// mysterythrehold.m:3-7 on the smoothed h4 x w4 maps (main.m:170-172).  One CTA per (picture, map).
// =================================================================================================
// One thread-block CLUSTER of MS_CLUSTER CTAs per (picture, map): the three sweeps over a map are latency bound, and one CTA per map
// (192 CTAs for 64 pictures: barely one per SM, 0.14 ms) left most of the GPU idle.  A CTA owns a contiguous slice (its re-reads hit
// the L1); the partial statistics of the slices meet in the rank-0 CTA's shared memory, one slot per rank, and every CTA adds them in
// rank order -- no atomics, so the result does not depend on timing.
constexpr int MS_CLUSTER = 8, MS_THREADS = 512;
struct MsShared { double s[2][MS_CLUSTER], sb[2][MS_CLUSTER]; float mx[2][MS_CLUSTER]; int nb[2][MS_CLUSTER]; };
// blockIdx.y = 0: the smoothed mv_norm map (planar); 1: the bit AND depth maps together -- they are components z / w of the
// interleaved ring, so one 16-byte load per cell serves both (one map per launch row read 4 of every 16 bytes, three times).
__global__ void __cluster_dims__(MS_CLUSTER, 1, 1) __launch_bounds__(MS_THREADS)
myst_small_kernel(int n4, double inv_hw, int RL, int first_poc, const float* __restrict__ mv_s,
                  const float4* __restrict__ sm4, float* __restrict__ Xmaps) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ MsShared sh;
  __shared__ float redf[32];
  __shared__ double redd[32];
  __shared__ int redi[32];
  MsShared* S0 = cluster.map_shared_rank(&sh, 0);
  const int crank = (int)cluster.block_rank();
  const int f = blockIdx.x / MS_CLUSTER;
  const bool two = blockIdx.y == 1;                                // uniform over the cluster
  const int nm = two ? 2 : 1;
  const float* __restrict__ p1 = mv_s + (size_t)f * n4;
  const float4* __restrict__ p4 = sm4 + (size_t)((first_poc + f) % RL) * n4;
  float* __restrict__ dst0 = Xmaps + ((size_t)f * 9 + (two ? 3 : 0)) * n4;      // maps 0 | 3 and 6
  float* __restrict__ dst1 = Xmaps + ((size_t)f * 9 + 6) * n4;
  const int per = (n4 + MS_CLUSTER - 1) / MS_CLUSTER, o0 = crank * per, o1 = min(n4, o0 + per);
  auto load2 = [&](int o, float& v0, float& v1) { if (two) { const float4 q = p4[o]; v0 = q.z; v1 = q.w; } else { v0 = p1[o]; v1 = 0.f; } };
  double s[2] = {0, 0}; float mx[2] = {-INFINITY, -INFINITY};
  for (int o = o0 + threadIdx.x; o < o1; o += MS_THREADS) {
    float v0, v1; load2(o, v0, v1);
    s[0] += v0; mx[0] = fmaxf(mx[0], v0); s[1] += v1; mx[1] = fmaxf(mx[1], v1);
  }
  for (int k = 0; k < nm; k++) {
    s[k] = block_reduce(s[k], redd, OpAddD());
    mx[k] = block_reduce(mx[k], redf, OpMax());
    if (threadIdx.x == 0) { S0->s[k][crank] = s[k]; S0->mx[k][crank] = mx[k]; }
  }
  cluster.sync();
  float lim[2] = {0.f, 0.f};
  for (int k = 0; k < nm; k++) {
    double t = 0; float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MS_CLUSTER; c++) { t += S0->s[k][c]; m = fmaxf(m, S0->mx[k][c]); }
    mx[k] = m;
    lim[k] = m - (float)(t * inv_hw);                              // max - ave, ave over im_height*im_width (:3-4)
  }
  double sb[2] = {0, 0}; int nb[2] = {0, 0};
  for (int o = o0 + threadIdx.x; o < o1; o += MS_THREADS) {
    float v0, v1; load2(o, v0, v1);
    if (v0 > lim[0]) { sb[0] += v0; nb[0]++; }
    if (two && v1 > lim[1]) { sb[1] += v1; nb[1]++; }
  }
  for (int k = 0; k < nm; k++) {
    sb[k] = block_reduce(sb[k], redd, OpAddD());
    nb[k] = block_reduce(nb[k], redi, OpAddI());
    if (threadIdx.x == 0) { S0->sb[k][crank] = sb[k]; S0->nb[k][crank] = nb[k]; }
  }
  cluster.sync();
  float thresh[2] = {NAN, NAN}, inv[2] = {0.f, 0.f};
  for (int k = 0; k < nm; k++) {
    double t = 0; int n = 0;
#pragma unroll
    for (int c = 0; c < MS_CLUSTER; c++) { t += S0->sb[k][c]; n += S0->nb[k][c]; }
    thresh[k] = n > 0 ? (float)(t / n) : NAN;                      // :5
    const float M = n > 0 ? fminf(thresh[k], mx[k]) : mx[k];       // max after the clamp
    inv[k] = 1.f / (M + 0.000001f);                                // :7
  }
  for (int o = o0 + threadIdx.x; o < o1; o += MS_THREADS) {
    float v0, v1; load2(o, v0, v1);
    if (v0 > thresh[0]) v0 = thresh[0];
    dst0[o] = v0 * inv[0];
    if (two) { if (v1 > thresh[1]) v1 = thresh[1]; dst1[o] = v1 * inv[1]; }
  }
  cluster.sync();                  // rank 0's shared memory must outlive every remote read
}

int myst_small_launch(int n4, int H, int W, int ring_len, int first_poc, int B, const float* mv_s, const float4* sm4,
                      float* Xmaps, cudaStream_t st) {
  dim3 grid(B * MS_CLUSTER, 2);
  myst_small_kernel<<<grid, MS_THREADS, 0, st>>>(n4, 1.0 / ((double)H * W), ring_len, first_poc, mv_s, sm4, Xmaps);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// host: tiled FP32 tensor map with zero fill outside (cuTensorMapEncodeTiled through the runtime's driver entry point)
static int encode_tensor_map_f32_3d(const void* base, const cuuint64_t dims[3], const cuuint64_t strides[2], const cuuint32_t box[3], CUtensorMap* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CDS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return CDS_ECUDA; }
    encode = (encode_fn)fn;
  }
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu", (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2]); return CDS_ECUDA; }
  return CDS_OK;
}

// =================================================================================================
// temporal difference with accumulated camera motion (main.m:174-199, immove.m)
// grid: x = picture in batch (fast, so the pictures that share ring data run together), y = cell tile
// =================================================================================================
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
constexpr int TMP_MAXJ = 1024;      // PT - 1 <= 7 * round(0.3 * 240) - 1 = 503
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// packed FP32 (sm_100 FADD2 / FFMA2: two IEEE operations per issued instruction; |.| and a scalar broadcast are operand modifiers)
__device__ __forceinline__ unsigned long long tp_pack(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void tp_unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long tp_sub2(unsigned long long a, unsigned long long b) {      // a - b
  float b0, b1; tp_unpack(b, b0, b1);
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(tp_pack(-b0, -b1))); return d;
}
__device__ __forceinline__ unsigned long long tp_fma2_abs(float w, unsigned long long d, unsigned long long acc) {   // acc + w * |d|
  float d0, d1; tp_unpack(d, d0, d1);
  unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(tp_pack(w, w)), "l"(tp_pack(fabsf(d0), fabsf(d1))), "l"(acc)); return r;
}

__global__ void __launch_bounds__(256)
temporal_kernel(int h4, int w4, int PT, int RL, int first_poc, const float4* __restrict__ sm4 /*{mvx, mvy, bit, depth} smoothed*/,
                const int* __restrict__ cam,
                const float* __restrict__ wtab /*[PT][4]: exp(-j/26), exp(-j/46), exp(-j/46)*/, float* __restrict__ Xmaps) {
  // per (picture, j): ring slot offset of picture poc-j and the accumulated camera shift, shared by every cell of the picture
  extern __shared__ int4 s_tab[];                 // [jmax + 1] = {slot offset in cells, sx, sy, 0}
  const int f = blockIdx.x, n4 = h4 * w4;
  const int poc = first_poc + f;
  const int jmax = min(PT - 1, poc);
  if (threadIdx.x < 32) {                         // warp 0: inclusive scan of x_cammov / y_cammov over j (main.m:189-190)
    const int lane = threadIdx.x;
    int cx = 0, cy = 0;
    for (int j0 = 1; j0 <= jmax; j0 += 32) {
      const int j = j0 + lane;
      int vx = 0, vy = 0;
      if (j <= jmax) { const int cs = ((poc - j + 1) % RL) * 2; vx = cam[cs]; vy = cam[cs + 1]; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tx = __shfl_up_sync(0xffffffffu, vx, o), ty = __shfl_up_sync(0xffffffffu, vy, o);
        if (lane >= o) { vx += tx; vy += ty; }
      }
      const int camX = cx + vx, camY = cy + vy;
      if (j <= jmax) {
        const int sx = camX >= 0 ? (camX + 2) >> 2 : -((-camX + 2) >> 2);   // round(camX/4), half away from zero
        const int sy = camY >= 0 ? (camY + 2) >> 2 : -((-camY + 2) >> 2);
        s_tab[j] = make_int4(((poc - j) % RL) * n4, sx, sy, 0);
      }
      cx += __shfl_sync(0xffffffffu, vx, 31); cy += __shfl_sync(0xffffffffu, vy, 31);
    }
  }
  __syncthreads();
  // two horizontally adjacent cells per thread (w4 is even): the (slot, shift, weight) of a term is fetched once for both
  const int o = (blockIdx.y * 256 + threadIdx.x) * 2;
  if (o >= n4) return;
  const int r = o / w4, c = o - r * w4;
  const size_t cur = (size_t)(poc % RL) * n4 + o;
  const float4 ref0 = sm4[cur], ref1 = sm4[cur + 1];
  float am0 = 0.f, ab0 = 0.f, ad0 = 0.f, am1 = 0.f, ab1 = 0.f, ad1 = 0.f;
  const float4* __restrict__ w4tab = reinterpret_cast<const float4*>(wtab);
#pragma unroll 2
  for (int j = 1; j <= jmax; j++) {
    const int4 t = s_tab[j];
    const int sr = r - max(t.z, 0), sc = c - t.y;                  // immove: +-x moves content right / left, +y down, zero fill; -y: no shift (immove.m:19 quirk)
    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
    if ((unsigned)sr < (unsigned)h4) {
      const float4* __restrict__ row = sm4 + (t.x + sr * w4);
      if ((unsigned)sc < (unsigned)w4) p0 = __ldg(row + sc);       // one LDG.128 for the four maps
      if ((unsigned)(sc + 1) < (unsigned)w4) p1 = __ldg(row + sc + 1);
    }
    const float4 wj = __ldg(w4tab + j);
    const float dx0 = ref0.x - p0.x, dy0 = ref0.y - p0.y, dx1 = ref1.x - p1.x, dy1 = ref1.y - p1.y;
    am0 = fmaf(wj.x, sqrt_approx(fmaf(dx0, dx0, dy0 * dy0)), am0);   // :195
    am1 = fmaf(wj.x, sqrt_approx(fmaf(dx1, dx1, dy1 * dy1)), am1);
    ab0 = fmaf(wj.y, fabsf(ref0.z - p0.z), ab0);                     // :196
    ab1 = fmaf(wj.y, fabsf(ref1.z - p1.z), ab1);
    ad0 = fmaf(wj.z, fabsf(ref0.w - p0.w), ad0);                     // :197
    ad1 = fmaf(wj.z, fabsf(ref1.w - p1.w), ad1);
  }
  float* X = Xmaps + (size_t)f * 9 * n4 + o;
  *reinterpret_cast<float2*>(X + (size_t)1 * n4) = make_float2(am0, am1);
  *reinterpret_cast<float2*>(X + (size_t)4 * n4) = make_float2(ab0, ab1);
  *reinterpret_cast<float2*>(X + (size_t)7 * n4) = make_float2(ad0, ad1);
}

// -------------------------------------------------------------------------------------------------
// Grouped variant (the default).  The kernel above streams, for every output picture, all PT-1 ring pictures from L2:
// B * (PT-1) * n4 * 16 B per batch (13.8 GB at 1080p / 50 fps / B = 64, ~11 TB/s: the L2 is the wall).  Consecutive output
// pictures read almost the same ring pictures, so here one CTA owns a 16 x 32 cell tile for a GROUP of TG consecutive
// pictures: it walks the ring pictures q = poc_last-1 ... poc_first-(PT-1) once, stages the part of picture q that the
// group's (shifted) reads touch in shared memory — one TMA box per ring picture into an 8-deep mbarrier ring, issued by a single
// thread, zero fill outside the picture = immove's zero padding; no block-wide barrier in the loop — and accumulates it into the
// TG pictures' registers.  L2 traffic drops by ~TG / halo
// overhead; what remains is one LDS.128 per (cell, picture, q) term, i.e. the shared-memory bandwidth.
// Every picture still adds its terms in ascending j with the same fmaf sequence, so the result is bit-identical to the
// kernel above and independent of where the picture sits in a batch (sharded == unsharded bytewise).
// The accumulated camera shift differs per (picture, q); the staged window covers the spread of the group's shifts up to
// TW_R x TW_C cells, pairs whose shifted tile falls outside it (fast pans, jitter) read from global memory instead.
// (The benchmark's synthetic clip jumps by up to 6 cells in a new direction every picture: 60 % of its pairs are inside the window
// and 14 % of the ring pictures run on the straight-line path.  A second, larger box chosen per ring picture -- spread 12 x 24,
// 98 % / 81 % -- was measured and is not kept: 0.745 against 0.720 ms; either way ~5.5 GB per launch cross from the L2 to the SMs,
// which is what bounds the kernel on that clip.  A static camera needs 2.9 GB.)
// -------------------------------------------------------------------------------------------------
// exp(-j / 26) (mv) and exp(-j / 46) (bit, depth), main.m:120-122,195-197: the same for every clip, so the grouped kernel reads them
// from constant memory through the uniform datapath (j is warp-uniform) instead of keeping a rotating window in vector registers
__constant__ float2 c_temporal_w[TMP_MAXJ];
constexpr int TG = 8;                          // pictures per group (the offset table below is read as two int4)
constexpr int TGT = 8;                         // pictures per thread (TGT = 4 = two thread sets, 992 threads at 64 registers: 0.95 ms against 0.89)
constexpr int TT_R = 15, TT_C = 32;            // tile: one warp per row of 32 cells; 15 consumer warps + the producer warp = 512 threads (128 registers each)
constexpr int TW_R = TT_R + 4, TW_C = TT_C + 8;    // staged window = the TMA box: tile + the spread of the group's shifts it can absorb
constexpr int T_STAGES = 8;                        // TMA ring: the windows of the next 8 ring pictures in flight (most come from HBM: the ring exceeds the L2)
constexpr int T_STAGE_BYTES = TW_R * TW_C * 16;
constexpr int T_SETS = TG / TGT, T_CONSUMERS = T_SETS * TT_R * TT_C, T_WARPS = T_CONSUMERS / 32;
constexpr int T_THREADS = T_CONSUMERS + 32;    // + one producer warp: a single thread that only keeps the TMA ring full, decoupled from the consumers' arithmetic
static_assert(TG == 8 && T_STAGES == TG && T_STAGES % TGT == 0 && T_STAGE_BYTES % 128 == 0 && T_THREADS <= 1024,
              "temporal_group_kernel: offset table read as two int4; static stage index; weight window rotation; TMA destination alignment");

__global__ void __launch_bounds__(T_THREADS, 1)
temporal_group_kernel(const __grid_constant__ CUtensorMap tmap /*ring as [RL][h4][w4 * 4] FP32, box [1][TW_R][TW_C * 4], zero fill outside*/,
                      int h4, int w4, int PT, int RL, int first_poc, int n, const float4* __restrict__ sm4,
                      const int* __restrict__ cam, const float* __restrict__ wtab, float* __restrict__ Xmaps) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char t_smem[];
  __shared__ __align__(8) uint64_t t_bars[2 * T_STAGES];                               // full[s] (TMA landed), empty[s] (all warps done with it)
  float4* stage = reinterpret_cast<float4*>(t_smem);                                   // [T_STAGES][TW_R * TW_C]
  float4* s_w = stage + T_STAGES * TW_R * TW_C;                                        // [PT] weights
  int2* s_shift = reinterpret_cast<int2*>(s_w + PT);                                   // [TG][PT] {sx, max(sy, 0)} per (picture, j)
  int4* s_win = reinterpret_cast<int4*>(s_shift + TG * PT);                            // [nq_max] {sxmax, symax, cols, rows} per ring picture
  int* s_off = reinterpret_cast<int*>(s_win + (PT - 1 + TG));                          // [nq_max][TG] offset of the pair's source tile in the window; -1 read from global; -2 no term
  const int n4 = h4 * w4;
  const int ntc = (w4 + TT_C - 1) / TT_C;
  // grid: x = picture group (fast: the CTAs of one tile run together and find each other's ring windows in the L2), y = tile
  const int r0 = (blockIdx.y / ntc) * TT_R, c0 = (blockIdx.y % ntc) * TT_C;
  const int f0 = blockIdx.x * TG, ng = min(TG, n - f0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int set = warp / TT_R, wrow = warp % TT_R;                 // thread set (pictures set*TGT ..), row of the tile
  const int poc0 = first_poc + f0;
  const int q_hi = poc0 + ng - 2;                                  // newest ring picture any picture of the group reads
  const int q_lo = max(0, poc0 - (PT - 1));
  const int nq = q_hi - q_lo + 1;                                  // may be <= 0 (first picture of a clip on its own)

  const uint32_t bar0 = smem_u32(t_bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < T_STAGES; s++) { mbar_init(bar0 + 8u * s, 1); mbar_init(bar0 + 8u * (T_STAGES + s), T_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp < ng) {                                 // warp g: inclusive scan of x_cammov / y_cammov over j for picture g (main.m:189-190)
    const int poc = poc0 + warp, jmax = min(PT - 1, poc);
    int cx = 0, cy = 0;
    for (int j0 = 1; j0 <= jmax; j0 += 32) {
      const int j = j0 + lane;
      int vx = 0, vy = 0;
      if (j <= jmax) { const int cs = ((poc - j + 1) % RL) * 2; vx = cam[cs]; vy = cam[cs + 1]; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tx = __shfl_up_sync(0xffffffffu, vx, o), ty = __shfl_up_sync(0xffffffffu, vy, o);
        if (lane >= o) { vx += tx; vy += ty; }
      }
      const int camX = cx + vx, camY = cy + vy;
      if (j <= jmax) {
        const int sx = camX >= 0 ? (camX + 2) >> 2 : -((-camX + 2) >> 2);   // round(camX/4), half away from zero
        const int sy = camY >= 0 ? (camY + 2) >> 2 : -((-camY + 2) >> 2);
        s_shift[warp * PT + j] = make_int2(sx, max(sy, 0));                // immove.m:19 quirk: a negative mov_y does not shift
      }
      cx += __shfl_sync(0xffffffffu, vx, 31); cy += __shfl_sync(0xffffffffu, vy, 31);
    }
  }
  __syncthreads();
  for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {        // window of ring picture q: the spread of the group's shifts
    const int q = q_hi - qi;
    int sxmin = 1 << 30, sxmax = -(1 << 30), symin = 1 << 30, symax = -(1 << 30);
    for (int g = 0; g < ng; g++) {
      const int poc = poc0 + g, j = poc - q;
      if (j >= 1 && j <= min(PT - 1, poc)) {
        const int2 sh = s_shift[g * PT + j];
        sxmin = min(sxmin, sh.x); sxmax = max(sxmax, sh.x); symin = min(symin, sh.y); symax = max(symax, sh.y);
      }
    }
    const int4 wn = sxmax >= sxmin ? make_int4(sxmax, symax, min(TT_C + sxmax - sxmin, TW_C), min(TT_R + symax - symin, TW_R)) : make_int4(0, 0, 0, 0);
    s_win[qi] = wn;
    for (int g = 0; g < TG; g++) {
      const int poc = poc0 + g, j = poc - q;
      int off = -2;
      if (g < ng && j >= 1 && j <= min(PT - 1, poc)) {
        const int2 sh = s_shift[g * PT + j];
        const int dx = wn.x - sh.x, dy = wn.y - sh.y;                  // >= 0
        // BYTE offset of the pair's source tile from the start of the ring buffers (ring slot qi % T_STAGES included), so that a
        // consumer adds it to its own cell address with one IADD; the TMA box always has the full capacity
        off = (dx + TT_C <= TW_C && dy + TT_R <= TW_R) ? ((qi % T_STAGES) * (TW_R * TW_C) + dy * TW_C + dx) * 16 : -1;
      }
      s_off[qi * TG + g] = off;
    }
  }
  __syncthreads();

  // One thread (warp 0, lane 0) keeps the ring full: per ring picture one TMA box at (c0 - sxmax, r0 - symax) of its ring slot;
  // cells outside the picture arrive as zeros (immove's zero padding).  The ring slot is tracked incrementally (q runs downwards).
  int st_slot = ((q_hi % RL) + RL) % RL;
  auto stage_q = [&](int qi) {                     // caller: the producer thread
    if (qi < nq) {
      const int s = qi % T_STAGES;
      mbar_wait(bar0 + 8u * (T_STAGES + s), (uint32_t)(((qi / T_STAGES) & 1) ^ 1));   // every warp is done with the window this one replaces
      const int4 wn = s_win[qi];
      mbar_expect_tx(bar0 + 8u * s, (uint32_t)T_STAGE_BYTES);
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(stage + s * (TW_R * TW_C))), "l"(&tmap), "r"((c0 - wn.x) * 4), "r"(r0 - wn.y), "r"(st_slot), "r"(bar0 + 8u * s) : "memory");
      st_slot = st_slot == 0 ? RL - 1 : st_slot - 1;
    }
  };

  const int r = r0 + wrow, c = c0 + lane;
  const bool inside = r < h4 && c < w4;
  const int g0 = T_SETS == 1 ? 0 : set * TGT;      // this thread's first picture of the group (a constant with one thread set: j and the weights stay uniform)
  float4 ref[TGT];
  float am[TGT];
  unsigned long long abd[TGT];                     // (bit, depth) sums as one packed pair: the two share weight and form
#pragma unroll
  for (int k = 0; k < TGT; k++) {
    ref[k] = (inside && g0 + k < ng) ? sm4[(size_t)((poc0 + g0 + k) % RL) * n4 + (size_t)r * w4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    am[k] = 0.f; abd[k] = 0ull;
  }
  if (warp == T_WARPS) {                          // ---- producer warp: one thread walks the ring pictures ahead of the consumers ----
    if (lane == 0) {
#pragma unroll 1
      for (int w = 0; w < nq; w++) stage_q(w);     // waits for the window's buffer to be released by every consumer warp, then issues the box
    }
    return;
  }
  // The loop is bound by issued instructions (15 warps, ~120 per ring picture, of which 56 are arithmetic), so its bookkeeping is
  // kept in loop-carried registers: 32-bit shared addresses made opaque to the compiler (it otherwise rebuilds them from
  // %cluster_ctaid / the window base in every iteration to save registers), the ring slot, its parity and j advance by increments.
  uint32_t mine, offp, barf;                       // this thread's cell in ring buffer 0; the offset-table row; full[0]
  asm volatile("mov.u32 %0, %1;" : "=r"(mine) : "r"(smem_u32(stage + wrow * TW_C + lane)));
  asm volatile("mov.u32 %0, %1;" : "=r"(offp) : "r"(smem_u32(s_off + g0)));
  asm volatile("mov.u32 %0, %1;" : "=r"(barf) : "r"(bar0));
  int q_slot = ((q_hi % RL) + RL) % RL;
  uint32_t u8 = 0, par = 0;                        // 8 * (qi % T_STAGES), (qi / T_STAGES) & 1
  int j0 = poc0 + g0 - q_hi;                       // j of this thread's first picture for ring picture q = q_hi - qi (>= 1 on the straight-line path)
  // (not unrolled over the ring: eight copies of the body do not fit the instruction cache — ncu showed no_inst stalls)
#pragma unroll 1
  for (int qi = 0; qi < nq; qi++, j0++, offp += TG * 4) {
    {
      {
        mbar_wait(barf + u8, par);                                       // window qi has landed
        const int q = q_hi - qi;
        int offs[TGT]; int any_neg = 0;
#pragma unroll
        for (int k4 = 0; k4 < TGT; k4 += 4) {
          int4 o4;
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(o4.x), "=r"(o4.y), "=r"(o4.z), "=r"(o4.w) : "r"(offp + k4 * 4));
          offs[k4] = o4.x; offs[k4 + 1] = o4.y; offs[k4 + 2] = o4.z; offs[k4 + 3] = o4.w;
          any_neg |= o4.x | o4.y | o4.z | o4.w;
        }
        if (any_neg >= 0) {
          // every picture of this thread has a term for q and every source tile lies in the staged window: straight-line code,
          // all shared-memory loads first so that their latencies overlap
          float4 pv[TGT];
#pragma unroll
          for (int k = 0; k < TGT; k++)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pv[k].x), "=f"(pv[k].y), "=f"(pv[k].z), "=f"(pv[k].w) : "r"(mine + (uint32_t)offs[k]));
#pragma unroll
          for (int k = 0; k < TGT; k++) {
            const float4 p = pv[k];
            const float2 wj = c_temporal_w[j0 + k];                              // uniform: constant bank
            float ddx, ddy;                                                      // the four differences as two FADD2
            tp_unpack(tp_sub2(tp_pack(ref[k].x, ref[k].y), tp_pack(p.x, p.y)), ddx, ddy);
            am[k] = fmaf(wj.x, sqrt_approx(fmaf(ddx, ddx, ddy * ddy)), am[k]);   // main.m:195
            abd[k] = tp_fma2_abs(wj.y, tp_sub2(tp_pack(ref[k].z, ref[k].w), tp_pack(p.z, p.w)), abd[k]);   // :196, :197
          }
        } else {
#pragma unroll
          for (int k = 0; k < TGT; k++) {
            const int off = offs[k];
            if (off != -2) {                                                  // uniform per thread set
              const int j = poc0 + g0 + k - q;
              float4 p;
              if (off >= 0) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w) : "r"(mine + (uint32_t)off));
              else {                                                          // shift spread beyond the window capacity: straight from L2
                const int2 sh = s_shift[(g0 + k) * PT + j];
                const int sr = r - sh.y, sc = c - sh.x;
                p = ((unsigned)sr < (unsigned)h4 && (unsigned)sc < (unsigned)w4) ? __ldg(sm4 + (size_t)q_slot * n4 + (size_t)sr * w4 + sc) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
              const float2 wj = c_temporal_w[j];
              const float ddx = ref[k].x - p.x, ddy = ref[k].y - p.y;
              am[k] = fmaf(wj.x, sqrt_approx(fmaf(ddx, ddx, ddy * ddy)), am[k]);
              abd[k] = tp_fma2_abs(wj.y, tp_sub2(tp_pack(ref[k].z, ref[k].w), tp_pack(p.z, p.w)), abd[k]);
            }
          }
        }
        q_slot = q_slot == 0 ? RL - 1 : q_slot - 1;
        __syncwarp();
        if (lane == 0) mbar_arrive(barf + 8u * T_STAGES + u8);          // this warp has read window qi
        u8 += 8; if (u8 == 8u * T_STAGES) { u8 = 0; par ^= 1u; }
      }
    }
  }
  if (inside) {
#pragma unroll
    for (int k = 0; k < TGT; k++)
      if (g0 + k < ng) {
        float* X = Xmaps + (size_t)(f0 + g0 + k) * 9 * n4 + (size_t)r * w4 + c;
        float ab, ad; tp_unpack(abd[k], ab, ad);
        X[(size_t)1 * n4] = am[k]; X[(size_t)4 * n4] = ab; X[(size_t)7 * n4] = ad;
      }
  }
}

int temporal_launch(int h4, int w4, int PT, int ring_len, int first_poc, int B, const float4* sm4, const int* cam_ring,
                    const float* wtab, float* Xmaps, cudaStream_t st) {
  if (PT > TMP_MAXJ) { set_error("temporal: PT %d exceeds %d", PT, TMP_MAXJ); return CDS_EINVAL; }
  static const int impl = [] { const char* e = getenv("CDS_TEMPORAL_IMPL"); return (e && !strcmp(e, "stream")) ? 0 : 1; }();
  if (impl == 0 && !(w4 & 1)) {
    dim3 grid(B, ceil_div(h4 * w4, 512));
    temporal_kernel<<<grid, 256, (size_t)PT * sizeof(int4), st>>>(h4, w4, PT, ring_len, first_poc, sm4, cam_ring, wtab, Xmaps);
    CDS_LAUNCH_CHECK();
    return CDS_OK;
  }
  {
    static std::mutex mu; static std::map<int, bool> filled;
    int dev = 0; CDS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (!filled[dev]) {
      std::vector<float2> wt(TMP_MAXJ);
      for (int j = 0; j < TMP_MAXJ; j++) wt[j] = make_float2((float)exp(-j / 26.0), (float)exp(-j / 46.0));
      CDS_CUDA(cudaMemcpyToSymbol(c_temporal_w, wt.data(), wt.size() * sizeof(float2)));
      filled[dev] = true;
    }
  }
  const int nq_max = PT - 1 + TG;
  const size_t smem = (size_t)T_STAGES * T_STAGE_BYTES + (size_t)PT * 16 + (size_t)TG * PT * 8 + (size_t)nq_max * 16 + (size_t)nq_max * TG * 4;
  if (int e = ensure_dyn_smem(temporal_group_kernel, smem)) return e;
  CUtensorMap tmap;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)w4 * 4, (cuuint64_t)h4, (cuuint64_t)ring_len};
    const cuuint64_t strides[2] = {(cuuint64_t)w4 * 16, (cuuint64_t)h4 * w4 * 16};
    const cuuint32_t box[3] = {(cuuint32_t)TW_C * 4, (cuuint32_t)TW_R, 1};
    if (int e = encode_tensor_map_f32_3d(sm4, dims, strides, box, &tmap)) return e;
  }
  dim3 grid(ceil_div(B, TG), ceil_div(h4, TT_R) * ceil_div(w4, TT_C));
  if (grid.y > 65535) { set_error("temporal: %u tiles", grid.y); return CDS_EINVAL; }
  temporal_group_kernel<<<grid, T_THREADS, smem, st>>>(tmap, h4, w4, PT, ring_len, first_poc, B, sm4, cam_ring, wtab, Xmaps);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// =================================================================================================
// contrast input preparation (cu_contrast5.m:8-32 block mean; Sudoku_contrastcpp5.m:18-34) and
// output post-processing (Sudoku :50, cu_contrast5.m:36-55/:60, main.m:237-241)
// =================================================================================================
__global__ void __launch_bounds__(256)
contrast_sub_kernel(const float* __restrict__ src, long frame_stride, int RL, int first_slot, int rows, int cols, int cus,
                    int srows, int scols, float* __restrict__ sub, unsigned* __restrict__ keys) {
  const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  const float* __restrict__ S = src + (size_t)(RL > 0 ? (first_slot + f) % RL : f) * frame_stride;
  float v = 0.f; bool ok = i < srows * scols;
  if (ok) {
    const int k = i / scols, j = i - k * scols;
    if (cus == 1) v = S[i];
    else {
      const int r0 = k * cus, c0 = j * cus, r1 = min(rows, r0 + cus), c1 = min(cols, c0 + cus);
      float acc = 0.f;
      for (int c = c0; c < c1; c++) {                               // mean(mean(temp)): column means first
        float cs = 0.f;
        for (int r = r0; r < r1; r++) cs += S[(size_t)r * cols + c];
        acc += cs / (float)(r1 - r0);
      }
      v = acc / (float)(c1 - c0);
    }
    v += 0.00000001f;                                               // Sudoku :18
    sub[(size_t)f * srows * scols + i] = v;
  }
  // min / max over the map (the zero padding joins in contrast_pad_kernel)
  float mn = ok ? v : INFINITY, mx = ok ? v : -INFINITY;
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mn <= mx) { atomicMin(keys + 2 * f, float_key(mn)); atomicMax(keys + 2 * f + 1, float_key(mx)); }
}

int contrast_sub_launch(const float* src, long src_frame_stride, int ring_len, int first_slot, int B, int rows, int cols,
                        int cusize, float* sub, unsigned* minmax_keys, cudaStream_t st) {
  const int srows = ceil_div(rows, cusize), scols = ceil_div(cols, cusize);
  dim3 grid(ceil_div(srows * scols, 256), B);
  contrast_sub_kernel<<<grid, 256, 0, st>>>(src, src_frame_stride, ring_len, first_slot, rows, cols, cusize, srows, scols, sub, minmax_keys);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

__global__ void __launch_bounds__(256)
contrast_pad_kernel(const float* __restrict__ sub, const unsigned* __restrict__ keys, int srows, int scols, int len,
                    float* __restrict__ pad) {
  const int f = blockIdx.y, pr = srows + 2 * len, pc = scols + 2 * len;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= pr * pc) return;
  const float mn = fminf(key_float(keys[2 * f]), 0.f), mx = fmaxf(key_float(keys[2 * f + 1]), 0.f);   // zero padding takes part in pm_norm
  const int r = i / pc, c = i - r * pc;
  float v = 0.f;
  if (r >= len && r < len + srows && c >= len && c < len + scols) v = sub[(size_t)f * srows * scols + (size_t)(r - len) * scols + (c - len)];
  pad[(size_t)f * pr * pc + i] = (v - mn) / (mx - mn);               // pm_norm.m (Sudoku :34)
}

int contrast_pad_launch(const float* sub, const unsigned* minmax_keys, int B, int srows, int scols, int len, float* pad,
                        cudaStream_t st) {
  dim3 grid(ceil_div((srows + 2 * len) * (scols + 2 * len), 256), B);
  contrast_pad_kernel<<<grid, 256, 0, st>>>(sub, minmax_keys, srows, scols, len, pad);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// ---- mv contrast input (cusize 1) in double ------------------------------------------------------
// Sudoku_contrastcpp5.m:18-34 on a signed map: +1e-8, zero pad, pm_norm.  The padded map's minimum cell
// becomes exactly 0 and computecontrast5 skips zeros (`p != 0`), so which cells tie with the minimum must
// be decided exactly as the reference's double arithmetic does; only the normalised result is cast to FP32.
__device__ __forceinline__ unsigned long long double_key(double d) {
  unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_double(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__global__ void __launch_bounds__(256)
contrast_minmax_f64_kernel(const double* __restrict__ src, int n, unsigned long long* __restrict__ keys) {
  __shared__ double redd[8];
  const int f = blockIdx.y;
  double mn = INFINITY, mx = -INFINITY;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const double v = src[(size_t)f * n + i] + 0.00000001;            // Sudoku :18
    mn = fmin(mn, v); mx = fmax(mx, v);
  }
  mn = block_reduce(mn, redd, OpMinD());
  mx = block_reduce(mx, redd, OpMaxD());
  if (threadIdx.x == 0 && mn <= mx) { atomicMin(keys + 2 * f, double_key(mn)); atomicMax(keys + 2 * f + 1, double_key(mx)); }
}
__global__ void __launch_bounds__(256)
contrast_pad_f64_kernel(const double* __restrict__ src, const unsigned long long* __restrict__ keys, int rows, int cols, int len,
                        float* __restrict__ pad) {
  const int f = blockIdx.y, pr = rows + 2 * len, pc = cols + 2 * len;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= pr * pc) return;
  const double mn = fmin(key_double(keys[2 * f]), 0.0), mx = fmax(key_double(keys[2 * f + 1]), 0.0);   // the zero padding takes part
  const int r = i / pc, c = i - r * pc;
  double v = 0.0;
  if (r >= len && r < len + rows && c >= len && c < len + cols) v = src[(size_t)f * rows * cols + (size_t)(r - len) * cols + (c - len)] + 0.00000001;
  pad[(size_t)f * pr * pc + i] = (float)((v - mn) / (mx - mn));     // pm_norm.m:2-4
}
int contrast_prep_f64_launch(const double* src, int nmaps, int rows, int cols, int len, unsigned long long* keys, float* pad,
                             cudaStream_t st) {
  // keys must have been initialised to {~0, 0} pairs by the caller
  dim3 g1(std::min(ceil_div(rows * cols, 256), 64), nmaps);
  contrast_minmax_f64_kernel<<<g1, 256, 0, st>>>(src, rows * cols, keys);
  CDS_LAUNCH_CHECK();
  dim3 g2(ceil_div((rows + 2 * len) * (cols + 2 * len), 256), nmaps);
  contrast_pad_f64_kernel<<<g2, 256, 0, st>>>(src, keys, rows, cols, len, pad);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

__global__ void __launch_bounds__(256)
contrast_post_kernel(const float* __restrict__ raw, const unsigned* __restrict__ maxbits, int srows, int scols, int cus,
                     int rows, int cols, float* __restrict__ Xmaps, int slot) {
  const int f = blockIdx.y, o = blockIdx.x * 256 + threadIdx.x;
  if (o >= rows * cols) return;
  const float mx = __uint_as_float(maxbits[f]);
  const int r = o / cols, c = o - r * cols;
  const float v = raw[(size_t)f * srows * scols + (size_t)(r / cus) * scols + (c / cus)];
  // contrast/max (Sudoku :50); an all-zero contrast map is NaN in the reference and ends as 0 (main.m:245-253)
  Xmaps[((size_t)f * 9 + slot) * rows * cols + o] = mx > 0.f ? v / mx : 0.f;
}

int contrast_post_launch(const float* raw, const unsigned* maxbits, int B, int srows, int scols, int cusize, int rows, int cols,
                         float* Xmaps, int slot, cudaStream_t st) {
  dim3 grid(ceil_div(rows * cols, 256), B);
  contrast_post_kernel<<<grid, 256, 0, st>>>(raw, maxbits, srows, scols, cusize, rows, cols, Xmaps, slot);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

__global__ void __launch_bounds__(256)
contrast_post_mv_kernel(const float* __restrict__ rawx, const float* __restrict__ rawy, const unsigned* __restrict__ maxx,
                        const unsigned* __restrict__ maxy, int n4, float* __restrict__ Xmaps, int slot) {
  const int f = blockIdx.y, o = blockIdx.x * 256 + threadIdx.x;
  if (o >= n4) return;
  const float mx = __uint_as_float(maxx[f]), my = __uint_as_float(maxy[f]);
  const float a = mx > 0.f ? rawx[(size_t)f * n4 + o] / mx : 0.f;    // NaN -> 0 (main.m:237-238)
  const float b = my > 0.f ? rawy[(size_t)f * n4 + o] / my : 0.f;
  Xmaps[((size_t)f * 9 + slot) * n4 + o] = sqrtf(a * a + b * b);     // main.m:241
}

int contrast_post_mv_launch(const float* rawx, const float* rawy, const unsigned* maxx, const unsigned* maxy, int B, int rows,
                            int cols, float* Xmaps, int slot, cudaStream_t st) {
  dim3 grid(ceil_div(rows * cols, 256), B);
  contrast_post_mv_kernel<<<grid, 256, 0, st>>>(rawx, rawy, maxx, maxy, rows * cols, Xmaps, slot);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// =================================================================================================
// full-resolution replicate + Gaussian as a polyphase filter
// =================================================================================================
// horizontal: T[m][r][4c+ph] = sum_d G[ph][d] X[m][r][c+d].  One CTA per (row, map); the zero-padded row sits in smem.
// A thread owns HP_CPT = 4 adjacent cells (16 outputs) and slides a 4-value window along the row: per 4 taps one
// conflict-free LDS.128 of inputs and four broadcast LDS.128 of weights feed 64 FFMAs.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float lo2(unsigned long long v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float hi2(unsigned long long v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
constexpr int HP_CPT = 4;
__global__ void __launch_bounds__(128)
fullres_hpass_kernel(const float* __restrict__ X, int h4, int w4, int dmin, int nt, int nt4, const float4* __restrict__ G,
                     float* __restrict__ T) {
  extern __shared__ __align__(16) float srow[];   // [w4p + nt4 + 4] inputs (zero padded), then nt4 float4 weights (zero padded to a multiple of 4 taps)
  const int w4p = ceil_div(w4, HP_CPT) * HP_CPT;
  const int nin = w4p + nt4 + 4;
  float4* sG = reinterpret_cast<float4*>(srow + nin);
  const int r = blockIdx.x, m = blockIdx.y;
  const float* __restrict__ xr = X + ((size_t)m * h4 + r) * w4;
  for (int i = threadIdx.x; i < nin; i += blockDim.x) { const int c = i + dmin; srow[i] = (c >= 0 && c < w4) ? xr[c] : 0.f; }
  for (int i = threadIdx.x; i < nt4; i += blockDim.x) sG[i] = i < nt ? G[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float4* __restrict__ out = reinterpret_cast<float4*>(T + ((size_t)m * h4 + r) * (size_t)w4 * 4);
  for (int c0 = threadIdx.x * HP_CPT; c0 < w4; c0 += blockDim.x * HP_CPT) {
    // packed FP32 (fma.rn.f32x2): the four phases of a cell are two register pairs, the input value is broadcast into a pair
    unsigned long long alo[HP_CPT], ahi[HP_CPT];
#pragma unroll
    for (int k = 0; k < HP_CPT; k++) { alo[k] = pack2(0.f, 0.f); ahi[k] = pack2(0.f, 0.f); }
    float4 w0 = *reinterpret_cast<const float4*>(srow + c0);          // x[c0 + t .. c0 + t + 3]
    for (int t = 0; t < nt4; t += 4) {
      const float4 w1 = *reinterpret_cast<const float4*>(srow + c0 + t + 4);
      const float win[7] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
      unsigned long long xx[7];
#pragma unroll
      for (int i = 0; i < 7; i++) xx[i] = pack2(win[i], win[i]);
#pragma unroll
      for (int tt = 0; tt < 4; tt++) {
        const float4 g = sG[t + tt];
        const unsigned long long glo = pack2(g.x, g.y), ghi = pack2(g.z, g.w);
#pragma unroll
        for (int k = 0; k < HP_CPT; k++) { alo[k] = ffma2(glo, xx[tt + k], alo[k]); ahi[k] = ffma2(ghi, xx[tt + k], ahi[k]); }
      }
      w0 = w1;
    }
    if (c0 + HP_CPT <= w4 && (w4 & 1) == 0) {      // the thread's 64 output bytes as two 256-bit stores (whole sectors)
#pragma unroll
      for (int k = 0; k < HP_CPT; k += 2)
        st_global_256(reinterpret_cast<float*>(out + c0 + k), make_float4(lo2(alo[k]), hi2(alo[k]), lo2(ahi[k]), hi2(ahi[k])),
                      make_float4(lo2(alo[k + 1]), hi2(alo[k + 1]), lo2(ahi[k + 1]), hi2(ahi[k + 1])));
    } else {
#pragma unroll
      for (int k = 0; k < HP_CPT; k++) if (c0 + k < w4) out[c0 + k] = make_float4(lo2(alo[k]), hi2(alo[k]), lo2(ahi[k]), hi2(ahi[k]));
    }
  }
}

int fullres_hpass_launch(const float* X, int nmaps, int h4, int w4, const Polyphase& P, float* T, cudaStream_t st) {
  dim3 grid(h4, nmaps);
  const int nt4 = ceil_div(P.nt, 4) * 4, w4p = ceil_div(w4, HP_CPT) * HP_CPT;
  const size_t smem = (size_t)(w4p + nt4 + 4) * 4 + (size_t)nt4 * 16;
  const int thr = std::min(128, ceil_div(ceil_div(w4, HP_CPT), 32) * 32);
  fullres_hpass_kernel<<<grid, thr, smem, st>>>(X, h4, w4, P.dmin, P.nt, nt4, reinterpret_cast<const float4*>(P.d_g), T);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// vertical + reductions: Y[m][4r+ph][x] = sum_d G[ph][d] T[m][r+d][x].
// One CTA owns a 128-column strip of one map and walks down it in tiles of VP_ROWS cell rows.  The (VP_ROWS + nt + 2)
// input rows of a tile are staged in shared memory with cp.async (out-of-picture rows zero-filled) into one of two
// buffers, so the next tile streams in from L2 while the current one is computed; staging cuts the L2 re-read factor
// from (nt+3)/4 to (VP_ROWS+nt)/VP_ROWS.  A thread owns 4 adjacent columns x VP_RPT consecutive cell rows (16 output
// rows): one LDS.128 of inputs and one new broadcast LDS.128 of weights per tap feed 16*VP_RPT FFMAs.  Taps run in
// ascending order for every output; the zero-weight pad entries add exact zeros.
constexpr int VP_RPT = 4;
constexpr int VP_ROWS = 32;                     // cell rows per tile
constexpr int VP_COLS4 = 32;                    // float4 columns per CTA
__global__ void __launch_bounds__(256, 2)
fullres_vpass_kernel(const float* __restrict__ T, int h4, int W, int dmin, int nt, const float4* __restrict__ G,
                     float* __restrict__ Ystore, const int* __restrict__ store_index, float* __restrict__ partial) {
  extern __shared__ float4 sm4[];               // weights [nt + 2*(VP_RPT-1)], then two input tiles [nin][VP_COLS4]
  __shared__ float redf[8];
  __shared__ double redd[8];
  const int ngw = nt + 2 * (VP_RPT - 1);
  const int nin = VP_ROWS + nt + VP_RPT - 2;    // input rows R0 + dmin .. R0 + dmin + nin - 1 of a tile
  float4* sGv = sm4;
  float4* tiles = sm4 + ngw;
  const int m = blockIdx.y, W4 = W >> 2;
  const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int x4 = blockIdx.x * VP_COLS4 + cg;
  const float4* __restrict__ Tm = reinterpret_cast<const float4*>(T + (size_t)m * h4 * W);
  auto stage = [&](int tile_idx, int buf) {     // warp rg copies whole 512-byte row segments
    float4* dst = tiles + (size_t)buf * nin * VP_COLS4;
    const int R0 = tile_idx * VP_ROWS;
    for (int i = rg; i < nin; i += 8) {
      const int gr = R0 + dmin + i;
      const bool ok = gr >= 0 && gr < h4 && x4 < W4;
      cp_async16_zfill(dst + i * VP_COLS4 + cg, ok ? (const void*)(Tm + (size_t)gr * W4 + x4) : (const void*)Tm, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int ntiles = ceil_div(h4, VP_ROWS);
  stage(0, 0);
  for (int i = threadIdx.x; i < ngw; i += 256) {
    const int t = i - (VP_RPT - 1);
    sGv[i] = (t >= 0 && t < nt) ? G[t] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int sidx = store_index ? store_index[m] : -1;
  float4* __restrict__ Ys = sidx >= 0 ? reinterpret_cast<float4*>(Ystore + (size_t)sidx * (size_t)h4 * 4 * W) : nullptr;
  float mn = INFINITY, mx = -INFINITY; double sum = 0;
  for (int ti = 0; ti < ntiles; ti++) {
    if (ti + 1 < ntiles) { stage(ti + 1, (ti + 1) & 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                            // tile ti has landed for every thread
    const int r0 = ti * VP_ROWS + rg * VP_RPT;
    if (x4 < W4 && r0 < h4) {
      // Packed FP32 (FFMA2, sm_100): an accumulator pair (x,y) / (z,w) of one phase is one 64-bit register pair and one
      // fma.rn.f32x2 does two IEEE FP32 FMAs, so the 64 FMAs per tap take 32 issue slots and leave the rest for the LDS.
      unsigned long long acc2[VP_RPT][4][2];    // [cell row][phase][xy | zw]
#pragma unroll
      for (int k = 0; k < VP_RPT; k++)
#pragma unroll
        for (int ph = 0; ph < 4; ph++) { acc2[k][ph][0] = 0ull; acc2[k][ph][1] = 0ull; }
      const float4* __restrict__ tp = tiles + (size_t)(ti & 1) * nin * VP_COLS4 + (rg * VP_RPT) * VP_COLS4 + cg;   // local input row of tap u of cell row r0
      unsigned long long gd[VP_RPT][4];         // gd[k][ph] = (w, w) with w = weight of tap u - k, phase ph
#pragma unroll
      for (int k = 1; k < VP_RPT; k++) { const float4 w = sGv[VP_RPT - 1 - k]; gd[k][0] = pack2(w.x, w.x); gd[k][1] = pack2(w.y, w.y); gd[k][2] = pack2(w.z, w.z); gd[k][3] = pack2(w.w, w.w); }
      auto step = [&](int u) {
        { const float4 w = sGv[u + VP_RPT - 1]; gd[0][0] = pack2(w.x, w.x); gd[0][1] = pack2(w.y, w.y); gd[0][2] = pack2(w.z, w.z); gd[0][3] = pack2(w.w, w.w); }
        const float4 v = tp[u * VP_COLS4];
        const unsigned long long vxy = pack2(v.x, v.y), vzw = pack2(v.z, v.w);
#pragma unroll
        for (int k = 0; k < VP_RPT; k++)
#pragma unroll
          for (int ph = 0; ph < 4; ph++) {
            acc2[k][ph][0] = ffma2(gd[k][ph], vxy, acc2[k][ph][0]);
            acc2[k][ph][1] = ffma2(gd[k][ph], vzw, acc2[k][ph][1]);
          }
#pragma unroll
        for (int k = VP_RPT - 1; k > 0; k--)
#pragma unroll
          for (int ph = 0; ph < 4; ph++) gd[k][ph] = gd[k - 1][ph];      // renamed away when the caller unrolls by VP_RPT
      };
      const int nu = nt + VP_RPT - 1;
      int u = 0;
      for (; u + VP_RPT <= nu; u += VP_RPT) {
#pragma unroll
        for (int q = 0; q < VP_RPT; q++) step(u + q);
      }
      for (; u < nu; u++) step(u);
      float4 acc[VP_RPT][4];
#pragma unroll
      for (int k = 0; k < VP_RPT; k++)
#pragma unroll
        for (int ph = 0; ph < 4; ph++) {
          acc[k][ph].x = lo2(acc2[k][ph][0]); acc[k][ph].y = hi2(acc2[k][ph][0]);
          acc[k][ph].z = lo2(acc2[k][ph][1]); acc[k][ph].w = hi2(acc2[k][ph][1]);
        }
      float tsum = 0.f;                        // 64 outputs of this tile in FP32 (the FIR itself is FP32), then FP64 across tiles / threads
#pragma unroll
      for (int k = 0; k < VP_RPT; k++) {
        if (r0 + k < h4) {
#pragma unroll
          for (int ph = 0; ph < 4; ph++) {
            const float4 a = acc[k][ph];
            mn = fminf(mn, fminf(fminf(a.x, a.y), fminf(a.z, a.w)));
            mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
            tsum += (a.x + a.y) + (a.z + a.w);
            if (Ys) Ys[(size_t)(4 * (r0 + k) + ph) * W4 + x4] = a;
          }
        }
      }
      sum += (double)tsum;
    }
    __syncthreads();                            // everyone is done with buffer ti & 1 before tile ti + 2 is staged into it
  }
  mn = block_reduce(mn, redf, OpMin());
  mx = block_reduce(mx, redf, OpMax());
  sum = block_reduce(sum, redd, OpAddD());
  if (threadIdx.x == 0) {
    float* p = partial + ((size_t)m * gridDim.x + blockIdx.x) * 4;
    p[0] = mn; p[1] = mx; reinterpret_cast<double*>(p + 2)[0] = sum;
  }
}

// =================================================================================================
// Vertical pass on the tensor cores (tcgen05 + TMEM).
//
// Y[4r+ph][x] = sum_t G[ph][t] T[r + dmin + t][x] is a banded GEMM: for a tile of 32 cell rows (128 output rows) the
// 128 x KP weight matrix A[m][k] = G[m % 4][k - m / 4] (KP = 32 + nt - 1 rounded up to 8) is THE SAME for every tile,
// strip and map, so it sits in shared memory for the life of the CTA (copied once with cp.async.bulk), and the input
// rows stream through as the B operand, 8 rows (one K step) at a time.  TF32 keeps 11 mantissa bits, so both operands
// are split hi + lo and three MMAs per K step accumulate lo*hi + hi*lo + hi*hi in FP32 (error ~2^-22, like the SVM
// kernel).  A CTA owns a 256-column strip of one map and walks down its row tiles: four staging warps load 8 x 256
// inputs per step, split them and transpose them into the K-major core-matrix layout the MMA reads; one thread issues the MMAs into
// one of two 256-column TMEM accumulators; four epilogue warps drain the other one (min / max / sum, and the store for
// the three temporal maps).  The FIR costs 3 x 72 x 2 tensor flops per output instead of 37 FFMA issue slots, which
// turns the pass from FP32-issue-bound into HBM-bound.
// =================================================================================================
namespace vtc {
using namespace tc;
constexpr int TM = 128, TN = 256, CELLS = TM / 4;       // N = 256: a tcgen05.mma with K = 8 has a ~140-cycle floor, smaller N only wastes it
constexpr int MAX_STAGES = 4;                     // operand ring (hi + lo of one K = 8 step per stage); 7 stages measured the same (shared-memory bandwidth bound)
constexpr int HALF_STAGE = 8 * TN * 4;            // one K = 8 step of hi (or lo): 8 rows x 128 columns
constexpr int STAGE_BYTES = 2 * HALF_STAGE;
constexpr int NCONV = 8;                          // converter warps: WPS per K step (2 k chunks x WPS/2 column parts), four steps in flight
constexpr int WPS = NCONV / 4;                    // measured: 8 warps 1.46 ms, 16 warps 1.66 ms per 64 pictures (shared-memory bandwidth is the wall)
constexpr int CBLK = 32 / (WPS / 2);              // 8-column blocks per converter warp
constexpr int THREADS = (6 + NCONV) * 32;         // warps 0-3 epilogue, warp 4 MMA + TMEM alloc, then NCONV converter warps, then the producer warp
constexpr int RSTAGES = 4, RSTAGES_MAX = 8;       // raw FP32 input ring filled by TMA: a multiple of the number of converter groups (4), so a stage always belongs to the same group, which then sees every phase of its barrier.  8 where shared memory allows: with 4 boxes of 8 KB in flight per SM the loop waited on TMA latency (~1000 cycles per K step against ~530 of shared-memory traffic)
constexpr int RAW_PITCH = TN * 4;                 // dense TMA box
constexpr int RAW_STAGE = 8 * RAW_PITCH;
constexpr int BAR_BYTES = 1024;
constexpr uint32_t TMEM_COLS = 512;               // two 256-column FP32 accumulators
constexpr uint32_t IDESC = idesc_tf32(TM, TN, 0, 0);   // both operands K-major
}  // namespace vtc

__global__ void __launch_bounds__(vtc::THREADS, 1)
fullres_vpass_tc_kernel(const __grid_constant__ CUtensorMap tmap /*T as [maps][h4][W] FP32, box [1][8][256]*/, int nmaps, int h4, int W, int dmin,
                        int kp, int nstages, int rstages, const float* __restrict__ aimg, float* __restrict__ Ystore, const int* __restrict__ store_index,
                        float* __restrict__ partial) {
  using namespace vtc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_half = kp * TM * 4;                 // bytes of A hi (then A lo)
  uint8_t* sA = smem;                             // [hi|lo][k chunk of 4][128 rows][4]   (K-major, no swizzle)
  uint8_t* sB = smem + 2 * a_half;                // STAGES x [hi|lo][k chunk of 4][256 columns][4]   (K-major, no swizzle)
  uint8_t* sR = sB + nstages * STAGE_BYTES;       // rstages x [8 rows][256] raw FP32 inputs (TMA boxes)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + rstages * RAW_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);
  float* red = reinterpret_cast<float*>(bars + 52);     // [4 warps][4]: min, max, sum as double
  const uint32_t bar0 = smem_u32(bars);
  constexpr int STAGES = MAX_STAGES;              // barrier slots; the ring itself uses `nstages` of them
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };
  const uint32_t abar = bar0 + 8u * (2 * STAGES + 4);
  auto rfull = [&](int s) { return bar0 + 8u * (2 * STAGES + 5 + s); };
  auto rempty = [&](int s) { return bar0 + 8u * (2 * STAGES + 5 + RSTAGES_MAX + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = 4 * h4;
  const int ntile = (H + TM - 1) / TM, ksteps = kp >> 3, total = ntile * ksteps;   // per work item
  const int nstrips = (W + TN - 1) / TN, nitems = nstrips * nmaps;
  // Persistent CTA (one per SM): work item = (column strip, map).  The weight matrix, the TMEM allocation and the barrier
  // phases live across items, so the pipeline never drains between items: every role walks the same item sequence with
  // global step / tile counters.

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; s++) { mbar_init(full(s), WPS); mbar_init(empty(s), 1); }   // WPS converter warps fill one stage
    for (int b = 0; b < 2; b++) { mbar_init(tfull(b), 1); mbar_init(tempty(b), 4); }
    mbar_init(abar, 1);
    for (int r = 0; r < rstages; r++) { mbar_init(rfull(r), 1); mbar_init(rempty(r), WPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(abar, 2u * a_half);
    bulk_g2s(smem_u32(sA), aimg, 2u * a_half, abar);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5 + NCONV) {
    if (lane == 0) {                                // ---- producer: one TMA box (8 rows x 256 columns) per K step; rows / columns outside the map arrive as zeros ----
      int g = 0;
      for (int wi = blockIdx.x; wi < nitems; wi += gridDim.x) {
        const int m = wi / nstrips, x0 = (wi - m * nstrips) * TN;
        for (int it = 0; it < total; it++, g++) {
          const int t = it / ksteps, ks = it - t * ksteps, rs = g % rstages;
          mbar_wait(rempty(rs), ((g / rstages) & 1) ^ 1);
          mbar_expect_tx(rfull(rs), (uint32_t)RAW_STAGE);
          const int r0 = CELLS * t + dmin + ks * 8;
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(smem_u32(sR + rs * RAW_STAGE)), "l"(&tmap), "r"(x0), "r"(r0), "r"(m), "r"(rfull(rs)) : "memory");
        }
      }
    }
  } else if (warp >= 5) {                           // ---- converter warps: raw FP32 -> split hi / lo -> UMMA K-major layout ----
    // Four groups of WPS warps take the K steps round-robin; a warp converts one 4-row k chunk of one column part.
    // Lane (kk, c) handles k = kk of the chunk and column c of an 8-column block.
    const int cw = warp - 5, pr = cw / WPS, kc = cw & 1, ch = (cw % WPS) >> 1;   // group pr converts steps g = pr, pr + 4, ...; a warp does one k chunk of one column part
    const int kk = lane & 3, c = lane >> 2;
    int nsteps = 0;
    for (int wi = blockIdx.x; wi < nitems; wi += gridDim.x) nsteps += total;
    for (int g = pr; g < nsteps; g += 4) {
      const int rs = g % rstages, s = g % nstages;
      mbar_wait(rfull(rs), (g / rstages) & 1);
      mbar_wait(empty(s), ((g / nstages) & 1) ^ 1);
      const uint8_t* raw = sR + rs * RAW_STAGE;
      const float* __restrict__ src = reinterpret_cast<const float*>(raw + (kc * 4 + kk) * RAW_PITCH) + ch * (CBLK * 8) + c;
      uint8_t* dst = sB + s * STAGE_BYTES + kc * (TN * 16) + (ch * (CBLK * 8) + c) * 16 + kk * 4;   // [k chunk][column][4 k]: this lane fills word kk of its column
      // The four kk lanes of a column read rows 1024 B apart (same bank), so each lane starts kk column blocks further
      // on: lanes then cover 32 different banks in the LDS and still 32 different banks in the transposing STS.
      float v[CBLK];                                // all loads first: the compiler will not hoist LDS above the STS below on its own
#pragma unroll
      for (int b8 = 0; b8 < CBLK; b8++) v[b8] = src[((b8 + kk) & (CBLK - 1)) * 8];
#pragma unroll
      for (int b8 = 0; b8 < CBLK; b8++) {
        const int nb = ((b8 + kk) & (CBLK - 1)) * 128;   // byte offset of column block (b8 + kk) % CBLK in the operand stage
        const uint32_t hi = tf32_rna(v[b8]);
        *reinterpret_cast<uint32_t*>(dst + nb) = hi;
        *reinterpret_cast<uint32_t*>(dst + HALF_STAGE + nb) = tf32_rna(v[b8] - __uint_as_float(hi));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(full(s)); mbar_arrive(rempty(rs)); }
    }
  } else if (warp == 4) {
    if (lane == 0) {                                // ---- MMA issuer ----
      mbar_wait(abar, 0);
      const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + a_half;
      int g = 0, tg = 0;
      for (int wi = blockIdx.x; wi < nitems; wi += gridDim.x) {
        for (int t = 0; t < ntile; t++, tg++) {
          const int buf = tg & 1;
          mbar_wait(tempty(buf), ((tg >> 1) & 1) ^ 1);
          const uint32_t d = tmem_base + (uint32_t)buf * TN;
          for (int ks = 0; ks < ksteps; ks++, g++) {
            const int s = g % nstages;
            mbar_wait(full(s), (g / nstages) & 1);
            tc_fence_after();
            const uint32_t b_hi = smem_u32(sB + s * STAGE_BYTES), b_lo = b_hi + HALF_STAGE;
            // A: K-major, LBO = distance between the two 16-byte k chunks of this step (128 rows x 16 B), SBO = 8-row groups
            const uint64_t ah = smem_desc(a_hi + ks * 2 * (TM * 16), TM * 16, 128), al = smem_desc(a_lo + ks * 2 * (TM * 16), TM * 16, 128);
            // B: K-major as well (the converter warps transpose on the fly; MN-major TF32 without swizzle gives zeros, tools/umma_probe.py)
            const uint64_t bh = smem_desc(b_hi, TN * 16, 128), bl = smem_desc(b_lo, TN * 16, 128);
            mma_tf32(d, al, bh, IDESC, ks > 0 ? 1u : 0u);
            mma_tf32(d, ah, bl, IDESC, 1u);
            mma_tf32(d, ah, bh, IDESC, 1u);
            mma_commit(empty(s));
          }
          mma_commit(tfull(buf));
        }
      }
    }
  } else {                                          // ---- epilogue warps 0-3: TMEM lanes 32*warp .. +31 = output rows of the tile ----
    int tg = 0;
    for (int wi = blockIdx.x; wi < nitems; wi += gridDim.x) {
      const int m = wi / nstrips, x0 = (wi - m * nstrips) * TN;
      const int sidx = store_index ? store_index[m] : -1;
      float* __restrict__ Ys = sidx >= 0 ? Ystore + (size_t)sidx * (size_t)H * W : nullptr;
      float mn = INFINITY, mx = -INFINITY; double sum = 0;
      for (int t = 0; t < ntile; t++, tg++) {
        const int buf = tg & 1;
        mbar_wait(tfull(buf), (tg >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * TN;
        const int y = t * TM + warp * 32 + lane;
        float tsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < TN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (y < H) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const int x = x0 + c + j;
              if (x < W) {                                     // W is a multiple of 8: a group of 8 columns is inside or outside as a whole
                const float4 a = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                const float4 b = make_float4(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]), __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                mn = fminf(mn, fminf(fminf(a.x, a.y), fminf(a.z, a.w)));
                mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
                tsum += (a.x + a.y) + (a.z + a.w);
                mn = fminf(mn, fminf(fminf(b.x, b.y), fminf(b.z, b.w)));
                mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
                tsum += (b.x + b.y) + (b.z + b.w);
                if (Ys) st_global_256(Ys + (size_t)y * W + x, a, b);      // a lane owns a row: 32 bytes = one whole sector per store
              }
            }
          }
        }
        sum += (double)tsum;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
      }
      // per-item statistics: the four epilogue warps meet on a named barrier (the other roles keep running)
      mn = warp_min(mn); mx = warp_max(mx); sum = warp_sum(sum);
      if (lane == 0) { red[warp * 4] = mn; red[warp * 4 + 1] = mx; reinterpret_cast<double*>(red + warp * 4 + 2)[0] = sum; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) {
        float m0 = red[0], m1 = red[1]; double sm = reinterpret_cast<double*>(red + 2)[0];
        for (int i = 1; i < 4; i++) { m0 = fminf(m0, red[i * 4]); m1 = fmaxf(m1, red[i * 4 + 1]); sm += reinterpret_cast<double*>(red + i * 4 + 2)[0]; }
        float* p = partial + (size_t)wi * 4;        // wi = m * nstrips + strip
        p[0] = m0; p[1] = m1; reinterpret_cast<double*>(p + 2)[0] = sm;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(vtc::TMEM_COLS) : "memory");
  }
}

// host: TMA descriptor of the T buffer viewed as [maps][h4][W] FP32 with an 8-row x 256-column box (zero fill outside)
static int vpass_tc_tensor_map(const float* T, int nmaps, int h4, int W, CUtensorMap* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CDS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return CDS_ECUDA; }
    encode = (encode_fn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)h4, (cuuint64_t)nmaps};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)h4 * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)vtc::TN, 8, 1}, estr[3] = {1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(T), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for %d maps of %d x %d", (int)r, nmaps, h4, W); return CDS_ECUDA; }
  return CDS_OK;
}

// host: the 128 x KP banded weight matrix of one tile, split hi / lo, as the shared-memory image the kernel bulk-copies
static int vpass_tc_prepare(const Polyphase& P, int* kp_out, const float** img_out) {
  Polyphase& PP = const_cast<Polyphase&>(P);
  const int kp = ceil_div(vtc::CELLS + P.nt - 1, 8) * 8;
  if (!PP.d_aimg) {
    auto tf32 = [](float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&v, &u, 4); return v; };
    std::vector<float> img((size_t)2 * kp * vtc::TM, 0.f);
    for (int mrow = 0; mrow < vtc::TM; mrow++)
      for (int k = 0; k < kp; k++) {
        const int tap = k - mrow / 4;
        const float wv = (tap >= 0 && tap < P.nt) ? P.g[(size_t)tap * 4 + (mrow % 4)] : 0.f;
        const float hi = tf32(wv), lo = tf32(wv - hi);
        const size_t o = ((size_t)(k / 4) * vtc::TM + mrow) * 4 + (k % 4);
        img[o] = hi; img[(size_t)kp * vtc::TM + o] = lo;
      }
    CDS_CUDA(cudaMalloc(&PP.d_aimg, img.size() * 4));
    CDS_CUDA(cudaMemcpy(PP.d_aimg, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  }
  *kp_out = kp; *img_out = PP.d_aimg;
  return CDS_OK;
}

int fir_threads(int W4) {                       // 96 or 128 threads, whichever wastes fewer columns
  const int w96 = ceil_div(W4, 96) * 96 - W4, w128 = ceil_div(W4, 128) * 128 - W4;
  return w96 < w128 ? 96 : 128;
}
int fullres_vpass_nblocks(int h4, int W) { (void)h4; return ceil_div(W / 4, VP_COLS4); }

int fullres_vpass_launch(const float* T, int nmaps, int h4, int W, const Polyphase& P, float* Ystore, const int* store_index,
                         float* partial, int* nblocks_out, cudaStream_t st) {
  static const int impl = [] { const char* e = getenv("CDS_VPASS_IMPL"); return (e && !strcmp(e, "fp32")) ? 0 : 1; }();
  if (impl == 1) {
    int kp = 0; const float* img = nullptr;
    if (int e = vpass_tc_prepare(P, &kp, &img)) return e;
    static const int rs_env = [] { const char* e = getenv("CDS_VPASS_RSTAGES"); return e ? atoi(e) : 0; }();
    int rstages = rs_env == 4 || rs_env == 8 ? rs_env : vtc::RSTAGES_MAX;
    size_t fixed = (size_t)2 * kp * vtc::TM * 4 + (size_t)rstages * vtc::RAW_STAGE + vtc::BAR_BYTES;
    if (fixed + (size_t)vtc::MAX_STAGES * vtc::STAGE_BYTES > 220 * 1024) {      // wide filters (4K: kp = 104): the weight tile leaves room for 4 raw stages only
      rstages = vtc::RSTAGES;
      fixed = (size_t)2 * kp * vtc::TM * 4 + (size_t)rstages * vtc::RAW_STAGE + vtc::BAR_BYTES;
    }
    const int nstages = (int)std::min<size_t>(vtc::MAX_STAGES, fixed < 220 * 1024 ? (220 * 1024 - fixed) / vtc::STAGE_BYTES : 0);
    const size_t smem = fixed + (size_t)nstages * vtc::STAGE_BYTES;
    if (nstages >= 4) {
      if (int e = ensure_dyn_smem(fullres_vpass_tc_kernel, smem)) return e;
      const int nstrips = ceil_div(W, vtc::TN);
      CUtensorMap tmap;
      if (int e = vpass_tc_tensor_map(T, nmaps, h4, W, &tmap)) return e;
      const int sms = device_sm_count();
      dim3 grid(std::min(sms, nstrips * nmaps));
fullres_vpass_tc_kernel<<<grid, vtc::THREADS, smem, st>>>(tmap, nmaps, h4, W, P.dmin, kp, nstages, rstages, img, Ystore, store_index, partial);
      CDS_LAUNCH_CHECK();
      *nblocks_out = nstrips;
      return CDS_OK;
    }
  }
  dim3 grid(ceil_div(W / 4, VP_COLS4), nmaps);
  const size_t smem = (size_t)(P.nt + 2 * (VP_RPT - 1)) * 16 + (size_t)2 * (VP_ROWS + P.nt + VP_RPT - 2) * VP_COLS4 * 16;
  if (int e = ensure_dyn_smem(fullres_vpass_kernel, smem)) return e;
  fullres_vpass_kernel<<<grid, 256, smem, st>>>(T, h4, W, P.dmin, P.nt, reinterpret_cast<const float4*>(P.d_g), Ystore, store_index, partial);
  CDS_LAUNCH_CHECK();
  *nblocks_out = grid.x;
  return CDS_OK;
}

// partial[m][nb][4] = (min, max, sum as double) -> stats[m][4] = (min, max, sum as double); one warp per map, fixed order.
__global__ void __launch_bounds__(32)
stats_finalize_kernel(const float* __restrict__ partial, int nb, float* __restrict__ stats) {
  const int m = blockIdx.x, lane = threadIdx.x;
  float mn = INFINITY, mx = -INFINITY; double s = 0;
  for (int i = lane; i < nb; i += 32) {
    const float* p = partial + ((size_t)m * nb + i) * 4;
    mn = fminf(mn, p[0]); mx = fmaxf(mx, p[1]); s += reinterpret_cast<const double*>(p + 2)[0];
  }
  mn = warp_min(mn); mx = warp_max(mx); s = warp_sum(s);
  if (lane == 0) { float* o = stats + (size_t)m * 4; o[0] = mn; o[1] = mx; reinterpret_cast<double*>(o + 2)[0] = s; }
}

int stats_finalize_launch(const float* partial, int nmaps, int nblocks, float* stats, cudaStream_t st) {
  stats_finalize_kernel<<<nmaps, 32, 0, st>>>(partial, nblocks, stats);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// =================================================================================================
// generic banded operators
// =================================================================================================
// Map selection shared by the row and column passes: with `outer_of_three` the launch covers only maps 0 and 2 of every three
// (the 200x200 features take the middle one -- the temporal map -- from the full-resolution path, main.m:284-294).
__device__ __forceinline__ int banded_map(int z, int outer_of_three) { return outer_of_three ? (z >> 1) * 3 + (z & 1) * 2 : z; }

// out[m][i][x] = sum_t w[i][t] * f(in[m][start[i]+t][x]);  f = identity or min((v-a)*b, thr) per map.
// A thread owns 4 adjacent columns (LDG.128 / STG.128); cols must be a multiple of 4.
template <bool XF>
__global__ void __launch_bounds__(128)
banded_rows_kernel(const float* __restrict__ in, int in_rows, int cols, const float* __restrict__ w, const int* __restrict__ start,
                   int nt, int n_out, const float* __restrict__ xf, float* __restrict__ out) {
  extern __shared__ float sw[];                 // nt weights of this output row
  const int i = blockIdx.y, m = blockIdx.z, C4 = cols >> 2;
  for (int t = threadIdx.x; t < nt; t += blockDim.x) sw[t] = w[(size_t)i * nt + t];
  __syncthreads();
  const int s0 = start[i];
  const int t_hi = min(nt, in_rows - s0);
  float a = 0.f, b = 1.f, thr = INFINITY;
  if (XF) { a = xf[4 * m]; b = xf[4 * m + 1]; thr = xf[4 * m + 2]; }
  const float4* __restrict__ src = reinterpret_cast<const float4*>(in + (size_t)m * in_rows * cols);
  float4* __restrict__ dst = reinterpret_cast<float4*>(out + ((size_t)m * n_out + i) * cols);
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < C4; x += gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int t = 0; t < t_hi; t++) {
      float4 v = __ldg(src + (size_t)(s0 + t) * C4 + x);
      if (XF) { v.x = fminf((v.x - a) * b, thr); v.y = fminf((v.y - a) * b, thr); v.z = fminf((v.z - a) * b, thr); v.w = fminf((v.w - a) * b, thr); }
      const float g = sw[t];
      acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y); acc.z = fmaf(g, v.z, acc.z); acc.w = fmaf(g, v.w, acc.w);
    }
    dst[x] = acc;
  }
}

// Four consecutive output rows per CTA (operators whose rows overlap heavily, e.g. resize o gaussian): the rows' weights
// are staged over their common input window as float4, so one LDG.128 of inputs + one broadcast LDS.128 feed 16 FFMAs
// and each input row is fetched once per 4 output rows.  Taps stay in ascending order; pad weights add exact zeros.
constexpr int BR_ROWS = 4;
template <bool XF>
__global__ void __launch_bounds__(128)
banded_rows4_kernel(const float* __restrict__ in, int in_rows, int cols, const float* __restrict__ w, const int* __restrict__ start,
                    int nt, int n_out, const float* __restrict__ xf, float* __restrict__ out, int outer_of_three) {
  extern __shared__ float4 sw4[];               // [window] x 4 rows
  const int i0 = blockIdx.y * BR_ROWS, m = banded_map(blockIdx.z, outer_of_three), C4 = cols >> 2;
  const int ilast = min(i0 + BR_ROWS - 1, n_out - 1);
  const int s_lo = start[i0], L = min(start[ilast] + nt, in_rows) - s_lo;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float g[BR_ROWS];
#pragma unroll
    for (int k = 0; k < BR_ROWS; k++) {
      const int i = i0 + k;
      const int t = i < n_out ? j + s_lo - start[min(i, n_out - 1)] : -1;
      g[k] = (t >= 0 && t < nt) ? w[(size_t)i * nt + t] : 0.f;
    }
    sw4[j] = make_float4(g[0], g[1], g[2], g[3]);
  }
  __syncthreads();
  const float4* __restrict__ src = reinterpret_cast<const float4*>(in + (size_t)m * in_rows * cols);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= C4) return;
  float fa = 0.f, fb = 1.f, fthr = INFINITY;
  if (XF) { fa = xf[4 * m]; fb = xf[4 * m + 1]; fthr = xf[4 * m + 2]; }
  float4 acc[BR_ROWS];
#pragma unroll
  for (int k = 0; k < BR_ROWS; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int j = 0; j < L; j++) {
    float4 v = __ldg(src + (size_t)(s_lo + j) * C4 + x);
    if (XF) { v.x = fminf((v.x - fa) * fb, fthr); v.y = fminf((v.y - fa) * fb, fthr); v.z = fminf((v.z - fa) * fb, fthr); v.w = fminf((v.w - fa) * fb, fthr); }
    const float4 g = sw4[j];
    acc[0].x = fmaf(g.x, v.x, acc[0].x); acc[0].y = fmaf(g.x, v.y, acc[0].y); acc[0].z = fmaf(g.x, v.z, acc[0].z); acc[0].w = fmaf(g.x, v.w, acc[0].w);
    acc[1].x = fmaf(g.y, v.x, acc[1].x); acc[1].y = fmaf(g.y, v.y, acc[1].y); acc[1].z = fmaf(g.y, v.z, acc[1].z); acc[1].w = fmaf(g.y, v.w, acc[1].w);
    acc[2].x = fmaf(g.z, v.x, acc[2].x); acc[2].y = fmaf(g.z, v.y, acc[2].y); acc[2].z = fmaf(g.z, v.z, acc[2].z); acc[2].w = fmaf(g.z, v.w, acc[2].w);
    acc[3].x = fmaf(g.w, v.x, acc[3].x); acc[3].y = fmaf(g.w, v.y, acc[3].y); acc[3].z = fmaf(g.w, v.z, acc[3].z); acc[3].w = fmaf(g.w, v.w, acc[3].w);
  }
#pragma unroll
  for (int k = 0; k < BR_ROWS; k++)
    if (i0 + k < n_out) reinterpret_cast<float4*>(out + ((size_t)m * n_out + i0 + k) * cols)[x] = acc[k];
}

// scalar-column variant for widths that are not a multiple of 4 (cell grids of pictures whose width is 8 mod 16)
__global__ void __launch_bounds__(128)
banded_rows_scalar_kernel(const float* __restrict__ in, int in_rows, int cols, const float* __restrict__ w, const int* __restrict__ start,
                          int nt, int n_out, const float* __restrict__ xf, float* __restrict__ out) {
  extern __shared__ float sw[];
  const int i = blockIdx.y, m = blockIdx.z;
  for (int t = threadIdx.x; t < nt; t += 128) sw[t] = w[(size_t)i * nt + t];
  __syncthreads();
  const int s0 = start[i];
  const int t_hi = min(nt, in_rows - s0);
  float a = 0.f, b = 1.f, thr = INFINITY;
  if (xf) { a = xf[4 * m]; b = xf[4 * m + 1]; thr = xf[4 * m + 2]; }
  const float* __restrict__ src = in + (size_t)m * in_rows * cols;
  for (int x = blockIdx.x * 128 + threadIdx.x; x < cols; x += gridDim.x * 128) {
    float acc = 0.f;
    for (int t = 0; t < t_hi; t++) {
      float v = __ldg(src + (size_t)(s0 + t) * cols + x);
      if (xf) v = fminf((v - a) * b, thr);
      acc = fmaf(sw[t], v, acc);
    }
    out[((size_t)m * n_out + i) * cols + x] = acc;
  }
}

int banded_rows_launch(const float* in, int nmaps, int in_rows, int cols, const Banded& B, const float* xf, float* out, cudaStream_t st,
                       int outer_of_three) {
  // `outer_of_three` is a hint (skipped maps are simply not needed): only the four-row kernel below implements it, the others do every map
  if (nmaps % 3 != 0) outer_of_three = 0;
  if (cols % 4) {
    dim3 grid(min(ceil_div(cols, 128), 64), B.n_out, nmaps);
    banded_rows_scalar_kernel<<<grid, 128, (size_t)B.nt * 4, st>>>(in, in_rows, cols, B.d_w, B.d_start, B.nt, B.n_out, xf, out);
    CDS_LAUNCH_CHECK();
    return CDS_OK;
  }
  const int thr = fir_threads(cols / 4);
  // overlapping rows: 4 output rows per CTA, each input row fetched once for the four.  Shrinking rows (resize 1080 -> 200: 12 taps,
  // starts 5.4 apart) overlap about two-fold; one output row per CTA re-fetched every input row twice through the L2, which was the
  // bound of that pass (1.6 GB at 4.2 TB/s).
  if (B.span4 >= 0 && B.span4 <= 2 * B.nt) {
    dim3 grid4(ceil_div(cols / 4, thr), ceil_div(B.n_out, BR_ROWS), nmaps);
    const size_t smem = (size_t)(B.nt + B.span4) * 16;
    if (outer_of_three) grid4.z = nmaps / 3 * 2;
    if (xf) banded_rows4_kernel<true><<<grid4, thr, smem, st>>>(in, in_rows, cols, B.d_w, B.d_start, B.nt, B.n_out, xf, out, outer_of_three);
    else    banded_rows4_kernel<false><<<grid4, thr, smem, st>>>(in, in_rows, cols, B.d_w, B.d_start, B.nt, B.n_out, xf, out, outer_of_three);
    CDS_LAUNCH_CHECK();
    return CDS_OK;
  }
  dim3 grid(ceil_div(cols / 4, thr), B.n_out, nmaps);
  if (xf) banded_rows_kernel<true><<<grid, thr, (size_t)B.nt * 4, st>>>(in, in_rows, cols, B.d_w, B.d_start, B.nt, B.n_out, xf, out);
  else    banded_rows_kernel<false><<<grid, thr, (size_t)B.nt * 4, st>>>(in, in_rows, cols, B.d_w, B.d_start, B.nt, B.n_out, xf, out);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// out[m][r][j] = sum_t w[j][t] * in[m][r][start[j]+t].
// One CTA per (group of 8 rows, map).  The rows are staged TRANSPOSED in shared memory -- 12 floats per input column: rows 0-3,
// rows 4-7, 4 floats of padding that spread the 16-byte bank groups -- so a thread, which owns output column j for all 8 rows,
// feeds 8 FFMAs from two LDS.128 and one coalesced weight load (transposed weights wT[t][j]); with the rows side by side it was one
// LDS.32 per FFMA and the kernel sat at the LSU limit.  Staging reads float4s along the rows (in_cols % 4 == 0) and turns
// 4 rows x 4 columns into four 16-byte stores in registers.  Measured against the row-major form (one LDS.32 per FFMA): 2.2x faster
// where the pass grows (200 -> 1920 columns: the lanes' windows coincide, broadcast loads), the same where it shrinks (480 -> 200,
// 1920 -> 200: windows 2.4 / 9.6 columns apart collide two-fold in the banks).
constexpr int BC_ROWS = 8;
__global__ void __launch_bounds__(256)
banded_cols_kernel(const float* __restrict__ in, int rows, int in_cols, const float* __restrict__ wT, const int* __restrict__ start,
                   int nt, int n_out, float* __restrict__ out, int outer_of_three) {
  extern __shared__ float4 scol[];              // [in_cols][3]
  const int r0 = blockIdx.x * BC_ROWS, m = banded_map(blockIdx.y, outer_of_three);
  const int nr = min(BC_ROWS, rows - r0);
  const float* __restrict__ src = in + ((size_t)m * rows + r0) * in_cols;
  if ((in_cols & 3) == 0) {
    const int C4 = in_cols >> 2;
    for (int idx = threadIdx.x; idx < 2 * C4; idx += 256) {
      const int g = idx / C4, c4 = idx - g * C4;
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        v[k] = 4 * g + k < nr ? __ldg(reinterpret_cast<const float4*>(src + (size_t)(4 * g + k) * in_cols) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4* d = scol + (size_t)(4 * c4) * 3 + g;
      d[0] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
      d[3] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
      d[6] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
      d[9] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
    }
  } else {
    for (int idx = threadIdx.x; idx < 2 * in_cols; idx += 256) {
      const int g = idx / in_cols, c = idx - g * in_cols;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = 4 * g + k < nr ? __ldg(src + (size_t)(4 * g + k) * in_cols + c) : 0.f;
      scol[(size_t)c * 3 + g] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_out; j += 256) {
    const int s0 = start[j], t_hi = min(nt, in_cols - s0);
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    const float4* __restrict__ sp = scol + (size_t)s0 * 3;
#pragma unroll 4
    for (int t = 0; t < t_hi; t++) {
      const float wv = __ldg(wT + (size_t)t * n_out + j);
      const float4 a = sp[3 * t], b = sp[3 * t + 1];
      lo.x = fmaf(wv, a.x, lo.x); lo.y = fmaf(wv, a.y, lo.y); lo.z = fmaf(wv, a.z, lo.z); lo.w = fmaf(wv, a.w, lo.w);
      hi.x = fmaf(wv, b.x, hi.x); hi.y = fmaf(wv, b.y, hi.y); hi.z = fmaf(wv, b.z, hi.z); hi.w = fmaf(wv, b.w, hi.w);
    }
    const float acc[BC_ROWS] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int r = 0; r < BC_ROWS; r++) if (r < nr) out[((size_t)m * rows + r0 + r) * n_out + j] = acc[r];
  }
}

int banded_cols_launch(const float* in, int nmaps, int rows, int in_cols, const Banded& B, float* out, cudaStream_t st, int outer_of_three) {
  if (nmaps % 3 != 0) outer_of_three = 0;
  dim3 grid(ceil_div(rows, BC_ROWS), outer_of_three ? nmaps / 3 * 2 : nmaps);
  const size_t smem = (size_t)in_cols * 48;
  if (int e = ensure_dyn_smem(banded_cols_kernel, smem)) return e;
  banded_cols_kernel<<<grid, 256, smem, st>>>(in, rows, in_cols, B.d_wT, B.d_start, B.nt, B.n_out, out, outer_of_three);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

}  // namespace cds
