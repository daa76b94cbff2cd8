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
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
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
  // F == 9 tensor-core path (svm_rbf9_tc_kernel): SVs grouped by coefficient sign and packed as ready-made shared-memory
  // images of 128-SV tiles, [tile][hi|lo][k chunk 0..3][128 SVs][4] TF32-rounded floats (16 KB per tile)
  float* d_tiles = nullptr;
  int ntiles = 0, ntiles_pos = 0;
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

// =================================================================================================
// Tensor-core path (tcgen05 + TMEM): the exponent of every (instance, SV) pair as one TF32 GEMM.
//
//   log2(|coef| * K(x, s)) = x'.s' - |x'|^2/2 - |s'|^2/2 + log2|coef|  =  xa . sa   with the augmented K = 16 vectors
//   xa = [x'0..x'8, 1, -|x'|^2/2, 0...],  sa = [s'0..s'8, -|s'|^2/2 + log2|coef|, 1, 0...]
// TF32 keeps 11 mantissa bits, far short of the 1e-4 tolerance on exp(.), so both operands are split
// hi + lo (each TF32-exact) and three MMAs are accumulated in FP32: lo*hi + hi*lo + hi*hi (the dropped
// lo*lo term is 2^-22 relative).  A CTA owns 128 instances (A tile, built once in shared memory) and streams
// all 128-SV tiles: a producer thread moves 16 KB tile images with cp.async.bulk onto an mbarrier ring, one
// thread issues 6 tcgen05.mma (128x128x8) per tile into one of two 128-column TMEM accumulators, and the four
// epilogue warps read the other accumulator with tcgen05.ld, apply MUFU.EX2 and add.  SVs are grouped by the
// sign of coef, so a tile adds to a "positive" or a "negative" sum and no coefficient is read in the epilogue:
// 1 MUFU + 1 FADD per pair; the kernel is MUFU-bound (16 ex2 / clk / SM), not FP32-issue-bound.
// =================================================================================================
namespace tc {
constexpr int TM = 128, TN = 128, KPAD = 16, STAGES = 4;
static_assert(TN == 128, "the epilogue below is unrolled for four 32-column TMEM loads");
constexpr int TILE_FLOATS = 2 * KPAD * TN;                 // hi + lo
constexpr int TILE_BYTES = TILE_FLOATS * 4;                // 16 KB
constexpr int HALF_BYTES = KPAD * TN * 4;                  // hi -> lo offset (8 KB); also A hi -> A lo
constexpr int KSTEP_BYTES = 2 * TN * 16;                   // two 16-byte k chunks of 128 rows = one K=8 MMA step
constexpr int A_BYTES = 2 * KPAD * TM * 4;                 // 16 KB
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = A_BYTES + STAGES * TILE_BYTES + BAR_BYTES;
constexpr int THREADS = 192;                               // warps 0-3 epilogue, warp 4 MMA + TMEM alloc, warp 5 producer
constexpr uint32_t TMEM_COLS = 256;                        // two 128-column FP32 accumulators
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = idesc_tf32(TM, TN, 0, 0);

// 2^x on the FMA pipe (round-to-nearest split, degree-5 minimax polynomial on [-0.5, 0.5], exponent patched in with an
// integer add): max relative error 2.3e-7, the same as MUFU.EX2's 2^-22.  The epilogue is MUFU-bound (16 ex2 / clk / SM)
// while the FMA pipe idles, so a fixed share of the exponentials is computed here instead.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);                                      // 2^-125 ~ 2e-38: indistinguishable from the flushed MUFU result
  const float fi = x + 12582912.f;                           // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (fi - 12582912.f);
  float p = 0.0013266971f;
  p = fmaf(p, f, 0.00967546f); p = fmaf(p, f, 0.055507425f); p = fmaf(p, f, 0.24022122f);
  p = fmaf(p, f, 0.69314694f); p = fmaf(p, f, 1.0000001f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(fi) << 23));
}
}  // namespace tc

// POLY_OF_8: of every 8 exponentials, this many are computed on the FMA pipe (ex2_poly) instead of MUFU.EX2
template <int POLY_OF_8>
__global__ void __launch_bounds__(tc::THREADS, 2)
svm_rbf9_tc_kernel(const float* __restrict__ X, long N, const float* __restrict__ tiles, int ntiles, int ntiles_pos,
                   float xscale, float rho, float* __restrict__ dec) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                        // [hi|lo][k chunk][128 instances][4]
  uint8_t* sB = smem + A_BYTES;                              // STAGES tile images
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + STAGES * TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row0 = (long)blockIdx.x * TM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; b++) { mbar_init(tfull(b), 1); mbar_init(tempty(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < TM) {                                    // one instance per thread: augmented, scaled, split hi/lo
    const long n = row0 + threadIdx.x;
    float xa[KPAD];
    float s2 = 0.f;
#pragma unroll
    for (int d = 0; d < 9; d++) { const float v = (n < N) ? __ldg(X + n * 9 + d) * xscale : 0.f; xa[d] = v; s2 = fmaf(v, v, s2); }
    xa[9] = 1.f; xa[10] = -0.5f * s2;
#pragma unroll
    for (int d = 11; d < KPAD; d++) xa[d] = 0.f;
#pragma unroll
    for (int kc = 0; kc < 4; kc++) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; j++) { hi[j] = tf32_rna(xa[4 * kc + j]); lo[j] = tf32_rna(xa[4 * kc + j] - __uint_as_float(hi[j])); }
      *reinterpret_cast<uint4*>(sA + (kc * TM + threadIdx.x) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + HALF_BYTES + (kc * TM + threadIdx.x) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor-core (async) proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    if (lane == 0) {                                         // ---- producer: SV tile images, global -> shared ----
      for (int t = 0; t < ntiles; t++) {
        const int s = t % STAGES;
        mbar_wait(empty(s), ((t / STAGES) & 1) ^ 1);
        mbar_expect_tx(full(s), TILE_BYTES);
        bulk_g2s(smem_u32(sB + s * TILE_BYTES), tiles + (size_t)t * TILE_FLOATS, TILE_BYTES, full(s));
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {                                         // ---- MMA issuer ----
      const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + HALF_BYTES;
      for (int t = 0; t < ntiles; t++) {
        const int s = t % STAGES, buf = t & 1;
        mbar_wait(tempty(buf), ((t >> 1) & 1) ^ 1);
        mbar_wait(full(s), (t / STAGES) & 1);
        tc_fence_after();
        const uint32_t b_hi = smem_u32(sB + s * TILE_BYTES), b_lo = b_hi + HALF_BYTES;
        const uint32_t d = tmem_base + (uint32_t)buf * TN;
        const uint32_t LBO = TN * 16, SBO = 128;             // TM == TN
#define CDS_DESC(base, ks) smem_desc((base) + (ks) * KSTEP_BYTES, LBO, SBO)
        mma_tf32(d, CDS_DESC(a_lo, 0), CDS_DESC(b_hi, 0), IDESC, 0u);   // small terms first
        mma_tf32(d, CDS_DESC(a_lo, 1), CDS_DESC(b_hi, 1), IDESC, 1u);
        mma_tf32(d, CDS_DESC(a_hi, 0), CDS_DESC(b_lo, 0), IDESC, 1u);
        mma_tf32(d, CDS_DESC(a_hi, 1), CDS_DESC(b_lo, 1), IDESC, 1u);
        mma_tf32(d, CDS_DESC(a_hi, 0), CDS_DESC(b_hi, 0), IDESC, 1u);
        mma_tf32(d, CDS_DESC(a_hi, 1), CDS_DESC(b_hi, 1), IDESC, 1u);
#undef CDS_DESC
        mma_commit(empty(s));                                // shared-memory stage free once these MMAs have read it
        mma_commit(tfull(buf));                              // accumulator ready for the epilogue
      }
    }
  } else {                                                   // ---- epilogue warps 0-3: TMEM lanes 32*warp .. +31 ----
    float accp[4] = {0.f, 0.f, 0.f, 0.f}, accn[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < ntiles; t++) {
      const int buf = t & 1;
      mbar_wait(tfull(buf), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * TN;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      auto consume = [&](const uint32_t (&v)[32]) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float e[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float x = __uint_as_float(v[j + q]);
            e[q] = (q < POLY_OF_8) ? ex2_poly(x) : ex2_approx(x);
          }
          a0 += e[0] + e[4]; a1 += e[1] + e[5]; a2 += e[2] + e[6]; a3 += e[3] + e[7];
        }
      };
      // register double buffering: the next 32 columns are in flight from TMEM while the MUFUs of the current ones issue
      uint32_t va[32], vb[32];
      tmem_ld32(taddr, va);
      tmem_ld_wait();
      tmem_ld32(taddr + 32, vb);
      consume(va);
      tmem_ld_wait();
      tmem_ld32(taddr + 64, va);
      consume(vb);
      tmem_ld_wait();
      tmem_ld32(taddr + 96, vb);
      consume(va);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));               // this warp's quarter of the accumulator is drained (all loads waited on)
      consume(vb);
      if (t < ntiles_pos) { accp[0] += a0; accp[1] += a1; accp[2] += a2; accp[3] += a3; }
      else                { accn[0] += a0; accn[1] += a1; accn[2] += a2; accn[3] += a3; }
    }
    const long n = row0 + threadIdx.x;
    if (n < N) dec[n] = ((accp[0] + accp[1]) + (accp[2] + accp[3])) - ((accn[0] + accn[1]) + (accn[2] + accn[3])) - rho;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// =================================================================================================
// Persistent variant of the kernel above: ONE CTA per SM walks the 128-instance tiles of the whole launch, so the kernel
// can share every SM with a kernel that is bound by a different pipe (the FP32-issue-bound contrast kernel of the next
// batch, pipeline.cu): it holds 10 warps, <= 128 registers per thread and 98 KB of shared memory per SM for its whole life
// and leaves the rest to the co-resident CTAs.  Roles: warps 0-7 epilogue (TMEM lane quadrant = warp % 4, warp set =
// warp / 4; a set takes every other SV tile, so the two warps of a scheduler run half a tile apart), warp 8 MMA issuer + TMEM
// allocation (four accumulators), warp 9 producer (lane 0 streams the SV tile images; the whole warp builds the A tile of the NEXT instance
// tile into the second A buffer, so the MMA pipe never waits for it).  mbarrier phases run on global counters across tiles.
// =================================================================================================
namespace tcp {
using namespace tc;
constexpr int EPI_WARPS = 8;
constexpr int NACC = 4;                                    // 128-column FP32 accumulators: all 512 TMEM columns (one CTA of this kernel per SM)
constexpr uint32_t ACC_COLS = NACC * TN;
static_assert(2 * STAGES + 2 * NACC + 4 <= 24, "mbarrier block");
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int RED_BYTES = 2 * 2 * TM * 4;                  // [tile parity][pos | neg][instance]
constexpr int SMEM_BYTES = 2 * A_BYTES + STAGES * TILE_BYTES + BAR_BYTES + RED_BYTES;
}  // namespace tcp

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ unsigned long long fadd2_pk(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

template <int POLY_OF_8>
__global__ void __maxnreg__(128)
svm_rbf9_tcp_kernel(const float* __restrict__ X, long N, const float* __restrict__ tiles, int ntiles, int ntiles_pos,
                    float xscale, float rho, float* __restrict__ dec) {
  using namespace tcp;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                        // 2 x [hi|lo][k chunk][128 instances][4]
  uint8_t* sB = smem + 2 * A_BYTES;                          // STAGES tile images
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty = [&](int b) { return bar0 + 8u * (2 * STAGES + NACC + b); };
  auto afull = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 * NACC + b); };
  auto aempty = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 * NACC + 2 + b); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long ntile_inst = (N + TM - 1) / TM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < NACC; b++) { mbar_init(tfull(b), 1); mbar_init(tempty(b), EPI_WARPS / 2); }
    for (int b = 0; b < 2; b++) { mbar_init(afull(b), 1); mbar_init(aempty(b), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ACC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == EPI_WARPS + 1) {                               // ---- producer warp ----
    auto build_a = [&](long tile, int k) {                   // the whole warp: 4 instances per lane -> A buffer k & 1
      const int ab = k & 1;
      mbar_wait(aempty(ab), ((k >> 1) & 1) ^ 1);
      uint8_t* A = sA + ab * A_BYTES;
#pragma unroll 1
      for (int i = 0; i < TM / 32; i++) {
        const int inst = lane + 32 * i;
        const long n = tile * TM + inst;
        float xa[KPAD];
        float s2 = 0.f;
#pragma unroll
        for (int d = 0; d < 9; d++) { const float v = (n < N) ? __ldg(X + n * 9 + d) * xscale : 0.f; xa[d] = v; s2 = fmaf(v, v, s2); }
        xa[9] = 1.f; xa[10] = -0.5f * s2;
#pragma unroll
        for (int d = 11; d < KPAD; d++) xa[d] = 0.f;
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; j++) { hi[j] = tf32_rna(xa[4 * kc + j]); lo[j] = tf32_rna(xa[4 * kc + j] - __uint_as_float(hi[j])); }
          *reinterpret_cast<uint4*>(A + (kc * TM + inst) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(A + HALF_BYTES + (kc * TM + inst) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(afull(ab));
    };
    long g = 0; int k = 0;
    if ((long)blockIdx.x < ntile_inst) build_a(blockIdx.x, 0);
    for (long tile = blockIdx.x; tile < ntile_inst; tile += gridDim.x, k++) {
      if (lane == 0) {
        for (int t = 0; t < ntiles; t++, g++) {
          const int s = (int)(g % STAGES);
          mbar_wait(empty(s), (uint32_t)((g / STAGES) & 1) ^ 1u);
          mbar_expect_tx(full(s), TILE_BYTES);
          bulk_g2s(smem_u32(sB + s * TILE_BYTES), tiles + (size_t)t * TILE_FLOATS, TILE_BYTES, full(s));
        }
      }
      __syncwarp();
      if (tile + gridDim.x < ntile_inst) build_a(tile + gridDim.x, k + 1);
    }
  } else if (warp == EPI_WARPS) {
    if (lane == 0) {                                         // ---- MMA issuer ----
      long g = 0; int k = 0;
      for (long tile = blockIdx.x; tile < ntile_inst; tile += gridDim.x, k++) {
        const int ab = k & 1;
        mbar_wait(afull(ab), (k >> 1) & 1);
        const uint32_t a_hi = smem_u32(sA + ab * A_BYTES), a_lo = a_hi + HALF_BYTES;
        for (int t = 0; t < ntiles; t++, g++) {
          const int s = (int)(g % STAGES), buf = (int)(g % NACC);
          mbar_wait(tempty(buf), (uint32_t)((g / NACC) & 1) ^ 1u);
          mbar_wait(full(s), (uint32_t)((g / STAGES) & 1));
          tc_fence_after();
          const uint32_t b_hi = smem_u32(sB + s * TILE_BYTES), b_lo = b_hi + HALF_BYTES;
          const uint32_t d = tmem_base + (uint32_t)buf * TN;
          const uint32_t LBO = TN * 16, SBO = 128;
#define CDS_DESC(base, ks) smem_desc((base) + (ks) * KSTEP_BYTES, LBO, SBO)
          mma_tf32(d, CDS_DESC(a_lo, 0), CDS_DESC(b_hi, 0), IDESC, 0u);   // small terms first
          mma_tf32(d, CDS_DESC(a_lo, 1), CDS_DESC(b_hi, 1), IDESC, 1u);
          mma_tf32(d, CDS_DESC(a_hi, 0), CDS_DESC(b_lo, 0), IDESC, 1u);
          mma_tf32(d, CDS_DESC(a_hi, 1), CDS_DESC(b_lo, 1), IDESC, 1u);
          mma_tf32(d, CDS_DESC(a_hi, 0), CDS_DESC(b_hi, 0), IDESC, 1u);
          mma_tf32(d, CDS_DESC(a_hi, 1), CDS_DESC(b_hi, 1), IDESC, 1u);
#undef CDS_DESC
          mma_commit(empty(s));
          mma_commit(tfull(buf));
        }
        mma_commit(aempty(ab));                              // every MMA that reads this A buffer has completed when this fires
      }
    }
  } else {                                                   // ---- epilogue warps 0-7 ----
    // Warp set s = warp / 4 takes every other SV tile (g & 1 == s), all 128 columns of it, from accumulator g % NACC; quad = warp % 4
    // is its TMEM lane quadrant.  With four accumulators the MMAs of a set's next tile are done long before the set asks for it.  The two warps that share a scheduler therefore work half a tile apart: while one waits for its next
    // accumulator (tcgen05.commit -> mbarrier -> first tcgen05.ld) the other is in the middle of its MUFU stream, so the XU pipe does
    // not idle at tile boundaries (with both warps on the SAME accumulator, one column half each, it did: 77 % of the MUFU rate).
    const int quad = warp & 3, set = warp >> 2;
    const uint32_t tquad = tmem_base + ((uint32_t)(quad * 32) << 16);
    int g = set, k = 0;                                      // g counts this CTA's SV tiles over all its instance tiles
    const int gtot = (ntile_inst > (long)blockIdx.x ? (int)((ntile_inst - 1 - blockIdx.x) / gridDim.x) + 1 : 0) * ntiles;
    uint32_t va[32], vb[32];
    auto first_chunk = [&](int gg) {                         // wait for accumulator gg % NACC and start loading its first 32 columns
      const int b = gg % NACC;
      mbar_wait(tfull(b), (uint32_t)((gg / NACC) & 1));
      tc_fence_after();
      tmem_ld32(tquad + (uint32_t)(b * TN), va);
    };
    if (g < gtot) first_chunk(g);
    for (long tile = blockIdx.x; tile < ntile_inst; tile += gridDim.x, k++) {
      float accp[4] = {0.f, 0.f, 0.f, 0.f}, accn[4] = {0.f, 0.f, 0.f, 0.f};
      const int g_end = (k + 1) * ntiles;                    // this instance tile's SV tiles are g in [k * ntiles, g_end)
      for (; g < g_end; g += 2) {
        const int t = g - k * ntiles;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        unsigned long long A01 = 0ull, A23 = 0ull;           // POLY_OF_8 < 0: the same four partial sums as two packed pairs (FADD2)
        auto consume = [&](const uint32_t (&v)[32]) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const float x = __uint_as_float(v[j + q]);
              e[q] = (q < POLY_OF_8) ? ex2_poly(x) : ex2_approx(x);
            }
            if (POLY_OF_8 < 0) {
              A01 = fadd2_pk(A01, fadd2_pk(pack2(e[0], e[1]), pack2(e[4], e[5])));
              A23 = fadd2_pk(A23, fadd2_pk(pack2(e[2], e[3]), pack2(e[6], e[7])));
            } else {
              a0 += e[0] + e[4]; a1 += e[1] + e[5]; a2 += e[2] + e[6]; a3 += e[3] + e[7];
            }
          }
        };
        const int buf = g % NACC;
        const uint32_t taddr = tquad + (uint32_t)(buf * TN);
        // register double buffering: the next 32 columns (of the set's next tile, at the end) are in flight from TMEM while the
        // MUFUs of the current ones issue
        tmem_ld_wait();
        tmem_ld32(taddr + 32, vb);
        consume(va);
        tmem_ld_wait();
        tmem_ld32(taddr + 64, va);
        consume(vb);
        tmem_ld_wait();
        tmem_ld32(taddr + 96, vb);
        consume(va);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));             // this warp's quarter of the accumulator is in registers
        if (g + 2 < gtot) first_chunk(g + 2);                // its MMAs were issued two tiles ago
        consume(vb);
        if (POLY_OF_8 < 0) {
          a0 = __uint_as_float((unsigned)A01); a1 = __uint_as_float((unsigned)(A01 >> 32));
          a2 = __uint_as_float((unsigned)A23); a3 = __uint_as_float((unsigned)(A23 >> 32));
        }
        if (t < ntiles_pos) { accp[0] += a0; accp[1] += a1; accp[2] += a2; accp[3] += a3; }
        else                { accn[0] += a0; accn[1] += a1; accn[2] += a2; accn[3] += a3; }
      }
      // the two sets' partial sums of an instance meet in shared memory (buffer k & 1: one named barrier per instance tile is enough)
      const float P = (accp[0] + accp[1]) + (accp[2] + accp[3]), Q = (accn[0] + accn[1]) + (accn[2] + accn[3]);
      float* rb = red + (k & 1) * 2 * TM;
      const int inst = quad * 32 + lane;
      if (set == 1) { rb[inst] = P; rb[TM + inst] = Q; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (set == 0) {
        const long n = tile * TM + inst;
        if (n < N) dec[n] = (P + rb[inst]) - (Q + rb[TM + inst]) - rho;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ACC_COLS) : "memory");
  }
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
  static const int impl = [] { const char* e = getenv("CDS_SVM_IMPL"); return (e && !strcmp(e, "fp32")) ? 0 : 1; }();
  if (m->nfeat == 9 && impl == 1 && m->d_tiles) {
    // share of exponentials on the FMA pipe; measured at nSV 4096: 0/8 1.40 ms, 1/8 1.28, 2/8 1.17, 3/8 1.22, 4/8 1.44 (CDS_SVM_POLY overrides)
    static const int poly = [] { const char* e = getenv("CDS_SVM_POLY"); const int v = e ? atoi(e) : 2; return v < 0 ? 0 : v > 4 ? 4 : v; }();
    auto kern = poly == 0 ? svm_rbf9_tc_kernel<0> : poly == 1 ? svm_rbf9_tc_kernel<1> : poly == 2 ? svm_rbf9_tc_kernel<2>
              : poly == 3 ? svm_rbf9_tc_kernel<3> : svm_rbf9_tc_kernel<4>;
    const int smem_bytes = tc::SMEM_BYTES;   // (padding it so that one CTA fits per SM, to leave room for overlapped kernels, measured 6 % slower)
    if (int e = ensure_dyn_smem(kern, smem_bytes)) return e;
    // CDS_SVM_PERSIST=0 selects the non-persistent kernel (two CTAs per SM, owns the whole SM); the default persistent kernel is
    // built to share the SMs with the FP32-bound contrast kernel of the next batch (pipeline.cu)
    static const int persist = [] { const char* e = getenv("CDS_SVM_PERSIST"); return (e && e[0] == '0') ? 0 : 1; }();
    if (persist) {
      // default (-1): every exponential on the MUFU and the partial sums added as packed pairs (FADD2: half the issue slots of the
      // scalar adds, same bits); 0 = scalar adds, 1..4 = that many of 8 exponentials as a polynomial on the FMA pipe (faster alone,
      // slower beside the FP32-bound contrast kernel)
      static const int polyp = [] { const char* e = getenv("CDS_SVM_POLY"); const int v = e ? atoi(e) : -1; return v < -1 ? -1 : v > 4 ? 4 : v; }();
      auto kp = polyp < 0 ? svm_rbf9_tcp_kernel<-1> : polyp == 0 ? svm_rbf9_tcp_kernel<0> : polyp == 1 ? svm_rbf9_tcp_kernel<1> : polyp == 2 ? svm_rbf9_tcp_kernel<2>
              : polyp == 3 ? svm_rbf9_tcp_kernel<3> : svm_rbf9_tcp_kernel<4>;
      if (int e = ensure_dyn_smem(kp, tcp::SMEM_BYTES)) return e;
      const long ntile_inst = (N + tc::TM - 1) / tc::TM;
      const unsigned gridp = (unsigned)std::min<long>(device_sm_count(), ntile_inst);
      kp<<<gridp, tcp::THREADS, tcp::SMEM_BYTES, st>>>(X, N, m->d_tiles, m->ntiles, m->ntiles_pos, (float)sqrt(2 * g2), (float)m->rho, dec);
      CDS_LAUNCH_CHECK();
      return CDS_OK;
    }
    const unsigned grid = (unsigned)((N + tc::TM - 1) / tc::TM);
    kern<<<grid, tc::THREADS, smem_bytes, st>>>(X, N, m->d_tiles, m->ntiles, m->ntiles_pos, (float)sqrt(2 * g2), (float)m->rho, dec);
  } else if (m->nfeat == 9) {
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
    // tensor-core path: SVs grouped by sign(coef) (zero coefficients dropped), each group padded to whole 128-SV tiles
    {
      auto tf32 = [](float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&v, &u, 4); return v; };   // cvt.rna.tf32
      std::vector<int> order;
      for (int i = 0; i < nsv; i++) if (coef[i] > 0) order.push_back(i);
      const int npos = (int)order.size(), tpos = ceil_div(npos, tc::TN);
      order.resize((size_t)tpos * tc::TN, -1);
      for (int i = 0; i < nsv; i++) if (coef[i] < 0) order.push_back(i);
      const int ttot = ceil_div((int)order.size(), tc::TN);
      order.resize((size_t)ttot * tc::TN, -1);
      std::vector<float> img((size_t)ttot * tc::TILE_FLOATS, 0.f);
      for (int t = 0; t < ttot; t++)
        for (int r = 0; r < tc::TN; r++) {
          const int i = order[(size_t)t * tc::TN + r];
          float sa[tc::KPAD] = {0};
          if (i >= 0) {
            double s2 = 0;
            for (int d = 0; d < 9; d++) { sa[d] = (float)(SV[(size_t)i * 9 + d] * sc); s2 += (double)sa[d] * sa[d]; }
            sa[9] = (float)(-0.5 * s2 + log2(fabs(coef[i])));
          } else {
            sa[9] = -1e30f;                                    // padding: 2^(-1e30) == 0
          }
          sa[10] = 1.f;
          float* base = img.data() + (size_t)t * tc::TILE_FLOATS;
          for (int k = 0; k < tc::KPAD; k++) {
            const float hi = tf32(sa[k]), lo = tf32(sa[k] - hi);
            const size_t o = ((size_t)(k / 4) * tc::TN + r) * 4 + (k % 4);
            base[o] = hi; base[(size_t)tc::KPAD * tc::TN + o] = lo;
          }
        }
      m->ntiles = ttot; m->ntiles_pos = tpos;
      if (ttot > 0) {
        CDS_CUDA(cudaMalloc(&m->d_tiles, img.size() * 4));
        CDS_CUDA(cudaMemcpy(m->d_tiles, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
      }
    }
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
      if (idx < 1 || idx > 64) {      // libsvm feature indices are 1-based; 0 marks a precomputed-kernel model, which this path does not serve
        fclose(f); set_error("cds_svm_model_load: SV %d has feature index %ld (expected 1..64)", i, idx); return CDS_EFORMAT;
      }
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
  if (m->d_tiles) cudaFree(m->d_tiles);
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
