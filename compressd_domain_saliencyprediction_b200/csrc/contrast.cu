// contrast.cu — stage (b): Gaussian-weighted centre-surround block contrast.
//
// Replaces computecontrast5.cpp:52-96 (mexFunction; img_extract :11, outer_pro :30, sumup :41):
//     out(r,c) = sum_{i,j<win} W(i,j) * g(P(r+i, c+j), P(r+len, c+len)),  g(p,ctr) = p != 0 ? |p-ctr| : 0
// on the zero-padded, pm_norm'ed map P built by Sudoku_contrastcpp5.m:18-34.
//
// B200 design: FP32 CUDA-core bound (2 FP32 lane operations per (cell, neighbour) pair:
// FADD d = p - c ; FFMA acc += |d| * w  with |.| as a free operand modifier; issued two pairs at a time
// as FADD2 / FFMA2, see sep_accumulate_packed).  The reference's W
// is a separable Gaussian (Sudoku_contrastcpp5.m:8-14), so the hot kernel accumulates each window
// row with the column factor gc[j] (a constant-bank FFMA operand: the loop over j is fully
// unrolled) and folds the row factor gr[i] in once per row.  Each thread owns CPT consecutive
// output columns so one 128-bit shared-memory load feeds 4*CPT pairs; lanes run down rows with a
// pitch chosen so the LDS.128 of a quarter-warp hit 8 distinct 16-byte bank groups.
// The p != 0 test costs one FSETP per loaded p (shared by CPT pairs) and is skipped per tile when
// the staged halo holds no zero.
#include "cds_common.cuh"
#include <math.h>
#include <algorithm>
#include <mutex>
#include <vector>

namespace cds {

constexpr int CONTRAST_SCHED_WORDS = 1 + 1024;     // [0] tile counter, [1 + smid] live CTAs per SM

template <int WIN>
struct SepWeights {
  float gr[WIN];   // factor along rows   (window row offset i)
  float gc[WIN];   // factor along columns (window col offset j)
};

template <int WIN, int CPT, int NWARPS>
struct SepCfg {
  static constexpr int TR = 16;
  static constexpr int TC = NWARPS * 2 * CPT;
  static constexpr int HR = TR + WIN - 1;
  static constexpr int NU = WIN + CPT - 1;
  static constexpr int NV4 = (NU + 3) / 4;
  static constexpr int HC = TC - CPT + 4 * NV4;            // columns the float4 walk can touch
  static constexpr int P4 = (HC + 3) / 4;
  static constexpr int PITCH = 4 * ((P4 % 2) ? P4 : P4 + 1);  // PITCH/4 odd -> conflict-free LDS.128 down rows
  static constexpr int SMEM_BYTES = HR * PITCH * 4;
  static constexpr int THREADS = NWARPS * 32;
};

template <int WIN, int CPT, bool ZTEST>
__device__ __forceinline__ void sep_accumulate(const float* __restrict__ sbase, const int pitch,
                                               const float (&c)[CPT], float (&acc)[CPT],
                                               const SepWeights<WIN>& w) {
  constexpr int NU = WIN + CPT - 1;
#pragma unroll 1
  for (int i = 0; i < WIN; i++) {
    const float* rowp = sbase + i * pitch;
    float racc[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) racc[k] = 0.f;
#pragma unroll
    for (int u4 = 0; u4 < NU; u4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(rowp + u4);
      const float pv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int u = u4 + t;
        if (u < NU) {
          const float p = pv[t];
          if (ZTEST) {
            if (p != 0.f) {
#pragma unroll
              for (int k = 0; k < CPT; k++) {
                const int j = u - k;
                if (j >= 0 && j < WIN) racc[k] = fmaf(fabsf(p - c[k]), w.gc[j], racc[k]);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < CPT; k++) {
              const int j = u - k;
              if (j >= 0 && j < WIN) racc[k] = fmaf(fabsf(p - c[k]), w.gc[j], racc[k]);
            }
          }
        }
      }
    }
    const float gi = w.gr[i];
#pragma unroll
    for (int k = 0; k < CPT; k++) acc[k] = fmaf(gi, racc[k], acc[k]);
  }
}

// ---- packed FP32 form of the same loop (sm_100 FADD2 / FFMA2: two IEEE FP32 operations per issued instruction) ----
// One packed operation serves the ordered pairs (centre 2m, neighbour u) and (centre 2m + 1, neighbour u + 1): both have the column
// offset j = u - 2m, so the weight gc[j] is ONE scalar (a uniform-register operand broadcast to both halves), the centres are a
// loop-invariant register pair and the neighbours are two adjacent values of the row.  d = p - c is one FADD2, racc += |d| * gc[j] one
// FFMA2 (|.| is an operand modifier there as well): 1 issued instruction per pair instead of 2, the same number of FP32-pipe lane
// cycles.  Standalone the kernel is bound by those lane cycles; the issue slots it gives up are what the MUFU-bound SVM kernel that
// shares the SMs with it needs (pipeline.cu).  Every centre still adds its terms in ascending u, each with one rounding for the
// product-sum: results are bitwise those of sep_accumulate.
// Zero test: the factor [p != 0] is folded into the subtraction as t = fma(-c, m, p) with m = 1 or 0 per neighbour -- p - c when
// p != 0, and p = 0 itself otherwise -- so the masked form costs one FSET per loaded neighbour and nothing per pair.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long abs2(unsigned long long v) { float a, b; upk2(v, a, b); return pk2(fabsf(a), fabsf(b)); }

template <int WIN, int CPT, bool ZTEST>
__device__ __forceinline__ void sep_accumulate_packed(const float* __restrict__ sbase, const int pitch,
                                                      const float (&c)[CPT], float (&acc)[CPT],
                                                      const SepWeights<WIN>& w) {
  static_assert(CPT % 2 == 0, "centres are processed in pairs");
  constexpr int NU = WIN + CPT - 1, NP = CPT / 2;
  unsigned long long nc2[NP];
#pragma unroll
  for (int m = 0; m < NP; m++) nc2[m] = pk2(-c[2 * m], -c[2 * m + 1]);
#pragma unroll 1
  for (int i = 0; i < WIN; i++) {
    const float* rowp = sbase + i * pitch;
    unsigned long long racc2[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) racc2[m] = 0ull;
    float p_prev = 0.f, m_prev = 0.f;                      // neighbour u4 - 1 (and its mask) of the previous 16-byte load
#pragma unroll
    for (int u4 = 0; u4 < NU; u4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(rowp + u4);
      const float pv[5] = {p_prev, v.x, v.y, v.z, v.w};
      float mk[5] = {m_prev, 1.f, 1.f, 1.f, 1.f};
      if (ZTEST) {
#pragma unroll
        for (int t = 1; t < 5; t++) mk[t] = pv[t] != 0.f ? 1.f : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 4; t++) {                        // the pair (neighbour u, neighbour u + 1), u = u4 - 1 + t
        const int u = u4 - 1 + t;
        if (u >= 0 && u + 1 < NU) {
          const unsigned long long p2 = pk2(pv[t], pv[t + 1]);
          const unsigned long long m2 = ZTEST ? pk2(mk[t], mk[t + 1]) : 0ull;
#pragma unroll
          for (int m = 0; m < NP; m++) {
            const int j = u - 2 * m;
            if (j >= 0 && j < WIN) {
              const unsigned long long d = ZTEST ? ffma2(nc2[m], m2, p2) : fadd2(p2, nc2[m]);
              const float g = w.gc[j];
              racc2[m] = ffma2(abs2(d), pk2(g, g), racc2[m]);
            }
          }
        }
      }
      p_prev = v.w; m_prev = mk[4];
    }
    const float gi = w.gr[i];
#pragma unroll
    for (int m = 0; m < NP; m++) {
      float r0, r1; upk2(racc2[m], r0, r1);
      acc[2 * m] = fmaf(gi, r0, acc[2 * m]); acc[2 * m + 1] = fmaf(gi, r1, acc[2 * m + 1]);
    }
  }
}

// Persistent: a CTA walks tiles t = blockIdx.x, + gridDim.x, ... of the (column tile, row tile, map) space, column tile fastest
// so that CTAs running together stage neighbouring halos.  The grid is (SMs x CTAs per SM) at most; the pipeline caps the CTAs
// per SM when the kernel has to leave room for the persistent SVM kernel that runs beside it (pipeline.cu, svm.cu).
// With `sched` (device words: [0] next tile, [1 + smid] live CTAs on that SM; zeroed before the launch) the tiles are handed out by
// an atomic counter and at most `cap` CTAs stay alive per SM — the launch over-subscribes every SM and the surplus CTAs retire
// at once, so the kernel occupies exactly `cap` CTA slots on every SM wherever the block scheduler put its CTAs.
// One tile per CTA when the grid covers the tile space (gridDim.x == ntiles: the default, and the fastest standalone form — the
// block scheduler's own hand-out staggers the staging phases of the CTAs that share an SM); persistent otherwise.
// 40 registers: three CTAs of this kernel then fit beside the 320-thread, 128-register SVM CTA that shares the SM with them
// (3 x 192 x 40 + 320 x 128 = 64 000 of 65 536); the packed form would take 58 without the cap and lose the third CTA.
template <int WIN, int CPT, int NWARPS, bool PACKED>
__global__ void __maxnreg__(40)
contrast_sep_kernel(const float* __restrict__ pad, float* __restrict__ out,
                    unsigned* __restrict__ out_max_bits, int R, int C, int pr, int pc, int ntx, int nty, int ntiles,
                    unsigned* __restrict__ sched, int cap, const __grid_constant__ SepWeights<WIN> w) {
  using Cfg = SepCfg<WIN, CPT, NWARPS>;
  constexpr int LEN = WIN / 2;
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_has_zero, s_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane & 15, lc = lane >> 4;
  const int cb = (warp * 2 + lc) * CPT;            // column base inside the tile
  const float* sbase = smem + lr * Cfg::PITCH + cb;
  if (sched) {
    if (threadIdx.x == 0) {
      unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
      const bool live = cap <= 0 || atomicAdd(sched + 1 + smid, 1u) < (unsigned)cap;
      s_tile = live ? (int)atomicAdd(sched, 1u) : ntiles;
    }
    __syncthreads();
  }
  for (int tile = sched ? s_tile : (int)blockIdx.x; tile < ntiles;) {
    const int tx = tile % ntx, trest = tile / ntx;
    const int tile_c0 = tx * Cfg::TC;
    const int tile_r0 = (trest % nty) * Cfg::TR;
    const int map = trest / nty;
    const float* __restrict__ P = pad + (size_t)map * pr * pc;
    if (threadIdx.x == 0) s_has_zero = 0;
    __syncthreads();                               // also: everyone has finished reading the previous tile
    // stage the halo tile (padded coordinates: output (r,c) sees padded rows r..r+WIN-1)
    int zero_seen = 0;
    for (int idx = threadIdx.x; idx < Cfg::HR * Cfg::PITCH; idx += Cfg::THREADS) {
      const int hr = idx / Cfg::PITCH, hc = idx - hr * Cfg::PITCH;
      const int gr_ = tile_r0 + hr, gc_ = tile_c0 + hc;
      float v = 1.f;   // out-of-range filler (never contributes to a stored output)
      if (gr_ < pr && gc_ < pc) { v = __ldg(P + (size_t)gr_ * pc + gc_); zero_seen |= (v == 0.f); }
      smem[idx] = v;
    }
    if (zero_seen) s_has_zero = 1;
    __syncthreads();

    float c[CPT], acc[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) { c[k] = sbase[LEN * Cfg::PITCH + LEN + k]; acc[k] = 0.f; }
    if (PACKED) {
      if (s_has_zero) sep_accumulate_packed<WIN, CPT, true>(sbase, Cfg::PITCH, c, acc, w);
      else            sep_accumulate_packed<WIN, CPT, false>(sbase, Cfg::PITCH, c, acc, w);
    } else {
      if (s_has_zero) sep_accumulate<WIN, CPT, true>(sbase, Cfg::PITCH, c, acc, w);
      else            sep_accumulate<WIN, CPT, false>(sbase, Cfg::PITCH, c, acc, w);
    }

    const int r = tile_r0 + lr, c0 = tile_c0 + cb;
    float mx = 0.f;
    if (r < R) {
      float* o = out + ((size_t)map * R + r) * C + c0;
      if (c0 + CPT <= C && (C % 4) == 0 && (CPT % 4) == 0) {
#pragma unroll
        for (int k = 0; k < CPT; k += 4)
          *reinterpret_cast<float4*>(o + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
#pragma unroll
        for (int k = 0; k < CPT; k++) mx = fmaxf(mx, acc[k]);
      } else {
#pragma unroll
        for (int k = 0; k < CPT; k++)
          if (c0 + k < C) { o[k] = acc[k]; mx = fmaxf(mx, acc[k]); }
      }
    }
    if (out_max_bits) {
      mx = warp_max(mx);
      if (lane == 0) atomic_max_nonneg(out_max_bits + map, mx);
    }
    if (sched) {
      __syncthreads();                             // everyone has read s_tile (and the shared tile) of this round
      if (threadIdx.x == 0) s_tile = (int)atomicAdd(sched, 1u);
      __syncthreads();
      tile = s_tile;
    } else tile += gridDim.x;
  }
}

// Generic fallback: arbitrary (non-separable) W, any win <= 64.  One thread per output cell,
// 16x16 outputs per block, W and the halo tile in shared memory.
__global__ void __launch_bounds__(256)
contrast_generic_kernel(const float* __restrict__ pad, const float* __restrict__ W /*[win][win] row=row offset*/,
                        float* __restrict__ out, unsigned* __restrict__ out_max_bits,
                        int R, int C, int pr, int pc, int win, int len) {
  extern __shared__ __align__(16) float smem[];
  const int HS = 16 + win - 1;
  float* sW = smem;                 // win*win
  float* sT = smem + win * win;     // HS*HS
  const int tile_c0 = blockIdx.x * 16, tile_r0 = blockIdx.y * 16, map = blockIdx.z;
  const float* __restrict__ P = pad + (size_t)map * pr * pc;
  for (int i = threadIdx.x; i < win * win; i += 256) sW[i] = W[i];
  for (int idx = threadIdx.x; idx < HS * HS; idx += 256) {
    const int hr = idx / HS, hc = idx - hr * HS;
    const int gr_ = tile_r0 + hr, gc_ = tile_c0 + hc;
    sT[idx] = (gr_ < pr && gc_ < pc) ? __ldg(P + (size_t)gr_ * pc + gc_) : 0.f;
  }
  __syncthreads();
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  const float ctr = sT[(tr + len) * HS + tc + len];
  float acc = 0.f;
  for (int i = 0; i < win; i++) {
    float racc = 0.f;
    for (int j = 0; j < win; j++) {
      const float p = sT[(tr + i) * HS + tc + j];
      if (p != 0.f) racc = fmaf(fabsf(p - ctr), sW[i * win + j], racc);
    }
    acc += racc;
  }
  const int r = tile_r0 + tr, c = tile_c0 + tc;
  float mx = 0.f;
  if (r < R && c < C) { out[((size_t)map * R + r) * C + c] = acc; mx = acc; }
  if (out_max_bits) {
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out_max_bits + map, mx);
  }
}

template <int WIN, int CPT, int NWARPS, bool PACKED>
static int launch_sep_cfg_p(const float* pad, float* out, unsigned* mx, int nmaps, int R, int C,
                          const float* gr, const float* gc, int ctas_per_sm, unsigned* sched, cudaStream_t st) {
  using Cfg = SepCfg<WIN, CPT, NWARPS>;
  SepWeights<WIN> w;
  for (int i = 0; i < WIN; i++) { w.gr[i] = gr[i]; w.gc[i] = gc[i]; }
  if (int e = ensure_dyn_smem(contrast_sep_kernel<WIN, CPT, NWARPS, PACKED>, Cfg::SMEM_BYTES)) return e;
  const int len = WIN / 2;
  const int ntx = ceil_div(C, Cfg::TC), nty = ceil_div(R, Cfg::TR);
  const long ntiles = (long)ntx * nty * nmaps;
  if (ntiles > 0x7fffffffL) { set_error("contrast: %ld tiles", ntiles); return CDS_EINVAL; }
  const int fit = std::max(1, (int)((227 * 1024) / (Cfg::SMEM_BYTES + 1024)));          // CTAs per SM by shared memory
  // default: one CTA per tile.  With a scheduler block: over-subscribe (`fit` CTAs per SM), the kernel keeps `ctas_per_sm` of them
  // alive on each SM.  With a cap but no scheduler block: a static persistent grid of that many CTAs per SM.
  const int per_sm = (!sched && ctas_per_sm > 0) ? std::min(ctas_per_sm, fit) : fit;
  const unsigned grid = (sched || ctas_per_sm > 0) ? (unsigned)std::min<long>(ntiles, (long)device_sm_count() * per_sm) : (unsigned)ntiles;
  if (sched) CDS_CUDA(cudaMemsetAsync(sched, 0, CONTRAST_SCHED_WORDS * sizeof(unsigned), st));
  contrast_sep_kernel<WIN, CPT, NWARPS, PACKED><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(
      pad, out, mx, R, C, R + 2 * len, C + 2 * len, ntx, nty, (int)ntiles, sched, ctas_per_sm, w);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// CDS_CONTRAST_PACKED=0 selects the scalar FADD + FFMA form of the inner loop (same results; kept for A/B timing)
static bool contrast_packed() {
  static const bool v = [] { const char* e = getenv("CDS_CONTRAST_PACKED"); return !(e && e[0] == '0'); }();
  return v;
}

template <int WIN, int CPT, int NWARPS>
static int launch_sep_cfg(const float* pad, float* out, unsigned* mx, int nmaps, int R, int C,
                          const float* gr, const float* gc, int ctas_per_sm, unsigned* sched, cudaStream_t st) {
  return contrast_packed() ? launch_sep_cfg_p<WIN, CPT, NWARPS, true>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st)
                           : launch_sep_cfg_p<WIN, CPT, NWARPS, false>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
}

template <int WIN, int CPT>
static int launch_sep_win(const float* pad, float* out, unsigned* mx, int nmaps, int R, int C,
                          const float* gr, const float* gc, int ctas_per_sm, unsigned* sched, cudaStream_t st) {
  // choose the block width (warps of 2*CPT columns) that wastes the fewest columns
  const int units = ceil_div(C, 2 * CPT);
  int best = 4, best_waste = 1 << 30;
  for (int nw = 8; nw >= 4; nw--) {
    const int waste = ceil_div(units, nw) * nw - units;
    if (waste < best_waste) { best_waste = waste; best = nw; }
  }
  switch (best) {
    case 4: return launch_sep_cfg<WIN, CPT, 4>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
    case 5: return launch_sep_cfg<WIN, CPT, 5>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
    case 6: return launch_sep_cfg<WIN, CPT, 6>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
    case 7: return launch_sep_cfg<WIN, CPT, 7>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
    default: return launch_sep_cfg<WIN, CPT, 8>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
  }
}

// W generic [win][win] (row offset major) on the device
int contrast_generic_launch(const float* pad, const float* dW, float* out, unsigned* mx, int nmaps,
                            int R, int C, int win, int len, cudaStream_t st) {
  if (win < 1 || win > 64) { set_error("contrast: win %d out of range 1..64", win); return CDS_EINVAL; }
  const int HS = 16 + win - 1;
  const size_t smem = (size_t)(win * win + HS * HS) * 4;
  if (int e = ensure_dyn_smem(contrast_generic_kernel, smem)) return e;
  dim3 grid(ceil_div(C, 16), ceil_div(R, 16), nmaps);
  contrast_generic_kernel<<<grid, 256, smem, st>>>(pad, dW, out, mx, R, C, R + 2 * len, C + 2 * len, win, len);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// Separable weights given as host arrays gr[win], gc[win].  dW_scratch (device, win*win floats) is
// only needed for window sizes without a specialised kernel.
// ctas_per_sm: 0 = as many persistent CTAs per SM as fit; > 0 caps them (room for a co-resident kernel).
// sched (optional, CONTRAST_SCHED_WORDS device words owned by the caller): dynamic tile hand-out and an exact per-SM cap.
int contrast_sep_launch(const float* pad, float* out, unsigned* mx, int nmaps, int R, int C, int win,
                        const float* gr, const float* gc, float* dW_scratch, cudaStream_t st, int ctas_per_sm, unsigned* sched) {
  // diagnosis: CDS_CONTRAST_CTAS=n runs every uncapped launch as a static persistent grid of n CTAs per SM (what the kernel can do at
  // the occupancy it is left with beside the SVM kernel)
  static const int diag_ctas = [] { const char* e = getenv("CDS_CONTRAST_CTAS"); return e ? atoi(e) : 0; }();
  if (!sched && ctas_per_sm == 0 && diag_ctas > 0) ctas_per_sm = diag_ctas;
  if (win == 48) return launch_sep_win<48, 8>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
  if (win == 24) return launch_sep_win<24, 8>(pad, out, mx, nmaps, R, C, gr, gc, ctas_per_sm, sched, st);
  if (!dW_scratch) { set_error("contrast: window %d needs a weight scratch buffer", win); return CDS_EINVAL; }
  std::vector<float> W((size_t)win * win);
  for (int i = 0; i < win; i++) for (int j = 0; j < win; j++) W[(size_t)i * win + j] = gr[i] * gc[j];
  CDS_CUDA(cudaMemcpyAsync(dW_scratch, W.data(), W.size() * 4, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaStreamSynchronize(st));   // W is a stack temporary
  return contrast_generic_launch(pad, dW_scratch, out, mx, nmaps, R, C, win, win / 2, st);
}

void gaussian_factor(int win, double delta, float* g) {
  const int len = win / 2;   // centre x0 = len+1 (1-based), Sudoku_contrastcpp5.m:7-10
  for (int i = 0; i < win; i++) { const double d = i - len; g[i] = (float)exp(-(d * d) / (2 * delta * delta)); }
}

// double -> float keeping "non-zero stays non-zero" (the kernel's p != 0 test must see the same
// zero set as the reference's double compare, computecontrast5.cpp:86)
__global__ void d2f_keepnz_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const double d = in[i];
    float f = (float)d;
    if (d != 0.0 && f == 0.f) f = d > 0 ? 1.4e-45f : -1.4e-45f;
    out[i] = f;
  }
}
__global__ void f2d_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

static std::mutex g_mu;
static DevBuf g_in, g_f, g_o, g_od, g_w;

}  // namespace cds

using namespace cds;

extern "C" int cds_contrast_sep_dev(const float* pad, float* out, unsigned* out_max_bits, int nmaps,
                                    int rows, int cols, int win, float delta, void* stream) {
  if (int e = ensure_device()) return e;
  if (!pad || !out || nmaps < 1 || rows < 1 || cols < 1 || win < 1 || win > 64) {
    set_error("cds_contrast_sep_dev: bad arguments"); return CDS_EINVAL;
  }
  float g[64];
  gaussian_factor(win, delta, g);
  std::lock_guard<std::mutex> lk(g_mu);
  if (int e = g_w.reserve(64 * 64 * 4)) return e;
  return contrast_sep_launch(pad, out, out_max_bits, nmaps, rows, cols, win, g, g, g_w.as<float>(), (cudaStream_t)stream, 0, nullptr);
}

extern "C" int cds_contrast5(const double* pad, int m1, int n1, const double* W, int len, int win,
                             int m, int n, double* out) {
  if (int e = ensure_device()) return e;
  if (!pad || !W || !out || m < 1 || n < 1 || win < 1 || win > 64 || len < 0 ||
      m1 < m + win - 1 || n1 < n + win - 1 || m1 != m + 2 * len || n1 != n + 2 * len) {
    set_error("cds_contrast5: bad arguments (need m1 = m + 2*len, n1 = n + 2*len, win-1 <= 2*len, win <= 64)");
    return CDS_EINVAL;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  // MATLAB column-major (m1 x n1) == row-major (n1 x m1): run on the transposed view with the
  // column-major W buffer read as [row offset = MATLAB column offset][col offset = MATLAB row offset].
  const int R = n, C = m, pr = n1, pc = m1;
  const size_t npad = (size_t)pr * pc, nout = (size_t)R * C;
  if (int e = g_in.reserve(npad * 8)) return e;
  if (int e = g_f.reserve(npad * 4)) return e;
  if (int e = g_o.reserve(nout * 4)) return e;
  if (int e = g_od.reserve(nout * 8)) return e;
  if (int e = g_w.reserve(64 * 64 * 4)) return e;
  cudaStream_t st = 0;
  CDS_CUDA(cudaMemcpyAsync(g_in.p, pad, npad * 8, cudaMemcpyHostToDevice, st));
  d2f_keepnz_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(g_in.as<double>(), g_f.as<float>(), npad);
  CDS_LAUNCH_CHECK();
  // separable?  W[a][b] * W[a0][b0] == W[a][b0] * W[a0][b]
  int a0 = 0, b0 = 0; double best = 0;
  for (int a = 0; a < win; a++) for (int b = 0; b < win; b++)
    if (fabs(W[(size_t)a * win + b]) > best) { best = fabs(W[(size_t)a * win + b]); a0 = a; b0 = b; }
  bool sep = best > 0 && (win == 48 || win == 24) && len == win / 2;
  if (sep) {
    const double w00 = W[(size_t)a0 * win + b0];
    for (int a = 0; a < win && sep; a++) for (int b = 0; b < win; b++) {
      const double lhs = W[(size_t)a * win + b] * w00, rhs = W[(size_t)a * win + b0] * W[(size_t)a0 * win + b];
      if (fabs(lhs - rhs) > 1e-12 * best * best) { sep = false; break; }
    }
  }
  int rc;
  if (sep) {
    float gr[64], gc[64];
    const double w00 = W[(size_t)a0 * win + b0];
    for (int a = 0; a < win; a++) gr[a] = (float)W[(size_t)a * win + b0];
    for (int b = 0; b < win; b++) gc[b] = (float)(W[(size_t)a0 * win + b] / w00);
    rc = contrast_sep_launch(g_f.as<float>(), g_o.as<float>(), nullptr, 1, R, C, win, gr, gc, g_w.as<float>(), st, 0, nullptr);
  } else {
    std::vector<float> Wf((size_t)win * win);
    for (size_t i = 0; i < Wf.size(); i++) Wf[i] = (float)W[i];
    CDS_CUDA(cudaMemcpyAsync(g_w.p, Wf.data(), Wf.size() * 4, cudaMemcpyHostToDevice, st));
    CDS_CUDA(cudaStreamSynchronize(st));
    rc = contrast_generic_launch(g_f.as<float>(), g_w.as<float>(), g_o.as<float>(), nullptr, 1, R, C, win, len, st);
  }
  if (rc) return rc;
  f2d_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(g_o.as<float>(), g_od.as<double>(), nout);
  CDS_LAUNCH_CHECK();
  CDS_CUDA(cudaMemcpyAsync(out, g_od.p, nout * 8, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  return CDS_OK;
}
