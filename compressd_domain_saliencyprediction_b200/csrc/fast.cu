// fast.cu — the reference's FAST version as CUDA kernels (sm_100a).
//
// Reference: HM_16.0_features/source/Lib/TLibDecoder/TDecGop.cpp:225-333 (per-slice tail), :486-656 (CameraDect, vote_alg,
// max, Umean) and salmat.cpp (ForwardFilter :39, MatrixNorm :119, MatrixNorm2 :130, MapThreshold :141, GenerateOriMap :181,
// GenerateTemporalMap :193 / _mv :242, GenerateSpatialMap :299 / _mv :361 / _mv2 :402, translateTransform :412).
//
// The reference is float32 through OpenCV 2.4.9 and its results depend on the ORDER of its float operations (several are long
// sequential float sums whose rounding error reaches 1e-3 relative).  These kernels reproduce that arithmetic exactly — every
// product and sum is rounded where the reference rounds it (__fmul_rn / __fadd_rn, never contracted to FMA) — and stay parallel:
//   * elementwise stages run one thread per cell; maxima / minima are order-free;
//   * cv::sum / cv::mean: groups of 4 consecutive floats added in float, groups accumulated in double (parallel tree; the
//     double sum's own order only matters below 1e-15);
//   * the sequential float sums of MapThreshold (salmat.cpp:155-162) and Umean (TDecGop.cpp:622-640) are emulated by ONE lane
//     per sum while the other 31 lanes of its warp select and compact the terms in raster order; maps that are replicated
//     blocks (bit map, the three full-resolution temporal maps) use a closed form for "add the same float n times" that is
//     exact inside a binade (seq_add_run), so a 1920x1080 map costs H x (selected blocks per block row) steps, not H x W;
//   * the H x W feature maps are never materialised: all nine are constant on 4x4-pixel cells (or 32 / 64-pixel blocks), so
//     the weighted sum is formed per cell and only the final Gaussian runs at pixel resolution, in the tap order of
//     cv::GaussianBlur (row pass: taps in order from zero; column pass: centre tap, then symmetric pairs).
// Data layout: raw ring [RL] of stage-(a) maps in their narrow types (int16 mv, u8 depth, u16 bytes per CTU) + 8 floats and
// 2 ints per picture; coarse temporal rings (CTU grid, 32-pixel grid); everything else is batch-local.
#include "fast.cuh"
#include <math.h>
#include <string.h>
#include <vector>

namespace cds {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr double PI_REF = 3.1415926;   // TDecGop.cpp:59

// per-phase cycle counts of block (0,0) of the latency-bound kernels (read with cds_debug_copy "fast_prof"); development aid
__device__ long long g_fast_prof[32];
#define FPROF(slot) do { if (blockIdx.x == P.prof_x && blockIdx.y == 0 && threadIdx.x == 0) { const long long _t = clock64(); g_fast_prof[slot] = _t - t_prof; t_prof = _t; } } while (0)

struct FastP {
  int prof_x;
  int w, h, h4, w4, n4, lw, lh, nctu, gw32, gh32, g32, B, PS, PT, RL;
  int ksize, half, hq, nq_h, nq_v;
  int16_t *r_mvx, *r_mvy; uint8_t* r_dep; uint16_t* r_bytes; float* r_sc; int* r_cam;
  float *r_t64, *r_t32;
  uint8_t* d_bin; float *d_sm, *d_smbit, *d_thr, *d_tacc, *d_tval, *d_sval, *d_sraw, *d_comb, *d_hrow, *d_blur;
  unsigned* d_mm;
  float *d_kh, *d_kv, *d_wt, *d_wm;
};

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ float rcp_scale(float mx) {       // Mat / max  ->  x * (float)(1. / max)   (matop.cpp operator/)
  return __double2float_rn(__ddiv_rn(1.0, (double)mx));
}
__device__ __forceinline__ float thr_apply(float v, float tprime, int mode) {   // salmat.cpp:166-173
  if (!mode) return v;
  return ((double)v >= (double)tprime) ? 1.0f : __double2float_rn(__ddiv_rn((double)v, (double)tprime));
}
__device__ __forceinline__ float sum4(float a, float b, float c, float d) {     // stat.cpp sum_: src[0]+src[1]+src[2]+src[3] in float
  return __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d);
}

template <typename T, typename Op>
__device__ T block_reduce(T v, Op op, T* scratch /*32*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL, v, o));
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T r = scratch[0];
  for (int i = 1; i < nw; i++) r = op(r, scratch[i]);
  return r;
}
struct OpMaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMaxI { __device__ int operator()(int a, int b) const { return max(a, b); } };
struct OpMinI { __device__ int operator()(int a, int b) const { return min(a, b); } };
struct OpAddI { __device__ int operator()(int a, int b) const { return a + b; } };
struct OpAddL { __device__ long long operator()(long long a, long long b) const { return a + b; } };
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };

// n sequential float additions of the same v > 0 to S >= 0, i.e. `for (i < n) S = S + v;` with round-to-nearest-even, in
// O(binade crossings) steps.  Inside a binade the rounded increment of S + v is a constant multiple q of ulp(S) (for a tie it
// is constant once S is even, which every tie result is), so k additions add k*q exactly as long as S + v stays below the next
// power of two; the first addition, ties with odd S, and the additions that cross a binade are done one at a time.
__device__ float seq_add_run(float S, float v, int n) {
  while (n > 0) {
    const float S1 = __fadd_rn(S, v);
    n--;
    if (n == 0) return S1;
    const unsigned b0 = __float_as_uint(S), b1 = __float_as_uint(S1);
    if (!(v <= S) || (b0 >> 23) != (b1 >> 23) || (b0 >> 23) == 0) { S = S1; continue; }
    const float S2 = __fadd_rn(S1, v);
    const unsigned b2 = __float_as_uint(S2);
    const unsigned q1 = b1 - b0, q2 = b2 - b1;              // same exponent: the difference of the bit patterns is the increment in ulps
    if ((b2 >> 23) != (b1 >> 23) || q1 != q2) { S = S1; continue; }
    if (q1 == 0) return S1;                                  // v < ulp/2: nothing is ever added
    const unsigned m1 = (b1 & 0x7fffffu) | 0x800000u;
    const unsigned kmax = (0xffffffu - m1) / q1;             // additions that provably stay inside the binade
    const unsigned k = min((unsigned)n, kmax);
    S = __uint_as_float(b1 + k * q1);
    n -= (int)k;
  }
  return S;
}

// lane 0 adds the `total` floats of list (zero padded to a multiple of 16) to S in order
__device__ __forceinline__ float lane0_add_list(float S, const float* list, int total) {
  const float4* l4 = reinterpret_cast<const float4*>(list);
  for (int i = 0; i < total; i += 16) {
    const float4 a = l4[i >> 2], b = l4[(i >> 2) + 1], c = l4[(i >> 2) + 2], d = l4[(i >> 2) + 3];
    S = __fadd_rn(S, a.x); S = __fadd_rn(S, a.y); S = __fadd_rn(S, a.z); S = __fadd_rn(S, a.w);
    S = __fadd_rn(S, b.x); S = __fadd_rn(S, b.y); S = __fadd_rn(S, b.z); S = __fadd_rn(S, b.w);
    S = __fadd_rn(S, c.x); S = __fadd_rn(S, c.y); S = __fadd_rn(S, c.z); S = __fadd_rn(S, c.w);
    S = __fadd_rn(S, d.x); S = __fadd_rn(S, d.y); S = __fadd_rn(S, d.z); S = __fadd_rn(S, d.w);
  }
  return S;
}

// One warp: compacts up to 4 selected terms per lane (in lane order = raster order) into list, lane 0 adds them.
__device__ __forceinline__ float warp_ordered_add(float S, unsigned sel, float t0, float t1, float t2, float t3, float* list, int* count) {
  const int lane = threadIdx.x & 31;
  const int nsel = __popc(sel);
  int inc = nsel;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
  const int total = __shfl_sync(FULL, inc, 31);
  if (total == 0) return S;
  int p = inc - nsel;
  if (sel & 1) list[p++] = t0;
  if (sel & 2) list[p++] = t1;
  if (sel & 4) list[p++] = t2;
  if (sel & 8) list[p++] = t3;
  const int padded = (total + 15) & ~15;
  for (int i = total + lane; i < padded; i += 32) list[i] = 0.f;    // S + 0 == S
  __syncwarp();
  if (lane == 0) S = lane0_add_list(S, list, total);
  __syncwarp();
  *count += total;
  return S;
}

// Umean's loops (TDecGop.cpp:622-640) over the members of the winning direction bin, raster order.  a = m * 0.25f.
// clamp = 0: sum of a;  clamp = 1: sum of ((flag && a < mean) || (!flag && a > mean)) ? a : mean.
__device__ float warp_member_sum(const uint8_t* __restrict__ bin, int best, const int16_t* __restrict__ m, int n4, int clamp,
                                 float mean, float* list) {
  const int lane = threadIdx.x & 31;
  const bool flag = mean > 0.f;
  float S = 0.f; int cnt = 0;
  for (int base = 0; base < n4; base += 128) {
    const int k = base + lane * 4;
    unsigned sel = 0; float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < n4) {
      const uchar4 bb = *reinterpret_cast<const uchar4*>(bin + k);
      const short4 mm = *reinterpret_cast<const short4*>(m + k);
      const unsigned char bs[4] = {bb.x, bb.y, bb.z, bb.w};
      const short ms[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (bs[j] == best) {
          const float a = __fmul_rn((float)ms[j], 0.25f);
          t[j] = clamp ? (((flag && a < mean) || (!flag && a > mean)) ? a : mean) : a;
          sel |= 1u << j;
        }
    }
    S = warp_ordered_add(S, sel, t[0], t[1], t[2], t[3], list, &cnt);
  }
  return __shfl_sync(FULL, S, 0);
}

// MapThreshold's first loop (salmat.cpp:155-162) on a dense float map: sum (float, raster order) and count of v > t
__device__ float warp_select_sum(const float* __restrict__ v, int n, double t, float* list, int* count) {
  const int lane = threadIdx.x & 31;
  float S = 0.f; int cnt = 0;
  for (int base = 0; base < n; base += 128) {
    const int k = base + lane * 4;
    unsigned sel = 0; float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < n) {
      q = *reinterpret_cast<const float4*>(v + k);
      sel = ((double)q.x > t ? 1u : 0u) | ((double)q.y > t ? 2u : 0u) | ((double)q.z > t ? 4u : 0u) | ((double)q.w > t ? 8u : 0u);
    }
    S = warp_ordered_add(S, sel, q.x, q.y, q.z, q.w, list, &cnt);
  }
  *count = __shfl_sync(FULL, cnt, 0);
  return __shfl_sync(FULL, S, 0);
}

// The same loop on a map of replicated blocks: block (i, j) of a gh x gw grid covers min(ms, rows - ms*i) x min(ms, cols - ms*j)
// elements of a rows x cols raster.  Lanes pick the selected blocks of a block row, lane 0 walks the element rows.
__device__ float warp_select_sum_runs(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, float* rv, int* rn,
                                      int* count) {
  const int lane = threadIdx.x & 31;
  float S = 0.f; int cnt = 0;
  for (int i = 0; i < gh; i++) {
    int L = 0;
    for (int j0 = 0; j0 < gw; j0 += 32) {
      const int j = j0 + lane;
      const float v = j < gw ? vals[i * gw + j] : 0.f;
      const bool s = j < gw && (double)v > t;
      const unsigned bal = __ballot_sync(FULL, s);
      if (s) { const int p = L + __popc(bal & ((1u << lane) - 1u)); rv[p] = v; rn[p] = min(ms, cols - ms * j); }
      L += __popc(bal);
    }
    __syncwarp();
    if (L > 0 && lane == 0) {
      const int nr = min(ms, rows - ms * i);
      int per_row = 0;
      for (int e = 0; e < L; e++) per_row += rn[e];
      for (int y = 0; y < nr; y++)
        for (int e = 0; e < L; e++) S = seq_add_run(S, rv[e], rn[e]);
      cnt += nr * per_row;
    }
    __syncwarp();
  }
  *count = __shfl_sync(FULL, cnt, 0);
  return __shfl_sync(FULL, S, 0);
}

// ------------------------------------------------------------------------------------------------ K1: CameraDect + vote_alg + Umean
constexpr int CAM_T = 512;

__device__ __forceinline__ int direction_bin(float y, float x) {          // TDecGop.cpp:584-587
  // index = ((atan2f(y, x) / PI) + 1) * 8.0 truncated; only values next to a bin edge need the exactly rounded float atan2
  const float af = atan2f(y, x);
  double tt = ((double)af / PI_REF + 1.0) * 8.0;
  if (fabs(tt - rint(tt)) < 1e-3) {
    const float a = __double2float_rn(atan2((double)y, (double)x));
    tt = ((double)a / PI_REF + 1.0) * 8.0;
  }
  int idx = (int)tt;
  return idx == 16 ? 15 : idx;
}

__global__ void __launch_bounds__(CAM_T)
fast_camera_kernel(FastP P, int first_poc, const int16_t* __restrict__ mvx, const int16_t* __restrict__ mvy,
                   const uint8_t* __restrict__ dep, const uint16_t* __restrict__ bytes, int* __restrict__ cam_out) {
  const int b = blockIdx.x, F = first_poc + b, slot = F % P.RL, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n4 = P.n4, w4 = P.w4, lw = P.lw;
  const int16_t* mx = mvx + (size_t)b * n4; const int16_t* my = mvy + (size_t)b * n4;
  const uint8_t* dp = dep + (size_t)b * n4; const uint16_t* by = bytes + (size_t)b * P.nctu;
  int16_t* rx = P.r_mvx + (size_t)slot * n4; int16_t* ry = P.r_mvy + (size_t)slot * n4;
  uint8_t* rd = P.r_dep + (size_t)slot * n4; uint16_t* rb = P.r_bytes + (size_t)slot * P.nctu;
  uint8_t* bin = P.d_bin + (size_t)b * n4;
  __shared__ int s_i[32]; __shared__ float s_f[32]; __shared__ long long s_l[32];
  __shared__ int s_hist[CAM_T / 32][16];
  __shared__ __align__(16) float s_list[2][144];
  __shared__ int s_best, s_len; __shared__ float s_mean[2], s_v[2];

  long long t_prof = clock64();
  // bytes: min / max; depth: max; copy this picture into the raw ring
  int mn = 65535, mxb = 0, mxd = 0;
  for (int i = tid; i < P.nctu; i += CAM_T) { const int v = by[i]; mn = min(mn, v); mxb = max(mxb, v); rb[i] = (uint16_t)v; }
  for (int i = tid; i < n4 / 4; i += CAM_T) {
    reinterpret_cast<int2*>(rx)[i] = reinterpret_cast<const int2*>(mx)[i];
    reinterpret_cast<int2*>(ry)[i] = reinterpret_cast<const int2*>(my)[i];
    const uchar4 d = reinterpret_cast<const uchar4*>(dp)[i];
    reinterpret_cast<uchar4*>(rd)[i] = d;
    mxd = max(mxd, max(max((int)d.x, (int)d.y), max((int)d.z, (int)d.w)));
  }
  mn = block_reduce(mn, OpMinI(), s_i); mxb = block_reduce(mxb, OpMaxI(), s_i); mxd = block_reduce(mxd, OpMaxI(), s_i);

  // background mask, BackNum, StaticNum (TDecGop.cpp:509-519): mvx^2 + mvy^2 <= 2 in pixels == <= 32 in quarter-pel units
  int back = 0, stat = 0;
  for (int k = tid; k < n4; k += CAM_T) {
    const int r = k / w4, c = k - r * w4;
    if ((int)by[(r >> 4) * lw + (c >> 4)] < mn + 20) {
      back++;
      const int x = mx[k], y = my[k];
      if (x * x + y * y <= 32) stat++;
    }
  }
  back = block_reduce(back, OpAddI(), s_i); stat = block_reduce(stat, OpAddI(), s_i);
  const bool moving = stat <= back / 2;
  float vx = 0.f, vy = 0.f;
  FPROF(0);
  if (moving) {
    // 16-bin direction histogram of the background cells (vote_alg :580-592)
    int mycount = 0;
    for (int k0 = wid * 32; k0 < n4; k0 += CAM_T) {
      const int k = k0 + lane;
      int bi = 255;
      if (k < n4) {
        const int r = k / w4, c = k - r * w4;
        if ((int)by[(r >> 4) * lw + (c >> 4)] < mn + 20) bi = direction_bin(__fmul_rn((float)my[k], 0.25f), __fmul_rn((float)mx[k], 0.25f));
        bin[k] = (uint8_t)bi;
      }
#pragma unroll
      for (int j = 0; j < 16; j++) { const unsigned m = __ballot_sync(FULL, bi == j); if (lane == j) mycount += __popc(m); }
    }
    if (lane < 16) s_hist[wid][lane] = mycount;
    __syncthreads();
    FPROF(1);
    if (tid == 0) {
      int best = 0, bc = -1;
      for (int j = 0; j < 16; j++) { int c = 0; for (int w = 0; w < CAM_T / 32; w++) c += s_hist[w][j]; if (c > bc) { bc = c; best = j; } }   // :600 first maximum
      s_best = best; s_len = bc;
    }
    __syncthreads();
    const int best = s_best, len = s_len;
    // first loop of Umean: exact in float whenever every partial sum is (sum of |m| < 2^24 quarter-pels)
    long long sx = 0, sy = 0, ax = 0, ay = 0;
    for (int k = tid; k < n4; k += CAM_T)
      if (bin[k] == best) { const int x = mx[k], y = my[k]; sx += x; sy += y; ax += abs(x); ay += abs(y); }
    sx = block_reduce(sx, OpAddL(), s_l); sy = block_reduce(sy, OpAddL(), s_l);
    ax = block_reduce(ax, OpAddL(), s_l); ay = block_reduce(ay, OpAddL(), s_l);
    if (wid < 2) {
      const int16_t* m = wid == 0 ? mx : my;
      const long long s = wid == 0 ? sx : sy, a = wid == 0 ? ax : ay;
      float sum = (a < (1ll << 24)) ? __fmul_rn((float)s, 0.25f) : warp_member_sum(bin, best, m, n4, 0, 0.f, s_list[wid]);
      const float mean = __fdiv_rn(sum, (float)len);
      sum = warp_member_sum(bin, best, m, n4, 1, mean, s_list[wid]);
      if (lane == 0) s_v[wid] = __fdiv_rn(sum, (float)len);
    }
    __syncthreads();
    vx = s_v[0]; vy = s_v[1];
    FPROF(2);
  }
  const float nx = moving ? -vx : -0.0f, ny = moving ? -vy : -0.0f;
  // m_norm (CalDist :552) of the compensated field: its maximum feeds MatrixNorm (TDecGop.cpp:294)
  float mxn = 0.f;
  for (int k = tid; k < n4; k += CAM_T) {
    const float x = __fadd_rn(__fmul_rn((float)mx[k], 0.25f), nx), y = __fadd_rn(__fmul_rn((float)my[k], 0.25f), ny);
    mxn = fmaxf(mxn, __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));
  }
  mxn = block_reduce(mxn, OpMaxF(), s_f);
  FPROF(3);
  if (tid == 0) {
    float* sc = P.r_sc + (size_t)slot * 8;
    sc[0] = nx; sc[1] = ny;
    sc[2] = mxb > 0 ? rcp_scale((float)mxb) : 1.f;
    sc[3] = mxd > 0 ? rcp_scale((float)mxd) : 1.f;
    sc[4] = mxn > 0.f ? rcp_scale(mxn) : 1.f;
    const int cx = moving ? __float2int_rn(vx) : 0, cy = moving ? __float2int_rn(vy) : 0;      // cvRound
    P.r_cam[slot * 2] = cx; P.r_cam[slot * 2 + 1] = cy;
    if (cam_out) { cam_out[b * 2] = cx; cam_out[b * 2 + 1] = cy; }
  }
}

// ------------------------------------------------------------------------------------------------ K2: ForwardFilter (salmat.cpp:39-55)
// tempmat = tempmat + SmoothFrame[i] / cnt for ring slots i = 0.. in SLOT order: fl(fl(x * (float)(1./cnt)) + acc)
__global__ void __launch_bounds__(256) fast_smooth_kernel(FastP P, int first_poc) {
  const int b = blockIdx.y, F = first_poc + b, k = blockIdx.x * 256 + threadIdx.x;
  const int n4 = P.n4, PS = P.PS;
  if (k >= n4 + P.nctu) return;
  const int cnt = F < PS ? F + 1 : PS;
  const float rc = __double2float_rn(__ddiv_rn(1.0, (double)cnt));
  float ad = 0.f, an = 0.f, ax = 0.f, ay = 0.f;
  for (int i = 0; i < cnt; i++) {
    const int f = F < PS ? i : F - ((F - i) % PS);        // the picture held by ring slot i
    const int slot = f % P.RL;
    const float* sc = P.r_sc + (size_t)slot * 8;
    if (k < n4) {
      const size_t o = (size_t)slot * n4 + k;
      const float x = __fadd_rn(__fmul_rn((float)P.r_mvx[o], 0.25f), sc[0]), y = __fadd_rn(__fmul_rn((float)P.r_mvy[o], 0.25f), sc[1]);
      const float nr = __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))), sc[4]);
      const float d = __fmul_rn((float)P.r_dep[o], sc[3]);
      ad = __fadd_rn(__fmul_rn(d, rc), ad); an = __fadd_rn(__fmul_rn(nr, rc), an);
      ax = __fadd_rn(__fmul_rn(x, rc), ax); ay = __fadd_rn(__fmul_rn(y, rc), ay);
    } else {
      const float v = __fmul_rn((float)P.r_bytes[(size_t)slot * P.nctu + (k - n4)], sc[2]);
      ad = __fadd_rn(__fmul_rn(v, rc), ad);
    }
  }
  const int ts = F % P.RL;
  if (k < n4) {
    float* sm = P.d_sm + (size_t)b * 4 * n4;
    sm[k] = ad; sm[n4 + k] = an; sm[2 * (size_t)n4 + k] = ax; sm[3 * (size_t)n4 + k] = ay;
    const int r = k / P.w4, c = k - r * P.w4;
    if (!(r & 7) && !(c & 7)) {                           // MapDownSample by MinSize/4 = 8 (salmat.cpp:205)
      float* t = P.r_t32 + (size_t)ts * 3 * P.g32 + (r >> 3) * P.gw32 + (c >> 3);
      t[0] = ad; t[P.g32] = ax; t[2 * P.g32] = ay;
    }
  } else {
    P.d_smbit[(size_t)b * P.nctu + (k - n4)] = ad;
    P.r_t64[(size_t)ts * P.nctu + (k - n4)] = ad;
  }
}

// ------------------------------------------------------------------------------------------------ K3: MapThreshold parameters
// job (map, picture): map 0 bit (CTU grid replicated x16), 1 depth, 2 mv norm, 3 mv x, 4 mv y.  Output d_thr[b][map] =
// {t', mode (0: src.copyTo(dst)), max of the map, scale of the following MatrixNorm of thr(map)}
__global__ void __launch_bounds__(256) fast_thr_kernel(FastP P) {
  const int map = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int n4 = P.n4;
  __shared__ double s_d[32]; __shared__ float s_f[32];
  __shared__ __align__(16) float s_list[144]; __shared__ float s_rv[128]; __shared__ int s_rn[128];
  const float* bs = P.d_smbit + (size_t)b * P.nctu;
  const float* v = P.d_sm + ((size_t)b * 4 + (map > 0 ? map - 1 : 0)) * n4;
  float mx = -INFINITY; double sum = 0.0;
  long long t_prof = clock64();
  if (map == 0) {
    for (int i = tid; i < P.nctu; i += 256) {
      const float x = bs[i];
      const int rows = min(16, P.h4 - 16 * (i / P.lw)), cols = min(16, P.w4 - 16 * (i % P.lw));
      sum += (double)sum4(x, x, x, x) * (double)(rows * cols / 4);
      mx = fmaxf(mx, x);
    }
  } else {
    for (int i = tid; i < n4 / 4; i += 256) {
      const float4 q = reinterpret_cast<const float4*>(v)[i];
      sum += (double)sum4(q.x, q.y, q.z, q.w);
      mx = fmaxf(mx, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
    }
  }
  sum = block_reduce(sum, OpAddD(), s_d); mx = block_reduce(mx, OpMaxF(), s_f);
  const double mean = __dmul_rn(sum, __ddiv_rn(1.0, (double)n4));       // cv::mean: sum * (1. / N)
  const double t = (double)mx - mean;
  int mode = !(t == 0.0 || t >= (double)mx);
  float tp = 0.f;
  FPROF(8);
  if (mode && tid < 32) {
    int count = 0;
    const float S = map == 0 ? warp_select_sum_runs(bs, P.lh, P.lw, P.h4, P.w4, 16, t, s_rv, s_rn, &count)
                             : warp_select_sum(v, n4, t, s_list, &count);
    tp = __fdiv_rn(S, (float)count);
    if (tp == 0.f) mode = 0;
    if (blockIdx.x == P.prof_x && blockIdx.y == 0 && tid == 0) g_fast_prof[10] = count;
  }
  FPROF(9);
  if (tid == 0) {
    float* o = P.d_thr + ((size_t)b * 5 + map) * 4;
    const float mt = thr_apply(mx, tp, mode);
    o[0] = tp; o[1] = (float)mode; o[2] = mx; o[3] = mt > 0.f ? rcp_scale(mt) : 1.f;
  }
}

// ------------------------------------------------------------------------------------------------ K5: GenerateTemporalMap accumulations
// job (map, picture): map 0 bit (CTU grid, MinSize 64, delta 46), 1 depth (32-pixel grid, 46), 2 mv x/y pair (32, 26)
__global__ void __launch_bounds__(128) fast_temporal_kernel(FastP P, int first_poc) {
  const int map = blockIdx.y, b = blockIdx.z, idx = blockIdx.x * 128 + threadIdx.x, F = first_poc + b;
  const int gw = map == 0 ? P.lw : P.gw32, gh = map == 0 ? P.lh : P.gh32, ms = map == 0 ? 64 : 32, g = gw * gh;
  if (idx >= g) return;
  const int PT = P.PT, RL = P.RL;
  const int last = F < PT ? (map == 2 ? F - 1 : F) : PT - 1;        // salmat.cpp:210 / :262 / :222
  const int r = idx / gw, c = idx - r * gw;
  const size_t stride = map == 0 ? (size_t)P.nctu : (size_t)3 * P.g32;
  const float* ring = map == 0 ? P.r_t64 : P.r_t32 + (map == 1 ? 0 : P.g32);
  const float Sx = ring[(size_t)(F % RL) * stride + idx];
  const float Sy = map == 2 ? ring[(size_t)(F % RL) * stride + P.g32 + idx] : 0.f;
  const float* wt = P.d_wt + (map == 2 ? 0 : PT);
  int sumx = 0, sumy = 0; float acc = 0.f;
  for (int i = 1; i <= last; i++) {
    const int f = F - i, slot = f % RL;
    sumx += P.r_cam[slot * 2]; sumy += P.r_cam[slot * 2 + 1];
    const int dx = sumx / ms, dy = sumy / ms;                        // cvRound(int / int)
    const int sr = r - dy, scn = c - dx;                             // translateTransform: dst(x + dx, y + dy) = src(x, y), zero fill
    const bool in = sr >= 0 && sr < gh && scn >= 0 && scn < gw;
    const size_t o = (size_t)slot * stride + (in ? sr * gw + scn : 0);
    float d;
    if (map == 2) {
      const float px = in ? ring[o] : 0.f, py = in ? ring[o + P.g32] : 0.f;
      const float ax = __fsub_rn(px, Sx), ay = __fsub_rn(py, Sy);
      d = __fsqrt_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
    } else {
      d = fabsf(__fsub_rn(in ? ring[o] : 0.f, Sx));
    }
    acc = __fadd_rn(__fmul_rn(d, wt[i]), acc);
  }
  P.d_tacc[((size_t)b * 3 + map) * P.g32 + idx] = acc;
}

// ------------------------------------------------------------------------------------------------ K6: MatrixNorm + MapThreshold of the
// full-resolution temporal map (salmat.cpp:233-235), evaluated on the block grid with multiplicities
__global__ void __launch_bounds__(128) fast_tthr_kernel(FastP P) {
  const int map = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int gw = map == 0 ? P.lw : P.gw32, gh = map == 0 ? P.lh : P.gh32, ms = map == 0 ? 64 : 32, g = gw * gh;
  extern __shared__ float s_v1[];
  __shared__ double s_d[32]; __shared__ float s_f[32]; __shared__ float s_rv[128]; __shared__ int s_rn[128];
  const float* acc = P.d_tacc + ((size_t)b * 3 + map) * P.g32;
  float mxa = 0.f;
  for (int i = tid; i < g; i += 128) mxa = fmaxf(mxa, acc[i]);
  mxa = block_reduce(mxa, OpMaxF(), s_f);
  const float s1 = mxa > 0.f ? rcp_scale(mxa) : 1.f;
  double sum = 0.0;
  for (int i = tid; i < g; i += 128) {
    const float x = mxa > 0.f ? __fmul_rn(acc[i], s1) : acc[i];
    s_v1[i] = x;
    const int rows = min(ms, P.h - ms * (i / gw)), cols = min(ms, P.w - ms * (i % gw));
    sum += (double)sum4(x, x, x, x) * (double)(rows * cols / 4);
  }
  sum = block_reduce(sum, OpAddD(), s_d);
  const float mx = mxa > 0.f ? __fmul_rn(mxa, s1) : mxa;
  const double mean = __dmul_rn(sum, __ddiv_rn(1.0, (double)P.h * (double)P.w));
  const double t = (double)mx - mean;
  int mode = !(t == 0.0 || t >= (double)mx);
  float tp = 0.f;
  if (mode && tid < 32) {
    int count = 0;
    const float S = warp_select_sum_runs(s_v1, gh, gw, P.h, P.w, ms, t, s_rv, s_rn, &count);
    tp = __fdiv_rn(S, (float)count);
    if (tp == 0.f) mode = 0;
  }
  __shared__ float s_tp; __shared__ int s_mode;
  if (tid == 0) { s_tp = tp; s_mode = mode; }
  __syncthreads();
  float* out = P.d_tval + ((size_t)b * 3 + map) * P.g32;
  for (int i = tid; i < g; i += 128) out[i] = thr_apply(s_v1[i], s_tp, s_mode);
}

// ------------------------------------------------------------------------------------------------ K7: GenerateSpatialMap (salmat.cpp:299-400)
// job (kind, picture): kind 0 bit (5x5, delta 3, CTU grid), 1 depth (12x12, delta 13), 2 mv x and y (12x12, delta 11) + _mv2
__global__ void __launch_bounds__(256) fast_spatial_kernel(FastP P) {
  const int kind = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int W = kind == 0 ? 5 : 12, len = W / 2, ms4 = kind == 0 ? 16 : 8;
  const int rn = kind == 0 ? P.lh : P.gh32, cn = kind == 0 ? P.lw : P.gw32, g = rn * cn;
  const int pr = rn + 2 * len, pc = cn + 2 * len;
  extern __shared__ float s_P[];                         // padded map pr x pc, then the window weights
  float* s_w = s_P + pr * pc;
  __shared__ float s_f[32];
  const float* wm = P.d_wm + (kind == 0 ? 0 : kind == 1 ? 25 : 25 + 144);
  for (int i = tid; i < W * W; i += 256) s_w[i] = wm[i];
  float* raw = P.d_sraw + ((size_t)b * 4 + (kind == 2 ? 2 : kind)) * P.g32;     // [b][bit | depth | mv x | mv y]
  for (int comp = 0; comp < (kind == 2 ? 2 : 1); comp++) {
    const int map = kind == 0 ? 0 : kind == 1 ? 1 : 3 + comp;
    const float* th = P.d_thr + ((size_t)b * 5 + map) * 4;
    const float tp = th[0]; const int mode = (int)th[1];
    const float* cur = kind == 0 ? P.d_smbit + (size_t)b * P.nctu : P.d_sm + ((size_t)b * 4 + map - 1) * P.n4;
    // S = down(thr(cur)); zero pad; MatrixNorm2 of the padded map
    float mn = 0.f, mx = 0.f;                            // the zero border takes part in the min / max
    __syncthreads();
    for (int i = tid; i < pr * pc; i += 256) s_P[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < g; i += 256) {
      const int r = i / cn, c = i - r * cn;
      const float v = thr_apply(kind == 0 ? cur[i] : cur[(size_t)(r * ms4) * P.w4 + c * ms4], tp, mode);
      s_P[(r + len) * pc + c + len] = v;
      mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    mn = block_reduce(mn, OpMinF(), s_f); mx = block_reduce(mx, OpMaxF(), s_f);
    const float neg = -mn;                               // dst = src - min  ->  fl(src + (float)(-min))
    const float mx2 = __fadd_rn(mx, neg);                // max of the shifted map (the shift is monotonic)
    const float s = mx2 > 0.f ? rcp_scale(mx2) : 1.f;
    for (int i = tid; i < pr * pc; i += 256) { const float x = __fadd_rn(s_P[i], neg); s_P[i] = mx2 > 0.f ? __fmul_rn(x, s) : x; }
    __syncthreads();
    // out(i,j) = sum(|window - centre| .* WeightMap): float products, cv::sum order (4 floats in float, groups in double)
    float omax = 0.f;
    for (int i = tid; i < g; i += 256) {
      const int r = i / cn, c = i - r * cn;
      const float ctr = s_P[(r + len) * pc + c + len];
      double acc = 0.0; float grp = 0.f; int q = 0;
      const int nfull = (W * W) & ~3;
      for (int e = 0; e < W * W; e++) {
        const int u = e / W, v = e - u * W;
        const float p = __fmul_rn(fabsf(__fsub_rn(s_P[(r + u) * pc + c + v], ctr)), s_w[e]);
        if (e < nfull) {
          grp = q == 0 ? p : __fadd_rn(grp, p);
          if (++q == 4) { acc += (double)grp; q = 0; }
        } else acc += (double)p;
      }
      const float o = __double2float_rn(acc);
      raw[(size_t)comp * P.g32 + i] = o;
      omax = fmaxf(omax, o);
    }
    omax = block_reduce(omax, OpMaxF(), s_f);
    const float s1 = omax > 0.f ? rcp_scale(omax) : 1.f;                 // MatrixNorm(TempOut)
    if (kind < 2) {
      const float m2 = omax > 0.f ? __fmul_rn(omax, s1) : omax;          // MatrixNorm(TempmatL, SpatialMap) after the replication
      const float s2 = m2 > 0.f ? rcp_scale(m2) : 1.f;
      float* out = P.d_sval + ((size_t)b * 3 + (kind == 0 ? 1 : 2)) * P.g32;
      for (int i = tid; i < g; i += 256) {
        float o = raw[i];
        if (omax > 0.f) o = __fmul_rn(o, s1);
        if (m2 > 0.f) o = __fmul_rn(o, s2);
        out[i] = o;
      }
    } else {
      for (int i = tid; i < g; i += 256) if (omax > 0.f) raw[(size_t)comp * P.g32 + i] = __fmul_rn(raw[(size_t)comp * P.g32 + i], s1);
    }
  }
  if (kind == 2) {                                       // GenerateSpatialMap_mv2: MatrixNorm(sqrt(x.^2 + y.^2))
    __syncthreads();
    float hm = 0.f;
    for (int i = tid; i < g; i += 256) {
      const float x = raw[i], y = raw[P.g32 + i];
      hm = fmaxf(hm, __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));
    }
    hm = block_reduce(hm, OpMaxF(), s_f);
    const float s = hm > 0.f ? rcp_scale(hm) : 1.f;
    float* out = P.d_sval + (size_t)b * 3 * P.g32;
    for (int i = tid; i < g; i += 256) {
      const float x = raw[i], y = raw[P.g32 + i];
      const float hv = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
      out[i] = hm > 0.f ? __fmul_rn(hv, s) : hv;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K8: weighted sum of the nine maps per cell
// FinalMap = A*w0 + B*w1 (addWeighted) then + C*w2 ... (scaleAdd), TDecGop.cpp:319
__constant__ float c_fw[9] = {0.058f, 0.171f, -0.095f, 0.068f, -0.01f, 0.509f, -0.051f, 0.154f, 0.196f};   // TDecGop.cpp:94

__global__ void __launch_bounds__(256) fast_combine_kernel(FastP P) {
  const int b = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
  if (k >= P.n4) return;
  const int r = k / P.w4, c = k - r * P.w4;
  const int ctu = (r >> 4) * P.lw + (c >> 4), c32 = (r >> 3) * P.gw32 + (c >> 3);
  const float* th = P.d_thr + (size_t)b * 20;
  const float* sm = P.d_sm + (size_t)b * 4 * P.n4;
  const float* tv = P.d_tval + (size_t)b * 3 * P.g32;
  const float* sv = P.d_sval + (size_t)b * 3 * P.g32;
  float m[9];
  m[0] = __fmul_rn(thr_apply(sm[P.n4 + k], th[8], (int)th[9]), th[11]);                 // MotinVectorNorm.OriMap
  m[1] = tv[2 * P.g32 + c32];                                                            // MvTemporal
  m[2] = sv[c32];                                                                        // MvSpatial
  m[3] = __fmul_rn(thr_apply(P.d_smbit[(size_t)b * P.nctu + ctu], th[0], (int)th[1]), th[3]);
  m[4] = tv[ctu];
  m[5] = sv[P.g32 + ctu];
  m[6] = __fmul_rn(thr_apply(sm[k], th[4], (int)th[5]), th[7]);
  m[7] = tv[P.g32 + c32];
  m[8] = sv[2 * P.g32 + c32];
  float s = __fadd_rn(__fmul_rn(m[0], c_fw[0]), __fmul_rn(m[1], c_fw[1]));
#pragma unroll
  for (int i = 2; i < 9; i++) s = __fadd_rn(__fmul_rn(m[i], c_fw[i]), s);
  P.d_comb[(size_t)b * P.n4 + k] = s;
}

// ------------------------------------------------------------------------------------------------ K9 / K10: cv::GaussianBlur(32F, BORDER_CONSTANT)
// Row pass (filter.cpp RowFilter): s = 0; s = fl(s + fl(x[t] * k[t])) for t = 0..K-1.  The input is constant on 4-pixel cells, so a
// thread computes the 4 pixels of one cell; the taps are padded in front with zeros (exact no-ops) so that the anchor is a multiple
// of 4 and the cell index of every (pixel, tap) pair is static.
__global__ void __launch_bounds__(256) fast_blur_h_kernel(FastP P) {
  const int r = blockIdx.x, b = blockIdx.y;
  extern __shared__ float s_row[];                       // hq zeros | w4 cells | nq_h + 1 zeros
  const int w4 = P.w4, hq = P.hq, nq = P.nq_h;
  const float* in = P.d_comb + ((size_t)b * P.h4 + r) * w4;
  for (int i = threadIdx.x; i < w4 + hq + nq + 2; i += 256) { const int c = i - hq; s_row[i] = (c >= 0 && c < w4) ? in[c] : 0.f; }
  __syncthreads();
  const float4* k4 = reinterpret_cast<const float4*>(P.d_kh);
  for (int p = threadIdx.x; p < w4; p += 256) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    float c0 = s_row[p];                                 // cell p - hq
    for (int m = 0; m < nq; m++) {
      const float c1 = s_row[p + m + 1];
      const float4 k = __ldg(k4 + m);
      // pixel j, tap 4m + i reads cell m + (j + i) / 4
      s0 = __fadd_rn(s0, __fmul_rn(c0, k.x)); s0 = __fadd_rn(s0, __fmul_rn(c0, k.y)); s0 = __fadd_rn(s0, __fmul_rn(c0, k.z)); s0 = __fadd_rn(s0, __fmul_rn(c0, k.w));
      s1 = __fadd_rn(s1, __fmul_rn(c0, k.x)); s1 = __fadd_rn(s1, __fmul_rn(c0, k.y)); s1 = __fadd_rn(s1, __fmul_rn(c0, k.z)); s1 = __fadd_rn(s1, __fmul_rn(c1, k.w));
      s2 = __fadd_rn(s2, __fmul_rn(c0, k.x)); s2 = __fadd_rn(s2, __fmul_rn(c0, k.y)); s2 = __fadd_rn(s2, __fmul_rn(c1, k.z)); s2 = __fadd_rn(s2, __fmul_rn(c1, k.w));
      s3 = __fadd_rn(s3, __fmul_rn(c0, k.x)); s3 = __fadd_rn(s3, __fmul_rn(c1, k.y)); s3 = __fadd_rn(s3, __fmul_rn(c1, k.z)); s3 = __fadd_rn(s3, __fmul_rn(c1, k.w));
      c0 = c1;
    }
    reinterpret_cast<float4*>(P.d_hrow + ((size_t)b * P.h4 + r) * P.w)[p] = make_float4(s0, s1, s2, s3);
  }
}

// Column pass (filter.cpp SymmColumnFilter): s = fl(x[0] * k[0]); s = fl(s + fl(fl(x[+t] + x[-t]) * k[t])), t = 1..half.  The row-pass
// output is constant over the 4 pixel rows of a cell row; a thread computes the 4 pixel rows of one cell row q for one column.
__global__ void __launch_bounds__(128) fast_blur_v_kernel(FastP P) {
  const int x = blockIdx.x * 128 + threadIdx.x, q = blockIdx.y, b = blockIdx.z;
  const int W = P.w, h4 = P.h4;
  __shared__ float s_f[32];
  float lo = INFINITY, hi = -INFINITY;
  if (x < W) {
    const float* hr = P.d_hrow + (size_t)b * h4 * W + x;
    const float4* k4 = reinterpret_cast<const float4*>(P.d_kv + 4);
    const float ctr = hr[(size_t)q * W];
    const float k0 = __ldg(P.d_kv);
    float s0 = __fmul_rn(ctr, k0), s1 = s0, s2 = s0, s3 = s0;
    float a0 = ctr, b0 = ctr;
    for (int m = 0; m < P.nq_v; m++) {
      const float a1 = (q + m + 1 < h4) ? hr[(size_t)(q + m + 1) * W] : 0.f;
      const float b1 = (q - m - 1 >= 0) ? hr[(size_t)(q - m - 1) * W] : 0.f;
      const float4 k = __ldg(k4 + m);
      // pixel row j, tap t = 4m + i (i = 1..4): above cell m + (j + i) / 4, below cell -(m + (i - j + 3) / 4)
      s0 = __fadd_rn(s0, __fmul_rn(__fadd_rn(a0, b1), k.x)); s0 = __fadd_rn(s0, __fmul_rn(__fadd_rn(a0, b1), k.y));
      s0 = __fadd_rn(s0, __fmul_rn(__fadd_rn(a0, b1), k.z)); s0 = __fadd_rn(s0, __fmul_rn(__fadd_rn(a1, b1), k.w));
      s1 = __fadd_rn(s1, __fmul_rn(__fadd_rn(a0, b0), k.x)); s1 = __fadd_rn(s1, __fmul_rn(__fadd_rn(a0, b1), k.y));
      s1 = __fadd_rn(s1, __fmul_rn(__fadd_rn(a1, b1), k.z)); s1 = __fadd_rn(s1, __fmul_rn(__fadd_rn(a1, b1), k.w));
      s2 = __fadd_rn(s2, __fmul_rn(__fadd_rn(a0, b0), k.x)); s2 = __fadd_rn(s2, __fmul_rn(__fadd_rn(a1, b0), k.y));
      s2 = __fadd_rn(s2, __fmul_rn(__fadd_rn(a1, b1), k.z)); s2 = __fadd_rn(s2, __fmul_rn(__fadd_rn(a1, b1), k.w));
      s3 = __fadd_rn(s3, __fmul_rn(__fadd_rn(a1, b0), k.x)); s3 = __fadd_rn(s3, __fmul_rn(__fadd_rn(a1, b0), k.y));
      s3 = __fadd_rn(s3, __fmul_rn(__fadd_rn(a1, b0), k.z)); s3 = __fadd_rn(s3, __fmul_rn(__fadd_rn(a1, b1), k.w));
      a0 = a1; b0 = b1;
    }
    float* out = P.d_blur + ((size_t)b * P.h + 4 * q) * W + x;
    out[0] = s0; out[W] = s1; out[2 * (size_t)W] = s2; out[3 * (size_t)W] = s3;
    lo = fminf(fminf(s0, s1), fminf(s2, s3)); hi = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
  }
  lo = block_reduce(lo, OpMinF(), s_f); hi = block_reduce(hi, OpMaxF(), s_f);
  if (threadIdx.x == 0) { atomicMin(P.d_mm + b * 2, float_key(lo)); atomicMax(P.d_mm + b * 2 + 1, float_key(hi)); }
}

// ------------------------------------------------------------------------------------------------ K11: MatrixNorm2, x255, convertTo(CV_8U)
template <typename OutT>
__global__ void __launch_bounds__(256) fast_output_kernel(FastP P, OutT* __restrict__ out) {
  const int b = blockIdx.y; const size_t HW = (size_t)P.h * P.w;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i * 4 >= HW) return;
  const float mn = key_float(P.d_mm[b * 2]), mx = key_float(P.d_mm[b * 2 + 1]);
  const float neg = -mn, mx2 = __fadd_rn(mx, neg);
  const float s = mx2 > 0.f ? rcp_scale(mx2) : 1.f;
  const float4 q = reinterpret_cast<const float4*>(P.d_blur + (size_t)b * HW)[i];
  float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; j++) { v[j] = __fadd_rn(v[j], neg); if (mx2 > 0.f) v[j] = __fmul_rn(v[j], s); }
  if constexpr (sizeof(OutT) == 1) {
    uchar4 o;
    o.x = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(v[0], 255.f))));
    o.y = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(v[1], 255.f))));
    o.z = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(v[2], 255.f))));
    o.w = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(v[3], 255.f))));
    reinterpret_cast<uchar4*>(out + (size_t)b * HW)[i] = o;
  } else {
    reinterpret_cast<float4*>(out + (size_t)b * HW)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void fast_mm_reset_kernel(unsigned* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mm[2 * i] = 0xffffffffu; mm[2 * i + 1] = 0u; }
}

int cv_round_host(double v) { return (int)nearbyint(v); }

// cv::getGaussianKernel(n, sigma, CV_32F) of OpenCV 2.4.9 (smooth.cpp): taps rounded to float before they are summed
void cv_gaussian_kernel_f32(int n, double sigma, std::vector<float>* k) {
  k->assign(n, 0.f);
  const double sx = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8, s2 = -0.5 / (sx * sx);
  double sum = 0;
  for (int i = 0; i < n; i++) { const double x = i - (n - 1) * 0.5; (*k)[i] = (float)exp(s2 * x * x); sum += (*k)[i]; }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) (*k)[i] = (float)((*k)[i] * sum);
}

}  // namespace

struct FastState {
  FastP p;
  std::vector<void*> allocs;
  size_t smem_tthr = 0, smem_sp[3] = {0, 0, 0}, smem_h = 0;
};

namespace {
template <typename T>
int falloc(FastState* s, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) { set_error("fast: cudaMalloc(%zu): %s", count * sizeof(T), cudaGetErrorString(e)); return CDS_ENOMEM; }
  s->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return CDS_OK;
}
}  // namespace

int fast_create(int w, int h, double fps, int B, FastState** out) {
  if ((w % 16) || (h % 4) || w < 64 || h < 64) { set_error("fast version: %dx%d unsupported (width multiple of 16, height of 4)", w, h); return CDS_EINVAL; }
  FastState* s = new FastState();
  FastP& p = s->p;
  memset(&p, 0, sizeof(p));
  p.w = w; p.h = h; p.h4 = h / 4; p.w4 = w / 4; p.n4 = p.h4 * p.w4;
  p.lw = (w + 63) / 64; p.lh = (h + 63) / 64; p.nctu = p.lw * p.lh;
  p.gw32 = (w + 31) / 32; p.gh32 = (h + 31) / 32; p.g32 = p.gw32 * p.gh32;
  p.B = B;
  { const char* e = getenv("CDS_FAST_PROF"); p.prof_x = e ? atoi(e) : -1; }
  p.PS = std::max(1, cv_round_host(fps * 0.3)); p.PT = std::max(1, cv_round_host(fps * 0.7));     // salmat.cpp:18-19
  p.RL = B + p.PS + p.PT;
  p.ksize = cv_round_host((double)(h / 15)); if (p.ksize % 2 == 0) p.ksize++;                    // TDecGop.cpp:320-322
  p.half = p.ksize / 2;
  if (p.gw32 > 128 || p.lw > 128) { delete s; set_error("fast version: width %d too large", w); return CDS_EINVAL; }
  int e = CDS_OK;
  const size_t n4 = p.n4, RL = p.RL, Bz = B;
#define A(ptr, count) if (!e) e = falloc(s, &p.ptr, (size_t)(count))
  A(r_mvx, RL * n4); A(r_mvy, RL * n4); A(r_dep, RL * n4); A(r_bytes, RL * p.nctu); A(r_sc, RL * 8); A(r_cam, RL * 2);
  A(r_t64, RL * p.nctu); A(r_t32, RL * 3 * p.g32);
  A(d_bin, Bz * n4); A(d_sm, Bz * 4 * n4); A(d_smbit, Bz * p.nctu); A(d_thr, Bz * 20);
  A(d_tacc, Bz * 3 * p.g32); A(d_tval, Bz * 3 * p.g32); A(d_sval, Bz * 3 * p.g32); A(d_sraw, Bz * 4 * p.g32);
  A(d_comb, Bz * n4); A(d_hrow, Bz * p.h4 * w); A(d_blur, Bz * (size_t)h * w); A(d_mm, Bz * 2);
#undef A
  // Gaussian taps (cv::GaussianBlur(FinalMap, Size(k, k), cvRound(H / 30), 0, BORDER_CONSTANT), TDecGop.cpp:323)
  std::vector<float> k;
  cv_gaussian_kernel_f32(p.ksize, (double)cv_round_host((double)(h / 30)), &k);
  const int z = (4 - p.half % 4) % 4;                      // leading zero taps: anchor becomes a multiple of 4
  p.hq = (p.half + z) / 4;
  p.nq_h = (p.ksize + z + 3) / 4;
  p.nq_v = (p.half + 3) / 4;
  std::vector<float> kh((size_t)p.nq_h * 4, 0.f), kv((size_t)p.nq_v * 4 + 4, 0.f);
  for (int t = 0; t < p.ksize; t++) kh[z + t] = k[t];
  kv[0] = k[p.half];
  for (int t = 1; t <= p.half; t++) kv[3 + t] = k[p.half + t];
  // temporal weights exp(-i / DeltaTemporal) (salmat.cpp:215): [0] delta 26 (mv), [1] delta 46 (bit, depth)
  std::vector<float> wt((size_t)2 * p.PT, 0.f);
  for (int i = 0; i < p.PT; i++) { wt[i] = (float)exp((double)(-i / 26.0f)); wt[p.PT + i] = (float)exp((double)(-i / 46.0f)); }
  // GenerateWeightMap (salmat.cpp:342): bit 5x5 delta 3, depth 12x12 delta 13, mv 12x12 delta 11
  std::vector<float> wm(25 + 144 + 144);
  const int Ws[3] = {5, 12, 12}; const float ds[3] = {3.f, 13.f, 11.f}; const int off[3] = {0, 25, 25 + 144};
  for (int m = 0; m < 3; m++)
    for (int i = 0; i < Ws[m]; i++) for (int j = 0; j < Ws[m]; j++) {
      const int len = Ws[m] / 2;
      const float d = (float)((i - len) * (i - len)) + (float)((j - len) * (j - len));
      const float den = 2 * (ds[m] * ds[m]);
      wm[off[m] + i * Ws[m] + j] = (float)exp((double)(-d / den));
    }
  auto up = [&](float** dst, const std::vector<float>& v) {
    if (e) return;
    e = falloc(s, dst, v.size());
    if (!e && cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("fast: cudaMemcpy tables"); e = CDS_ECUDA; }
  };
  up(&p.d_kh, kh); up(&p.d_kv, kv); up(&p.d_wt, wt); up(&p.d_wm, wm);
  s->smem_tthr = (size_t)p.g32 * 4;
  s->smem_sp[0] = ((size_t)(p.lh + 4) * (p.lw + 4) + 25) * 4;
  s->smem_sp[1] = s->smem_sp[2] = ((size_t)(p.gh32 + 12) * (p.gw32 + 12) + 144) * 4;
  s->smem_h = ((size_t)p.w4 + p.hq + p.nq_h + 2) * 4;
  if (!e && (s->smem_tthr > 48000 || s->smem_sp[1] > 48000)) { set_error("fast version: picture too large for the block-grid kernels"); e = CDS_EINVAL; }
  if (e) { fast_destroy(s); return e; }
  *out = s;
  return fast_clear(s, nullptr);
}

void fast_destroy(FastState* s) {
  if (!s) return;
  for (void* q : s->allocs) cudaFree(q);
  delete s;
}

int fast_halo(const FastState* s) { return s->p.PS + s->p.PT - 2; }

int fast_clear(FastState* s, cudaStream_t st) {
  const FastP& p = s->p;
  const size_t RL = p.RL, n4 = p.n4;
  CDS_CUDA(cudaMemsetAsync(p.r_mvx, 0, RL * n4 * 2, st)); CDS_CUDA(cudaMemsetAsync(p.r_mvy, 0, RL * n4 * 2, st));
  CDS_CUDA(cudaMemsetAsync(p.r_dep, 0, RL * n4, st)); CDS_CUDA(cudaMemsetAsync(p.r_bytes, 0, RL * p.nctu * 2, st));
  CDS_CUDA(cudaMemsetAsync(p.r_sc, 0, RL * 32, st)); CDS_CUDA(cudaMemsetAsync(p.r_cam, 0, RL * 8, st));
  CDS_CUDA(cudaMemsetAsync(p.r_t64, 0, RL * p.nctu * 4, st)); CDS_CUDA(cudaMemsetAsync(p.r_t32, 0, RL * 3 * p.g32 * 4, st));
  return CDS_OK;
}

int fast_run_batch(FastState* s, int first_poc, int n, const int16_t* mvx, const int16_t* mvy, const uint8_t* dep,
                   const uint16_t* bytes, void* out, int out_u8, int* cam_out, cudaStream_t st) {
  const FastP& p = s->p;
  if (n < 1 || n > p.B || first_poc < 0) { set_error("fast_run_batch: bad batch"); return CDS_EINVAL; }
  fast_camera_kernel<<<n, CAM_T, 0, st>>>(p, first_poc, mvx, mvy, dep, bytes, cam_out);
  CDS_LAUNCH_CHECK();
  fast_smooth_kernel<<<dim3(ceil_div(p.n4 + p.nctu, 256), n), 256, 0, st>>>(p, first_poc);
  CDS_LAUNCH_CHECK();
  if (!out) return CDS_OK;
  fast_thr_kernel<<<dim3(5, n), 256, 0, st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_temporal_kernel<<<dim3(ceil_div(p.g32, 128), 3, n), 128, 0, st>>>(p, first_poc);
  CDS_LAUNCH_CHECK();
  fast_tthr_kernel<<<dim3(3, n), 128, s->smem_tthr, st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_spatial_kernel<<<dim3(3, n), 256, s->smem_sp[1], st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_combine_kernel<<<dim3(ceil_div(p.n4, 256), n), 256, 0, st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_mm_reset_kernel<<<ceil_div(n, 128), 128, 0, st>>>(p.d_mm, n);
  CDS_LAUNCH_CHECK();
  fast_blur_h_kernel<<<dim3(p.h4, n), 256, s->smem_h, st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_blur_v_kernel<<<dim3(ceil_div(p.w, 128), p.h4, n), 128, 0, st>>>(p);
  CDS_LAUNCH_CHECK();
  const int nb = ceil_div((int)(((size_t)p.h * p.w / 4)), 256);
  if (out_u8) fast_output_kernel<uint8_t><<<dim3(nb, n), 256, 0, st>>>(p, reinterpret_cast<uint8_t*>(out));
  else fast_output_kernel<float><<<dim3(nb, n), 256, 0, st>>>(p, reinterpret_cast<float*>(out));
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

int fast_get_cam(const FastState* s, int poc, int* xy) {
  CDS_CUDA(cudaMemcpy(xy, s->p.r_cam + 2 * (poc % s->p.RL), 8, cudaMemcpyDeviceToHost));
  return CDS_OK;
}

int fast_debug_copy(FastState* s, const char* which, int f, void* dst, size_t bytes, cudaStream_t st) {
  const FastP& p = s->p;
  const void* src = nullptr; size_t have = 0;
  const std::string wname(which);
  if (wname == "fast_sm") { src = p.d_sm + (size_t)f * 4 * p.n4; have = (size_t)4 * p.n4 * 4; }
  else if (wname == "fast_smbit") { src = p.d_smbit + (size_t)f * p.nctu; have = (size_t)p.nctu * 4; }
  else if (wname == "fast_thr") { src = p.d_thr + (size_t)f * 20; have = 80; }
  else if (wname == "fast_tacc") { src = p.d_tacc + (size_t)f * 3 * p.g32; have = (size_t)3 * p.g32 * 4; }
  else if (wname == "fast_tval") { src = p.d_tval + (size_t)f * 3 * p.g32; have = (size_t)3 * p.g32 * 4; }
  else if (wname == "fast_sval") { src = p.d_sval + (size_t)f * 3 * p.g32; have = (size_t)3 * p.g32 * 4; }
  else if (wname == "fast_comb") { src = p.d_comb + (size_t)f * p.n4; have = (size_t)p.n4 * 4; }
  else if (wname == "fast_prof") {
    if (bytes > sizeof(long long) * 32) { set_error("fast_debug_copy: fast_prof is 256 bytes"); return CDS_EINVAL; }
    CDS_CUDA(cudaStreamSynchronize(st));
    CDS_CUDA(cudaMemcpyFromSymbol(dst, g_fast_prof, bytes));
    return CDS_OK;
  }
  else if (wname == "fast_blur") { src = p.d_blur + (size_t)f * p.h * p.w; have = (size_t)p.h * p.w * 4; }
  else { set_error("fast_debug_copy: unknown '%s'", which); return CDS_EINVAL; }
  if (bytes > have) { set_error("fast_debug_copy: %zu > %zu bytes", bytes, have); return CDS_EINVAL; }
  CDS_CUDA(cudaStreamSynchronize(st));
  CDS_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return CDS_OK;
}

}  // namespace cds
