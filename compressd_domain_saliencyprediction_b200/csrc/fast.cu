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
//   * the sequential float sums of MapThreshold (salmat.cpp:155-162) and Umean (TDecGop.cpp:622-640) are evaluated in parallel
//     and still bit-exactly: inside one binade of the running sum every addition moves it by an integer number of ulps that
//     does not depend on the sum, so pieces of the sequence reduce as integers (block_seq_sum); only the additions that cross
//     a binade, ties, and terms comparable to the sum are done one float at a time.  Maps that are replicated blocks (bit map,
//     the three full-resolution temporal maps) use the same idea on runs "add v n times" (warp_select_sum_runs, seq_add_run),
//     so a 1920x1080 map costs about one step per block row, not H x W additions;
//   * the H x W feature maps are never materialised: all nine are constant on 4x4-pixel cells (or 32 / 64-pixel blocks), so
//     the weighted sum is formed per cell and only the final Gaussian runs at pixel resolution, in the tap order of
//     cv::GaussianBlur (row pass: taps in order from zero; column pass: centre tap, then symmetric pairs).
// Data layout: raw ring [RL = 2B + PS + PT] of stage-(a) maps in their narrow types (int16 mv, u8 depth, u16 bytes per CTU) + 8 floats
// and 2 ints per picture; coarse temporal rings (CTU grid, 32-pixel grid); everything else is batch-local.  fast_front (camera
// stages) and fast_back (smoothing .. map) are the two halves the clip context may run on different streams.
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
  unsigned* d_mm; int* d_cami; float* d_camf;
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

// ---- exact emulation of a sequential float sum S = S + t_i (raster order) without walking it one term at a time ------------
// Inside one binade of S every addition moves S by a whole number of ulps that does not depend on S: fl(S + v) = S + rn_u(v),
// rn_u = round to the binade's grid u (ties excepted: they depend on the parity of S).  A piece of the sequence can therefore be
// applied as an INTEGER sum of increments, in any order, provided (a) no term is a tie on this grid, (b) every partial sum stays
// inside the binade (bounded by the sums of the positive and of the negative increments), (c) every |v| < 2^(e-1), so that the
// probe below is valid.  Pieces that fail are added one float at a time by one thread — that is where S crosses into the next
// binade — and the rest is re-evaluated on the new grid.  The result is bit-identical to the sequential loop.

// increment (in ulps of binade eb) of one term; probe C = 1.5 * 2^e has an even mantissa, C + ulp an odd one: they disagree on ties
__device__ __forceinline__ bool term_increment(float v, unsigned Cb, float lim, int* q) {
  if (!(fabsf(v) < lim)) return false;
  const int q0 = (int)(__float_as_uint(__fadd_rn(__uint_as_float(Cb), v)) - Cb);
  const int q1 = (int)(__float_as_uint(__fadd_rn(__uint_as_float(Cb + 1u), v)) - (Cb + 1u));
  *q = q0;
  return q0 == q1;
}

// A round covers 32 pieces of 128 * SEQ_G consecutive cells (one warp x one chunk of T * 4 * SEQ_G cells; a thread owns SEQ_G groups
// of 4 consecutive cells per chunk); K = 32 / (T/32) chunks per round.
constexpr int SEQ_G = 1;     // 2 (8 cells per thread and chunk, half as many rounds) measured 10 % slower: the rounds are work bound, not latency bound
template <int T>
struct SeqSumSmem {
  static constexpr int NW = T / 32, K = 32 / NW;
  float terms[K][T * 4 * SEQ_G];
  unsigned char sel[K][T];             // SEQ_G * 4 selection bits per thread
  int pp[32], pn[32], pf[32];          // per piece: positive ulps, negative ulps, flag (1: needs the sequential path, -1: empty)
  float S; int restart;
};

// The whole CTA (T threads) sums ngroups groups of 4 consecutive cells; load(g) fetches the raw words of group g, eval(raw, t)
// returns its selection bits and four terms.  Returns the sum to every thread; *count_out = number of selected terms.
template <int T, typename LoadFn, typename EvalFn>
__device__ float block_seq_sum(int ngroups, LoadFn load, EvalFn eval, SeqSumSmem<T>& sm, int* count_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = T / 32, K = 32 / NW, G = SEQ_G;
  int cnt = 0;
  __syncthreads();
  if (tid == 0) sm.S = 0.f;
  decltype(load(0)) raw[K][G];
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int u = 0; u < G; u++) raw[k][u] = load(min((k * T + tid) * G + u, ngroups - 1));   // raw words of the next round stay in registers
  for (int g0 = 0; g0 < ngroups; g0 += K * T * G) {
    float t[K][G][4]; unsigned sel[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      sel[k] = 0u;
#pragma unroll
      for (int u = 0; u < G; u++) {
        const int g = g0 + (k * T + tid) * G + u;
        const unsigned su = g < ngroups ? eval(raw[k][u], t[k][u]) : 0u;
        sel[k] |= su << (4 * u);
        raw[k][u] = load(min(g + K * T * G, ngroups - 1));                               // prefetch
      }
      cnt += __popc(sel[k]);
    }
    __syncthreads();                                   // the previous round's reads of sm.terms are over
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
      for (int u = 0; u < G; u++)
        *reinterpret_cast<float4*>(&sm.terms[k][(tid * G + u) * 4]) = make_float4(t[k][u][0], t[k][u][1], t[k][u][2], t[k][u][3]);
      sm.sel[k][tid] = (unsigned char)sel[k];
    }
    int first = 0;                                     // pieces before `first` are already in S
    int regrids = 0;                                   // times this round had to be re-evaluated on a new grid
    while (true) {
      __syncthreads();
      const unsigned Sb = __float_as_uint(sm.S);
      const unsigned eb = (Sb >> 23) & 0xffu; const bool neg = (Sb >> 31) != 0u;
      const bool bad = eb < 2u || eb >= 0xfeu;
      const unsigned Cb = (eb << 23) | 0x400000u;
      const float lim = __uint_as_float((eb - 1u) << 23);
#pragma unroll
      for (int k = 0; k < K; k++) {
        const int piece = k * NW + wid;
        int qp = 0, qn = 0, flag = bad ? 1 : 0;
        if (piece >= first && !bad) {
#pragma unroll
          for (int u = 0; u < G; u++)
#pragma unroll
            for (int j = 0; j < 4; j++)
              if ((sel[k] >> (4 * u + j)) & 1u) {
                int q;
                if (!term_increment(neg ? -t[k][u][j] : t[k][u][j], Cb, lim, &q)) flag = 1;
                else if (q > 0) qp += q; else qn -= q;
              }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { qp += __shfl_xor_sync(FULL, qp, o); qn += __shfl_xor_sync(FULL, qn, o); flag |= __shfl_xor_sync(FULL, flag, o); }
        const unsigned anysel = __ballot_sync(FULL, sel[k] != 0u);
        if (lane == 0) { sm.pp[piece] = qp; sm.pn[piece] = qn; sm.pf[piece] = anysel ? flag : -1; }
      }
      __syncthreads();
      if (wid == 0) {
        // lane l owns piece l: prefix sums of the increments decide how far the round can be applied in one step
        float S = sm.S; int restart = 32;
        const int pf = sm.pf[lane];
        const unsigned pp = (lane >= first && pf >= 0) ? (unsigned)sm.pp[lane] : 0u, pn = (lane >= first && pf >= 0) ? (unsigned)sm.pn[lane] : 0u;
        int from = first;
        while (from < 32) {
          const unsigned sb = __float_as_uint(S), ab = sb & 0x7fffffffu, m = (ab & 0x7fffffu) | 0x800000u;
          unsigned cp = lane >= from ? pp : 0u, cn = lane >= from ? pn : 0u;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const unsigned up = __shfl_up_sync(FULL, cp, o), un = __shfl_up_sync(FULL, cn, o); if (lane >= o) { cp += up; cn += un; } }
          // a sum that keeps changing binade (it hovers around zero, or it is still tiny) is walked float by float for the rest of the round
          const bool walk = regrids > 2;
          const bool inval = lane >= from && pf >= 0 && (walk || pf > 0 || m + cp > 0xffffffu || (cn != 0u && m < 0x800001u + cn));
          const unsigned bal = __ballot_sync(FULL, inval);
          const int stop = bal ? __ffs(bal) - 1 : 32;              // first piece that cannot be applied in closed form
          const int lastok = stop - 1;
          if (lastok >= from) {
            const unsigned ap = __shfl_sync(FULL, cp, lastok), an = __shfl_sync(FULL, cn, lastok);
            S = __uint_as_float((sb & 0x80000000u) | (ab + ap - an));
          }
          if (stop == 32) break;
          // piece `stop` one float at a time: lane l fetches thread l's terms, the chain of additions reads them by shuffle
          const int k = stop / NW, w = stop - k * NW;
          {
            float4 tq[G];
#pragma unroll
            for (int u = 0; u < G; u++) tq[u] = *reinterpret_cast<const float4*>(&sm.terms[k][((w * 32 + lane) * G + u) * 4]);
            const unsigned sq = sm.sel[k][w * 32 + lane];
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
              const unsigned sl = __shfl_sync(FULL, sq, i);
#pragma unroll
              for (int u = 0; u < G; u++) {
                const float a0 = __shfl_sync(FULL, tq[u].x, i), a1 = __shfl_sync(FULL, tq[u].y, i), a2 = __shfl_sync(FULL, tq[u].z, i), a3 = __shfl_sync(FULL, tq[u].w, i);
                if (sl & (1u << (4 * u))) S = __fadd_rn(S, a0);
                if (sl & (2u << (4 * u))) S = __fadd_rn(S, a1);
                if (sl & (4u << (4 * u))) S = __fadd_rn(S, a2);
                if (sl & (8u << (4 * u))) S = __fadd_rn(S, a3);
              }
            }
          }
          from = stop + 1;
          const unsigned nb = __float_as_uint(S);
          if (!walk && (((nb >> 23) & 0xffu) != eb || (nb >> 31) != (Sb >> 31))) { restart = from; break; }     // new grid: re-evaluate the rest
        }
        if (lane == 0) { sm.S = S; sm.restart = restart; }
      }
      __syncthreads();
      first = sm.restart;
      if (first >= 32) break;
      regrids++;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  __syncthreads();
  if (lane == 0) sm.pp[wid] = cnt;
  __syncthreads();
  int total = 0;
  for (int w = 0; w < NW; w++) total += sm.pp[w];
  *count_out = total;
  return sm.S;
}

// MapThreshold's first loop (salmat.cpp:155-162) on a map of replicated blocks: block (i, j) of a gh x gw grid covers
// min(ms, rows - ms*i) x min(ms, cols - ms*j) elements of a rows x cols raster, so an element row of block row i is a sequence of
// runs "add v_e n_e times".  One warp: inside a binade a whole element row adds Q = sum n_e * q_e ulps, and k rows add k * Q; the run
// in which S leaves the binade (found with a prefix scan), ties and terms too large for the probe go through seq_add_run.
__device__ float warp_select_sum_runs(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, float* rv, int* rn,
                                      int* rq, int* count) {
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
    if (L == 0) continue;
    const int nr = min(ms, rows - ms * i);
    int per_row = 0;
    for (int e = lane; e < L; e += 32) per_row += rn[e];
    per_row = warp_sum(per_row);
    cnt += nr * per_row;
    int rows_left = nr, pos = 0;
    while (rows_left > 0) {
      const unsigned Sb = __float_as_uint(S), eb = (Sb >> 23) & 0xffu;
      const bool bad = eb < 2u || eb >= 0xfeu;
      const unsigned Cb = (eb << 23) | 0x400000u;
      const float lim = __uint_as_float((eb - 1u) << 23);
      const unsigned m = (Sb & 0x7fffffu) | 0x800000u;
      int qsum = 0, anyflag = bad ? 1 : 0;
      for (int e = lane; e < L; e += 32) {
        int q = 0;
        const bool ok = !bad && term_increment(rv[e], Cb, lim, &q);
        const int qr = q > 0x1000000 / rn[e] ? 0x1000000 : q * rn[e];      // >= 2^24 ulps leaves the binade whatever S is
        rq[e] = ok ? qr : -1;
        if (ok) qsum += qr; else anyflag = 1;
      }
      __syncwarp();
      qsum = warp_sum(qsum); anyflag = __any_sync(FULL, anyflag);           // at most 128 runs x 2^24: fits
      if (pos == 0 && !anyflag) {
        if (qsum == 0) break;                                   // every remaining addition leaves S unchanged
        const unsigned k = min((unsigned)rows_left, (0xffffffu - m) / (unsigned)qsum);
        S = __uint_as_float(Sb + k * (unsigned)qsum);
        rows_left -= (int)k;
        if (rows_left == 0) break;
        if (k > 0) continue;                                    // S moved: recompute the bounds before walking the crossing row
      }
      // walk the current element row from run `pos`: everything before the first run that is flagged or leaves the binade is applied at once
      unsigned running = 0; int found = -1; unsigned before = 0;
      for (int e0 = 0; e0 < L && found < 0; e0 += 32) {
        const int e = e0 + lane;
        const bool act = e >= pos && e < L;
        const int rqe = act ? rq[e] : 0;
        unsigned pref = rqe > 0 ? (unsigned)rqe : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, pref, o); if (lane >= o) pref += u; }
        const bool cross = act && (rqe < 0 || m + running + pref > 0xffffffu);
        const unsigned bal = __ballot_sync(FULL, cross);
        if (bal) {
          const int l = __ffs(bal) - 1;
          found = e0 + l;
          const unsigned incl = __shfl_sync(FULL, pref, l);
          const int own = __shfl_sync(FULL, rqe, l);
          before = running + incl - (own > 0 ? (unsigned)own : 0u);
        } else running += __shfl_sync(FULL, pref, 31);
      }
      if (found < 0) { S = __uint_as_float(__float_as_uint(S) + running); pos = 0; rows_left--; continue; }
      S = __uint_as_float(__float_as_uint(S) + before);
      S = seq_add_run(S, rv[found], rn[found]);
      pos = found + 1;
      if (pos == L) { pos = 0; rows_left--; }
    }
    __syncwarp();
  }
  *count = cnt;
  return S;
}

// ------------------------------------------------------------------------------------------------ K1: CameraDect + vote_alg + Umean
__device__ __forceinline__ int direction_bin(float y, float x) {          // TDecGop.cpp:584-587
  // index = ((atan2f(y, x) / PI) + 1) * 8.0 truncated; only values next to a bin edge need the exactly rounded float atan2
  const float tf = (atan2f(y, x) * 0.318309886f + 1.0f) * 8.0f;
  if (fabsf(tf - rintf(tf)) >= 1e-3f) return (int)tf;
  const float a = __double2float_rn(atan2((double)y, (double)x));
  const int idx = (int)(((double)a / PI_REF + 1.0) * 8.0);
  return idx == 16 ? 15 : idx;
}

// K1a, grid (chunks of 1024 cells, pictures): copy the picture into the raw ring, background mask, BackNum / StaticNum
// (TDecGop.cpp:509-519), direction bin of every background cell (vote_alg :580-592) and the 16-bin histogram.
// d_cami[b]: {BackNum, StaticNum, hist[16], max depth, max m_norm (float bits), max bytes}
__global__ void __launch_bounds__(256)
fast_cam_scan_kernel(FastP P, int first_poc, const int16_t* __restrict__ mvx, const int16_t* __restrict__ mvy,
                     const uint8_t* __restrict__ dep, const uint16_t* __restrict__ bytes) {
  const int b = blockIdx.y, F = first_poc + b, slot = F % P.RL, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n4 = P.n4, w4 = P.w4, lw = P.lw;
  const uint16_t* by = bytes + (size_t)b * P.nctu;
  __shared__ int s_i[32]; __shared__ int s_hist[8][16];
  int mn = 65535, mxb = 0;
  for (int i = tid; i < P.nctu; i += 256) { mn = min(mn, (int)by[i]); mxb = max(mxb, (int)by[i]); }
  mn = block_reduce(mn, OpMinI(), s_i);
  if (blockIdx.x == 0) {
    uint16_t* rb = P.r_bytes + (size_t)slot * P.nctu;
    for (int i = tid; i < P.nctu; i += 256) rb[i] = by[i];
    mxb = block_reduce(mxb, OpMaxI(), s_i);
    if (tid == 0) P.d_cami[b * 24 + 20] = mxb;
  }
  const int k = (blockIdx.x * 256 + tid) * 4;
  int back = 0, stat = 0, mxd = 0; int bi[4] = {255, 255, 255, 255};
  if (k < n4) {
    const size_t o = (size_t)b * n4 + k, ro = (size_t)slot * n4 + k;
    const short4 mx = *reinterpret_cast<const short4*>(mvx + o), my = *reinterpret_cast<const short4*>(mvy + o);
    const uchar4 d = *reinterpret_cast<const uchar4*>(dep + o);
    *reinterpret_cast<short4*>(P.r_mvx + ro) = mx; *reinterpret_cast<short4*>(P.r_mvy + ro) = my;
    *reinterpret_cast<uchar4*>(P.r_dep + ro) = d;
    mxd = max(max((int)d.x, (int)d.y), max((int)d.z, (int)d.w));
    const int r = k / w4, c = k - r * w4;                   // the 4 cells share a row (w4 % 4 == 0) and a CTU column (c % 4 == 0)
    if ((int)by[(r >> 4) * lw + (c >> 4)] < mn + 20) {
      const short xs[4] = {mx.x, mx.y, mx.z, mx.w}, ys[4] = {my.x, my.y, my.z, my.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        back++;
        if ((unsigned)((int)xs[j] * xs[j]) + (unsigned)((int)ys[j] * ys[j]) <= 32u) stat++;   // mvx^2 + mvy^2 <= 2 pixels^2, in quarter-pel units
        bi[j] = direction_bin(__fmul_rn((float)ys[j], 0.25f), __fmul_rn((float)xs[j], 0.25f));
      }
    }
    *reinterpret_cast<uchar4*>(P.d_bin + o) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1], (unsigned char)bi[2], (unsigned char)bi[3]);
  }
  int mycount = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int c = __popc(__ballot_sync(FULL, bi[0] == j)) + __popc(__ballot_sync(FULL, bi[1] == j)) +
                  __popc(__ballot_sync(FULL, bi[2] == j)) + __popc(__ballot_sync(FULL, bi[3] == j));
    if (lane == j) mycount = c;
  }
  if (lane < 16) s_hist[wid][lane] = mycount;
  back = block_reduce(back, OpAddI(), s_i); stat = block_reduce(stat, OpAddI(), s_i); mxd = block_reduce(mxd, OpMaxI(), s_i);
  int* ci = P.d_cami + b * 24;
  if (tid == 0) { if (back) atomicAdd(ci, back); if (stat) atomicAdd(ci + 1, stat); atomicMax(ci + 18, mxd); }
  if (tid < 16) { int c = 0; for (int w = 0; w < 8; w++) c += s_hist[w][tid]; if (c) atomicAdd(ci + 2 + tid, c); }
}

// K1b, grid (component x / y, pictures): static / moving decision, winning bin (:600 first maximum), Umean of its members for one
// component (exact sequential float sums).  d_camf[b] = {vx, vy}.
constexpr int CAM_T = 512;

__global__ void __launch_bounds__(CAM_T) fast_cam_vote_kernel(FastP P, int first_poc) {
  const int comp = blockIdx.x, b = blockIdx.y, F = first_poc + b, slot = F % P.RL, tid = threadIdx.x;
  const int n4 = P.n4;
  const int16_t* m = (comp == 0 ? P.r_mvx : P.r_mvy) + (size_t)slot * n4;
  const uint8_t* bin = P.d_bin + (size_t)b * n4;
  const int* ci = P.d_cami + b * 24;
  __shared__ long long s_l[32];
  __shared__ SeqSumSmem<CAM_T> s_seq;
  const int back = ci[0], stat = ci[1];
  float v = 0.f;
  long long t_prof = clock64();
  if (stat <= back / 2) {                                   // TDecGop.cpp:529 (integer division)
    int best = 0, bc = -1;
    for (int j = 0; j < 16; j++) if (ci[2 + j] > bc) { bc = ci[2 + j]; best = j; }
    const int len = bc;
    // first loop of Umean: exact in float whenever every partial sum is (sum of |m| < 2^24 quarter-pels)
    long long s = 0, a = 0;
    for (int g0 = 0; g0 < n4 / 4; g0 += 4 * CAM_T) {
      uchar4 bb[4]; short4 mm[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int g = min(g0 + u * CAM_T + tid, n4 / 4 - 1); bb[u] = reinterpret_cast<const uchar4*>(bin)[g]; mm[u] = reinterpret_cast<const short4*>(m)[g]; }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (g0 + u * CAM_T + tid < n4 / 4) {
          if (bb[u].x == best) { s += mm[u].x; a += abs((int)mm[u].x); }
          if (bb[u].y == best) { s += mm[u].y; a += abs((int)mm[u].y); }
          if (bb[u].z == best) { s += mm[u].z; a += abs((int)mm[u].z); }
          if (bb[u].w == best) { s += mm[u].w; a += abs((int)mm[u].w); }
        }
    }
    s = block_reduce(s, OpAddL(), s_l); a = block_reduce(a, OpAddL(), s_l);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) { const long long t_ = clock64(); g_fast_prof[0] = t_ - t_prof; t_prof = t_; }
    int cnt;
    struct Raw { uchar4 b; short4 m; };
    auto load = [&](int g) -> Raw { return Raw{reinterpret_cast<const uchar4*>(bin)[g], reinterpret_cast<const short4*>(m)[g]}; };
    auto plain = [&](const Raw& r, float* t) -> unsigned {
      t[0] = __fmul_rn((float)r.m.x, 0.25f); t[1] = __fmul_rn((float)r.m.y, 0.25f); t[2] = __fmul_rn((float)r.m.z, 0.25f); t[3] = __fmul_rn((float)r.m.w, 0.25f);
      return (r.b.x == best ? 1u : 0u) | (r.b.y == best ? 2u : 0u) | (r.b.z == best ? 4u : 0u) | (r.b.w == best ? 8u : 0u);
    };
    const float sum1 = (a < (1ll << 24)) ? __fmul_rn((float)s, 0.25f) : block_seq_sum<CAM_T>(n4 / 4, load, plain, s_seq, &cnt);
    const float mean = __fdiv_rn(sum1, (float)len);
    const bool flag = mean > 0.f;
    auto clamped = [&](const Raw& r, float* t) -> unsigned {          // TDecGop.cpp:633-639
      const unsigned sel = plain(r, t);
#pragma unroll
      for (int j = 0; j < 4; j++) t[j] = ((flag && t[j] < mean) || (!flag && t[j] > mean)) ? t[j] : mean;
      return sel;
    };
    const float sum2 = block_seq_sum<CAM_T>(n4 / 4, load, clamped, s_seq, &cnt);
    v = __fdiv_rn(sum2, (float)len);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) { const long long t_ = clock64(); g_fast_prof[2] = t_ - t_prof; g_fast_prof[1] = len; g_fast_prof[3] = (a < (1ll << 24)); }
  }
  if (tid == 0) P.d_camf[b * 2 + comp] = v;
}

// K1c, grid (chunks, pictures): maximum of the compensated m_norm (CalDist :552), which feeds MatrixNorm (TDecGop.cpp:294)
__global__ void __launch_bounds__(256) fast_cam_norm_kernel(FastP P, int first_poc) {
  const int b = blockIdx.y, slot = (first_poc + b) % P.RL, g = blockIdx.x * 256 + threadIdx.x;
  __shared__ float s_f[32];
  const int* ci = P.d_cami + b * 24;
  const bool moving = ci[1] <= ci[0] / 2;
  const float nx = moving ? -P.d_camf[b * 2] : -0.0f, ny = moving ? -P.d_camf[b * 2 + 1] : -0.0f;
  float mxn = 0.f;
  if (g < P.n4 / 4) {
    const short4 x4 = reinterpret_cast<const short4*>(P.r_mvx + (size_t)slot * P.n4)[g], y4 = reinterpret_cast<const short4*>(P.r_mvy + (size_t)slot * P.n4)[g];
    const short xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float x = __fadd_rn(__fmul_rn((float)xs[j], 0.25f), nx), y = __fadd_rn(__fmul_rn((float)ys[j], 0.25f), ny);
      mxn = fmaxf(mxn, __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));
    }
  }
  mxn = block_reduce(mxn, OpMaxF(), s_f);
  if (threadIdx.x == 0 && mxn > 0.f) atomicMax(reinterpret_cast<unsigned*>(P.d_cami + b * 24 + 19), __float_as_uint(mxn));
}

// K1d, one thread per picture: the per-picture scalars the later kernels read from the ring
__global__ void fast_cam_final_kernel(FastP P, int first_poc, int n, int* __restrict__ cam_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const int slot = (first_poc + b) % P.RL;
  const int* ci = P.d_cami + b * 24;
  const bool moving = ci[1] <= ci[0] / 2;
  const float vx = P.d_camf[b * 2], vy = P.d_camf[b * 2 + 1];
  const int mxb = ci[20], mxd = ci[18];
  const float mxn = __uint_as_float((unsigned)ci[19]);
  float* sc = P.r_sc + (size_t)slot * 8;
  sc[0] = moving ? -vx : -0.0f; sc[1] = moving ? -vy : -0.0f;
  sc[2] = mxb > 0 ? rcp_scale((float)mxb) : 1.f;
  sc[3] = mxd > 0 ? rcp_scale((float)mxd) : 1.f;
  sc[4] = mxn > 0.f ? rcp_scale(mxn) : 1.f;
  const int cx = moving ? __float2int_rn(vx) : 0, cy = moving ? __float2int_rn(vy) : 0;        // cvRound
  P.r_cam[slot * 2] = cx; P.r_cam[slot * 2 + 1] = cy;
  if (cam_out) { cam_out[b * 2] = cx; cam_out[b * 2 + 1] = cy; }
}

// ------------------------------------------------------------------------------------------------ K2: ForwardFilter (salmat.cpp:39-55)
// tempmat = tempmat + SmoothFrame[i] / cnt for ring slots i = 0.. in SLOT order: fl(fl(x * (float)(1./cnt)) + acc).
// A CTA owns 128 cells x SM_G consecutive pictures: the raw values of the SM_G + PS - 1 ring pictures they read are converted
// once (compensated mv, its norm, normalised depth) into shared memory, then each thread walks its picture's slots.
constexpr int SM_G = 4;

__global__ void __launch_bounds__(128 * SM_G) fast_smooth_kernel(FastP P, int first_poc, int n) {
  extern __shared__ float4 s_raw[];                      // [SM_G + PS - 1][128] {depth, norm, x, y}
  const int c = threadIdx.x, g = threadIdx.y, tid = g * 128 + c;
  const int k = blockIdx.x * 128 + c, b0 = blockIdx.y * SM_G, F0 = first_poc + b0;
  const int n4 = P.n4, PS = P.PS, nf = SM_G + PS - 1;
  const int fbase = F0 - (PS - 1);                       // picture held by tile row 0
  const int slot0 = ((fbase % P.RL) + P.RL) % P.RL;
  for (int i = tid; i < nf * 128; i += 128 * SM_G) {
    const int fr = i >> 7, cc = i & 127, f = fbase + fr, kk = blockIdx.x * 128 + cc;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f >= 0 && kk < n4 && f < first_poc + n) {
      const int slot = slot0 + fr < P.RL ? slot0 + fr : slot0 + fr - P.RL;
      const float* sc = P.r_sc + (size_t)slot * 8;
      const size_t o = (size_t)slot * n4 + kk;
      const float x = __fadd_rn(__fmul_rn((float)P.r_mvx[o], 0.25f), sc[0]), y = __fadd_rn(__fmul_rn((float)P.r_mvy[o], 0.25f), sc[1]);
      v.x = __fmul_rn((float)P.r_dep[o], sc[3]);
      v.y = __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))), sc[4]);
      v.z = x; v.w = y;
    }
    s_raw[i] = v;
  }
  __syncthreads();
  const int b = b0 + g, F = F0 + g;
  if (b >= n || k >= n4) return;
  const int cnt = F < PS ? F + 1 : PS;
  const float rc = __double2float_rn(__ddiv_rn(1.0, (double)cnt));
  float ad = 0.f, an = 0.f, ax = 0.f, ay = 0.f;
  const int i0 = F % PS;                                   // ring slot i holds picture F - i0 + i (i <= i0) or F - i0 + i - PS (i > i0)
  const float4* col = s_raw + (F - i0 - fbase) * 128 + c;
  for (int i = 0; i < cnt; i++) {
    const float4 v = col[(i - (i > i0 ? PS : 0)) * 128];
    ad = __fadd_rn(__fmul_rn(v.x, rc), ad); an = __fadd_rn(__fmul_rn(v.y, rc), an);
    ax = __fadd_rn(__fmul_rn(v.z, rc), ax); ay = __fadd_rn(__fmul_rn(v.w, rc), ay);
  }
  float* sm = P.d_sm + (size_t)b * 4 * n4;
  sm[k] = ad; sm[n4 + k] = an; sm[2 * (size_t)n4 + k] = ax; sm[3 * (size_t)n4 + k] = ay;
  const int r = k / P.w4, cc = k - r * P.w4;
  if (!(r & 7) && !(cc & 7)) {                            // MapDownSample by MinSize/4 = 8 (salmat.cpp:205)
    float* t = P.r_t32 + (size_t)(F % P.RL) * 3 * P.g32 + (r >> 3) * P.gw32 + (cc >> 3);
    t[0] = ad; t[P.g32] = ax; t[2 * P.g32] = ay;
  }
}

// the bit map lives on the CTU grid (the x16 replication commutes with the elementwise filter)
__global__ void __launch_bounds__(128) fast_smooth_bit_kernel(FastP P, int first_poc) {
  const int b = blockIdx.y, F = first_poc + b, k = blockIdx.x * 128 + threadIdx.x;
  if (k >= P.nctu) return;
  const int PS = P.PS, cnt = F < PS ? F + 1 : PS;
  const float rc = __double2float_rn(__ddiv_rn(1.0, (double)cnt));
  float acc = 0.f;
  for (int i = 0; i < cnt; i++) {
    const int f = F < PS ? i : F - ((F - i) % PS), slot = f % P.RL;
    const float v = __fmul_rn((float)P.r_bytes[(size_t)slot * P.nctu + k], P.r_sc[(size_t)slot * 8 + 2]);
    acc = __fadd_rn(__fmul_rn(v, rc), acc);
  }
  P.d_smbit[(size_t)b * P.nctu + k] = acc;
  P.r_t64[(size_t)(F % P.RL) * P.nctu + k] = acc;
}

// ------------------------------------------------------------------------------------------------ K3: MapThreshold parameters
// job (map, picture): map 0 bit (CTU grid replicated x16), 1 depth, 2 mv norm, 3 mv x, 4 mv y.  Output d_thr[b][map] =
// {t', mode (0: src.copyTo(dst)), max of the map, scale of the following MatrixNorm of thr(map)}
__global__ void __launch_bounds__(256) fast_thr_kernel(FastP P) {
  const int map = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int n4 = P.n4;
  __shared__ double s_d[32]; __shared__ float s_f[32];
  __shared__ float s_rv[128]; __shared__ int s_rn[128], s_rq[128];
  __shared__ SeqSumSmem<256> s_seq;
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
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (int i0 = 0; i0 < n4 / 4; i0 += 1024) {                 // 4 independent loads in flight per thread
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int i = i0 + u * 256 + tid; q[u] = i < n4 / 4 ? v4[i] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (i0 + u * 256 + tid < n4 / 4) {
          sum += (double)sum4(q[u].x, q[u].y, q[u].z, q[u].w);
          mx = fmaxf(mx, fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w)));
        }
    }
  }
  sum = block_reduce(sum, OpAddD(), s_d); mx = block_reduce(mx, OpMaxF(), s_f);
  const double mean = __dmul_rn(sum, __ddiv_rn(1.0, (double)n4));       // cv::mean: sum * (1. / N)
  const double t = (double)mx - mean;
  int mode = !(t == 0.0 || t >= (double)mx);
  float tp = 0.f;
  FPROF(8);
  if (mode) {
    int count = 0; float S = 0.f;
    if (map == 0) {
      if (tid < 32) S = warp_select_sum_runs(bs, P.lh, P.lw, P.h4, P.w4, 16, t, s_rv, s_rn, s_rq, &count);
    } else {
      auto load = [&](int g) -> float4 { return reinterpret_cast<const float4*>(v)[g]; };
      auto sel_gt = [&](const float4& q, float* tt) -> unsigned {
        tt[0] = q.x; tt[1] = q.y; tt[2] = q.z; tt[3] = q.w;
        return ((double)q.x > t ? 1u : 0u) | ((double)q.y > t ? 2u : 0u) | ((double)q.z > t ? 4u : 0u) | ((double)q.w > t ? 8u : 0u);
      };
      S = block_seq_sum<256>(n4 / 4, load, sel_gt, s_seq, &count);
    }
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
  __shared__ double s_d[32]; __shared__ float s_f[32]; __shared__ float s_rv[128]; __shared__ int s_rn[128], s_rq[128];
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
    const float S = warp_select_sum_runs(s_v1, gh, gw, P.h, P.w, ms, t, s_rv, s_rn, s_rq, &count);
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
// job (kind, picture): kind 0 bit (5x5, delta 3, CTU grid), 1 depth (12x12, delta 13), 2 / 3 mv x / y (12x12, delta 11)
constexpr int SP_T = 512;

__global__ void __launch_bounds__(SP_T) fast_spatial_kernel(FastP P) {
  const int kind = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int W = kind == 0 ? 5 : 12, len = W / 2, ms4 = kind == 0 ? 16 : 8;
  const int rn = kind == 0 ? P.lh : P.gh32, cn = kind == 0 ? P.lw : P.gw32, g = rn * cn;
  const int pr = rn + 2 * len, pc = cn + 2 * len;
  extern __shared__ float s_P[];                         // padded map pr x pc, then the window weights
  float* s_w = s_P + pr * pc;
  __shared__ float s_f[32];
  const float* wm = P.d_wm + (kind == 0 ? 0 : kind == 1 ? 25 : 25 + 144);
  for (int i = tid; i < W * W; i += SP_T) s_w[i] = wm[i];
  const int map = kind == 0 ? 0 : kind == 1 ? 1 : kind + 1;          // thr job of this map (mv x: 3, mv y: 4)
  const float* th = P.d_thr + ((size_t)b * 5 + map) * 4;
  const float tp = th[0]; const int mode = (int)th[1];
  const float* cur = kind == 0 ? P.d_smbit + (size_t)b * P.nctu : P.d_sm + ((size_t)b * 4 + map - 1) * P.n4;
  // S = down(thr(cur)); zero pad; MatrixNorm2 of the padded map
  float mn = 0.f, mx = 0.f;                              // the zero border takes part in the min / max
  for (int i = tid; i < pr * pc; i += SP_T) s_P[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < g; i += SP_T) {
    const int r = i / cn, c = i - r * cn;
    const float v = thr_apply(kind == 0 ? cur[i] : cur[(size_t)(r * ms4) * P.w4 + c * ms4], tp, mode);
    s_P[(r + len) * pc + c + len] = v;
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  mn = block_reduce(mn, OpMinF(), s_f); mx = block_reduce(mx, OpMaxF(), s_f);
  const float neg = -mn;                                 // dst = src - min  ->  fl(src + (float)(-min))
  const float mx2 = __fadd_rn(mx, neg);                  // max of the shifted map (the shift is monotonic)
  const float s = mx2 > 0.f ? rcp_scale(mx2) : 1.f;
  for (int i = tid; i < pr * pc; i += SP_T) { const float x = __fadd_rn(s_P[i], neg); s_P[i] = mx2 > 0.f ? __fmul_rn(x, s) : x; }
  __syncthreads();
  // out(i,j) = sum(|window - centre| .* WeightMap): float products, cv::sum order (4 floats in float, groups in double)
  float* raw = P.d_sraw + ((size_t)b * 4 + kind) * P.g32;
  float omax = 0.f;
  const int nfull = (W * W) & ~3;
  for (int i = tid; i < g; i += SP_T) {
    const int r = i / cn, c = i - r * cn;
    const float ctr = s_P[(r + len) * pc + c + len];
    double acc = 0.0; float grp = 0.f; int q = 0, e = 0;
    for (int u = 0; u < W; u++) {
      const float* row = s_P + (r + u) * pc + c;
      for (int v = 0; v < W; v++, e++) {
        const float p = __fmul_rn(fabsf(__fsub_rn(row[v], ctr)), s_w[e]);
        if (e < nfull) {
          grp = q == 0 ? p : __fadd_rn(grp, p);
          if (++q == 4) { acc += (double)grp; q = 0; }
        } else acc += (double)p;
      }
    }
    const float o = __double2float_rn(acc);
    raw[i] = o;
    omax = fmaxf(omax, o);
  }
  omax = block_reduce(omax, OpMaxF(), s_f);
  const float s1 = omax > 0.f ? rcp_scale(omax) : 1.f;                   // MatrixNorm(TempOut)
  if (kind < 2) {
    const float m2 = omax > 0.f ? __fmul_rn(omax, s1) : omax;            // MatrixNorm(TempmatL, SpatialMap) after the replication
    const float s2 = m2 > 0.f ? rcp_scale(m2) : 1.f;
    float* out = P.d_sval + ((size_t)b * 3 + (kind == 0 ? 1 : 2)) * P.g32;
    for (int i = tid; i < g; i += SP_T) {
      float o = raw[i];
      if (omax > 0.f) o = __fmul_rn(o, s1);
      if (m2 > 0.f) o = __fmul_rn(o, s2);
      out[i] = o;
    }
  } else if (omax > 0.f) {
    for (int i = tid; i < g; i += SP_T) raw[i] = __fmul_rn(raw[i], s1);
  }
}

// GenerateSpatialMap_mv2 (salmat.cpp:402): MatrixNorm(sqrt(x.^2 + y.^2)) on the 32-pixel grid
__global__ void __launch_bounds__(256) fast_mvs_kernel(FastP P) {
  const int b = blockIdx.x, tid = threadIdx.x, g = P.g32;
  __shared__ float s_f[32];
  const float* rx = P.d_sraw + ((size_t)b * 4 + 2) * g; const float* ry = rx + g;
  float hm = 0.f;
  for (int i = tid; i < g; i += 256) hm = fmaxf(hm, __fsqrt_rn(__fadd_rn(__fmul_rn(rx[i], rx[i]), __fmul_rn(ry[i], ry[i]))));
  hm = block_reduce(hm, OpMaxF(), s_f);
  const float s = hm > 0.f ? rcp_scale(hm) : 1.f;
  float* out = P.d_sval + (size_t)b * 3 * g;
  for (int i = tid; i < g; i += 256) {
    const float hv = __fsqrt_rn(__fadd_rn(__fmul_rn(rx[i], rx[i]), __fmul_rn(ry[i], ry[i])));
    out[i] = hm > 0.f ? __fmul_rn(hv, s) : hv;
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
      // pixel j, tap 4m + i reads cell m + (j + i) / 4: seven distinct products in the 16 (pixel, tap) pairs
      const float x0 = __fmul_rn(c0, k.x), y0 = __fmul_rn(c0, k.y), z0 = __fmul_rn(c0, k.z), w0 = __fmul_rn(c0, k.w);
      const float y1 = __fmul_rn(c1, k.y), z1 = __fmul_rn(c1, k.z), w1 = __fmul_rn(c1, k.w);
      s0 = __fadd_rn(s0, x0); s0 = __fadd_rn(s0, y0); s0 = __fadd_rn(s0, z0); s0 = __fadd_rn(s0, w0);
      s1 = __fadd_rn(s1, x0); s1 = __fadd_rn(s1, y0); s1 = __fadd_rn(s1, z0); s1 = __fadd_rn(s1, w1);
      s2 = __fadd_rn(s2, x0); s2 = __fadd_rn(s2, y0); s2 = __fadd_rn(s2, z1); s2 = __fadd_rn(s2, w1);
      s3 = __fadd_rn(s3, x0); s3 = __fadd_rn(s3, y1); s3 = __fadd_rn(s3, z1); s3 = __fadd_rn(s3, w1);
      c0 = c1;
    }
    reinterpret_cast<float4*>(P.d_hrow + ((size_t)b * P.h4 + r) * P.w)[p] = make_float4(s0, s1, s2, s3);
  }
}

// Column pass (filter.cpp SymmColumnFilter): s = fl(x[0] * k[0]); s = fl(s + fl(fl(x[+t] + x[-t]) * k[t])), t = 1..half.  The row-pass
// output is constant over the 4 pixel rows of a cell row; a thread computes the 4 pixel rows of one cell row q for one column.
// A CTA stages BV_Q cell rows plus the nq_v + 1 rows above and below of a 128-column strip in shared memory.
constexpr int BV_Q = 16;

__global__ void __launch_bounds__(256) fast_blur_v_kernel(FastP P) {
  extern __shared__ __align__(16) float s_t[];            // [BV_Q + 2 * (nq_v + 1)][128] cell rows, then the taps
  const int x0 = blockIdx.x * 128, q0 = blockIdx.y * BV_Q, b = blockIdx.z;
  const int W = P.w, h4 = P.h4, nq = P.nq_v, nr = BV_Q + 2 * (nq + 1);
  float* s_k = s_t + nr * 128;
  __shared__ float s_f[32];
  const float* hr = P.d_hrow + (size_t)b * h4 * W;
  for (int i = threadIdx.x; i < nr * 32; i += 256) {        // float4 columns: W % 4 == 0 and x0 % 128 == 0
    const int rr = i >> 5, c4 = i & 31, q = q0 - (nq + 1) + rr, x = x0 + c4 * 4;
    reinterpret_cast<float4*>(s_t)[i] = (q >= 0 && q < h4 && x < W) ? *reinterpret_cast<const float4*>(hr + (size_t)q * W + x) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < 4 + 4 * nq; i += 256) s_k[i] = P.d_kv[i];
  __syncthreads();
  float lo = INFINITY, hi = -INFINITY;
  const int cc = threadIdx.x & 127, x = x0 + cc;
  for (int ql = threadIdx.x >> 7; ql < BV_Q; ql += 2) {
    const int q = q0 + ql;
    if (q >= h4 || x >= W) continue;
    const float* col = s_t + (ql + nq + 1) * 128 + cc;    // cell row q
    const float ctr = col[0];
    float s0 = __fmul_rn(ctr, s_k[0]), s1 = s0, s2 = s0, s3 = s0;
    float a0 = ctr, b0 = ctr;
#pragma unroll 3
    for (int m = 0; m < nq; m++) {
      const float a1 = col[(m + 1) * 128], b1 = col[-(m + 1) * 128];
      const float4 k = *reinterpret_cast<const float4*>(s_k + 4 + 4 * m);
      // pixel row j, tap t = 4m + i (i = 1..4): above cell m + (j + i) / 4, below cell -(m + (i - j + 3) / 4); only four
      // distinct (above + below) sums occur in the 16 (row, tap) pairs
      const float p00 = __fadd_rn(a0, b0), p01 = __fadd_rn(a0, b1), p10 = __fadd_rn(a1, b0), p11 = __fadd_rn(a1, b1);
      s0 = __fadd_rn(s0, __fmul_rn(p01, k.x)); s0 = __fadd_rn(s0, __fmul_rn(p01, k.y)); s0 = __fadd_rn(s0, __fmul_rn(p01, k.z)); s0 = __fadd_rn(s0, __fmul_rn(p11, k.w));
      s1 = __fadd_rn(s1, __fmul_rn(p00, k.x)); s1 = __fadd_rn(s1, __fmul_rn(p01, k.y)); s1 = __fadd_rn(s1, __fmul_rn(p11, k.z)); s1 = __fadd_rn(s1, __fmul_rn(p11, k.w));
      s2 = __fadd_rn(s2, __fmul_rn(p00, k.x)); s2 = __fadd_rn(s2, __fmul_rn(p10, k.y)); s2 = __fadd_rn(s2, __fmul_rn(p11, k.z)); s2 = __fadd_rn(s2, __fmul_rn(p11, k.w));
      s3 = __fadd_rn(s3, __fmul_rn(p10, k.x)); s3 = __fadd_rn(s3, __fmul_rn(p10, k.y)); s3 = __fadd_rn(s3, __fmul_rn(p10, k.z)); s3 = __fadd_rn(s3, __fmul_rn(p11, k.w));
      a0 = a1; b0 = b1;
    }
    float* out = P.d_blur + ((size_t)b * P.h + 4 * q) * W + x;
    out[0] = s0; out[W] = s1; out[2 * (size_t)W] = s2; out[3 * (size_t)W] = s3;
    lo = fminf(lo, fminf(fminf(s0, s1), fminf(s2, s3))); hi = fmaxf(hi, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
  }
  lo = block_reduce(lo, OpMinF(), s_f); hi = block_reduce(hi, OpMaxF(), s_f);
  if (threadIdx.x == 0 && lo <= hi) { atomicMin(P.d_mm + b * 2, float_key(lo)); atomicMax(P.d_mm + b * 2 + 1, float_key(hi)); }
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
  cudaStream_t side = nullptr;          // the temporal chain (temporal -> tthr) runs beside the threshold / spatial chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  void (*mark)(void* user, int stage, cudaStream_t st) = nullptr;   // per-stage timing hook of the clip context (cds_profile_*)
  void* mark_user = nullptr; bool profiling = false;
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
  p.RL = 2 * B + p.PS + p.PT;           // the front half (camera stages) of batch k+1 may fill its ring slots while the back half of batch k still reads
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
  A(d_comb, Bz * n4); A(d_hrow, Bz * p.h4 * w); A(d_blur, Bz * (size_t)h * w); A(d_mm, Bz * 2); A(d_cami, Bz * 24); A(d_camf, Bz * 2);
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
  if (!e && (cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) != cudaSuccess ||
             cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
             cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess)) { set_error("fast: stream / event creation"); e = CDS_ECUDA; }
  const size_t smem_sm = (size_t)(SM_G + p.PS - 1) * 128 * sizeof(float4);
  if (!e && smem_sm > 200 * 1024) { set_error("fast version: fps %.1f needs a %zu-byte smoothing tile", fps, smem_sm); e = CDS_EINVAL; }
  if (!e) e = ensure_dyn_smem(fast_smooth_kernel, smem_sm);      // monotonic per device: a later, smaller context must not lower it
  if (e) { fast_destroy(s); return e; }
  *out = s;
  return fast_clear(s, nullptr);
}

void fast_destroy(FastState* s) {
  if (!s) return;
  if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
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
  if (int e = fast_front(s, first_poc, n, mvx, mvy, dep, bytes, cam_out, st)) return e;
  return fast_back(s, first_poc, n, out, out_u8, st);
}

// front half: CameraDect for the batch (raw ring slots, per-picture scalars, camera motion) - everything later pictures need from the camera stage
int fast_front(FastState* s, int first_poc, int n, const int16_t* mvx, const int16_t* mvy, const uint8_t* dep, const uint16_t* bytes,
               int* cam_out, cudaStream_t st) {
  const FastP& p = s->p;
  if (n < 1 || n > p.B || first_poc < 0) { set_error("fast_run_batch: bad batch"); return CDS_EINVAL; }
  auto M = [&](int stage) { if (s->profiling && s->mark) s->mark(s->mark_user, stage, st); };
  M(FAST_ST_CAM_SCAN);
  CDS_CUDA(cudaMemsetAsync(p.d_cami, 0, (size_t)n * 24 * sizeof(int), st));
  fast_cam_scan_kernel<<<dim3(ceil_div(p.n4, 1024), n), 256, 0, st>>>(p, first_poc, mvx, mvy, dep, bytes);
  CDS_LAUNCH_CHECK();
  M(FAST_ST_CAM_VOTE);
  fast_cam_vote_kernel<<<dim3(2, n), CAM_T, 0, st>>>(p, first_poc);
  CDS_LAUNCH_CHECK();
  fast_cam_norm_kernel<<<dim3(ceil_div(p.n4 / 4, 256), n), 256, 0, st>>>(p, first_poc);
  CDS_LAUNCH_CHECK();
  fast_cam_final_kernel<<<ceil_div(n, 64), 64, 0, st>>>(p, first_poc, n, cam_out);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// back half: smoothing (also the temporal rings of the block grids) and, when `out` is given, everything up to the map
int fast_back(FastState* s, int first_poc, int n, void* out, int out_u8, cudaStream_t st) {
  const FastP& p = s->p;
  if (n < 1 || n > p.B || first_poc < 0) { set_error("fast_run_batch: bad batch"); return CDS_EINVAL; }
  auto M = [&](int stage) { if (s->profiling && s->mark) s->mark(s->mark_user, stage, st); };
  M(FAST_ST_SMOOTH);
  fast_smooth_kernel<<<dim3(ceil_div(p.n4, 128), ceil_div(n, SM_G)), dim3(128, SM_G), (size_t)(SM_G + p.PS - 1) * 128 * sizeof(float4), st>>>(p, first_poc, n);
  CDS_LAUNCH_CHECK();
  fast_smooth_bit_kernel<<<dim3(ceil_div(p.nctu, 128), n), 128, 0, st>>>(p, first_poc);
  CDS_LAUNCH_CHECK();
  if (!out) return CDS_OK;
  // two independent latency-bound chains after the smoothing: MapThreshold -> spatial maps, and temporal maps -> their threshold
  // (side by side on two streams; one after the other when the stages are being timed)
  cudaStream_t sd = s->profiling ? st : s->side;
  M(FAST_ST_TEMPORAL);
  if (!s->profiling) {
    CDS_CUDA(cudaEventRecord(s->ev_fork, st));
    CDS_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
  }
  fast_temporal_kernel<<<dim3(ceil_div(p.g32, 128), 3, n), 128, 0, sd>>>(p, first_poc);
  CDS_LAUNCH_CHECK();
  fast_tthr_kernel<<<dim3(3, n), 128, s->smem_tthr, sd>>>(p);
  CDS_LAUNCH_CHECK();
  if (!s->profiling) CDS_CUDA(cudaEventRecord(s->ev_join, s->side));
  M(FAST_ST_THR_SPATIAL);
  fast_thr_kernel<<<dim3(5, n), 256, 0, st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_spatial_kernel<<<dim3(4, n), SP_T, s->smem_sp[1], st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_mvs_kernel<<<n, 256, 0, st>>>(p);
  CDS_LAUNCH_CHECK();
  if (!s->profiling) CDS_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
  M(FAST_ST_COMBINE_BLUR_H);
  fast_combine_kernel<<<dim3(ceil_div(p.n4, 256), n), 256, 0, st>>>(p);
  CDS_LAUNCH_CHECK();
  fast_mm_reset_kernel<<<ceil_div(n, 128), 128, 0, st>>>(p.d_mm, n);
  CDS_LAUNCH_CHECK();
  fast_blur_h_kernel<<<dim3(p.h4, n), 256, s->smem_h, st>>>(p);
  CDS_LAUNCH_CHECK();
  M(FAST_ST_BLUR_V);
  fast_blur_v_kernel<<<dim3(ceil_div(p.w, 128), ceil_div(p.h4, BV_Q), n), 256, ((size_t)(BV_Q + 2 * (p.nq_v + 1)) * 128 + 4 + 4 * p.nq_v) * sizeof(float), st>>>(p);
  CDS_LAUNCH_CHECK();
  M(FAST_ST_OUTPUT);
  const int nb = ceil_div((int)(((size_t)p.h * p.w / 4)), 256);
  if (out_u8) fast_output_kernel<uint8_t><<<dim3(nb, n), 256, 0, st>>>(p, reinterpret_cast<uint8_t*>(out));
  else fast_output_kernel<float><<<dim3(nb, n), 256, 0, st>>>(p, reinterpret_cast<float*>(out));
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// ---- probes: the two sequential-sum emulations on caller-supplied data (tests feed them ties, binade crossings, mixed signs) ----
namespace {
__global__ void __launch_bounds__(256) probe_seq_sum_kernel(const float* __restrict__ v, const unsigned char* __restrict__ sel, int ngroups,
                                                            float* __restrict__ out, int* __restrict__ count) {
  __shared__ SeqSumSmem<256> s_seq;
  struct Raw { float4 v; uchar4 s; };
  auto load = [&](int g) -> Raw { return Raw{reinterpret_cast<const float4*>(v)[g], reinterpret_cast<const uchar4*>(sel)[g]}; };
  auto eval = [&](const Raw& r, float* t) -> unsigned {
    t[0] = r.v.x; t[1] = r.v.y; t[2] = r.v.z; t[3] = r.v.w;
    return (r.s.x ? 1u : 0u) | (r.s.y ? 2u : 0u) | (r.s.z ? 4u : 0u) | (r.s.w ? 8u : 0u);
  };
  int cnt = 0;
  const float S = block_seq_sum<256>(ngroups, load, eval, s_seq, &cnt);
  if (threadIdx.x == 0) { *out = S; *count = cnt; }
}
__global__ void __launch_bounds__(32) probe_run_sum_kernel(const float* __restrict__ vals, int gh, int gw, int rows, int cols, int ms, double t,
                                                           float* __restrict__ out, int* __restrict__ count) {
  __shared__ float s_rv[128]; __shared__ int s_rn[128], s_rq[128];
  int cnt = 0;
  const float S = warp_select_sum_runs(vals, gh, gw, rows, cols, ms, t, s_rv, s_rn, s_rq, &cnt);
  if (threadIdx.x == 0) { *out = S; *count = cnt; }
}
}  // namespace

int fast_probe_seq_sum(const float* v, const unsigned char* sel, int n, float* sum, int* count) {
  if (int e = ensure_device()) return e;
  if (!v || !sel || !sum || !count || n < 4 || (n % 4)) { set_error("cds_probe_seq_sum: n must be a positive multiple of 4"); return CDS_EINVAL; }
  float* dv = nullptr; unsigned char* ds = nullptr; float* dout = nullptr;
  CDS_CUDA(cudaMalloc(&dv, (size_t)n * 4)); CDS_CUDA(cudaMalloc(&ds, (size_t)n)); CDS_CUDA(cudaMalloc(&dout, 16));
  CDS_CUDA(cudaMemcpy(dv, v, (size_t)n * 4, cudaMemcpyHostToDevice)); CDS_CUDA(cudaMemcpy(ds, sel, (size_t)n, cudaMemcpyHostToDevice));
  probe_seq_sum_kernel<<<1, 256>>>(dv, ds, n / 4, dout, reinterpret_cast<int*>(dout + 1));
  CDS_LAUNCH_CHECK();
  float h[2];
  CDS_CUDA(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
  *sum = h[0]; memcpy(count, &h[1], 4);
  cudaFree(dv); cudaFree(ds); cudaFree(dout);
  return CDS_OK;
}

int fast_probe_run_sum(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, float* sum, int* count) {
  if (int e = ensure_device()) return e;
  if (!vals || !sum || !count || gh < 1 || gw < 1 || gw > 128 || ms < 1 || rows <= ms * (gh - 1) || rows > ms * gh || cols <= ms * (gw - 1) || cols > ms * gw) {
    set_error("cds_probe_run_sum: bad geometry"); return CDS_EINVAL;
  }
  float* dv = nullptr; float* dout = nullptr;
  CDS_CUDA(cudaMalloc(&dv, (size_t)gh * gw * 4)); CDS_CUDA(cudaMalloc(&dout, 16));
  CDS_CUDA(cudaMemcpy(dv, vals, (size_t)gh * gw * 4, cudaMemcpyHostToDevice));
  probe_run_sum_kernel<<<1, 32>>>(dv, gh, gw, rows, cols, ms, t, dout, reinterpret_cast<int*>(dout + 1));
  CDS_LAUNCH_CHECK();
  float h[2];
  CDS_CUDA(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
  *sum = h[0]; memcpy(count, &h[1], 4);
  cudaFree(dv); cudaFree(dout);
  return CDS_OK;
}

void fast_set_marker(FastState* s, void (*mark)(void*, int, cudaStream_t), void* user) { s->mark = mark; s->mark_user = user; }
void fast_profile(FastState* s, bool on) { s->profiling = on; }

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
