// vote.cu — stage (c): camera-motion detection and removal.
//
// Replaces main.m:77-105 with cm_judge.m, mv_vote.m, direction_cluster.m (and the unused
// magnitude_vote.m as an optional op).  One CTA per picture, everything in exact integer arithmetic
// until the final division, so the voted camera motion is bit-exact:
//   * medfilt2(V,[2 2]) of V = mv/4 is (m1+m2)/8 with m1,m2 the middle two of four quarter-pel ints
//     (zero padded, window anchored top-left) -> integer key k = m1+m2 in 1/8-pixel units;
//   * the 16-bin direction histogram (direction_cluster.m:9-17) is built from exact octant tests
//     ((|x|+|y|)^2 vs 2x^2 decides the pi/8 boundaries; rational boundaries fall where the
//     reference's double arithmetic puts them) with warp-aggregated ballots into shared memory;
//   * the half-sorted means (direction_cluster.m:25-50) are sums of the m smallest keys, found by a
//     two-level radix select on shared-memory histograms (no sort), exact in int64;
//   * x_ave = S/8/count is one double division -> identical to MATLAB's sum(...)/n.
#include "cds_common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#include <math.h>
#include <mutex>
#include <vector>

namespace cds {

constexpr int CM_THREADS = 512;      // 2 CTAs per SM (60 registers): more 8-CTA clusters in flight (1024 threads: 0.56 ms per 64 pictures, 512: 0.45, 256: 0.49)

__device__ __forceinline__ int med4_key(int a, int b, int c, int d) {
  // sum of the two middle values of four
  const int lo1 = min(a, b), hi1 = max(a, b), lo2 = min(c, d), hi2 = max(c, d);
  const int mn = min(lo1, lo2), mx = max(hi1, hi2);
  return a + b + c + d - mn - mx;
}

// floor((atan2(y,x)/pi + 1) * 8) with 16 -> 15, exactly (direction_cluster.m:9-11)
__device__ __forceinline__ int dir_bin16(int x, int y) {
  if (y == 0) return x < 0 ? 15 : 8;
  if (x == 0) return y > 0 ? 12 : 4;
  const long long ax = x < 0 ? -(long long)x : x, ay = y < 0 ? -(long long)y : y;
  const long long s2 = (ax + ay) * (ax + ay);
  const int q = (s2 >= 2 * ax * ax) + (ay >= ax) + (s2 <= 2 * ay * ay);   // floor(alpha / (pi/8)), alpha = atan(|y|/|x|)
  const bool diag = ax == ay;                                              // alpha == pi/4 exactly
  if (x > 0) return y > 0 ? 8 + q : (diag ? 6 : 7 - q);
  return y > 0 ? (diag ? 14 : 15 - q) : q;
}

struct CmShared {
  int hist[16];
  unsigned hx[512], hy[512];      // radix-select histograms (counts)
  long long red_ll[32 * 2];
  int red_i[32 * 4];
  float red_f[32];
  int bmin, bmax, dmax, N, T, best, flag;
  int xmin, xmax, ymin, ymax;
  int pivx_hi, pivy_hi, belowx, belowy;     // level-1 pivot bin and count below it
  int pivx, pivy, lessx, lessy;             // pivot key (offset form) and count of keys < pivot
  long long sumx_all, sumy_all, sumx_less, sumy_less;
  double x_ave, y_ave;
  float mvmax;
};

__device__ __forceinline__ int block_reduce_sum_i(int v, int* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int r = 0;
  for (int i = 0; i < nw; i++) r += red[i];
  return r;
}
__device__ __forceinline__ long long block_reduce_sum_ll(long long v, long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  long long r = 0;
  for (int i = 0; i < nw; i++) r += red[i];
  return r;
}
__device__ __forceinline__ int block_reduce_min_i(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int r = red[0];
  for (int i = 1; i < nw; i++) r = min(r, red[i]);
  return r;
}
__device__ __forceinline__ int block_reduce_max_i(int v, int* red) { return -block_reduce_min_i(-v, red); }
__device__ __forceinline__ float block_reduce_max_f(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; i++) r = fmaxf(r, red[i]);
  return r;
}

struct CmCell {
  int kx, ky;     // medfilt keys (1/8 px)
  bool eff;       // background cell: Bm < min(Bm)+20  (main.m:81)
};

__device__ __forceinline__ CmCell cm_cell(const int16_t* __restrict__ mvx, const int16_t* __restrict__ mvy,
                                          const uint16_t* __restrict__ byts, int h4, int w4, int ncol, int bmin, int o) {
  const int r = o / w4, c = o - r * w4;
  const bool dn = r + 1 < h4, rt = c + 1 < w4;
  CmCell out;
  out.kx = med4_key(mvx[o], dn ? mvx[o + w4] : 0, rt ? mvx[o + 1] : 0, (dn && rt) ? mvx[o + w4 + 1] : 0);
  out.ky = med4_key(mvy[o], dn ? mvy[o + w4] : 0, rt ? mvy[o + 1] : 0, (dn && rt) ? mvy[o + w4 + 1] : 0);
  out.eff = (int)byts[(r >> 4) * ncol + (c >> 4)] < bmin + 20;
  return out;
}

// level-1/level-2 radix-select pivot search over a 512/256-bin count histogram: smallest bin b with
// cumulative count > rank (0-based rank of the wanted order statistic)
__device__ __forceinline__ void find_pivot(const unsigned* hist, int nbins, int rank, int* bin_out, int* below_out) {
  int cum = 0, b = 0;
  for (; b < nbins; b++) { if (cum + (int)hist[b] > rank) break; cum += hist[b]; }
  *bin_out = b < nbins ? b : nbins - 1; *below_out = cum;
}

// warp-parallel radix-select pivot search (same result as find_pivot): executed by one full warp
__device__ __forceinline__ void find_pivot_warp(const unsigned* hist, int nbins, int rank, int* bin_out, int* below_out) {
  const int lane = threadIdx.x & 31, per = nbins >> 5;          // nbins is 512 or 256
  int mine = 0;
  for (int i = 0; i < per; i++) mine += (int)hist[lane * per + i];
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const int excl = incl - mine;
  const unsigned holds = __ballot_sync(0xffffffffu, excl <= rank && rank < incl);
  if (holds == 0) { if (lane == 0) { *bin_out = nbins - 1; *below_out = __shfl_sync(0xffffffffu, incl, 31); } return; }
  const int owner = __ffs(holds) - 1;
  if (lane == owner) {
    int cum = excl, b = lane * per;
    for (; b < lane * per + per; b++) { if (cum + (int)hist[b] > rank) break; cum += hist[b]; }
    *bin_out = b; *below_out = cum;
  }
}

struct CmCluster {                 // cluster-wide accumulators: the copy in the rank-0 CTA is the one everybody adds to / reads
  int hist[16];
  unsigned hx[512], hy[512];       // level-1 radix-select histograms
  unsigned hx2[256], hy2[256];     // level-2
  int dmax, N, T;
  int xmin, xmax, ymin, ymax;
  unsigned long long sumx_all, sumy_all, sumx_less, sumy_less;   // two's-complement sums
  int pivx_hi, pivy_hi, belowx, belowy, pivx, pivy, lessx, lessy;
  unsigned vmax_bits;
};

// grid: CM_CLUSTER CTAs (one thread-block cluster) per picture f; the CTAs split the cells, partial results meet in the
// rank-0 CTA's shared memory through distributed-shared-memory atomics, and every CTA derives the (identical) decisions.
// VT = double for the clip pipeline: the smoothed mv maps feed Sudoku_contrastcpp5's `p != 0` test after a
// pm_norm, so ties at the map minimum must round exactly as the reference's double arithmetic does.
constexpr int CM_CLUSTER = 8;
template <typename VT>
__global__ void __cluster_dims__(CM_CLUSTER, 1, 1) __launch_bounds__(CM_THREADS)
camera_motion_kernel(int w, int h, const int16_t* __restrict__ mvx_all, const int16_t* __restrict__ mvy_all,
                     const uint16_t* __restrict__ bytes_all, const uint8_t* __restrict__ depth_all,
                     VT* __restrict__ Vx_ring, VT* __restrict__ Vy_ring, float* __restrict__ mvn_ring,
                     float* __restrict__ bitn_ring, float* __restrict__ depn_ring, int* __restrict__ cam_ring,
                     double* __restrict__ ave_out, int* __restrict__ flag_out, int first_slot, int ring_len) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ CmShared S;           // CTA-local scratch (reductions, local histograms in S.hx / S.hy)
  __shared__ CmCluster Csm;
  CmCluster* C0 = cluster.map_shared_rank(&Csm, 0);
  const int crank = (int)cluster.block_rank();
  const int f = blockIdx.x / CM_CLUSTER;
  const int h4 = h >> 2, w4 = w >> 2, ncol = (w + 63) >> 6, nctu = ncol * ((h + 63) >> 6);
  const int n4 = h4 * w4;
  const int16_t* __restrict__ mvx = mvx_all + (size_t)f * n4;
  const int16_t* __restrict__ mvy = mvy_all + (size_t)f * n4;
  const uint16_t* __restrict__ byts = bytes_all + (size_t)f * nctu;
  const uint8_t* __restrict__ dep = depth_all + (size_t)f * n4;
  const int slot = (first_slot + f) % ring_len;
  VT* Vx = Vx_ring + (size_t)slot * n4; VT* Vy = Vy_ring + (size_t)slot * n4;
  float* mvn = mvn_ring + (size_t)slot * n4; float* bitn = bitn_ring + (size_t)slot * n4; float* depn = depn_ring + (size_t)slot * n4;
  const int tid = threadIdx.x, lane = tid & 31;
  const int chunk = (n4 + CM_CLUSTER - 1) / CM_CLUSTER;
  const int o_lo = min(n4, crank * chunk), o_hi = min(n4, o_lo + chunk);      // this CTA's cells

  // ---- zero the accumulators (only rank 0's copy is used, every CTA clears its own) ----
  if (tid < 16) Csm.hist[tid] = 0;
  for (int i = tid; i < 512; i += CM_THREADS) { Csm.hx[i] = 0; Csm.hy[i] = 0; }
  for (int i = tid; i < 256; i += CM_THREADS) { Csm.hx2[i] = 0; Csm.hy2[i] = 0; }
  if (tid == 0) {
    Csm.dmax = 0; Csm.N = 0; Csm.T = 0;
    Csm.xmin = 1 << 30; Csm.xmax = -(1 << 30); Csm.ymin = 1 << 30; Csm.ymax = -(1 << 30);
    Csm.sumx_all = 0; Csm.sumy_all = 0; Csm.sumx_less = 0; Csm.sumy_less = 0; Csm.vmax_bits = 0;
  }
  cluster.sync();

  // ---- bytes min/max over CTUs (tiny: every CTA does all of it), depth max over this CTA's cells ----
  int bmn = 1 << 30, bmx = 0, dmx = 0;
  for (int i = tid; i < nctu; i += CM_THREADS) { const int b = byts[i]; bmn = min(bmn, b); bmx = max(bmx, b); }
  for (int i = o_lo + tid; i < o_hi; i += CM_THREADS) dmx = max(dmx, (int)dep[i]);
  bmn = block_reduce_min_i(bmn, S.red_i);
  bmx = block_reduce_max_i(bmx, S.red_i);
  dmx = block_reduce_max_i(dmx, S.red_i);

  // ---- pass 1: background count N, static count T, direction histogram ----
  int N = 0, T = 0, hacc = 0;
  for (int o0 = o_lo; o0 < o_hi; o0 += CM_THREADS) {
    const int o = o0 + tid;
    bool eff = false; int bin = -1;
    if (o < o_hi) {
      const CmCell c = cm_cell(mvx, mvy, byts, h4, w4, ncol, bmn, o);
      eff = c.eff;
      if (eff) {
        N++;
        T += ((long long)c.kx * c.kx + (long long)c.ky * c.ky) <= 128;       // Vx^2+Vy^2 <= 2  (cm_judge.m:5-6)
        bin = dir_bin16(c.kx, c.ky);
      }
    }
#pragma unroll
    for (int b = 0; b < 16; b++) {                                            // warp-aggregated histogram
      const unsigned m = __ballot_sync(0xffffffffu, bin == b);
      if (lane == b) hacc += __popc(m);
    }
  }
  if (lane < 16 && hacc) atomicAdd(&C0->hist[lane], hacc);
  N = block_reduce_sum_i(N, S.red_i);
  T = block_reduce_sum_i(T, S.red_i);
  if (tid == 0) { atomicAdd(&C0->N, N); atomicAdd(&C0->T, T); atomicMax(&C0->dmax, dmx); }
  cluster.sync();
  if (tid == 0) {
    N = C0->N; T = C0->T;
    S.dmax = C0->dmax;
    S.flag = !(2 * (long long)T > N);                                         // cm_judge.m:8-12 (T > 0.5*N -> static)
    int hv[16];
    for (int k = 0; k < 16; k++) hv[k] = C0->hist[k];
    int best = 0;
    for (int k = 1; k < 16; k++) if (hv[k] > hv[best]) best = k;              // first maximum
    S.best = best; S.N = hv[best]; S.x_ave = 0; S.y_ave = 0;
  }
  __syncthreads();
  const int flag = S.flag, best = S.best, Len = S.N;
  dmx = S.dmax;

  if (flag && Len > 0) {                                                      // uniform over the cluster
    // ---- pass 2: level-1 histograms (high 9 bits of key+65536), min/max, total sums ----
    for (int i = tid; i < 512; i += CM_THREADS) { S.hx[i] = 0; S.hy[i] = 0; }
    __syncthreads();
    int xmn = 1 << 30, xmx = -(1 << 30), ymn = 1 << 30, ymx = -(1 << 30);
    long long sx = 0, sy = 0;
    for (int o = o_lo + tid; o < o_hi; o += CM_THREADS) {
      const CmCell c = cm_cell(mvx, mvy, byts, h4, w4, ncol, bmn, o);
      if (c.eff && dir_bin16(c.kx, c.ky) == best) {
        xmn = min(xmn, c.kx); xmx = max(xmx, c.kx); ymn = min(ymn, c.ky); ymx = max(ymx, c.ky);
        sx += c.kx; sy += c.ky;
        atomicAdd(&S.hx[(c.kx + 65536) >> 8], 1u);
        atomicAdd(&S.hy[(c.ky + 65536) >> 8], 1u);
      }
    }
    xmn = block_reduce_min_i(xmn, S.red_i); xmx = block_reduce_max_i(xmx, S.red_i);
    ymn = block_reduce_min_i(ymn, S.red_i); ymx = block_reduce_max_i(ymx, S.red_i);
    sx = block_reduce_sum_ll(sx, S.red_ll); sy = block_reduce_sum_ll(sy, S.red_ll);
    __syncthreads();
    for (int i = tid; i < 512; i += CM_THREADS) {
      if (S.hx[i]) atomicAdd(&C0->hx[i], S.hx[i]);
      if (S.hy[i]) atomicAdd(&C0->hy[i], S.hy[i]);
    }
    if (tid == 0) {
      atomicMin(&C0->xmin, xmn); atomicMax(&C0->xmax, xmx); atomicMin(&C0->ymin, ymn); atomicMax(&C0->ymax, ymx);
      atomicAdd(&C0->sumx_all, (unsigned long long)sx); atomicAdd(&C0->sumy_all, (unsigned long long)sy);
    }
    cluster.sync();
    xmn = C0->xmin; xmx = C0->xmax; ymn = C0->ymin; ymx = C0->ymax;
    sx = (long long)C0->sumx_all; sy = (long long)C0->sumy_all;
    // number of smallest keys to sum: direction_cluster.m:40-50
    const int Lencut = Len / 2;
    const bool topx = abs(xmn) > abs(xmx), topy = abs(ymn) > abs(ymx);
    const int mx_ = topx ? Lencut - 1 : Lencut, my_ = topy ? Lencut - 1 : Lencut;    // sum of the m smallest
    if (crank == 0 && tid < 32) {
      if (mx_ > 0) find_pivot_warp(Csm.hx, 512, mx_ - 1, &Csm.pivx_hi, &Csm.belowx); else if (tid == 0) { Csm.pivx_hi = -1; Csm.belowx = 0; }
      if (my_ > 0) find_pivot_warp(Csm.hy, 512, my_ - 1, &Csm.pivy_hi, &Csm.belowy); else if (tid == 0) { Csm.pivy_hi = -1; Csm.belowy = 0; }
    }
    cluster.sync();
    const int phx = C0->pivx_hi, phy = C0->pivy_hi;
    // ---- pass 3: level-2 histograms (low 8 bits) inside the pivot bins ----
    for (int i = tid; i < 256; i += CM_THREADS) { S.hx[i] = 0; S.hy[i] = 0; }
    __syncthreads();
    for (int o = o_lo + tid; o < o_hi; o += CM_THREADS) {
      const CmCell c = cm_cell(mvx, mvy, byts, h4, w4, ncol, bmn, o);
      if (c.eff && dir_bin16(c.kx, c.ky) == best) {
        const int ux = c.kx + 65536, uy = c.ky + 65536;
        if ((ux >> 8) == phx) atomicAdd(&S.hx[ux & 255], 1u);
        if ((uy >> 8) == phy) atomicAdd(&S.hy[uy & 255], 1u);
      }
    }
    __syncthreads();
    for (int i = tid; i < 256; i += CM_THREADS) {
      if (S.hx[i]) atomicAdd(&C0->hx2[i], S.hx[i]);
      if (S.hy[i]) atomicAdd(&C0->hy2[i], S.hy[i]);
    }
    cluster.sync();
    if (crank == 0 && tid < 32) {
      int b = 0, below = 0;
      if (phx >= 0) { find_pivot_warp(Csm.hx2, 256, mx_ - 1 - Csm.belowx, &S.pivx, &S.lessx); __syncwarp(); if (tid == 0) { b = S.pivx; below = S.lessx; Csm.pivx = (phx << 8) | b; Csm.lessx = Csm.belowx + below; } }
      else if (tid == 0) { Csm.pivx = 0; Csm.lessx = 0; }
      __syncwarp();
      if (phy >= 0) { find_pivot_warp(Csm.hy2, 256, my_ - 1 - Csm.belowy, &S.pivy, &S.lessy); __syncwarp(); if (tid == 0) { b = S.pivy; below = S.lessy; Csm.pivy = (phy << 8) | b; Csm.lessy = Csm.belowy + below; } }
      else if (tid == 0) { Csm.pivy = 0; Csm.lessy = 0; }
    }
    cluster.sync();
    const int pvx = C0->pivx, pvy = C0->pivy, lessx = C0->lessx, lessy = C0->lessy;
    // ---- pass 4: exact sums of the keys strictly below the pivots ----
    long long lx = 0, ly = 0;
    for (int o = o_lo + tid; o < o_hi; o += CM_THREADS) {
      const CmCell c = cm_cell(mvx, mvy, byts, h4, w4, ncol, bmn, o);
      if (c.eff && dir_bin16(c.kx, c.ky) == best) {
        if (c.kx + 65536 < pvx) lx += c.kx;
        if (c.ky + 65536 < pvy) ly += c.ky;
      }
    }
    lx = block_reduce_sum_ll(lx, S.red_ll); ly = block_reduce_sum_ll(ly, S.red_ll);
    if (tid == 0) { atomicAdd(&C0->sumx_less, (unsigned long long)lx); atomicAdd(&C0->sumy_less, (unsigned long long)ly); }
    cluster.sync();
    if (tid == 0) {
      lx = (long long)C0->sumx_less; ly = (long long)C0->sumy_less;
      // sum of the m smallest = (keys < pivot) + (m - #less) copies of the pivot
      const long long smx = mx_ > 0 ? lx + (long long)(mx_ - lessx) * (pvx - 65536) : 0;
      const long long smy = my_ > 0 ? ly + (long long)(my_ - lessy) * (pvy - 65536) : 0;
      double xa, ya;
      if (Len < 2) { xa = (double)sx / 8.0; ya = (double)sy / 8.0; }          // reference undefined (index 0 / 0-division): the value itself
      else {
        xa = topx ? ((double)(sx - smx) / 8.0) / (double)(Len - Lencut + 1) : ((double)smx / 8.0) / (double)Lencut;
        ya = topy ? ((double)(sy - smy) / 8.0) / (double)(Len - Lencut + 1) : ((double)smy / 8.0) / (double)Lencut;
      }
      S.x_ave = xa; S.y_ave = ya;
    }
    __syncthreads();
  }
  const double x_ave = S.x_ave, y_ave = S.y_ave;

  // ---- pass 5: compensated field, |V| max ----
  float vmax = 0.f;
  for (int o = o_lo + tid; o < o_hi; o += CM_THREADS) {
    double dvx, dvy;
    if (flag) {
      const CmCell c = cm_cell(mvx, mvy, byts, h4, w4, ncol, bmn, o);
      dvx = (double)c.kx * 0.125 - x_ave; dvy = (double)c.ky * 0.125 - y_ave;                   // Vxf - x_ave  (main.m:89-90)
    } else {
      dvx = (double)mvx[o] * 0.25; dvy = (double)mvy[o] * 0.25;                                 // unfiltered (main.m:78-79,93)
    }
    Vx[o] = (VT)dvx; Vy[o] = (VT)dvy;
    const float vx = (float)dvx, vy = (float)dvy;
    const float m = sqrtf(vx * vx + vy * vy);
    mvn[o] = m;
    vmax = fmaxf(vmax, m);
  }
  vmax = block_reduce_max_f(vmax, S.red_f);
  if (tid == 0) atomicMax(&C0->vmax_bits, __float_as_uint(vmax));              // vmax >= 0: bit patterns order like the values
  cluster.sync();
  vmax = __uint_as_float(C0->vmax_bits);
  const float eps = 2.220446049250313e-16f;
  const float inv_b = 1.f / ((float)bmx + eps), inv_d = 1.f / ((float)dmx + eps);
  for (int o = o_lo + tid; o < o_hi; o += CM_THREADS) {
    const int r = o / w4, c = o - r * w4;
    mvn[o] = mvn[o] / (vmax + eps);                                            // main.m:102
    bitn[o] = (float)byts[(r >> 4) * ncol + (c >> 4)] * inv_b;                // main.m:103
    depn[o] = (float)dep[o] * inv_d;                                           // main.m:104
  }
  if (crank == 0 && tid == 0) {
    // round() half away from zero (main.m:98-99)
    cam_ring[2 * slot + 0] = (int)(x_ave < 0 ? -floor(-x_ave + 0.5) : floor(x_ave + 0.5));
    cam_ring[2 * slot + 1] = (int)(y_ave < 0 ? -floor(-y_ave + 0.5) : floor(y_ave + 0.5));
    if (ave_out) { ave_out[2 * f] = x_ave; ave_out[2 * f + 1] = y_ave; }
    if (flag_out) flag_out[f] = flag;
  }
  cluster.sync();                  // rank 0's shared memory must outlive every remote access
}

int camera_motion_launch(int w, int h, int nframes, const int16_t* mvx, const int16_t* mvy, const uint16_t* byts,
                         const uint8_t* depth, float* Vx, float* Vy, float* mvn, float* bitn, float* depn, int* cam,
                         double* ave, int* flag, int first_slot, int ring_len, cudaStream_t st) {
  camera_motion_kernel<float><<<nframes * CM_CLUSTER, CM_THREADS, 0, st>>>(w, h, mvx, mvy, byts, depth, Vx, Vy, mvn, bitn, depn, cam, ave,
                                                              flag, first_slot, ring_len);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}
int camera_motion_launch_f64(int w, int h, int nframes, const int16_t* mvx, const int16_t* mvy, const uint16_t* byts,
                             const uint8_t* depth, double* Vx, double* Vy, float* mvn, float* bitn, float* depn, int* cam,
                             double* ave, int* flag, int first_slot, int ring_len, cudaStream_t st) {
  camera_motion_kernel<double><<<nframes * CM_CLUSTER, CM_THREADS, 0, st>>>(w, h, mvx, mvy, byts, depth, Vx, Vy, mvn, bitn, depn, cam, ave,
                                                               flag, first_slot, ring_len);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// mv_vote on an explicit vector list (1/8-px integer keys): same algorithm, one CTA.
__global__ void __launch_bounds__(CM_THREADS)
mv_vote_kernel(const int* __restrict__ vx8, const int* __restrict__ vy8, int n, double* __restrict__ out) {
  __shared__ int hist[16];
  __shared__ unsigned hx[512], hy[512];
  __shared__ int red_i[32];
  __shared__ long long red_ll[32];
  __shared__ int sh[8];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 16) hist[tid] = 0;
  __syncthreads();
  int hacc = 0;
  for (int o0 = 0; o0 < n; o0 += CM_THREADS) {
    const int o = o0 + tid;
    const int bin = o < n ? dir_bin16(vx8[o], vy8[o]) : -1;
#pragma unroll
    for (int b = 0; b < 16; b++) { const unsigned m = __ballot_sync(0xffffffffu, bin == b); if (lane == b) hacc += __popc(m); }
  }
  if (lane < 16 && hacc) atomicAdd(&hist[lane], hacc);
  __syncthreads();
  int best = 0;
  for (int k = 1; k < 16; k++) if (hist[k] > hist[best]) best = k;
  const int Len = hist[best];
  if (Len == 0) { if (tid == 0) { out[0] = 0; out[1] = 0; } return; }
  double res[2];
  for (int q = 0; q < 2; q++) {
    const int* v = q == 0 ? vx8 : vy8;
    unsigned* hh = q == 0 ? hx : hy;
    __syncthreads();
    for (int i = tid; i < 512; i += CM_THREADS) hh[i] = 0;
    __syncthreads();
    int mn = 1 << 30, mx = -(1 << 30); long long s = 0;
    for (int o = tid; o < n; o += CM_THREADS)
      if (dir_bin16(vx8[o], vy8[o]) == best) { const int k = v[o]; mn = min(mn, k); mx = max(mx, k); s += k; atomicAdd(&hh[(k + 65536) >> 8], 1u); }
    mn = block_reduce_min_i(mn, red_i); mx = block_reduce_max_i(mx, red_i); s = block_reduce_sum_ll(s, red_ll);
    const int Lencut = Len / 2; const bool top = abs(mn) > abs(mx);
    const int m = top ? Lencut - 1 : Lencut;
    long long sm = 0;
    if (m > 0) {
      if (tid == 0) find_pivot(hh, 512, m - 1, &sh[0], &sh[1]);
      __syncthreads();
      const int ph = sh[0], below = sh[1];
      __syncthreads();
      for (int i = tid; i < 256; i += CM_THREADS) hh[i] = 0;
      __syncthreads();
      for (int o = tid; o < n; o += CM_THREADS)
        if (dir_bin16(vx8[o], vy8[o]) == best) { const int u = v[o] + 65536; if ((u >> 8) == ph) atomicAdd(&hh[u & 255], 1u); }
      __syncthreads();
      if (tid == 0) { int b, bl; find_pivot(hh, 256, m - 1 - below, &b, &bl); sh[2] = (ph << 8) | b; sh[3] = below + bl; }
      __syncthreads();
      const int pv = sh[2], less = sh[3];
      long long l = 0;
      for (int o = tid; o < n; o += CM_THREADS)
        if (dir_bin16(vx8[o], vy8[o]) == best && v[o] + 65536 < pv) l += v[o];
      l = block_reduce_sum_ll(l, red_ll);
      sm = l + (long long)(m - less) * (pv - 65536);
    }
    if (Len < 2) res[q] = (double)s / 8.0;
    else res[q] = top ? ((double)(s - sm) / 8.0) / (double)(Len - Lencut + 1) : ((double)sm / 8.0) / (double)Lencut;
  }
  if (tid == 0) { out[0] = res[0]; out[1] = res[1]; }
}

// magnitude_vote.m:1-27 — 20-class value histogram, mean of the modal class(es).  Tiny; one CTA, FP64.
__global__ void __launch_bounds__(256)
magnitude_vote_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double sred[256];
  __shared__ int ired[256];
  __shared__ double range[64];
  __shared__ int label[64];
  __shared__ int slen;
  const int tid = threadIdx.x;
  double mn = INFINITY, mx = -INFINITY;
  for (int i = tid; i < n; i += 256) { mn = fmin(mn, v[i]); mx = fmax(mx, v[i]); }
  sred[tid] = mn; __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (tid < s) sred[tid] = fmin(sred[tid], sred[tid + s]); __syncthreads(); }
  mn = sred[0]; __syncthreads();
  sred[tid] = mx; __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (tid < s) sred[tid] = fmax(sred[tid], sred[tid + s]); __syncthreads(); }
  mx = sred[0]; __syncthreads();
  const double step = (mx - mn + 1) / 20;                                      // :5
  if (tid == 0) {
    int cnt = (int)floor((mx - mn) / step + 1e-10) + 1;                        // numel(min:step:max)
    if (cnt > 64) cnt = 64;
    for (int i = 0; i < cnt; i++) range[i] = mn + i * step;
    slen = cnt;
  }
  __syncthreads();
  const int len = slen;
  for (int ii = 0; ii < len; ii++) {                                           // :10-15: #{range(ii) <= v < range(ii+1)}
    int c = 0;
    for (int i = tid; i < n; i += 256) c += (v[i] >= range[ii]) && (ii == len - 1 || !(v[i] >= range[ii + 1]));
    ired[tid] = c; __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (tid < s) ired[tid] += ired[tid + s]; __syncthreads(); }
    if (tid == 0) label[ii] = ired[0];
    __syncthreads();
  }
  int lmax = label[0];
  for (int ii = 1; ii < len; ii++) lmax = max(lmax, label[ii]);
  double v_ave = 0; int nmax = 0;
  for (int ii = 0; ii < len; ii++) {
    if (label[ii] != lmax) continue;
    double s = 0; int c = 0;
    for (int i = tid; i < n; i += 256)
      if (v[i] >= range[ii] && (ii == len - 1 || v[i] < range[ii + 1])) { s += v[i]; c++; }
    sred[tid] = s; ired[tid] = c; __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (tid < st) { sred[tid] += sred[tid + st]; ired[tid] += ired[tid + st]; } __syncthreads(); }
    v_ave += sred[0] / ired[0]; nmax++;
    __syncthreads();
  }
  if (tid == 0) out[0] = v_ave / nmax;
}

static std::mutex g_cm_mu;
static DevBuf g_b[12];

}  // namespace cds

using namespace cds;

extern "C" int cds_camera_motion(int w, int h, const int16_t* mvx, const int16_t* mvy, const uint16_t* ctu_bytes,
                                 const uint8_t* depth, float* Vx, float* Vy, float* mv_norm, float* bit_norm,
                                 float* depth_norm, int* cam, double* ave, int* flag) {
  if (int e = ensure_device()) return e;
  if (!mvx || !mvy || !ctu_bytes || !depth || !Vx || !Vy || !mv_norm || !bit_norm || !depth_norm || !cam ||
      w < 8 || h < 8 || (w % 8) || (h % 8)) { set_error("cds_camera_motion: bad arguments"); return CDS_EINVAL; }
  std::lock_guard<std::mutex> lk(g_cm_mu);
  const size_t n4 = (size_t)(h / 4) * (w / 4), nctu = (size_t)((w + 63) / 64) * ((h + 63) / 64);
  const size_t sizes[12] = {n4 * 2, n4 * 2, nctu * 2, n4, n4 * 4, n4 * 4, n4 * 4, n4 * 4, n4 * 4, 8, 16, 4};
  for (int i = 0; i < 12; i++) if (int e = g_b[i].reserve(sizes[i])) return e;
  cudaStream_t st = 0;
  CDS_CUDA(cudaMemcpyAsync(g_b[0].p, mvx, n4 * 2, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemcpyAsync(g_b[1].p, mvy, n4 * 2, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemcpyAsync(g_b[2].p, ctu_bytes, nctu * 2, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemcpyAsync(g_b[3].p, depth, n4, cudaMemcpyHostToDevice, st));
  if (int e = camera_motion_launch(w, h, 1, g_b[0].as<int16_t>(), g_b[1].as<int16_t>(), g_b[2].as<uint16_t>(), g_b[3].as<uint8_t>(),
                                   g_b[4].as<float>(), g_b[5].as<float>(), g_b[6].as<float>(), g_b[7].as<float>(), g_b[8].as<float>(),
                                   g_b[9].as<int>(), g_b[10].as<double>(), g_b[11].as<int>(), 0, 1, st)) return e;
  float* outs[5] = {Vx, Vy, mv_norm, bit_norm, depth_norm};
  for (int i = 0; i < 5; i++) CDS_CUDA(cudaMemcpyAsync(outs[i], g_b[4 + i].p, n4 * 4, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaMemcpyAsync(cam, g_b[9].p, 8, cudaMemcpyDeviceToHost, st));
  double have[2]; int hflag = 0;
  CDS_CUDA(cudaMemcpyAsync(have, g_b[10].p, 16, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaMemcpyAsync(&hflag, g_b[11].p, 4, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  if (ave) { ave[0] = have[0]; ave[1] = have[1]; }
  if (flag) *flag = hflag;
  return CDS_OK;
}

extern "C" int cds_mv_vote(const int32_t* vx8, const int32_t* vy8, int n, double* x_ave, double* y_ave) {
  if (int e = ensure_device()) return e;
  if (!vx8 || !vy8 || !x_ave || !y_ave || n < 0) { set_error("cds_mv_vote: bad arguments"); return CDS_EINVAL; }
  for (int i = 0; i < n; i++)
    if (vx8[i] < -65536 || vx8[i] > 65535 || vy8[i] < -65536 || vy8[i] > 65535) {
      set_error("cds_mv_vote: key %d outside the 1/8-pel range of int16 motion vectors", i); return CDS_EINVAL;
    }
  if (n == 0) { *x_ave = 0; *y_ave = 0; return CDS_OK; }
  std::lock_guard<std::mutex> lk(g_cm_mu);
  if (int e = g_b[0].reserve((size_t)n * 4)) return e;
  if (int e = g_b[1].reserve((size_t)n * 4)) return e;
  if (int e = g_b[10].reserve(16)) return e;
  cudaStream_t st = 0;
  CDS_CUDA(cudaMemcpyAsync(g_b[0].p, vx8, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemcpyAsync(g_b[1].p, vy8, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  mv_vote_kernel<<<1, CM_THREADS, 0, st>>>(g_b[0].as<int>(), g_b[1].as<int>(), n, g_b[10].as<double>());
  CDS_LAUNCH_CHECK();
  double r[2];
  CDS_CUDA(cudaMemcpyAsync(r, g_b[10].p, 16, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  *x_ave = r[0]; *y_ave = r[1];
  return CDS_OK;
}

extern "C" int cds_magnitude_vote(const double* value, int n, double* v_ave) {
  if (int e = ensure_device()) return e;
  if (!value || !v_ave || n < 1) { set_error("cds_magnitude_vote: bad arguments"); return CDS_EINVAL; }
  std::lock_guard<std::mutex> lk(g_cm_mu);
  if (int e = g_b[0].reserve((size_t)n * 8)) return e;
  if (int e = g_b[10].reserve(16)) return e;
  cudaStream_t st = 0;
  CDS_CUDA(cudaMemcpyAsync(g_b[0].p, value, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  magnitude_vote_kernel<<<1, 256, 0, st>>>(g_b[0].as<double>(), n, g_b[10].as<double>());
  CDS_LAUNCH_CHECK();
  CDS_CUDA(cudaMemcpyAsync(v_ave, g_b[10].p, 8, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  return CDS_OK;
}
