// getframe.cu — stage (a): per-CTU syntax tokens -> 4x4-block mv_x / mv_y / depth maps.
//
// Replaces getFrame() (HM_16.0_features/source/Lib/TLibDecoder/code_mat_h.h:575-649) with its helpers
// batch_64 (:55-451), cupu_f (:452), my_data (:464-573) and func.h.  The reference paints a 64x64
// label map per CTU with ~10 full sweeps per PU and then samples every 4th pixel; here one warp
// owns one CTU and works directly on the 16x16 grid of 4x4 cells that is actually sampled:
//   1. the warp walks the depth-first CU token stream in lockstep (<= 85 tokens) and paints, per
//      leaf CU, the PU label each sampled pixel would carry after the reference's PU split
//      (:281-385) and PartSize-15 compaction (:386-450);
//   2. my_data's stream expansion is evaluated in closed form per PU (lanes in parallel);
//   3. the sequential in-place label->MV substitution (:595-603) becomes a per-cell chain walk.
// All integer quirks of the reference are reproduced (see oracle/cds_oracle.c for the derivation);
// outputs are bit-exact.  Integer/latency-bound, tiny next to stages (b)/(d); batched over frames.
#include "cds_common.cuh"
#include <mutex>

namespace cds {

constexpr int GF_WARPS = 4;
constexpr int GF_MAX_CUPU = 96;
constexpr int GF_MAX_PRED = 256;
constexpr int GF_MAX_MV = 1280;
constexpr int GF_EXT_CAP = 2048;   // same caps as the oracle (the reference overflows a 1000-slot array instead)
constexpr int GF_MAX_X = 500;

struct GfWarpSmem {
  int16_t cupu[GF_MAX_CUPU];
  int16_t pred[GF_MAX_PRED];
  int16_t mv[GF_MAX_MV];
  int16_t label[256];
  uint8_t depth[256];
  int16_t pux[512];
  int16_t puy[512];
  int16_t mvlen[GF_MAX_PRED];   // raw-stream offset of intra PU ii   (mv_length, code_mat_h.h:520)
  int16_t epos[GF_MAX_PRED];    // its slot in the expanded stream     (mv_length[ii] + 3*ii, :525)
};

__device__ __forceinline__ int gf_pu_index(int part, int s, int r, int c) {
  switch (part) {                                   // code_mat_h.h:340-384 reduced to geometry
    case 1: return r >= s / 2;
    case 2: return (r + c * s) >= (s * s) / 2;
    case 3: return (r >= s / 2 ? 2 : 0) + (c >= s / 2 ? 1 : 0);
    case 4: return r >= s / 4;
    case 5: return r >= s / 4 * 3;
    case 6: return (r + c * s) >= (s * s) / 4;
    case 7: return (r + c * s) >= (s * s) / 4 * 3;
    default: return 0;
  }
}
__device__ __forceinline__ int gf_nlabels(int part) {
  if (part <= 0 || part == 99) return 1;
  if (part == 3) return 4;
  if (part == 15) return 1;
  return 2;
}
__device__ __forceinline__ int gf_tokread(const int16_t* mv, int n, int i) {
  if (i < 0) return 0;
  if (i < n) return mv[i];
  return i == n ? 999 : 0;
}

// value of the expanded MV stream at slot i (my_data, code_mat_h.h:477-563)
__device__ __forceinline__ int gf_ext_at(const GfWarpSmem& S, int i, int nmv, int npred, int T, int a0) {
  if (T == 0) return gf_tokread(S.mv, nmv, i);
  if (T == 1) {
    if (i < 4 * a0) return gf_tokread(S.mv, nmv, i);
    if (i < 4 * a0 + 3) return 0;
    if (i == 4 * a0 + 3) return gf_tokread(S.mv, nmv, a0);        // QUIRK :500
    return gf_tokread(S.mv, nmv, i - 3);
  }
  int lo = 0, hi = T;                                              // q = #{ii : epos[ii] <= i}
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (S.epos[mid] <= i) lo = mid + 1; else hi = mid; }
  const int q = lo;
  if (q > 0 && i < S.epos[q - 1] + 4) return (i - S.epos[q - 1] < 3) ? 0 : gf_tokread(S.mv, nmv, S.mvlen[q - 1]);
  if (q == T) {                                                     // tail, QUIRK :554
    if (S.mvlen[T - 1] + 1 != npred) return gf_tokread(S.mv, nmv, i - 3 * T);
    return 0;
  }
  return gf_tokread(S.mv, nmv, i - 3 * q);
}

__device__ __forceinline__ int gf_subst(int L, const int16_t* pu, int x) {   // getFrame :595-603
  if (L < 1 || L > x) return L;
  int cur = L - 1, v = pu[cur];
  while (v - 1 > cur && v - 1 < x) { cur = v - 1; v = pu[cur]; }
  return v;
}

__global__ void __launch_bounds__(GF_WARPS * 32)
getframe_kernel(int w, int h, int nframes, const int32_t* __restrict__ hdr, const int16_t* __restrict__ tok,
                const int64_t* __restrict__ tok_base, int16_t* __restrict__ mvx, int16_t* __restrict__ mvy,
                uint8_t* __restrict__ depth, int* __restrict__ err) {
  __shared__ GfWarpSmem smem[GF_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncol = (w + 63) >> 6, nrow = (h + 63) >> 6, nctu = ncol * nrow;
  const long gw = (long)blockIdx.x * GF_WARPS + warp;
  if (gw >= (long)nframes * nctu) return;
  const int f = (int)(gw / nctu), line = (int)(gw - (long)f * nctu);
  GfWarpSmem& S = smem[warp];
  const int32_t* hd = hdr + ((size_t)f * nctu + line) * 4;
  const int off = hd[0], ncupu = hd[1], npred = hd[2], nmv = hd[3];
  const int h4 = h >> 2, w4 = w >> 2;
  const int R40 = (line / ncol) * 16, C40 = (line % ncol) * 16;
  bool bad = ncupu < 1 || ncupu > GF_MAX_CUPU || npred < 0 || npred > GF_MAX_PRED || nmv < 0 || nmv > GF_MAX_MV || off < 0;
  // the three streams must lie inside this picture's slice of the token buffer (a corrupt offset must not read a neighbour's tokens)
  if (!bad && tok_base) bad = (long)off + ncupu + npred + nmv > (long)(tok_base[f + 1] - tok_base[f]);
  if (!bad) {
    const int16_t* src = tok + (tok_base ? tok_base[f] : 0) + off;
    for (int i = lane; i < ncupu; i += 32) S.cupu[i] = src[i];
    for (int i = lane; i < npred; i += 32) S.pred[i] = src[ncupu + i];
    for (int i = lane; i < nmv; i += 32) S.mv[i] = src[ncupu + npred + i];
    for (int i = lane; i < 256; i += 32) { S.label[i] = 0; S.depth[i] = 0; }
    __syncwarp();
    // ---- 1. lockstep walk of the CU quad-tree (z-order: TL, TR, BL, BR) ----
    const bool unsplit = S.cupu[0] != 99;
    int st_r[12], st_c[12], st_s[12], sp = 0;
    st_r[0] = 0; st_c[0] = 0; st_s[0] = 64; sp = 1;
    int pos = 0, next = 1, next15 = 1;
    while (sp > 0) {
      --sp;
      const int r0 = st_r[sp], c0 = st_c[sp], size = st_s[sp];
      if (pos >= ncupu) { bad = true; break; }
      int t = S.cupu[pos++];
      if (t == 99) {
        if (size <= 8) { bad = true; break; }
        const int s = size >> 1;
        st_r[sp] = r0 + s; st_c[sp] = c0 + s; st_s[sp] = s; sp++;
        st_r[sp] = r0 + s; st_c[sp] = c0;     st_s[sp] = s; sp++;
        st_r[sp] = r0;     st_c[sp] = c0 + s; st_s[sp] = s; sp++;
        st_r[sp] = r0;     st_c[sp] = c0;     st_s[sp] = s; sp++;
      } else {
        if (unsplit) t = 0;                    // QUIRK func.h:177: an unsplit CTU's PartSize is ignored
        const int nl = gf_nlabels(t);
        const int base = next, lab15 = next15;
        next15 += nl;
        next += (t == 15) ? 0 : nl;
        const int d = size == 64 ? 0 : size == 32 ? 1 : size == 16 ? 2 : 3;
        const int n4 = size >> 2;
        for (int idx = lane; idx < n4 * n4; idx += 32) {
          const int rr = idx / n4, cc = idx - rr * n4;
          const int L = (t == 15) ? -lab15 : base + gf_pu_index(t, size, rr * 4, cc * 4);
          const int cell = ((r0 >> 2) + rr) * 16 + (c0 >> 2) + cc;
          S.label[cell] = (int16_t)L; S.depth[cell] = (uint8_t)d;
        }
      }
    }
    if (!bad && pos != ncupu) bad = true;
  }
  if (bad) {
    if (lane == 0 && err) atomicOr(err, 1);
    for (int cell = lane; cell < 256; cell += 32) {
      const int R4 = R40 + (cell >> 4), C4 = C40 + (cell & 15);
      if (R4 < h4 && C4 < w4) {
        const size_t o = ((size_t)f * h4 + R4) * w4 + C4;
        mvx[o] = 0; mvy[o] = 0; depth[o] = 0;
      }
    }
    return;
  }
  // ---- 2. my_data: intra bookkeeping by warp scan, then closed-form expansion per PU ----
  int T = 0, inter_run = 0, a0 = 0;
  for (int p0 = 0; p0 < npred; p0 += 32) {
    const int p = p0 + lane;
    const int v = p < npred ? S.pred[p] : 0;
    const unsigned m_intra = __ballot_sync(0xffffffffu, v == 2);
    const unsigned m_inter = __ballot_sync(0xffffffffu, v == 1);
    const unsigned below = (1u << lane) - 1u;
    if (v == 2) {
      const int ii = T + __popc(m_intra & below);
      const int ninter = inter_run + __popc(m_inter & below);
      const int ml = 4 * ninter + ii;
      S.mvlen[ii] = (int16_t)ml; S.epos[ii] = (int16_t)(ml + 3 * ii);
      if (ii == 0) S.pux[0] = (int16_t)p;      // remember addr_intra[0] (scratch; overwritten below)
    }
    T += __popc(m_intra); inter_run += __popc(m_inter);
  }
  __syncwarp();
  if (T > 0) a0 = S.pux[0];
  __syncwarp();
  int ext_l = nmv + 3 * T;
  if (ext_l > GF_EXT_CAP) ext_l = GF_EXT_CAP;
  int x = ext_l >> 2;
  if (x > GF_MAX_X) x = GF_MAX_X;
  for (int j = lane; j < x; j += 32) {
    S.pux[j] = (int16_t)gf_ext_at(S, 4 * j + 2, nmv, npred, T, a0);
    S.puy[j] = (int16_t)gf_ext_at(S, 4 * j + 3, nmv, npred, T, a0);
  }
  __syncwarp();
  // ---- 3. substitute and store: a lane owns two adjacent cells (w4 is even), so mv_x / mv_y leave as 4-byte and depth as
  //         2-byte vector stores; 8 lanes cover one 16-cell row of the CTU (32 contiguous bytes per mv map) ----
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int cell = 2 * (lane + 32 * k);
    const int R4 = R40 + (cell >> 4), C4 = C40 + (cell & 15);
    if (R4 < h4 && C4 < w4) {
      const int L0 = S.label[cell], L1 = S.label[cell + 1];
      const size_t o = ((size_t)f * h4 + R4) * w4 + C4;
      *reinterpret_cast<short2*>(mvx + o) = make_short2((short)gf_subst(L0, S.pux, x), (short)gf_subst(L1, S.pux, x));
      *reinterpret_cast<short2*>(mvy + o) = make_short2((short)gf_subst(L0, S.puy, x), (short)gf_subst(L1, S.puy, x));
      *reinterpret_cast<uchar2*>(depth + o) = make_uchar2(S.depth[cell], S.depth[cell + 1]);
    }
  }
}

int getframe_launch(int w, int h, int nframes, const int32_t* hdr, const int16_t* tok, const int64_t* tok_base,
                    int16_t* mvx, int16_t* mvy, uint8_t* depth, int* err, cudaStream_t st) {
  const long warps = (long)nframes * ((w + 63) / 64) * ((h + 63) / 64);
  const unsigned grid = (unsigned)((warps + GF_WARPS - 1) / GF_WARPS);
  getframe_kernel<<<grid, GF_WARPS * 32, 0, st>>>(w, h, nframes, hdr, tok, tok_base, mvx, mvy, depth, err);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

static std::mutex g_gf_mu;
static DevBuf g_hdr, g_tok, g_mx, g_my, g_dp, g_err;

}  // namespace cds

using namespace cds;

static int geometry_ok(int w, int h) {
  if (w < 8 || h < 8 || (w % 8) || (h % 8) || w > 8192 || h > 8192) {
    set_error("picture size %dx%d unsupported (multiples of 8, <= 8192)", w, h); return 0;
  }
  return 1;
}

extern "C" int cds_getframe_dev(int w, int h, int nframes, const int32_t* hdr, const int16_t* tok,
                                const int64_t* tok_base, int16_t* mvx, int16_t* mvy, uint8_t* depth, void* stream) {
  if (int e = ensure_device()) return e;
  if (!geometry_ok(w, h)) return CDS_EINVAL;
  if (!hdr || !tok || !mvx || !mvy || !depth || nframes < 1) { set_error("cds_getframe_dev: bad arguments"); return CDS_EINVAL; }
  return getframe_launch(w, h, nframes, hdr, tok, tok_base, mvx, mvy, depth, nullptr, (cudaStream_t)stream);
}

extern "C" int cds_getframe(int w, int h, const int32_t* hdr, const int16_t* tok, size_t ntok,
                            int16_t* mvx, int16_t* mvy, uint8_t* depth) {
  if (int e = ensure_device()) return e;
  if (!geometry_ok(w, h)) return CDS_EINVAL;
  if (!hdr || !tok || !mvx || !mvy || !depth) { set_error("cds_getframe: null argument"); return CDS_EINVAL; }
  const int nctu = ((w + 63) / 64) * ((h + 63) / 64);
  for (int a = 0; a < nctu; a++) {
    const long end = (long)hdr[a * 4] + hdr[a * 4 + 1] + hdr[a * 4 + 2] + hdr[a * 4 + 3];
    if (hdr[a * 4] < 0 || hdr[a * 4 + 1] < 0 || hdr[a * 4 + 2] < 0 || hdr[a * 4 + 3] < 0 || end > (long)ntok) {
      set_error("cds_getframe: CTU %d header points outside the token buffer", a); return CDS_EFORMAT;
    }
  }
  std::lock_guard<std::mutex> lk(g_gf_mu);
  const size_t ncell = (size_t)(h / 4) * (w / 4);
  if (int e = g_hdr.reserve((size_t)nctu * 16)) return e;
  if (int e = g_tok.reserve(ntok * 2 + 16)) return e;
  if (int e = g_mx.reserve(ncell * 2)) return e;
  if (int e = g_my.reserve(ncell * 2)) return e;
  if (int e = g_dp.reserve(ncell)) return e;
  if (int e = g_err.reserve(4)) return e;
  cudaStream_t st = 0;
  CDS_CUDA(cudaMemcpyAsync(g_hdr.p, hdr, (size_t)nctu * 16, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemcpyAsync(g_tok.p, tok, ntok * 2, cudaMemcpyHostToDevice, st));
  CDS_CUDA(cudaMemsetAsync(g_err.p, 0, 4, st));
  if (int e = getframe_launch(w, h, 1, g_hdr.as<int32_t>(), g_tok.as<int16_t>(), nullptr, g_mx.as<int16_t>(),
                              g_my.as<int16_t>(), g_dp.as<uint8_t>(), g_err.as<int>(), st)) return e;
  int herr = 0;
  CDS_CUDA(cudaMemcpyAsync(mvx, g_mx.p, ncell * 2, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaMemcpyAsync(mvy, g_my.p, ncell * 2, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaMemcpyAsync(depth, g_dp.p, ncell, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaMemcpyAsync(&herr, g_err.p, 4, cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  if (herr) { set_error("cds_getframe: malformed CU token stream"); return CDS_EFORMAT; }
  return CDS_OK;
}

// ---- "semantic" stage (a): the hook hands over the CTU's 256 z-scan partitions (depth, prediction mode, L0 / L1
// motion vector of 4x4-pixel units, what TComDataCU::getDepth / getPredictionMode / getCUMvField hold) instead of the token streams
// of TDecCu::MVoutput; the kernel is a z-scan -> raster transpose.  Depth is identical to getFrame's; the motion maps differ from
// the reference's in the few percent of cells where getFrame's aliasing quirks (code_mat_h.h:500,533) fire - an intentional deviation.
namespace {
__device__ __forceinline__ unsigned morton4(unsigned x, unsigned y) {      // HM g_auiRasterToZscan for 4x4 units inside a 64x64 CTU
  x = (x | (x << 2)) & 0x33u; x = (x | (x << 1)) & 0x55u;
  y = (y | (y << 2)) & 0x33u; y = (y | (y << 1)) & 0x55u;
  return x | (y << 1);
}
__global__ void __launch_bounds__(256) semantic_scatter_kernel(int w4, int h4, int ncol, int nctu, const uint8_t* __restrict__ depth256,
                                                               const uint8_t* __restrict__ pred256, const int16_t* __restrict__ mv256,
                                                               int16_t* __restrict__ mvx, int16_t* __restrict__ mvy, uint8_t* __restrict__ depth) {
  const int f = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x, n4 = w4 * h4;
  if (k >= n4) return;
  const int r = k / w4, c = k - r * w4;
  const size_t part = ((size_t)f * nctu + (r >> 4) * ncol + (c >> 4)) * 256 + morton4(c & 15, r & 15);
  const bool inter = pred256[part] == 0;                                 // MODE_INTER (TypeDef.h:440)
  const size_t o = (size_t)f * n4 + k;
  mvx[o] = inter ? mv256[part * 2] : (int16_t)0;
  mvy[o] = inter ? mv256[part * 2 + 1] : (int16_t)0;
  depth[o] = depth256[part];
}
}  // namespace

extern "C" int cds_getframe_semantic_dev(int w, int h, int nframes, const uint8_t* depth256, const uint8_t* pred256, const int16_t* mv256,
                                         int16_t* mvx, int16_t* mvy, uint8_t* depth, void* stream) {
  if (int e = ensure_device()) return e;
  if (!geometry_ok(w, h)) return CDS_EINVAL;
  if (!depth256 || !pred256 || !mv256 || !mvx || !mvy || !depth || nframes < 1) { set_error("cds_getframe_semantic_dev: bad arguments"); return CDS_EINVAL; }
  const int w4 = w / 4, h4 = h / 4, ncol = (w + 63) / 64, nctu = ncol * ((h + 63) / 64);
  semantic_scatter_kernel<<<dim3(ceil_div(w4 * h4, 256), nframes), 256, 0, (cudaStream_t)stream>>>(w4, h4, ncol, nctu, depth256, pred256, mv256, mvx, mvy, depth);
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

extern "C" int cds_getframe_semantic(int w, int h, int nframes, const uint8_t* depth256, const uint8_t* pred256, const int16_t* mv256,
                                     int16_t* mvx, int16_t* mvy, uint8_t* depth) {
  if (int e = ensure_device()) return e;
  if (!geometry_ok(w, h)) return CDS_EINVAL;
  if (!depth256 || !pred256 || !mv256 || !mvx || !mvy || !depth || nframes < 1) { set_error("cds_getframe_semantic: bad arguments"); return CDS_EINVAL; }
  const size_t nctu = (size_t)((w + 63) / 64) * ((h + 63) / 64), np = (size_t)nframes * nctu * 256, nc = (size_t)nframes * (w / 4) * (h / 4);
  uint8_t *d_d = nullptr, *d_p = nullptr, *o_d = nullptr; int16_t *d_m = nullptr, *o_x = nullptr, *o_y = nullptr;
  int rc = CDS_OK;
  if (cudaMalloc(&d_d, np) != cudaSuccess || cudaMalloc(&d_p, np) != cudaSuccess || cudaMalloc(&d_m, np * 4) != cudaSuccess ||
      cudaMalloc(&o_d, nc) != cudaSuccess || cudaMalloc(&o_x, nc * 2) != cudaSuccess || cudaMalloc(&o_y, nc * 2) != cudaSuccess) {
    set_error("cds_getframe_semantic: cudaMalloc"); rc = CDS_ENOMEM;
  }
  if (!rc && (cudaMemcpy(d_d, depth256, np, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(d_p, pred256, np, cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpy(d_m, mv256, np * 4, cudaMemcpyHostToDevice) != cudaSuccess)) { set_error("cds_getframe_semantic: copy in"); rc = CDS_ECUDA; }
  if (!rc) rc = cds_getframe_semantic_dev(w, h, nframes, d_d, d_p, d_m, o_x, o_y, o_d, nullptr);
  if (!rc && (cudaMemcpy(mvx, o_x, nc * 2, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(mvy, o_y, nc * 2, cudaMemcpyDeviceToHost) != cudaSuccess ||
              cudaMemcpy(depth, o_d, nc, cudaMemcpyDeviceToHost) != cudaSuccess)) { set_error("cds_getframe_semantic: copy out"); rc = CDS_ECUDA; }
  cudaFree(d_d); cudaFree(d_p); cudaFree(d_m); cudaFree(o_d); cudaFree(o_x); cudaFree(o_y);
  return rc;
}
