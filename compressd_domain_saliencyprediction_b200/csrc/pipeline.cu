// pipeline.cu — per-clip context and batch orchestration of the whole hot path
// (main.m:77-305 with the RBF-SVM branch, or the linear branch; TDecGop.cpp:225-333 hook surface).
//
// One context owns every device buffer for batches of up to B pictures plus ring buffers holding the
// temporal history (PS raw pictures for smoothing, PT smoothed pictures for the temporal features).
// A batch is ~35 kernel launches on one stream; nothing synchronises with the host inside
// cds_clip_process_dev.  Stages: (a) getframe.cu, (c) vote.cu, (e) maps.cu, (b) contrast.cu,
// (d) svm.cu, final map below.
#include "maps.cuh"
#include "fast.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>

struct cds_svm_model;

namespace cds {
// from the other translation units
int getframe_launch(int w, int h, int nframes, const int32_t* hdr, const int16_t* tok, const int64_t* tok_base,
                    int16_t* mvx, int16_t* mvy, uint8_t* depth, int* err, cudaStream_t st);
int camera_motion_launch_f64(int w, int h, int nframes, const int16_t* mvx, const int16_t* mvy, const uint16_t* byts,
                             const uint8_t* depth, double* Vx, double* Vy, float* mvn, float* bitn, float* depn, int* cam,
                             double* ave, int* flag, int first_slot, int ring_len, cudaStream_t st);
int contrast_sep_launch(const float* pad, float* out, unsigned* mx, int nmaps, int R, int C, int win,
                        const float* gr, const float* gc, float* dW_scratch, cudaStream_t st, int ctas_per_sm, unsigned* sched);
void gaussian_factor(int win, double delta, float* g);
int contrast_generic_launch(const float* pad, const float* dW, float* out, unsigned* mx, int nmaps,
                            int R, int C, int win, int len, cudaStream_t st);
int svm_predict_launch(const cds_svm_model* m, const float* X, long N, float* dec, cudaStream_t st);
}  // namespace cds

using namespace cds;

static const float kFeatureWeight[9] = {0.058f, 0.171f, -0.095f, 0.068f, -0.01f, 0.509f, -0.051f, 0.154f, 0.196f};   // main.m:115

struct cds_ctx {
  cds_params p;
  // fast version, device-resident entry point: front half (stage a + camera stages) of batch k+1 on fast_front_stream beside the back half of batch k
  cudaStream_t fast_front_stream = nullptr; cudaEvent_t ev_ff[2] = {nullptr, nullptr}, ev_fb[2] = {nullptr, nullptr}, ev_fstart = nullptr;
  int ff_flip = 0; bool fast_ovl = false, fast_ovl_active = false;
  bool maps_in = false;                // cds_clip_process_maps: d_mvx / d_mvy / d_dep already hold stage (a)'s output
  FastState* fast = nullptr;           // use_rbf == CDS_FUSION_FAST: the reference's fast version (fast.cu) replaces stages (c)-(e), (b), (d)
  int w, h, h4, w4, n4, ncol, nrow, nctu, B, RL, PS, PT, gk;
  double gs;
  int next_poc = 0;
  long HW;
  // operators
  Polyphase poly;                      // fourX2 + gaussian, both directions (square kernel)
  Banded Av, Ah;                       // imresize(.,[200 200]) o imfilter o fourX2 on h4 / w4 cells
  Banded Rv, Rh;                       // imresize(.,[200 200]) on H / W pixels (temporal features)
  Banded Bv, Bh;                       // imfilter o imresize(.,[H W]) from 200 (final map)
  float *d_dist = nullptr, *d_wtab = nullptr, *d_w5 = nullptr;
  // stage (a) batch outputs
  int16_t *d_mvx = nullptr, *d_mvy = nullptr; uint8_t* d_dep = nullptr; int* d_err = nullptr;
  // rings [RL][n4]
  double *r_Vx = nullptr, *r_Vy = nullptr;            // mv_x / mv_y stay double up to the contrast's pm_norm (zero-test ties)
  float *r_mvn = nullptr, *r_bitn = nullptr, *r_depn = nullptr;
  float4* r_sm4 = nullptr;             // smoothed {mvx, mvy, bit, depth} interleaved: the temporal pass reads all four with one LDG.128
  int* r_cam = nullptr;                // [RL][2]
  // batch buffers
  float *d_mv_s = nullptr, *d_X = nullptr;           // [B][n4], [B][9][n4]
  double* d_cxy = nullptr;                            // smoothed mvx | mvy of the batch, [2][B][n4]
  unsigned long long* d_keys64 = nullptr;             // [2B][2] double min/max keys
  float *d_subB = nullptr, *d_padB = nullptr, *d_rawB = nullptr;   // bit contrast (17x30 @1080p)
  float *d_subD = nullptr, *d_padD = nullptr, *d_rawD = nullptr;   // depth contrast
  float *d_padM = nullptr, *d_rawM = nullptr;        // mvx | mvy: [2B]
  unsigned *d_keys = nullptr, *d_maxbits = nullptr;   // [4][B][2], [4][B]
  unsigned* d_sched = nullptr;                        // 2 x 2048 words: tile counter + per-SM CTA counts of the two capped contrast launches
  float *d_T = nullptr;                // [9B][h4][W]
  float *d_Y = nullptr;                // [3B or 9B][H][W]
  int* d_store_index = nullptr;        // [9B]
  float *d_partial = nullptr, *d_stats9 = nullptr;   // [9B][nblk][4], [9B][4]
  float *d_partial2 = nullptr, *d_xf = nullptr;      // temporal thresholds: [3B][nblk2][4], [3B][4]
  float *d_A1 = nullptr, *d_F200 = nullptr;          // [9B][200][w4], [9B][200][200]
  float *d_R1 = nullptr, *d_Ft = nullptr;            // [3B][200][W], [3B][200][200]
  float *d_xsvm = nullptr, *d_dec = nullptr;         // [B][40000][9], [B][40000]
  float* d_xsvm2[2] = {nullptr, nullptr};            // d_xsvm and a second copy: the back half of batch i (SVM -> map) overlaps the front half of batch i+1
  cudaStream_t back_stream = nullptr;
  cudaEvent_t ev_front[2] = {nullptr, nullptr}, ev_back[2] = {nullptr, nullptr}, ev_mid = nullptr;
  int xs_flip = 0, last_xs = 0; bool overlap = true;
  // Pre stage on its own stream (RBF branch, overlap on): stages (a), (c), smoothing, thresholds, temporal and the contrast inputs of batch
  // i+1 — latency- and bandwidth-bound kernels — run beside the map passes of batch i.  The batch-local buffers that cross from the pre
  // stage to the main stage exist twice (set = batch parity): d_X, the three padded contrast inputs.
  cudaStream_t pre_stream = nullptr; cudaEvent_t ev_pre[2] = {nullptr, nullptr}, ev_c48 = nullptr, ev_call = nullptr;
  float* d_X2[2] = {nullptr, nullptr}; float* d_padB2[2] = {nullptr, nullptr}; float* d_padD2[2] = {nullptr, nullptr}; float* d_padM2[2] = {nullptr, nullptr};
  bool pre_ovl = true, call_open = false; int last_set = 0;
  cudaStream_t last_pre_stream = nullptr;          // the stream the most recent batch's pre stage ran on (its camera-motion gather included)
  // Deferred back half (RBF branch, overlap on): the SVM + final-map chain of batch i is launched from the MIDDLE of batch
  // i+1's front half, right before its contrast kernels, so that the MUFU-bound persistent SVM kernel and the FP32-bound
  // contrast kernel start together and share every SM for their whole duration (or from join_back when no batch follows).
  struct PendingBack { bool valid = false; int slot = 0, n = 0, tag = -1; void* maps_out = nullptr; cudaEvent_t before_back = nullptr; } pb;
  int retired_tag = -1; cudaEvent_t retired_ev = nullptr;     // set when a back half has been enqueued: its tag and completion event
  // CDS_TRACE=1: timed events around the co-scheduled kernels of the first batches, printed by cds_destroy (development aid)
  bool trace = false; std::vector<cudaEvent_t> tr_ev; int tr_batches = 0;
  float *d_fparam = nullptr, *d_C = nullptr;         // [9][4] mean, 1/std, weight; [B][200][200]
  float *d_T2 = nullptr, *d_raw = nullptr;           // [B][200][W], [B][H][W]
  float *d_fpartial = nullptr, *d_fstats = nullptr;  // final min/max
  void* d_out[2] = {nullptr, nullptr};   // two [B][H][W] f32 or u8 output buffers: batch i+1 renders while batch i drains to the host
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  int last_buf = 0, out_flip = 0;
  int* d_cam_stage = nullptr;            // [B][2] camera motion of one batch (cds_clip_process_maps: one batch in flight at a time)
  int* d_cam_call = nullptr; size_t d_cam_cap = 0;   // [nframes][2] camera motion of a whole cds_clip_process call: every batch gathers into its
                                                     // own slice, so no batch can overwrite values a queued device->host copy has not read yet
  int* h_cam_pin = nullptr; size_t h_cam_cap = 0;   // pinned staging: a device->host copy into pageable memory would block the host and serialise the pipeline
  // staging for host-buffer entry points
  uint16_t* d_bytes = nullptr; int32_t* d_hdr = nullptr; int16_t* d_tok = nullptr; int64_t* d_tokbase = nullptr;
  size_t tok_cap = 0;
  // ---- hook API state (cds_acquire_frame_buffers / cds_submit_frame / cds_get_map / cds_flush) ----
  // Two PINNED staging batches: the decoder thread writes picture k+1's tokens straight into one while the other is on its way to
  // the device (asynchronous H2D + kernel chain on the stream; nothing waits for the GPU until a map is asked for).
  struct HookStage {
    uint16_t* bytes = nullptr; int32_t* hdr = nullptr; int16_t* tok = nullptr; int64_t* base = nullptr;   // cudaHostAlloc
    size_t tok_cap = 0, tok_used = 0; int n = 0, first = 0; cudaEvent_t h2d_done = nullptr; bool in_flight = false;
  } hs[2];
  // Two PINNED result batches (maps + camera motion): a batch's maps stay retrievable until two further batches were launched.
  struct HookOut {
    void* maps = nullptr; int* cam = nullptr; int first = -1, n = 0; cudaEvent_t ready = nullptr, cam_ready = nullptr; bool copy_queued = false;
  } ho[2];
  int* d_cam_hook = nullptr;           // [2][B][2]
  int hs_cur = 0, ho_cur = 0; bool hook_ready = false, acquired = false;
  size_t hook_pic_tok = 0;             // worst-case tokens of one picture (1300 per CTU)
  int pending = 0, pending_first = 0, last_first = -1, last_n = 0;   // pending: pictures submitted but not launched yet
  cudaStream_t stream = nullptr;
  int nblk = 0, nblk2 = 0, nblkf = 0;
  std::vector<void*> allocs;
  // per-stage device timing (cds_profile_*): events recorded on the launching stream around each stage
  bool prof_on = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<std::pair<int, cudaEvent_t>> ev_marks;     // (stage id that STARTS here, event); id -1 closes a batch
  double prof_ms[24] = {0};
  long prof_calls[24] = {0};
};

enum Stage { ST_GETFRAME = 0, ST_VOTE, ST_SMOOTH, ST_THRESH, ST_TEMPORAL, ST_CONTRAST_PREP, ST_CONTRAST_W5, ST_CONTRAST_W24,
             ST_CONTRAST_W48, ST_CONTRAST_POST, ST_FULLRES_H, ST_FULLRES_V, ST_TEMPORAL_STATS, ST_RESIZE200, ST_SVM, ST_FINAL,
             ST_FAST0 /* + FastStage: the fast version's stages (fast.cuh) */, ST_COUNT = ST_FAST0 + FAST_ST_COUNT };
static_assert(ST_COUNT <= 24, "prof_ms / prof_calls hold 24 stages");
static const char* kStageNames = "getframe,vote,smooth,thresh_small,temporal,contrast_prep,contrast_win5,contrast_win24,contrast_win48,"
                                 "contrast_post,fullres_hpass,fullres_vpass,temporal_stats,resize200,svm_rbf,final_map,"
                                 "fast_cam_scan,fast_cam_vote,fast_smooth,fast_temporal,fast_thr_spatial,fast_combine_blur_h,fast_blur_v,fast_output";

static void prof_mark(cds_ctx* c, int stage, cudaStream_t st) {
  if (!c->prof_on) return;
  if (c->ev_used == c->ev_pool.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; c->ev_pool.push_back(e); }
  cudaEvent_t e = c->ev_pool[c->ev_used++];
  cudaEventRecord(e, st);
  c->ev_marks.push_back({stage, e});
}

namespace {

enum { TR_MID = 0, TR_C48_START, TR_C48_END, TR_FRONT_END, TR_SVM_START, TR_SVM_END, TR_BACK_END, TR_N };
constexpr int TR_MAX_BATCHES = 24;
void trace_mark(cds_ctx* c, int batch, int which, cudaStream_t st) {
  if (!c->trace || batch < 0 || batch >= TR_MAX_BATCHES) return;
  if (c->tr_ev.empty()) { c->tr_ev.resize((size_t)TR_MAX_BATCHES * TR_N, nullptr); for (auto& e : c->tr_ev) cudaEventCreate(&e); }
  cudaEventRecord(c->tr_ev[(size_t)batch * TR_N + which], st);
}
void trace_dump(cds_ctx* c) {
  if (!c->trace || c->tr_ev.empty()) return;
  cudaDeviceSynchronize();
  static const char* nm[TR_N] = {"mid", "c48_start", "c48_end", "front_end", "svm_start", "svm_end", "back_end"};
  for (int b = 1; b < std::min(c->tr_batches, TR_MAX_BATCHES); b++) {
    fprintf(stderr, "[cds trace] batch %2d (ms after its mid-point; svm/back = previous batch's):", b);
    for (int k = 0; k < TR_N; k++) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, c->tr_ev[(size_t)b * TR_N + TR_MID], c->tr_ev[(size_t)b * TR_N + k]) == cudaSuccess) fprintf(stderr, " %s %.3f", nm[k], t);
    }
    fprintf(stderr, "\n");
  }
  cudaGetLastError();
  for (auto& e : c->tr_ev) cudaEventDestroy(e);
  c->tr_ev.clear();
}

template <typename T>
int dalloc(cds_ctx* c, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); return CDS_ENOMEM; }
  c->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return CDS_OK;
}

double mround(double x) { return x < 0 ? -floor(-x + 0.5) : floor(x + 0.5); }

// ------------------------------------------------------------------------------------------------
// small kernels private to the pipeline
// ------------------------------------------------------------------------------------------------
__global__ void fill_u32_kernel(unsigned* p, size_t n, unsigned v0, unsigned v1) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (i & 1) ? v1 : v0;
}

__global__ void gather_cam_kernel(const int* __restrict__ ring, int RL, int first_poc, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) out[i] = ring[2 * ((first_poc + (i >> 1)) % RL) + (i & 1)];
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v0, unsigned long long v1) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (i & 1) ? v1 : v0;
}

// temporal features: pm_norm -> NaN->0 -> mysterythrehold statistics on the stored full-res maps (main.m:255-257)
__global__ void __launch_bounds__(256)
temporal_thresh_partial_kernel(const float* __restrict__ Y, long HW, const float* __restrict__ stats9, float* __restrict__ partial2,
                               int nine_stored /*1: Y holds all nine maps per picture (linear branch)*/) {
  __shared__ double redd[8];
  __shared__ int redi[8];
  const int tm = blockIdx.y;                     // f*3 + j
  const int f = tm / 3, j = tm - 3 * f;
  const float* st = stats9 + (size_t)(f * 9 + 1 + 3 * j) * 4;
  const float mn = st[0], mx = st[1]; const double sum = reinterpret_cast<const double*>(st + 2)[0];
  const bool valid = mx > mn;
  const float inv = valid ? 1.f / (mx - mn) : 0.f;
  const double sumn = valid ? (sum - (double)HW * mn) / ((double)mx - (double)mn) : 0.0;
  const float lim = (valid ? 1.f : 0.f) - (float)(sumn / (double)HW);       // max(Yn) - ave
  const float* __restrict__ y = Y + (size_t)(nine_stored ? f * 9 + 1 + 3 * j : tm) * HW;
  double sb = 0; int nb = 0;
  const float4* __restrict__ y4 = reinterpret_cast<const float4*>(y);      // HW is a multiple of 4 (w % 8 == 0); two 16-byte loads in flight per thread
  const long n4v = HW >> 2, stride = (long)gridDim.x * 256;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n4v; i += 2 * stride) {
    const float4 a = y4[i];
    const bool two = i + stride < n4v;
    const float4 b = two ? y4[i + stride] : make_float4(mn, mn, mn, mn);
    const float va[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float v = valid ? (va[k] - mn) * inv : 0.f;
      if (v > lim && (k < 4 || two)) { sb += v; nb++; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sb += __shfl_xor_sync(0xffffffffu, sb, o); nb += __shfl_xor_sync(0xffffffffu, nb, o); }
  if ((threadIdx.x & 31) == 0) { redd[threadIdx.x >> 5] = sb; redi[threadIdx.x >> 5] = nb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0; int n = 0;
    for (int i = 0; i < 8; i++) { s += redd[i]; n += redi[i]; }
    float* p = partial2 + ((size_t)tm * gridDim.x + blockIdx.x) * 4;
    reinterpret_cast<double*>(p)[0] = s; reinterpret_cast<int*>(p)[2] = n;
  }
}
__global__ void __launch_bounds__(32)
temporal_thresh_finalize_kernel(const float* __restrict__ partial2, int nb2, const float* __restrict__ stats9, float* __restrict__ xf) {
  const int tm = blockIdx.x, lane = threadIdx.x;
  const int f = tm / 3, j = tm - 3 * f;
  double s = 0; int n = 0;
  for (int i = lane; i < nb2; i += 32) {
    const float* p = partial2 + ((size_t)tm * nb2 + i) * 4;
    s += reinterpret_cast<const double*>(p)[0]; n += reinterpret_cast<const int*>(p)[2];
  }
  s = warp_sum(s); n = warp_sum(n);
  if (lane == 0) {
    const float* st = stats9 + (size_t)(f * 9 + 1 + 3 * j) * 4;
    const float mn = st[0], mx = st[1];
    const bool valid = mx > mn;
    const float mxn = valid ? 1.f : 0.f;
    const float thresh = n > 0 ? (float)(s / n) : INFINITY;                // mysterythrehold.m:5 (empty set -> no clamp)
    const float M = n > 0 ? fminf(thresh, mxn) : mxn;
    float* o = xf + (size_t)tm * 4;
    o[0] = mn; o[1] = valid ? 1.f / (mx - mn) : 0.f; o[2] = thresh; o[3] = 1.f / (M + 0.000001f);   // :6-7
  }
}

// featuresTest (main.m:284-294): 9 features at 200x200, z-scored, row-major pixel order n = r*200 + c
__global__ void __launch_bounds__(256)
assemble_features_kernel(const float* __restrict__ F200, const float* __restrict__ Ft, const float* __restrict__ stats9,
                         const float* __restrict__ xf, const float* __restrict__ fparam /*[9][4] = mean, 1/std, weight, 0*/, float* __restrict__ xsvm) {
  // a thread computes the nine features of one pixel; the CTA's 256 x 9 floats are one contiguous piece of xsvm, so they go out
  // through shared memory as coalesced stores (nine scalar stores per thread 36 bytes apart wrote 4 bytes per sector per instruction)
  __shared__ float so[256 * 9];
  const int f = blockIdx.y, n0 = blockIdx.x * 256, n = n0 + threadIdx.x;
  if (n < 40000) {
#pragma unroll
    for (int i = 0; i < 9; i++) {
      float v;
      if (i % 3 == 1) {
        const int tm = f * 3 + i / 3;
        v = Ft[(size_t)tm * 40000 + n] * xf[(size_t)tm * 4 + 3];
      } else {
        const float* st = stats9 + (size_t)(f * 9 + i) * 4;
        const float mn = st[0], mx = st[1];
        v = mx > mn ? (F200[(size_t)(f * 9 + i) * 40000 + n] - mn) / (mx - mn) : 0.f;     // pm_norm, NaN -> 0
      }
      so[threadIdx.x * 9 + i] = (v - fparam[4 * i]) * fparam[4 * i + 1];
    }
  }
  __syncthreads();
  const int cnt = min(256, 40000 - n0) * 9;
  float* __restrict__ o = xsvm + ((size_t)f * 40000 + n0) * 9;
  for (int i = threadIdx.x; i < cnt; i += 256) o[i] = so[i];
}

__global__ void __launch_bounds__(1024)
dec_minmax_norm_kernel(const float* __restrict__ dec, float* __restrict__ C) {
  __shared__ float red[32];
  const int f = blockIdx.x;
  const float* d = dec + (size_t)f * 40000;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < 40000; i += 1024) { const float v = d[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
  __syncthreads();
  float m0 = red[0]; for (int i = 1; i < 32; i++) m0 = fminf(m0, red[i]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  float m1 = red[0]; for (int i = 1; i < 32; i++) m1 = fmaxf(m1, red[i]);
  const float inv = m1 > m0 ? 1.f / (m1 - m0) : 0.f;                        // pm_norm (main.m:299); flat -> 0 (reference: NaN)
  for (int i = threadIdx.x; i < 40000; i += 1024) C[(size_t)f * 40000 + i] = (d[i] - m0) * inv;
}

// final map: Y = Bv * T2 (imfilter o imresize), times the centre prior, stored raw + min/max partials (main.m:300-304).
// Register-tiled banded FIR: a thread owns 4 columns x 4 consecutive output rows; the CTA stages the four rows'
// weights over their common input window as float4 (one broadcast LDS.128 per input row feeds 16 FFMAs).
constexpr int FR_ROWS = 4;
__global__ void __launch_bounds__(128)
final_rows_kernel(const float* __restrict__ T2, int W, int H, const float* __restrict__ w, const int* __restrict__ start, int nt,
                  const float* __restrict__ dist, float* __restrict__ raw, float* __restrict__ partial) {
  extern __shared__ float4 swf4[];                // [window] x 4 rows
  __shared__ float redf[4];
  const int y0 = blockIdx.y * FR_ROWS, f = blockIdx.z, W4 = W >> 2;
  const int ylast = min(y0 + FR_ROWS - 1, H - 1);
  const int s_lo = start[y0], L = min(start[ylast] + nt, 200) - s_lo;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float g[FR_ROWS];
#pragma unroll
    for (int k = 0; k < FR_ROWS; k++) {
      const int y = y0 + k;
      const int t = y < H ? j + s_lo - start[min(y, H - 1)] : -1;
      g[k] = (t >= 0 && t < nt) ? w[(size_t)y * nt + t] : 0.f;
    }
    swf4[j] = make_float4(g[0], g[1], g[2], g[3]);
  }
  __syncthreads();
  const float4* __restrict__ src = reinterpret_cast<const float4*>(T2 + (size_t)f * 200 * W);
  float mn = INFINITY, mx = -INFINITY;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < W4) {
    float4 acc[FR_ROWS];
#pragma unroll
    for (int k = 0; k < FR_ROWS; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int j = 0; j < L; j++) {
      const float4 v = __ldg(src + (size_t)(s_lo + j) * W4 + x);
      const float4 g = swf4[j];
      acc[0].x = fmaf(g.x, v.x, acc[0].x); acc[0].y = fmaf(g.x, v.y, acc[0].y); acc[0].z = fmaf(g.x, v.z, acc[0].z); acc[0].w = fmaf(g.x, v.w, acc[0].w);
      acc[1].x = fmaf(g.y, v.x, acc[1].x); acc[1].y = fmaf(g.y, v.y, acc[1].y); acc[1].z = fmaf(g.y, v.z, acc[1].z); acc[1].w = fmaf(g.y, v.w, acc[1].w);
      acc[2].x = fmaf(g.z, v.x, acc[2].x); acc[2].y = fmaf(g.z, v.y, acc[2].y); acc[2].z = fmaf(g.z, v.z, acc[2].z); acc[2].w = fmaf(g.z, v.w, acc[2].w);
      acc[3].x = fmaf(g.w, v.x, acc[3].x); acc[3].y = fmaf(g.w, v.y, acc[3].y); acc[3].z = fmaf(g.w, v.z, acc[3].z); acc[3].w = fmaf(g.w, v.w, acc[3].w);
    }
#pragma unroll
    for (int k = 0; k < FR_ROWS; k++) {
      const int y = y0 + k;
      if (y < H) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(dist) + (size_t)y * W4 + x);
        float4 a = acc[k];
        a.x *= d.x; a.y *= d.y; a.z *= d.z; a.w *= d.w;
        reinterpret_cast<float4*>(raw)[((size_t)f * H + y) * W4 + x] = a;
        mn = fminf(mn, fminf(fminf(a.x, a.y), fminf(a.z, a.w)));
        mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
      }
    }
  }
  mn = warp_min(mn); mx = warp_max(mx);
  const int nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mn;
  __syncthreads();
  float m0 = redf[0]; for (int i = 1; i < nw; i++) m0 = fminf(m0, redf[i]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mx;
  __syncthreads();
  float m1 = redf[0]; for (int i = 1; i < nw; i++) m1 = fmaxf(m1, redf[i]);
  if (threadIdx.x == 0) {
    const size_t nb = (size_t)gridDim.x * gridDim.y;
    float* p = partial + ((size_t)f * nb + (size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
    p[0] = m0; p[1] = m1; reinterpret_cast<double*>(p + 2)[0] = 0.0;
  }
}

// linear fusion (main.m:258-272) from the nine stored full-res maps, times the centre prior
__global__ void __launch_bounds__(256)
linear_comb_kernel(const float* __restrict__ Y, long HW, const float* __restrict__ stats9, const float* __restrict__ xf,
                   const float* __restrict__ fparam, const float* __restrict__ dist, float* __restrict__ raw,
                   float* __restrict__ partial) {
  __shared__ float redf[8];
  const int f = blockIdx.y;
  float a[9], b[9], thr[9], sc[9];
  for (int i = 0; i < 9; i++) {
    const float* st = stats9 + (size_t)(f * 9 + i) * 4;
    const bool valid = st[1] > st[0];
    a[i] = st[0]; b[i] = valid ? 1.f / (st[1] - st[0]) : 0.f; thr[i] = INFINITY; sc[i] = 1.f;
    if (i % 3 == 1) { const float* x = xf + (size_t)(f * 3 + i / 3) * 4; thr[i] = x[2]; sc[i] = x[3]; }
  }
  float mn = INFINITY, mx = -INFINITY;
  for (long p = blockIdx.x * 256L + threadIdx.x; p < HW; p += (long)gridDim.x * 256) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const float v = fminf((Y[((size_t)f * 9 + i) * HW + p] - a[i]) * b[i], thr[i]) * sc[i];
      s += (v - fparam[4 * i]) * fparam[4 * i + 1] * fparam[4 * i + 2];
    }
    s *= dist[p];
    raw[(size_t)f * HW + p] = s;
    mn = fminf(mn, s); mx = fmaxf(mx, s);
  }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mn;
  __syncthreads();
  float m0 = redf[0]; for (int i = 1; i < 8; i++) m0 = fminf(m0, redf[i]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mx;
  __syncthreads();
  float m1 = redf[0]; for (int i = 1; i < 8; i++) m1 = fmaxf(m1, redf[i]);
  if (threadIdx.x == 0) { float* p = partial + ((size_t)f * gridDim.x + blockIdx.x) * 4; p[0] = m0; p[1] = m1; reinterpret_cast<double*>(p + 2)[0] = 0.0; }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
normalize_out_kernel(const float* __restrict__ raw, long HW, const float* __restrict__ fstats, OutT* __restrict__ out) {
  const int f = blockIdx.y;
  const float mn = fstats[(size_t)f * 4], mx = fstats[(size_t)f * 4 + 1];
  const float den = mx - mn;                                                // pm_norm (main.m:304): a true division, so the maximum maps to exactly 1
  for (long p = blockIdx.x * 256L + threadIdx.x; p < HW; p += (long)gridDim.x * 256) {
    const float v = den > 0.f ? (raw[(size_t)f * HW + p] - mn) / den : 0.f;
    if (sizeof(OutT) == 1) out[(size_t)f * HW + p] = (OutT)__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 255.f);
    else out[(size_t)f * HW + p] = (OutT)v;
  }
}

// the same, four pixels per thread (HW is a multiple of 4; used when `out` is aligned for 4-wide stores)
template <typename OutT>
__global__ void __launch_bounds__(256)
normalize_out4_kernel(const float* __restrict__ raw, long HW, const float* __restrict__ fstats, OutT* __restrict__ out) {
  const int f = blockIdx.y;
  const float mn = fstats[(size_t)f * 4], mx = fstats[(size_t)f * 4 + 1];
  const float den = mx - mn;
  const float4* __restrict__ r4 = reinterpret_cast<const float4*>(raw + (size_t)f * HW);
  for (long p = blockIdx.x * 256L + threadIdx.x; p < (HW >> 2); p += (long)gridDim.x * 256) {
    const float4 q = r4[p];
    float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = den > 0.f ? (v[k] - mn) / den : 0.f;
    if (sizeof(OutT) == 1) {
      uchar4 o;
      o.x = (unsigned char)__float2int_rn(fminf(fmaxf(v[0], 0.f), 1.f) * 255.f); o.y = (unsigned char)__float2int_rn(fminf(fmaxf(v[1], 0.f), 1.f) * 255.f);
      o.z = (unsigned char)__float2int_rn(fminf(fmaxf(v[2], 0.f), 1.f) * 255.f); o.w = (unsigned char)__float2int_rn(fminf(fmaxf(v[3], 0.f), 1.f) * 255.f);
      reinterpret_cast<uchar4*>(out + (size_t)f * HW)[p] = o;
    } else {
      reinterpret_cast<float4*>(out + (size_t)f * HW)[p] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

}  // namespace

// =================================================================================================
// context
// =================================================================================================
namespace {

int build_operators(cds_ctx* c) {
  const int H = c->h, W = c->w, h4 = c->h4, w4 = c->w4;
  build_gauss_polyphase(c->gk, c->gs, &c->poly);
  if (int e = c->poly.upload()) return e;
  // composite 200 x cells operators:  imresize o imfilter o fourX2
  {
    std::vector<double> Kv = dense_gauss_rep4(h4, c->gk, c->gs), Rv = dense_imresize(H, 200);
    banded_from_dense(dense_mul(Rv, 200, H, Kv, h4), 200, h4, &c->Av);
    banded_from_dense(Rv, 200, H, &c->Rv);
    std::vector<double> Kh = dense_gauss_rep4(w4, c->gk, c->gs), Rh = dense_imresize(W, 200);
    banded_from_dense(dense_mul(Rh, 200, W, Kh, w4), 200, w4, &c->Ah);
    banded_from_dense(Rh, 200, W, &c->Rh);
  }
  // final: imfilter o imresize(200 -> H / W)
  {
    std::vector<double> Gv = dense_gauss(H, c->gk, c->gs), Uv = dense_imresize(200, H);
    banded_from_dense(dense_mul(Gv, H, H, Uv, 200), H, 200, &c->Bv);
    std::vector<double> Gh = dense_gauss(W, c->gk, c->gs), Uh = dense_imresize(200, W);
    banded_from_dense(dense_mul(Gh, W, W, Uh, 200), W, 200, &c->Bh);
  }
  for (Banded* b : {&c->Av, &c->Ah, &c->Rv, &c->Rh, &c->Bv, &c->Bh}) if (int e = b->upload()) return e;
  if (c->Bv.span4 < 0) { set_error("cds_create: final-map operator is not banded monotonically"); return CDS_EINVAL; }
  // centre prior, getdistMatrix.m:1-17
  {
    std::vector<float> dist((size_t)H * W);
    const int mx_ = H / 2, my_ = W / 2; double dmax = 0;
    std::vector<double> d((size_t)H * W);
    for (int x = 1; x <= H; x++) for (int y = 1; y <= W; y++) {
      const double v = floor(sqrt((double)(x - mx_) * (x - mx_) + (double)(y - my_) * (y - my_)));
      d[(size_t)(x - 1) * W + (y - 1)] = v; if (v > dmax) dmax = v;
    }
    for (size_t i = 0; i < d.size(); i++) dist[i] = (float)(1.0 - d[i] / dmax);
    if (int e = dalloc(c, &c->d_dist, dist.size())) return e;
    CDS_CUDA(cudaMemcpy(c->d_dist, dist.data(), dist.size() * 4, cudaMemcpyHostToDevice));
  }
  // 5x5 bit-contrast weights exp(-R^2/(2*3^2)), Sudoku_contrastcpp5.m:8-14 with delta_bit_spacial = 3
  {
    float g[8], W5[25];
    gaussian_factor(5, 3.0, g);
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) W5[i * 5 + j] = g[i] * g[j];
    CDS_CUDA(cudaMemcpy(c->d_w5, W5, sizeof(W5), cudaMemcpyHostToDevice));
  }
  // temporal weights exp(-j/delta), main.m:120-122,195-197
  {
    std::vector<float> wt((size_t)(c->PT + 1) * 4, 0.f);
    for (int j = 0; j <= c->PT; j++) { wt[4 * j] = (float)exp(-j / 26.0); wt[4 * j + 1] = (float)exp(-j / 46.0); wt[4 * j + 2] = (float)exp(-j / 46.0); }
    if (int e = dalloc(c, &c->d_wtab, wt.size())) return e;
    CDS_CUDA(cudaMemcpy(c->d_wtab, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
  }
  return CDS_OK;
}

}  // namespace

namespace { int join_back(cds_ctx* c, cudaStream_t st); void hook_reset(cds_ctx* c); void hook_free(cds_ctx* c); }

extern "C" int cds_create(const cds_params* p, cds_ctx** out) {
  if (int e = ensure_device()) return e;
  if (!p || !out) { set_error("cds_create: null argument"); return CDS_EINVAL; }
  if (p->w < 64 || p->h < 64 || (p->w % 8) || (p->h % 8) || p->w > 8192 || p->h > 8192) {
    set_error("cds_create: %dx%d unsupported (multiples of 8, 64..8192)", p->w, p->h); return CDS_EINVAL;
  }
  if (!(p->fps >= 4) || p->fps > 240) { set_error("cds_create: fps %.3f out of range", p->fps); return CDS_EINVAL; }
  if (p->use_rbf < 0 || p->use_rbf > CDS_FUSION_FAST) { set_error("cds_create: use_rbf %d unknown", p->use_rbf); return CDS_EINVAL; }
  const bool fm = p->use_rbf == CDS_FUSION_FAST;
  if (p->use_rbf == CDS_FUSION_RBF && (!p->model || cds_svm_model_nfeat(p->model) != 9)) {
    set_error("cds_create: use_rbf needs a 9-feature RBF model (svmRBFmodel_all.mat stand-in)"); return CDS_EINVAL;
  }
  for (int i = 0; i < 9 && !fm; i++) if (!(p->std_vec[i] != 0)) { set_error("cds_create: std_vec[%d] is zero", i); return CDS_EINVAL; }
  cds_ctx* c = new cds_ctx();
  c->p = *p;
  c->w = p->w; c->h = p->h; c->h4 = p->h / 4; c->w4 = p->w / 4; c->n4 = c->h4 * c->w4;
  c->ncol = (p->w + 63) / 64; c->nrow = (p->h + 63) / 64; c->nctu = c->ncol * c->nrow;
  c->HW = (long)p->w * p->h;
  c->B = p->max_batch > 0 ? p->max_batch : 32;
  c->PS = (int)mround(p->fps * 0.3);                       // main.m:119
  c->PT = (int)mround(c->PS * 7.0);                        // main.m:129
  c->gk = (int)mround(p->h / 7.5); c->gs = mround(p->h / 30.0);   // main.m:211
  if (c->gk < 1 || c->gs <= 0) { delete c; set_error("cds_create: picture too small for the Gaussian"); return CDS_EINVAL; }
  c->RL = c->B + c->PT + c->PS;
  const int B = c->B, n4 = c->n4, H = c->h, W = c->w, h4 = c->h4, w4 = c->w4;
  const size_t HW = (size_t)c->HW;
  int e = CDS_OK;
  cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete c; set_error("cudaStreamCreate: %s", cudaGetErrorString(ce)); return CDS_ECUDA; }
#define A(ptr, count) if (!e) e = dalloc(c, &c->ptr, (size_t)(count))
  A(d_mvx, (size_t)B * n4); A(d_mvy, (size_t)B * n4); A(d_dep, (size_t)B * n4); A(d_err, 4);
  if (!e && cudaMemset(c->d_err, 0, 16) != cudaSuccess) { set_error("cudaMemset d_err"); e = CDS_ECUDA; }
  if (fm) {
    if (!e) e = fast_create(p->w, p->h, p->fps, B, &c->fast);
    if (!e) { c->PS = fast_halo(c->fast) + 2; c->PT = 0; }
    if (!e && (cudaStreamCreateWithFlags(&c->fast_front_stream, cudaStreamNonBlocking) != cudaSuccess ||
               cudaEventCreateWithFlags(&c->ev_fstart, cudaEventDisableTiming) != cudaSuccess)) { set_error("cudaStreamCreate fast front"); e = CDS_ECUDA; }
    for (int b = 0; b < 2 && !e; b++)
      if (cudaEventCreateWithFlags(&c->ev_ff[b], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_fb[b], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate"); e = CDS_ECUDA; }
    { const char* ov = getenv("CDS_OVERLAP"); c->fast_ovl = !(ov && ov[0] == '0'); }
    if (!e) fast_set_marker(c->fast, [](void* user, int stage, cudaStream_t st) { prof_mark(static_cast<cds_ctx*>(user), ST_FAST0 + stage, st); }, c);
  }
#undef A
#define A(ptr, count) if (!e && !fm) e = dalloc(c, &c->ptr, (size_t)(count))
  A(r_Vx, (size_t)c->RL * n4); A(r_Vy, (size_t)c->RL * n4); A(r_mvn, (size_t)c->RL * n4); A(r_bitn, (size_t)c->RL * n4);
  A(r_depn, (size_t)c->RL * n4); A(r_sm4, (size_t)c->RL * n4); A(r_cam, (size_t)c->RL * 2);
  A(d_mv_s, (size_t)B * n4); A(d_X, (size_t)B * 9 * n4); A(d_cxy, (size_t)2 * B * n4); A(d_keys64, (size_t)4 * B);
  const int sbR = (h4 + 15) / 16, sbC = (w4 + 15) / 16, sdR = (h4 + 1) / 2, sdC = (w4 + 1) / 2;
  A(d_subB, (size_t)B * sbR * sbC); A(d_padB, (size_t)B * (sbR + 4) * (sbC + 4)); A(d_rawB, (size_t)B * sbR * sbC);
  A(d_subD, (size_t)B * sdR * sdC); A(d_padD, (size_t)B * (sdR + 24) * (sdC + 24)); A(d_rawD, (size_t)B * sdR * sdC);
  A(d_padM, (size_t)2 * B * (h4 + 48) * (w4 + 48)); A(d_rawM, (size_t)2 * B * n4);
  c->d_X2[0] = c->d_X; c->d_padB2[0] = c->d_padB; c->d_padD2[0] = c->d_padD; c->d_padM2[0] = c->d_padM;
  c->d_X2[1] = c->d_X; c->d_padB2[1] = c->d_padB; c->d_padD2[1] = c->d_padD; c->d_padM2[1] = c->d_padM;
  { const char* ov = getenv("CDS_OVERLAP"); const char* po = getenv("CDS_PRE_OVERLAP");
    c->pre_ovl = !(ov && ov[0] == '0') && !(po && po[0] == '0') && p->use_rbf == CDS_FUSION_RBF; }
  if (c->pre_ovl && !fm) {
    A(d_X2[1], (size_t)B * 9 * n4); A(d_padB2[1], (size_t)B * (sbR + 4) * (sbC + 4)); A(d_padD2[1], (size_t)B * (sdR + 24) * (sdC + 24));
    A(d_padM2[1], (size_t)2 * B * (h4 + 48) * (w4 + 48));
  }
  A(d_keys, (size_t)4 * B * 2); A(d_maxbits, (size_t)4 * B); A(d_w5, 64 * 64); A(d_sched, 4096);
  A(d_T, (size_t)9 * B * h4 * W);
  A(d_Y, (size_t)(p->use_rbf ? 3 : 9) * B * HW);
  A(d_store_index, (size_t)9 * B);
  c->nblk = fullres_vpass_nblocks(h4, W);
  c->nblk2 = 64;
  c->nblkf = p->use_rbf ? ceil_div(W / 4, fir_threads(W / 4)) * ceil_div(H, FR_ROWS) : 128;
  A(d_partial, (size_t)9 * B * c->nblk * 4); A(d_stats9, (size_t)9 * B * 4);
  A(d_partial2, (size_t)3 * B * c->nblk2 * 4); A(d_xf, (size_t)3 * B * 4);
  A(d_A1, (size_t)9 * B * 200 * w4); A(d_F200, (size_t)9 * B * 40000);
  A(d_R1, (size_t)3 * B * 200 * W); A(d_Ft, (size_t)3 * B * 40000);
  A(d_xsvm, (size_t)B * 40000 * 9); A(d_dec, (size_t)B * 40000); A(d_C, (size_t)B * 40000);
  if (p->use_rbf) { A(d_xsvm2[1], (size_t)B * 40000 * 9); }
  c->d_xsvm2[0] = c->d_xsvm;
  A(d_T2, (size_t)B * 200 * W); A(d_raw, (size_t)B * HW);
  A(d_fpartial, (size_t)B * c->nblkf * 4); A(d_fstats, (size_t)B * 4);
#undef A
#define A(ptr, count) if (!e) e = dalloc(c, &c->ptr, (size_t)(count))
  A(d_cam_stage, (size_t)2 * B);
  A(d_bytes, (size_t)B * c->nctu); A(d_hdr, (size_t)B * c->nctu * 4); A(d_tokbase, (size_t)B + 1);
  c->tok_cap = (size_t)B * c->nctu * 160;
  A(d_tok, c->tok_cap);
  for (int b = 0; b < 2 && !e; b++) {
    void* q = nullptr; ce = cudaMalloc(&q, (size_t)B * HW * (p->out_u8 ? 1 : 4) + 256);
    if (ce != cudaSuccess) { set_error("cudaMalloc out"); e = CDS_ENOMEM; } else { c->allocs.push_back(q); c->d_out[b] = q; }
    if (!e && (cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming) != cudaSuccess ||
               cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming) != cudaSuccess)) { set_error("cudaEventCreate"); e = CDS_ECUDA; }
  }
  if (!e && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate copy"); e = CDS_ECUDA; }
  if (!e && (cudaStreamCreateWithFlags(&c->pre_stream, cudaStreamNonBlocking) != cudaSuccess ||
             cudaEventCreateWithFlags(&c->ev_pre[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_pre[1], cudaEventDisableTiming) != cudaSuccess ||
             cudaEventCreateWithFlags(&c->ev_c48, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_call, cudaEventDisableTiming) != cudaSuccess)) {
    set_error("cudaStreamCreate pre"); e = CDS_ECUDA;
  }
  if (!e) {           // the back stream outranks the front stream: the persistent SVM CTAs are placed before the contrast CTAs that start with them
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&c->back_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_mid, cudaEventDisableTiming) != cudaSuccess) { set_error("cudaStreamCreate back"); e = CDS_ECUDA; }
  }
  for (int b = 0; b < 2 && !e; b++)
    if (cudaEventCreateWithFlags(&c->ev_front[b], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_back[b], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate"); e = CDS_ECUDA; }
  { const char* ov = getenv("CDS_OVERLAP"); c->overlap = !(ov && ov[0] == '0') && p->use_rbf == CDS_FUSION_RBF; }
  { const char* tr = getenv("CDS_TRACE"); c->trace = tr && tr[0] == '1'; }
#undef A
  float* d_fparam = nullptr;
  if (!e && !fm) e = dalloc(c, &d_fparam, 36);
  if (!e && !fm) {
    float fp[36];
    for (int i = 0; i < 9; i++) { fp[4 * i] = (float)p->mean_vec[i]; fp[4 * i + 1] = (float)(1.0 / p->std_vec[i]); fp[4 * i + 2] = kFeatureWeight[i]; fp[4 * i + 3] = 0.f; }
    if (cudaMemcpy(d_fparam, fp, sizeof(fp), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("cudaMemcpy fparam"); e = CDS_ECUDA; }
    c->d_fparam = d_fparam;
  }
  if (!e && !fm) e = build_operators(c);
  if (!e && !fm) {
    std::vector<int> si((size_t)9 * B);
    for (int f = 0; f < B; f++) for (int i = 0; i < 9; i++)
      si[(size_t)f * 9 + i] = p->use_rbf ? ((i % 3 == 1) ? f * 3 + i / 3 : -1) : f * 9 + i;
    if (cudaMemcpy(c->d_store_index, si.data(), si.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("cudaMemcpy store_index"); e = CDS_ECUDA; }
  }
  if (e) { cds_destroy(c); return e; }
  cds_seek(c, 0);
  *out = c;
  return CDS_OK;
}

extern "C" void cds_destroy(cds_ctx* c) {
  if (!c) return;
  trace_dump(c);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->back_stream) { cudaStreamSynchronize(c->back_stream); cudaStreamDestroy(c->back_stream); }
  if (c->pre_stream) { cudaStreamSynchronize(c->pre_stream); cudaStreamDestroy(c->pre_stream); }
  for (cudaEvent_t ev : {c->ev_pre[0], c->ev_pre[1], c->ev_c48, c->ev_call}) if (ev) cudaEventDestroy(ev);
  if (c->fast_front_stream) { cudaStreamSynchronize(c->fast_front_stream); cudaStreamDestroy(c->fast_front_stream); }
  if (c->ev_fstart) cudaEventDestroy(c->ev_fstart);
  if (c->ev_mid) cudaEventDestroy(c->ev_mid);
  for (int b = 0; b < 2; b++) { if (c->ev_ff[b]) cudaEventDestroy(c->ev_ff[b]); if (c->ev_fb[b]) cudaEventDestroy(c->ev_fb[b]); }
  for (int b = 0; b < 2; b++) { if (c->ev_front[b]) cudaEventDestroy(c->ev_front[b]); if (c->ev_back[b]) cudaEventDestroy(c->ev_back[b]); }
  for (int b = 0; b < 2; b++) { if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]); if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]); }
  if (c->h_cam_pin) cudaFreeHost(c->h_cam_pin);
  hook_free(c);
  fast_destroy(c->fast);
  for (void* q : c->allocs) cudaFree(q);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  c->poly.release();
  for (Banded* b : {&c->Av, &c->Ah, &c->Rv, &c->Rh, &c->Bv, &c->Bh}) b->release();
  delete c;
}

extern "C" int cds_seek(cds_ctx* c, int poc) {
  if (!c || poc < 0) { set_error("cds_seek: bad arguments"); return CDS_EINVAL; }
  join_back(c, c->stream);
  cudaStreamSynchronize(c->stream);
  if (c->pre_stream) cudaStreamSynchronize(c->pre_stream);
  if (c->back_stream) cudaStreamSynchronize(c->back_stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->fast) {
    if (int e = fast_clear(c->fast, c->stream)) return e;
    cudaStreamSynchronize(c->stream);
    c->next_poc = poc; c->pending = 0; c->pending_first = poc; c->last_first = -1; c->last_n = 0;
    hook_reset(c);
    return CDS_OK;
  }
  const size_t rb = (size_t)c->RL * c->n4 * 4;
  for (float* r : {c->r_mvn, c->r_bitn, c->r_depn})
    CDS_CUDA(cudaMemset(r, 0, rb));
  CDS_CUDA(cudaMemset(c->r_sm4, 0, 4 * rb));
  for (double* r : {c->r_Vx, c->r_Vy}) CDS_CUDA(cudaMemset(r, 0, 2 * rb));
  CDS_CUDA(cudaMemset(c->r_cam, 0, (size_t)c->RL * 8));
  CDS_CUDA(cudaStreamSynchronize(0));                  // the context's streams are non-blocking: they do not wait for the null stream
  c->next_poc = poc; c->pending = 0; c->pending_first = poc; c->last_first = -1; c->last_n = 0;
  hook_reset(c);
  return CDS_OK;
}
extern "C" int cds_reset(cds_ctx* c) { return cds_seek(c, 0); }

namespace {

// The stream on which this API call's ring-touching work (host->device staging, stages (a), (c), smoothing, ...) runs: the context's pre
// stream when the pre stage overlaps the previous batch's map passes, else `st`.  The first use inside an API call orders the pre
// stream after everything the caller had enqueued on `st` before the call; pre_join() closes the call.
cudaStream_t pre_stream_for(cds_ctx* c, cudaStream_t st) {
  const bool pov = c->overlap && c->pre_ovl && c->p.use_rbf == CDS_FUSION_RBF && !c->prof_on && !c->maps_in && !c->fast;
  if (!pov) return st;
  if (!c->call_open) {
    if (cudaEventRecord(c->ev_call, st) != cudaSuccess || cudaStreamWaitEvent(c->pre_stream, c->ev_call, 0) != cudaSuccess) return st;
    c->call_open = true;
  }
  return c->pre_stream;
}
// end of an API call: `st` covers whatever is still running on the pre stream
int pre_join(cds_ctx* c, cudaStream_t st) {
  if (!c->call_open) return CDS_OK;
  c->call_open = false;
  CDS_CUDA(cudaEventRecord(c->ev_call, c->pre_stream));
  CDS_CUDA(cudaStreamWaitEvent(st, c->ev_call, 0));
  return CDS_OK;
}

// stages (a), (c) and smoothing: everything later pictures need from this one as temporal context
int run_context_stages(cds_ctx* c, int first_poc, int n, const uint16_t* bytes, const int32_t* hdr, const int16_t* tok,
                       const int64_t* tok_base, cudaStream_t st) {
  prof_mark(c, ST_GETFRAME, st);
  if (!c->maps_in) if (int e = getframe_launch(c->w, c->h, n, hdr, tok, tok_base, c->d_mvx, c->d_mvy, c->d_dep, c->d_err, st)) return e;
  if (c->fast) return fast_run_batch(c->fast, first_poc, n, c->d_mvx, c->d_mvy, c->d_dep, bytes, nullptr, 0, nullptr, st);
  prof_mark(c, ST_VOTE, st);
  if (int e = camera_motion_launch_f64(c->w, c->h, n, c->d_mvx, c->d_mvy, bytes, c->d_dep, c->r_Vx, c->r_Vy, c->r_mvn, c->r_bitn,
                                       c->r_depn, c->r_cam, nullptr, nullptr, first_poc % c->RL, c->RL, st)) return e;
  prof_mark(c, ST_SMOOTH, st);
  return smooth_launch(c->n4, c->PS, c->RL, first_poc, n, c->r_Vx, c->r_Vy, c->r_mvn, c->r_bitn, c->r_depn, c->r_sm4,
                       c->d_mv_s, c->d_cxy, st);
}

template <typename OutT>
int launch_normalize(cds_ctx* c, int n, void* out, cudaStream_t st) {
  dim3 grid(256, n);
  if ((c->HW & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & (4 * sizeof(OutT) - 1)) == 0)
    normalize_out4_kernel<OutT><<<grid, 256, 0, st>>>(c->d_raw, c->HW, c->d_fstats, reinterpret_cast<OutT*>(out));
  else
    normalize_out_kernel<OutT><<<grid, 256, 0, st>>>(c->d_raw, c->HW, c->d_fstats, reinterpret_cast<OutT*>(out));
  CDS_LAUNCH_CHECK();
  return CDS_OK;
}

// (d) RBF-SVM, then pm_norm -> imresize -> imfilter -> centre prior -> pm_norm (main.m:297-304) for the batch whose z-scored
// features sit in d_xsvm2[slot].  On `sb`; prof marks only when sb is the profiled stream.
int launch_back_rbf(cds_ctx* c, int slot, int n, void* maps_out, cudaStream_t sb, bool mark) {
  const int H = c->h, W = c->w;
  if (mark) prof_mark(c, ST_SVM, sb);
  trace_mark(c, c->tr_batches, TR_SVM_START, sb);
  if (int e = svm_predict_launch(c->p.model, c->d_xsvm2[slot], (long)n * 40000, c->d_dec, sb)) return e;
  trace_mark(c, c->tr_batches, TR_SVM_END, sb);
  if (mark) prof_mark(c, ST_FINAL, sb);
  dec_minmax_norm_kernel<<<n, 1024, 0, sb>>>(c->d_dec, c->d_C);
  CDS_LAUNCH_CHECK();
  if (int e = banded_cols_launch(c->d_C, n, 200, 200, c->Bh, c->d_T2, sb)) return e;
  const int thr = fir_threads(W / 4);
  dim3 grid(ceil_div(W / 4, thr), ceil_div(H, FR_ROWS), n);
  final_rows_kernel<<<grid, thr, (size_t)(c->Bv.nt + c->Bv.span4) * 16, sb>>>(c->d_T2, W, H, c->Bv.d_w, c->Bv.d_start, c->Bv.nt, c->d_dist, c->d_raw, c->d_fpartial);
  CDS_LAUNCH_CHECK();
  if (int e = stats_finalize_launch(c->d_fpartial, n, grid.x * grid.y, c->d_fstats, sb)) return e;
  if (maps_out) {
    if (int e = c->p.out_u8 ? launch_normalize<uint8_t>(c, n, maps_out, sb) : launch_normalize<float>(c, n, maps_out, sb)) return e;
  }
  return CDS_OK;
}

// enqueue the deferred back half on the back stream; `with` (optional) is an event of the front stream it should start with
int flush_pending_back(cds_ctx* c, cudaEvent_t with) {
  if (!c->pb.valid) return CDS_OK;
  const cds_ctx::PendingBack pb = c->pb;
  c->pb.valid = false;
  cudaStream_t sb = c->back_stream;
  CDS_CUDA(cudaStreamWaitEvent(sb, c->ev_front[pb.slot], 0));
  if (with) CDS_CUDA(cudaStreamWaitEvent(sb, with, 0));
  if (pb.before_back) CDS_CUDA(cudaStreamWaitEvent(sb, pb.before_back, 0));
  if (int e = launch_back_rbf(c, pb.slot, pb.n, pb.maps_out, sb, false)) return e;
  trace_mark(c, c->tr_batches, TR_BACK_END, sb);
  CDS_CUDA(cudaEventRecord(c->ev_back[pb.slot], sb));
  c->retired_tag = pb.tag; c->retired_ev = c->ev_back[pb.slot];
  return CDS_OK;
}

// The batch has a front half (stages a, c, e, b and the 200x200 features; FP32-pipe bound) and a back half (SVM -> final map;
// MUFU / tensor bound).  With `overlap` the back half is deferred (cds_ctx::pb) and runs on the context's high-priority back
// stream beside the contrast kernels of the NEXT batch; xsvm is double buffered and events order the halves.  `before_back`
// (optional) must have happened before the back half writes maps_out.  When a back half (this batch's or the previous one's)
// has been enqueued, c->retired_tag / c->retired_ev name it (retired_ev == nullptr: it ended on `st`).
int run_batch(cds_ctx* c, int first_poc, int n, const uint16_t* bytes, const int32_t* hdr, const int16_t* tok,
              const int64_t* tok_base, void* maps_out, int* cam_out, cudaStream_t st, cudaEvent_t before_back = nullptr,
              int tag = 0) {
  if (c->fast) {          // TDecGop.cpp:270-332 for n pictures: stage (a), then the fast version's kernels (fast.cu)
    c->retired_tag = tag; c->retired_ev = nullptr;
    if (c->fast_ovl_active && !c->prof_on && !c->maps_in) {
      // front half on its own stream: it waited for the start of this API call (inputs resident) and waits for the back half of
      // batch k-2, whose ring reads end where this batch's ring writes begin; the back half follows on `st`
      const int pz = c->ff_flip; c->ff_flip ^= 1;
      cudaStream_t sf = c->fast_front_stream;
      CDS_CUDA(cudaStreamWaitEvent(sf, c->ev_fb[pz], 0));
      if (int e = getframe_launch(c->w, c->h, n, hdr, tok, tok_base, c->d_mvx, c->d_mvy, c->d_dep, c->d_err, sf)) return e;
      if (int e = fast_front(c->fast, first_poc, n, c->d_mvx, c->d_mvy, c->d_dep, bytes, cam_out, sf)) return e;
      CDS_CUDA(cudaEventRecord(c->ev_ff[pz], sf));
      CDS_CUDA(cudaStreamWaitEvent(st, c->ev_ff[pz], 0));
      if (int e = fast_back(c->fast, first_poc, n, maps_out, c->p.out_u8, st)) return e;
      CDS_CUDA(cudaEventRecord(c->ev_fb[pz], st));
      return CDS_OK;
    }
    prof_mark(c, ST_GETFRAME, st);
    if (!c->maps_in) if (int e = getframe_launch(c->w, c->h, n, hdr, tok, tok_base, c->d_mvx, c->d_mvy, c->d_dep, c->d_err, st)) return e;
    if (before_back) CDS_CUDA(cudaStreamWaitEvent(st, before_back, 0));
    if (int e = fast_run_batch(c->fast, first_poc, n, c->d_mvx, c->d_mvy, c->d_dep, bytes, maps_out, c->p.out_u8, cam_out, st)) return e;
    prof_mark(c, -1, st);
    return CDS_OK;
  }
  const bool ov = c->overlap && c->p.use_rbf && !c->prof_on;
  const int slot = ov ? c->xs_flip : 0;
  if (ov) c->xs_flip ^= 1;
  c->last_xs = slot;
  float* xsvm = c->d_xsvm2[slot];
  if (!ov) if (int e = flush_pending_back(c, nullptr)) return e;          // a mode switch (profiling on) must not strand a deferred half
  const int h4 = c->h4, w4 = c->w4, n4 = c->n4, H = c->h, W = c->w, RL = c->RL;
  const float* fparam = c->d_fparam;
  // ---- PRE stage: stages (a), (c), smoothing, thresholds, temporal, contrast inputs.  With `pov` it runs on the context's pre
  //      stream, so that batch i+1's pre stage (latency / bandwidth bound) overlaps the map passes of batch i on `st`.  It may start
  //      when (1) the caller's inputs are ready (ev_call, once per API call), (2) the main stage of batch i-1 (same buffer set two
  //      batches ago: ev_front[slot]) has finished reading this set of d_X / padded inputs, (3) the contrast kernels of the previous
  //      batch are done (ev_c48: the issue-bound phase stays undisturbed).  Everything that touches the temporal rings lives here,
  //      in one stream, in picture order. ----
  const bool pov = ov && c->pre_ovl && !c->maps_in;
  const int set = pov ? slot : 0;                                        // buffer set = xsvm slot: both flip once per batch
  c->last_set = set;
  cudaStream_t sp = pov ? pre_stream_for(c, st) : st;
  c->last_pre_stream = sp;
  float* dX = c->d_X2[set]; float* dPadB = c->d_padB2[set]; float* dPadD = c->d_padD2[set]; float* dPadM = c->d_padM2[set];
  if (pov) {
    CDS_CUDA(cudaStreamWaitEvent(sp, c->ev_front[set], 0));
    CDS_CUDA(cudaStreamWaitEvent(sp, c->ev_c48, 0));
  }
  // Where the previous batch's deferred back half (persistent MUFU-bound SVM, then the final-map chain) joins this batch's front
  // half (CDS_BACK_AT): 1 (default) = at the mid-point, beside the FP32-bound contrast kernels; 0 = at its start; 2 = after the
  // contrast kernels (beside the H pass; the V pass needs whole SMs and waits); 3 = after the V pass, beside the HBM / L2-bound
  // statistics and resize passes.  Measured at 1080p / nSV 4096 / 64 pictures: 10.35 ms per batch for 1, 10.47 for 0, 11.4 for 2 and
  // for 3, 13 without any overlap.  The bandwidth-bound passes are poor partners: the resident SVM CTA holds 63 % of the register
  // file, and with a third of their usual threads in flight they crawl until it leaves (3: the 1.3 ms of passes end 1.2 ms AFTER
  // the 3.0 ms SVM); the FP32-bound contrast kernel needs few threads to keep its pipe busy and keeps 58 % of its speed.
  static const int back_at = [] { const char* e = getenv("CDS_BACK_AT"); const int v = e ? atoi(e) : 1; return v < 0 || v > 3 ? 1 : v; }();
  if (ov && c->pb.valid && back_at == 0) {
    trace_mark(c, c->tr_batches, TR_MID, st);
    CDS_CUDA(cudaEventRecord(c->ev_mid, st));
    if (int e = flush_pending_back(c, c->ev_mid)) return e;
  }
  if (int e = run_context_stages(c, first_poc, n, bytes, hdr, tok, tok_base, sp)) return e;
  if (cam_out) {      // cam_out is a DEVICE pointer here: gather the ring slots of this batch (the pre stage owns the rings)
    gather_cam_kernel<<<ceil_div(2 * n, 64), 64, 0, sp>>>(c->r_cam, RL, first_poc, n, cam_out);
    CDS_LAUNCH_CHECK();
  }
  // ---- thresholded smoothed maps and temporal features -> X slots 0,3,6 / 1,4,7 ----
  prof_mark(c, ST_THRESH, sp);
  if (int e = myst_small_launch(n4, H, W, RL, first_poc, n, c->d_mv_s, c->r_sm4, dX, sp)) return e;
  prof_mark(c, ST_TEMPORAL, sp);
  if (int e = temporal_launch(h4, w4, c->PT, RL, first_poc, n, c->r_sm4, c->r_cam, c->d_wtab, dX, sp)) return e;
  // ---- (b) contrast: bit (cusize 16, delta 3, win 5), depth (2, 13, 24), mvx / mvy (1, 11, 48)  main.m:226-235 ----
  {
    const size_t nk = (size_t)2 * n * 2;
    fill_u32_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, sp>>>(c->d_keys, nk, 0xffffffffu, 0u);
    CDS_LAUNCH_CHECK();
    unsigned *kB = c->d_keys, *kD = c->d_keys + 2 * n;
    unsigned *mB = c->d_maxbits, *mD = c->d_maxbits + n, *mX = c->d_maxbits + 2 * n, *mY = c->d_maxbits + 3 * n;
    const int sbR = (h4 + 15) / 16, sbC = (w4 + 15) / 16, sdR = (h4 + 1) / 2, sdC = (w4 + 1) / 2;
    float g[64];
    prof_mark(c, ST_CONTRAST_PREP, sp);
    // block mean (+1e-8), zero pad, pm_norm of the padded map (cu_contrast5.m:8-32, Sudoku_contrastcpp5.m:18-34)
    if (int e = contrast_sub_launch(dX + (size_t)3 * n4, (long)9 * n4, 0, 0, n, h4, w4, 16, c->d_subB, kB, sp)) return e;
    if (int e = contrast_pad_launch(c->d_subB, kB, n, sbR, sbC, 2, dPadB, sp)) return e;
    if (int e = contrast_sub_launch(dX + (size_t)6 * n4, (long)9 * n4, 0, 0, n, h4, w4, 2, c->d_subD, kD, sp)) return e;
    if (int e = contrast_pad_launch(c->d_subD, kD, n, sdR, sdC, 12, dPadD, sp)) return e;
    // mvx | mvy (cusize 1) in double up to the pm_norm; d_cxy holds [mvx: n maps | mvy: n maps] (written by smooth_launch)
    {
      const size_t nk64 = (size_t)2 * n * 2;
      fill_u64_kernel<<<(unsigned)((nk64 + 255) / 256), 256, 0, sp>>>(c->d_keys64, nk64, ~0ull, 0ull);
      CDS_LAUNCH_CHECK();
      if (int e = contrast_prep_f64_launch(c->d_cxy, 2 * n, h4, w4, 24, c->d_keys64, dPadM, sp)) return e;
    }
    // ---- MAIN stage on `st`: waits for this batch's pre stage and for the back half of batch i-2 (it read xsvm[slot]) ----
    if (pov) { CDS_CUDA(cudaEventRecord(c->ev_pre[set], sp)); CDS_CUDA(cudaStreamWaitEvent(st, c->ev_pre[set], 0)); }
    CDS_CUDA(cudaMemsetAsync(c->d_maxbits, 0, (size_t)4 * n * 4, st));
    // ---- mid-point: the previous batch's persistent SVM kernel starts together with this batch's contrast kernels ----
    int co_ctas = 0;
    if (back_at == 1) trace_mark(c, c->tr_batches, TR_MID, st);
    if (ov && c->pb.valid && back_at == 1) {
      CDS_CUDA(cudaEventRecord(c->ev_mid, st));
      if (int e = flush_pending_back(c, c->ev_mid)) return e;
      // CDS_CO_CTAS=n > 0: persistent contrast kernels capped at n CTAs per SM (a scheduler block hands out the tiles).  Default 0:
      // the ordinary one-tile-per-CTA grid -- the SVM CTAs are placed first (high-priority stream, launched just above) and the
      // registers they leave (65 536 - 320 x 128) hold three 192-thread, 40-register contrast CTAs per SM anyway; measured
      // 10.49 ms per step against 10.57 (n = 3) and 10.55 (n = 2) with the packed contrast kernel.
      static const int co = [] { const char* e = getenv("CDS_CO_CTAS"); const int v = e ? atoi(e) : 0; return v < 0 ? 0 : v; }();
      co_ctas = co;
    }
    // bit: 5x5 window on the CTU grid
    prof_mark(c, ST_CONTRAST_W5, st);
    if (int e = contrast_generic_launch(dPadB, c->d_w5, c->d_rawB, mB, n, sbR, sbC, 5, 2, st)) return e;
    // depth: 24x24 window on the 8-px grid
    prof_mark(c, ST_CONTRAST_W24, st);
    gaussian_factor(24, 13.0, g);
    if (int e = contrast_sep_launch(dPadD, c->d_rawD, mD, n, sdR, sdC, 24, g, g, nullptr, st, co_ctas, co_ctas ? c->d_sched : nullptr)) return e;
    // mvx | mvy as 2n maps in one launch: 48x48 window on the 4-px grid (the dominant FP32 kernel)
    prof_mark(c, ST_CONTRAST_W48, st);
    trace_mark(c, c->tr_batches, TR_C48_START, st);
    gaussian_factor(48, 11.0, g);
    if (int e = contrast_sep_launch(dPadM, c->d_rawM, mX, 2 * n, h4, w4, 48, g, g, nullptr, st, co_ctas, co_ctas ? c->d_sched + 2048 : nullptr)) return e;   // mX,mY contiguous
    trace_mark(c, c->tr_batches, TR_C48_END, st);
    if (pov) CDS_CUDA(cudaEventRecord(c->ev_c48, st));                    // the next batch's pre stage may join from here on
    if (ov && c->pb.valid && back_at == 2) {
      trace_mark(c, c->tr_batches, TR_MID, st);
      CDS_CUDA(cudaEventRecord(c->ev_mid, st));
      if (int e = flush_pending_back(c, c->ev_mid)) return e;
    }
    prof_mark(c, ST_CONTRAST_POST, st);
    if (int e = contrast_post_launch(c->d_rawB, mB, n, sbR, sbC, 16, h4, w4, dX, 5, st)) return e;
    if (int e = contrast_post_launch(c->d_rawD, mD, n, sdR, sdC, 2, h4, w4, dX, 8, st)) return e;
    if (int e = contrast_post_mv_launch(c->d_rawM, c->d_rawM + (size_t)n * n4, mX, mY, n, h4, w4, dX, 2, st)) return e;
  }
  // ---- fourX2 + imfilter(gaussian) at full resolution: statistics (and the temporal maps themselves) ----
  int nb = 0;
  prof_mark(c, ST_FULLRES_H, st);
  if (int e = fullres_hpass_launch(dX, 9 * n, h4, w4, c->poly, c->d_T, st)) return e;
  prof_mark(c, ST_FULLRES_V, st);
  if (int e = fullres_vpass_launch(c->d_T, 9 * n, h4, W, c->poly, c->d_Y, c->d_store_index, c->d_partial, &nb, st)) return e;
  if (ov && c->pb.valid && back_at == 3) {        // after the V pass (which needs whole SMs): beside the HBM / L2-bound statistics and resize passes
    trace_mark(c, c->tr_batches, TR_MID, st);
    CDS_CUDA(cudaEventRecord(c->ev_mid, st));
    if (int e = flush_pending_back(c, c->ev_mid)) return e;
  }
  prof_mark(c, ST_TEMPORAL_STATS, st);
  if (int e = stats_finalize_launch(c->d_partial, 9 * n, nb, c->d_stats9, st)) return e;
  {
    dim3 grid(c->nblk2, 3 * n);
    temporal_thresh_partial_kernel<<<grid, 256, 0, st>>>(c->d_Y, c->HW, c->d_stats9, c->d_partial2, c->p.use_rbf ? 0 : 1);
    CDS_LAUNCH_CHECK();
    temporal_thresh_finalize_kernel<<<3 * n, 32, 0, st>>>(c->d_partial2, c->nblk2, c->d_stats9, c->d_xf);
    CDS_LAUNCH_CHECK();
  }
  if (c->p.use_rbf) {
    // ---- features at 200x200 (main.m:275-294) ----
    prof_mark(c, ST_RESIZE200, st);
    // the cell path serves features 0, 2 of every triple only: the middle one (temporal) comes from the full-resolution path below
    if (int e = banded_rows_launch(dX, 9 * n, h4, w4, c->Av, nullptr, c->d_A1, st, 1)) return e;
    if (int e = banded_cols_launch(c->d_A1, 9 * n, 200, w4, c->Ah, c->d_F200, st, 1)) return e;
    if (int e = banded_rows_launch(c->d_Y, 3 * n, H, W, c->Rv, c->d_xf, c->d_R1, st)) return e;
    if (int e = banded_cols_launch(c->d_R1, 3 * n, 200, W, c->Rh, c->d_Ft, st)) return e;
    if (ov) CDS_CUDA(cudaStreamWaitEvent(st, c->ev_back[slot], 0));     // the back half of batch i-2 read xsvm[slot]
    {
      dim3 grid(ceil_div(40000, 256), n);
      assemble_features_kernel<<<grid, 256, 0, st>>>(c->d_F200, c->d_Ft, c->d_stats9, c->d_xf, fparam, xsvm);
      CDS_LAUNCH_CHECK();
    }
    if (ov) {          // the back half is deferred to the next batch's mid-point (or to join_back)
      trace_mark(c, c->tr_batches, TR_FRONT_END, st);
      c->tr_batches++;
      CDS_CUDA(cudaEventRecord(c->ev_front[slot], st));
      c->pb.valid = true; c->pb.slot = slot; c->pb.n = n; c->pb.tag = tag; c->pb.maps_out = maps_out; c->pb.before_back = before_back;
      return CDS_OK;
    }
    if (before_back) CDS_CUDA(cudaStreamWaitEvent(st, before_back, 0));
    if (int e = launch_back_rbf(c, slot, n, maps_out, st, true)) return e;
  } else {
    if (before_back) CDS_CUDA(cudaStreamWaitEvent(st, before_back, 0));
    prof_mark(c, ST_FINAL, st);
    dim3 grid(128, n);
    linear_comb_kernel<<<grid, 256, 0, st>>>(c->d_Y, c->HW, c->d_stats9, c->d_xf, fparam, c->d_dist, c->d_raw, c->d_fpartial);
    CDS_LAUNCH_CHECK();
    if (int e = stats_finalize_launch(c->d_fpartial, n, 128, c->d_fstats, st)) return e;
    if (maps_out) {
      if (int e = c->p.out_u8 ? launch_normalize<uint8_t>(c, n, maps_out, st) : launch_normalize<float>(c, n, maps_out, st)) return e;
    }
  }
  c->retired_tag = tag; c->retired_ev = nullptr;
  prof_mark(c, -1, st);
  return CDS_OK;
}

// make `st` wait for every back half still in flight (end of an API call: the caller's stream then covers all the work)
int join_back(cds_ctx* c, cudaStream_t st) {
  if (int e = pre_join(c, st)) return e;
  if (int e = flush_pending_back(c, nullptr)) return e;     // no batch follows: the last back half runs on its own
  if (c->overlap && c->p.use_rbf)
    for (int b = 0; b < 2; b++) CDS_CUDA(cudaStreamWaitEvent(st, c->ev_back[b], 0));
  return CDS_OK;
}

}  // namespace

extern "C" int cds_clip_context_dev(cds_ctx* c, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                                    const int16_t* tok, const int64_t* tok_base, void* stream) {
  if (int e = ensure_device()) return e;
  if (!c || !ctu_bytes || !hdr || !tok || !tok_base || nframes < 1) { set_error("cds_clip_context_dev: bad arguments"); return CDS_EINVAL; }
  if (first_poc != c->next_poc) { set_error("cds_clip_context_dev: expected poc %d, got %d (use cds_seek)", c->next_poc, first_poc); return CDS_EINVAL; }
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  cudaStream_t sp = pre_stream_for(c, st);
  for (int done = 0; done < nframes; done += c->B) {
    const int n = std::min(c->B, nframes - done);
    if (int e = run_context_stages(c, first_poc + done, n, ctu_bytes + (size_t)done * c->nctu, hdr + (size_t)done * c->nctu * 4, tok,
                                   tok_base + done, sp)) return e;
  }
  if (int e = pre_join(c, st)) return e;
  c->next_poc = first_poc + nframes;
  return CDS_OK;
}

extern "C" int cds_clip_process_dev(cds_ctx* c, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                                    const int16_t* tok, const int64_t* tok_base, void* maps_out, int* cam_out, void* stream) {
  if (int e = ensure_device()) return e;
  if (!c || !ctu_bytes || !hdr || !tok || !tok_base || nframes < 1) { set_error("cds_clip_process_dev: bad arguments"); return CDS_EINVAL; }
  if (first_poc != c->next_poc) { set_error("cds_clip_process_dev: expected poc %d, got %d (use cds_seek)", c->next_poc, first_poc); return CDS_EINVAL; }
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const size_t px = (size_t)c->HW * (c->p.out_u8 ? 1 : 4);
  if (c->fast && c->fast_ovl && nframes > c->B) {          // several batches in one call: let the halves of consecutive batches overlap
    CDS_CUDA(cudaEventRecord(c->ev_fstart, st));
    CDS_CUDA(cudaStreamWaitEvent(c->fast_front_stream, c->ev_fstart, 0));
    c->fast_ovl_active = true;
  }
  int rc = CDS_OK;
  for (int done = 0; done < nframes && !rc; done += c->B) {
    const int n = std::min(c->B, nframes - done);
    rc = run_batch(c, first_poc + done, n, ctu_bytes + (size_t)done * c->nctu, hdr + (size_t)done * c->nctu * 4, tok,
                   tok_base + done, maps_out ? (char*)maps_out + (size_t)done * px : nullptr,
                   cam_out ? cam_out + 2 * done : nullptr, st);
  }
  c->fast_ovl_active = false;
  if (rc) return rc;
  if (int e = join_back(c, st)) return e;
  c->next_poc = first_poc + nframes;
  return CDS_OK;
}

// host buffers in; batch i renders into d_out[i & 1] on the compute stream while batch i-1 drains to host_out on the copy
// stream (events order the two), so the device->host copy of the maps overlaps the next batch's kernels.
static int process_host(cds_ctx* c, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                        const int16_t* tok, const int64_t* tok_base, size_t ntok, void* host_out, int* cam_out,
                        bool context_only = false) {
  cudaStream_t st = c->stream, cs = c->copy_stream;
  const size_t px = (size_t)c->HW * (c->p.out_u8 ? 1 : 4);
  const int nb = (nframes + c->B - 1) / c->B;
  // validate the token ranges and size the staging buffer once (worst case 1237 tokens per CTU)
  size_t need = 0;
  for (int done = 0; done < nframes; done += c->B) {
    const int n = std::min(c->B, nframes - done);
    const int64_t t0 = tok_base[done], t1 = tok_base[done + n];
    if (t0 < 0 || t1 < t0 || (size_t)t1 > ntok) { set_error("cds_clip_process: tok_base out of range"); return CDS_EFORMAT; }
    need = std::max(need, (size_t)(t1 - t0));
  }
  if (need > c->tok_cap) {
    CDS_CUDA(cudaStreamSynchronize(st));
    const size_t want = need + need / 4;
    void* q = nullptr;
    if (cudaMalloc(&q, want * 2 + 256) != cudaSuccess) { set_error("cds_clip_process: cannot stage %zu tokens", want); return CDS_ENOMEM; }
    for (void*& a : c->allocs) if (a == (void*)c->d_tok) { cudaFree(a); a = q; }
    c->d_tok = reinterpret_cast<int16_t*>(q); c->tok_cap = want;
  }
  std::vector<int64_t> base((size_t)nb * (c->B + 1));       // lives until the final synchronisation below
  if (cam_out && !context_only && (size_t)2 * nframes > c->h_cam_cap) {
    if (c->h_cam_pin) cudaFreeHost(c->h_cam_pin);
    c->h_cam_pin = nullptr; c->h_cam_cap = 0;
    CDS_CUDA(cudaHostAlloc((void**)&c->h_cam_pin, (size_t)2 * nframes * sizeof(int), cudaHostAllocDefault));
    c->h_cam_cap = (size_t)2 * nframes;
  }
  if (cam_out && !context_only && (size_t)2 * nframes > c->d_cam_cap) {      // every entry point returns synchronised: nothing is in flight here
    void* q = nullptr;
    if (cudaMalloc(&q, (size_t)2 * nframes * sizeof(int) + 256) != cudaSuccess) { set_error("cds_clip_process: cannot allocate the camera-motion buffer"); return CDS_ENOMEM; }
    bool swapped = false;
    for (void*& a : c->allocs) if (c->d_cam_call && a == (void*)c->d_cam_call) { cudaFree(a); a = q; swapped = true; }
    if (!swapped) c->allocs.push_back(q);
    c->d_cam_call = reinterpret_cast<int*>(q); c->d_cam_cap = (size_t)2 * nframes;
  }
  // a batch's maps leave for the host when its back half has been enqueued ("retired"): at once without overlap, from the middle of
  // the next batch (or from join_back) with it
  struct HostBatch { int buf, done, n; };
  std::vector<HostBatch> hb((size_t)nb);
  auto drain_retired = [&]() -> int {
    if (c->retired_tag < 0) return CDS_OK;
    const HostBatch b = hb[(size_t)c->retired_tag];
    c->retired_tag = -1;
    if (c->retired_ev) CDS_CUDA(cudaStreamWaitEvent(cs, c->retired_ev, 0));
    else { CDS_CUDA(cudaEventRecord(c->ev_done[b.buf], st)); CDS_CUDA(cudaStreamWaitEvent(cs, c->ev_done[b.buf], 0)); }
    if (host_out) CDS_CUDA(cudaMemcpyAsync((char*)host_out + (size_t)b.done * px, c->d_out[b.buf], (size_t)b.n * px, cudaMemcpyDeviceToHost, cs));
    CDS_CUDA(cudaEventRecord(c->ev_copied[b.buf], cs));
    return CDS_OK;
  };
  int bi = 0;
  for (int done = 0; done < nframes; done += c->B, bi++) {
    const int n = std::min(c->B, nframes - done);
    const int64_t t0 = tok_base[done], t1 = tok_base[done + n];
    int64_t* bs = base.data() + (size_t)bi * (c->B + 1);
    for (int i = 0; i <= n; i++) bs[i] = tok_base[done + i] - t0;
    cudaStream_t sp = pre_stream_for(c, st);        // the staging buffers are read by stages (a) / (c): same stream, in order
    CDS_CUDA(cudaMemcpyAsync(c->d_bytes, ctu_bytes + (size_t)done * c->nctu, (size_t)n * c->nctu * 2, cudaMemcpyHostToDevice, sp));
    CDS_CUDA(cudaMemcpyAsync(c->d_hdr, hdr + (size_t)done * c->nctu * 4, (size_t)n * c->nctu * 16, cudaMemcpyHostToDevice, sp));
    CDS_CUDA(cudaMemcpyAsync(c->d_tok, tok + t0, (size_t)(t1 - t0) * 2, cudaMemcpyHostToDevice, sp));
    CDS_CUDA(cudaMemcpyAsync(c->d_tokbase, bs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, sp));
    if (context_only) {
      if (int e = run_context_stages(c, first_poc + done, n, c->d_bytes, c->d_hdr, c->d_tok, c->d_tokbase, sp)) return e;
      continue;
    }
    const int buf = c->out_flip; c->out_flip ^= 1;
    hb[bi] = {buf, done, n};
    // before_back = the copy that last read d_out[buf] has finished (a never-recorded event is a no-op)
    c->retired_tag = -1;
    if (int e = run_batch(c, first_poc + done, n, c->d_bytes, c->d_hdr, c->d_tok, c->d_tokbase, c->d_out[buf],
                          cam_out ? c->d_cam_call + (size_t)2 * done : nullptr, st, c->ev_copied[buf], bi)) return e;
    if (int e = drain_retired()) return e;
    c->last_first = first_poc + done; c->last_n = n; c->last_buf = buf;
  }
  c->retired_tag = -1;
  int herr = 0;
  if (int e = join_back(c, st)) return e;
  if (!context_only) if (int e = drain_retired()) return e;
  CDS_CUDA(cudaMemcpyAsync(&herr, c->d_err, 4, cudaMemcpyDeviceToHost, st));
  if (cam_out && !context_only)     // one copy for the call, after every front half (the gathers run on st)
    CDS_CUDA(cudaMemcpyAsync(c->h_cam_pin, c->d_cam_call, (size_t)2 * nframes * sizeof(int), cudaMemcpyDeviceToHost, st));
  CDS_CUDA(cudaStreamSynchronize(st));
  CDS_CUDA(cudaStreamSynchronize(cs));
  if (cam_out && !context_only) memcpy(cam_out, c->h_cam_pin, (size_t)2 * nframes * sizeof(int));
  if (herr) {
    CDS_CUDA(cudaMemset(c->d_err, 0, 4));
    set_error("cds_clip_process: malformed CU token stream in pictures %d..%d", first_poc, first_poc + nframes - 1); return CDS_EFORMAT;
  }
  c->next_poc = first_poc + nframes;
  return CDS_OK;
}

extern "C" int cds_clip_context(cds_ctx* c, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                                const int16_t* tok, const int64_t* tok_base, size_t ntok) {
  if (int e = ensure_device()) return e;
  if (!c || !ctu_bytes || !hdr || !tok || !tok_base || nframes < 1) { set_error("cds_clip_context: bad arguments"); return CDS_EINVAL; }
  if (first_poc != c->next_poc) { set_error("cds_clip_context: expected poc %d, got %d (use cds_seek)", c->next_poc, first_poc); return CDS_EINVAL; }
  if (c->pending) { set_error("cds_clip_context: pictures submitted through cds_submit_frame are still pending"); return CDS_EINVAL; }
  if (int e = process_host(c, first_poc, nframes, ctu_bytes, hdr, tok, tok_base, ntok, nullptr, nullptr, true)) return e;
  c->pending_first = c->next_poc;
  return CDS_OK;
}

extern "C" int cds_clip_process(cds_ctx* c, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                                const int16_t* tok, const int64_t* tok_base, size_t ntok, void* maps_out, int* cam_out) {
  if (int e = ensure_device()) return e;
  if (!c || !ctu_bytes || !hdr || !tok || !tok_base || nframes < 1) { set_error("cds_clip_process: bad arguments"); return CDS_EINVAL; }
  if (first_poc != c->next_poc) { set_error("cds_clip_process: expected poc %d, got %d (use cds_seek)", c->next_poc, first_poc); return CDS_EINVAL; }
  if (c->pending) { set_error("cds_clip_process: %d pictures submitted through cds_submit_frame are still pending (cds_flush first)", c->pending); return CDS_EINVAL; }
  if (int e = process_host(c, first_poc, nframes, ctu_bytes, hdr, tok, tok_base, ntok, maps_out, cam_out)) return e;
  c->pending_first = c->next_poc;
  return CDS_OK;
}

// stage-(a) maps in (main.m:25-27,38-55 exchange format), maps out; one batch at a time, no copy / compute overlap
extern "C" int cds_clip_process_maps(cds_ctx* c, int first_poc, int nframes, const int16_t* mvx, const int16_t* mvy,
                                     const uint8_t* depth, const uint16_t* ctu_bytes, void* maps_out, int* cam_out) {
  if (int e = ensure_device()) return e;
  if (!c || !mvx || !mvy || !depth || !ctu_bytes || nframes < 1) { set_error("cds_clip_process_maps: bad arguments"); return CDS_EINVAL; }
  if (first_poc != c->next_poc) { set_error("cds_clip_process_maps: expected poc %d, got %d (use cds_seek)", c->next_poc, first_poc); return CDS_EINVAL; }
  if (c->pending) { set_error("cds_clip_process_maps: pictures submitted through cds_submit_frame are still pending"); return CDS_EINVAL; }
  cudaStream_t st = c->stream;
  const size_t px = (size_t)c->HW * (c->p.out_u8 ? 1 : 4), n4 = c->n4;
  std::vector<int> cam((size_t)2 * c->B);
  int rc = CDS_OK;
  c->maps_in = true;
  for (int done = 0; done < nframes && !rc; done += c->B) {
    const int n = std::min(c->B, nframes - done);
    cudaError_t ce = cudaMemcpyAsync(c->d_mvx, mvx + (size_t)done * n4, (size_t)n * n4 * 2, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->d_mvy, mvy + (size_t)done * n4, (size_t)n * n4 * 2, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->d_dep, depth + (size_t)done * n4, (size_t)n * n4, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->d_bytes, ctu_bytes + (size_t)done * c->nctu, (size_t)n * c->nctu * 2, cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) { set_error("cds_clip_process_maps: copy in: %s", cudaGetErrorString(ce)); rc = CDS_ECUDA; break; }
    if (!maps_out) { rc = run_context_stages(c, first_poc + done, n, c->d_bytes, nullptr, nullptr, nullptr, st); }
    else {
      const int buf = c->out_flip; c->out_flip ^= 1;
      rc = run_batch(c, first_poc + done, n, c->d_bytes, nullptr, nullptr, nullptr, c->d_out[buf], c->d_cam_stage, st);
      if (!rc) rc = join_back(c, st);
      if (!rc) {
        ce = cudaMemcpyAsync((char*)maps_out + (size_t)done * px, c->d_out[buf], (size_t)n * px, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(cam.data(), c->d_cam_stage, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
        if (ce != cudaSuccess) { set_error("cds_clip_process_maps: copy out: %s", cudaGetErrorString(ce)); rc = CDS_ECUDA; }
      }
      c->last_first = first_poc + done; c->last_n = n; c->last_buf = buf;
    }
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { set_error("cds_clip_process_maps: %s", cudaGetErrorString(cudaGetLastError())); rc = CDS_ECUDA; }
    if (!rc && maps_out && cam_out) memcpy(cam_out + (size_t)2 * done, cam.data(), (size_t)n * 8);
  }
  c->maps_in = false;
  if (rc) return rc;
  c->next_poc = first_poc + nframes; c->pending_first = c->next_poc;
  return CDS_OK;
}

extern "C" int cds_probe_seq_sum(const float* v, const unsigned char* sel, int n, float* sum, int* count) { return fast_probe_seq_sum(v, sel, n, sum, count); }
extern "C" int cds_probe_run_sum(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, float* sum, int* count) {
  return fast_probe_run_sum(vals, gh, gw, rows, cols, ms, t, sum, count);
}

// ---- hook-style API (TDecCu.cpp:153-289 capture, TDecGop.cpp:270-332 per-slice tail): one picture at a time, batched internally,
//      asynchronous: pinned staging in, stream-ordered H2D + kernels + D2H into pinned result buffers, no host wait before cds_get_map ----
namespace {

void hook_reset(cds_ctx* c) {
  for (auto& S : c->hs) { S.n = 0; S.tok_used = 0; S.in_flight = false; }
  for (auto& O : c->ho) { O.first = -1; O.n = 0; O.copy_queued = false; }
  c->acquired = false;
}

void hook_free(cds_ctx* c) {
  for (auto& S : c->hs) {
    if (S.bytes) cudaFreeHost(S.bytes); if (S.hdr) cudaFreeHost(S.hdr); if (S.tok) cudaFreeHost(S.tok); if (S.base) cudaFreeHost(S.base);
    if (S.h2d_done) cudaEventDestroy(S.h2d_done);
    S = cds_ctx::HookStage();
  }
  for (auto& O : c->ho) {
    if (O.maps) cudaFreeHost(O.maps); if (O.cam) cudaFreeHost(O.cam);
    if (O.ready) cudaEventDestroy(O.ready); if (O.cam_ready) cudaEventDestroy(O.cam_ready);
    O = cds_ctx::HookOut();
  }
  c->hook_ready = false;
}

// first use of the hook API: pinned staging / result buffers (contexts that only see the batch entry points never pay for them)
int hook_init(cds_ctx* c) {
  if (c->hook_ready) return CDS_OK;
  const size_t px = (size_t)c->HW * (c->p.out_u8 ? 1 : 4);
  c->hook_pic_tok = (size_t)c->nctu * 1300;                  // worst case 1237 tokens per CTU (85 + 256 + 896)
  const size_t cap = (size_t)c->B * c->nctu * 320 + c->hook_pic_tok;   // typical streams: < 320 per CTU; a batch closes early when a worst-case picture might not fit
  cudaError_t e = cudaSuccess;
  for (auto& S : c->hs) {
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&S.bytes, (size_t)c->B * c->nctu * 2, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&S.hdr, (size_t)c->B * c->nctu * 16, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&S.tok, cap * 2, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&S.base, (size_t)(c->B + 1) * 8, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.h2d_done, cudaEventDisableTiming);
    S.tok_cap = cap;
  }
  for (auto& O : c->ho) {
    if (e == cudaSuccess) e = cudaHostAlloc(&O.maps, (size_t)c->B * px, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&O.cam, (size_t)c->B * 8, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&O.ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&O.cam_ready, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) { set_error("hook buffers: %s", cudaGetErrorString(e)); hook_free(c); return CDS_ENOMEM; }
  if (cap > c->tok_cap) {                                     // device token staging holds a whole staging batch
    CDS_CUDA(cudaStreamSynchronize(c->stream));
    void* q = nullptr;
    if (cudaMalloc(&q, cap * 2 + 256) != cudaSuccess) { set_error("hook buffers: cannot stage %zu tokens on the device", cap); hook_free(c); return CDS_ENOMEM; }
    for (void*& a : c->allocs) if (a == (void*)c->d_tok) { cudaFree(a); a = q; }
    c->d_tok = reinterpret_cast<int16_t*>(q); c->tok_cap = cap;
  }
  if (!c->d_cam_hook) if (int er = dalloc(c, &c->d_cam_hook, (size_t)4 * c->B)) { hook_free(c); return er; }
  hook_reset(c);
  c->hook_ready = true;
  return CDS_OK;
}

// a back half has been enqueued (run_batch / join_back set retired_tag = result slot): queue its maps for the host
int hook_drain_retired(cds_ctx* c) {
  if (c->retired_tag < 0) return CDS_OK;
  const int slot = c->retired_tag & 1, buf = c->retired_tag >> 1;
  c->retired_tag = -1;
  cds_ctx::HookOut& O = c->ho[slot];
  cudaStream_t cs = c->copy_stream;
  const size_t px = (size_t)c->HW * (c->p.out_u8 ? 1 : 4);
  if (c->retired_ev) CDS_CUDA(cudaStreamWaitEvent(cs, c->retired_ev, 0));
  else { CDS_CUDA(cudaEventRecord(c->ev_done[buf], c->stream)); CDS_CUDA(cudaStreamWaitEvent(cs, c->ev_done[buf], 0)); }
  CDS_CUDA(cudaMemcpyAsync(O.maps, c->d_out[buf], (size_t)O.n * px, cudaMemcpyDeviceToHost, cs));
  CDS_CUDA(cudaEventRecord(c->ev_copied[buf], cs));
  CDS_CUDA(cudaEventRecord(O.ready, cs));
  O.copy_queued = true;
  return CDS_OK;
}

// enqueue the staged pictures: H2D, kernel chain, D2H — and return
int hook_launch(cds_ctx* c) {
  cds_ctx::HookStage& S = c->hs[c->hs_cur];
  if (S.n == 0) return CDS_OK;
  cudaStream_t st = c->stream;
  const int n = S.n, first = S.first;
  S.base[n] = (int64_t)S.tok_used;
  // whatever happens below, these pictures are consumed: the next picture expected is first + n (a failed batch does not wedge the context)
  c->pending = 0; c->pending_first = first + n; c->acquired = false;
  c->hs_cur ^= 1;
  int rc = CDS_OK;
  auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && !rc) { set_error("cds_submit_frame: %s: %s", what, cudaGetErrorString(e)); rc = CDS_ECUDA; } };
  cudaStream_t sp = pre_stream_for(c, st);          // staging is read by stages (a) / (c): same stream, in order
  cu(cudaMemcpyAsync(c->d_bytes, S.bytes, (size_t)n * c->nctu * 2, cudaMemcpyHostToDevice, sp), "copy in");
  cu(cudaMemcpyAsync(c->d_hdr, S.hdr, (size_t)n * c->nctu * 16, cudaMemcpyHostToDevice, sp), "copy in");
  cu(cudaMemcpyAsync(c->d_tok, S.tok, S.tok_used * 2, cudaMemcpyHostToDevice, sp), "copy in");
  cu(cudaMemcpyAsync(c->d_tokbase, S.base, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, sp), "copy in");
  cu(cudaEventRecord(S.h2d_done, sp), "event");
  S.in_flight = true; S.n = 0; S.tok_used = 0;
  if (rc) return rc;
  const int slot = c->ho_cur; c->ho_cur ^= 1;
  const int buf = c->out_flip; c->out_flip ^= 1;
  cds_ctx::HookOut& O = c->ho[slot];
  O.first = first; O.n = n; O.copy_queued = false;
  int* dcam = c->d_cam_hook + (size_t)slot * 2 * c->B;
  c->retired_tag = -1;
  rc = run_batch(c, first, n, c->d_bytes, c->d_hdr, c->d_tok, c->d_tokbase, c->d_out[buf], dcam, st, c->ev_copied[buf], slot | (buf << 1));
  if (rc) { O.first = -1; O.n = 0; return rc; }
  // camera motion: gathered by the pre stage; its copy stays on that stream, so the gather of batch k+2 (same slot) cannot overtake it
  cudaStream_t sg = c->last_pre_stream ? c->last_pre_stream : st;
  cu(cudaMemcpyAsync(O.cam, dcam, (size_t)n * 8, cudaMemcpyDeviceToHost, sg), "copy out");
  cu(cudaEventRecord(O.cam_ready, sg), "event");
  if (rc) return rc;
  if (int e = hook_drain_retired(c)) return e;
  c->next_poc = first + n;
  c->last_first = first; c->last_n = n; c->last_buf = buf;
  return CDS_OK;
}

}  // namespace

/* Hook 2: pinned buffers for the NEXT picture (TDecCu::decompressCU / MVoutput write their tokens here instead of the
 * float[LCUNUM][1000] statics).  Valid until the matching cds_submit_frame. */
extern "C" int cds_acquire_frame_buffers(cds_ctx* c, uint16_t** ctu_bytes, int32_t** hdr, int16_t** tok, size_t* tok_capacity) {
  if (int e = ensure_device()) return e;
  if (!c || !ctu_bytes || !hdr || !tok) { set_error("cds_acquire_frame_buffers: null argument"); return CDS_EINVAL; }
  if (int e = hook_init(c)) return e;
  cds_ctx::HookStage* S = &c->hs[c->hs_cur];
  if (S->n > 0 && S->tok_cap - S->tok_used < c->hook_pic_tok) {          // a worst-case picture might not fit: close the batch early
    if (int e = hook_launch(c)) return e;
    S = &c->hs[c->hs_cur];
  }
  if (S->n == 0 && S->in_flight) { CDS_CUDA(cudaEventSynchronize(S->h2d_done)); S->in_flight = false; }   // its previous batch has left for the device
  *ctu_bytes = S->bytes + (size_t)S->n * c->nctu;
  *hdr = S->hdr + (size_t)S->n * c->nctu * 4;
  *tok = S->tok + S->tok_used;
  if (tok_capacity) *tok_capacity = S->tok_cap - S->tok_used;
  c->acquired = true;
  return CDS_OK;
}

extern "C" int cds_submit_frame(cds_ctx* c, int poc, const uint16_t* ctu_bytes, const int32_t* hdr, const int16_t* tok, size_t ntok) {
  if (int e = ensure_device()) return e;
  if (!c || !ctu_bytes || !hdr || !tok) { set_error("cds_submit_frame: null argument"); return CDS_EINVAL; }
  if (poc != c->pending_first + c->pending) { set_error("cds_submit_frame: pictures must arrive in POC order (expected %d, got %d)", c->pending_first + c->pending, poc); return CDS_EINVAL; }
  if (ntok > (size_t)c->nctu * 1300) { set_error("cds_submit_frame: %zu tokens for %d CTUs", ntok, c->nctu); return CDS_EFORMAT; }
  if (int e = hook_init(c)) return e;
  uint16_t* sb; int32_t* sh; int16_t* stk; size_t cap;
  const bool was_acquired = c->acquired;
  if (int e = cds_acquire_frame_buffers(c, &sb, &sh, &stk, &cap)) return e;      // (re)computes the slot of this picture; may close a full batch
  if (!(was_acquired && ctu_bytes == sb && hdr == sh && tok == stk)) {          // caller's own buffers: one copy into the pinned slot
    memcpy(sb, ctu_bytes, (size_t)c->nctu * 2);
    memcpy(sh, hdr, (size_t)c->nctu * 16);
    memcpy(stk, tok, ntok * 2);
  }
  cds_ctx::HookStage& S = c->hs[c->hs_cur];
  if (S.n == 0) S.first = poc;
  S.base[S.n] = (int64_t)S.tok_used;
  S.tok_used += ntok; S.n++;
  c->pending = S.n; c->pending_first = S.first; c->acquired = false;
  if (S.n == c->B) return hook_launch(c);
  return CDS_OK;
}

/* Enqueue whatever is staged (a partial batch) and every deferred back half; does not wait for the GPU. */
extern "C" int cds_flush(cds_ctx* c) {
  if (!c) { set_error("cds_flush: null context"); return CDS_EINVAL; }
  if (c->hook_ready) if (int e = hook_launch(c)) return e;
  c->retired_tag = -1;
  if (int e = join_back(c, c->stream)) return e;
  if (c->hook_ready) if (int e = hook_drain_retired(c)) return e;
  return CDS_OK;
}

extern "C" int cds_get_map(cds_ctx* c, int poc, void* dst, int* cam_xy) {
  if (!c || !dst) { set_error("cds_get_map: null argument"); return CDS_EINVAL; }
  const size_t px = (size_t)c->HW * (c->p.out_u8 ? 1 : 4);
  if (c->hook_ready) {
    if (poc >= c->pending_first && poc < c->pending_first + c->pending) { if (int e = hook_launch(c)) return e; }
    for (auto& O : c->ho) {
      if (poc < O.first || poc >= O.first + O.n) continue;
      if (!O.copy_queued) {                                   // its back half is still deferred: no batch followed yet
        c->retired_tag = -1;
        if (int e = join_back(c, c->stream)) return e;
        if (int e = hook_drain_retired(c)) return e;
      }
      if (!O.copy_queued) { set_error("cds_get_map: picture %d was never rendered", poc); return CDS_EINVAL; }
      CDS_CUDA(cudaEventSynchronize(O.ready));
      CDS_CUDA(cudaEventSynchronize(O.cam_ready));
      int herr = 0;                                           // malformed token streams are reported where the maps are collected
      CDS_CUDA(cudaMemcpy(&herr, c->d_err, 4, cudaMemcpyDeviceToHost));
      if (herr) { CDS_CUDA(cudaMemset(c->d_err, 0, 4)); set_error("cds_get_map: malformed CU token stream at or before picture %d", poc); return CDS_EFORMAT; }
      memcpy(dst, (const char*)O.maps + (size_t)(poc - O.first) * px, px);
      if (cam_xy) { cam_xy[0] = O.cam[2 * (poc - O.first)]; cam_xy[1] = O.cam[2 * (poc - O.first) + 1]; }
      return CDS_OK;
    }
  }
  if (c->hook_ready && (c->ho[0].first >= 0 || c->ho[1].first >= 0)) {
    set_error("cds_get_map: picture %d is not in the two result batches that stay retrievable", poc); return CDS_EINVAL;
  }
  // batch entry points with maps_out == NULL leave their last batch on the device
  if (poc < c->last_first || poc >= c->last_first + c->last_n) {
    set_error("cds_get_map: picture %d is not in the most recent batch [%d, %d)", poc, c->last_first, c->last_first + c->last_n); return CDS_EINVAL;
  }
  CDS_CUDA(cudaMemcpy(dst, (char*)c->d_out[c->last_buf] + (size_t)(poc - c->last_first) * px, px, cudaMemcpyDeviceToHost));
  if (cam_xy && c->fast) return fast_get_cam(c->fast, poc, cam_xy);
  if (cam_xy) CDS_CUDA(cudaMemcpy(cam_xy, c->r_cam + 2 * (poc % c->RL), 8, cudaMemcpyDeviceToHost));
  return CDS_OK;
}

// ---- per-stage device timing -------------------------------------------------------------------
extern "C" const char* cds_profile_stage_names(void) { return kStageNames; }

extern "C" int cds_profile_enable(cds_ctx* c, int on) {
  if (!c) { set_error("cds_profile_enable: null context"); return CDS_EINVAL; }
  CDS_CUDA(cudaStreamSynchronize(c->stream));
  c->prof_on = on != 0; c->ev_used = 0; c->ev_marks.clear();
  if (c->fast) fast_profile(c->fast, c->prof_on);
  for (int i = 0; i < 24; i++) { c->prof_ms[i] = 0; c->prof_calls[i] = 0; }
  return CDS_OK;
}

// Synchronises the device, folds every recorded stage interval into per-stage totals and returns them
// (milliseconds and number of timed intervals per stage, ST_COUNT entries each) since cds_profile_enable.
extern "C" int cds_profile_read(cds_ctx* c, double* ms, long* calls, int n) {
  if (!c || !ms || n < ST_COUNT) { set_error("cds_profile_read: need room for %d stages", (int)ST_COUNT); return CDS_EINVAL; }
  CDS_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i + 1 < c->ev_marks.size(); i++) {
    const int id = c->ev_marks[i].first;
    if (id < 0) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, c->ev_marks[i].second, c->ev_marks[i + 1].second) == cudaSuccess) { c->prof_ms[id] += t; c->prof_calls[id]++; }
  }
  c->ev_marks.clear(); c->ev_used = 0;
  for (int i = 0; i < ST_COUNT; i++) { ms[i] = c->prof_ms[i]; if (calls) calls[i] = c->prof_calls[i]; }
  return CDS_OK;
}

// Test hook: copy an intermediate of the most recent batch to the host.  which: "xmaps" [9][n4], "stats" [9][4],
// "feat200" [9][40000] (raw, before pm_norm), "ft" [3][40000], "xsvm" [40000][9], "dec" [40000], "raw" [H][W].
extern "C" int cds_debug_copy(cds_ctx* c, const char* which, int frame_in_batch, void* dst, size_t bytes) {
  if (!c || !which || !dst || frame_in_batch < 0 || frame_in_batch >= c->B) { set_error("cds_debug_copy: bad arguments"); return CDS_EINVAL; }
  const int f = frame_in_batch; const void* src = nullptr; size_t n = 0;
  if (c->fast) return fast_debug_copy(c->fast, which, f, dst, bytes, c->stream);
  if (!strcmp(which, "xmaps")) { src = c->d_X2[c->last_set] + (size_t)f * 9 * c->n4; n = (size_t)9 * c->n4 * 4; }
  else if (!strcmp(which, "stats")) { src = c->d_stats9 + (size_t)f * 36; n = 36 * 4; }
  else if (!strcmp(which, "feat200")) { src = c->d_F200 + (size_t)f * 9 * 40000; n = (size_t)9 * 40000 * 4; }
  else if (!strcmp(which, "ft")) { src = c->d_Ft + (size_t)f * 3 * 40000; n = (size_t)3 * 40000 * 4; }
  else if (!strcmp(which, "xf")) { src = c->d_xf + (size_t)f * 12; n = 12 * 4; }
  else if (!strcmp(which, "xsvm")) { src = c->d_xsvm2[c->last_xs] + (size_t)f * 40000 * 9; n = (size_t)40000 * 9 * 4; }
  else if (!strcmp(which, "dec")) { src = c->d_dec + (size_t)f * 40000; n = (size_t)40000 * 4; }
  else if (!strcmp(which, "ymaps")) { src = c->d_Y + (size_t)f * (c->p.use_rbf ? 3 : 9) * c->HW; n = (size_t)(c->p.use_rbf ? 3 : 9) * c->HW * 4; }
  else if (!strcmp(which, "raw")) { src = c->d_raw + (size_t)f * c->HW; n = (size_t)c->HW * 4; }
  else { set_error("cds_debug_copy: unknown intermediate %s", which); return CDS_EINVAL; }
  if (bytes < n) { set_error("cds_debug_copy: %s needs %zu bytes", which, n); return CDS_EINVAL; }
  CDS_CUDA(cudaStreamSynchronize(c->stream));
  CDS_CUDA(cudaStreamSynchronize(c->back_stream));
  CDS_CUDA(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost));
  return CDS_OK;
}
