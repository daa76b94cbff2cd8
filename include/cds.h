/* cds.h — C ABI of libcds_b200.so: the B200-native compressed-domain saliency hot path.
 *
 * Plain pointers and sizes only (no torch / C++ types).  Every entry point returns 0 on success or
 * a negative CDS_E* code; cds_last_error() gives the text.  The library has NO CPU fallback: every
 * compute entry point fails with CDS_ENODEVICE when no sm_100 device is usable.
 *
 * Host-buffer entry points (cds_contrast5, cds_getframe, cds_svm_rbf_predict, cds_camera_motion,
 * cds_clip_process) copy host<->device inside the call.  The *_dev entry points take device
 * pointers and a cudaStream_t (passed as void*) and never synchronise.
 *
 * Reference interfaces replaced (paths relative to remega/Compressd_Domain_SaliencyPrediction):
 *   cds_contrast5        <- computecontrast5.cpp:52 mexFunction(pad, W, len, winsize, m, n)
 *                           called from Sudoku_contrastcpp5.m:37
 *   cds_sudoku_contrast  <- Sudoku_contrastcpp5.m:1 ; cds_cu_contrast5 <- cu_contrast5.m:1
 *   cds_getframe         <- HM_16.0_features/source/Lib/TLibDecoder/code_mat_h.h:575 getFrame()
 *   cds_syntax_pack_ref  <- TDecCu.h:100-106 statics bits/MV/CUPU/Pred (float[LCUNUM][1000], 999-terminated)
 *   cds_svm_*            <- libsvm-3.17/svm.h:87 svm_predict_values, matlab/svmpredict.c:272 libsvmpredict,
 *                           matlab/svm_model_matlab.c:211 matlab_matrix_to_model, svm.cpp:2641 svm_load_model
 *   cds_camera_motion    <- main.m:77-105 + cm_judge.m + mv_vote.m + direction_cluster.m
 *   cds_magnitude_vote   <- magnitude_vote.m:1
 *   cds_create(use_rbf = CDS_FUSION_FAST) <- the decoder's own per-slice arithmetic: TDecGop.cpp:225-333 (tail of decompressSlice),
 *                           :486-656 (CameraDect, vote_alg, max, Umean), salmat.cpp (ForwardFilter, MapThreshold, Generate*Map)
 *   cds_clip_process_maps <- main.m:25-27,38-55: the per-picture mv_x / mv_y / depth / bitmap arrays MATLAB reads from the decoder's
 *                           text files or .mat cell arrays (writer code_mat_h.cpp:28-57)
 *   cds_getframe_semantic <- TComDataCU::getDepth / getPredictionMode / getCUMvField (the partitions TDecCu::MVoutput walks, TDecCu.cpp:165-289)
 *   cds_clip_*           <- main.m:113-305 (per-clip loop) / TDecGop.cpp:225-333 (per-slice hook)
 */
#ifndef CDS_H
#define CDS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CDS_OK 0
#define CDS_EINVAL (-1)     /* bad argument */
#define CDS_ENODEVICE (-2)  /* no usable CUDA device (the library never falls back to the CPU) */
#define CDS_ECUDA (-3)      /* CUDA runtime error, see cds_last_error() */
#define CDS_ENOMEM (-4)
#define CDS_EFORMAT (-5)    /* malformed token stream / model */

const char* cds_last_error(void);
int cds_version(void);
/* 1 when a compute-capability 10.x device is present and usable. */
int cds_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * (b) centre-surround contrast
 * ------------------------------------------------------------------------------------------ */
/* Drop-in for the computecontrast5 mex: pad is (m1 x n1) column-major double (the zero-padded,
 * pm_norm'ed map), W is (win x win) column-major, out is (m x n) column-major.
 * out(r,c) = sum_{i,j<win} W(i,j) * g(pad(r+i, c+j), pad(r+len, c+len)),  g(p,ctr) = p != 0 ? |p-ctr| : 0.
 * Arithmetic is FP32 on the device. */
int cds_contrast5(const double* pad, int m1, int n1, const double* W, int len, int win,
                  int m, int n, double* out);
/* Sudoku_contrastcpp5(subsaliencymap, delta, cusize, patchsize): map is rows x cols ROW-major. */
int cds_sudoku_contrast(const double* map, int rows, int cols, double delta, double cusize,
                        double patchsize, double* out);
/* cu_contrast5(salmap, cusize, delta, patchsize): rows x cols ROW-major. */
int cds_cu_contrast5(const double* map, int rows, int cols, int cusize, double delta,
                     double patchsize, double* out);
/* Device-resident batched core: nmaps padded FP32 maps [nmaps][pr][pc] (row-major, pr = rows+2*len,
 * pc = cols+2*len) -> raw contrast [nmaps][rows][cols] and per-map max (float bits, must be zeroed
 * by the caller or pass NULL).  Separable Gaussian weights exp(-d^2/(2 delta^2)) centred at `len`. */
int cds_contrast_sep_dev(const float* pad, float* out, unsigned* out_max_bits, int nmaps, int rows,
                         int cols, int win, float delta, void* stream);

/* ------------------------------------------------------------------------------------------
 * (a) syntax tokens -> 4x4 maps
 * ------------------------------------------------------------------------------------------ */
/* Packed wire format produced by the HM hooks for one picture:
 *   hdr[nctu][4]  int32 : {token offset (in int16 units, relative to `tok`), n_cupu, n_pred, n_mv}
 *   tok[]         int16 : per CTU  [cupu tokens | pred tokens | mv tokens]  (no 999 terminators)
 *   ctu_bytes[nctu] uint16 : CABAC bytes consumed per CTU (TDecSlice.cpp:333-341)
 * CTUs are in raster order, nctu = ceil(h/64)*ceil(w/64). */
int cds_getframe(int w, int h, const int32_t* hdr, const int16_t* tok, size_t ntok,
                 int16_t* mvx, int16_t* mvy, uint8_t* depth);
/* Convert the reference's static arrays (float[nctu][1000], 999-terminated) to the packed format.
 * hdr must hold nctu*4 ints, tok at least 3000*nctu int16.  Returns the token count (>= 0). */
long cds_syntax_pack_ref(int nctu, const float* MV, const float* Pred, const float* CUPU,
                         int32_t* hdr, int16_t* tok);
/* Batched, device-resident: nframes pictures; hdr [nframes][nctu][4] with offsets relative to
 * tok + tok_base[f]. */
int cds_getframe_dev(int w, int h, int nframes, const int32_t* hdr, const int16_t* tok,
                     const int64_t* tok_base, int16_t* mvx, int16_t* mvy, uint8_t* depth, void* stream);

/* ------------------------------------------------------------------------------------------
 * (d) libsvm C-SVC / RBF decision values
 * ------------------------------------------------------------------------------------------ */
typedef struct cds_svm_model cds_svm_model;
/* Build from dense arrays (what matlab_matrix_to_model reads from the MATLAB struct):
 * SV row-major nsv x nfeat, coef[nsv] (= sv_coef(:,1)), rho (= rho(1)), gamma (= Parameters(4)),
 * label0/label1 (= Label(1:2)).  Only svm_type 0 (C-SVC), kernel_type 2 (RBF), nr_class 2. */
int cds_svm_model_create(int nsv, int nfeat, const double* SV, const double* coef, double rho,
                         double gamma, int label0, int label1, cds_svm_model** out);
/* Load a libsvm text model (svm.cpp:2756 svm_save_model format). */
int cds_svm_model_load(const char* path, cds_svm_model** out);
void cds_svm_model_free(cds_svm_model* m);
int cds_svm_model_nsv(const cds_svm_model* m);
int cds_svm_model_nfeat(const cds_svm_model* m);
/* libsvmpredict core: X is N x F COLUMN-major (MATLAB), dec[N], label[N] (label may be NULL). */
int cds_svm_rbf_predict(const cds_svm_model* m, const double* X, int N, int F, double* dec, double* label);
/* Device-resident: X row-major [N][F] FP32 (already z-scored), dec[N] FP32. */
int cds_svm_rbf_predict_dev(const cds_svm_model* m, const float* X, long N, float* dec, void* stream);

/* ------------------------------------------------------------------------------------------
 * (c) camera-motion removal
 * ------------------------------------------------------------------------------------------ */
/* One picture.  Inputs: stage-(a) maps + bytes per CTU.  Outputs (h/4 x w/4 row-major FP32): Vx, Vy
 * (pixels; camera motion removed when the camera moves), mv_norm, bit_norm, depth_norm;
 * cam[2] = {x_cammov, y_cammov} (exact), ave[2] = {x_ave, y_ave} (exact doubles), flag. */
int cds_camera_motion(int w, int h, const int16_t* mvx, const int16_t* mvy, const uint16_t* ctu_bytes,
                      const uint8_t* depth, float* Vx, float* Vy, float* mv_norm, float* bit_norm,
                      float* depth_norm, int* cam, double* ave, int* flag);
/* mv_vote(eVx, eVy) = direction_cluster(.,.,16): n vectors in 1/8-pixel integer units (the exact
 * representation of medfilt2'ed quarter-pel MVs / 4). */
int cds_mv_vote(const int32_t* vx8, const int32_t* vy8, int n, double* x_ave, double* y_ave);
/* magnitude_vote(value) */
int cds_magnitude_vote(const double* value, int n, double* v_ave);

/* ------------------------------------------------------------------------------------------
 * whole path: per-clip context (main.m:113-305 with the RBF-SVM branch; TDecGop.cpp:225-333 hook)
 * ------------------------------------------------------------------------------------------ */
typedef struct cds_ctx cds_ctx;
#define CDS_FUSION_LINEAR 0
#define CDS_FUSION_RBF 1
#define CDS_FUSION_FAST 2
typedef struct {
  int w, h;            /* luma size; multiples of 8 */
  double fps;          /* PreframSmooth = round(0.3 fps), PreframTemporal = round(7 PS) (main.m:119,129) */
  int use_rbf;         /* CDS_FUSION_RBF 1: RBF-SVM fusion (main.m:273-301); CDS_FUSION_LINEAR 0: linear fusion (main.m:258-272);
                        * CDS_FUSION_FAST 2: the FAST version's arithmetic (TDecGop.cpp:225-333, 486-656 + salmat.cpp: float32,
                        * CameraDect / Umean vote, PS = cvRound(0.3 fps), PT = cvRound(0.7 fps), coarse-grid contrast, fixed feature
                        * weights, cv::GaussianBlur, 8-bit map) — reproduces the decoder's ../HMdata/<POC>.png; width % 16 == 0 */
  const cds_svm_model* model;   /* stand-in for svmRBFmodel_all.mat:model (required when use_rbf) */
  double mean_vec[9], std_vec[9];   /* svmRBFmodel_all.mat: meanVec, stdVec */
  int max_batch;       /* pictures processed per launch group (0 = default 32) */
  int out_u8;          /* 1: output maps as uint8 (x255, like writeVideo / imwrite), 0: FP32 */
} cds_params;

int cds_create(const cds_params* p, cds_ctx** out);
void cds_destroy(cds_ctx* ctx);
/* Hook 2 (TDecCu::decompressCU / MVoutput, TDecCu.cpp:153-161, 165-289): PINNED host buffers for the next picture's packed
 * syntax, owned by the context — ctu_bytes[nctu], hdr[nctu][4], tok[*tok_capacity] (room for a worst-case picture, 1300 tokens
 * per CTU).  The decoder thread appends its int16 tokens and per-CTU counts there directly (no 999-terminated float[1000] rows,
 * no conversion pass) and then calls cds_submit_frame with these same pointers, which makes the hand-off zero-copy.  The pointers
 * are valid until that cds_submit_frame.  May close a batch early (and launch it) when the token pool is nearly full. */
int cds_acquire_frame_buffers(cds_ctx* ctx, uint16_t** ctu_bytes, int32_t** hdr, int16_t** tok, size_t* tok_capacity);
/* Hook 3 (the per-slice tail of TDecGop::decompressSlice, TDecGop.cpp:270-332): hand one picture's packed syntax to the GPU.
 * Pictures must arrive in POC order (starting at 0 or at the cds_seek position).  With the pointers of
 * cds_acquire_frame_buffers nothing is copied; any other host memory is copied once into the pinned staging batch.
 * ASYNCHRONOUS: the call only records the picture; when a batch of max_batch pictures is complete it enqueues the host->device
 * copies, the kernel chain and the device->host copy of the maps on the context's streams and returns without waiting for the
 * GPU (the decoder thread goes on parsing the next pictures while the batch runs). */
int cds_submit_frame(cds_ctx* ctx, int poc, const uint16_t* ctu_bytes, const int32_t* hdr,
                     const int16_t* tok, size_t ntok);
/* Waits until picture `poc` is finished (only for its own batch) and copies its saliency map (h x w, row-major; FP32 or u8
 * according to out_u8) and camera motion (may be NULL) out of the context's pinned result buffer.  A still-staged partial batch
 * that contains `poc` is launched first.  The maps of a batch stay retrievable until two further batches have been launched.
 * A malformed token stream in the batch is reported here (CDS_EFORMAT). */
int cds_get_map(cds_ctx* ctx, int poc, void* dst, int* cam_xy);
/* Launch any partially filled batch and every deferred stage; does not wait for the GPU (cds_get_map does). */
int cds_flush(cds_ctx* ctx);
/* Batch entry points used by bench.py.  Process `nframes` pictures [first .. first+nframes) whose
 * packed syntax is resident on the DEVICE (dev pointers; layout as cds_getframe_dev, ctu_bytes
 * [nframes][nctu]) and write maps to a device buffer [nframes][h][w]; cam_out (may be NULL) is a
 * DEVICE buffer [nframes][2].  `stream` is a cudaStream_t (NULL = the context's own stream). */
int cds_clip_process_dev(cds_ctx* ctx, int first_poc, int nframes, const uint16_t* ctu_bytes,
                         const int32_t* hdr, const int16_t* tok, const int64_t* tok_base,
                         void* maps_out, int* cam_out, void* stream);
/* Same with HOST buffers (pinned or pageable): copies in, computes, copies maps out (maps_out may be
 * NULL: the last batch then stays retrievable through cds_get_map). */
int cds_clip_process(cds_ctx* ctx, int first_poc, int nframes, const uint16_t* ctu_bytes,
                     const int32_t* hdr, const int16_t* tok, const int64_t* tok_base, size_t ntok,
                     void* maps_out, int* cam_out);
/* The same clip step from stage-(a) MAPS instead of syntax — the reference's normal version exchanges exactly these between its
 * decoder and MATLAB (main.m:25-27,38-55: mv_x / mv_y / depth h/4 x w/4 per picture and the bytes-per-CTU bitmap, as text files or
 * .mat cell arrays; formats.py reads both).  HOST buffers: mvx, mvy [nframes][h/4*w/4] quarter-pel int16, depth u8 (0..3),
 * ctu_bytes [nframes][nctu] u16.  Works in every fusion mode; maps_out NULL feeds the pictures as temporal context only. */
int cds_clip_process_maps(cds_ctx* ctx, int first_poc, int nframes, const int16_t* mvx, const int16_t* mvy,
                          const uint8_t* depth, const uint16_t* ctu_bytes, void* maps_out, int* cam_out);
/* "Semantic" stage (a): the CTU's 256 z-scan partitions as TComDataCU holds them — depth256 / pred256
 * [nframes][nctu][256] u8 (getDepth(), getPredictionMode(): 0 inter, 1 intra), mv256 [nframes][nctu][256][2] int16 quarter-pel
 * (getCUMvField(L0 or L1)->getMv()) — transposed to the h/4 x w/4 maps of cds_getframe.  Depth equals getFrame's; the motion maps
 * differ where getFrame's stream-aliasing quirks fire (an intentional deviation).  Feed the result to cds_clip_process_maps. */
int cds_getframe_semantic(int w, int h, int nframes, const uint8_t* depth256, const uint8_t* pred256, const int16_t* mv256,
                          int16_t* mvx, int16_t* mvy, uint8_t* depth);
int cds_getframe_semantic_dev(int w, int h, int nframes, const uint8_t* depth256, const uint8_t* pred256, const int16_t* mv256,
                              int16_t* mvx, int16_t* mvy, uint8_t* depth, void* stream);
/* Probes of the fast version's sequential-float-sum emulation (csrc/fast.cu): the float32 result of
 * `S = 0; for (i < n) if (sel[i]) S = S + v[i];` (salmat.cpp:155-162, TDecGop.cpp:622-640; n a multiple of 4), and of the same loop
 * over a rows x cols raster of replicated blocks (gh x gw values, block size ms, elements > t selected).  Host buffers. */
int cds_probe_seq_sum(const float* v, const unsigned char* sel, int n, float* sum, int* count);
int cds_probe_run_sum(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, float* sum, int* count);
/* Drop all temporal state (start a new clip at POC 0). */
int cds_reset(cds_ctx* ctx);
/* Drop all temporal state and continue at picture `poc` (frame-range sharding: seek to
 * first - (PS + PT - 2), feed that halo through cds_clip_context_dev, then process the owned range). */
int cds_seek(cds_ctx* ctx, int poc);
/* Run only the stages later pictures depend on (syntax->maps, camera motion, smoothing) for `nframes`
 * pictures of temporal context; no saliency maps are produced.  Device pointers, as cds_clip_process_dev. */
int cds_clip_context_dev(cds_ctx* ctx, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                         const int16_t* tok, const int64_t* tok_base, void* stream);
/* Same with HOST buffers (as cds_clip_process). */
int cds_clip_context(cds_ctx* ctx, int first_poc, int nframes, const uint16_t* ctu_bytes, const int32_t* hdr,
                     const int16_t* tok, const int64_t* tok_base, size_t ntok);
/* Test hook: copy an intermediate of the most recent batch to host memory.  which = "xmaps" ([9][h/4*w/4]),
 * "stats" ([9][4]), "feat200" ([9][40000]), "ft" ([3][40000]), "xf" ([3][4]), "xsvm" ([40000][9]), "dec" ([40000]),
 * "raw" ([h][w]); all FP32. */
int cds_debug_copy(cds_ctx* ctx, const char* which, int frame_in_batch, void* dst, size_t bytes);
/* Per-stage device timing of the clip pipeline: CUDA events recorded on the launching stream around
 * each stage of every batch.  cds_profile_stage_names() is a comma-separated list; cds_profile_read
 * synchronises and returns accumulated milliseconds / interval counts per stage (same order). */
const char* cds_profile_stage_names(void);
int cds_profile_enable(cds_ctx* ctx, int on);
int cds_profile_read(cds_ctx* ctx, double* ms, long* calls, int n);
/* Measure this device's issue ceilings with micro-kernels (lane-instructions per second): the
 * contrast kernel's FADD+FFMA pair mix, pure FFMA, and MUFU.EX2.  Used as roofline denominators
 * for the FP32/MUFU-bound kernels (MEASURED_PEAKS.json has only HBM and bf16 tensor numbers). */
int cds_measure_peaks(double* fp32_pair_ips, double* ffma_ips, double* mufu_ips);
/* The same plus a fourth micro-kernel: vals[0..n) = {FADD -> FFMA pair mix (the contrast pair), pure FFMA, MUFU.EX2, FADD and FFMA
 * streams WITHOUT a dependence between them}, n <= 4.  The fourth separates "mixing two opcodes on the FMA pipe" from "every FFMA
 * waits for its own FADD" as the reason why the pair mix issues below the pure-FFMA rate. */
int cds_measure_peaks_ex(double* vals, int n);
/* Development probes (tools/mix_probe.py, tools/umma_probe.py): instruction-mix throughput and one tcgen05.mma with
 * caller-supplied shared-memory operand images / descriptor strides.  Not part of the drop-in surface. */
int cds_probe_mix(int which, int ctas_per_sm, double* fma_lane_ops_per_s, double* mufu_per_s);
/* Asynchronous spinner on `stream` for co-residency experiments (tools/corun_probe.py): which = 0 MUFU.EX2, 1 FFMA, 2 FFMA2,
 * 3 MUFU.EX2 + FADD; `ctas_per_sm` CTAs of `threads` (32..256) threads per SM, `iters` rounds of 128 operations per thread. */
int cds_probe_spin(int which, int ctas_per_sm, int threads, int iters, void* stream);
int cds_probe_umma(const float* a_img, int a_bytes, const float* b_img, int b_bytes, unsigned a_lbo, unsigned a_sbo,
                   unsigned b_lbo, unsigned b_sbo, int a_mn_major, int b_mn_major, float* D);
/* Number of kernel launches issued by this library since the last call (bench bookkeeping). */
long cds_take_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
