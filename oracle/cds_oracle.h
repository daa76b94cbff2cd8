/* cds_oracle.h — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the timed CPU baseline — never as the product.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   stage (a) orc_getframe      : pinned — bit-exact vs the authors' golden .mat maps
 *                                 (tests/golden/stage_a_*.npz) and vs the verbatim reference
 *                                 getFrame() compiled from /root/reference (oracle/_ref).
 *   stage (b) orc_computecontrast5: pinned vs the verbatim reference computecontrast5.cpp
 *                                 (oracle/_ref/libref_contrast.so); no golden vectors exist.
 *   stage (d) orc_svm_*         : pinned vs the reference's vendored libsvm-3.17
 *                                 (oracle/_ref/libref_svm.so, heart_scale known answer).
 *   stages (c),(e), MATLAB glue : PARITY UNPINNED — restated from the .m text; the reference
 *                                 ships no outputs for them and MATLAB is not available.
 *   fast version orc_fast_*     : pinned — the reference's 39 golden PNGs (bin/HMdata/0..38.png),
 *                                 see cds_oracle_fast.c and tests/test_oracle_fast_cpu.py.
 *
 * Conventions: maps are row-major [rows][cols] doubles unless a function says "col-major"
 * (the two mex drop-ins keep MATLAB's column-major layout).
 */
#ifndef CDS_ORACLE_H
#define CDS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------- stage (a): syntax tokens -> 4x4 maps (code_mat_h.h:575 getFrame) ---------- */
/* hdr[nctu][4] = {token offset, n_cupu, n_pred, n_mv}; tok = int16 stream [cupu|pred|mv] per CTU.
 * Outputs h/4 x w/4 row-major.  Returns 0, or <0 on malformed input. */
int orc_getframe(int w, int h, const int32_t* hdr, const int16_t* tok,
                 int16_t* mvx, int16_t* mvy, uint8_t* depth);

/* ---------- stage (b): centre-surround contrast ---------- */
/* computecontrast5.cpp:52-96 — col-major, exactly the mex arguments. */
void orc_computecontrast5(const double* pad, int m1, int n1, const double* W, int len, int win,
                          int m, int n, double* out);
/* Sudoku_contrastcpp5.m:1-52 on a row-major rows x cols map. */
void orc_sudoku_contrast(const double* map, int rows, int cols, double delta, double cusize,
                         double patchsize, double* out);
/* cu_contrast5.m:3-63 on a row-major rows x cols map. */
void orc_cu_contrast5(const double* map, int rows, int cols, int cusize, double delta,
                      double patchsize, double* out);

/* ---------- stage (d): libsvm C-SVC / RBF decision values (svm.cpp:2501-2575, 325-365) ---------- */
/* SV row-major nsv x nfeat; X col-major N x F (svmpredict.c:165-170); label0/1 = model.Label. */
void orc_svm_rbf_predict(int nsv, int nfeat, const double* SV, const double* coef, double rho,
                         double gamma, int label0, int label1, const double* X, int N,
                         double* dec, double* label);

/* ---------- stage (c): camera motion (main.m:77-105, cm_judge.m, mv_vote.m, direction_cluster.m) */
/* mvx,mvy: h4 x w4 quarter-pel ints; bitmap: ceil(h/64) x ceil(w/64) bytes per CTU; depth h4 x w4.
 * Outputs (h4 x w4 doubles): Vx, Vy (pixels, camera motion removed when moving), mv_norm, bit_norm,
 * depth_norm; cam[2] = {x_cammov, y_cammov}; ave[2] = {x_ave, y_ave}; flag = 1 if camera moving. */
void orc_camera_motion(int w, int h, const int16_t* mvx, const int16_t* mvy, const uint16_t* bitmap,
                       const uint8_t* depth, double* Vx, double* Vy, double* mv_norm,
                       double* bit_norm, double* depth_norm, int* cam, double* ave, int* flag);
/* direction_cluster.m:1-50 (k = 16).  Returns 0; n == 0 gives ave = 0. */
void orc_direction_cluster(const double* vx, const double* vy, int n, double* x_ave, double* y_ave);
/* magnitude_vote.m:1-27 (dead code in the reference; kept as an optional op). */
double orc_magnitude_vote(const double* v, int n);
/* medfilt2(A,[2 2]) with MATLAB's even-window convention, zero padding. */
void orc_medfilt2_2x2(const double* a, int rows, int cols, double* out);

/* ---------- stage (e) helpers (pm_norm.m, mysterythrehold.m, immove.m, fourX2.m, upsample.m,
 *            getdistMatrix.m, imfilter+fspecial('gaussian'), imresize bilinear) ---------- */
void orc_pm_norm(double* a, long n);
void orc_mysterythrehold(double* a, long n, int im_height, int im_width);
void orc_immove(const double* im, int rows, int cols, int mov_x, int mov_y, double* out);
void orc_replicate(const double* a, int rows, int cols, int k, int out_rows, int out_cols, double* out);
void orc_distmatrix(int hei, int wid, double* out);
void orc_gauss_kernel(int k, double sigma, double* g /* k taps, normalised 1-D factor */);
void orc_imfilter_gauss(const double* a, int rows, int cols, int k, double sigma, double* out);
void orc_imresize_bilinear(const double* a, int rows, int cols, int out_rows, int out_cols, double* out);

/* ---------- whole clip, main.m:77-305 with linesvm = 0 (RBF branch) ---------- */
typedef struct {
  int w, h, nframes;
  double fps;
  /* stand-in for svmRBFmodel_all.mat (model, meanVec, stdVec) */
  int nsv, nfeat;
  const double* SV; const double* coef; double rho, gamma; int label0, label1;
  const double* meanVec; const double* stdVec;
  int use_rbf;              /* 0: linear fusion main.m:258-272, 1: RBF main.m:273-301 */
} orc_clip_params;
/* inputs: per-frame integer maps (stage a outputs).  out: nframes x h x w saliency maps (row-major
 * doubles), cam: nframes x 2.  feat200 (optional, may be NULL): nframes x 9 x 200 x 200.
 * first_out..nframes-1 are written (earlier frames are processed as temporal context only). */
int orc_process_clip(const orc_clip_params* p, const int16_t* mvx, const int16_t* mvy,
                     const uint8_t* depth, const uint16_t* bitmap, int first_out,
                     double* out, int* cam, double* feat200);

/* CPU-baseline option: run the two heavy kernels through the reference's own code (oracle/_ref: ref_computecontrast5 =
 * the verbatim mexFunction, ref_svm_predict = libsvm-3.17) in bands / chunks on the thread pool.  NULLs restore the port. */
typedef void (*orc_ref_contrast_fn)(const double* pad, int m1, int n1, const double* W, int len, int win, int m, int n, double* out);
typedef void (*orc_ref_svm_fn)(void* model, const double* X, int N, int F, double* dec, double* label);
void orc_use_reference_kernels(orc_ref_contrast_fn contrast, orc_ref_svm_fn svm, void* svm_model);

/* timing of the last orc_process_clip (CPU-baseline bookkeeping): which = 0 -> seconds in pass 1 +
 * smoothing over all frames, 1 -> seconds in the per-output-frame work.  orc_threads() = worker threads in use. */
double orc_last_seconds(int which);
void orc_set_threads(int n);   /* pthread parallel-for width (default 1; results do not depend on it) */
int orc_threads(void);

/* ---------- the reference's FAST version (cds_oracle_fast.c): TDecGop.cpp:225-333,486-656 + salmat.cpp, float32 ----------
 * Pinned by the 39 golden PNGs HM_16.0_features/bin/HMdata/0..38.png (tests/golden/fast_BasketballDrill.npz). */
typedef struct orc_fast orc_fast;
orc_fast* orc_fast_create(int w, int h, double fps);          /* NULL unless w, h are positive multiples of 4 */
void orc_fast_destroy(orc_fast* o);
int orc_fast_PS(const orc_fast* o);
int orc_fast_PT(const orc_fast* o);
/* One picture (FrameIndex = pictures fed so far): stage-(a) maps h/4 x w/4, bytes per CTU.  out_u8: h x w (what imwrite
 * stores); out_f32 (may be NULL): the normalised map before x255; cam[2] (may be NULL); feat9 (may be NULL): 9 x h x w in the
 * order of FeatureWeight (TDecGop.cpp:319). */
int orc_fast_frame(orc_fast* o, const int16_t* mvx, const int16_t* mvy, const uint8_t* depth, const uint16_t* ctu_bytes,
                   uint8_t* out_u8, float* out_f32, int* cam, float* feat9);
/* cv::getGaussianKernel(n, sigma, CV_32F) and cv::GaussianBlur(CV_32F, BORDER_CONSTANT) as OpenCV 2.4.9 evaluates them */
void orc_cv_gaussian_kernel(int n, double sigma, float* k);
void orc_cv_gaussian_blur(const float* src, int rows, int cols, int ksize, double sigma, float* dst, float* tmp);
/* the sequential float loops of MapThreshold / Umean, plainly (checkers for cds_probe_seq_sum / cds_probe_run_sum) */
float orc_seq_sum_f32(const float* v, const unsigned char* sel, long n, int* count);
float orc_run_sum_f32(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, int* count);

#ifdef __cplusplus
}
#endif
#endif
