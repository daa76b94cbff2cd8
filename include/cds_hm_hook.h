/* cds_hm_hook.h — reference-side binding of libcds_b200.so for the HM_16.0_features decoder.
 *
 * Header-only C++ (no CUDA, no link-time dependency: the library is dlopen'ed), meant to be included by the
 * reference's TLibDecoder/TDecGop.cpp in place of its OpenCV saliency block (TDecGop.cpp:47-118 globals,
 * :225-333 per-slice block) and by TLibDecoder/TDecCu.cpp for the capture hook.
 *
 * Capture (hook 2, TDecCu::decompressCU / MVoutput, TDecCu.cpp:153-289): instead of the float[LCUNUM][1000] statics
 * TDecCu::MV / CUPU / Pred with their 999 terminators (TDecCu.h:100-106), the walk appends int16 tokens to three small
 * per-CTU scratch arrays and CdsHmHook::onCtu() copies them, with their lengths and the CTU's byte count
 * (TDecSlice.cpp:333-341), straight into the PINNED staging buffers the library handed out (cds_acquire_frame_buffers).
 * Hand-off (hook 3, the tail of TDecGop::decompressSlice, TDecGop.cpp:270-332): CdsHmHook::onSliceDone() calls
 * cds_submit_frame with those same pointers (zero copy; asynchronous: the GPU works on batch k while HM parses batch k+1)
 * and drains the maps of the batch BEFORE the one just launched (cds_get_map waits for that batch only) into <CDS_OUT>
 * (raw maps, POC order); the tail is drained when the process exits.  Two clocks keep HM's CABAC parse / reconstruction
 * (the host producer) apart from the saliency hand-off + GPU wait, as the north star asks.
 * CdsHmHook::onSlice() is the older entry that converts the reference's statics (cds_syntax_pack_ref) — kept for decoders that
 * leave TDecCu.cpp untouched.
 *
 * Environment:
 *   CDS_LIB        path of libcds_b200.so                       (required)
 *   CDS_OUT        output file for the maps                     (default: none, maps are dropped after the copy)
 *   CDS_FPS        frame rate, sets PS / PT (main.m:119,129)    (default 50)
 *   CDS_SVM_MODEL  libsvm text model with 9 features            (default: none -> linear fusion, main.m:258-272)
 *   CDS_MEAN, CDS_STD  nine comma-separated values (meanVec / stdVec of svmRBFmodel_all.mat; default 0 / 1)
 *   CDS_BATCH      pictures per GPU batch                       (default 32)
 *   CDS_U8         1: 8-bit maps like the reference's imwrite (TDecGop.cpp:327-332), 0: FP32 (default 0; 1 with CDS_MODE=fast)
 *   CDS_MODE       "fast": the fast version's arithmetic, i.e. what this decoder's removed OpenCV block computed
 *                  (TDecGop.cpp:225-333 + salmat.cpp); no SVM model needed; the 8-bit maps equal its ../HMdata/<POC>.png
 */
#ifndef CDS_HM_HOOK_H
#define CDS_HM_HOOK_H
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>
#include "cds.h"

class CdsHmHook {
 public:
  /* Hook 2, once per CTU from TDecCu::decompressCU: the three token streams of CTU `addr` (int16, no terminators) and its bytes. */
  void onCtu(int width, int height, int nctu, int addr, int bytes, const short* cupu, int ncupu, const short* pred, int npred, const short* mv, int nmv) {
    const double t0 = now();
    if (!open(width, height, nctu)) return;
    if (!acq_) {
      if (!acquire_ || acquire_(ctx_, &a_bytes_, &a_hdr_, &a_tok_, &a_cap_) != CDS_OK) { fail("cds_acquire_frame_buffers"); return; }
      acq_ = true; ntok_ = 0;
    }
    if (addr < 0 || addr >= nctu || ntok_ + (size_t)(ncupu + npred + nmv) > a_cap_) { fail("token buffer overflow"); return; }
    int32_t* h = a_hdr_ + 4 * (size_t)addr;
    h[0] = (int32_t)ntok_; h[1] = ncupu; h[2] = npred; h[3] = nmv;
    memcpy(a_tok_ + ntok_, cupu, 2 * (size_t)ncupu); ntok_ += ncupu;
    memcpy(a_tok_ + ntok_, pred, 2 * (size_t)npred); ntok_ += npred;
    memcpy(a_tok_ + ntok_, mv, 2 * (size_t)nmv); ntok_ += nmv;
    a_bytes_[addr] = (uint16_t)(bytes < 0 ? 0 : bytes > 65535 ? 65535 : bytes);
    t_hook_ += now() - t0;
    if (t_first_ == 0) t_first_ = t0;
  }
  /* Hook 3, once per slice NAL (one slice per picture): submit the picture captured by onCtu and drain finished maps. */
  void onSliceDone(int poc) {
    const double t0 = now();
    if (!ctx_ || failed_ || !acq_) return;
    acq_ = false;
    if (submit_(ctx_, poc, a_bytes_, a_hdr_, a_tok_, ntok_) != CDS_OK) { fail("cds_submit_frame"); return; }
    submitted_ = poc + 1;
    /* the batch launched last is still running: collect the one before it (already finished or nearly so) */
    if (submitted_ - drained_ >= 2 * batch_) drain(drained_ + batch_);
    t_hook_ += now() - t0;
  }
  /* Older entry: MV / Pred / CUPU are the reference's float[nctu][1000] statics (999-terminated); bits: float[nctu]. */
  void onSlice(int poc, int width, int height, int nctu, const float* bits, const float* MV, const float* Pred, const float* CUPU) {
    const double t0 = now();
    if (!open(width, height, nctu)) return;
    hdr_.resize((size_t)nctu * 4); tok_.resize((size_t)nctu * 3000); bytes_.resize(nctu);
    const long ntok = pack_(nctu, MV, Pred, CUPU, hdr_.data(), tok_.data());
    if (ntok < 0) { fail("cds_syntax_pack_ref"); return; }
    for (int a = 0; a < nctu; a++) bytes_[a] = (uint16_t)bits[a];
    if (submit_(ctx_, poc, bytes_.data(), hdr_.data(), tok_.data(), (size_t)ntok) != CDS_OK) { fail("cds_submit_frame"); return; }
    submitted_ = poc + 1;
    if (submitted_ - drained_ >= 2 * batch_) drain(drained_ + batch_);
    t_hook_ += now() - t0;
    if (t_first_ == 0) t_first_ = t0;
  }
  /* Drain the tail and report.  Runs from atexit(), registered AFTER the library created its CUDA context, so it runs
   * BEFORE the CUDA runtime's own exit handlers (a static destructor would run after them and every call would fail). */
  void finish() {
    if (!ctx_) return;
    const double t0 = now();
    drain(submitted_);                                             /* cds_get_map flushes the partial batch */
    t_hook_ += now() - t0;
    const double wall = now() - t_first_;
    const double hand = t_hook_ - t_init_;                         /* one-off: dlopen, CUDA context, model, cds_create */
    fprintf(stderr, "[cds hook] %d pictures: total %.3f s, one-off start-up %.3f s, saliency hand-off + GPU wait %.3f s (%.3f ms/picture, %.1f %% of the wall time), "
            "HM parse + reconstruction %.3f s (%.3f ms/picture)\n", submitted_, wall, t_init_, hand, 1e3 * hand / (submitted_ ? submitted_ : 1),
            100.0 * hand / (wall > 0 ? wall : 1), wall - t_hook_, 1e3 * (wall - t_hook_) / (submitted_ ? submitted_ : 1));
    destroy_(ctx_); ctx_ = NULL;
    if (model_) model_free_(model_);
    if (out_) fclose(out_);
  }
  static CdsHmHook*& instance() { static CdsHmHook* p = NULL; return p; }
  static CdsHmHook& get() { if (!instance()) instance() = new CdsHmHook(); return *instance(); }

 private:
  typedef long (*pack_fn)(int, const float*, const float*, const float*, int32_t*, int16_t*);
  typedef int (*create_fn)(const cds_params*, cds_ctx**);
  typedef void (*destroy_fn)(cds_ctx*);
  typedef int (*submit_fn)(cds_ctx*, int, const uint16_t*, const int32_t*, const int16_t*, size_t);
  typedef int (*acquire_fn)(cds_ctx*, uint16_t**, int32_t**, int16_t**, size_t*);
  typedef int (*getmap_fn)(cds_ctx*, int, void*, int*);
  typedef int (*load_fn)(const char*, cds_svm_model**);
  typedef void (*mfree_fn)(cds_svm_model*);
  typedef const char* (*err_fn)(void);

  static void atExit() { if (instance()) instance()->finish(); }
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
  void fail(const char* what) { if (!failed_) fprintf(stderr, "[cds hook] %s failed: %s\n", what, err_ ? err_() : "?"); failed_ = true; }
  static void parse9(const char* env, double* dst, double dflt) {
    for (int i = 0; i < 9; i++) dst[i] = dflt;
    const char* s = getenv(env);
    for (int i = 0; s && i < 9; i++) { char* end; dst[i] = strtod(s, &end); if (end == s) { dst[i] = dflt; break; } s = (*end == ',') ? end + 1 : end; }
  }
  bool open(int w, int h, int nctu) {
    if (ctx_) return !failed_;
    if (failed_) return false;
    const double t_open0 = now();
    struct InitClock { double t0; double* acc; ~InitClock() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); *acc += ts.tv_sec + 1e-9 * ts.tv_nsec - t0; } } ic = {t_open0, &t_init_};
    const char* lib = getenv("CDS_LIB");
    void* so = lib ? dlopen(lib, RTLD_NOW | RTLD_LOCAL) : NULL;
    if (!so) { fprintf(stderr, "[cds hook] cannot load CDS_LIB (%s): %s\n", lib ? lib : "unset", dlerror()); failed_ = true; return false; }
    pack_ = (pack_fn)dlsym(so, "cds_syntax_pack_ref"); create_ = (create_fn)dlsym(so, "cds_create"); destroy_ = (destroy_fn)dlsym(so, "cds_destroy");
    submit_ = (submit_fn)dlsym(so, "cds_submit_frame"); getmap_ = (getmap_fn)dlsym(so, "cds_get_map"); load_ = (load_fn)dlsym(so, "cds_svm_model_load");
    model_free_ = (mfree_fn)dlsym(so, "cds_svm_model_free"); err_ = (err_fn)dlsym(so, "cds_last_error");
    acquire_ = (acquire_fn)dlsym(so, "cds_acquire_frame_buffers");
    if (!pack_ || !create_ || !destroy_ || !submit_ || !getmap_ || !load_ || !model_free_ || !err_) { fprintf(stderr, "[cds hook] missing symbols in %s\n", lib); failed_ = true; return false; }
    cds_params p; memset(&p, 0, sizeof(p));
    p.w = w; p.h = h; p.fps = getenv("CDS_FPS") ? atof(getenv("CDS_FPS")) : 50.0;
    parse9("CDS_MEAN", p.mean_vec, 0.0); parse9("CDS_STD", p.std_vec, 1.0);
    if (getenv("CDS_SVM_MODEL")) {
      if (load_(getenv("CDS_SVM_MODEL"), &model_) != CDS_OK) { fail("cds_svm_model_load"); return false; }
      p.use_rbf = 1; p.model = model_;
    }
    const bool fast = getenv("CDS_MODE") && !strcmp(getenv("CDS_MODE"), "fast");
    if (fast) p.use_rbf = CDS_FUSION_FAST;     // the arithmetic this decoder's own OpenCV block had (TDecGop.cpp:225-333, salmat.cpp)
    batch_ = getenv("CDS_BATCH") ? atoi(getenv("CDS_BATCH")) : 32;
    p.max_batch = batch_; p.out_u8 = getenv("CDS_U8") ? atoi(getenv("CDS_U8")) : (fast ? 1 : 0);
    if (create_(&p, &ctx_) != CDS_OK) { fail("cds_create"); ctx_ = NULL; return false; }
    px_ = (size_t)w * h * (p.out_u8 ? 1 : 4);
    map_.resize(px_);
    if (getenv("CDS_OUT")) out_ = fopen(getenv("CDS_OUT"), "wb");
    (void)nctu;
    atexit(&CdsHmHook::atExit);
    return true;
  }
  void drain(int upto) {
    for (; drained_ < upto && !failed_; drained_++) {
      int cam[2];
      if (getmap_(ctx_, drained_, map_.data(), cam) != CDS_OK) { fail("cds_get_map"); return; }
      if (out_) fwrite(map_.data(), 1, px_, out_);
    }
  }

  cds_ctx* ctx_ = NULL; cds_svm_model* model_ = NULL; FILE* out_ = NULL;
  pack_fn pack_ = NULL; create_fn create_ = NULL; destroy_fn destroy_ = NULL; submit_fn submit_ = NULL; getmap_fn getmap_ = NULL;
  load_fn load_ = NULL; mfree_fn model_free_ = NULL; err_fn err_ = NULL; acquire_fn acquire_ = NULL;
  uint16_t* a_bytes_ = NULL; int32_t* a_hdr_ = NULL; int16_t* a_tok_ = NULL; size_t a_cap_ = 0, ntok_ = 0; bool acq_ = false;
  std::vector<int32_t> hdr_; std::vector<int16_t> tok_; std::vector<uint16_t> bytes_; std::vector<char> map_;
  size_t px_ = 0; int batch_ = 32, submitted_ = 0, drained_ = 0; bool failed_ = false;
  double t_hook_ = 0, t_first_ = 0, t_init_ = 0;
};
#endif
