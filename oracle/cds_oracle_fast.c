/* cds_oracle_fast.c — CPU restatement of the reference's FAST version.  TEST INFRASTRUCTURE ONLY (see cds_oracle.h).
 *
 * Follows HM_16.0_features/source/Lib/TLibDecoder/TDecGop.cpp:225-333 (per-slice tail), :486-656 (CameraDect, vote_alg,
 * max, Umean) and salmat.cpp (ForwardFilter :39, MatrixDuplication :57, MapDownSample :93, MatrixNorm :119, MatrixNorm2
 * :130, MapThreshold :141, GenerateOriMap :181, GenerateTemporalMap :193, GenerateTemporalMap_mv :242,
 * GenerateSpatialMap :299, GenerateWeightMap :342, GenerateSpatialMap_mv :361, _mv2 :402, translateTransform :412).
 * The reference computes in float32 through OpenCV 2.4.9 (not under /root/reference); the arithmetic of the OpenCV calls it
 * makes is restated op by op, with the rounding points OpenCV 2.4.9 has:
 *   Mat / s            -> x * (float)(1.0 / s)                      (matop.cpp operator/ -> convertTo with alpha = 1./s)
 *   acc + M * s        -> fl(fl(M * (float)s) + acc)                 (scaleAdd, no fused multiply-add)
 *   A*a + B*b          -> fl(fl(A*a) + fl(B*b))                      (addWeighted), further terms as scaleAdd
 *   sum(), mean()      -> groups of 4 consecutive floats added in float, groups accumulated in double; mean = sum*(1./N)
 *   GaussianBlur 32F   -> float kernel of getGaussianKernel; row pass taps in order from zero, column pass centre tap then
 *                         symmetric pairs fl(s + fl(fl(a+b)*k)); BORDER_CONSTANT zeros
 *   convertTo(CV_8U)   -> round half to even, saturate
 * Pinned by the 39 golden PNGs HM_16.0_features/bin/HMdata/0..38.png (BasketballDrill 832x480, tests/golden/fast_*.npz):
 * see tests/test_oracle_fast_cpu.py for the measured agreement.  Sub-LSB choices above cannot be resolved by 8-bit goldens.
 * Everything is sequential and at the reference's own resolutions (full H x W feature maps); no attempt at speed. */
#include "cds_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_REF 3.1415926 /* TDecGop.cpp:59 */

typedef struct {
  float dS, dT;          /* DeltaSpatial, DeltaTemporal */
  int ssize, minsize;    /* SpatialSize, MinSize */
  int gh, gw;            /* ceil(H/MinSize) x ceil(W/MinSize) */
  float* cur;            /* ThisFrameValue h4 x w4 */
  float* smooth;         /* SmoothFrame[PS] */
  float* tring;          /* TemporalFrame[PT], gh x gw each */
  float* ori;            /* OriMap H x W */
  float* tmap;           /* TemporalMap H x W */
  float* smap;           /* SpatialMap H x W */
} salmat_t;

struct orc_fast {
  int w, h, h4, w4, lh, lw, PS, PT, frame, cap;
  salmat_t bit, dep, mvx, mvy, mvn;
  int *camx, *camy;
  float *fmvx, *fmvy, *fdep, *fbits, *bitmap, *mask, *mvnorm, *tmpS, *tmpL, *final_, *rowbuf, *mvT, *mvS;
};

static int cv_round(double v) { return (int)nearbyint(v); }      /* cvRound: SSE2 cvtsd2si, round half to even */

/* cv::sum for CV_32F (stat.cpp sum_<float,double>): 4 consecutive floats added in float, then into the double accumulator */
static double cv_sum(const float* a, long n) {
  double s = 0; long i = 0;
  for (; i + 4 <= n; i += 4) { float g = a[i] + a[i + 1]; g = g + a[i + 2]; g = g + a[i + 3]; s += g; }
  for (; i < n; i++) s += a[i];
  return s;
}
static double cv_mean(const float* a, long n) { return cv_sum(a, n) * (1.0 / (double)n); }
static void cv_minmax(const float* a, long n, double* mn, double* mx) {
  float lo = a[0], hi = a[0];
  for (long i = 1; i < n; i++) { if (a[i] < lo) lo = a[i]; if (a[i] > hi) hi = a[i]; }
  *mn = lo; *mx = hi;
}
static void scale_to(const float* a, long n, double s, float* d) { float f = (float)(1.0 / s); for (long i = 0; i < n; i++) d[i] = a[i] * f; }

/* salmat.cpp:119 */
static void mat_norm(const float* src, long n, float* dst) {
  double mn, mx; cv_minmax(src, n, &mn, &mx);
  if (mx > 0) scale_to(src, n, mx, dst); else if (dst != src) memcpy(dst, src, n * sizeof(float));
}
/* salmat.cpp:130 */
static void mat_norm2(float* a, long n) {
  double mn, mx; cv_minmax(a, n, &mn, &mx);
  float neg = (float)(-mn);
  for (long i = 0; i < n; i++) a[i] = a[i] + neg;
  cv_minmax(a, n, &mn, &mx);
  if (mx > 0) scale_to(a, n, mx, a);
}
/* salmat.cpp:57 */
static void dup_to(const float* src, int sr, int sc, float* dst, int dr, int dc, int k) {
  (void)sr;
  for (int i = 0; i < dr; i++) for (int j = 0; j < dc; j++) dst[(long)i * dc + j] = src[(long)(i / k) * sc + j / k];
}
/* salmat.cpp:93 */
static void down_to(const float* src, int sc, float* dst, int dr, int dc, int k) {
  for (int i = 0; i < dr; i++) for (int j = 0; j < dc; j++) dst[(long)i * dc + j] = src[(long)i * k * sc + (long)j * k];
}
/* salmat.cpp:141 (dst may alias src) */
static void map_threshold(const float* src, long n, float* dst) {
  double mean = cv_mean(src, n), mn, mx; cv_minmax(src, n, &mn, &mx);
  double t = mx - mean;
  if (t == 0 || t >= mx) { if (dst != src) memcpy(dst, src, n * sizeof(float)); return; }
  float sum = 0; int count = 0;
  for (long k = 0; k < n; k++) if (src[k] > t) { sum = sum + src[k]; count++; }
  t = sum / count;                                   /* float / int -> float, kept in a double */
  if (t == 0) { if (dst != src) memcpy(dst, src, n * sizeof(float)); return; }
  for (long k = 0; k < n; k++) dst[k] = (src[k] >= t) ? 1.0f : (float)(src[k] / t);
}
/* salmat.cpp:39 */
static void forward_filter(salmat_t* s, long n, int F, int PS) {
  memcpy(s->smooth + (long)(F % PS) * n, s->cur, n * sizeof(float));
  int cnt = F < PS ? F + 1 : PS;
  float f = (float)(1.0 / (double)cnt);
  for (long k = 0; k < n; k++) {
    float acc = 0;
    for (int i = 0; i < cnt; i++) acc = acc + s->smooth[(long)i * n + k] * f;
    s->cur[k] = acc;
  }
}
/* salmat.cpp:412: dst(x+dx, y+dy) = src(x, y), zero fill */
static void translate(const float* src, int rows, int cols, int dx, int dy, float* dst) {
  if (dx == 0 && dy == 0) { memcpy(dst, src, (long)rows * cols * sizeof(float)); return; }
  memset(dst, 0, (long)rows * cols * sizeof(float));
  if (abs(dx) < cols && abs(dy) < rows)
    for (int y = 0; y < rows - abs(dy); y++)
      for (int x = 0; x < cols - abs(dx); x++)
        dst[(long)(y + (dy <= 0 ? 0 : dy)) * cols + x + (dx <= 0 ? 0 : dx)] = src[(long)(y + (dy > 0 ? 0 : -dy)) * cols + x + (dx > 0 ? 0 : -dx)];
}

/* salmat.cpp:181 */
static void gen_ori(orc_fast* o, salmat_t* s) {
  long n4 = (long)o->h4 * o->w4, N = (long)o->h * o->w;
  map_threshold(s->cur, n4, o->tmpS);
  dup_to(o->tmpS, o->h4, o->w4, o->tmpL, o->h, o->w, 4);
  mat_norm(o->tmpL, N, s->ori);
}

/* salmat.cpp:193 (mv = 0) and :242 (mv = 1, sy = the y component's object) */
static void gen_temporal(orc_fast* o, salmat_t* s, salmat_t* sy, int F, float* out) {
  const int gh = s->gh, gw = s->gw, PT = o->PT, ms = s->minsize;
  const long g = (long)gh * gw, N = (long)o->h * o->w;
  float* Sx = (float*)malloc(g * sizeof(float)); float* Sy = (float*)malloc(g * sizeof(float));
  float* tx = (float*)malloc(g * sizeof(float)); float* ty = (float*)malloc(g * sizeof(float));
  float* acc = (float*)calloc(g, sizeof(float));
  down_to(s->cur, o->w4, Sx, gh, gw, ms / 4);
  memcpy(s->tring + (long)(F % PT) * g, Sx, g * sizeof(float));
  if (sy) { down_to(sy->cur, o->w4, Sy, gh, gw, ms / 4); memcpy(sy->tring + (long)(F % PT) * g, Sy, g * sizeof(float)); }
  int last = F < PT ? (sy ? F - 1 : F) : PT - 1;      /* :210 i<FrameIndex+1, :262 i<FrameIndex, :222/:278 i<PreFrameTemp */
  int sumx = 0, sumy = 0;
  for (int i = 1; i <= last; i++) {
    int slot = (F - i) % PT;
    sumx += o->camx[F - i]; sumy += o->camy[F - i];
    int dx = sumx / ms, dy = sumy / ms;                /* integer division, cvRound of an int */
    float wgt = (float)exp((double)(-i / s->dT));      /* exp(float) -> float */
    translate(s->tring + (long)slot * g, gh, gw, dx, dy, tx);
    if (sy) translate(sy->tring + (long)slot * g, gh, gw, dx, dy, ty);
    for (long k = 0; k < g; k++) {
      float d;
      if (sy) { float ax = tx[k] - Sx[k], ay = ty[k] - Sy[k]; float xx = ax * ax, yy = ay * ay; d = sqrtf(xx + yy); }
      else d = fabsf(tx[k] - Sx[k]);
      float p = d * wgt;
      acc[k] = p + acc[k];
    }
  }
  dup_to(acc, gh, gw, o->tmpL, o->h, o->w, ms);
  mat_norm(o->tmpL, N, out);
  map_threshold(out, N, out);
  free(Sx); free(Sy); free(tx); free(ty); free(acc);
}

/* salmat.cpp:299 (mv = 0) and :361 (mv = 1) */
static void gen_spatial(orc_fast* o, salmat_t* s, int mv) {
  const int ms4 = s->minsize / 4, W = s->ssize, len = W / 2;
  const int rn = (o->h4 + ms4 - 1) / ms4, cn = (o->w4 + ms4 - 1) / ms4;
  const int pr = rn + 2 * len, pc = cn + 2 * len;
  const long n4 = (long)o->h4 * o->w4, N = (long)o->h * o->w;
  float* S = (float*)malloc((long)rn * cn * sizeof(float));
  float* P = (float*)calloc((long)pr * pc, sizeof(float));
  float* out = (float*)malloc((long)rn * cn * sizeof(float));
  float wm[32 * 32], tmp[32 * 32];
  for (int i = 0; i < W; i++) for (int j = 0; j < W; j++) {                     /* salmat.cpp:342 */
    float d = (float)((i - len) * (i - len)) + (float)((j - len) * (j - len));
    float den = 2 * (s->dS * s->dS);
    wm[i * W + j] = (float)exp((double)(-d / den));
  }
  map_threshold(s->cur, n4, o->tmpS);
  down_to(o->tmpS, o->w4, S, rn, cn, ms4);
  for (int i = 0; i < rn; i++) memcpy(P + (long)(i + len) * pc + len, S + (long)i * cn, cn * sizeof(float));
  mat_norm2(P, (long)pr * pc);
  for (int i = 0; i < rn; i++) for (int j = 0; j < cn; j++) {
    float c = P[(long)(i + len) * pc + j + len];
    for (int u = 0; u < W; u++) for (int v = 0; v < W; v++) tmp[u * W + v] = fabsf(P[(long)(i + u) * pc + j + v] - c) * wm[u * W + v];
    out[(long)i * cn + j] = (float)cv_sum(tmp, W * W);
  }
  mat_norm(out, (long)rn * cn, out);
  dup_to(out, rn, cn, o->tmpL, o->h, o->w, s->minsize);
  if (mv) memcpy(s->smap, o->tmpL, N * sizeof(float)); else mat_norm(o->tmpL, N, s->smap);
  free(S); free(P); free(out);
}

/* TDecGop.cpp:616 */
static float umean(const float* a, int len) {
  float sum = 0, mean;
  for (int i = 0; i < len; i++) sum = sum + a[i];
  mean = sum / len;
  sum = 0;
  int flag = mean > 0;
  for (int i = 0; i < len; i++) sum = ((flag && a[i] < mean) || (!flag && a[i] > mean)) ? sum + a[i] : sum + mean;
  return sum / len;
}

/* TDecGop.cpp:486 CameraDect + :559 vote_alg.  mvx/mvy are divided by 4 in place (:504-505). */
static void camera_dect(orc_fast* o, int* mov_x, int* mov_y) {
  const long n4 = (long)o->h4 * o->w4;
  double mn, mx; int back = 0, stat = 0;
  for (long k = 0; k < n4; k++) { o->fmvx[k] = o->fmvx[k] * 0.25f; o->fmvy[k] = o->fmvy[k] * 0.25f; o->mask[k] = 0; }
  cv_minmax(o->bitmap, n4, &mn, &mx);
  for (long k = 0; k < n4; k++)
    if (o->bitmap[k] < mn + 20) {
      back++; o->mask[k] = 1;
      float t = o->fmvx[k] * o->fmvx[k] + o->fmvy[k] * o->fmvy[k];
      if (t <= 2) stat++;
    }
  *mov_x = 0; *mov_y = 0;
  if (stat <= back / 2) {
    int count[16] = {0};
    unsigned char* bin = (unsigned char*)malloc((size_t)o->h4 * (size_t)o->w4);
    for (long k = 0; k < n4; k++) if (o->mask[k] != 0) {
      int idx = (int)(((atan2f(o->fmvy[k], o->fmvx[k]) / PI_REF) + 1) * 8.0);
      if (idx == 16) idx--;
      bin[k] = (unsigned char)idx; count[idx]++;
    }
    int best = 0; for (int i = 1; i < 16; i++) if (count[i] > count[best]) best = i;          /* :600 first maximum */
    float* ax = (float*)malloc((count[best] + 1) * sizeof(float)); float* ay = (float*)malloc((count[best] + 1) * sizeof(float));
    int m = 0;
    for (long k = 0; k < n4; k++) if (o->mask[k] != 0 && bin[k] == best) { ax[m] = o->fmvx[k]; ay[m] = o->fmvy[k]; m++; }
    float vx = umean(ax, m), vy = umean(ay, m);
    float nx = (float)(-(double)vx), ny = (float)(-(double)vy);
    for (long k = 0; k < n4; k++) { o->fmvx[k] = o->fmvx[k] + nx; o->fmvy[k] = o->fmvy[k] + ny; }
    *mov_x = cv_round(vx); *mov_y = cv_round(vy);
    free(bin); free(ax); free(ay);
  }
  for (long k = 0; k < n4; k++) { float xx = o->fmvx[k] * o->fmvx[k], yy = o->fmvy[k] * o->fmvy[k]; o->mvnorm[k] = sqrtf(xx + yy); }
}

/* cv::GaussianBlur on CV_32F, BORDER_CONSTANT (smooth.cpp getGaussianKernel, filter.cpp RowFilter / SymmColumnFilter) */
void orc_cv_gaussian_kernel(int n, double sigma, float* k) {
  double sx = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8, s2 = -0.5 / (sx * sx), sum = 0;
  for (int i = 0; i < n; i++) { double x = i - (n - 1) * 0.5; k[i] = (float)exp(s2 * x * x); sum += k[i]; }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) k[i] = (float)(k[i] * sum);
}
void orc_cv_gaussian_blur(const float* src, int rows, int cols, int ksize, double sigma, float* dst, float* tmp) {
  float k[512]; const int half = ksize / 2;
  orc_cv_gaussian_kernel(ksize, sigma, k);
  for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
    float s = 0;
    for (int t = 0; t < ksize; t++) { int xx = x + t - half; float v = (xx >= 0 && xx < cols) ? src[(long)y * cols + xx] : 0.0f; s = s + v * k[t]; }
    tmp[(long)y * cols + x] = s;
  }
  for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
    float s = tmp[(long)y * cols + x] * k[half];
    for (int t = 1; t <= half; t++) {
      float a = (y + t < rows) ? tmp[(long)(y + t) * cols + x] : 0.0f, b = (y - t >= 0) ? tmp[(long)(y - t) * cols + x] : 0.0f;
      float p = (a + b) * k[half + t];
      s = s + p;
    }
    dst[(long)y * cols + x] = s;
  }
}

static void salmat_init(salmat_t* s, float dS, float dT, int ssize, int minsize, const orc_fast* o) {
  const long n4 = (long)o->h4 * o->w4, N = (long)o->h * o->w;
  s->dS = dS; s->dT = dT; s->ssize = ssize; s->minsize = minsize;
  s->gh = (o->h + minsize - 1) / minsize; s->gw = (o->w + minsize - 1) / minsize;
  s->cur = (float*)calloc(n4, sizeof(float));
  s->smooth = (float*)calloc(n4 * o->PS, sizeof(float));
  s->tring = (float*)calloc((long)s->gh * s->gw * o->PT, sizeof(float));
  s->ori = (float*)calloc(N, sizeof(float)); s->tmap = (float*)calloc(N, sizeof(float)); s->smap = (float*)calloc(N, sizeof(float));
}
static void salmat_free(salmat_t* s) { free(s->cur); free(s->smooth); free(s->tring); free(s->ori); free(s->tmap); free(s->smap); }

orc_fast* orc_fast_create(int w, int h, double fps) {
  if (w <= 0 || h <= 0 || (w % 4) || (h % 4)) return NULL;
  orc_fast* o = (orc_fast*)calloc(1, sizeof(orc_fast));
  o->w = w; o->h = h; o->h4 = h / 4; o->w4 = w / 4; o->lh = (h + 63) / 64; o->lw = (w + 63) / 64;
  o->PS = cv_round(fps * 0.3); o->PT = cv_round(fps * 0.7);                    /* salmat.cpp:18-19 */
  if (o->PS < 1) o->PS = 1;
  if (o->PT < 1) o->PT = 1;
  /* TDecGop.cpp:63-71,79-83 */
  salmat_init(&o->bit, 3, 46, 5, 64, o); salmat_init(&o->dep, 13, 46, 12, 32, o);
  salmat_init(&o->mvx, 11, 26, 12, 32, o); salmat_init(&o->mvy, 11, 26, 12, 32, o); salmat_init(&o->mvn, 11, 26, 12, 32, o);
  const long n4 = (long)o->h4 * o->w4, N = (long)w * h;
  o->fmvx = (float*)malloc(n4 * 4); o->fmvy = (float*)malloc(n4 * 4); o->fdep = (float*)malloc(n4 * 4);
  o->fbits = (float*)malloc((long)o->lh * o->lw * 4); o->bitmap = (float*)malloc(n4 * 4); o->mask = (float*)malloc(n4 * 4);
  o->mvnorm = (float*)malloc(n4 * 4); o->tmpS = (float*)malloc(n4 * 4); o->tmpL = (float*)malloc(N * 4);
  o->final_ = (float*)malloc(N * 4); o->rowbuf = (float*)malloc(N * 4); o->mvT = (float*)malloc(N * 4); o->mvS = (float*)malloc(N * 4);
  o->cap = 1024; o->camx = (int*)calloc(o->cap, sizeof(int)); o->camy = (int*)calloc(o->cap, sizeof(int));
  return o;
}
void orc_fast_destroy(orc_fast* o) {
  if (!o) return;
  salmat_free(&o->bit); salmat_free(&o->dep); salmat_free(&o->mvx); salmat_free(&o->mvy); salmat_free(&o->mvn);
  free(o->fmvx); free(o->fmvy); free(o->fdep); free(o->fbits); free(o->bitmap); free(o->mask); free(o->mvnorm); free(o->tmpS);
  free(o->tmpL); free(o->final_); free(o->rowbuf); free(o->mvT); free(o->mvS); free(o->camx); free(o->camy); free(o);
}
int orc_fast_PS(const orc_fast* o) { return o->PS; }
int orc_fast_PT(const orc_fast* o) { return o->PT; }

/* TDecGop.cpp:225-333 for picture FrameIndex = number of pictures fed so far */
int orc_fast_frame(orc_fast* o, const int16_t* mvx, const int16_t* mvy, const uint8_t* depth, const uint16_t* ctu_bytes,
                   uint8_t* out_u8, float* out_f32, int* cam, float* feat9) {
  const int F = o->frame, h4 = o->h4, w4 = o->w4;
  const long n4 = (long)h4 * w4, N = (long)o->h * o->w, nl = (long)o->lh * o->lw;
  if (F >= o->cap) { o->cap *= 2; o->camx = (int*)realloc(o->camx, o->cap * sizeof(int)); o->camy = (int*)realloc(o->camy, o->cap * sizeof(int)); }
  for (long k = 0; k < n4; k++) { o->fmvx[k] = mvx[k]; o->fmvy[k] = mvy[k]; o->fdep[k] = depth[k]; }
  for (long k = 0; k < nl; k++) o->fbits[k] = ctu_bytes[k];
  dup_to(o->fbits, o->lh, o->lw, o->bitmap, h4, w4, 16);                        /* :273-274 */
  camera_dect(o, &o->camx[F], &o->camy[F]);                                    /* :277 */
  mat_norm(o->fbits, nl, o->fbits);                                            /* :278 */
  dup_to(o->fbits, o->lh, o->lw, o->bit.cur, h4, w4, 16);                      /* :280 */
  forward_filter(&o->bit, n4, F, o->PS);
  mat_norm(o->fdep, n4, o->dep.cur); forward_filter(&o->dep, n4, F, o->PS);   /* :284-285 */
  memcpy(o->mvx.cur, o->fmvx, n4 * 4); forward_filter(&o->mvx, n4, F, o->PS);
  memcpy(o->mvy.cur, o->fmvy, n4 * 4); forward_filter(&o->mvy, n4, F, o->PS);
  gen_ori(o, &o->bit); gen_ori(o, &o->dep);                                    /* :292-293 */
  mat_norm(o->mvnorm, n4, o->mvn.cur); forward_filter(&o->mvn, n4, F, o->PS); gen_ori(o, &o->mvn);
  gen_temporal(o, &o->bit, NULL, F, o->bit.tmap); gen_temporal(o, &o->dep, NULL, F, o->dep.tmap);
  gen_temporal(o, &o->mvx, &o->mvy, F, o->mvT);
  gen_spatial(o, &o->bit, 0); gen_spatial(o, &o->dep, 0); gen_spatial(o, &o->mvx, 1); gen_spatial(o, &o->mvy, 1);
  for (long k = 0; k < N; k++) { float xx = o->mvx.smap[k] * o->mvx.smap[k], yy = o->mvy.smap[k] * o->mvy.smap[k]; o->tmpL[k] = sqrtf(xx + yy); }
  mat_norm(o->tmpL, N, o->mvS);                                                /* salmat.cpp:402 */
  static const float wgt[9] = {0.058f, 0.171f, -0.095f, 0.068f, -0.01f, 0.509f, -0.051f, 0.154f, 0.196f};   /* TDecGop.cpp:94 */
  const float* maps[9] = {o->mvn.ori, o->mvT, o->mvS, o->bit.ori, o->bit.tmap, o->bit.smap, o->dep.ori, o->dep.tmap, o->dep.smap};
  for (long k = 0; k < N; k++) {
    float a = maps[0][k] * wgt[0], b = maps[1][k] * wgt[1];
    float s = a + b;
    for (int i = 2; i < 9; i++) { float p = maps[i][k] * wgt[i]; s = p + s; }
    o->final_[k] = s;
  }
  if (feat9) for (int i = 0; i < 9; i++) memcpy(feat9 + (long)i * N, maps[i], N * 4);
  int fs = cv_round((double)(o->h / 15)); if (fs % 2 == 0) fs++;                /* :320-322 */
  orc_cv_gaussian_blur(o->final_, o->h, o->w, fs, (double)cv_round((double)(o->h / 30)), o->final_, o->rowbuf);
  mat_norm2(o->final_, N);
  if (out_f32) memcpy(out_f32, o->final_, N * 4);
  for (long k = 0; k < N; k++) {
    int v = cv_round((double)(o->final_[k] * 255.0f));
    out_u8[k] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
  }
  if (cam) { cam[0] = o->camx[F]; cam[1] = o->camy[F]; }
  o->frame++;
  return 0;
}

/* the two sequential float loops the GPU emulates, for the probe tests */
float orc_seq_sum_f32(const float* v, const unsigned char* sel, long n, int* count) {
  float s = 0; int c = 0;
  for (long i = 0; i < n; i++) if (sel[i]) { s = s + v[i]; c++; }
  *count = c;
  return s;
}
float orc_run_sum_f32(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, int* count) {
  float s = 0; int c = 0; (void)gh;
  for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) { float v = vals[(y / ms) * gw + x / ms]; if (v > t) { s = s + v; c++; } }
  *count = c;
  return s;
}
