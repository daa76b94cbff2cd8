/* cds_oracle.c — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 * See cds_oracle.h for the pinning status of each stage.  Every function cites the reference
 * file:line it restates (paths relative to the reference checkout).
 */
#include "cds_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXTOK 2048

/* =====================================================================================
 * Stage (a) — HM_16.0_features/source/Lib/TLibDecoder/code_mat_h.h + func.h
 *
 * The reference builds a 64x64 label map by repeatedly folding the token stream
 * (batch_64 :55-451), splits CU labels into PU labels (:281-385, func.h:189 pu()), removes
 * out-of-picture CUs (PartSize 15, :386-450), expands the MV stream (my_data :464-573) and
 * substitutes label -> MV in place (getFrame :595-603).  For legal quad-trees (depth <= 3) the
 * folding is equivalent to a depth-first walk that numbers leaf CUs in z-order; this
 * restatement performs that walk directly and reproduces every integer quirk of the
 * original explicitly (each is marked QUIRK below).  Equivalence is enforced by tests against
 * the verbatim reference build (oracle/_ref) and the authors' golden .mat maps.
 * ===================================================================================== */
typedef struct { int r0, c0, size, part; } orc_cu;

static int orc_walk(const int16_t* cupu, int n, int* pos, int r0, int c0, int size, orc_cu* cus, int* ncu) {
  if (*pos >= n) return -1;
  int t = cupu[(*pos)++];
  if (t == 99) {                       /* split node, TDecCu.cpp:279 */
    if (size <= 8) return -2;
    int s = size / 2;                  /* z-order: TL, TR, BL, BR (code_mat_h.h:113-116) */
    if (orc_walk(cupu, n, pos, r0, c0, s, cus, ncu)) return -1;
    if (orc_walk(cupu, n, pos, r0, c0 + s, s, cus, ncu)) return -1;
    if (orc_walk(cupu, n, pos, r0 + s, c0, s, cus, ncu)) return -1;
    if (orc_walk(cupu, n, pos, r0 + s, c0 + s, s, cus, ncu)) return -1;
    return 0;
  }
  if (*ncu >= 256) return -3;
  cus[*ncu].r0 = r0; cus[*ncu].c0 = c0; cus[*ncu].size = size; cus[*ncu].part = t;
  (*ncu)++;
  return 0;
}

/* number of labels a leaf CU occupies after the PU stage (len_pu func.h:26 plus one) */
static int orc_npu_labels(int part) {
  if (part <= 0 || part == 99) return 1;
  if (part == 3) return 4;
  if (part == 15) return 1;            /* stays one label until the :386-450 compaction */
  return 2;
}

/* PU index of pixel (r,c) inside a CU of side s — code_mat_h.h:340-384 reduced to geometry
 * (the member list is column-major, radii == s for a square CU). */
static int orc_pu_index(int part, int s, int r, int c) {
  switch (part) {
    case 1: return r >= s / 2;                                   /* 2NxN  :367 */
    case 2: return (r + c * s) >= (s * s) / 2;                   /* Nx2N  :340 */
    case 3: return (r >= s / 2 ? 2 : 0) + (c >= s / 2 ? 1 : 0);  /* NxN   :349 */
    case 4: return r >= s / 4;                                   /* 2NxnU :373 */
    case 5: return r >= s / 4 * 3;                               /* 2NxnD :379 */
    case 6: return (r + c * s) >= (s * s) / 4;                   /* nLx2N :343 */
    case 7: return (r + c * s) >= (s * s) / 4 * 3;               /* nRx2N :346 */
    default: return 0;
  }
}

static int orc_tokread(const int16_t* a, int n, int i) {
  /* the reference reads float[1000] rows terminated by 999 (TDecCu.cpp:159-161); reads past the
   * terminator are not reachable for legal streams.  Defined here as 999 then 0. */
  if (i < 0) return 0;
  if (i < n) return a[i];
  return i == n ? 999 : 0;
}

/* my_data, code_mat_h.h:464-573.  Returns the number of PUs x; fills pux/puy (size >= 512). */
static int orc_my_data(const int16_t* mv, int nummv, const int16_t* pred, int numpred, int* pux, int* puy) {
  static const int CAP = ORC_MAXTOK;
  int ext[ORC_MAXTOK];
  int addr_intra[512], total_intra = 0;
  memset(ext, 0, sizeof(ext));
  for (int i = 0; i < numpred && total_intra < 512; i++)
    if (pred[i] == 2) addr_intra[total_intra++] = i;           /* :470-476 */
  int ext_l = nummv + total_intra * 3;                         /* :477 */
  if (ext_l > CAP) ext_l = CAP;   /* the reference overflows mv_extend[1000] here (UB); not reachable on the bundled streams */
#define EXT_SET(i, v) do { int _i = (i); if (_i >= 0 && _i < CAP) ext[_i] = (v); } while (0)
  if (total_intra == 0) {                                      /* :482-486 */
    for (int i = 0; i < ext_l; i++) ext[i] = orc_tokread(mv, nummv, i);
  } else if (total_intra == 1) {                               /* :487-510 */
    int a = addr_intra[0];
    for (int i = 0; i < a * 4; i++) EXT_SET(i, orc_tokread(mv, nummv, i));
    EXT_SET(a * 4 + 0, 0); EXT_SET(a * 4 + 1, 0); EXT_SET(a * 4 + 2, 0);
    EXT_SET(a * 4 + 3, orc_tokread(mv, nummv, a));             /* QUIRK :500 — raw slot `a`, not 4a */
    int mvc = a * 4 + 1;
    for (int i = (a + 1) * 4; i < ext_l; i++) EXT_SET(i, orc_tokread(mv, nummv, mvc++));
  } else {                                                     /* :511-563 */
    int mv_length[512];
    for (int ii = 0; ii < total_intra; ii++) {
      int num_inter = 0, num_intra = 0;
      for (int i = 0; i < addr_intra[ii]; i++) {
        if (pred[i] == 1) num_inter++; else if (pred[i] == 2) num_intra++;
      }
      mv_length[ii] = 4 * num_inter + num_intra;               /* raw-stream offset of intra PU ii */
    }
    int end_point = 0, len_temp = 0;   /* `len_temp` is re-declared per iteration in the reference (:524) but keeps its slot */
    for (int ii = 0; ii < total_intra; ii++) {
      int base = mv_length[ii] + ii * 3;
      EXT_SET(base, 0); EXT_SET(base + 1, 0); EXT_SET(base + 2, 0);
      EXT_SET(base + 3, orc_tokread(mv, nummv, mv_length[ii]));
      if (ii == 0) {
        for (int i = 0; i < mv_length[0]; i++) EXT_SET(i, orc_tokread(mv, nummv, i));
        len_temp = mv_length[0];
      } else {
        int start_point = len_temp + 4;
        end_point = start_point + mv_length[ii] - mv_length[ii - 1] - 1;
        len_temp = end_point;
        int mcv = mv_length[ii - 1] + 1;
        for (int i = start_point; i < end_point; i++) EXT_SET(i, orc_tokread(mv, nummv, mcv++));
      }
    }
    if (mv_length[total_intra - 1] + 1 != numpred) {           /* QUIRK :554 — stream offset vs PU count */
      int mvc = mv_length[total_intra - 1] + 1;
      for (int i = end_point + 4; i < ext_l; i++) EXT_SET(i, orc_tokread(mv, nummv, mvc++));
    }
  }
#undef EXT_SET
  int x = ext_l / 4;                                           /* :564 */
  if (x > 500) x = 500;
  for (int j = 0; j < x; j++) { pux[j] = ext[2 + j * 4]; puy[j] = ext[3 + j * 4]; }
  return x;
}

/* in-place sequential substitution getFrame :595-603, as a per-cell chain (QUIRK: a freshly
 * written MV that equals a later label is substituted again; labels beyond x keep the label) */
static int orc_subst(int L, const int* mv, int x) {
  if (L < 1 || L > x) return L;
  int cur = L - 1, v = mv[cur];
  while (v - 1 > cur && v - 1 < x) { cur = v - 1; v = mv[cur]; }
  return v;
}

int orc_getframe(int w, int h, const int32_t* hdr, const int16_t* tok, int16_t* mvx, int16_t* mvy, uint8_t* depth) {
  int ncol = (w + 63) / 64, nrow = (h + 63) / 64;   /* == int(w/64.0+0.5) for every supported size, code_mat_h.h:20-21 */
  int h4 = h / 4, w4 = w / 4;
  orc_cu cus[256];
  int base[256], pux[512], puy[512];
  for (int line = 0; line < nrow * ncol; line++) {
    const int16_t* cupu = tok + hdr[line * 4 + 0];
    int ncupu = hdr[line * 4 + 1], npred = hdr[line * 4 + 2], nmv = hdr[line * 4 + 3];
    const int16_t* pred = cupu + ncupu;
    const int16_t* mv = pred + npred;
    if (ncupu > ORC_MAXTOK || npred > ORC_MAXTOK || nmv > ORC_MAXTOK) return -10;
    int ncu = 0, pos = 0;
    if (orc_walk(cupu, ncupu, &pos, 0, 0, 64, cus, &ncu)) return -1;
    /* label bookkeeping; `lab15` = label a PartSize-15 CU carries before compaction (negated, :409) */
    int next = 1, next15 = 1;
    int lab15[256];
    for (int c = 0; c < ncu; c++) {
      int part = cus[c].part;
      if (ncu == 1) part = 0;            /* QUIRK func.h:177 — token 0 is overwritten by 99: an unsplit CTU's PartSize is ignored */
      base[c] = next; lab15[c] = next15;
      next15 += orc_npu_labels(part);
      next += (part == 15) ? 0 : orc_npu_labels(part);
      cus[c].part = part;
    }
    int x = orc_my_data(mv, nmv, pred, npred, pux, puy);
    int R0 = 64 * (line / ncol), C0 = 64 * (line % ncol);
    for (int c = 0; c < ncu; c++) {
      int s = cus[c].size, part = cus[c].part;
      int d = s == 64 ? 0 : s == 32 ? 1 : s == 16 ? 2 : s == 8 ? 3 : 0;   /* depth_cal func.h:11 */
      for (int r = 0; r < s; r += 4) {
        int R = R0 + cus[c].r0 + r;
        if (R >= h) break;
        for (int cc = 0; cc < s; cc += 4) {
          int C = C0 + cus[c].c0 + cc;
          if (C >= w) break;
          int L = (part == 15) ? -lab15[c] : base[c] + orc_pu_index(part, s, r, cc);
          size_t o = (size_t)(R / 4) * w4 + C / 4;
          if (R / 4 < h4 && C / 4 < w4) {
            mvx[o] = (int16_t)orc_subst(L, pux, x);
            mvy[o] = (int16_t)orc_subst(L, puy, x);
            depth[o] = (uint8_t)d;
          }
        }
      }
    }
  }
  return 0;
}

/* =====================================================================================
 * Stage (b)
 * ===================================================================================== */
/* computecontrast5.cpp:52-96 (img_extract :11, outer_pro :30, sumup :41); column-major. */
void orc_computecontrast5(const double* pad, int m1, int n1, const double* W, int len, int win,
                          int m, int n, double* out) {
  (void)n1;
  for (int row = 1 + len; row <= m + len; row++) {
    for (int col = 1 + len; col <= n + len; col++) {
      double centre = pad[(size_t)(col - 1) * m1 + row - 1];                      /* :83 */
      double acc = 0;
      for (int j = 0; j < win; j++) {                                              /* patch column */
        const double* src = pad + (size_t)(col - len - 1 + j) * m1 + (row - len - 1);  /* :24 */
        for (int i = 0; i < win; i++) {
          double p = src[i];
          if (p != 0) p = fabs(p - centre);                                        /* :86-87 */
          acc = acc + p * W[j * win + i];                                          /* :89-90 */
        }
      }
      out[(size_t)(col - len - 1) * m + row - len - 1] = acc;                      /* :91 */
    }
  }
}

static double orc_round(double x) { return x < 0 ? -floor(-x + 0.5) : floor(x + 0.5); }  /* MATLAB round */

/* Sudoku_contrastcpp5.m:1-52.  A row-major rows x cols array is the column-major cols x rows
 * transpose; W is symmetric and the window is the same in both directions, so the core can
 * run on the transposed view unchanged. */
void orc_sudoku_contrast(const double* map, int rows, int cols, double delta, double cusize,
                         double patchsize, double* out) {
  int win = (int)orc_round(patchsize / cusize);                                    /* :6 */
  int len = win / 2;                                                               /* :7 */
  double* W = (double*)malloc(sizeof(double) * win * win);
  for (int i = 1; i <= win; i++)
    for (int j = 1; j <= win; j++) {
      double dx = i - (len + 1), dy = j - (len + 1);
      double R = sqrt(dx * dx + dy * dy);                                          /* :12 */
      W[(j - 1) * win + (i - 1)] = exp(-(R * R) / (2 * delta * delta));            /* :14 */
    }
  int pr = rows + 2 * len, pc = cols + 2 * len;
  double* pad = (double*)calloc((size_t)pr * pc, sizeof(double));
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++)
      pad[(size_t)(r + len) * pc + c + len] = map[(size_t)r * cols + c] + 0.00000001;   /* :18, :26-32 */
  orc_pm_norm(pad, (long)pr * pc);                                                 /* :34 */
  /* transposed view: column-major (pc x pr) */
  orc_computecontrast5(pad, pc, pr, W, len, win, cols, rows, out);                 /* :37 */
  double mx = -INFINITY;
  for (long i = 0; i < (long)rows * cols; i++) if (out[i] > mx) mx = out[i];       /* MATLAB max ignores NaN */
  for (long i = 0; i < (long)rows * cols; i++) out[i] = out[i] / mx;               /* :50 */
  free(W); free(pad);
}

/* cu_contrast5.m:3-63 */
void orc_cu_contrast5(const double* map, int rows, int cols, int cusize, double delta,
                      double patchsize, double* out) {
  if (cusize != 1) {
    int numrow = (rows + cusize - 1) / cusize, numcol = (cols + cusize - 1) / cusize;   /* :7-8 */
    double* sub = (double*)malloc(sizeof(double) * numrow * numcol);
    double* con = (double*)malloc(sizeof(double) * numrow * numcol);
    for (int k = 0; k < numrow; k++)
      for (int j = 0; j < numcol; j++) {
        int r0 = k * cusize, c0 = j * cusize;
        int r1 = (k == numrow - 1) ? rows : (k + 1) * cusize;
        int c1 = (j == numcol - 1) ? cols : (j + 1) * cusize;
        double s = 0;                                     /* mean(mean(temp)) :29 — column means first */
        for (int c = c0; c < c1; c++) {
          double cs = 0;
          for (int r = r0; r < r1; r++) cs += map[(size_t)r * cols + c];
          s += cs / (r1 - r0);
        }
        sub[k * numcol + j] = s / (c1 - c0);
      }
    orc_sudoku_contrast(sub, numrow, numcol, delta, cusize, patchsize, con);       /* :34 */
    double mx = -INFINITY;
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        double v = con[(r / cusize) * numcol + (c / cusize)];                      /* :36-54 */
        out[(size_t)r * cols + c] = v;
        if (v > mx) mx = v;
      }
    for (long i = 0; i < (long)rows * cols; i++) out[i] = out[i] / mx;             /* :55 */
    free(sub); free(con);
  } else {
    orc_sudoku_contrast(map, rows, cols, delta, cusize, patchsize, out);           /* :59 */
    double mx = -INFINITY;
    for (long i = 0; i < (long)rows * cols; i++) if (out[i] > mx) mx = out[i];
    for (long i = 0; i < (long)rows * cols; i++) out[i] = out[i] / mx;             /* :60 */
  }
}

/* =====================================================================================
 * Stage (d) — libsvm-3.17/svm.cpp:2501-2575 (C-SVC, nr_class = 2) + Kernel::k_function RBF :325-365
 * with the dense instance rows built by matlab/svmpredict.c:165-170.  A dense x against a sparse
 * SV walks every index 1..F in order; a missing SV entry contributes x_j*x_j == (x_j-0)^2.
 * ===================================================================================== */
void orc_svm_rbf_predict(int nsv, int nfeat, const double* SV, const double* coef, double rho,
                         double gamma, int label0, int label1, const double* X, int N,
                         double* dec, double* label) {
  for (int n = 0; n < N; n++) {
    double sum = 0;
    for (int i = 0; i < nsv; i++) {
      double d2 = 0;
      for (int j = 0; j < nfeat; j++) {
        double d = X[(size_t)j * N + n] - SV[(size_t)i * nfeat + j];
        d2 += d * d;                                                                /* :332-333 */
      }
      sum += coef[i] * exp(-gamma * d2);                                           /* :364, :2549-2552 */
    }
    sum -= rho;                                                                    /* :2553 */
    dec[n] = sum;
    label[n] = sum > 0 ? label0 : label1;                                          /* :2556-2559 */
  }
}

/* =====================================================================================
 * Stage (e) helpers — MATLAB semantics (PARITY UNPINNED: no MATLAB here, no reference outputs)
 * ===================================================================================== */
/* pm_norm.m:1-5 — MATLAB min/max skip NaN */
void orc_pm_norm(double* a, long n) {
  double mn = INFINITY;
  for (long i = 0; i < n; i++) if (a[i] < mn) mn = a[i];
  double mx = -INFINITY;
  for (long i = 0; i < n; i++) { a[i] -= mn; if (a[i] > mx) mx = a[i]; }
  for (long i = 0; i < n; i++) a[i] = a[i] / mx;
}

/* mysterythrehold.m:3-7 */
void orc_mysterythrehold(double* a, long n, int im_height, int im_width) {
  double s = 0, mx = -INFINITY;
  for (long i = 0; i < n; i++) { s += a[i]; if (a[i] > mx) mx = a[i]; }
  double ave = s / ((double)im_height * im_width);                                 /* :3 (full-res pixel count) */
  double sb = 0; long nb = 0;
  for (long i = 0; i < n; i++) if (a[i] > mx - ave) { sb += a[i]; nb++; }          /* :4 */
  double thresh = sb / (double)nb;                                                 /* :5 (NaN when empty) */
  double mx2 = -INFINITY;
  for (long i = 0; i < n; i++) { if (a[i] > thresh) a[i] = thresh; if (a[i] > mx2) mx2 = a[i]; }   /* :6 */
  for (long i = 0; i < n; i++) a[i] = a[i] / (mx2 + 0.000001);                     /* :7 */
}

/* immove.m:1-35 — positive mov_x shifts content right, positive mov_y shifts it down; zero fill */
void orc_immove(const double* im, int rows, int cols, int mov_x, int mov_y, double* out) {
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      int sr = r - mov_y, sc = c - mov_x;
      out[(size_t)r * cols + c] = (sr >= 0 && sr < rows && sc >= 0 && sc < cols) ? im[(size_t)sr * cols + sc] : 0.0;
    }
}

/* upsample.m:1-10 / fourX2.m:1-17 — replicate each element k x k, crop to out_rows x out_cols */
void orc_replicate(const double* a, int rows, int cols, int k, int out_rows, int out_cols, double* out) {
  for (int r = 0; r < out_rows; r++)
    for (int c = 0; c < out_cols; c++) {
      int sr = r / k, sc = c / k;
      out[(size_t)r * out_cols + c] = (sr < rows && sc < cols) ? a[(size_t)sr * cols + sc] : 0.0;
    }
}

/* getdistMatrix.m:1-17 */
void orc_distmatrix(int hei, int wid, double* out) {
  int mx_ = hei / 2, my_ = wid / 2;
  double mx = -INFINITY;
  for (int x = 1; x <= hei; x++)
    for (int y = 1; y <= wid; y++) {
      double d = floor(sqrt((double)(x - mx_) * (x - mx_) + (double)(y - my_) * (y - my_)));
      out[(size_t)(x - 1) * wid + (y - 1)] = d;
      if (d > mx) mx = d;
    }
  for (long i = 0; i < (long)hei * wid; i++) out[i] = 1 - out[i] / mx;
}

/* fspecial('gaussian', k, sigma) is exp(-(x^2+y^2)/(2 sigma^2)) / sum on x,y in -(k-1)/2..(k-1)/2;
 * the eps*max truncation never triggers for k ~ 4 sigma, so the kernel is the outer product of
 * this normalised 1-D factor. */
void orc_gauss_kernel(int k, double sigma, double* g) {
  double siz = (k - 1) / 2.0, s = 0;
  for (int i = 0; i < k; i++) { double x = i - siz; g[i] = exp(-(x * x) / (2 * sigma * sigma)); s += g[i]; }
  for (int i = 0; i < k; i++) g[i] /= s;
}

/* imfilter(A, h): correlation, zero padding, 'same'; kernel origin floor((k+1)/2) (1-based) */
void orc_imfilter_gauss(const double* a, int rows, int cols, int k, double sigma, double* out) {
  double* g = (double*)malloc(sizeof(double) * k);
  double* tmp = (double*)malloc(sizeof(double) * (size_t)rows * cols);
  orc_gauss_kernel(k, sigma, g);
  int c0 = (k + 1) / 2 - 1;   /* 0-based index of the origin tap */
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      double s = 0;
      for (int u = 0; u < k; u++) { int sc = c + u - c0; if (sc >= 0 && sc < cols) s += g[u] * a[(size_t)r * cols + sc]; }
      tmp[(size_t)r * cols + c] = s;
    }
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      double s = 0;
      for (int u = 0; u < k; u++) { int sr = r + u - c0; if (sr >= 0 && sr < rows) s += g[u] * tmp[(size_t)sr * cols + c]; }
      out[(size_t)r * cols + c] = s;
    }
  free(g); free(tmp);
}

/* imresize(A,[out_rows out_cols],'bilinear') — MATLAB's contributions(): triangle kernel, widened
 * by 1/scale when shrinking (antialiasing), weights normalised, symmetric boundary indices. */
static void orc_resize_weights(int in_len, int out_len, int* P_out, int** idx_out, double** w_out) {
  double scale = (double)out_len / in_len;
  double kw = 2.0;
  int aa = scale < 1;
  if (aa) kw = kw / scale;
  int P = (int)ceil(kw) + 2;
  int* idx = (int*)malloc(sizeof(int) * (size_t)out_len * P);
  double* wt = (double*)malloc(sizeof(double) * (size_t)out_len * P);
  for (int x = 1; x <= out_len; x++) {
    double u = x / scale + 0.5 * (1 - 1 / scale);
    int left = (int)floor(u - kw / 2);
    double s = 0;
    for (int p = 0; p < P; p++) {
      int ind = left + p;
      double t = u - ind;
      if (aa) t = scale * t;
      double v = (t >= -1 && t < 0) ? t + 1 : (t >= 0 && t <= 1) ? 1 - t : 0;
      if (aa) v = scale * v;
      wt[(size_t)(x - 1) * P + p] = v; s += v;
      int mm = 2 * in_len;
      int q = ((ind - 1) % mm + mm) % mm;            /* aux = [1:n, n:-1:1] */
      idx[(size_t)(x - 1) * P + p] = q < in_len ? q : mm - 1 - q;   /* 0-based */
    }
    for (int p = 0; p < P; p++) wt[(size_t)(x - 1) * P + p] /= s;
  }
  *P_out = P; *idx_out = idx; *w_out = wt;
}

void orc_imresize_bilinear(const double* a, int rows, int cols, int out_rows, int out_cols, double* out) {
  int Pr, Pc; int *ir, *ic; double *wr, *wc;
  orc_resize_weights(rows, out_rows, &Pr, &ir, &wr);
  orc_resize_weights(cols, out_cols, &Pc, &ic, &wc);
  double* tmp = (double*)malloc(sizeof(double) * (size_t)out_rows * cols);
  for (int r = 0; r < out_rows; r++)
    for (int c = 0; c < cols; c++) {
      double s = 0;
      for (int p = 0; p < Pr; p++) s += wr[(size_t)r * Pr + p] * a[(size_t)ir[(size_t)r * Pr + p] * cols + c];
      tmp[(size_t)r * cols + c] = s;
    }
  for (int r = 0; r < out_rows; r++)
    for (int c = 0; c < out_cols; c++) {
      double s = 0;
      for (int p = 0; p < Pc; p++) s += wc[(size_t)c * Pc + p] * tmp[(size_t)r * cols + ic[(size_t)c * Pc + p]];
      out[(size_t)r * out_cols + c] = s;
    }
  free(tmp); free(ir); free(ic); free(wr); free(wc);
}
