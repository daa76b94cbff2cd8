// maps.cuh — declarations shared by maps.cu (stage (e) kernels) and pipeline.cu (orchestration).
#pragma once
#include "cds_common.cuh"
#include <vector>

struct cds_svm_model;

namespace cds {

// A banded linear operator out[i] = sum_t w[i][t] * in[start[i] + t]  (rows of a sparse matrix with
// contiguous support).  Host copy + device copy.
struct Banded {
  int n_out = 0, n_in = 0, nt = 0;
  int span4 = 0;             // max over i of start[min(i+3, n_out-1)] - start[i]; -1 when start is not non-decreasing
  std::vector<float> w;      // [n_out][nt]
  std::vector<int> start;    // [n_out]  (start + nt may exceed n_in; weights there are 0 and reads are clamped)
  float* d_w = nullptr;
  float* d_wT = nullptr;     // [nt][n_out] transposed copy (coalesced reads when a thread owns an output column)
  int* d_start = nullptr;
  int upload();
  void release();
};

// Polyphase Gaussian for imfilter(fourX2(X), fspecial('gaussian',k,sigma)): full-res sample 4c+ph
// = sum_d G[ph][d - dmin] * X[c + d], X zero outside (main.m:208-213, fourX2.m, imfilter zero padding).
struct Polyphase {
  int dmin = 0, nt = 0;
  std::vector<float> g;      // [nt][4]  (tap-major so one float4 holds the 4 phases of a tap)
  float* d_g = nullptr;
  float* d_aimg = nullptr;   // tensor-core V pass: banded 128 x KP weight matrix, hi | lo, shared-memory image (built on first use)
  int upload();
  void release();
};

void build_gauss_polyphase(int k, double sigma, Polyphase* out);
// dense helpers in double, row-major [n_out][n_in]
std::vector<double> dense_gauss_rep4(int n_cells, int k, double sigma);              // (4*n_cells) x n_cells
std::vector<double> dense_gauss(int n, int k, double sigma);                         // n x n, zero padded correlation
std::vector<double> dense_imresize(int n_in, int n_out);                             // n_out x n_in, MATLAB bilinear (+antialias)
std::vector<double> dense_mul(const std::vector<double>& A, int ar, int ac, const std::vector<double>& B, int bc);
void banded_from_dense(const std::vector<double>& A, int n_out, int n_in, Banded* out);

// ---- launches (all asynchronous on `st`) ----
// Vx / Vy rings are double (the reference's mv_x / mv_y cell arrays); cxy (optional) receives the smoothed
// mvx_crop | mvy_crop of this batch in double, [2][B][n4], for contrast_prep_f64_launch
int smooth_launch(int n4, int PS, int ring_len, int first_poc, int B, const double* Vx, const double* Vy, const float* mvn,
                  const float* bitn, const float* depn, float4* sm4 /*ring [RL][n4] {mvx, mvy, bit, depth} smoothed*/, float* mv_s,
                  double* cxy, cudaStream_t st);
// Sudoku :18-34 for the signed mv maps in double -> padded FP32 maps; keys [nmaps][2] preset to {~0ull, 0}
int contrast_prep_f64_launch(const double* src, int nmaps, int rows, int cols, int len, unsigned long long* keys, float* pad,
                             cudaStream_t st);
// mysterythrehold on the three smoothed h4 x w4 maps -> Xmaps slots 0,3,6
int myst_small_launch(int n4, int H, int W, int ring_len, int first_poc, int B, const float* mv_s, const float4* sm4,
                      float* Xmaps, cudaStream_t st);
int temporal_launch(int h4, int w4, int PT, int ring_len, int first_poc, int B, const float4* sm4, const int* cam_ring,
                    const float* wtab, float* Xmaps, cudaStream_t st);
// block-mean downsample (cusize) + 1e-8 + min/max of one map type; src_stride/src_off select the map inside a batch buffer
int contrast_sub_launch(const float* src, long src_frame_stride, int ring_len, int first_slot, int B, int rows, int cols,
                        int cusize, float* sub, unsigned* minmax_keys, cudaStream_t st);
int contrast_pad_launch(const float* sub, const unsigned* minmax_keys, int B, int srows, int scols, int len, float* pad,
                        cudaStream_t st);
// raw contrast -> X maps: replicate by cusize, /max; for mv: hypot of the two normalised components
int contrast_post_launch(const float* raw, const unsigned* maxbits, int B, int srows, int scols, int cusize, int rows, int cols,
                         float* Xmaps, int slot, cudaStream_t st);
int contrast_post_mv_launch(const float* rawx, const float* rawy, const unsigned* maxx, const unsigned* maxy, int B, int rows,
                            int cols, float* Xmaps, int slot, cudaStream_t st);
// full-res: T = X * Gh^T (h4 x W), then Y = Gv * T (H x W) reduced to min/max/sum, optionally stored
int fullres_hpass_launch(const float* X, int nmaps, int h4, int w4, const Polyphase& P, float* T, cudaStream_t st);
int fullres_vpass_launch(const float* T, int nmaps, int h4, int W, const Polyphase& P, float* Ystore /*or null*/,
                         const int* store_index /*[nmaps] -> slot in Ystore or -1*/, float* partial, int* nblocks_out, cudaStream_t st);
int fir_threads(int W4);                      // CTA width (96 or 128 threads of 4 columns) wasting the fewest columns
int fullres_vpass_nblocks(int h4, int W);   // partial blocks per map written by fullres_vpass_launch
int stats_finalize_launch(const float* partial, int nmaps, int nblocks, float* stats /*[nmaps][4] min,max,sum_lo,sum_hi*/, cudaStream_t st);
// generic banded operators, batched over nmaps; xf (optional, [nmaps][4] = a,b,thr,unused): v = min((in - a) * b, thr)
int banded_rows_launch(const float* in, int nmaps, int in_rows, int cols, const Banded& B, const float* xf, float* out, cudaStream_t st,
                       int outer_of_three = 0);
int banded_cols_launch(const float* in, int nmaps, int rows, int in_cols, const Banded& B, float* out, cudaStream_t st, int outer_of_three = 0);

}  // namespace cds
