// fast.cuh — the reference's FAST version (TDecGop.cpp:225-333, 486-656; salmat.cpp) behind the clip context.
#pragma once
#include "cds_common.cuh"

namespace cds {

struct FastState;

// PS = cvRound(0.3 fps), PT = cvRound(0.7 fps) (salmat.cpp:18-19).  B = pictures per launch group.
int fast_create(int w, int h, double fps, int B, FastState** out);
void fast_destroy(FastState* s);
int fast_clear(FastState* s, cudaStream_t st);          // zero the rings (cds_seek)
int fast_halo(const FastState* s);                      // pictures of temporal context a frame-range shard needs: PS + PT - 2
// n <= B pictures with FrameIndex first_poc..first_poc+n-1.  mvx / mvy / dep: stage (a) outputs [n][h4*w4]; bytes [n][nctu].
// out ([n][h][w], u8 or f32) may be null: temporal context only.  cam_out (device, [n][2]) may be null.
int fast_run_batch(FastState* s, int first_poc, int n, const int16_t* mvx, const int16_t* mvy, const uint8_t* dep,
                   const uint16_t* bytes, void* out, int out_u8, int* cam_out, cudaStream_t st);
// the two halves of fast_run_batch: the front half of the next batch may run on another stream beside the back half of this one
// (the ring holds 2 B + PS + PT pictures); the caller orders them with events
int fast_front(FastState* s, int first_poc, int n, const int16_t* mvx, const int16_t* mvy, const uint8_t* dep, const uint16_t* bytes,
               int* cam_out, cudaStream_t st);
int fast_back(FastState* s, int first_poc, int n, void* out, int out_u8, cudaStream_t st);
// per-stage timing: the clip context's marker is called at the start of each stage while profiling is on (the two chains then run
// one after the other instead of side by side)
enum FastStage { FAST_ST_CAM_SCAN = 0, FAST_ST_CAM_VOTE, FAST_ST_SMOOTH, FAST_ST_TEMPORAL, FAST_ST_THR_SPATIAL, FAST_ST_COMBINE_BLUR_H,
                 FAST_ST_BLUR_V, FAST_ST_OUTPUT, FAST_ST_COUNT };
void fast_set_marker(FastState* s, void (*mark)(void* user, int stage, cudaStream_t st), void* user);
void fast_profile(FastState* s, bool on);
int fast_get_cam(const FastState* s, int poc, int* xy);   // camera motion of a picture still in the ring (host copy)
// probes (tests): `for (i < n) if (sel[i]) S = S + v[i];` and MapThreshold's sum over a map of replicated blocks, both in float
int fast_probe_seq_sum(const float* v, const unsigned char* sel, int n, float* sum, int* count);
int fast_probe_run_sum(const float* vals, int gh, int gw, int rows, int cols, int ms, double t, float* sum, int* count);
// debug / tests: copy one intermediate of the last batch (see fast.cu) to the host
int fast_debug_copy(FastState* s, const char* which, int frame_in_batch, void* dst, size_t bytes, cudaStream_t st);

}  // namespace cds
