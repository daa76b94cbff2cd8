// C-callable wrapper around the reference's verbatim getFrame()
// (HM_16.0_features/source/Lib/TLibDecoder/code_mat_h.h:575).  The reference header is
// #included from /root/reference at build time (never copied).  TEST INFRASTRUCTURE ONLY.
#include "code_mat_h.h"
#undef width
#undef height
extern "C" {
int ref_getframe_width(void) { return VIDEO_WIDTH; }
int ref_getframe_height(void) { return VIDEO_HIGHT; }
int ref_getframe_nctu(void) { return num_patch; }
// mv/pred/cupu: nctu x 1000 floats, each row terminated by 999 (reference layout, TDecCu.h:100-106)
void ref_getframe(float* mv, float* pred, float* cupu, float* mvx, float* mvy, float* depth) {
  getFrame((float(*)[1000])mv, (float(*)[1000])pred, (float(*)[1000])cupu,
           (float(*)[VIDEO_WIDTH / 4])mvx, (float(*)[VIDEO_WIDTH / 4])mvy,
           (float(*)[VIDEO_WIDTH / 4])depth);
}
}
