/* Stand-in for HM_16.0_features/source/Lib/TLibDecoder/TDecCu.h:41-45 — only the five
 * compile-time geometry macros that code_mat_h.h consumes.  REF_W / REF_H come from -D.
 * Force-included (-include) with -D__TDECCU__ so the real TDecCu.h (same directory as
 * code_mat_h.h, hence found first) is skipped through its own include guard. */
#ifndef CDS_ORACLE_GEOMETRY_MACROS_H
#define CDS_ORACLE_GEOMETRY_MACROS_H
#define VIDEO_HIGHT REF_H
#define VIDEO_WIDTH REF_W
#define LCUNUM (((REF_H + 63) / 64) * ((REF_W + 63) / 64))
#define FRAMENUM 500
#define FRAMERATE 50
#endif
