// pack.cpp — the reference decoder's capture statics -> packed wire format (host).
// TDecCu::bits / MV / CUPU / Pred are float[LCUNUM][1000] arrays terminated by 999 (TDecCu.h:100-106, writers
// TDecCu.cpp:153-161, 165-289); the path's wire format is hdr[nctu][4] = {token offset, n_cupu, n_pred, n_mv} + int16 tokens.
#include <stdint.h>
#include <string.h>
#include "../../include/cds.h"

extern "C" long cds_syntax_pack_ref(int nctu, const float* MV, const float* Pred, const float* CUPU,
                                    int32_t* hdr, int16_t* tok) {
  if (nctu < 1 || !MV || !Pred || !CUPU || !hdr || !tok) return CDS_EINVAL;
  long off = 0;
  for (int a = 0; a < nctu; a++) {
    const float* src[3] = {CUPU + (size_t)a * 1000, Pred + (size_t)a * 1000, MV + (size_t)a * 1000};
    hdr[a * 4 + 0] = (int32_t)off;
    for (int s = 0; s < 3; s++) {
      int n = 0;
      while (n < 1000 && src[s][n] != 999.f) { tok[off + n] = (int16_t)src[s][n]; n++; }   // 999 sentinel, TDecCu.cpp:159-161
      if (n == 1000) return CDS_EFORMAT;
      hdr[a * 4 + 1 + s] = n; off += n;
    }
  }
  return off;
}
