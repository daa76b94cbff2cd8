// synth.cpp — synthetic per-CTU HEVC syntax (host only; built into libcds_synth.so, which holds no CUDA code, so the
// CPU arms of bench.py and the CPU tests can generate inputs without mapping the product library).  Stands in for the HM16.0 CABAC parser (the
// reference host producer, timed separately) in throughput runs and fuzz tests; statistics follow
// SURVEY.md section 8(d).  The streams are LEGAL in the sense of TDecCu::MVoutput
// (TDecCu.cpp:165-289): depth-first CU tokens (99 = split, else PartSize, 15 = block outside the
// picture after HM's forced boundary splits), one Pred token per PU (1 inter / 2 intra) and
// [interDir, refPOC, mvHor, mvVer] per inter PU / [0] per intra PU.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../include/cds_synth.h"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  }
};

uint64_t mix(uint64_t a, uint64_t b) {
  Rng r(a * 0x9E3779B97F4A7C15ull + b + 0x1234567ull);
  return r.next();
}

struct Gen {
  int w, h, poc, gx, gy;
  int X0, Y0;             // CTU origin in the picture
  Rng rng;
  int16_t *cupu, *pred, *mv;
  bool adv = false;        // adversarial statistics for differential fuzzing (seed bit 60)
  int ncupu = 0, npred = 0, nmv = 0;
  Gen(uint64_t seed) : rng(seed) {}

  int clip16(long v) { return v > 32767 ? 32767 : v < -32768 ? -32768 : (int)v; }

  void leaf(int size) {
    const bool intra = poc == 0 || rng.uni() < (adv ? 0.35 : 0.1);
    int part = 0, npu = 1;
    if (intra) {
      if (size == 8 && rng.uni() < (adv ? 0.5 : 0.2)) { part = 3; npu = 4; }
    } else {
      double u = rng.uni();
      if (adv) u = 0.4 + 0.6 * u;   // more 2NxN / Nx2N / AMP
      if (u < 0.70) part = 0;
      else if (u < 0.78) part = 1;
      else if (u < 0.86) part = 2;
      else if (u < 0.98) part = size >= 16 ? 4 + (int)((u - 0.86) / 0.03) : 0;
      else part = 0;       // NxN is intra-only at 8x8
      if (part > 7) part = 7;
      npu = part == 0 ? 1 : 2;
    }
    cupu[ncupu++] = (int16_t)part;
    for (int i = 0; i < npu; i++) {
      if (intra) { pred[npred++] = 2; mv[nmv++] = 0; }
      else {
        pred[npred++] = 1;
        mv[nmv++] = 1;
        mv[nmv++] = (int16_t)(poc - 1);
        const double sg = adv ? 4.0 : 12.0;   // small MVs collide with PU labels (substitution quirk)
        mv[nmv++] = (int16_t)clip16(gx + lround(rng.normal() * sg));
        mv[nmv++] = (int16_t)clip16(gy + lround(rng.normal() * sg));
      }
    }
  }

  void node(int x0, int y0, int size, int depth) {
    static const double psplit[3] = {0.55, 0.45, 0.35};
    const int ax = X0 + x0, ay = Y0 + y0;
    if (ax >= w || ay >= h) { cupu[ncupu++] = 15; return; }              // outside: never decoded
    bool split;
    if (ax + size > w || ay + size > h) split = true;                      // HM forced boundary split
    else split = size > 8 && rng.uni() < (adv ? 0.8 : psplit[depth]);
    if (!split) { leaf(size); return; }
    cupu[ncupu++] = 99;
    const int s = size / 2;
    node(x0, y0, s, depth + 1); node(x0 + s, y0, s, depth + 1);
    node(x0, y0 + s, s, depth + 1); node(x0 + s, y0 + s, s, depth + 1);
  }
};

}  // namespace

extern "C" long cds_synth_frame(int w, int h, uint64_t seed, int poc, int32_t* hdr, int16_t* tok, size_t tok_cap,
                                uint16_t* ctu_bytes) {
  if (!hdr || !tok || !ctu_bytes || w < 8 || h < 8 || (w % 8) || (h % 8)) return -1;
  const int ncol = (w + 63) / 64, nrow = (h + 63) / 64;
  Rng fr(mix(seed, (uint64_t)poc));
  const bool adv = (seed >> 60) & 1;
  int gx = (int)(fr.next() % 65) - 32, gy = (int)(fr.next() % 65) - 32;
  if (adv) { gx = gx / 6 + 4; gy = gy / 6 + 3; }
  size_t off = 0;
  int16_t cupu[96], pred[260], mv[1100];
  for (int a = 0; a < nrow * ncol; a++) {
    Gen g(mix(mix(seed, (uint64_t)poc), (uint64_t)a + 77));
    g.w = w; g.h = h; g.poc = poc; g.gx = gx; g.gy = gy; g.adv = adv;
    g.X0 = 64 * (a % ncol); g.Y0 = 64 * (a / ncol);
    g.cupu = cupu; g.pred = pred; g.mv = mv;
    g.node(0, 0, 64, 0);
    const size_t n = (size_t)g.ncupu + g.npred + g.nmv;
    if (off + n > tok_cap) return -2;
    hdr[a * 4 + 0] = (int32_t)off; hdr[a * 4 + 1] = g.ncupu; hdr[a * 4 + 2] = g.npred; hdr[a * 4 + 3] = g.nmv;
    memcpy(tok + off, cupu, g.ncupu * 2); off += g.ncupu;
    memcpy(tok + off, pred, g.npred * 2); off += g.npred;
    memcpy(tok + off, mv, g.nmv * 2); off += g.nmv;
    double e = -40.0 * log(1.0 - g.rng.uni());
    long b = lround(e);
    ctu_bytes[a] = (uint16_t)(b > 4095 ? 4095 : b);
  }
  return (long)off;
}

