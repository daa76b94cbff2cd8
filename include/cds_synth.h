/* cds_synth.h — synthetic per-CTU HEVC syntax producer (libcds_synth.so: host-only C++, no CUDA).
 *
 * Stands in for the reference's host producer — HM16.0 CABAC parsing with the capture hooks of TDecSlice.cpp:333-341 and
 * TDecCu.cpp:153-289 — in throughput runs and fuzz tests (synthetic syntax buffers of the BASELINE shapes).
 * It is NOT part of the drop-in surface and lives outside libcds_b200.so so that bench.py's reference arm and the CPU-only
 * tests never map the product library. */
#ifndef CDS_SYNTH_H
#define CDS_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Fills hdr[nctu*4] = {token offset, n_cupu, n_pred, n_mv}, tok (capacity tok_cap int16) and ctu_bytes[nctu] for picture `poc`
 * of the clip identified by `seed` (the packed wire format of include/cds.h).  Seed bit 60 selects adversarial statistics
 * (deep trees, many intra / NxN / AMP CUs, small MVs that collide with PU labels) used by the differential fuzz tests.
 * Returns the number of tokens written, -1 on bad arguments, -2 when tok_cap is too small. */
long cds_synth_frame(int w, int h, uint64_t seed, int poc, int32_t* hdr, int16_t* tok, size_t tok_cap,
                     uint16_t* ctu_bytes);
#ifdef __cplusplus
}
#endif
#endif
