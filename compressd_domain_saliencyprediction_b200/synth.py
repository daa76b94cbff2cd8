"""Synthetic per-CTU syntax from libcds_synth.so — a host-only library without CUDA code, so the
CPU arms of bench.py and the CPU tests generate their inputs without mapping libcds_b200.so."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_L = None


def synth_lib():
    global _L
    if _L is None:
        path = os.path.join(HERE, "libcds_synth.so")
        if not os.path.exists(path):
            from . import build
            build.build_synth()
        L = C.CDLL(path)
        L.cds_synth_frame.restype = C.c_long
        L.cds_synth_frame.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _L = L
    return _L


def nctu(w, h):
    return ((w + 63) // 64) * ((h + 63) // 64)


def synth_frame(w, h, seed, poc):
    """Synthetic per-CTU syntax for throughput runs.  Returns hdr, tok, ctu_bytes."""
    n = nctu(w, h)
    hdr = np.zeros((n, 4), np.int32)
    tok = np.zeros(1300 * n, np.int16)
    byts = np.zeros(n, np.uint16)
    nt = synth_lib().cds_synth_frame(w, h, seed, poc, hdr.ctypes.data, tok.ctypes.data, tok.size, byts.ctypes.data)
    if nt < 0:
        raise ValueError(f"cds_synth_frame({w}x{h}, poc {poc}) failed: {nt}")
    return hdr, tok[:nt].copy(), byts


def synth_clip(w, h, seed, nframes, first_poc=0):
    """nframes pictures packed for the batch entry points: hdr[F][nctu][4], tok, tok_base[F+1], bytes[F][nctu]."""
    hdrs, toks, byts, base = [], [], [], [0]
    for f in range(nframes):
        hd, tk, by = synth_frame(w, h, seed, first_poc + f)
        hdrs.append(hd); toks.append(tk); byts.append(by); base.append(base[-1] + tk.size)
    return np.stack(hdrs), np.concatenate(toks), np.array(base, np.int64), np.stack(byts)
