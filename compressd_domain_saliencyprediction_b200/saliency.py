"""Per-clip driver: the loop of main.m:113-305 (RBF-SVM or linear fusion) behind cds_ctx."""
import ctypes as C

import numpy as np

from ._lib import CdsError, check, lib
from .ops import SvmModel, nctu


class _Params(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("fps", C.c_double), ("use_rbf", C.c_int),
                ("model", C.c_void_p), ("mean_vec", C.c_double * 9), ("std_vec", C.c_double * 9),
                ("max_batch", C.c_int), ("out_u8", C.c_int)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class SaliencyClip:
    """Saliency maps for one video from packed per-CTU syntax.

    model / meanVec / stdVec are the contents of svmRBFmodel_all.mat (main.m:113); the file is absent
    from the reference checkout, so callers pass a stand-in."""

    def __init__(self, w, h, fps, model=None, meanVec=None, stdVec=None, use_rbf=True, max_batch=32, out_u8=False, fast=False):
        """fast=True selects the reference's FAST version (TDecGop.cpp:225-333 + salmat.cpp; no model needed), whose 8-bit maps
        are what its decoder writes to ../HMdata/<POC>.png."""
        self.w, self.h, self.fps = int(w), int(h), float(fps)
        self.out_u8 = bool(out_u8)
        self.model = model
        self.fast = bool(fast)
        p = _Params()
        p.w, p.h, p.fps, p.use_rbf = self.w, self.h, self.fps, (2 if fast else int(bool(use_rbf)))
        p.model = model.handle if model is not None else None
        mv = np.zeros(9) if meanVec is None else np.asarray(meanVec, np.float64).reshape(9)
        sv = np.ones(9) if stdVec is None else np.asarray(stdVec, np.float64).reshape(9)
        for i in range(9):
            p.mean_vec[i] = mv[i]; p.std_vec[i] = sv[i]
        p.max_batch, p.out_u8 = int(max_batch), int(self.out_u8)
        self._h = C.c_void_p()
        check(lib().cds_create(C.byref(p), C.byref(self._h)), "cds_create")
        self.max_batch = int(max_batch)
        if fast:      # cvRound (half to even), salmat.cpp:18-19
            self.PS = max(1, int(np.rint(self.fps * 0.3))); self.PT = max(1, int(np.rint(self.fps * 0.7)))
        else:         # main.m:119,129
            self.PS = int(np.floor(self.fps * 0.3 + 0.5)); self.PT = int(np.floor(self.PS * 7 + 0.5))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().cds_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seek(self, poc):
        check(lib().cds_seek(self._h, int(poc)), "cds_seek")

    def process(self, first_poc, hdr, tok, tok_base, ctu_bytes):
        """Host buffers in, host maps out: (maps [F][h][w] float32 or uint8, cam [F][2])."""
        hdr = np.ascontiguousarray(hdr, np.int32); tok = np.ascontiguousarray(tok, np.int16)
        tok_base = np.ascontiguousarray(tok_base, np.int64); byts = np.ascontiguousarray(ctu_bytes, np.uint16)
        F = hdr.shape[0]
        if hdr.shape != (F, nctu(self.w, self.h), 4) or byts.shape != (F, nctu(self.w, self.h)) or tok_base.size != F + 1:
            raise CdsError("process: packed syntax shapes do not match the picture size")
        maps = np.zeros((F, self.h, self.w), np.uint8 if self.out_u8 else np.float32)
        cam = np.zeros((F, 2), np.int32)
        check(lib().cds_clip_process(self._h, int(first_poc), F, _ptr(byts), _ptr(hdr), _ptr(tok), _ptr(tok_base), tok.size,
                                     _ptr(maps), _ptr(cam)), "cds_clip_process")
        return maps, cam

    def process_maps(self, first_poc, mvx, mvy, depth, ctu_bytes, context_only=False):
        """Stage-(a) maps in (the mv_x / mv_y / depth / bitmap arrays main.m:25-55 reads; see formats.py), saliency maps out."""
        mvx = np.ascontiguousarray(mvx, np.int16); mvy = np.ascontiguousarray(mvy, np.int16)
        depth = np.ascontiguousarray(depth, np.uint8); byts = np.ascontiguousarray(ctu_bytes, np.uint16)
        F = mvx.shape[0]; h4, w4 = self.h // 4, self.w // 4
        if mvx.shape != (F, h4, w4) or mvy.shape != mvx.shape or depth.shape != mvx.shape or byts.reshape(F, -1).shape[1] != nctu(self.w, self.h):
            raise CdsError("process_maps: map shapes do not match the picture size")
        maps = None if context_only else np.zeros((F, self.h, self.w), np.uint8 if self.out_u8 else np.float32)
        cam = np.zeros((F, 2), np.int32)
        check(lib().cds_clip_process_maps(self._h, int(first_poc), F, _ptr(mvx), _ptr(mvy), _ptr(depth), _ptr(byts),
                                          None if context_only else _ptr(maps), _ptr(cam)), "cds_clip_process_maps")
        return maps, cam

    def context(self, first_poc, hdr, tok, tok_base, ctu_bytes):
        """Feed pictures as temporal context only (frame-range sharding halo): no maps are produced."""
        hdr = np.ascontiguousarray(hdr, np.int32); tok = np.ascontiguousarray(tok, np.int16)
        tok_base = np.ascontiguousarray(tok_base, np.int64); byts = np.ascontiguousarray(ctu_bytes, np.uint16)
        F = hdr.shape[0]
        if hdr.shape != (F, nctu(self.w, self.h), 4) or byts.shape != (F, nctu(self.w, self.h)) or tok_base.size != F + 1:
            raise CdsError("context: packed syntax shapes do not match the picture size")
        check(lib().cds_clip_context(self._h, int(first_poc), F, _ptr(byts), _ptr(hdr), _ptr(tok), _ptr(tok_base), tok.size),
              "cds_clip_context")

    # hook-style surface (TDecCu.cpp:153-289 capture, TDecGop.cpp:270-332 per-slice tail): one picture per call
    def acquire_frame_buffers(self):
        """Pinned views (ctu_bytes[nctu] u16, hdr[nctu][4] i32, tok[capacity] i16) for the next picture: fill them in place and
        pass them to submit_frame_acquired(poc, ntok) — the zero-copy hand-off the decoder hook uses."""
        pb, ph, pt, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(lib().cds_acquire_frame_buffers(self._h, C.byref(pb), C.byref(ph), C.byref(pt), C.byref(cap)), "cds_acquire_frame_buffers")
        n = nctu(self.w, self.h)
        byts = np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_uint16)), (n,))
        hdr = np.ctypeslib.as_array(C.cast(ph, C.POINTER(C.c_int32)), (n, 4))
        tok = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_int16)), (cap.value,))
        self._acq = (pb, ph, pt)
        return byts, hdr, tok

    def submit_frame_acquired(self, poc, ntok):
        pb, ph, pt = self._acq
        check(lib().cds_submit_frame(self._h, int(poc), pb, ph, pt, int(ntok)), "cds_submit_frame")

    def submit_frame(self, poc, ctu_bytes, hdr, tok):
        hdr = np.ascontiguousarray(hdr, np.int32); tok = np.ascontiguousarray(tok, np.int16)
        byts = np.ascontiguousarray(ctu_bytes, np.uint16)
        check(lib().cds_submit_frame(self._h, int(poc), _ptr(byts), _ptr(hdr), _ptr(tok), tok.size), "cds_submit_frame")

    def flush(self):
        check(lib().cds_flush(self._h), "cds_flush")

    def get_map(self, poc):
        m = np.zeros((self.h, self.w), np.uint8 if self.out_u8 else np.float32)
        cam = np.zeros(2, np.int32)
        check(lib().cds_get_map(self._h, int(poc), _ptr(m), _ptr(cam)), "cds_get_map")
        return m, cam

    def debug(self, which, frame_in_batch, shape):
        a = np.zeros(shape, np.float32)
        check(lib().cds_debug_copy(self._h, which.encode(), int(frame_in_batch), _ptr(a), a.nbytes), "cds_debug_copy")
        return a
