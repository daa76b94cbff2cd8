"""Many bitstreams -> one GPU-owning process ("multi-process decode to feed the GPUs").

Each video is decoded by its own producer PROCESS — the reference's hooked HM-16.0 decoder, parse-only (CDS_PARSE_ONLY=1), writing the
hooks' per-CTU syntax in the packed wire format of include/cds.h to an inherited pipe (HM_PACKED_FD; patch in
oracle/hm_dump/build_hm_dump.py).  This process owns the GPU: one clip context and one reader thread per video, which batches the
pictures and calls cds_clip_process (ctypes releases the GIL, every context has its own stream).  Hooked decoders that each own a
CUDA context time-slice the GPU instead (tools/multi_decode.py: ~520 pictures/s for 16 processes)."""
import os
import subprocess
import threading
import time

import numpy as np

from .saliency import SaliencyClip


def _read_exact(f, n):
    b = f.read(n)
    while b is not None and len(b) < n:
        more = f.read(n - len(b))
        if not more:
            return None if not b else b
        b += more
    return b


def read_packed_record(f):
    """One picture from the producer's pipe -> (poc, hdr [nctu][4] i32, ctu_bytes [nctu] u16, tok i16) or None at end of stream."""
    head = _read_exact(f, 12)
    if not head or len(head) < 12:
        return None
    poc, nctu, ntok = np.frombuffer(head, np.int32)
    body = _read_exact(f, int(nctu) * 18 + int(ntok) * 2)
    if body is None or len(body) < nctu * 18 + ntok * 2:
        raise IOError("truncated syntax record")
    hdr = np.frombuffer(body, np.int32, nctu * 4).reshape(nctu, 4)
    byts = np.frombuffer(body, np.uint16, nctu, offset=nctu * 16)
    tok = np.frombuffer(body, np.int16, ntok, offset=nctu * 18)
    return int(poc), hdr, byts, tok


class _Video(threading.Thread):
    def __init__(self, producer, stream, w, h, fps, clip_kwargs, batch, collect, parse_only):
        super().__init__(daemon=True)
        self.w, self.h, self.batch, self.collect = w, h, batch, collect
        self.clip = SaliencyClip(w, h, fps, max_batch=batch, **clip_kwargs)
        rfd, wfd = os.pipe()
        env = dict(os.environ, HM_PACKED_FD=str(wfd), CDS_PARSE_ONLY="1" if parse_only else "0")
        self.proc = subprocess.Popen([producer, "-b", stream], env=env, pass_fds=(wfd,), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.close(wfd)
        self.pipe = os.fdopen(rfd, "rb", buffering=1 << 20)
        self.pictures = 0
        self.maps, self.cams, self.error = [], [], None

    def _flush(self, first, recs):
        hdr = np.stack([r[1] for r in recs]); byts = np.stack([r[2] for r in recs])
        base = np.zeros(len(recs) + 1, np.int64); base[1:] = np.cumsum([r[3].size for r in recs])
        tok = np.concatenate([r[3] for r in recs]) if base[-1] else np.zeros(1, np.int16)
        maps, cam = self.clip.process(first, hdr, tok, base, byts)
        self.pictures += len(recs)
        if self.collect:
            self.maps.append(maps); self.cams.append(cam)

    def run(self):
        try:
            recs, first = [], 0
            while True:
                r = read_packed_record(self.pipe)
                if r is None:
                    break
                if r[0] != first + len(recs):
                    raise IOError(f"pictures out of order: got POC {r[0]}, expected {first + len(recs)} (low-delay streams only)")
                recs.append(r)
                if len(recs) == self.batch:
                    self._flush(first, recs); first += len(recs); recs = []
            if recs:
                self._flush(first, recs)
            self.proc.wait()
        except Exception as e:  # reported by run_streams
            self.error = e
            self.proc.kill()
        finally:
            self.clip.close()


def run_streams(producer, streams, fps=50.0, batch=32, collect=False, parse_only=True, **clip_kwargs):
    """streams: list of (path, width, height).  Decodes them concurrently (one producer process each) and runs every picture through
    a clip context of this process.  clip_kwargs go to SaliencyClip (fast=True, or model= / meanVec= / stdVec=, out_u8=...).
    Returns dict(pictures, seconds, pictures_per_s, maps / cams per stream when collect)."""
    t0 = time.perf_counter()
    vids = [_Video(producer, p, w, h, fps, clip_kwargs, batch, collect, parse_only) for (p, w, h) in streams]
    for v in vids:
        v.start()
    for v in vids:
        v.join()
    sec = time.perf_counter() - t0
    for v in vids:
        if v.error:
            raise v.error
    n = sum(v.pictures for v in vids)
    out = dict(pictures=n, seconds=sec, pictures_per_s=n / sec, per_stream=[v.pictures for v in vids])
    if collect:
        out["maps"] = [np.concatenate(v.maps) for v in vids]; out["cams"] = [np.concatenate(v.cams) for v in vids]
    return out
