#!/usr/bin/env python3
"""Generate tests/golden/stage_a_*.npz  (run once in the build container; needs /root/reference).

Inputs : per-CTU syntax of the bundled HM16.0 streams, captured by the reference's own hooks
         (TDecCu.cpp:153-289, TDecSlice.cpp:333-341) through oracle/_ref/hm_syntax_dump
         (built by oracle/hm_dump/build_hm_dump.py from the reference decoder sources).
Outputs: the AUTHORS' golden maps HEVCfiles/<video>/{bitmap,depth,mv_x,mv_y}.mat (main.m:52-55)
         for the same frames.  These pin stage (a) (syntax -> 4x4 maps), SURVEY.md section 0.4.

Token packing (the repo's wire format, include/cds.h): per CTU hdr = (offset, n_cupu, n_pred, n_mv)
into an int16 token stream laid out [cupu | pred | mv] per CTU.
"""
import os, struct, subprocess, sys
import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
DUMPER = os.path.join(ROOT, "oracle", "_ref", "hm_syntax_dump")
WORK = os.environ.get("CDS_DUMP_DIR", "/tmp/dumps")

SETS = {  # video: (w, h, frames)
    "BasketballDrill": (832, 480, [0, 1, 2, 3, 4, 17, 64, 128, 250, 333, 498, 499]),
    "BasketballDrive": (1920, 1080, [0, 1, 2, 130, 140, 499]),
}


def read_dump(path):
    b = open(path, "rb").read(); off = 0; frames = {}
    while off < len(b):
        poc, nctu = struct.unpack_from("<ii", b, off); off += 8
        ctus = []
        for _ in range(nctu):
            by, nc, npd, nm = struct.unpack_from("<iiii", b, off); off += 16
            cu = np.frombuffer(b, "<i2", nc, off); off += 2 * nc
            pr = np.frombuffer(b, "<i2", npd, off); off += 2 * npd
            mv = np.frombuffer(b, "<i2", nm, off); off += 2 * nm
            ctus.append((by, cu, pr, mv))
        frames[poc] = ctus
    return frames


def pack(frames_ctus):
    hdr, toks, base, byts = [], [], [0], []
    for ctus in frames_ctus:
        h = np.zeros((len(ctus), 4), np.int32); off = 0; t = []
        for a, (by, cu, pr, mv) in enumerate(ctus):
            h[a] = (off, len(cu), len(pr), len(mv)); t += [cu, pr, mv]; off += len(cu) + len(pr) + len(mv)
        hdr.append(h); toks.append(np.concatenate(t).astype(np.int16)); base.append(base[-1] + off)
        byts.append(np.array([c[0] for c in ctus], np.uint16))
    return np.stack(hdr), np.concatenate(toks), np.array(base, np.int64), np.stack(byts)


def main():
    os.makedirs(WORK, exist_ok=True)
    for name, (w, h, sel) in SETS.items():
        dump = os.path.join(WORK, name + ".bin")
        if not os.path.exists(dump):
            env = dict(os.environ, HM_SYNTAX_DUMP=dump)
            subprocess.check_call([DUMPER, "-b", f"{REF}/HEVCfiles/{name}/str.bin"], env=env, stdout=subprocess.DEVNULL)
        fr = read_dump(dump)
        g = {k: sio.loadmat(f"{REF}/HEVCfiles/{name}/{k}.mat") for k in ("bitmap", "depth", "mv_x", "mv_y")}
        gb = g["bitmap"]["video_bitmap"][0]; gd = g["depth"]["video_depth"][0]
        gx = g["mv_x"]["video_mv_x"][0]; gy = g["mv_y"]["video_mv_y"][0]
        hdr, tok, base, byts = pack([fr[p] for p in sel])
        out = os.path.join(HERE, f"stage_a_{name}.npz")
        np.savez_compressed(
            out, w=w, h=h, pocs=np.array(sel, np.int32), hdr=hdr, tok=tok, tok_base=base, ctu_bytes=byts,
            mvx=np.stack([gx[p] for p in sel]).astype(np.int16), mvy=np.stack([gy[p] for p in sel]).astype(np.int16),
            depth=np.stack([gd[p] for p in sel]).astype(np.uint8), bitmap=np.stack([gb[p] for p in sel]).astype(np.uint16))
        print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
