#!/usr/bin/env python3
"""Generate tests/golden/real_*.npz: packed per-CTU syntax of CONSECUTIVE pictures of the reference's bundled
HM16.0 streams (run once in the build container; needs /root/reference and oracle/_ref/hm_syntax_dump).

The syntax comes from the reference's own decoder hooks (TDecCu.cpp:153-289, TDecSlice.cpp:333-341) via
oracle/hm_dump; where the authors shipped golden stage-(a) maps (HEVCfiles/<video>/*.mat, main.m:52-55) the maps
of the same pictures are stored beside it.  The GPU tests run the whole path on these clips (north_star:
"Correctness uses the bundled HEVCfiles streams").
"""
import os, subprocess, sys
import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_stage_a import read_dump, pack, DUMPER, WORK, REF  # noqa: E402

CLIPS = {  # name: (w, h, fps, first picture, count, has golden .mat)
    "BasketballPass": (416, 240, 50.0, 0, 40, False),
    "BasketballDrill": (832, 480, 50.0, 0, 14, True),
    "BasketballDrive": (1920, 1080, 50.0, 126, 12, True),
}


def main():
    os.makedirs(WORK, exist_ok=True)
    for name, (w, h, fps, first, count, golden) in CLIPS.items():
        dump = os.path.join(WORK, name + ".bin")
        if not os.path.exists(dump):
            env = dict(os.environ, HM_SYNTAX_DUMP=dump)
            subprocess.check_call([DUMPER, "-b", f"{REF}/HEVCfiles/{name}/str.bin"], env=env, stdout=subprocess.DEVNULL)
        fr = read_dump(dump)
        sel = list(range(first, first + count))
        hdr, tok, base, byts = pack([fr[p] for p in sel])
        rec = dict(w=w, h=h, fps=fps, pocs=np.array(sel, np.int32), hdr=hdr, tok=tok, tok_base=base, ctu_bytes=byts)
        if golden:
            g = {k: sio.loadmat(f"{REF}/HEVCfiles/{name}/{k}.mat") for k in ("bitmap", "depth", "mv_x", "mv_y")}
            rec["mvx"] = np.stack([g["mv_x"]["video_mv_x"][0][p] for p in sel]).astype(np.int16)
            rec["mvy"] = np.stack([g["mv_y"]["video_mv_y"][0][p] for p in sel]).astype(np.int16)
            rec["depth"] = np.stack([g["depth"]["video_depth"][0][p] for p in sel]).astype(np.uint8)
            rec["bitmap"] = np.stack([g["bitmap"]["video_bitmap"][0][p] for p in sel]).astype(np.uint16)
        out = os.path.join(HERE, f"real_{name}.npz")
        np.savez_compressed(out, **rec)
        print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
