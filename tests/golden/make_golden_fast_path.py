#!/usr/bin/env python3
"""Generate tests/golden/fast_BasketballDrill.npz: the reference's ONLY whole-path golden outputs.

HM_16.0_features/bin/HMdata/0..38.png are the saliency maps the authors' fast-version decoder (TDecGop.cpp:225-333) wrote for
the first 39 pictures of BasketballDrill (832x480, TDecCu.h:41-45: FRAMERATE 50).  The fixture stores them next to the packed
per-CTU syntax of the same pictures, captured by the reference's own decoder hooks (oracle/_ref/hm_syntax_dump).
Run once in the build container (needs /root/reference and `make -C oracle ref`).
"""
import os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_stage_a import read_dump, pack, DUMPER, WORK, REF  # noqa: E402
import cv2  # noqa: E402


def main():
    os.makedirs(WORK, exist_ok=True)
    name, w, h, fps, n = "BasketballDrill", 832, 480, 50.0, 39
    dump = os.path.join(WORK, name + ".bin")
    if not os.path.exists(dump):
        subprocess.check_call([DUMPER, "-b", f"{REF}/HEVCfiles/{name}/str.bin"], env=dict(os.environ, HM_SYNTAX_DUMP=dump), stdout=subprocess.DEVNULL)
    fr = read_dump(dump)
    hdr, tok, base, byts = pack([fr[p] for p in range(n)])
    png = np.stack([cv2.imread(f"{REF}/HM_16.0_features/bin/HMdata/{p}.png", cv2.IMREAD_UNCHANGED) for p in range(n)])
    assert png.shape == (n, h, w) and png.dtype == np.uint8
    out = os.path.join(HERE, f"fast_{name}.npz")
    np.savez_compressed(out, w=w, h=h, fps=fps, hdr=hdr, tok=tok, tok_base=base, ctu_bytes=byts, png=png)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
