#!/usr/bin/env python3
"""Generate tests/golden/all_streams.npz: two pictures (an early P picture and a late one) of EVERY bundled HM16.0 stream
(HEVCfiles/*/str.bin) as packed per-CTU syntax, with the maps the reference's own getFrame (code_mat_h.h, compiled verbatim
under oracle/_ref) produces for them.  Run once in the build container (needs /root/reference and oracle/_ref).

The eight 1280x720 clips are left out: the reference's CTU-row formula int(h/64.0+0.5) (code_mat_h.h:20-21) gives 11 rows
for 720 lines instead of 12, so its own output is not a usable reference there."""
import os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden_stage_a import read_dump, pack, DUMPER, WORK, REF  # noqa: E402
import _oracle as o  # noqa: E402
import cv2  # noqa: E402


def main():
    os.makedirs(WORK, exist_ok=True)
    rec, names = {}, []
    for name in sorted(os.listdir(f"{REF}/HEVCfiles")):
        stream = f"{REF}/HEVCfiles/{name}/str.bin"
        if not os.path.exists(stream):
            continue
        cap = cv2.VideoCapture(f"{REF}/videos/{name}.avi")
        w, h = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        if w == 0 or h == 720 or o.ref(f"getframe_{w}x{h}") is None:
            print("skip", name, w, h); continue
        dump = os.path.join(WORK, name + ".bin")
        if not os.path.exists(dump):
            subprocess.check_call([DUMPER, "-b", stream], env=dict(os.environ, HM_SYNTAX_DUMP=dump), stdout=subprocess.DEVNULL)
        fr = read_dump(dump)
        sel = [3, max(fr) - 2]
        hdr, tok, base, byts = pack([fr[p] for p in sel])
        maps = [o.ref_getframe(w, h, hdr[i], tok[base[i]:base[i + 1]]) for i in range(2)]
        k = len(names); names.append(name)
        rec[f"w{k}"] = w; rec[f"h{k}"] = h; rec[f"hdr{k}"] = hdr; rec[f"tok{k}"] = tok; rec[f"base{k}"] = base; rec[f"bytes{k}"] = byts
        rec[f"mvx{k}"] = np.stack([m[0] for m in maps]).astype(np.int16); rec[f"mvy{k}"] = np.stack([m[1] for m in maps]).astype(np.int16)
        rec[f"dep{k}"] = np.stack([m[2] for m in maps]).astype(np.uint8)
        print(name, w, h, "pocs", sel, "tokens", tok.size)
    out = os.path.join(HERE, "all_streams.npz")
    np.savez_compressed(out, names=np.array(names), **rec)
    print(out, os.path.getsize(out), "bytes,", len(names), "streams")


if __name__ == "__main__":
    sys.exit(main())
