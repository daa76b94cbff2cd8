#!/usr/bin/env python3
"""Generate tests/golden/semantic_BasketballDrill.npz: the 256 z-scan partitions per CTU (depth, prediction mode, motion vector) that
TComDataCU holds for BasketballDrill pictures 0..13, captured by the hooked reference decoder (oracle/_ref/hm_syntax_dump with
HM_SEMANTIC_DUMP set; patch in oracle/hm_dump/build_hm_dump.py).  The same pictures' reference stage-(a) maps (the authors' .mat
files) are in real_BasketballDrill.npz.  SURVEY 8(f).2.  Run once in the build container."""
import os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_stage_a import DUMPER, WORK, REF  # noqa: E402


def main():
    name, n = "BasketballDrill", 14
    os.makedirs(WORK, exist_ok=True)
    sem = os.path.join(WORK, name + ".sem")
    subprocess.check_call([DUMPER, "-b", f"{REF}/HEVCfiles/{name}/str.bin"], env=dict(os.environ, HM_SYNTAX_DUMP=os.path.join(WORK, name + ".bin"), HM_SEMANTIC_DUMP=sem),
                          stdout=subprocess.DEVNULL)
    raw = np.fromfile(sem, np.uint8)
    nctu = 104
    rec = 8 + nctu * (256 + 256 + 1024)
    assert raw.size % rec == 0
    raw = raw.reshape(-1, rec)
    poc = raw[:, :4].copy().view(np.int32).reshape(-1)
    order = np.argsort(poc)[:n]
    assert list(poc[order]) == list(range(n))
    body = raw[order, 8:].reshape(n, nctu, 1536)
    depth = body[:, :, :256].copy(); pred = body[:, :, 256:512].copy()
    mv = body[:, :, 512:].copy().view(np.int16).reshape(n, nctu, 256, 2)
    out = os.path.join(HERE, f"semantic_{name}.npz")
    np.savez_compressed(out, w=832, h=480, pocs=np.arange(n, dtype=np.int32), depth256=depth, pred256=pred, mv256=mv)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
