#!/usr/bin/env python3
"""Generate tests/golden/fixations_*.npz: the eye-tracking fixations of the reference's database for the pictures of the
real-stream fixtures (run once in the build container; needs /root/reference/video_database.mat).

video_database.mat is a MATLAB v7.3 (HDF5) file and this image has no HDF5 reader.  Its only large dataset is
`fixdata` (N x 6 doubles: subject, video index, start ms, duration ms, x, y — main.m:325-346 uses columns 2-6), stored as
deflate-compressed chunks of 6 x 1345 values; they are found by scanning for zlib streams that inflate to exactly that
size, in file order.  The video names are small UTF-16 datasets in cell order.  Both readings are validated below:
per-video maximum coordinates must fit the known picture size of every clip.

Per-picture assignment follows main.m:331-344: a fixation covers pictures ceil(start/T) .. ceil((start+duration)/T)
(1-based, T = 1000/fps), clipped to the clip; positions outside the picture are dropped.
"""
import os, re, sys, zlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DB = "/root/reference/video_database.mat"
SIZES = {"BasketballPass": (416, 240, 50.0), "BasketballDrill": (832, 480, 50.0), "BasketballDrive": (1920, 1080, 50.0)}


def read_database():
    b = open(DB, "rb").read()
    chunks, i = [], 0
    while i < len(b) - 2:
        if b[i] == 0x78 and b[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            try:
                d = zlib.decompressobj(); out = d.decompress(b[i:])
                if len(out) == 6 * 1345 * 8 and d.eof:
                    chunks.append(np.frombuffer(out, "<f8").reshape(6, 1345)); i += len(b[i:]) - len(d.unused_data); continue
            except zlib.error:
                pass
        i += 1
    fix = np.concatenate(chunks, axis=1).T
    fix = fix[fix[:, 1] > 0]
    names = [m.group().decode("utf-16le") for m in re.finditer(rb"(?:[A-Za-z0-9_]\x00){3,}", b[:20000])][:33]
    assert len(names) == 33 and names[0] == "Kimono" and names[-1] == "PeopleOnStreet", names
    return fix, names


def main():
    fix, names = read_database()
    # validation of the reading: every clip's fixations fit its picture
    for name, (w, h, _) in SIZES.items():
        f = fix[fix[:, 1] == names.index(name) + 1]
        inside = np.mean((f[:, 4] > 0) & (f[:, 4] <= w) & (f[:, 5] > 0) & (f[:, 5] <= h))
        assert len(f) > 300 and inside > 0.97, (name, len(f), inside)
    for name, (w, h, fps) in SIZES.items():
        clip = np.load(os.path.join(HERE, f"real_{name}.npz"))
        pocs = clip["pocs"]; nframes = 500
        f = fix[fix[:, 1] == names.index(name) + 1]
        T = 1000.0 / fps
        per = {int(p): [] for p in pocs}
        for row in f:
            x, y = row[4], row[5]
            if not (0 < x <= w and 0 < y <= h):
                continue
            s = max(1, int(np.ceil(row[2] / T))); e = min(nframes, int(np.ceil((row[2] + row[3]) / T)))
            for k in range(s, e + 1):                      # 1-based picture index k <-> poc k-1
                if k - 1 in per:
                    per[k - 1].append((x, y))
        xy = np.array([p for poc in pocs for p in per[int(poc)]], np.float32).reshape(-1, 2)
        off = np.cumsum([0] + [len(per[int(poc)]) for poc in pocs]).astype(np.int32)
        out = os.path.join(HERE, f"fixations_{name}.npz")
        np.savez_compressed(out, pocs=pocs, xy=xy, offsets=off, w=w, h=h, fps=fps, video_index=names.index(name) + 1)
        print(out, len(xy), "fixation points over", len(pocs), "pictures")


if __name__ == "__main__":
    sys.exit(main())
