#!/usr/bin/env python3
"""Generate the fixation fixtures from the reference's eye-tracking database (run once in the build container; needs
/root/reference/video_database.mat, a MATLAB v7.3 = HDF5 file, read with the package's own minimal HDF5 reader mat73.py):

  tests/golden/fixations_all.npz       all 33 clips: names, picture sizes, frame counts, frame rates and the 26 914 x 6 fixation
                                       table (main.m:325-346 uses columns 2-6) — what the AUC / NSS tooling needs for any clip
  tests/golden/fixations_<clip>.npz    per-picture fixations of the three real-stream fixtures (main.m:331-344 assignment)
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from compressd_domain_saliencyprediction_b200 import metrics  # noqa: E402

DB = "/root/reference/video_database.mat"
CLIPS = ["BasketballPass", "BasketballDrill", "BasketballDrive"]


def main():
    db = metrics.load_fixation_database(DB)
    assert len(db["names"]) == 33 and db["names"][0] == "Kimono" and db["names"][-1] == "PeopleOnStreet"
    for k, name in enumerate(db["names"]):                     # the reading is sane: every clip's fixations fit its picture
        f = db["fixdata"][db["fixdata"][:, 1] == k + 1]
        w, h = db["size"][k]
        inside = np.mean((f[:, 4] > 0) & (f[:, 4] <= w) & (f[:, 5] > 0) & (f[:, 5] <= h))
        assert len(f) > 300 and inside > 0.9, (name, len(f), inside)
    np.savez_compressed(os.path.join(HERE, "fixations_all.npz"), names=np.array(db["names"]), size=db["size"].astype(np.int32),
                        frames=db["frames"].astype(np.int32), fps=db["fps"].astype(np.float32), fixdata=db["fixdata"].astype(np.float32))
    print("fixations_all.npz:", len(db["names"]), "clips,", len(db["fixdata"]), "fixations")
    for name in CLIPS:
        clip = np.load(os.path.join(HERE, f"real_{name}.npz"))
        k = db["names"].index(name)
        xy, off = metrics.fixations_per_picture(db, name, clip["pocs"])
        out = os.path.join(HERE, f"fixations_{name}.npz")
        np.savez_compressed(out, pocs=clip["pocs"], xy=xy, offsets=off, w=int(db["size"][k][0]), h=int(db["size"][k][1]),
                            fps=float(db["fps"][k]), video_index=k + 1)
        print(out, len(xy), "fixation points over", len(clip["pocs"]), "pictures")


if __name__ == "__main__":
    sys.exit(main())
