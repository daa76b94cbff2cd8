#!/usr/bin/env python3
"""Generate tests/golden/standin_svmRBFmodel.{model,npz}: a deterministic stand-in for the reference's missing
svmRBFmodel_all.mat (main.m:113: model, meanVec, stdVec), trained with the reference's OWN libsvm-3.17 svm-train
(oracle/_ref/svm-train) on the reference's OWN eye-tracking fixations.  Run once in the build container.

Samples: the nine 200x200 HEVC features (main.m:275-292, computed by the CPU oracle) of the three real-stream fixture clips;
label +1 for grid cells within 6 cells of a fixation of that picture, -1 for cells at least 25 cells away from every
fixation; 1000 + 1000 samples drawn with a fixed seed.  meanVec / stdVec are the sample statistics (main.m:293-294 z-scores
with them); C-SVC, RBF, gamma = 1/9, C = 1 (libsvm defaults for 9 features).
It is a stand-in, trained and evaluated on the same few pictures: it makes the SVM stage responsive to the real features
(unlike a random model), nothing more.
"""
import os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import _oracle as o  # noqa: E402

TRAIN = os.path.join(ROOT, "oracle", "_ref", "svm-train")


def features(name):
    d = np.load(os.path.join(HERE, f"real_{name}.npz")); fx = np.load(os.path.join(HERE, f"fixations_{name}.npz"))
    w, h, F = int(d["w"]), int(d["h"]), len(d["pocs"])
    ints = [o.orc_getframe(w, h, d["hdr"][f], d["tok"][d["tok_base"][f]:d["tok_base"][f + 1]]) for f in range(F)]
    mvx = np.stack([m[0] for m in ints]); mvy = np.stack([m[1] for m in ints]); dep = np.stack([m[2] for m in ints])
    dummy = o.synth_model(4)
    first = max(0, F - 12)
    o.oracle().orc_set_threads(8)
    _, _, feat = o.orc_process_clip(w, h, 10.0, mvx, mvy, dep, d["ctu_bytes"], dummy, np.zeros(9), np.ones(9), first_out=first, use_rbf=1, want_feat=True)
    feat = np.nan_to_num(feat)                                    # [F-first][9][200][200]
    X, y = [], []
    yy, xx = np.mgrid[0:200, 0:200]
    for k in range(feat.shape[0]):
        pts = fx["xy"][fx["offsets"][first + k]:fx["offsets"][first + k + 1]]
        if len(pts) == 0:
            continue
        gx = (pts[:, 0] - 0.5) * 200.0 / w; gy = (pts[:, 1] - 0.5) * 200.0 / h
        d2 = np.min((xx[None] - gx[:, None, None]) ** 2 + (yy[None] - gy[:, None, None]) ** 2, axis=0)
        f9 = feat[k].reshape(9, -1).T
        pos = np.flatnonzero(d2.ravel() <= 36); neg = np.flatnonzero(d2.ravel() >= 625)
        X += [f9[pos], f9[neg]]; y += [np.ones(len(pos)), -np.ones(len(neg))]
    return np.concatenate(X), np.concatenate(y)


def main():
    Xs, ys = zip(*[features(n) for n in ("BasketballPass", "BasketballDrill", "BasketballDrive")])
    X, y = np.concatenate(Xs), np.concatenate(ys)
    rng = np.random.default_rng(2017)
    pick = np.concatenate([rng.choice(np.flatnonzero(y > 0), 1000, replace=False), rng.choice(np.flatnonzero(y < 0), 1000, replace=False)])
    X, y = X[pick], y[pick]
    mean, std = X.mean(0), X.std(0) + 1e-9
    Z = (X - mean) / std
    tr = "/tmp/standin_train.txt"
    with open(tr, "w") as f:
        for zi, yi in zip(Z, y):
            f.write(f"{int(yi):+d} " + " ".join(f"{j + 1}:{v:.8g}" for j, v in enumerate(zi)) + "\n")
    out = os.path.join(HERE, "standin_svmRBFmodel.model")
    subprocess.check_call([TRAIN, "-s", "0", "-t", "2", "-g", str(1.0 / 9), "-c", "1", "-q", tr, out])
    np.savez(os.path.join(HERE, "standin_svmRBFmodel.npz"), meanVec=mean, stdVec=std)
    nsv = [l for l in open(out) if l.startswith("total_sv")][0].split()[1]
    print(out, os.path.getsize(out), "bytes, total_sv", nsv, "; meanVec", np.round(mean, 3), "stdVec", np.round(std, 3))


if __name__ == "__main__":
    sys.exit(main())
