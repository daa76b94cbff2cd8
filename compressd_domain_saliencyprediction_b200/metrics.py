"""Saliency evaluation against eye-tracking fixations: AUC-Judd, NSS and CC.

The reference extracts per-picture fixations (main.m:325-346) but ships no metric code;
north_star asks for "AUC/NSS matching to 3 decimals" between the reference path and this one, so the same
implementations are applied to the oracle's maps and to the GPU's maps.  Host-side evaluation tooling (numpy);
nothing here is on the hot path.

Fixations are (x, y) in the reference's 1-based pixel coordinates (x = column, y = row).
"""
import numpy as np


def _fix_indices(shape, xy):
    h, w = shape
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    c = np.clip(np.ceil(xy[:, 0]).astype(np.int64) - 1, 0, w - 1)
    r = np.clip(np.ceil(xy[:, 1]).astype(np.int64) - 1, 0, h - 1)
    return r, c


def auc_judd(sal, xy):
    """Area under the ROC curve with thresholds at the saliency values of the fixated pixels (Judd et al.).
    True positives: fixated pixels above the threshold; false positives: non-fixated pixels above it."""
    sal = np.asarray(sal, np.float64)
    r, c = _fix_indices(sal.shape, xy)
    if len(r) == 0:
        return float("nan")
    fixmap = np.zeros(sal.shape, bool); fixmap[r, c] = True
    s_fix = np.sort(sal[fixmap])[::-1]
    nfix, npx = len(s_fix), sal.size
    if nfix == npx:
        return float("nan")
    all_sorted = np.sort(sal.ravel())
    above = npx - np.searchsorted(all_sorted, s_fix, side="left")          # pixels with value >= threshold
    tp = np.concatenate([[0.0], np.arange(1, nfix + 1) / nfix, [1.0]])
    fp = np.concatenate([[0.0], (above - np.arange(1, nfix + 1)) / (npx - nfix), [1.0]])
    return float(np.trapezoid(tp, fp))


def nss(sal, xy):
    """Normalised scanpath saliency: mean of the z-scored map at the fixated pixels."""
    sal = np.asarray(sal, np.float64)
    r, c = _fix_indices(sal.shape, xy)
    sd = sal.std()
    if len(r) == 0 or sd == 0:
        return float("nan")
    return float(np.mean((sal[r, c] - sal.mean()) / sd))


def cc(sal, other):
    """Pearson correlation between two maps (e.g. a saliency map and a blurred fixation density)."""
    a = np.asarray(sal, np.float64).ravel(); b = np.asarray(other, np.float64).ravel()
    a = a - a.mean(); b = b - b.mean()
    d = np.sqrt((a * a).sum() * (b * b).sum())
    return float((a * b).sum() / d) if d > 0 else float("nan")


def evaluate_clip(maps, xy, offsets):
    """Mean AUC-Judd / NSS over the pictures that have fixations.  offsets[f]..offsets[f+1] index xy for picture f."""
    aucs, nsss = [], []
    for f in range(len(maps)):
        pts = xy[offsets[f]:offsets[f + 1]]
        if len(pts) == 0:
            continue
        a, n = auc_judd(maps[f], pts), nss(maps[f], pts)
        if np.isfinite(a) and np.isfinite(n):
            aucs.append(a); nsss.append(n)
    return (float(np.mean(aucs)) if aucs else float("nan")), (float(np.mean(nsss)) if nsss else float("nan")), len(aucs)


# ------------------------------------------------------------------------------------------------
# the reference's fixation database (video_database.mat, MATLAB v7.3; main.m:3-18, 325-346)
# ------------------------------------------------------------------------------------------------
def load_fixation_database(path):
    """video_database.mat (through the minimal HDF5 reader mat73.py) or the .npz extract of it (tests/golden/fixations_all.npz)
    -> {"names": [33 str], "size": [33][2] (w, h), "frames": [33], "fps": [33], "fixdata": [N][6]} with fixdata columns
    SubjectIndex, VideoIndex (1-based), Timestamp ms, GazeDuration ms, FixationPointX, FixationPointY."""
    if str(path).endswith(".npz"):
        d = np.load(path, allow_pickle=False)
        return {"names": [str(s) for s in d["names"]], "size": d["size"].astype(np.float64), "frames": d["frames"].astype(np.int64),
                "fps": d["fps"].astype(np.float64), "fixdata": d["fixdata"].astype(np.float64)}
    from . import mat73
    db = mat73.load(path)["video_database"]
    vi = db["videos_info"]
    return {"names": list(db["video_names_list"]), "size": np.asarray(vi["size"], np.float64).reshape(-1, 2),
            "frames": np.asarray(vi["frames"]).reshape(-1).astype(np.int64), "fps": np.asarray(vi["framerate_fps"], np.float64).reshape(-1),
            "fixdata": np.asarray(db["fixdata"], np.float64)}


def fixations_per_picture(db, name, pocs=None):
    """main.m:325-346 for one clip: a fixation covers pictures ceil(start / T) .. ceil((start + duration) / T) (1-based,
    T = 1000 / fps), clipped to the clip; positions outside the picture are dropped.
    -> (xy [M][2] float32 in picture order, offsets [len(pocs) + 1]) for the given 0-based pictures (default: all)."""
    k = db["names"].index(name)
    w, h = db["size"][k]
    nframes = int(db["frames"][k]); T = 1000.0 / float(db["fps"][k])
    pocs = np.arange(nframes) if pocs is None else np.asarray(pocs)
    per = {int(p): [] for p in pocs}
    rows = db["fixdata"][db["fixdata"][:, 1] == k + 1]
    for row in rows:
        x, y = row[4], row[5]
        if not (0 < x <= w and 0 < y <= h):
            continue
        s = max(1, int(np.ceil(row[2] / T))); e = min(nframes, int(np.ceil((row[2] + row[3]) / T)))
        for kk in range(s, e + 1):                             # 1-based picture index kk <-> poc kk - 1
            if kk - 1 in per:
                per[kk - 1].append((x, y))
    xy = np.array([p for poc in pocs for p in per[int(poc)]], np.float32).reshape(-1, 2)
    off = np.cumsum([0] + [len(per[int(poc)]) for poc in pocs]).astype(np.int32)
    return xy, off
