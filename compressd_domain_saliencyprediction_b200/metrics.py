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
