"""AUC-Judd / NSS / CC (host-side evaluation tooling) on hand-checkable cases and on the fixation fixtures."""
import os

import numpy as np

from compressd_domain_saliencyprediction_b200 import metrics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_auc_perfect_chance_and_inverted():
    rng = np.random.default_rng(0)
    sal = rng.uniform(0, 1, (40, 60))
    top = np.argsort(sal.ravel())[-20:]
    xy = np.stack([top % 60 + 1, top // 60 + 1], axis=1)                 # fixations on the 20 most salient pixels
    assert metrics.auc_judd(sal, xy) > 0.999
    low = np.argsort(sal.ravel())[:20]
    assert metrics.auc_judd(sal, np.stack([low % 60 + 1, low // 60 + 1], axis=1)) < 0.05
    rnd = rng.integers(0, 2400, 600)
    assert abs(metrics.auc_judd(sal, np.stack([rnd % 60 + 1, rnd // 60 + 1], axis=1)) - 0.5) < 0.05


def test_nss_and_cc_known_values():
    sal = np.zeros((10, 10)); sal[2, 3] = 1.0
    z = (1.0 - sal.mean()) / sal.std()
    assert abs(metrics.nss(sal, [(4, 3)]) - z) < 1e-12                   # x = column 4, y = row 3 (1-based)
    assert abs(metrics.cc(sal, 3 * sal + 2) - 1.0) < 1e-12 and abs(metrics.cc(sal, -sal) + 1.0) < 1e-12


def test_fixation_fixtures_are_consistent_with_the_clips():
    for name in ("BasketballPass", "BasketballDrill", "BasketballDrive"):
        f = np.load(os.path.join(GOLD, f"fixations_{name}.npz")); c = np.load(os.path.join(GOLD, f"real_{name}.npz"))
        assert np.array_equal(f["pocs"], c["pocs"]) and int(f["w"]) == int(c["w"]) and int(f["h"]) == int(c["h"])
        xy, off = f["xy"], f["offsets"]
        assert off[-1] == len(xy) and len(off) == len(f["pocs"]) + 1
        assert np.all(xy[:, 0] > 0) and np.all(xy[:, 0] <= int(f["w"])) and np.all(xy[:, 1] > 0) and np.all(xy[:, 1] <= int(f["h"]))
        assert np.min(np.diff(off)) >= 10                                   # 32 observers: every picture has fixations
