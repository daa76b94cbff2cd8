"""SURVEY 8(f).2, "semantic" stage (a): the CTU's 256 z-scan partitions (TComDataCU::getDepth / getPredictionMode / getCUMvField, captured
from the reference decoder: tests/golden/semantic_BasketballDrill.npz) transposed to the h/4 x w/4 maps.  Depth must equal the authors'
golden maps exactly; the motion maps are an intentional deviation from getFrame (its stream-aliasing quirks are not reproduced) and
differ in a few percent of the cells."""
import os

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(o.ROOT, "tests", "golden")


def _morton_table():
    z = np.zeros((16, 16), np.int64)
    for r in range(16):
        for c in range(16):
            v = 0
            for i in range(4):
                v |= ((c >> i) & 1) << (2 * i) | ((r >> i) & 1) << (2 * i + 1)
            z[r, c] = v
    return z


def _transpose_ref(w, h, depth256, pred256, mv256):
    """numpy restatement of the z-scan -> raster transpose (g_auiRasterToZscan of 4x4 units in a 64x64 CTU is the Morton order)"""
    h4, w4, ncol = h // 4, w // 4, (w + 63) // 64
    rr, cc = np.mgrid[0:h4, 0:w4]
    ctu = (rr >> 4) * ncol + (cc >> 4)
    z = _morton_table()[rr & 15, cc & 15]
    dep = depth256[:, ctu, z]
    inter = pred256[:, ctu, z] == 0
    mvx = np.where(inter, mv256[:, ctu, z, 0], 0).astype(np.int16); mvy = np.where(inter, mv256[:, ctu, z, 1], 0).astype(np.int16)
    return mvx, mvy, dep


def test_semantic_scatter_matches_the_transpose_and_the_golden_depth():
    import compressd_domain_saliencyprediction_b200 as cds
    assert cds.device_ok()
    s = np.load(os.path.join(GOLD, "semantic_BasketballDrill.npz")); r = np.load(os.path.join(GOLD, "real_BasketballDrill.npz"))
    w, h = int(s["w"]), int(s["h"])
    mvx, mvy, dep = cds.ops.getframe_semantic(w, h, s["depth256"], s["pred256"], s["mv256"])
    wx, wy, wd = _transpose_ref(w, h, s["depth256"], s["pred256"], s["mv256"])
    assert np.array_equal(mvx, wx) and np.array_equal(mvy, wy) and np.array_equal(dep, wd)
    assert np.array_equal(dep, r["depth"])                               # identical to the reference's depth maps (authors' .mat)
    dx, dy = float((mvx != r["mvx"]).mean()), float((mvy != r["mvy"]).mean())
    print(f"semantic vs reference getFrame over 14 pictures: mv_x differs in {100 * dx:.2f} % of the cells, mv_y in {100 * dy:.2f} %")
    assert dx < 0.06 and dy < 0.06
    assert np.array_equal(mvx[0], r["mvx"][0]) and np.array_equal(mvy[0], r["mvy"][0])     # the intra picture: all zero either way


def test_semantic_maps_feed_the_clip_pipeline():
    """The composition a maintainer would use: semantic hook -> cds_getframe_semantic -> cds_clip_process_maps (here the fast version)."""
    import compressd_domain_saliencyprediction_b200 as cds
    s = np.load(os.path.join(GOLD, "semantic_BasketballDrill.npz")); r = np.load(os.path.join(GOLD, "real_BasketballDrill.npz"))
    w, h = int(s["w"]), int(s["h"])
    mvx, mvy, dep = cds.ops.getframe_semantic(w, h, s["depth256"], s["pred256"], s["mv256"])
    clip = cds.SaliencyClip(w, h, 50.0, fast=True, out_u8=True, max_batch=8)
    got, cam = clip.process_maps(0, mvx, mvy, dep, r["ctu_bytes"])
    clip.close()
    fo = o.FastOracle(w, h, 50.0)
    want = np.stack([fo.frame(mvx[f], mvy[f], dep[f], r["ctu_bytes"][f])[0] for f in range(mvx.shape[0])])
    fo.close()
    assert np.array_equal(got, want)
    ref_clip = cds.SaliencyClip(w, h, 50.0, fast=True, out_u8=True, max_batch=8)
    ref_maps, _ = ref_clip.process(0, r["hdr"], r["tok"], r["tok_base"], r["ctu_bytes"])
    ref_clip.close()
    d = np.abs(got.astype(np.int16) - ref_maps.astype(np.int16))
    print(f"saliency maps from semantic vs reference stage (a): mean |diff| {d.mean():.3f} LSB, max {int(d.max())}")
    assert d.mean() < 3.0
