"""GPU parity tests for stage (c): camera-motion detection / removal (main.m:77-105, cm_judge.m,
mv_vote.m, direction_cluster.m, magnitude_vote.m).  Votes and camera motion must be bit-exact
(north_star: "Integer feature maps and MV votes must be bit-exact"); the normalised FP32 maps are
held to 1e-4 relative."""
import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


@pytest.fixture(scope="module")
def cds():
    import compressd_domain_saliencyprediction_b200 as pkg
    assert pkg.device_ok(), pkg.lib().cds_last_error()
    return pkg


def _frame(cds, w, h, seed, poc):
    hdr, tok, byts = cds.ops.synth_frame(w, h, seed, poc)
    mx, my, dp = o.orc_getframe(w, h, hdr, tok)
    return mx, my, dp, byts


@pytest.mark.parametrize("wh", [(416, 240), (832, 480), (1920, 1080), (72, 136)])
def test_camera_motion_vs_oracle(cds, wh):
    w, h = wh
    moving = 0
    for poc in range(6):
        mx, my, dp, byts = _frame(cds, w, h, 1234, poc)
        got = cds.ops.camera_motion(w, h, mx, my, byts, dp)
        outs, cam, ave, flag = o.orc_camera_motion(w, h, mx, my, byts, dp)
        assert got["flag"] == flag, (wh, poc)
        assert np.array_equal(got["cam"], cam), (wh, poc, got["cam"], cam)
        assert got["ave"][0] == ave[0] and got["ave"][1] == ave[1], (wh, poc, got["ave"], ave)   # exact doubles
        moving += flag
        for name, want in zip(("Vx", "Vy", "mv_norm", "bit_norm", "depth_norm"), outs):
            scale = max(np.max(np.abs(want)), 1e-30)
            assert np.max(np.abs(got[name] - want)) <= REL_TOL * scale, (wh, poc, name)
    assert moving >= 3      # the synthetic pan must exercise the vote


def test_camera_motion_static_scene_keeps_unfiltered_field(cds):
    """cm_judge: more than half of the background vectors within sqrt(2) px -> no compensation, and the
    UNFILTERED field is kept (main.m:93-96)."""
    w, h = 416, 240
    rng = np.random.default_rng(0)
    mx = rng.integers(-3, 4, (h // 4, w // 4)).astype(np.int16)
    my = rng.integers(-3, 4, (h // 4, w // 4)).astype(np.int16)
    mx[10:20, 30:50] = 57
    dp = rng.integers(0, 4, (h // 4, w // 4)).astype(np.uint8)
    byts = rng.integers(0, 300, cds.ops.nctu(w, h)).astype(np.uint16)
    got = cds.ops.camera_motion(w, h, mx, my, byts, dp)
    outs, cam, ave, flag = o.orc_camera_motion(w, h, mx, my, byts, dp)
    assert flag == 0 and got["flag"] == 0 and tuple(got["cam"]) == (0, 0)
    assert np.array_equal(got["Vx"], (mx / 4.0).astype(np.float32))
    assert np.max(np.abs(got["mv_norm"] - outs[2])) <= REL_TOL


def test_camera_motion_all_zero_inputs(cds):
    w, h = 128, 64
    z16 = np.zeros((h // 4, w // 4), np.int16); z8 = np.zeros((h // 4, w // 4), np.uint8)
    byts = np.zeros(cds.ops.nctu(w, h), np.uint16)
    got = cds.ops.camera_motion(w, h, z16, z16, byts, z8)
    assert got["flag"] == 0 and np.all(got["mv_norm"] == 0) and np.all(got["bit_norm"] == 0) and np.all(got["depth_norm"] == 0)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 50000])
@pytest.mark.parametrize("seed", [0, 1])
def test_mv_vote_bit_exact_vs_oracle(cds, n, seed):
    rng = np.random.default_rng(seed * 100 + n)
    k = rng.integers(-400, 401, (2, n))
    if seed:                                  # a dominant direction with heavy ties, plus axis-aligned vectors
        k[0, : n // 2] = 24; k[1, : n // 2] = rng.integers(-2, 3, n // 2)
        k[:, n // 2: n // 2 + 3] = 0
    vx, vy = k[0] / 8.0, k[1] / 8.0
    xa, ya = cds.ops.mv_vote(vx, vy)
    wx, wy = o.orc_direction_cluster(vx, vy)
    assert xa == wx and ya == wy, (n, seed, (xa, ya), (wx, wy))


def test_mv_vote_direction_bin_boundaries(cds):
    """Vectors exactly on the 16 bin boundaries (multiples of pi/8 that are representable: axes and
    diagonals) must land where the reference's double atan2 arithmetic puts them."""
    pts = [(8, 0), (8, 8), (0, 8), (-8, 8), (-8, 0), (-8, -8), (0, -8), (8, -8)]
    for (x, y) in pts:
        vx = np.array([x, x, 3 * x + 1, 1], float) / 8.0
        vy = np.array([y, y, 3 * y, -7], float) / 8.0
        assert cds.ops.mv_vote(vx, vy) == o.orc_direction_cluster(vx, vy), (x, y)


@pytest.mark.parametrize("n", [1, 5, 400, 20000])
def test_magnitude_vote_vs_oracle(cds, n):
    v = np.random.default_rng(n).gamma(2.0, 3.0, n)
    got = cds.ops.magnitude_vote(v)
    want = o.orc_magnitude_vote(v)
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want))


def test_mv_vote_rejects_non_eighth_pel(cds):
    with pytest.raises(cds.CdsError):
        cds.ops.mv_vote([0.3], [0.1])
