"""Second opinion for the oracle's restated MATLAB / Image Processing Toolbox primitives (SURVEY 8c rows marked
"restatement only"): every helper of oracle/cds_oracle.c that follows .m text is compared here with an independent
numpy / scipy / torch evaluation of the same documented semantics (tests/matlab_semantics.py), and the whole normal
version (main.m:77-305, both fusion branches) is run twice — once by the C oracle, once by the numpy transcription —
on the same clip.  No GPU."""
import ctypes as C

import numpy as np
import pytest

import _oracle as o
import matlab_semantics as ms


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _orc2(fn, a, *args, out_shape=None):
    a = np.ascontiguousarray(a, np.float64)
    out = np.zeros(out_shape if out_shape else a.shape)
    getattr(o.oracle(), fn)(_p(a), a.shape[0], a.shape[1], *args, _p(out))
    return out


def test_round_is_half_away_from_zero():
    assert list(ms.m_round([0.5, -0.5, 1.5, -1.5, 2.4999, -2.5, 0.0])) == [1, -1, 2, -2, 2, -3, 0]


@pytest.mark.parametrize("mx", [-3, -1, 0, 2, 5])
@pytest.mark.parametrize("my", [-4, -1, 0, 1, 3])
def test_immove_follows_the_m_file_for_every_sign(mx, my):
    """immove.m:19 measures temp (before the row padding), so a negative mov_y leaves the rows where they are."""
    rng = np.random.default_rng(100 + mx * 10 + my)
    im = rng.standard_normal((9, 13))
    want = ms.immove(im, mx, my)
    got = _orc2("orc_immove", im, mx, my)
    assert np.array_equal(got, want)
    if my < 0:
        assert np.array_equal(got, ms.immove(im, mx, 0))          # the quirk itself
    if my > 0 and mx == 0:
        assert np.array_equal(got[my:], im[:-my]) and not got[:my].any()


def test_medfilt2_even_window_bruteforce():
    rng = np.random.default_rng(1)
    for shape in [(7, 9), (1, 5), (6, 1), (16, 16)]:
        a = rng.integers(-40, 40, shape) / 4.0
        assert np.array_equal(_orc2("orc_medfilt2_2x2", a), ms.medfilt2_2x2(a))


def test_scipy_correlate_is_imfilter_definition():
    """the scipy call used as the second opinion really is 'zero-padded correlation centred at floor((k+1)/2)'"""
    rng = np.random.default_rng(2)
    a = rng.standard_normal((12, 15))
    for k in (4, 5, 8):
        h = rng.standard_normal((k, k))
        assert np.allclose(ms.imfilter(a, h), ms.imfilter_bruteforce(a, h), atol=1e-12)


@pytest.mark.parametrize("k,sigma,shape", [(32, 8.0, (40, 56)), (9, 2.0, (20, 31)), (64, 16.0, (96, 72)), (13, 3.0, (30, 30))])
def test_imfilter_gaussian_vs_scipy_2d(k, sigma, shape):
    a = np.random.default_rng(k).standard_normal(shape)
    want = ms.imfilter(a, ms.fspecial_gaussian(k, sigma))
    got = _orc2("orc_imfilter_gauss", a, k, C.c_double(sigma))
    assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("h", [240, 480, 768, 800, 1080, 1600, 2160])
def test_fspecial_truncation_never_triggers_for_the_path_kernels(h):
    """fspecial zeroes entries < eps*max; at k = round(h/7.5), sigma = round(h/30) the corner is exp(-4) of the centre,
    so the 2-D kernel stays the outer product of the normalised 1-D factor the oracle and the CUDA path use."""
    k = int(ms.m_round(h / 7.5)); s = float(ms.m_round(h / 30))
    G = ms.fspecial_gaussian(k, s)
    assert np.count_nonzero(G == 0) == 0
    g = np.zeros(k)
    o.oracle().orc_gauss_kernel(k, C.c_double(s), _p(g))
    assert np.max(np.abs(np.outer(g, g) - G)) < 1e-15
    assert abs(g.sum() - 1) < 1e-14


@pytest.mark.parametrize("shape,out", [((270, 480), (200, 200)), ((240, 416), (200, 200)), ((200, 200), (240, 416)),
                                       ((200, 200), (1080, 1920)), ((1080, 1920), (200, 200)), ((96, 128), (200, 200))])
def test_imresize_vs_contributions_and_torch(shape, out):
    a = np.random.default_rng(shape[0]).standard_normal(shape)
    got = _orc2("orc_imresize_bilinear", a, out[0], out[1], out_shape=out)
    want = ms.imresize_bilinear(a, out[0], out[1])
    assert np.max(np.abs(got - want)) < 1e-12
    # independent implementation: identical away from the border (torch clips the window there, MATLAB mirrors it)
    t = ms.imresize_torch(a, out[0], out[1])
    mr = int(np.ceil(max(1.0, shape[0] / out[0]) * out[0] / shape[0])) + 2
    mc = int(np.ceil(max(1.0, shape[1] / out[1]) * out[1] / shape[1])) + 2
    assert np.max(np.abs(got[mr:-mr, mc:-mc] - t[mr:-mr, mc:-mc])) < 1e-9


def test_pm_norm_mysterythrehold_distmatrix():
    rng = np.random.default_rng(5)
    a = rng.random((30, 52))
    b = a.copy(); o.oracle().orc_pm_norm(_p(b), b.size)
    assert np.allclose(b, ms.pm_norm(a), atol=1e-15)
    for hw in [(120, 208), (30, 52)]:
        b = a.copy(); o.oracle().orc_mysterythrehold(_p(b), b.size, hw[0], hw[1])
        assert np.allclose(b, ms.mysterythrehold(a, *hw), atol=1e-14)
    d = np.zeros((50, 70)); o.oracle().orc_distmatrix(50, 70, _p(d))
    assert np.array_equal(d, ms.getdistMatrix(50, 70))


def test_direction_cluster_vs_literal():
    rng = np.random.default_rng(6)
    for n in (17, 400, 5000, 40, 9):
        vx = rng.integers(-60, 60, n) / 8.0 + 3; vy = rng.integers(-60, 60, n) / 8.0 - 1
        want = ms.direction_cluster(vx, vy)
        if np.isnan(want[0]):
            continue                                   # a winning bin with a single member, see the next test
        assert o.orc_direction_cluster(vx, vy) == want
    # ties between bins: the first maximal bin wins
    vx = np.array([1.0, 1.0, -1.0, -1.0]); vy = np.array([0.5, 0.5, -0.5, -0.5])
    assert o.orc_direction_cluster(vx, vy) == ms.direction_cluster(vx, vy)


def test_direction_cluster_single_member_bin_is_a_documented_deviation():
    """Len = 1 -> Lencut = 0 -> direction_cluster.m:43/49 evaluates sum([])/0 = NaN and every later map of the clip turns
    NaN.  The oracle and the CUDA path return the member itself instead (known limits)."""
    with np.errstate(invalid="ignore"):
        assert np.isnan(ms.direction_cluster(np.array([2.5]), np.array([-1.0]))[0])
    assert o.orc_direction_cluster(np.array([2.5]), np.array([-1.0])) == (2.5, -1.0)


def _random_clip(rng, w, h, F, pan):
    h4, w4 = h // 4, w // 4
    nrow, ncol = (h + 63) // 64, (w + 63) // 64
    mvx = np.zeros((F, h4, w4), np.int16); mvy = np.zeros_like(mvx)
    dep = rng.integers(0, 4, (F, h4, w4)).astype(np.uint8)
    byt = rng.integers(5, 400, (F, nrow, ncol)).astype(np.uint16)
    for f in range(F):
        px, py = pan[f % len(pan)]
        blk = rng.integers(-6, 7, (2, (h4 + 3) // 4, (w4 + 3) // 4))
        mvx[f] = np.kron(blk[0], np.ones((4, 4), np.int64))[:h4, :w4] + px
        mvy[f] = np.kron(blk[1], np.ones((4, 4), np.int64))[:h4, :w4] + py
    return mvx, mvy, dep, byt


def test_camera_motion_vs_literal():
    rng = np.random.default_rng(7)
    w, h = 192, 128
    for pan in [(0, 0), (37, -22), (-18, 9), (5, 5)]:
        mvx, mvy, dep, byt = _random_clip(rng, w, h, 1, [pan])
        outs, cam, ave, flag = o.orc_camera_motion(w, h, mvx[0], mvy[0], byt[0].ravel(), dep[0])
        Vx, Vy, mn, bn, dn, cx, cy = ms.camera_motion(mvx[0], mvy[0], byt[0], dep[0], w, h)
        assert (cam[0], cam[1]) == (cx, cy)
        for got, want in zip(outs, (Vx, Vy, mn, bn, dn)):
            assert np.array_equal(got, want)
    assert flag in (0, 1)


@pytest.mark.parametrize("cusize,delta,patch", [(16, 3.0, 80), (2, 13.0, 48), (1, 11.0, 48)])
def test_cu_contrast5_vs_literal(cusize, delta, patch):
    rng = np.random.default_rng(cusize)
    a = rng.random((30, 52)) - (0.4 if cusize == 1 else 0.0)          # the mv maps are signed
    got = o.orc_cu_contrast5(a, cusize, delta, patch)
    want = ms.cu_contrast5(a, cusize, delta, patch)
    assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("use_rbf", [1, 0])
def test_whole_normal_version_second_opinion(use_rbf):
    """main.m:77-305 twice: C oracle vs the numpy / scipy transcription, on a clip whose camera pans in every
    direction (so that medfilt2, the vote and immove's four branches are all exercised)."""
    rng = np.random.default_rng(11)
    w, h, F, fps = 128, 96, 7, 10.0
    pan = [(0, 0), (33, 21), (35, -26), (-29, -31), (-30, 18), (41, -37), (0, 0)]
    mvx, mvy, dep, byt = _random_clip(rng, w, h, F, pan)
    SV, coef, rho, gamma, labels = o.synth_model(48)
    SV = SV * 0.5 + 1.0
    mean9, std9 = np.full(9, 0.3), np.full(9, 0.2)
    o.use_reference_kernels(None)
    maps, cam, feat = o.orc_process_clip(w, h, fps, mvx, mvy, dep, byt.reshape(F, -1), (SV, coef, rho, gamma, labels), mean9, std9,
                                         use_rbf=use_rbf, want_feat=True)
    wmaps, wcam, wfeat = ms.process_clip(w, h, fps, mvx, mvy, dep, byt, (SV, coef, rho, gamma, labels), mean9, std9,
                                         use_rbf=use_rbf, want_feat=bool(use_rbf))
    assert np.array_equal(cam, wcam)
    assert np.any(cam[:, 1] < 0) and np.any(cam[:, 1] > 0) and np.any(cam[:, 0] < 0)
    if use_rbf:
        assert np.max(np.abs(np.nan_to_num(feat) - np.nan_to_num(wfeat))) < 1e-9
    assert np.max(np.abs(np.nan_to_num(maps) - np.nan_to_num(wmaps))) < 1e-8
