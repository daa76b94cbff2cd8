"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI of
libcds_b200.so and is checked against the CPU oracle / golden fixtures.  Integer stages must be
bit-exact; FP32 stages within 1e-4 relative (north_star tolerance), written per test."""
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


# ----------------------------------------------------------------------------------------- (a)
@pytest.mark.parametrize("name", ["BasketballDrill", "BasketballDrive"])
def test_getframe_bit_exact_vs_authors_golden(cds, name):
    d = np.load(f"{o.ROOT}/tests/golden/stage_a_{name}.npz")
    w, h = int(d["w"]), int(d["h"])
    for f in range(len(d["pocs"])):
        tok = d["tok"][d["tok_base"][f]:d["tok_base"][f + 1]]
        mx, my, dp = cds.ops.getframe(w, h, d["hdr"][f], tok)
        assert np.array_equal(mx, d["mvx"][f]), (name, int(d["pocs"][f]))
        assert np.array_equal(my, d["mvy"][f])
        assert np.array_equal(dp, d["depth"][f])


@pytest.mark.parametrize("wh", [(416, 240), (832, 480), (1280, 800), (1920, 1080), (2560, 1600), (3840, 2160), (64, 64), (72, 136)])
@pytest.mark.parametrize("adv", [0, 1])
def test_getframe_bit_exact_vs_oracle_fuzz(cds, wh, adv):
    w, h = wh
    for poc in (0, 1, 2, 9):
        hdr, tok, _ = cds.ops.synth_frame(w, h, 4321 + (adv << 60), poc)
        got = cds.ops.getframe(w, h, hdr, tok)
        want = o.orc_getframe(w, h, hdr, tok)
        for g, x in zip(got, want):
            assert np.array_equal(g, x), (wh, adv, poc)


def test_getframe_reference_array_interface(cds):
    """getFrame(TDecCu::MV, TDecCu::Pred, TDecCu::CUPU, ...) with the reference's 999-terminated float arrays."""
    w, h = 832, 480
    hdr, tok, _ = cds.ops.synth_frame(w, h, 5, 3)
    MV, PR, CU = o.unpack_to_ref_arrays(hdr, tok)
    mx, my, dp = cds.ops.getFrame(MV, PR, CU, w, h)
    want = o.orc_getframe(w, h, hdr, tok)
    assert np.array_equal(mx, want[0]) and np.array_equal(my, want[1]) and np.array_equal(dp, want[2])


def test_getframe_rejects_malformed_stream(cds):
    w, h = 128, 64
    hdr, tok, _ = cds.ops.synth_frame(w, h, 5, 3)
    bad = tok.copy(); bad[hdr[0, 0]] = 99; bad[hdr[0, 0] + 1:hdr[0, 0] + hdr[0, 1]] = 99   # endless splits
    with pytest.raises(cds.CdsError):
        cds.ops.getframe(w, h, hdr, bad)


# ----------------------------------------------------------------------------------------- (b)
def _rand_map(rng, rows, cols, signed):
    a = rng.standard_normal((rows, cols)) if signed else rng.uniform(0, 1, (rows, cols))
    a[rng.uniform(size=a.shape) < 0.1] = a.min() if signed else 0.0
    return a


@pytest.mark.parametrize("case", [
    (60, 104, 11.0, 1, 48, True),     # mv component map of a 416x240 frame: win 48, zeros inside (signed map)
    (60, 104, 11.0, 1, 48, False),    # non-negative map: zeros only in the padding
    (67, 45, 13.0, 2, 48, False),     # depth: win 24
    (17, 30, 3.0, 16, 80, False),     # bit: win 5 (generic kernel)
    (33, 50, 7.0, 1, 31, True),       # odd window (generic kernel)
])
def test_computecontrast5_mex_signature_vs_oracle(cds, case):
    rows, cols, delta, cusize, patch, signed = case
    rng = np.random.default_rng(11)
    a = _rand_map(rng, rows, cols, signed)
    pad, W, ln, win = o.sudoku_inputs(a, delta, cusize, patch)
    got = cds.ops.computecontrast5(pad, W, ln, win, rows, cols)
    want = o.orc_computecontrast5(pad, W, ln, win, rows, cols)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= REL_TOL * np.max(np.abs(want))
    big = want > 1e-3 * want.max()
    assert np.max(np.abs(got[big] - want[big]) / want[big]) <= REL_TOL


def test_computecontrast5_non_separable_weights(cds):
    rng = np.random.default_rng(2)
    rows, cols, win, ln = 20, 26, 24, 12
    pad = np.pad(rng.uniform(0.1, 1, (rows, cols)), ln)
    W = rng.uniform(0, 1, (win, win))
    got = cds.ops.computecontrast5(pad, W, ln, win, rows, cols)
    want = o.orc_computecontrast5(pad, W, ln, win, rows, cols)
    assert np.max(np.abs(got - want)) <= REL_TOL * np.max(np.abs(want))


def test_computecontrast5_constant_map_is_exactly_zero(cds):
    """A flat map gives contrast 0 everywhere in the reference (-> NaN after /max, main.m:237)."""
    a = np.full((40, 70), 0.25)
    pad, W, ln, win = o.sudoku_inputs(a, 11.0, 1, 48)
    got = cds.ops.computecontrast5(pad, W, ln, win, 40, 70)
    assert np.all(got == 0.0)


def test_computecontrast5_rejects_bad_arguments(cds):
    with pytest.raises(cds.CdsError):
        cds.ops.computecontrast5(np.zeros((10, 10)), np.ones((5, 5)), 2, 5, 9, 9)


# ----------------------------------------------------------------------------------------- (d)
@pytest.mark.parametrize("nsv", [1, 255, 1024, 4096])
def test_svm_rbf_decision_values_vs_oracle(cds, nsv):
    SV, coef, rho, gamma, labels = o.synth_model(nsv)
    X = np.random.default_rng(3).standard_normal((2003, 9)) * 1.5
    m = cds.ops.SvmModel(SVs=SV, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)
    assert m.totalSV == nsv and m.nfeat == 9
    dec, lab = m.decision_values(X)
    want, wlab = o.orc_svm_predict(SV, coef, rho, gamma, labels, X)
    scale = np.max(np.abs(want + rho)) + abs(rho)
    assert np.max(np.abs(dec - want)) <= REL_TOL * scale
    sure = np.abs(want) > 10 * REL_TOL * scale
    assert np.array_equal(lab[sure], wlab[sure])


def test_svm_heart_scale_known_answer(cds):
    d = np.load(f"{o.ROOT}/tests/golden/svm_heart_scale.npz")
    m = cds.ops.SvmModel(SVs=d["SV"], sv_coef=d["coef"], rho=float(d["rho"]), gamma=float(d["gamma"]), Label=d["labels"])
    lab, acc, dec = cds.ops.libsvmpredict(d["y"], d["X"], m)
    assert lab.shape == (270, 1) and dec.shape == (270, 1) and acc.shape == (3, 1)
    assert abs(acc[0, 0] - 86.6667) < 1e-3                       # Accuracy = 86.6667% (234/270)
    assert np.max(np.abs(dec[:, 0] - d["dec_ref"])) <= REL_TOL * np.max(np.abs(d["dec_ref"]))


def test_svm_text_model_loader(cds):
    d = np.load(f"{o.ROOT}/tests/golden/svm_heart_scale.npz")
    m = cds.ops.SvmModel(path=f"{o.ROOT}/tests/golden/svm_heart_scale.model", Label=d["labels"])
    assert m.totalSV == 130 and m.nfeat == 13
    dec, _ = m.decision_values(d["X"])
    assert np.max(np.abs(dec - d["dec_ref"])) <= REL_TOL * np.max(np.abs(d["dec_ref"]))
