"""CPU-only: pin the oracle (oracle/cds_oracle.c) against the reference's golden vectors and, where
oracle/_ref exists (the reference's own sources compiled in the build container), against the
reference code itself."""
import os

import numpy as np
import pytest

import _oracle as o

GOLDEN = ["BasketballDrill", "BasketballDrive"]


@pytest.mark.parametrize("name", GOLDEN)
def test_getframe_oracle_matches_authors_golden_mat(name):
    d = np.load(f"{o.ROOT}/tests/golden/stage_a_{name}.npz")
    w, h = int(d["w"]), int(d["h"])
    for f in range(len(d["pocs"])):
        tok = d["tok"][d["tok_base"][f]:d["tok_base"][f + 1]]
        mx, my, dp = o.orc_getframe(w, h, d["hdr"][f], tok)
        assert np.array_equal(mx, d["mvx"][f]), (name, int(d["pocs"][f]))
        assert np.array_equal(my, d["mvy"][f])
        assert np.array_equal(dp, d["depth"][f])
        assert np.array_equal(d["ctu_bytes"][f].reshape(d["bitmap"][f].shape), d["bitmap"][f])


@pytest.mark.parametrize("name", GOLDEN)
def test_reference_getframe_matches_golden(name):
    d = np.load(f"{o.ROOT}/tests/golden/stage_a_{name}.npz")
    w, h = int(d["w"]), int(d["h"])
    if o.ref(f"getframe_{w}x{h}") is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for f in range(min(3, len(d["pocs"]))):
        tok = d["tok"][d["tok_base"][f]:d["tok_base"][f + 1]]
        mx, my, dp = o.ref_getframe(w, h, d["hdr"][f], tok)
        assert np.array_equal(mx, d["mvx"][f]) and np.array_equal(my, d["mvy"][f]) and np.array_equal(dp, d["depth"][f])


@pytest.mark.parametrize("wh", [(416, 240), (832, 480), (1280, 800), (1920, 1080)])
@pytest.mark.parametrize("adv", [0, 1])
def test_getframe_oracle_vs_verbatim_reference_fuzz(wh, adv):
    w, h = wh
    if o.ref(f"getframe_{w}x{h}") is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from compressd_domain_saliencyprediction_b200 import ops
    for poc in (0, 1, 5):
        hdr, tok, _ = ops.synth_frame(w, h, 99 + (adv << 60), poc)
        mx, my, dp = o.orc_getframe(w, h, hdr, tok)
        r = o.ref_getframe(w, h, hdr, tok)
        assert np.array_equal(mx, r[0]) and np.array_equal(my, r[1]) and np.array_equal(dp, r[2])


def _rand_map(rng, rows, cols, signed):
    a = rng.standard_normal((rows, cols)) if signed else rng.uniform(0, 1, (rows, cols))
    a[rng.uniform(size=a.shape) < 0.1] = a.min() if signed else 0.0   # ties at the minimum -> exact zeros after pm_norm
    return a


@pytest.mark.parametrize("case", [(23, 31, 11.0, 1, 48, True), (20, 17, 13.0, 2, 48, False), (9, 12, 3.0, 16, 80, False)])
def test_contrast_oracle_vs_verbatim_reference(case):
    rows, cols, delta, cusize, patch, signed = case
    if o.ref("contrast") is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    a = _rand_map(rng, rows, cols, signed)
    pad, W, ln, win = o.sudoku_inputs(a, delta, cusize, patch)
    # MATLAB view: the map is m x n = rows x cols
    got = o.orc_computecontrast5(pad, W, ln, win, rows, cols)
    want = o.ref_computecontrast5(pad, W, ln, win, rows, cols)
    assert np.array_equal(got, want)      # same arithmetic, same order -> identical doubles
    # the row-major Sudoku restatement must agree with the mex-level composition
    sud = o.orc_sudoku(a, delta, cusize, patch)
    np.testing.assert_allclose(sud, want / want.max(), rtol=1e-12, atol=0)


def test_contrast_constant_map_gives_nan_like_reference():
    a = np.zeros((12, 14))
    with np.errstate(all="ignore"):
        out = o.orc_cu_contrast5(a, 1, 11.0, 48)
    assert np.isnan(out).all()          # 0/0, then main.m:237-238 turns NaN into 0


def test_cu_contrast5_block_mean_and_replication():
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, (21, 27))
    out = o.orc_cu_contrast5(a, 4, 3.0, 20)
    assert out.shape == a.shape and abs(out.max() - 1.0) < 1e-12
    assert np.all(out[0:4, 0:4] == out[0, 0]) and np.all(out[20:21, 24:27] == out[20, 24])


@pytest.mark.parametrize("nsv,nfeat", [(300, 9), (130, 13)])
def test_svm_oracle_vs_reference_libsvm(nsv, nfeat):
    if o.ref("svm") is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    SV, coef, rho, gamma, labels = o.synth_model(nsv, nfeat)
    SV[::7, 2] = 0.0                       # zeros are dropped from the sparse SV rows like MATLAB does
    X = np.random.default_rng(1).standard_normal((500, nfeat))
    dec, lab = o.orc_svm_predict(SV, coef, rho, gamma, labels, X)
    rdec, rlab = o.ref_svm_predict(SV, coef, rho, gamma, labels, X)
    assert np.array_equal(dec, rdec) and np.array_equal(lab, rlab)


def test_svm_heart_scale_known_answer():
    """libsvm-3.17/matlab/README:179-181: svm-train -c 1 -g 0.07 heart_scale -> 86.6667 % (234/270)."""
    d = np.load(f"{o.ROOT}/tests/golden/svm_heart_scale.npz")
    dec, lab = o.orc_svm_predict(d["SV"], d["coef"], float(d["rho"]), float(d["gamma"]), d["labels"], d["X"])
    assert d["SV"].shape[0] == 130
    assert int((lab == d["y"]).sum()) == 234
    assert np.array_equal(dec, d["dec_ref"])


@pytest.mark.parametrize("name", ["BasketballDrill", "BasketballDrive"])
def test_oracle_getframe_on_consecutive_real_pictures_vs_authors_golden(name):
    """tests/golden/real_*.npz: consecutive pictures of the bundled streams with the authors' .mat maps."""
    d = np.load(os.path.join(o.ROOT, "tests", "golden", f"real_{name}.npz"))
    w, h = int(d["w"]), int(d["h"])
    for f in range(len(d["pocs"])):
        got = o.orc_getframe(w, h, d["hdr"][f], d["tok"][d["tok_base"][f]:d["tok_base"][f + 1]])
        assert np.array_equal(got[0], d["mvx"][f]) and np.array_equal(got[1], d["mvy"][f]) and np.array_equal(got[2], d["depth"][f]), f


def test_oracle_camera_motion_fires_on_basketballdrive_pan():
    d = np.load(os.path.join(o.ROOT, "tests", "golden", "real_BasketballDrive.npz"))
    w, h = int(d["w"]), int(d["h"])
    for f in range(4):
        outs, cam, ave, flag = o.orc_camera_motion(w, h, d["mvx"][f], d["mvy"][f], d["ctu_bytes"][f], d["depth"][f])
        assert flag == 1 and cam[0] >= 6 and cam[1] == 0


def test_clip_through_reference_kernels_matches_the_port():
    """bench.py --impl reference routes contrast and SVM through the reference's own compiled code (oracle/_ref); the maps
    must agree with the restatements (same algorithm, same double arithmetic up to summation order)."""
    import compressd_domain_saliencyprediction_b200 as cds
    if o.ref("contrast") is None or o.ref("svm") is None:
        pytest.skip("oracle/_ref not built")
    w, h, F, fps = 416, 240, 4, 10.0
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 31, F)
    ints = [o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]]) for f in range(F)]
    mvx = np.stack([m[0] for m in ints]); mvy = np.stack([m[1] for m in ints]); dep = np.stack([m[2] for m in ints])
    SV, coef, rho, gamma, labels = o.synth_model(64)
    SV = SV * 0.5 + 1.0
    mdl = (SV, coef, rho, gamma, labels)
    mean9, std9 = np.full(9, 0.3), np.full(9, 0.2)
    o.oracle().orc_set_threads(4)
    a, acam, _ = o.orc_process_clip(w, h, fps, mvx, mvy, dep, byts, mdl, mean9, std9, first_out=F - 1)
    assert o.use_reference_kernels(mdl)
    try:
        b, bcam, _ = o.orc_process_clip(w, h, fps, mvx, mvy, dep, byts, mdl, mean9, std9, first_out=F - 1)
    finally:
        o.use_reference_kernels(None)
    assert np.array_equal(acam, bcam)
    assert np.max(np.abs(np.nan_to_num(a) - np.nan_to_num(b))) <= 1e-9


def test_oracle_getframe_on_every_bundled_stream_vs_reference_getframe():
    """tests/golden/all_streams.npz: two pictures of each of the 24 bundled HM16.0 streams whose geometry the reference handles
    (the 1280x720 clips are excluded: its CTU-row formula is wrong there), with the maps of the verbatim reference getFrame."""
    d = np.load(os.path.join(o.ROOT, "tests", "golden", "all_streams.npz"))
    assert len(d["names"]) == 24
    for k, name in enumerate(d["names"]):
        w, h = int(d[f"w{k}"]), int(d[f"h{k}"])
        for i in range(2):
            got = o.orc_getframe(w, h, d[f"hdr{k}"][i], d[f"tok{k}"][d[f"base{k}"][i]:d[f"base{k}"][i + 1]])
            assert np.array_equal(got[0], d[f"mvx{k}"][i]) and np.array_equal(got[1], d[f"mvy{k}"][i]) and np.array_equal(got[2], d[f"dep{k}"][i]), (name, i)


def test_standin_trained_model_fixture_loads_and_matches_reference_libsvm():
    """tests/golden/standin_svmRBFmodel.model: trained by the reference's own svm-train on the reference's fixations
    (make_golden_standin_model.py).  The text parser, the oracle's decision values and libsvm-3.17's agree."""
    g = os.path.join(o.ROOT, "tests", "golden")
    SV, coef, rho, gamma, labels = o.load_libsvm_model(os.path.join(g, "standin_svmRBFmodel.model"))
    st = np.load(os.path.join(g, "standin_svmRBFmodel.npz"))
    assert SV.shape[1] == 9 and SV.shape[0] > 100 and abs(gamma - 1 / 9) < 1e-5 and labels == (1, -1)
    assert st["meanVec"].shape == (9,) and np.all(st["stdVec"] > 0)
    assert np.all(np.abs(coef) <= 1.0 + 1e-12)                       # C = 1 box constraint
    X = np.random.default_rng(5).standard_normal((400, 9))
    dec, lab = o.orc_svm_predict(SV, coef, rho, gamma, labels, X)
    assert 0 < int((lab > 0).sum()) < 400
    r = o.ref_svm_predict(SV, coef, rho, gamma, labels, X, nsv0=int((coef > 0).sum()))
    if r is not None:
        assert np.array_equal(dec, r[0]) and np.array_equal(lab, r[1])


def test_parse_only_hm_producer_captures_identical_syntax(tmp_path):
    """the reference's hooked decoder with CDS_PARSE_ONLY=1 (reconstruction, deblocking and SAO skipped; patch applied by
    oracle/hm_dump/build_hm_dump.py to a scratch copy) hands the hooks byte-identical per-CTU syntax.  All 32 bundled streams were
    compared in the build container (profiles/r01_hm_parse_only.txt); this runs the one small stream that travels with oracle/_ref."""
    import filecmp, subprocess
    dumper = os.path.join(o.ROOT, "oracle", "_ref", "hm_syntax_dump")
    stream = os.path.join(o.ROOT, "oracle", "_ref", "streams", "BasketballPass.bin")
    if not (os.path.exists(dumper) and os.path.exists(stream)):
        pytest.skip("hooked HM decoder / bitstream not built (needs the reference tree at build time)")
    a, b = str(tmp_path / "full.bin"), str(tmp_path / "parse.bin")
    for out, po in ((a, "0"), (b, "1")):
        r = subprocess.run([dumper, "-b", stream], env=dict(os.environ, HM_SYNTAX_DUMP=out, CDS_PARSE_ONLY=po), stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=300)
        assert r.returncode == 0
    assert os.path.getsize(a) > 1000000 and filecmp.cmp(a, b, shallow=False)


def test_reference_libsvmpredict_mexfunction_through_the_mex_standin():
    """The reference's own matlab/svmpredict.c + svm_model_matlab.c, compiled verbatim against oracle/shim_mx/mex.h (struct arrays, sparse
    SVs, transpose call-back), reproduce libsvm's heart_scale known answer (matlab/README:179-181) and the oracle's decision values:
    this pins the stand-in the GPU binding test (tests/test_gpu_bindings.py) relies on."""
    import ctypes as C
    so = os.path.join(o.ORACLE_DIR, "_ref", "libref_svmpredict.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    L = C.CDLL(so)
    d = np.load(os.path.join(o.ROOT, "tests", "golden", "svm_heart_scale.npz"))
    SV = np.ascontiguousarray(d["SV"], np.float64); coef = np.ascontiguousarray(d["coef"], np.float64)
    X = np.asfortranarray(d["X"], np.float64); y = np.ascontiguousarray(d["y"], np.float64)
    N, F = X.shape
    lab = np.zeros(N); acc = np.zeros(3); dec = np.zeros(N); msg = C.create_string_buffer(1024)
    L.mx_libsvmpredict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    rc = L.mx_libsvmpredict(o._ptr(y), o._ptr(X), N, F, SV.shape[0], o._ptr(SV), o._ptr(coef), float(d["rho"]), float(d["gamma"]), int(d["labels"][0]),
                            int(d["labels"][1]), int((coef > 0).sum()), b"", o._ptr(lab), o._ptr(acc), o._ptr(dec), msg, 1024)
    assert rc == 0, msg.value
    assert int((lab == y).sum()) == 234 and abs(acc[0] - 100.0 * 234 / 270) < 1e-9
    assert np.array_equal(dec, d["dec_ref"])
