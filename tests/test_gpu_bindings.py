"""The reference-side bindings compiled for real (bindings/): the computecontrast5 mex drop-in is called
through a minimal mex.h stand-in exactly as MATLAB would call it, next to the reference's own verbatim mexFunction."""
import ctypes as C
import os

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mexlib():
    import compressd_domain_saliencyprediction_b200 as pkg
    assert pkg.device_ok(), pkg.lib().cds_last_error()
    so = os.path.join(o.ORACLE_DIR, "_build", "libmex_computecontrast5_cds.so")
    if not os.path.exists(so):
        pytest.skip("mex binding not built (run __graft_entry__.build())")
    L = C.CDLL(so)
    L.mexcds_last_error.restype = C.c_char_p
    return L


def _call(L, pad, W, ln, win, m, n):
    pad = np.asfortranarray(pad, np.float64); W = np.asfortranarray(W, np.float64)
    out = np.zeros((m, n), np.float64, order="F")
    rc = L.mexcds_computecontrast5(o._ptr(pad), pad.shape[0], pad.shape[1], o._ptr(W), W.shape[0], ln, win, m, n, o._ptr(out))
    return rc, out


@pytest.mark.parametrize("case", [(60, 104, 11.0, 1, 48), (67, 120, 13.0, 2, 48), (17, 30, 3.0, 16, 80)])
def test_mex_dropin_matches_reference_mexfunction(mexlib, case):
    rows, cols, delta, cus, patch = case
    a = np.random.default_rng(rows).standard_normal((rows, cols))
    pad, W, ln, win = o.sudoku_inputs(a, delta, cus, patch)
    rc, got = _call(mexlib, pad, W, ln, win, rows, cols)
    assert rc == 0, mexlib.mexcds_last_error()
    want = o.ref_computecontrast5(pad, W, ln, win, rows, cols)          # the reference's own mexFunction, compiled verbatim
    if want is None:
        want = o.orc_computecontrast5(pad, W, ln, win, rows, cols)
    assert np.max(np.abs(got - want)) <= 1e-4 * np.max(np.abs(want))


def test_mex_dropin_reports_bad_arguments(mexlib):
    a = np.random.default_rng(0).uniform(0, 1, (20, 20))
    pad, W, ln, win = o.sudoku_inputs(a, 3.0, 1, 8)
    rc, _ = _call(mexlib, pad, W, ln, win, 25, 20)                       # m does not match the padded size
    assert rc != 0 and b"cds:computecontrast5" in mexlib.mexcds_last_error()
    rc, _ = _call(mexlib, pad, W[:-1, :-1], ln, win, 20, 20)             # W is not winsize x winsize
    assert rc != 0


# ---------------------------------------------------------------------------------------------------------------------------------
# libsvmpredict: bindings/libsvmpredict_cds.cpp next to the reference's own matlab/svmpredict.c + svm_model_matlab.c (verbatim, against
# the struct-capable mex.h stand-in oracle/shim_mx/mex.h).  Both are driven by oracle/ref_wrap/mx_harness.c, which builds the MATLAB
# arguments (label vector, dense instance matrix, the 11-field model struct with sparse SVs) and calls mexFunction.
_SIG = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
        C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]


def _predict(L, labels, X, SV, coef, rho, gamma, lab2, options=b""):
    labels = np.ascontiguousarray(labels, np.float64); X = np.asfortranarray(X, np.float64)
    SV = np.ascontiguousarray(SV, np.float64); coef = np.ascontiguousarray(coef, np.float64)
    N, F = X.shape
    out_l = np.zeros(N); acc = np.zeros(3); dec = np.zeros(N); msg = C.create_string_buffer(1024)
    L.mx_libsvmpredict.argtypes = _SIG
    rc = L.mx_libsvmpredict(o._ptr(labels), o._ptr(X), N, F, SV.shape[0], o._ptr(SV), o._ptr(coef), float(rho), float(gamma), int(lab2[0]), int(lab2[1]),
                            int((coef > 0).sum()), options, o._ptr(out_l), o._ptr(acc), o._ptr(dec), msg, 1024)
    return rc, out_l, acc, dec, msg.value.decode()


@pytest.fixture(scope="module")
def predictors():
    import compressd_domain_saliencyprediction_b200 as pkg
    assert pkg.device_ok(), pkg.lib().cds_last_error()
    mine = os.path.join(o.ORACLE_DIR, "_build", "libmex_libsvmpredict_cds.so")
    ref = os.path.join(o.ORACLE_DIR, "_ref", "libref_svmpredict.so")
    if not (os.path.exists(mine) and os.path.exists(ref)):
        pytest.skip("libsvmpredict bindings not built (needs the reference tree at build time)")
    return C.CDLL(mine), C.CDLL(ref)


def test_libsvmpredict_dropin_matches_reference_mexfunction_on_heart_scale(predictors):
    mine, ref = predictors
    d = np.load(os.path.join(o.ROOT, "tests", "golden", "svm_heart_scale.npz"))
    args = (d["y"], d["X"], d["SV"], d["coef"], float(d["rho"]), float(d["gamma"]), d["labels"])
    rc_r, lab_r, acc_r, dec_r, _ = _predict(ref, *args)
    rc_m, lab_m, acc_m, dec_m, msg = _predict(mine, *args)
    assert rc_r == 0 and rc_m == 0, msg
    assert abs(acc_r[0] - 86.66666667) < 1e-6                              # libsvm-3.17/matlab/README:179-181, 234/270
    assert np.array_equal(lab_m, lab_r) and np.allclose(acc_m, acc_r, rtol=1e-12, atol=0)
    assert np.max(np.abs(dec_m - dec_r)) <= 1e-4 * np.max(np.abs(dec_r))


def test_libsvmpredict_dropin_on_the_papers_instance_shape(predictors):
    """main.m:297: 40 000 x 9 z-scored instances, labels all zero (so accuracy is the share of label-0 predictions: none)."""
    mine, ref = predictors
    SV, coef, rho, gamma, labels = o.synth_model(600)
    X = np.random.default_rng(2).standard_normal((40000, 9)) * 0.7
    y = np.zeros(40000)
    rc_r, lab_r, acc_r, dec_r, _ = _predict(ref, y, X, SV, coef, rho, gamma, labels)
    rc_m, lab_m, acc_m, dec_m, msg = _predict(mine, y, X, SV, coef, rho, gamma, labels)
    assert rc_r == 0 and rc_m == 0, msg
    close = np.abs(dec_r) > 1e-4 * np.max(np.abs(dec_r))                   # labels may only differ where the decision value is ~0
    assert np.array_equal(lab_m[close], lab_r[close])
    assert np.max(np.abs(dec_m - dec_r)) <= 1e-4 * np.max(np.abs(dec_r))
    assert abs(acc_m[1] - acc_r[1]) <= 1e-3 * max(acc_r[1], 1e-9)


def test_libsvmpredict_dropin_error_style(predictors):
    """Like svmpredict.c: a printed message and empty outputs (fake_answer), no exception."""
    mine, ref = predictors
    SV, coef, rho, gamma, labels = o.synth_model(16)
    X = np.zeros((10, 9))
    for L in (ref, mine):
        rc, _, _, _, msg = _predict(L, np.zeros(10), X, SV, coef, rho, gamma, labels, options=b"-b 1")
        assert rc == -1 and msg                                            # the reference: model has no probability information
        rc, _, _, _, msg = _predict(L, np.zeros(7), X[:7], SV, coef, rho, gamma, labels)       # sanity: a valid call after an error still works
        assert rc == 0
