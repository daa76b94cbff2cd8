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
