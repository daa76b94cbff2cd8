"""Test-side loader for the CPU oracle (oracle/cds_oracle.c) and, when present, the reference's own
code compiled under oracle/_ref.  TEST INFRASTRUCTURE: never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_ORC = None
_REF = {}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle():
    global _ORC
    if _ORC is None:
        so = os.path.join(ORACLE_DIR, "_build", "libcds_oracle.so")
        src = os.path.join(ORACLE_DIR, "cds_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        L = C.CDLL(so)
        if hasattr(L, "orc_magnitude_vote"):
            L.orc_magnitude_vote.restype = C.c_double
        _ORC = L
    return _ORC


def ref(name):
    """oracle/_ref/libref_<name>.so or None (built only where /root/reference exists)."""
    if name not in _REF:
        so = os.path.join(ORACLE_DIR, "_ref", f"libref_{name}.so")
        _REF[name] = C.CDLL(so) if os.path.exists(so) else None
    return _REF[name]


# ---------------------------------------------------------------------------- stage (a)
def orc_getframe(w, h, hdr, tok):
    hdr = np.ascontiguousarray(hdr, np.int32); tok = np.ascontiguousarray(tok, np.int16)
    mvx = np.zeros((h // 4, w // 4), np.int16); mvy = np.zeros_like(mvx); dp = np.zeros((h // 4, w // 4), np.uint8)
    rc = oracle().orc_getframe(w, h, _ptr(hdr), _ptr(tok), _ptr(mvx), _ptr(mvy), _ptr(dp))
    assert rc == 0, rc
    return mvx, mvy, dp


def unpack_to_ref_arrays(hdr, tok):
    """packed tokens -> the reference's float[nctu][1000] arrays with 999 terminators."""
    n = hdr.shape[0]
    MV = np.zeros((n, 1000), np.float32); PR = np.zeros((n, 1000), np.float32); CU = np.zeros((n, 1000), np.float32)
    for a in range(n):
        o, nc, npd, nm = (int(v) for v in hdr[a])
        CU[a, :nc] = tok[o:o + nc]; CU[a, nc] = 999
        PR[a, :npd] = tok[o + nc:o + nc + npd]; PR[a, npd] = 999
        MV[a, :nm] = tok[o + nc + npd:o + nc + npd + nm]; MV[a, nm] = 999
    return MV, PR, CU


def ref_getframe(w, h, hdr, tok):
    L = ref(f"getframe_{w}x{h}")
    if L is None:
        return None
    MV, PR, CU = unpack_to_ref_arrays(hdr, tok)
    mx = np.zeros((h // 4, w // 4), np.float32); my = np.zeros_like(mx); dp = np.zeros_like(mx)
    L.ref_getframe(_ptr(MV), _ptr(PR), _ptr(CU), _ptr(mx), _ptr(my), _ptr(dp))
    return mx, my, dp


# ---------------------------------------------------------------------------- stage (b)
def orc_computecontrast5(pad, W, len_, win, m, n):
    pad = np.asfortranarray(pad, np.float64); W = np.asfortranarray(W, np.float64)
    out = np.zeros((m, n), np.float64, order="F")
    oracle().orc_computecontrast5(_ptr(pad), pad.shape[0], pad.shape[1], _ptr(W), len_, win, m, n, _ptr(out))
    return out


def ref_computecontrast5(pad, W, len_, win, m, n):
    L = ref("contrast")
    if L is None:
        return None
    pad = np.asfortranarray(pad, np.float64); W = np.asfortranarray(W, np.float64)
    out = np.zeros((m, n), np.float64, order="F")
    L.ref_computecontrast5(_ptr(pad), pad.shape[0], pad.shape[1], _ptr(W), len_, win, m, n, _ptr(out))
    return out


def orc_sudoku(a, delta, cusize, patchsize):
    a = np.ascontiguousarray(a, np.float64); out = np.zeros_like(a)
    oracle().orc_sudoku_contrast(_ptr(a), a.shape[0], a.shape[1], C.c_double(delta), C.c_double(cusize),
                                 C.c_double(patchsize), _ptr(out))
    return out


def orc_cu_contrast5(a, cusize, delta, patchsize):
    a = np.ascontiguousarray(a, np.float64); out = np.zeros_like(a)
    oracle().orc_cu_contrast5(_ptr(a), a.shape[0], a.shape[1], int(cusize), C.c_double(delta),
                              C.c_double(patchsize), _ptr(out))
    return out


def sudoku_inputs(a, delta, cusize, patchsize):
    """(pad, W, len, win) exactly as Sudoku_contrastcpp5.m:6-34 builds them (numpy float64)."""
    win = int(np.floor(patchsize / cusize + 0.5)); ln = win // 2
    i = np.arange(1, win + 1)
    X, Y = np.meshgrid(i, i)
    R = np.sqrt((X - (ln + 1.0)) ** 2 + (Y - (ln + 1.0)) ** 2)
    W = np.exp(-R ** 2 / (2 * delta ** 2))
    p = np.pad(np.asarray(a, np.float64) + 0.00000001, ln)
    p = p - p.min(); p = p / p.max()
    return p, W, ln, win


# ---------------------------------------------------------------------------- stage (d)
def orc_svm_predict(SV, coef, rho, gamma, labels, X):
    SV = np.ascontiguousarray(SV, np.float64); coef = np.ascontiguousarray(coef, np.float64)
    X = np.asfortranarray(X, np.float64); N = X.shape[0]
    dec = np.zeros(N); lab = np.zeros(N)
    oracle().orc_svm_rbf_predict(SV.shape[0], SV.shape[1], _ptr(SV), _ptr(coef), C.c_double(rho), C.c_double(gamma),
                                 int(labels[0]), int(labels[1]), _ptr(X), N, _ptr(dec), _ptr(lab))
    return dec, lab


def ref_svm_predict(SV, coef, rho, gamma, labels, X, nsv0=None):
    L = ref("svm")
    if L is None:
        return None
    L.ref_svm_build.restype = C.c_void_p
    SV = np.ascontiguousarray(SV, np.float64); coef = np.ascontiguousarray(coef, np.float64)
    X = np.asfortranarray(X, np.float64); N, F = X.shape
    m = C.c_void_p(L.ref_svm_build(SV.shape[0], SV.shape[1], _ptr(SV), _ptr(coef), C.c_double(rho), C.c_double(gamma),
                                   int(labels[0]), int(labels[1]), int(nsv0 if nsv0 is not None else SV.shape[0] // 2)))
    dec = np.zeros(N); lab = np.zeros(N)
    L.ref_svm_predict(m, _ptr(X), N, F, _ptr(dec), _ptr(lab))
    return dec, lab


def synth_model(nsv, nfeat=9, seed=7):
    """SURVEY.md 8(d) stand-in for the missing svmRBFmodel_all.mat: gamma 1/9, SV ~ N(0,1), coef ~ +-U(0,1), rho 0.1."""
    r = np.random.default_rng(seed)
    SV = r.standard_normal((nsv, nfeat))
    coef = r.uniform(0, 1, nsv) * np.where(np.arange(nsv) < nsv // 2, 1.0, -1.0)
    return SV, coef, 0.1, 1.0 / nfeat, (1, -1)
