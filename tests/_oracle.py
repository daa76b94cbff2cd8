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
SAN = os.environ.get("CDS_ORACLE_SAN") == "1"      # load the ASan + UBSan builds (oracle/Makefile `sanitize`; needs LD_PRELOAD=libasan)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle():
    global _ORC
    if _ORC is None:
        so = os.path.join(ORACLE_DIR, "_build", "san" if SAN else "", "libcds_oracle.so")
        src = os.path.join(ORACLE_DIR, "cds_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "sanitize" if SAN else "oracle"])
        L = C.CDLL(so)
        if hasattr(L, "orc_magnitude_vote"):
            L.orc_magnitude_vote.restype = C.c_double
        _ORC = L
    return _ORC


def ref(name):
    """oracle/_ref/libref_<name>.so or None (built only where /root/reference exists)."""
    if name not in _REF:
        so = os.path.join(ORACLE_DIR, "_ref", "san" if SAN else "", f"libref_{name}.so")
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


def load_libsvm_model(path):
    """Parse the text written by svm_save_model (svm.cpp:2756) for a 2-class C-SVC / RBF model -> (SV, coef, rho, gamma, labels)."""
    hdr = {}
    with open(path) as f:
        for line in f:
            k = line.split()
            if k[0] == "SV":
                break
            hdr[k[0]] = k[1:]
        rows = [l.split() for l in f if l.strip()]
    assert hdr["svm_type"] == ["c_svc"] and hdr["kernel_type"] == ["rbf"] and hdr["nr_class"] == ["2"]
    nfeat = max(int(t.split(":")[0]) for r in rows for t in r[1:])
    SV = np.zeros((len(rows), nfeat)); coef = np.zeros(len(rows))
    for i, r in enumerate(rows):
        coef[i] = float(r[0])
        for t in r[1:]:
            j, v = t.split(":"); SV[i, int(j) - 1] = float(v)
    assert len(rows) == int(hdr["total_sv"][0])
    return SV, coef, float(hdr["rho"][0]), float(hdr["gamma"][0]), (int(hdr["label"][0]), int(hdr["label"][1]))


# ---------------------------------------------------------------------------- stage (c) / clip
def orc_camera_motion(w, h, mvx, mvy, bitmap, depth):
    h4, w4 = h // 4, w // 4
    mvx = np.ascontiguousarray(mvx, np.int16); mvy = np.ascontiguousarray(mvy, np.int16)
    bitmap = np.ascontiguousarray(bitmap, np.uint16); depth = np.ascontiguousarray(depth, np.uint8)
    outs = [np.zeros((h4, w4)) for _ in range(5)]
    cam = np.zeros(2, np.int32); ave = np.zeros(2); flag = C.c_int()
    oracle().orc_camera_motion(w, h, _ptr(mvx), _ptr(mvy), _ptr(bitmap), _ptr(depth), *[_ptr(a) for a in outs],
                               _ptr(cam), _ptr(ave), C.byref(flag))
    return outs, cam, ave, flag.value


def orc_direction_cluster(vx, vy):
    vx = np.ascontiguousarray(vx, np.float64); vy = np.ascontiguousarray(vy, np.float64)
    xa = C.c_double(); ya = C.c_double()
    oracle().orc_direction_cluster(_ptr(vx), _ptr(vy), vx.size, C.byref(xa), C.byref(ya))
    return xa.value, ya.value


def orc_magnitude_vote(v):
    v = np.ascontiguousarray(v, np.float64)
    return float(oracle().orc_magnitude_vote(_ptr(v), v.size))


class ClipParams(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("nframes", C.c_int), ("fps", C.c_double),
                ("nsv", C.c_int), ("nfeat", C.c_int), ("SV", C.c_void_p), ("coef", C.c_void_p),
                ("rho", C.c_double), ("gamma", C.c_double), ("label0", C.c_int), ("label1", C.c_int),
                ("meanVec", C.c_void_p), ("stdVec", C.c_void_p), ("use_rbf", C.c_int)]


def orc_process_clip(w, h, fps, mvx, mvy, depth, bitmap, model, mean_vec, std_vec, first_out=0, use_rbf=1, want_feat=False):
    """mvx/mvy/depth: [F][h4][w4]; bitmap [F][nctu].  Returns (maps [F-first_out][h][w], cam [F][2], feat200 or None)."""
    F = mvx.shape[0]
    SV, coef, rho, gamma, labels = model
    SV = np.ascontiguousarray(SV, np.float64); coef = np.ascontiguousarray(coef, np.float64)
    mv = np.ascontiguousarray(mean_vec, np.float64); sv = np.ascontiguousarray(std_vec, np.float64)
    p = ClipParams(w, h, F, float(fps), SV.shape[0], SV.shape[1], SV.ctypes.data, coef.ctypes.data, float(rho), float(gamma),
                   int(labels[0]), int(labels[1]), mv.ctypes.data, sv.ctypes.data, int(use_rbf))
    mvx = np.ascontiguousarray(mvx, np.int16); mvy = np.ascontiguousarray(mvy, np.int16)
    depth = np.ascontiguousarray(depth, np.uint8); bitmap = np.ascontiguousarray(bitmap, np.uint16)
    out = np.zeros((F - first_out, h, w)); cam = np.zeros((F, 2), np.int32)
    feat = np.zeros((F - first_out, 9, 200, 200)) if (want_feat and use_rbf) else None
    rc = oracle().orc_process_clip(C.byref(p), _ptr(mvx), _ptr(mvy), _ptr(depth), _ptr(bitmap), first_out, _ptr(out), _ptr(cam),
                                   _ptr(feat) if feat is not None else None)
    assert rc == 0, rc
    return out, cam, feat


# ---------------------------------------------------------------------------- CPU baseline with the reference's own kernels
_REF_KEEP = []


def use_reference_kernels(model=None):
    """Route the oracle's contrast and SVM calls through oracle/_ref (the reference's verbatim computecontrast5 mexFunction
    and libsvm-3.17).  model = (SV, coef, rho, gamma, labels) builds the libsvm model; None restores the restatements.
    Returns True when the reference kernels are in use."""
    L = oracle()
    L.orc_use_reference_kernels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    if model is None or ref("contrast") is None or ref("svm") is None:
        L.orc_use_reference_kernels(None, None, None)
        return False
    SV, coef, rho, gamma, labels = model
    Ls = ref("svm")
    Ls.ref_svm_build.restype = C.c_void_p
    SV = np.ascontiguousarray(SV, np.float64); coef = np.ascontiguousarray(coef, np.float64)
    m = Ls.ref_svm_build(SV.shape[0], SV.shape[1], _ptr(SV), _ptr(coef), C.c_double(rho), C.c_double(gamma),
                         int(labels[0]), int(labels[1]), int(np.sum(coef > 0)))
    _REF_KEEP.append((SV, coef))
    L.orc_use_reference_kernels(C.cast(ref("contrast").ref_computecontrast5, C.c_void_p), C.cast(Ls.ref_svm_predict, C.c_void_p), C.c_void_p(m))
    return True


# ---------------------------------------------------------------------------- fast version (cds_oracle_fast.c)
class FastOracle:
    """The reference's fast version (TDecGop.cpp:225-333 + salmat.cpp), one picture at a time."""

    def __init__(self, w, h, fps):
        L = oracle()
        L.orc_fast_create.restype = C.c_void_p
        L.orc_fast_create.argtypes = [C.c_int, C.c_int, C.c_double]
        L.orc_fast_destroy.argtypes = [C.c_void_p]
        L.orc_fast_frame.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.orc_fast_PS.argtypes = [C.c_void_p]; L.orc_fast_PT.argtypes = [C.c_void_p]
        self.L, self.w, self.h = L, w, h
        self.o = L.orc_fast_create(w, h, float(fps))
        assert self.o, "orc_fast_create failed"
        self.PS, self.PT = L.orc_fast_PS(self.o), L.orc_fast_PT(self.o)

    def frame(self, mvx, mvy, depth, ctu_bytes, want_feat=False):
        """-> (u8 map [h][w], f32 map [h][w], cam (2,), feat9 [9][h][w] or None)"""
        mvx = np.ascontiguousarray(mvx, np.int16); mvy = np.ascontiguousarray(mvy, np.int16)
        depth = np.ascontiguousarray(depth, np.uint8); byts = np.ascontiguousarray(ctu_bytes, np.uint16)
        u8 = np.zeros((self.h, self.w), np.uint8); f32 = np.zeros((self.h, self.w), np.float32); cam = np.zeros(2, np.int32)
        feat = np.zeros((9, self.h, self.w), np.float32) if want_feat else None
        rc = self.L.orc_fast_frame(self.o, _ptr(mvx), _ptr(mvy), _ptr(depth), _ptr(byts), _ptr(u8), _ptr(f32), _ptr(cam),
                                   _ptr(feat) if want_feat else None)
        assert rc == 0
        return u8, f32, cam, feat

    def close(self):
        if self.o:
            self.L.orc_fast_destroy(self.o); self.o = None

    def __del__(self):
        self.close()
