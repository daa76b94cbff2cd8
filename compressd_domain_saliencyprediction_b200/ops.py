"""Host-side mirrors of the reference's operator surfaces, bound to the C ABI (include/cds.h).

Names and argument meaning follow the reference so that parity tests read like its own calls:
  computecontrast5        <- computecontrast5.cpp:52 (mex), called by Sudoku_contrastcpp5.m:37
  Sudoku_contrastcpp5     <- Sudoku_contrastcpp5.m:1
  cu_contrast5            <- cu_contrast5.m:1
  getFrame / getframe     <- HM_16.0_features/.../code_mat_h.h:575
  SvmModel / libsvmpredict<- libsvm-3.17/matlab/svmpredict.c:272, svm.cpp:2501
  camera_motion, mv_vote, magnitude_vote <- main.m:77-105, mv_vote.m, magnitude_vote.m
Everything computes on the GPU; nothing here falls back to numpy.
"""
import ctypes as C

import numpy as np

from ._lib import CdsError, check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------ (b)
def computecontrast5(subsaliencymap3, W, len_, WINDOW_SIZE, m, n):
    """contrastmap = computecontrast5(pad, W, len, WINDOW_SIZE, m, n)   (Sudoku_contrastcpp5.m:37)"""
    pad = np.asfortranarray(subsaliencymap3, dtype=np.float64)
    Wf = np.asfortranarray(W, dtype=np.float64)
    m, n, len_, win = int(m), int(n), int(len_), int(WINDOW_SIZE)
    if Wf.shape != (win, win):
        raise CdsError("W must be WINDOW_SIZE x WINDOW_SIZE")
    out = np.zeros((m, n), dtype=np.float64, order="F")
    check(lib().cds_contrast5(_ptr(pad), pad.shape[0], pad.shape[1], _ptr(Wf), len_, win, m, n, _ptr(out)),
          "cds_contrast5")
    return out


def Sudoku_contrastcpp5(subsaliencymap, delta, cusize, patchsize):
    a = np.ascontiguousarray(subsaliencymap, dtype=np.float64)
    out = np.zeros_like(a)
    check(lib().cds_sudoku_contrast(_ptr(a), a.shape[0], a.shape[1], float(delta), float(cusize),
                                    float(patchsize), _ptr(out)), "cds_sudoku_contrast")
    return out


def cu_contrast5(salmap, cusize, delta, patchsize):
    a = np.ascontiguousarray(salmap, dtype=np.float64)
    out = np.zeros_like(a)
    check(lib().cds_cu_contrast5(_ptr(a), a.shape[0], a.shape[1], int(cusize), float(delta),
                                 float(patchsize), _ptr(out)), "cds_cu_contrast5")
    return out


# ------------------------------------------------------------------------------------------ (a)
def nctu(w, h):
    return ((h + 63) // 64) * ((w + 63) // 64)


def pack_ref_arrays(MV, Pred, CUPU):
    """TDecCu::MV/Pred/CUPU (float[nctu][1000], 999-terminated) -> packed (hdr, tok)."""
    MV = np.ascontiguousarray(MV, np.float32)
    Pred = np.ascontiguousarray(Pred, np.float32)
    CUPU = np.ascontiguousarray(CUPU, np.float32)
    n = MV.shape[0]
    hdr = np.zeros((n, 4), np.int32)
    tok = np.zeros(3000 * n, np.int16)
    nt = lib().cds_syntax_pack_ref(n, _ptr(MV), _ptr(Pred), _ptr(CUPU), _ptr(hdr), _ptr(tok))
    check(nt, "cds_syntax_pack_ref")
    return hdr, tok[:nt].copy()


def getframe(w, h, hdr, tok):
    """Packed syntax of one picture -> (mv_x, mv_y, depth) 4x4-block maps (int16, int16, uint8)."""
    hdr = np.ascontiguousarray(hdr, np.int32)
    tok = np.ascontiguousarray(tok, np.int16)
    if hdr.shape != (nctu(w, h), 4):
        raise CdsError(f"hdr must be ({nctu(w, h)}, 4)")
    mvx = np.zeros((h // 4, w // 4), np.int16)
    mvy = np.zeros_like(mvx)
    depth = np.zeros((h // 4, w // 4), np.uint8)
    check(lib().cds_getframe(w, h, _ptr(hdr), _ptr(tok), tok.size, _ptr(mvx), _ptr(mvy), _ptr(depth)),
          "cds_getframe")
    return mvx, mvy, depth


def getframe_semantic(w, h, depth256, pred256, mv256):
    """z-scan partition arrays of TComDataCU ([F][nctu][256] depth / prediction mode, [F][nctu][256][2] mv) ->
    (mvx, mvy, depth) maps [F][h/4][w/4]."""
    d = np.ascontiguousarray(depth256, np.uint8); p = np.ascontiguousarray(pred256, np.uint8); m = np.ascontiguousarray(mv256, np.int16)
    F = d.shape[0]
    if d.shape != (F, nctu(w, h), 256) or p.shape != d.shape or m.shape != (F, nctu(w, h), 256, 2):
        raise CdsError("getframe_semantic: partition arrays do not match the picture size")
    mvx = np.zeros((F, h // 4, w // 4), np.int16); mvy = np.zeros_like(mvx); depth = np.zeros((F, h // 4, w // 4), np.uint8)
    check(lib().cds_getframe_semantic(w, h, F, _ptr(d), _ptr(p), _ptr(m), _ptr(mvx), _ptr(mvy), _ptr(depth)), "cds_getframe_semantic")
    return mvx, mvy, depth


def getFrame(mv_arr, pred_arr, cupu_arr, w, h):
    """getFrame(TDecCu::MV, TDecCu::Pred, TDecCu::CUPU, mv_x, mv_y, depth)   (TDecGop.cpp:270)"""
    hdr, tok = pack_ref_arrays(mv_arr, pred_arr, cupu_arr)
    mvx, mvy, depth = getframe(w, h, hdr, tok)
    return mvx.astype(np.float32), mvy.astype(np.float32), depth.astype(np.float32)


from .synth import synth_frame, synth_clip  # noqa: E402,F401  (host-only libcds_synth.so; kept here for the tests' ops.synth_* spelling)


# ------------------------------------------------------------------------------------------ (d)
class SvmModel:
    """2-class C-SVC / RBF libsvm model (the 'model' of svmRBFmodel_all.mat, main.m:113)."""

    def __init__(self, SVs=None, sv_coef=None, rho=0.0, gamma=None, Label=(1, -1), path=None):
        self._h = C.c_void_p()
        if path is not None:
            check(lib().cds_svm_model_load(str(path).encode(), C.byref(self._h)), "cds_svm_model_load")
        else:
            SV = np.ascontiguousarray(SVs, np.float64)
            coef = np.ascontiguousarray(np.asarray(sv_coef).reshape(-1), np.float64)
            if SV.ndim != 2 or coef.size != SV.shape[0]:
                raise CdsError("SVs must be nsv x nfeat and sv_coef nsv")
            check(lib().cds_svm_model_create(SV.shape[0], SV.shape[1], _ptr(SV), _ptr(coef), float(rho),
                                             float(gamma), int(Label[0]), int(Label[1]), C.byref(self._h)),
                  "cds_svm_model_create")
        self.Label = tuple(Label)

    @classmethod
    def from_mat(cls, path):
        """(model, meanVec, stdVec) from a svmRBFmodel_all.mat (main.m:113; layout of svm_model_matlab.c:17-29)."""
        from . import formats
        m = formats.load_svm_model_mat(path)
        return cls(SVs=m["SVs"], sv_coef=m["sv_coef"], rho=m["rho"], gamma=m["gamma"], Label=m["Label"]), m["meanVec"], m["stdVec"]

    @property
    def handle(self):
        return self._h

    @property
    def totalSV(self):
        return lib().cds_svm_model_nsv(self._h)

    @property
    def nfeat(self):
        return lib().cds_svm_model_nfeat(self._h)

    def decision_values(self, X):
        X = np.asfortranarray(X, dtype=np.float64)
        N, F = X.shape
        dec = np.zeros(N, np.float64)
        lab = np.zeros(N, np.float64)
        check(lib().cds_svm_rbf_predict(self._h, _ptr(X), N, F, _ptr(dec), _ptr(lab)), "cds_svm_rbf_predict")
        return dec, lab

    def __del__(self):
        try:
            if self._h:
                lib().cds_svm_model_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def libsvmpredict(testing_label_vector, testing_instance_matrix, model):
    """[predicted_label, accuracy, decision_values] = libsvmpredict(y, X, model)   (svmpredict.c:258)"""
    y = np.asarray(testing_label_vector, np.float64).reshape(-1)
    dec, lab = model.decision_values(testing_instance_matrix)
    if y.size != lab.size:
        raise CdsError("label vector and instance matrix disagree")
    n = y.size
    correct = float(np.sum(lab == y))
    err = lab - y
    sump, sumt = lab.sum(), y.sum()
    sumpp, sumtt, sumpt = (lab * lab).sum(), (y * y).sum(), (lab * y).sum()
    den = (n * sumpp - sump * sump) * (n * sumtt - sumt * sumt)
    scc = ((n * sumpt - sump * sumt) ** 2) / den if den != 0 else float("nan")
    accuracy = np.array([correct / n * 100.0, float((err * err).sum()) / n, scc])   # svmpredict.c:233-238
    return lab.reshape(-1, 1), accuracy.reshape(3, 1), dec.reshape(-1, 1)


# ------------------------------------------------------------------------------------------ (c)
def camera_motion(w, h, mvx, mvy, ctu_bytes, depth):
    """main.m:77-105 for one picture.  Returns dict(Vx, Vy, mv_norm, bit_norm, depth_norm, cam, ave, flag)."""
    h4, w4 = h // 4, w // 4
    mvx = np.ascontiguousarray(mvx, np.int16); mvy = np.ascontiguousarray(mvy, np.int16)
    byts = np.ascontiguousarray(ctu_bytes, np.uint16).reshape(-1); depth = np.ascontiguousarray(depth, np.uint8)
    if mvx.shape != (h4, w4) or byts.size != nctu(w, h):
        raise CdsError("camera_motion: map shapes do not match the picture size")
    outs = [np.zeros((h4, w4), np.float32) for _ in range(5)]
    cam = np.zeros(2, np.int32); ave = np.zeros(2, np.float64); flag = C.c_int()
    check(lib().cds_camera_motion(w, h, _ptr(mvx), _ptr(mvy), _ptr(byts), _ptr(depth), *[_ptr(a) for a in outs],
                                  _ptr(cam), _ptr(ave), C.byref(flag)), "cds_camera_motion")
    return dict(Vx=outs[0], Vy=outs[1], mv_norm=outs[2], bit_norm=outs[3], depth_norm=outs[4], cam=cam, ave=ave,
                flag=flag.value)


def mv_vote(eVx, eVy):
    """[x_ave, y_ave] = mv_vote(eVx, eVy)   (mv_vote.m; values are multiples of 1/8 pixel)."""
    vx = np.asarray(eVx, np.float64).reshape(-1) * 8; vy = np.asarray(eVy, np.float64).reshape(-1) * 8
    if np.any(vx != np.round(vx)) or np.any(vy != np.round(vy)):
        raise CdsError("mv_vote: inputs must be multiples of 1/8 pixel (medfilt2 of quarter-pel MVs / 4)")
    vx8 = np.ascontiguousarray(vx, np.int32); vy8 = np.ascontiguousarray(vy, np.int32)
    xa = C.c_double(); ya = C.c_double()
    check(lib().cds_mv_vote(_ptr(vx8), _ptr(vy8), vx8.size, C.byref(xa), C.byref(ya)), "cds_mv_vote")
    return xa.value, ya.value


def magnitude_vote(value):
    v = np.ascontiguousarray(value, np.float64).reshape(-1)
    out = C.c_double()
    check(lib().cds_magnitude_vote(_ptr(v), v.size, C.byref(out)), "cds_magnitude_vote")
    return out.value
