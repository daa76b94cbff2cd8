"""The reference's wire formats at either side of the hot path, so that a MATLAB user
can swap one stage at a time.  Host-side I/O only.

  * feature text files   mv_x.txt / mv_y.txt / depth.txt / bitmap.txt written by code_mat_h.exe
                         (code_mat_h.cpp:28-57 saveres) and read by main.m:38-51: per picture a line with the picture
                         index, then one line per row of space-terminated integers;
  * feature .mat files   mv_x.mat / mv_y.mat / depth.mat / bitmap.mat with 1 x N cell arrays video_mv_x, ... (main.m:52-55);
  * Saliencymap_<name>.mat  1 x N cell array Sal_map of H x W doubles (main.m:309);
  * <POC>.png            8-bit maps, one file per picture (TDecGop.cpp:327-332);
  * svmRBFmodel_all.mat  model (libsvm's MATLAB struct, svm_model_matlab.c:17-29) + meanVec + stdVec (main.m:113).
"""
import os

import numpy as np

FEATURE_FILES = (("mv_x", "video_mv_x"), ("mv_y", "video_mv_y"), ("depth", "video_depth"), ("bitmap", "video_bitmap"))


def write_feature_txt(directory, mvx, mvy, depth, bitmap, first_index=0):
    """mvx / mvy / depth: [N][h/4][w/4] integers, bitmap: [N][ceil(h/64)][ceil(w/64)] (bytes per CTU)."""
    os.makedirs(directory, exist_ok=True)
    for (stem, _), arr in zip(FEATURE_FILES, (mvx, mvy, depth, bitmap)):
        arr = np.asarray(arr)
        with open(os.path.join(directory, stem + ".txt"), "w") as f:
            for k in range(arr.shape[0]):
                f.write(f"{first_index + k}\n")
                for row in arr[k]:
                    f.write(" ".join(str(int(v)) for v in row) + " \n")


def read_feature_txt(directory, w, h, nframes):
    """Inverse of write_feature_txt with main.m:38-51's shapes.  Returns (mvx, mvy, depth, bitmap) as int32 arrays."""
    shapes = {"mv_x": (h // 4, w // 4), "mv_y": (h // 4, w // 4), "depth": (h // 4, w // 4), "bitmap": (-(-h // 64), -(-w // 64))}
    out = []
    for stem, _ in FEATURE_FILES:
        r, c = shapes[stem]
        vals = np.array(open(os.path.join(directory, stem + ".txt")).read().split(), np.int64)
        if vals.size != nframes * (1 + r * c):
            raise ValueError(f"{stem}.txt holds {vals.size} numbers, expected {nframes} pictures of 1 + {r}x{c}")
        out.append(vals.reshape(nframes, 1 + r * c)[:, 1:].reshape(nframes, r, c).astype(np.int32))
    return tuple(out)


def _cell_row(arrays):
    c = np.empty((1, len(arrays)), dtype=object)
    for i, a in enumerate(arrays):
        c[0, i] = a
    return c


def save_feature_mats(directory, mvx, mvy, depth, bitmap):
    """mv_x.mat ... bitmap.mat as main.m:52-55 saves them (MAT v5, 1 x N cell arrays of doubles)."""
    import scipy.io as sio
    os.makedirs(directory, exist_ok=True)
    for (stem, var), arr in zip(FEATURE_FILES, (mvx, mvy, depth, bitmap)):
        sio.savemat(os.path.join(directory, stem + ".mat"), {var: _cell_row([np.asarray(a, np.float64) for a in arr])})


def load_feature_mats(directory):
    """Reads the four .mat files (e.g. the authors' HEVCfiles/<video>/*.mat).  Returns dict name -> list of 2-D arrays."""
    import scipy.io as sio
    out = {}
    for stem, var in FEATURE_FILES:
        cell = sio.loadmat(os.path.join(directory, stem + ".mat"))[var]
        out[stem] = [np.asarray(cell[0, i]) for i in range(cell.shape[1])]
    return out


def save_saliency_mat(path, maps):
    """Saliencymap_<name>.mat: Sal_map{k} = H x W double (main.m:309)."""
    import scipy.io as sio
    sio.savemat(path, {"Sal_map": _cell_row([np.asarray(m, np.float64) for m in maps])}, do_compression=True)


def load_saliency_mat(path):
    import scipy.io as sio
    cell = sio.loadmat(path)["Sal_map"]
    return [np.asarray(cell[0, i]) for i in range(cell.shape[1])]


def write_png_per_poc(directory, maps, first_poc=0):
    """<POC>.png like TDecGop.cpp:327-332; float maps in [0, 1] are scaled by 255 and rounded, uint8 maps written as is."""
    import cv2
    os.makedirs(directory, exist_ok=True)
    for k, m in enumerate(maps):
        m = np.asarray(m)
        if m.dtype != np.uint8:
            m = np.rint(np.clip(m, 0.0, 1.0) * 255.0).astype(np.uint8)
        if not cv2.imwrite(os.path.join(directory, f"{first_poc + k}.png"), m):
            raise IOError(f"cannot write {first_poc + k}.png")


# ---- svmRBFmodel_all.mat (main.m:113: load('svmRBFmodel_all','model','meanVec','stdVec')) -------------------------------------
# `model` is the struct libsvm's MATLAB interface builds (libsvm-3.17/matlab/svm_model_matlab.c:17-29, field order fixed):
# Parameters [svm_type kernel_type degree gamma coef0]', nr_class, totalSV, rho, Label, sv_indices, ProbA, ProbB, nSV,
# sv_coef (totalSV x (nr_class-1)), SVs (sparse totalSV x nfeat).  The file is absent from the reference checkout; these two
# functions read / write that layout (MAT v5) so that a trained model drops in as the file the reference loads.
_MODEL_FIELDS = ("Parameters", "nr_class", "totalSV", "rho", "Label", "sv_indices", "ProbA", "ProbB", "nSV", "sv_coef", "SVs")


def save_svm_model_mat(path, SVs, sv_coef, rho, gamma, Label=(1, -1), meanVec=None, stdVec=None):
    """Write model / meanVec / stdVec as svmRBFmodel_all.mat holds them (2-class C-SVC, RBF)."""
    import scipy.io as sio
    import scipy.sparse as sp
    SV = np.asarray(SVs, np.float64); coef = np.asarray(sv_coef, np.float64).reshape(-1, 1)
    if SV.ndim != 2 or coef.shape[0] != SV.shape[0]:
        raise ValueError("SVs must be totalSV x nfeat and sv_coef totalSV")
    npos = int(np.sum(coef > 0))
    model = {
        "Parameters": np.array([[0.0], [2.0], [3.0], [float(gamma)], [0.0]]),      # C-SVC, RBF, degree 3 (unused), gamma, coef0
        "nr_class": np.array([[2.0]]), "totalSV": np.array([[float(SV.shape[0])]]), "rho": np.array([[float(rho)]]),
        "Label": np.array([[float(Label[0])], [float(Label[1])]]), "sv_indices": np.arange(1, SV.shape[0] + 1, dtype=np.float64).reshape(-1, 1),
        "ProbA": np.zeros((0, 0)), "ProbB": np.zeros((0, 0)), "nSV": np.array([[float(npos)], [float(SV.shape[0] - npos)]]),
        "sv_coef": coef, "SVs": sp.csc_matrix(SV),
    }
    out = {"model": model}
    if meanVec is not None:
        out["meanVec"] = np.asarray(meanVec, np.float64).reshape(1, -1)
    if stdVec is not None:
        out["stdVec"] = np.asarray(stdVec, np.float64).reshape(1, -1)
    sio.savemat(path, out, format="5", do_compression=True)


def load_svm_model_mat(path):
    """-> dict(SVs [totalSV][nfeat], sv_coef [totalSV], rho, gamma, Label (2,), meanVec (nfeat,), stdVec (nfeat,)) from svmRBFmodel_all.mat.
    Only what the reference uses is accepted: svm_type 0 (C-SVC), kernel_type 2 (RBF), two classes (main.m:297, svm.cpp:2501)."""
    import scipy.io as sio
    m = sio.loadmat(path, squeeze_me=False, struct_as_record=True)
    if "model" not in m:
        raise ValueError(f"{path}: no variable 'model'")
    s = m["model"][0, 0]
    missing = [f for f in ("Parameters", "nr_class", "rho", "Label", "sv_coef", "SVs") if f not in s.dtype.names]
    if missing:
        raise ValueError(f"{path}: model struct lacks {missing}")
    par = np.asarray(s["Parameters"], np.float64).reshape(-1)
    if int(par[0]) != 0 or int(par[1]) != 2 or int(np.asarray(s["nr_class"]).reshape(-1)[0]) != 2:
        raise ValueError(f"{path}: only 2-class C-SVC with an RBF kernel is supported")
    SV = s["SVs"]
    SV = np.asarray(SV.todense() if hasattr(SV, "todense") else SV, np.float64)
    out = dict(SVs=np.ascontiguousarray(SV), sv_coef=np.asarray(s["sv_coef"], np.float64).reshape(-1),
               rho=float(np.asarray(s["rho"]).reshape(-1)[0]), gamma=float(par[3]),
               Label=tuple(int(v) for v in np.asarray(s["Label"]).reshape(-1)[:2]))
    for k in ("meanVec", "stdVec"):
        out[k] = np.asarray(m[k], np.float64).reshape(-1) if k in m else None
    return out
