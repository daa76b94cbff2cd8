"""Wire formats at either side of the hot path (main.m:25-55, :309; code_mat_h.cpp:28-57; TDecGop.cpp:327-332)."""
import os

import numpy as np

import _oracle as o
from compressd_domain_saliencyprediction_b200 import formats

GOLD = os.path.join(o.ROOT, "tests", "golden")


def _golden():
    d = np.load(os.path.join(GOLD, "real_BasketballDrill.npz"))
    return d["mvx"][:4], d["mvy"][:4], d["depth"][:4], d["bitmap"][:4], int(d["w"]), int(d["h"])


def test_feature_txt_round_trip_and_layout(tmp_path):
    mvx, mvy, dep, bm, w, h = _golden()
    formats.write_feature_txt(tmp_path, mvx, mvy, dep, bm)
    lines = open(tmp_path / "mv_x.txt").read().split("\n")
    assert lines[0] == "0" and lines[1].endswith(" ") and len(lines[1].split()) == w // 4      # saveres(): index line, then h/4 rows
    assert lines[1 + h // 4] == "1"
    rx, ry, rd, rb = formats.read_feature_txt(tmp_path, w, h, 4)
    assert np.array_equal(rx, mvx) and np.array_equal(ry, mvy) and np.array_equal(rd, dep) and np.array_equal(rb, bm)


def test_feature_mats_round_trip_like_the_authors_files(tmp_path):
    mvx, mvy, dep, bm, w, h = _golden()
    formats.save_feature_mats(tmp_path, mvx, mvy, dep, bm)
    back = formats.load_feature_mats(tmp_path)
    assert len(back["mv_x"]) == 4 and back["mv_x"][0].shape == (h // 4, w // 4) and back["bitmap"][0].shape == bm[0].shape
    assert all(np.array_equal(back["mv_y"][k], mvy[k]) for k in range(4))
    assert all(np.array_equal(back["depth"][k], dep[k]) for k in range(4))


def test_saliency_mat_and_png(tmp_path):
    import cv2
    maps = np.random.default_rng(1).uniform(0, 1, (3, 48, 64)).astype(np.float32)
    formats.save_saliency_mat(tmp_path / "Saliencymap_x.mat", maps)
    back = formats.load_saliency_mat(tmp_path / "Saliencymap_x.mat")
    assert len(back) == 3 and np.array_equal(back[1], maps[1].astype(np.float64))
    formats.write_png_per_poc(tmp_path / "HMdata", maps, first_poc=5)
    img = cv2.imread(str(tmp_path / "HMdata" / "6.png"), cv2.IMREAD_GRAYSCALE)
    assert img.shape == (48, 64) and np.array_equal(img, np.rint(maps[1] * 255).astype(np.uint8))


def test_svmrbfmodel_all_mat_round_trip(tmp_path):
    """svmRBFmodel_all.mat (main.m:113) as libsvm's MATLAB interface lays the model struct out (svm_model_matlab.c:17-29): the 11 fields
    in their fixed order, sparse SVs, plus meanVec / stdVec; and the committed stand-in file equals the libsvm text model it came from."""
    import os
    import scipy.io as sio
    import scipy.sparse as sp
    import _oracle as o
    from compressd_domain_saliencyprediction_b200 import formats
    r = np.random.default_rng(4)
    SV = r.standard_normal((17, 9)); SV[::3, 4] = 0.0
    coef = np.concatenate([r.random(9), -r.random(8)])
    p = str(tmp_path / "svmRBFmodel_all.mat")
    formats.save_svm_model_mat(p, SV, coef, 0.25, 1 / 9, (1, -1), np.arange(9) * 0.1, np.arange(1, 10) * 0.5)
    raw = sio.loadmat(p)
    assert raw["model"].dtype.names == ("Parameters", "nr_class", "totalSV", "rho", "Label", "sv_indices", "ProbA", "ProbB", "nSV", "sv_coef", "SVs")
    s = raw["model"][0, 0]
    assert sp.issparse(s["SVs"]) and s["SVs"].shape == (17, 9) and s["sv_coef"].shape == (17, 1)
    assert s["Parameters"].reshape(-1).tolist()[:2] == [0.0, 2.0] and s["nSV"].reshape(-1).tolist() == [9.0, 8.0]
    m = formats.load_svm_model_mat(p)
    assert np.array_equal(m["SVs"], SV) and np.array_equal(m["sv_coef"], coef) and m["rho"] == 0.25 and m["Label"] == (1, -1)
    assert np.array_equal(m["meanVec"], np.arange(9) * 0.1) and np.array_equal(m["stdVec"], np.arange(1, 10) * 0.5)
    g = os.path.join(o.ROOT, "tests", "golden")
    SVt, coeft, rhot, gammat, labelst = o.load_libsvm_model(os.path.join(g, "standin_svmRBFmodel.model"))
    mm = formats.load_svm_model_mat(os.path.join(g, "standin_svmRBFmodel_all.mat"))
    assert np.array_equal(mm["SVs"], SVt) and np.array_equal(mm["sv_coef"], coeft) and mm["rho"] == rhot and mm["gamma"] == gammat
