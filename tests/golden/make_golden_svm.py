#!/usr/bin/env python3
"""Generate tests/golden/svm_heart_scale.npz with the reference's own libsvm-3.17 (oracle/_ref/libref_svm.so):
svm-train -s 0 -t 2 -c 1 -g 0.07 on libsvm-3.17/heart_scale, then svm_predict_values on the same data
(the known-answer recipe of libsvm-3.17/matlab/README:177-188: Accuracy = 86.6667% (234/270))."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_svm.so"))
    L.ref_svm_train.restype = C.c_void_p
    rows, y = [], []
    for line in open("/root/reference/libsvm-3.17/heart_scale"):
        parts = line.split()
        y.append(float(parts[0])); r = np.zeros(13)
        for kv in parts[1:]:
            k, v = kv.split(":"); r[int(k) - 1] = float(v)
        rows.append(r)
    X = np.ascontiguousarray(np.array(rows)); y = np.array(y)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    m = C.c_void_p(L.ref_svm_train(p(X), p(y), X.shape[0], 13, C.c_double(1.0), C.c_double(0.07)))
    nsv = L.ref_svm_nsv(m)
    SV = np.zeros((nsv, 13)); coef = np.zeros(nsv); rho = C.c_double(); gamma = C.c_double(); labels = np.zeros(2, np.int32)
    L.ref_svm_export(m, 13, p(SV), p(coef), C.byref(rho), C.byref(gamma), p(labels))
    Xf = np.asfortranarray(X); dec = np.zeros(X.shape[0]); lab = np.zeros(X.shape[0])
    L.ref_svm_predict(m, p(Xf), X.shape[0], 13, p(dec), p(lab))
    acc = int((lab == y).sum())
    print("nSV", nsv, "accuracy", acc, "/", len(y))
    assert nsv == 130 and acc == 234
    L.ref_svm_save(m, os.path.join(HERE, "svm_heart_scale.model").encode())
    np.savez_compressed(os.path.join(HERE, "svm_heart_scale.npz"), X=X, y=y, SV=SV, coef=coef, rho=rho.value,
                        gamma=gamma.value, labels=labels, dec_ref=dec, pred_ref=lab)


if __name__ == "__main__":
    sys.exit(main())
