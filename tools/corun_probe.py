#!/usr/bin/env python3
"""Co-residency probe: how fast does the 48x48 contrast kernel run while another kernel shares the SMs with it?

The partner runs on a high-priority stream and is launched first (as the deferred SVM is in pipeline.cu); the contrast kernel
(one tile per CTA) is timed with CUDA events on its own stream.  Partners: the persistent tcgen05 SVM kernel itself, and spinners
that issue only MUFU.EX2, MUFU.EX2 + FADD (the SVM epilogue's mix without tensor cores / TMEM / shared memory), FFMA or FFMA2
from 8 warps per SM.  Prints one JSON line per pairing:
    python tools/corun_probe.py [--frames 64] [--nsv 4096]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import compressd_domain_saliencyprediction_b200 as pkg  # noqa: E402
from compressd_domain_saliencyprediction_b200._lib import check  # noqa: E402


def vp(t):
    return C.c_void_p(t.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--nsv", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    L = pkg.lib()
    assert pkg.device_ok(), L.cds_last_error()
    dev = torch.device("cuda:0")
    F, R, Cc, win, ln = a.frames, 270, 480, 48, 24
    nm = 2 * F
    pad = torch.rand(nm, R + 2 * ln, Cc + 2 * ln, device=dev) * 0.98 + 0.01
    pad[:, :ln, :] = 0; pad[:, -ln:, :] = 0; pad[:, :, :ln] = 0; pad[:, :, -ln:] = 0
    out = torch.empty(nm, R, Cc, device=dev); mx = torch.zeros(nm, dtype=torch.int32, device=dev)
    rng = np.random.default_rng(7)
    SV = rng.standard_normal((a.nsv, 9)); coef = rng.uniform(0, 1, a.nsv) * np.where(np.arange(a.nsv) < a.nsv // 2, 1.0, -1.0)
    m = pkg.ops.SvmModel(SVs=SV, sv_coef=coef, rho=0.1, gamma=1.0 / 9, Label=(1, -1))
    N = 40000 * F
    X = torch.randn(N, 9, device=dev); dec = torch.empty(N, device=dev)
    s_c = torch.cuda.Stream(priority=0)
    s_p = torch.cuda.Stream(priority=-1)
    pc, pp = C.c_void_p(s_c.cuda_stream), C.c_void_p(s_p.cuda_stream)

    def contrast():
        check(L.cds_contrast_sep_dev(vp(pad), vp(out), vp(mx), nm, R, Cc, win, C.c_float(11.0), pc), "contrast")

    partners = {
        "none": None,
        "svm_tcgen05_persistent": lambda: check(L.cds_svm_rbf_predict_dev(m.handle, vp(X), N, vp(dec), pp), "svm"),
        "spin_mufu_8warps": lambda: check(L.cds_probe_spin(0, 1, 256, 6000, pp), "spin"),
        "spin_mufu_fadd_8warps": lambda: check(L.cds_probe_spin(3, 1, 256, 6000, pp), "spin"),
        "spin_mufu_fadd2_8warps": lambda: check(L.cds_probe_spin(4, 1, 256, 6000, pp), "spin"),
        "spin_ffma_8warps": lambda: check(L.cds_probe_spin(1, 1, 256, 12000, pp), "spin"),
        "spin_ffma2_8warps": lambda: check(L.cds_probe_spin(2, 2, 128, 16000, pp), "spin"),
    }
    for name, partner in partners.items():
        tc, tp = [], []
        for r in range(a.reps + 1):
            torch.cuda.synchronize()
            ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if partner:
                ep0.record(s_p); partner(); ep1.record(s_p)
            ec0.record(s_c); contrast(); ec1.record(s_c)
            torch.cuda.synchronize()
            if r:
                tc.append(ec0.elapsed_time(ec1))
                if partner:
                    tp.append(ep0.elapsed_time(ep1))
        rec = {"partner": name, "contrast_ms": float(np.median(tc)), "partner_ms": float(np.median(tp)) if tp else None}
        if partner:          # the partner alone
            ts = []
            for r in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s_p); partner(); e1.record(s_p)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rec["partner_alone_ms"] = float(np.median(ts))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
