#!/usr/bin/env python3
"""Per-kernel timing harness (CUDA events on the launching stream, L2 flushed between iterations).
Used for roofline fractions of the individual kernels and as the ncu target:
    python tools/kernel_bench.py [--kernel contrast|svm|getframe|all] [--w 1920 --h 1080] [--frames 32]
Prints one JSON line per kernel."""
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


def timeit(fn, iters, warmup, flush):
    st = torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()            # > L2 (126 MB): evict
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="all")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--nsv", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--peaks", action="store_true")
    a = ap.parse_args()
    L = pkg.lib()
    assert pkg.device_ok(), L.cds_last_error()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    if a.peaks or a.kernel == "all":
        p = (C.c_double * 3)()
        check(L.cds_measure_peaks(C.byref(p, 0), C.byref(p, 8), C.byref(p, 16)), "peaks")
        peaks = {"fp32_pair_ips": p[0], "ffma_ips": p[1], "mufu_ips": p[2]}
        print(json.dumps({"kernel": "peaks", **peaks}), flush=True)
    h4, w4, F = a.h // 4, a.w // 4, a.frames
    if a.kernel in ("contrast", "all"):
        for win, delta, R, Cc, nm in ((48, 11.0, h4, w4, 2 * F), (24, 13.0, (h4 + 1) // 2, (w4 + 1) // 2, F)):
            ln = win // 2
            pad = torch.rand(nm, R + 2 * ln, Cc + 2 * ln, device=dev) * 0.98 + 0.01
            pad[:, :ln, :] = 0; pad[:, -ln:, :] = 0; pad[:, :, :ln] = 0; pad[:, :, -ln:] = 0
            out = torch.empty(nm, R, Cc, device=dev); mx = torch.zeros(nm, dtype=torch.int32, device=dev)
            fn = lambda: check(L.cds_contrast_sep_dev(vp(pad), vp(out), vp(mx), nm, R, Cc, win, C.c_float(delta), st), "contrast")
            med, mn = timeit(fn, a.iters, a.warmup, flush)
            pairs = float(nm) * R * Cc * win * win
            rec = {"kernel": f"contrast_sep_win{win}", "maps": nm, "rows": R, "cols": Cc, "ms_median": med, "ms_min": mn,
                   "pairs": pairs, "pairs_per_s": pairs / (med * 1e-3), "fp32_instr_per_s": 2 * pairs / (med * 1e-3)}
            if peaks:
                rec["frac_of_fp32_pair_peak"] = rec["fp32_instr_per_s"] / peaks["fp32_pair_ips"]
            print(json.dumps(rec), flush=True)
    if a.kernel in ("svm", "all"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        rng = np.random.default_rng(7)
        nsv = a.nsv
        SV = rng.standard_normal((nsv, 9)); coef = rng.uniform(0, 1, nsv) * np.where(np.arange(nsv) < nsv // 2, 1.0, -1.0)
        m = pkg.ops.SvmModel(SVs=SV, sv_coef=coef, rho=0.1, gamma=1.0 / 9, Label=(1, -1))
        N = 40000 * F
        X = torch.randn(N, 9, device=dev); dec = torch.empty(N, device=dev)
        fn = lambda: check(L.cds_svm_rbf_predict_dev(m.handle, vp(X), N, vp(dec), st), "svm")
        med, mn = timeit(fn, a.iters, a.warmup, flush)
        pairs = float(N) * nsv
        rec = {"kernel": "svm_rbf9", "instances": N, "nsv": nsv, "ms_median": med, "ms_min": mn, "pairs_per_s": pairs / (med * 1e-3),
               "fp32_instr_per_s": 11 * pairs / (med * 1e-3), "mufu_per_s": pairs / (med * 1e-3)}
        if peaks:
            rec["frac_of_ffma_peak"] = rec["fp32_instr_per_s"] / peaks["ffma_ips"]
            rec["frac_of_mufu_peak"] = rec["mufu_per_s"] / peaks["mufu_ips"]
        print(json.dumps(rec), flush=True)
    if a.kernel in ("getframe", "all"):
        hdr, tok, base, byts = pkg.ops.synth_clip(a.w, a.h, 1234, F)
        d_hdr = torch.from_numpy(hdr).to(dev); d_tok = torch.from_numpy(tok).to(dev); d_base = torch.from_numpy(base).to(dev)
        mvx = torch.empty(F, h4, w4, dtype=torch.int16, device=dev); mvy = torch.empty_like(mvx)
        dp = torch.empty(F, h4, w4, dtype=torch.uint8, device=dev)
        fn = lambda: check(L.cds_getframe_dev(a.w, a.h, F, vp(d_hdr), vp(d_tok), vp(d_base), vp(mvx), vp(mvy), vp(dp), st), "getframe")
        med, mn = timeit(fn, a.iters, a.warmup, flush)
        byts_alg = tok.nbytes + hdr.nbytes + F * h4 * w4 * 5
        print(json.dumps({"kernel": "getframe", "frames": F, "ms_median": med, "ms_min": mn, "alg_bytes": byts_alg,
                          "GBps": byts_alg / (med * 1e-3) / 1e9, "frames_per_s": F / (med * 1e-3)}), flush=True)
    print(json.dumps({"gpu_launches": pkg.take_launch_count()}))


if __name__ == "__main__":
    main()
