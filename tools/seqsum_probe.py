#!/usr/bin/env python3
"""Run cds_probe_seq_sum on a few term distributions (time the kernel with `ncu --metrics gpu__time_duration.sum`)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import compressd_domain_saliencyprediction_b200 as cds
L = cds.lib()
r = np.random.default_rng(1)
n = 129600
cases = {
    "uniform01_all": (r.random(n, dtype=np.float32), np.ones(n, np.uint8)),
    "uniform01_sparse3pct": (r.random(n, dtype=np.float32), (r.random(n) < 0.03).astype(np.uint8)),
    "pan_clamped": (np.minimum((r.integers(20, 40, n) * 0.25).astype(np.float32), np.float32(7.3125001)), (r.random(n) < 0.5).astype(np.uint8)),
    "mixed_sign": ((r.integers(-40, 41, n) * 0.25).astype(np.float32), (r.random(n) < 0.5).astype(np.uint8)),
}
for name, (v, sel) in cases.items():
    s = C.c_float(0); cnt = C.c_int(0)
    L.cds_probe_seq_sum(v.ctypes.data_as(C.c_void_p), sel.ctypes.data_as(C.c_void_p), n, C.byref(s), C.byref(cnt))
    print(name, s.value, cnt.value)
