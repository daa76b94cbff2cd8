#!/usr/bin/env python3
"""Instruction-mix probes on the device (FFMA / FFMA2 co-issue with MUFU.EX2 and LDS); prints JSON lines."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import compressd_domain_saliencyprediction_b200 as pkg
L = pkg.lib()
L.cds_probe_mix.restype = C.c_int
L.cds_probe_mix.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
names = {0: "11 FFMA(+FADD) : 1 MUFU", 1: "pure FFMA2", 2: "10 FFMA2 : 2 MUFU (packed pair)", 3: "11 FFMA : 1 MUFU : LDS.128"}
for which in (0, 1, 2, 3):
    for ctas in (2, 4, 8):
        a, b = C.c_double(), C.c_double()
        rc = L.cds_probe_mix(which, ctas, C.byref(a), C.byref(b))
        print(json.dumps({"mix": names[which], "ctas_per_sm": ctas, "rc": rc, "fma_lane_ops_per_s": a.value, "mufu_per_s": b.value}), flush=True)
