#!/usr/bin/env python3
"""Bitstream -> maps throughput of the feeder (N parse-only decoder processes -> one GPU-owning process).
usage: feeder_bench.py [N] [fast|svm] [stream.bin WxH]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import compressd_domain_saliencyprediction_b200 as cds

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
stream = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "oracle", "_ref", "streams", "BasketballDrill.bin")
w, h = map(int, (sys.argv[4] if len(sys.argv) > 4 else "832x480").split("x"))
kw = dict(fast=True, out_u8=True)
if mode == "svm":
    import _oracle as o
    SV, coef, rho, gamma, labels = o.synth_model(4096)
    kw = dict(model=cds.ops.SvmModel(SVs=SV * 0.5 + 1.0, sv_coef=coef, rho=rho, gamma=gamma, Label=labels), meanVec=np.full(9, 0.3), stdVec=np.full(9, 0.2), out_u8=True)
dumper = os.path.join(ROOT, "oracle", "_ref", "hm_syntax_dump")
cds.feeder.run_streams(dumper, [(stream, w, h)], **kw)         # warm-up: library, CUDA context
r = cds.feeder.run_streams(dumper, [(stream, w, h)] * n, **kw)
print(f"{n} parse-only decoder processes -> 1 GPU process ({mode}, {os.path.basename(stream)} {w}x{h}): {r['pictures']} pictures in {r['seconds']:.2f} s = "
      f"{r['pictures_per_s']:.0f} pictures/s (host cores: {os.cpu_count()})")
