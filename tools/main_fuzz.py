#!/usr/bin/env python3
"""Randomised geometry / rate / batch sweep of the nine-feature + RBF-SVM path, GPU vs oracle within 1e-4 (diagnostic).
usage: main_fuzz.py [cases] [seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as o
import compressd_domain_saliencyprediction_b200 as cds

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
r = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
MEAN9, STD9 = np.full(9, 0.3), np.full(9, 0.2)
SV, coef, rho, gamma, labels = o.synth_model(200); SV = SV * 0.5 + 1.0
gm = cds.ops.SvmModel(SVs=SV, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)
o.oracle().orc_set_threads(16)
bad = 0; t0 = time.time()
for i in range(cases):
    w = int(r.integers(9, 90)) * 8; h = int(r.integers(9, 70)) * 8
    fps = float(r.choice([7.0, 10.0, 13.0, 24.0])); F = int(r.integers(2, 9)); batch = int(r.integers(1, 6)); seed = int(r.integers(0, 10000))
    rbf = int(r.integers(0, 4) > 0)
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
    maps = [o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]]) for f in range(F)]
    mvx = np.stack([m[0] for m in maps]); mvy = np.stack([m[1] for m in maps]); dep = np.stack([m[2] for m in maps])
    want, wcam, _ = o.orc_process_clip(w, h, fps, mvx, mvy, dep, byts, (SV, coef, rho, gamma, labels), MEAN9, STD9, first_out=0, use_rbf=rbf)
    try:
        clip = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, use_rbf=bool(rbf), max_batch=batch)
        got, cam = clip.process(0, hdr, tok, base, byts); clip.close()
        err = float(np.max(np.abs(got - np.nan_to_num(want))))
        ok = np.array_equal(cam, wcam) and err <= 1e-4
        msg = f"err {err:.2e}" + ("" if np.array_equal(cam, wcam) else " cam differs")
    except Exception as e:
        ok = False; msg = repr(e)
    bad += not ok
    print(f"{i:3d} {w}x{h} fps {fps} F {F} batch {batch} rbf {rbf} seed {seed}: {'ok ' if ok else 'FAIL '}{msg}", flush=True)
print(f"{cases - bad} of {cases} cases within 1e-4 in {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
