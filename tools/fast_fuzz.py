#!/usr/bin/env python3
"""Randomised geometry / rate / batch sweep of the fast version, GPU vs oracle, bit-exact (diagnostic).  usage: fast_fuzz.py [cases] [seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as o
import compressd_domain_saliencyprediction_b200 as cds

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
r = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
t0 = time.time()
for i in range(cases):
    w = int(r.integers(4, 60)) * 16; h = int(r.integers(8, 90)) * 8
    fps = float(r.choice([7.0, 10.0, 15.0, 24.0, 25.0, 29.97, 30.0, 50.0, 60.0]))
    F = int(r.integers(1, 14)); batch = int(r.integers(1, 9)); seed = int(r.integers(0, 10000))
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
    fo = o.FastOracle(w, h, fps)
    want = []; wcam = []
    for f in range(F):
        m = o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]])
        u8, f32, cam, _ = fo.frame(m[0], m[1], m[2], byts[f]); want.append(f32); wcam.append(cam.copy())
    fo.close()
    try:
        clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=False, max_batch=batch)
        got, cam = clip.process(0, hdr, tok, base, byts); clip.close()
        ok = np.array_equal(cam, np.stack(wcam)) and got.tobytes() == np.stack(want).tobytes()
        msg = "" if ok else f"max|d| {np.nanmax(np.abs(got - np.stack(want))):.3e} cam {np.array_equal(cam, np.stack(wcam))}"
    except Exception as e:
        ok = False; msg = repr(e)
    bad += not ok
    print(f"{i:3d} {w}x{h} fps {fps} F {F} batch {batch} seed {seed}: {'ok' if ok else 'FAIL ' + msg}", flush=True)
print(f"{cases - bad} of {cases} cases bit-identical in {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
