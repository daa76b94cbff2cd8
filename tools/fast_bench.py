#!/usr/bin/env python3
"""Device-resident timing of the fast-version path (use_rbf = CDS_FUSION_FAST): pictures/s and ms per batch.
usage: fast_bench.py [WxH] [fps] [batch] [steps]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import compressd_domain_saliencyprediction_b200 as cds
from compressd_domain_saliencyprediction_b200._lib import check

w, h = map(int, (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
fps = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
L = cds.lib(); nctu = cds.ops.nctu(w, h)
clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=B)
F = (2 + steps) * B
hdr, tok, base, byts = cds.ops.synth_clip(w, h, 1234, F, first_poc=1)
dev = torch.device("cuda", 0)
d_hdr, d_tok, d_base, d_byts = (torch.from_numpy(x).to(dev) for x in (hdr, tok, base, byts))
out = torch.empty((steps * B, h, w), dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = C.c_void_p(stream.cuda_stream)
P = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
def step(i, nb=1):        # nb batches in ONE call: the library may overlap the halves of consecutive batches
    check(L.cds_clip_process_dev(clip.handle, i * B, nb * B, P(d_byts, i * B * nctu * 2), P(d_hdr, i * B * nctu * 16), P(d_tok), P(d_base, i * B * 8), P(out), None, st), "process")
step(0); step(1); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
step(2, steps)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"fast version {w}x{h} @{fps} fps, batch {B}: {ms:.3f} ms per batch, {B / ms * 1e3:.0f} pictures/s (device-resident, u8 maps)")

if os.environ.get("CDS_FAST_PROF") is not None:
    a = np.zeros(32, np.int64)
    check(L.cds_debug_copy(clip.handle, b"fast_prof", 0, a.ctypes.data_as(C.c_void_p), a.nbytes), "prof")
    print("cam_vote block (0,0) cycles: member sums", a[0], "Umean", a[2], "members", a[1], "first loop exact", a[3])
    print("thr job", os.environ["CDS_FAST_PROF"], "(cycles): stats", a[8], "sequential sum", a[9], "selected", a[10])
