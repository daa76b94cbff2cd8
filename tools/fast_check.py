#!/usr/bin/env python3
"""Fast-version pipeline, GPU vs oracle, stage by stage (diagnostic; tests/test_gpu_fast.py is the gate).

usage: fast_check.py [fixture.npz | synth:WxH:frames[:seed]] [fps] [max_batch]
Prints, per picture, the first intermediate that differs (smoothed maps, ori / temporal / spatial features, weighted sum,
final map) with the largest absolute difference, then the number of differing u8 pixels."""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as o  # noqa: E402
import compressd_domain_saliencyprediction_b200 as cds  # noqa: E402

W9 = np.array([0.058, 0.171, -0.095, 0.068, -0.01, 0.509, -0.051, 0.154, 0.196], np.float32)


def load(spec):
    if spec.startswith("synth:"):
        parts = spec.split(":")
        w, h = map(int, parts[1].split("x")); F = int(parts[2]); seed = int(parts[3]) if len(parts) > 3 else 5
        hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
        return w, h, hdr, tok, base, byts
    d = np.load(spec)
    return int(d["w"]), int(d["h"]), d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"]


def main():
    spec = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "fast_BasketballDrill.npz")
    fps = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
    mb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    w, h, hdr, tok, base, byts = load(spec)
    F = hdr.shape[0]; h4, w4 = h // 4, w // 4
    fo = o.FastOracle(w, h, fps)
    clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=mb)
    clipf = cds.SaliencyClip(w, h, fps, fast=True, out_u8=False, max_batch=mb)
    lw, lh, gw, gh = (w + 63) // 64, (h + 63) // 64, (w + 31) // 32, (h + 31) // 32
    g32 = gw * gh
    bad = 0
    for f0 in range(0, F, mb):
        n = min(mb, F - f0)
        sl = slice(f0, f0 + n)
        tk = tok[base[f0]:base[f0 + n]]; bs = base[f0:f0 + n + 1] - base[f0]
        u8, cam = clip.process(f0, hdr[sl], tk, bs, byts[sl])
        f32, _ = clipf.process(f0, hdr[sl], tk, bs, byts[sl])
        for b in range(n):
            f = f0 + b
            mvx, mvy, dep = o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]])
            ou8, of32, ocam, feat = fo.frame(mvx, mvy, dep, byts[f], want_feat=True)
            msgs = []
            if not np.array_equal(cam[b], ocam): msgs.append(f"cam gpu {cam[b]} oracle {ocam}")
            tv = clip.debug("fast_tval", b, (3, g32)); sv = clip.debug("fast_sval", b, (3, g32))
            comb = clip.debug("fast_comb", b, (h4, w4))
            def up(a, gh_, gw_, k):  # replicate a block grid to cells
                return np.repeat(np.repeat(a[:gh_ * gw_].reshape(gh_, gw_), k, 0), k, 1)[:h4, :w4]
            cell = lambda m: feat[m][::4, ::4]
            for name, got, want in (("mvT", up(tv[2], gh, gw, 8), cell(1)), ("bitT", up(tv[0], lh, lw, 16), cell(4)), ("depT", up(tv[1], gh, gw, 8), cell(7)),
                                    ("mvS", up(sv[0], gh, gw, 8), cell(2)), ("bitS", up(sv[1], lh, lw, 16), cell(5)), ("depS", up(sv[2], gh, gw, 8), cell(8))):
                if not np.array_equal(got, want): msgs.append(f"{name} max|d| {np.max(np.abs(got - want)):.3e} ({int((got != want).sum())} cells)")
            s = feat[0] * W9[0] + feat[1] * W9[1]
            for i in range(2, 9): s = feat[i] * W9[i] + s
            wc = s[::4, ::4]
            if not np.array_equal(comb, wc): msgs.append(f"comb max|d| {np.max(np.abs(comb - wc)):.3e} ({int((comb != wc).sum())} cells)")
            if not np.array_equal(f32[b], of32): msgs.append(f"f32 max|d| {np.max(np.abs(f32[b] - of32)):.3e} ({int((f32[b] != of32).sum())} px)")
            nd = int((u8[b] != ou8).sum())
            if nd: msgs.append(f"u8 {nd} px differ (max {int(np.max(np.abs(u8[b].astype(int) - ou8.astype(int))))})")
            bad += bool(msgs)
            print(f"picture {f}: " + ("identical (u8, f32, cam, features)" if not msgs else "; ".join(msgs)), flush=True)
    print("RESULT", "ALL IDENTICAL" if bad == 0 else f"{bad} pictures differ")


if __name__ == "__main__":
    main()
