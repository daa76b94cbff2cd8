"""Compare the tensor-core V pass with the FP32 one on the same clip: python tools/vpass_check.py [w h]  (runs itself twice)."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child(w, h, out):
    import compressd_domain_saliencyprediction_b200 as cds
    F = 3
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 7, F)
    clip = cds.SaliencyClip(w, h, 10.0, use_rbf=False, max_batch=4)
    clip.process(0, hdr, tok, base, byts)
    np.savez(out, stats=clip.debug("stats", F - 1, (9, 4)), y=clip.debug("ymaps", F - 1, (9, h, w)))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); sys.exit(0)
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (416, 240)
    for impl in ("fp32", "tc"):
        subprocess.check_call([sys.executable, __file__, "--child", str(w), str(h), f"/tmp/vp_{impl}.npz"], env=dict(os.environ, CDS_VPASS_IMPL=impl))
    a, b = np.load("/tmp/vp_fp32.npz"), np.load("/tmp/vp_tc.npz")
    print("stats fp32\n", a["stats"][:, :2], "\nstats tc\n", b["stats"][:, :2])
    ya, yb = a["y"], b["y"]
    for m in range(9):
        d = np.abs(ya[m] - yb[m]); print("map", m, "max |diff|", d.max(), "max |y|", np.abs(ya[m]).max(), "rows with diff>1e-5:", np.unique(np.where(d > 1e-5)[0])[:12], "cols:", np.unique(np.where(d > 1e-5)[1])[:12])
    m = 0
    print("fp32 y[0, 100:108, 200:204]\n", ya[m, 100:108, 200:204], "\ntc\n", yb[m, 100:108, 200:204])
