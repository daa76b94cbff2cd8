"""Debug helper: compare per-feature 200x200 intermediates of one picture between the CUDA path and the oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as o
import compressd_domain_saliencyprediction_b200 as cds

w, h, fps, seed = 416, 240, 10.0, 1234
F = int(sys.argv[1]) if len(sys.argv) > 1 else 22
MEAN9 = np.full(9, 0.3); STD9 = np.full(9, 0.2)
hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
maps = [o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]]) for f in range(F)]
mvx = np.stack([m[0] for m in maps]); mvy = np.stack([m[1] for m in maps]); dep = np.stack([m[2] for m in maps])
SV, coef, rho, gamma, labels = o.synth_model(300)
SV = SV * 0.5 + 1.0
gm = cds.ops.SvmModel(SVs=SV, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)
want, wcam, wfeat = o.orc_process_clip(w, h, fps, mvx, mvy, dep, byts, (SV, coef, rho, gamma, labels), MEAN9, STD9, use_rbf=1, want_feat=True)
clip = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=8)
got, cam = clip.process(0, hdr, tok, base, byts)
for f in range(F):
    print("frame", f, "map err", np.max(np.abs(got[f] - np.nan_to_num(want[f]))), "cam", cam[f], wcam[f])
fb = (F - 1) % 8
xs = clip.debug("xsvm", fb, (40000, 9))
feats = xs * STD9.astype(np.float32) + MEAN9.astype(np.float32)
names = ["mv_ori", "mv_t", "mv_s", "bit_ori", "bit_t", "bit_s", "dep_ori", "dep_t", "dep_s"]
for i in range(9):
    wi = np.nan_to_num(wfeat[F - 1, i]).reshape(-1)
    print(names[i], "feat err", np.max(np.abs(feats[:, i] - wi)), "range", wi.min(), wi.max())
print("xf", clip.debug("xf", fb, (3, 4)))
print("stats", clip.debug("stats", fb, (9, 4)))
