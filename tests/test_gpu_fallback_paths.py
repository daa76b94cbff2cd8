"""The CUDA-core fallbacks behind the tensor-core kernels (FP32 SVM kernel, FFMA2 vertical FIR) stay parity-checked:
a child process with CDS_SVM_IMPL=fp32 / CDS_VPASS_IMPL=fp32 must reproduce the oracle like the default path does."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import compressd_domain_saliencyprediction_b200 as cds
d = np.load(sys.argv[2])
gm = cds.ops.SvmModel(SVs=d["SV"], sv_coef=d["coef"], rho=float(d["rho"]), gamma=float(d["gamma"]), Label=(1, -1))
clip = cds.SaliencyClip(int(d["w"]), int(d["h"]), 10.0, model=gm, meanVec=np.full(9, 0.3), stdVec=np.full(9, 0.2), max_batch=4)
maps, cam = clip.process(0, d["hdr"], d["tok"], d["base"], d["byts"])
np.savez(sys.argv[3], maps=maps, cam=cam)
'''


@pytest.mark.parametrize("env", [{"CDS_SVM_IMPL": "fp32"}, {"CDS_VPASS_IMPL": "fp32"}, {"CDS_SVM_POLY": "0"}])
def test_fallback_kernels_match_the_oracle(tmp_path, env):
    import compressd_domain_saliencyprediction_b200 as cds
    assert cds.device_ok()
    w, h, F = 416, 240, 6
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 404, F)
    SV, coef, rho, gamma, labels = o.synth_model(200)
    SV = SV * 0.5 + 1.0
    inp = tmp_path / "in.npz"; out = tmp_path / "out.npz"
    np.savez(inp, w=w, h=h, hdr=hdr, tok=tok, base=base, byts=byts, SV=SV, coef=coef, rho=rho, gamma=gamma)
    subprocess.check_call([sys.executable, "-c", CHILD, o.ROOT, str(inp), str(out)], env=dict(os.environ, **env), timeout=300)
    got = np.load(out)
    ints = [o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]]) for f in range(F)]
    mvx = np.stack([m[0] for m in ints]); mvy = np.stack([m[1] for m in ints]); dep = np.stack([m[2] for m in ints])
    want, wcam, _ = o.orc_process_clip(w, h, 10.0, mvx, mvy, dep, byts, (SV, coef, rho, gamma, labels), np.full(9, 0.3), np.full(9, 0.2))
    assert np.array_equal(got["cam"], wcam)
    assert np.max(np.abs(got["maps"] - np.nan_to_num(want))) <= 1e-4
