"""The environment-selected forms of the two dominant kernels give the SAME BITS (run with -m gpu on a B200).

The packed FP32 loop of the contrast kernel (FADD2 / FFMA2, the default) against the scalar FADD + FFMA loop
(CDS_CONTRAST_PACKED=0), and the packed partial sums of the persistent SVM kernel (default) against scalar adds
(CDS_SVM_POLY=0): each centre / instance adds the same terms in the same order with one rounding per operation, so the results must
be bitwise equal, not just close.  The switches are read once per process, so every form runs in its own child process."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import _oracle as o
import compressd_domain_saliencyprediction_b200 as pkg
assert pkg.device_ok(), pkg.lib().cds_last_error()
rng = np.random.default_rng(5)
out = {{}}
for i, (rows, cols, signed, patch) in enumerate([(60, 104, True, 48), (135, 240, False, 48), (67, 45, False, 24)]):
    a = rng.standard_normal((rows, cols)) if signed else rng.uniform(0, 1, (rows, cols))
    a[rng.uniform(size=a.shape) < 0.1] = a.min() if signed else 0.0          # zeros inside the map: the masked loop
    pad, W, ln, win = o.sudoku_inputs(a, 11.0, 48 // patch, 48)
    out["c%d" % i] = pkg.ops.computecontrast5(pad, W, ln, win, pad.shape[0] - 2 * ln, pad.shape[1] - 2 * ln)
SV, coef, rho, gamma, labels = o.synth_model(1000)
X = rng.standard_normal((40000 * 3 + 77, 9)) * 1.5
m = pkg.ops.SvmModel(SVs=SV, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)
out["dec"] = m.decision_values(X)[0]
np.savez({path!r}, **out)
"""


def _run(tmp_path, name, env):
    path = str(tmp_path / (name + ".npz"))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=o.ROOT, path=path)], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(path)


def test_packed_and_scalar_kernel_forms_are_bitwise_equal(tmp_path):
    default = _run(tmp_path, "default", {})
    scalar = _run(tmp_path, "scalar", {"CDS_CONTRAST_PACKED": "0", "CDS_SVM_POLY": "0"})
    for k in default.files:
        a, b = default[k], scalar[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes(), (k, float(np.max(np.abs(a - b))))
