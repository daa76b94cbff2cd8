"""CPU-only: the C-ABI library loads, exports every symbol include/cds.h declares, the ctypes table
covers them all, and (without a GPU) compute entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as o


def _declared(header="cds.h"):
    txt = open(os.path.join(o.ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cds_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import compressd_domain_saliencyprediction_b200 as pkg
    L = pkg.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cds.h but not exported by libcds_b200.so"
    assert L._cds_missing == []


def test_synth_library_is_host_only_and_exports_its_header():
    """include/cds_synth.h is served by libcds_synth.so, which must not pull in the CUDA runtime or the product library
    (bench.py's reference arm and the CPU tests generate their inputs with it)."""
    import subprocess
    from compressd_domain_saliencyprediction_b200 import synth
    L = synth.synth_lib()
    for n in _declared("cds_synth.h"):
        assert hasattr(L, n), n
    so = os.path.join(os.path.dirname(synth.__file__), "libcds_synth.so")
    needed = subprocess.run(["objdump", "-p", so], capture_output=True, text=True).stdout
    assert "libcudart" not in needed and "libcuda" not in needed and "libcds_b200" not in needed


def test_ctypes_table_covers_header():
    from compressd_domain_saliencyprediction_b200 import _lib
    src = open(_lib.__file__).read()
    for n in _declared():
        assert f'"{n}"' in src, f"{n} has no ctypes signature in _lib.py"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import compressd_domain_saliencyprediction_b200 as pkg
    assert not pkg.device_ok()
    with pytest.raises(pkg.CdsError):
        pkg.ops.getframe(64, 64, np.zeros((1, 4), np.int32), np.zeros(4, np.int16))
    with pytest.raises(pkg.CdsError):
        pkg.ops.computecontrast5(np.zeros((52, 52)), np.ones((5, 5)), 2, 5, 48, 48)
    with pytest.raises(pkg.CdsError):
        pkg.SaliencyClip(416, 240, 25.0, use_rbf=False)
    assert b"no CPU fallback" in pkg.lib().cds_last_error() or b"CUDA" in pkg.lib().cds_last_error()


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, load or link it."""
    pk = os.path.join(o.ROOT, "compressd_domain_saliencyprediction_b200")
    for dp, _, fs in os.walk(pk):
        if "build" in dp.split(os.sep):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                for needle in ("libcds_oracle", "cds_oracle.h", "import _oracle", "orc_", "oracle/_ref", "oracle/_build"):
                    assert needle not in t, (os.path.join(dp, f), needle)


def test_synth_producer_is_deterministic_and_legal():
    from compressd_domain_saliencyprediction_b200 import ops
    a = ops.synth_frame(416, 240, 1234, 3)
    b = ops.synth_frame(416, 240, 1234, 3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    hdr, tok, byts = a
    assert hdr.shape == (28, 4) and byts.shape == (28,)
    assert int(hdr[-1, 0] + hdr[-1, 1:].sum()) == tok.size
    mx, my, dp = o.orc_getframe(416, 240, hdr, tok)
    assert dp.max() <= 3 and mx.shape == (60, 104)
