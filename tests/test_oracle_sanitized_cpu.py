"""The oracle suite once more on AddressSanitizer + UndefinedBehaviorSanitizer builds of the checker: oracle/cds_oracle.c,
cds_oracle_fast.c and — where /root/reference exists — the reference's own getFrame (code_mat_h.h + func.h, whose index
expressions are out of range on illegal streams: func.h:150,165,180; code_mat_h.h:396,524,557), computecontrast5.cpp and
libsvm-3.17.  A finding aborts the child process, so a green run means: on the golden streams, the legal
fuzz streams and the whole normal / fast paths the checker never reads or writes out of bounds and hits no UB the
sanitizers know.  The child is a fresh interpreter with the ASan runtime preloaded; the parent only reads its verdict."""
import os
import subprocess
import sys

import pytest

import _oracle as o


SANCC = "/usr/bin/gcc"      # the distribution compiler ships the sanitizer runtimes (oracle/Makefile: SANCC / SANCXX)


def _runtime(name):
    if not os.path.exists(SANCC):
        return None
    p = subprocess.run([SANCC, f"-print-file-name={name}"], capture_output=True, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


def test_oracle_suite_under_asan_ubsan():
    asan = _runtime("libasan.so")
    if asan is None:
        pytest.skip("no libasan.so next to gcc")
    r = subprocess.run(["make", "-s", "-C", o.ORACLE_DIR, "sanitize"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    env = dict(os.environ, CDS_ORACLE_SAN="1", LD_PRELOAD=asan,
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=1:allocator_may_return_null=1",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    tests = ["tests/test_oracle_cpu.py", "tests/test_oracle_fast_cpu.py",
             "tests/test_matlab_semantics_cpu.py::test_whole_normal_version_second_opinion",
             "tests/test_matlab_semantics_cpu.py::test_camera_motion_vs_literal",
             "tests/test_matlab_semantics_cpu.py::test_cu_contrast5_vs_literal"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider"] + tests, cwd=o.ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    out = r.stdout + r.stderr
    assert "AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]
    assert r.returncode == 0, out[-4000:]
    assert " passed" in out


def test_sanitizer_is_armed():
    """Control: a deliberate out-of-bounds read inside the instrumented oracle must abort the child (otherwise a green
    run of the suite above would prove nothing)."""
    asan = _runtime("libasan.so")
    if asan is None:
        pytest.skip("no libasan.so next to gcc")
    assert subprocess.run(["make", "-s", "-C", o.ORACLE_DIR, "sanitize"], capture_output=True).returncode == 0
    env = dict(os.environ, CDS_ORACLE_SAN="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=0")
    code = ("import numpy as np, _oracle as o\n"
            "a = np.ones(16); out = np.zeros(64)\n"
            "o.oracle().orc_immove(a.ctypes.data, 8, 8, 0, 0, out.ctypes.data)\n"      # claims 64 inputs, 16 exist
            "print('survived')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.join(o.ROOT, "tests"), env=env, capture_output=True, text=True, timeout=300)
    assert "AddressSanitizer" in r.stderr and "survived" not in r.stdout, (r.stdout, r.stderr[-2000:])
