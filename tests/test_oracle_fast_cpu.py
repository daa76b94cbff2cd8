"""Pins the fast-version oracle (oracle/cds_oracle_fast.c) against the reference's only whole-path golden outputs:
HM_16.0_features/bin/HMdata/0..38.png, the maps the authors' decoder wrote for BasketballDrill pictures 0..38
(tests/golden/fast_BasketballDrill.npz, made by make_golden_fast_path.py from the PNGs and the reference hooks' syntax dump).

Measured: 36 of the 39 maps byte-identical, the other three differ in ONE pixel each by 1 LSB (a value within 3 float ulps of a
rounding tie; a 1-ulp change of a single exp() weight moves them, i.e. the residue is at the level of the MSVC CRT's expf,
which cannot be resolved here).  Any other summation order or division form tried produces 1-13 differing pixels per picture.
"""
import os

import numpy as np

import _oracle as o

GOLD = os.path.join(o.ROOT, "tests", "golden")


def test_fast_oracle_reproduces_the_39_golden_pngs():
    d = np.load(os.path.join(GOLD, "fast_BasketballDrill.npz"))
    w, h = int(d["w"]), int(d["h"])
    fo = o.FastOracle(w, h, float(d["fps"]))
    assert (fo.PS, fo.PT) == (15, 35)                                # cvRound(50*0.3), cvRound(50*0.7)  salmat.cpp:18-19
    exact, off = 0, 0
    for f in range(d["png"].shape[0]):
        mvx, mvy, dep = o.orc_getframe(w, h, d["hdr"][f], d["tok"][d["tok_base"][f]:d["tok_base"][f + 1]])
        u8, f32, cam, _ = fo.frame(mvx, mvy, dep, d["ctu_bytes"][f])
        diff = np.abs(u8.astype(np.int16) - d["png"][f].astype(np.int16))
        n = int((diff > 0).sum())
        assert diff.max() <= 1 and n <= 1, (f, int(diff.max()), n)
        exact += n == 0; off += n
        assert tuple(cam) == (0, 0)                                  # the camera-motion branch never fires in these pictures
        assert f32.min() == 0.0 and f32.max() <= 1.0
    assert exact >= 36 and off <= 3, (exact, off)
    fo.close()


def test_cv_gaussian_kernel_close_to_current_opencv():
    """getGaussianKernel(n, sigma, CV_32F) as OpenCV 2.4.9 computes it (taps rounded to float before they are summed).  The cv2 4.x
    in this image keeps the taps in double until the end, so the two agree to a float ulp, not bit for bit; against the golden
    PNGs the 2.4.9 form gives 3 differing pixels in 39 pictures, the 4.x form 1-7 per picture."""
    cv2 = __import__("pytest").importorskip("cv2")
    import ctypes as C
    L = o.oracle()
    L.orc_cv_gaussian_kernel.argtypes = [C.c_int, C.c_double, C.c_void_p]
    for n, s in ((17, 8.0), (33, 16.0), (73, 36.0), (145, 72.0)):
        k = np.zeros(n, np.float32)
        L.orc_cv_gaussian_kernel(n, s, k.ctypes.data)
        r = cv2.getGaussianKernel(n, s, cv2.CV_32F).reshape(-1)
        assert np.max(np.abs(k - r) / r) < 2.5e-7 and abs(float(k.sum()) - 1.0) < 1e-6


def test_cv_gaussian_blur_close_to_current_opencv():
    """cv2 4.x vectorises the blur differently (last-bit differences), so this is a tolerance check of border handling and taps."""
    cv2 = __import__("pytest").importorskip("cv2")
    import ctypes as C
    L = o.oracle()
    L.orc_cv_gaussian_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    a = np.random.default_rng(3).random((60, 104), dtype=np.float32)
    out = np.zeros_like(a); tmp = np.zeros_like(a)
    L.orc_cv_gaussian_blur(a.ctypes.data, 60, 104, 17, 8.0, out.ctypes.data, tmp.ctypes.data)
    ref = cv2.GaussianBlur(a, (17, 17), 8.0, borderType=cv2.BORDER_CONSTANT)
    assert np.max(np.abs(out - ref)) < 2e-6
