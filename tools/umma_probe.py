"""Check tcgen05.mma operand layouts against numpy: one 128 x 256 x 8 TF32 MMA with hand-built shared-memory images."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import compressd_domain_saliencyprediction_b200 as pkg
L = pkg.lib()
L.cds_probe_umma.restype = C.c_int
L.cds_probe_umma.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(0)
A = rng.integers(-4, 5, (128, 8)).astype(np.float32)        # small integers: exact in TF32
B = rng.integers(-4, 5, (8, 256)).astype(np.float32)        # B[k][n]
want = A @ B
# A: K-major, no swizzle: [k chunk of 4][row][4]
a_img = np.zeros((2, 128, 4), np.float32)
for k in range(8):
    a_img[k // 4, :, k % 4] = A[:, k]
def run(b_img, lbo, sbo, mn):
    D = np.zeros((128, 256), np.float32)
    rc = L.cds_probe_umma(a_img.ctypes.data, a_img.nbytes, b_img.ctypes.data, b_img.nbytes, 2048, 128, lbo, sbo, 0, mn, D.ctypes.data)
    assert rc == 0, L.cds_last_error()
    return D
# candidate 1: B K-major (like A): [k chunk][n][4]
b1 = np.zeros((2, 256, 4), np.float32)
for k in range(8):
    b1[k // 4, :, k % 4] = B[k, :]
D = run(b1, 4096, 128, 0)
print("B K-major       : max err", np.abs(D - want).max())
# candidate 2: B MN-major, [n group of 4][k 0..7][4 n]
b2 = np.zeros((64, 8, 4), np.float32)
for k in range(8):
    b2[:, k, :] = B[k, :].reshape(64, 4)
for lbo, sbo in ((8192, 128), (128, 8192), (128, 128), (16, 128), (128, 16)):
    D = run(b2, lbo, sbo, 1)
    e = np.abs(D - want)
    print(f"B MN-major lbo={lbo} sbo={sbo}: max err {e.max()}  wrong cols {np.unique(np.where(e > 0)[1])[:16]}  D[0,:8]={D[0,:8]} want={want[0,:8]}")
