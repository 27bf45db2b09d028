import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib
l = lib()
rng = np.random.default_rng(0)
def run(M, N, K, bn, act, iters=50):
    A = rng.standard_normal((M, K), dtype=np.float32); W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, act, bn, iters, C.byref(ms)) == 0, l.prisma_last_error()
    return ms.value * 1e3
for (M, N, K, bn) in [(37888, 256, 384, 0), (37888, 256, 384, 256), (37888, 256, 1920, 0), (37888, 128, 1920, 0), (37888, 32, 512, 0), (18944, 256, 384, 256), (148 * 128, 256, 64, 256)]:
    print("M %5d N %4d K %4d bn %3d : f16-out %6.1f us   f32+sigmoid %6.1f us" % (M, N, K, bn, run(M, N, K, bn, -2), run(M, N, K, bn, 3)), flush=True)
