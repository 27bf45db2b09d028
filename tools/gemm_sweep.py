import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib
l = lib()
rng = np.random.default_rng(0)
def run(M, N, K, bn, act, iters=30):
    A = rng.standard_normal((M, K), dtype=np.float32); W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, act, bn, iters, C.byref(ms)) == 0, l.prisma_last_error()
    return ms.value * 1e3
print("us per launch (back-to-back): rows = K, cols = [f32-out, f16-out, no-out]")
for (M, N, bn) in [(2443, 3072, 256), (2443, 3072, 512), (9772, 3072, 256), (9772, 3072, 512), (9772, 1024, 256), (9772, 1024, 512), (8192, 8192, 256), (8192, 8192, 512)]:
    print("M", M, "N", N, "bn", bn)
    for K in ((1024, 4096) if M < 8192 else (8192,)):
        print("   K=%5d  %7.1f %7.1f %7.1f" % (K, run(M, N, K, bn, 0), run(M, N, K, bn, -2), run(M, N, K, bn, -1)))
