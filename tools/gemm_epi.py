"""Epilogue cost of the encoder GEMM shapes: us per launch for [no-out, f16-out, f32-out, f32 residual in place]."""
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
M = int(sys.argv[1]) if len(sys.argv) > 1 else 9772
for (N, K) in [(1024, 1024), (1024, 4096), (3072, 1024), (4096, 1024)]:
    for bn in (256, 512, 128, 64):
        r = [run(M, N, K, bn, a) for a in (-1, -2, 0, -3)]
        gf = 2.0 * M * N * K / 1e9
        print("M %5d N %4d K %4d bn %3d : none %6.1f  f16 %6.1f  f32 %6.1f  resid %6.1f us   (%.0f TF/s at resid)" % (M, N, K, bn, *r, gf / r[3] / 1e3), flush=True)
