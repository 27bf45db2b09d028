"""Which unit paces the shifted-row GEMM on the RAFT update-block shapes (8 waves per launch)?  Run under PRISMA_GEMM_DBG=
0 full / 2 no epilogue work / 3 loads only / 5 MMAs + epilogue without loads / 6 MMAs only."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib
l = lib()
rng = np.random.default_rng(0)
def run(M, N, K, bn, act, iters=30):
    A = np.ones((M, K), np.float32); A[::7] = 0.5
    W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, act, bn, iters, C.byref(ms)) == 0, l.prisma_last_error()
    return ms.value * 1e3
M = 148 * 8 * 128
for (N, K, bn) in [(128, 1920, 0), (128, 1920, 384), (256, 1920, 0), (256, 1920, 256), (64, 1152, 0), (256, 1024, 0)]:
    t = run(M, N, K, bn, -2)
    print("M %6d N %4d K %4d bn %3d : f16-out %7.1f us  = %6.1f clk per 64-wide K block per tile at 1.8 GHz, %6.0f TF/s" %
          (M, N, K, bn, t, t * 1e-6 * 1.8e9 / (8 * (K // 64)), 2.0 * M * N * K / t * 1e-6), flush=True)
