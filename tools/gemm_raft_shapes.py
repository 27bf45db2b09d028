"""Per-launch cost of the RAFT update-block GEMM shapes, back to back (no launch gaps): us per launch for the dense
[no-out, f16-out, f32-out] epilogues.  M = 37 888 rows (1080p x 0.75 at 1/8 resolution, both directions)."""
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
M = 37888
for (N, K) in [(256, 384), (256, 1920), (256, 2304), (128, 1920), (128, 2304), (64, 1152), (32, 512), (256, 256)]:
    for bn in (0, 256, 128):
        if bn and bn > 2 * N: continue
        r = [run(M, N, K, bn, a) for a in (-1, -2, 0, 3)]
        gf = 2.0 * M * N * K / 1e9
        print("M %5d N %4d K %4d bn %3d : none %6.1f  f16 %6.1f  f32 %6.1f  f32+sigmoid %6.1f us   (%.0f TF/s at f16)" % (M, N, K, bn, *r, gf / r[1] / 1e3), flush=True)
