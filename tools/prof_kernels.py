"""A few launches of the two tensor-core kernels at ViT-L shapes, for ncu (run under `ncu ... python tools/prof_kernels.py`)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib
l = lib()
rng = np.random.default_rng(0)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("gemm", "both"):
    # the encoder linears of a 12-frame pass (M = 12 x 2443): fc1 with fp16 output (act -2), fc2 with the in-place fp32
    # residual epilogue (act -3); bn 0 = the engine's own tile choice (CTA pairs, 256-wide tiles)
    for (M, N, K, act) in [(29316, 4096, 1024, -2), (29316, 1024, 4096, -3)]:
        A = rng.standard_normal((M, K), dtype=np.float32)
        W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
        b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
        assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, act, 0, 3, C.byref(ms)) == 0
        print("gemm", M, N, K, act, "ms", ms.value, "TF", 2.0 * M * N * K / ms.value / 1e9)
if which in ("attn", "both"):
    T, heads = 2443, 16 * 4   # heads x frames: the grid of a 4-frame pass (the harness is single-image)
    qkv = rng.standard_normal((T, 3 * heads * 64), dtype=np.float32)
    out = np.empty((T, heads * 64), np.float32); ms = C.c_float()
    assert l.prisma_debug_attention(0, fptr(qkv), fptr(out), T, heads, 3, C.byref(ms)) == 0
    print("attn ms", ms.value, "TF", 4.0 * heads * T * T * 64 / ms.value / 1e9)
if which in ("corr",):
    h8, w8 = 102, 180
    fm1 = rng.standard_normal((2, 256, h8, w8), dtype=np.float32); fm2 = rng.standard_normal((2, 256, h8, w8), dtype=np.float32)
    h = C.c_void_p(); ms = C.c_float()
    assert l.prisma_flowcorr_create(0, 2, h8, w8, C.byref(h)) == 0, l.prisma_last_error()
    assert l.prisma_flowcorr_set_fmaps(h, fptr(fm1), fptr(fm2)) == 0
    assert l.prisma_flowcorr_build(h, 2, C.byref(ms)) == 0
    work = (C.c_double * 2)(); l.prisma_flowcorr_work(h, work)
    print("corr build ms", ms.value, "GB/s", work[1] / ms.value / 1e6, "TF", work[0] / ms.value / 1e9)
if which in ("gemm_tma",):
    # fc1 of a 12-frame pass through the TMA-store epilogue (bias + GELU in registers, bulk stores): act -6
    M, N, K = 29316, 4096, 1024
    A = rng.standard_normal((M, K), dtype=np.float32)
    W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, -6, 0, 3, C.byref(ms)) == 0
    print("gemm fc1 tma-store", M, N, K, "ms", ms.value, "TF", 2.0 * M * N * K / ms.value / 1e9)
if which in ("gemm_swap",):
    # the RAFT GRU q-conv size class (N = 128, K = 1920, 8 waves) on the transposed tiles
    M, N, K = 148 * 8 * 128, 128, 1920
    A = np.ones((M, K), np.float32); A[::7] = 0.5
    W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, -2, 0, 3, C.byref(ms)) == 0
    print("gemm swap", M, N, K, "ms", ms.value, "TF", 2.0 * M * N * K / ms.value / 1e9)
if which in ("gemm_red",):
    # fc2 of a 12-frame pass through the TMA reduce-add epilogue (x += acc + bias in place, the add done by the L2): act -8
    M, N, K = 29316, 1024, 4096
    A = rng.standard_normal((M, K), dtype=np.float32)
    W = (rng.standard_normal((N, K), dtype=np.float32) / 64).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, -8, 0, 3, C.byref(ms)) == 0
    print("gemm fc2 tma reduce-add", M, N, K, "ms", ms.value, "TF", 2.0 * M * N * K / ms.value / 1e9)
