"""GPU diagnostic: sweep shapes x BLOCK_N through prisma_debug_gemm, print error / NaN statistics and timings."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib

l = lib()


def run(M, N, K, bn, act=0, iters=1, use_bias=True):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((M, K), dtype=np.float32)
    W = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N, dtype=np.float32)
    D = np.full((M, N), 7.0, np.float32)
    ms = C.c_float()
    rc = l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b) if use_bias else None, fptr(D), M, N, K, act, bn, iters, C.byref(ms))
    if rc != 0:
        print(f"M={M} N={N} K={K} bn={bn}: rc={rc} {l.prisma_last_error().decode()}")
        return
    ref = A.astype(np.float16).astype(np.float32) @ W.astype(np.float16).astype(np.float32).T + (b if use_bias else 0)
    nan = np.isnan(D)
    err = np.nanmax(np.abs(D - ref)) if not nan.all() else float("nan")
    tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0
    msg = f"M={M} N={N} K={K} bn={bn} bias={use_bias}: err={err:.3e} nan={int(nan.sum())}/{D.size} ms={ms.value:.4f} TF={tf:.1f}"
    if nan.any():
        r, c = np.where(nan)
        msg += f" nan rows[{r.min()},{r.max()}] cols[{c.min()},{c.max()}] uniq_cols={len(np.unique(c))} uniq_rows={len(np.unique(r))}"
    print(msg, flush=True)


if __name__ == "__main__":
    for bias in (True, False):
        for bn in (128, 256, 64, 32):
            run(128, 128, 64, bn, use_bias=bias)
    for bn in (0, 32, 64, 128, 256):
        run(256, 256, 256, bn)
        run(300, 384, 200, bn)
        run(2443, 1152, 384, bn)
    run(64, 64, 64, 0)
    run(40000, 256, 128, 0)
    # throughput at ViT-L shapes
    for (M, N, K) in [(2443, 3072, 1024), (2443, 1024, 1024), (2443, 4096, 1024), (2443, 1024, 4096), (4886, 4096, 1024), (8192, 8192, 8192)]:
        for bn in (128, 256):
            run(M, N, K, bn, iters=20)
