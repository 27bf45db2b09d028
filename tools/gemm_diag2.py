import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib
l = lib()
def run(A, W, bias=None, act=0, bn=0):
    M, K = A.shape; N = W.shape[0]
    D = np.full((M, N), 7.0, np.float32); ms = C.c_float()
    rc = l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(bias) if bias is not None else None, fptr(D), M, N, K, act, bn, 1, C.byref(ms))
    assert rc == 0, l.prisma_last_error()
    return D
np.set_printoptions(linewidth=200, precision=4, suppress=True)
A = np.ones((128, 64), np.float32); W = np.ones((128, 64), np.float32)
D = run(A, W); print("ones:", D[0, :8], D[127, 120:], "unique", np.unique(D)[:10])
A = np.zeros((128, 64), np.float32); A[:, 0] = np.arange(128); W = np.zeros((128, 64), np.float32); W[:, 0] = 1
D = run(A, W); print("row id:", D[:6, 0], D[60:66, 5], "ok", np.array_equal(D, np.repeat(np.arange(128, dtype=np.float32)[:, None], 128, 1)))
A = np.zeros((128, 64), np.float32); A[:, 0] = 1; W = np.zeros((128, 64), np.float32); W[:, 0] = np.arange(128)
D = run(A, W); print("col id:", D[0, :6], D[5, 60:66], "ok", np.array_equal(D, np.repeat(np.arange(128, dtype=np.float32)[None, :], 128, 0)))
A = np.zeros((128, 64), np.float32); W = np.zeros((128, 64), np.float32)
for k in range(64): A[:, k] = k; 
W[:, :] = 0; W[:, 3] = 1
D = run(A, W); print("k sel 3:", np.unique(D))
rng = np.random.default_rng(0)
A = rng.standard_normal((128, 64), dtype=np.float32); W = rng.standard_normal((128, 64), dtype=np.float32)
D = run(A, W); ref = A.astype(np.float16).astype(np.float32) @ W.astype(np.float16).astype(np.float32).T
print("rand:", D[0, :6], ref[0, :6], "nan", np.isnan(D).sum(), "A nan", np.isnan(A).sum())
A2 = np.ascontiguousarray(A); print(A.flags['C_CONTIGUOUS'], A.dtype, A.strides)
