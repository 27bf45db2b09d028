"""Phase cycle counters of one softmax warp of the attention kernel (build with PRISMA_NVCC_EXTRA=-DPRISMA_ATTN_PROFILE, run
with PRISMA_ATTN_PROF=1) and the kernel's sustained TFLOP/s at the ViT-L 1080p shape (2443 tokens, 16 heads x 4 frames)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib
l = lib()
rng = np.random.default_rng(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2443
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 64
qkv = rng.standard_normal((T, 3 * heads * 64), dtype=np.float32)
out = np.empty((T, heads * 64), np.float32); ms = C.c_float()
for it in (3, 20):
    assert l.prisma_debug_attention(0, fptr(qkv), fptr(out), T, heads, it, C.byref(ms)) == 0, l.prisma_last_error()
    print("attn T", T, "heads", heads, "ms", ms.value, "TF", 4.0 * heads * T * T * 64 / ms.value / 1e9, flush=True)
# reference check on a few heads (fp32 softmax of the fp16-rounded operands)
q = qkv.astype(np.float16).astype(np.float32)
err = 0.0
for h in (0, heads - 1):
    Q = (q[:, h * 64:(h + 1) * 64] * 0.125).astype(np.float16).astype(np.float32)
    K = q[:, heads * 64 + h * 64: heads * 64 + (h + 1) * 64]; V = q[:, 2 * heads * 64 + h * 64: 2 * heads * 64 + (h + 1) * 64]
    S = Q @ K.T; P = np.exp(S - S.max(1, keepdims=True)); O = (P / P.sum(1, keepdims=True)) @ V
    err = max(err, float(np.abs(O - out[:, h * 64:(h + 1) * 64]).max()))
print("max abs err vs fp32 softmax:", err)
