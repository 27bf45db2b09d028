"""Clock / board power while one kernel runs back to back (B200 is power-capped under these kernels: fps = P / energy-per-frame).
usage: python tools/power_probe.py [gemm|gemm2|attn|pass]..."""
import ctypes as C, os, subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200._lib import fptr, lib

l = lib()
rng = np.random.default_rng(0)


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,power.limit,temperature.gpu", "--format=csv,noheader,nounits",
                                   "-i", "0", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._r, daemon=True).start()

    def _r(self):
        for line in self.p.stdout:
            try:
                self.rows.append([float(c) for c in line.split(",")])
            except ValueError:
                pass

    def stop(self):
        self.p.terminate()
        a = np.array(self.rows[len(self.rows) // 3:]) if len(self.rows) > 3 else np.zeros((1, 4))
        return "clk %4.0f MHz  power %4.0f W (limit %4.0f)  temp %2.0f C  [%d samples]" % (*np.median(a, axis=0), len(self.rows))


def gemm(M, N, K, bn, act, secs):
    A = rng.standard_normal((M, K), dtype=np.float32); W = (rng.standard_normal((N, K), dtype=np.float32) / 32).astype(np.float32)
    b = np.zeros(N, np.float32); D = np.empty((M, N), np.float32); ms = C.c_float()
    l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, act, bn, 20, C.byref(ms))
    iters = max(int(secs * 1e3 / ms.value), 20)
    s = Sampler()
    assert l.prisma_debug_gemm(0, fptr(A), fptr(W), fptr(b), fptr(D), M, N, K, act, bn, iters, C.byref(ms)) == 0
    print("gemm %dx%dx%d bn %d act %d: %.1f us  %.0f TF/s | %s" % (M, N, K, bn, act, ms.value * 1e3, 2.0 * M * N * K / ms.value / 1e9, s.stop()), flush=True)


def attn(T, heads, secs):
    qkv = (rng.standard_normal((T, 3 * heads * 64), dtype=np.float32)).astype(np.float32)
    out = np.empty((T, heads * 64), np.float32); ms = C.c_float()
    l.prisma_debug_attention(0, fptr(qkv), fptr(out), T, heads, 20, C.byref(ms))
    iters = max(int(secs * 1e3 / ms.value), 20)
    s = Sampler()
    assert l.prisma_debug_attention(0, fptr(qkv), fptr(out), T, heads, iters, C.byref(ms)) == 0
    print("attention T %d heads %d: %.1f us  %.0f TF/s | %s" % (T, heads, ms.value * 1e3, 4.0 * heads * T * T * 64 / ms.value / 1e9, s.stop()), flush=True)


def full_pass(batch, secs):
    from prisma_b200.depth import DepthAnythingEngine
    from prisma_b200.seeded_weights import make_da_weights
    eng = DepthAnythingEngine("vitl", make_da_weights("vitl", 0))
    ms = eng.time_resident(720, 1280, 5, batch)
    s = Sampler()
    ms = eng.time_resident(720, 1280, max(int(secs * 1e3 / ms), 5), batch)
    print("full pass batch %d: %.2f ms  %.1f fps | %s" % (batch, ms, batch / ms * 1e3, s.stop()), flush=True)


which = sys.argv[1:] or ["gemm", "attn", "pass"]
secs = 3.0
if "idle" in which:
    s = Sampler(); time.sleep(2); print("idle |", s.stop())
if "gemm" in which:
    gemm(8192, 8192, 8192, 256, -2, secs)
    gemm(8192, 8192, 8192, 512, -2, secs)
    gemm(8192, 8192, 8192, 128, -2, secs)
    gemm(9772, 4096, 1024, 256, -2, secs)
    gemm(9772, 1024, 4096, 256, -3, secs)
if "attn" in which:
    attn(2443, 16 * 4, secs)    # heads x frames: same grid as a 4-frame pass (the harness is single-image, heads are independent)
    attn(2443, 16 * 12, secs)
if "pass" in which:
    full_pass(4, secs)
    full_pass(12, secs)
