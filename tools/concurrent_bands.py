"""depth_anything and flow_raft over the same 1080p clip: one after the other vs concurrently (two host threads, one engine and
one CUDA stream each).  Wall-clock frames/s through the public clip APIs, pinned host buffers (bench.py's e2e leg)."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prisma_b200.depth import DepthAnythingEngine, pinned_empty
from prisma_b200.flow import RaftFlowEngine
from prisma_b200.seeded_weights import make_da_weights, make_raft_weights
from prisma_b200.synthetic import synthetic_frame
H, W, N, B = 1080, 1920, 24, 12
da = DepthAnythingEngine("vitl", make_da_weights("vitl", 0))
raft = RaftFlowEngine(make_raft_weights(0), iterations=12, scale=0.75)
base = [synthetic_frame(H, W, t) for t in range(4)]
clip = pinned_empty((N, H, W, 3), np.uint8)
clip[...] = np.stack([np.roll(base[i % 4], 7 * (i // 4), axis=1) for i in range(N)])
hs, ws = raft.out_size(H, W)
od = pinned_empty((N, H, W, 3), np.uint8)
of = {"fwd_rgb": pinned_empty((N, hs, ws, 3), np.uint8), "bwd_rgb": pinned_empty((N, hs, ws, 3), np.uint8)}
def d(): da.infer_clip(clip, pass_frames=B, out_rgb=od)
def f(cont=True): raft.infer_clip(clip, continue_clip=cont, want_flow=False, want_rgb=True, out=of)
d(); f(False); d(); f()
def seq(steps):
    t0 = time.perf_counter()
    for _ in range(steps): d(); f()
    torch.cuda.synchronize(); return time.perf_counter() - t0
def par(steps):
    t0 = time.perf_counter()
    for _ in range(steps):
        a = threading.Thread(target=d); b = threading.Thread(target=f)
        a.start(); b.start(); a.join(); b.join()
    torch.cuda.synchronize(); return time.perf_counter() - t0
def par_free(steps):   # each band runs its own loop: no per-step join
    def loop(fn):
        for _ in range(steps): fn()
    t0 = time.perf_counter()
    a = threading.Thread(target=loop, args=(d,)); b = threading.Thread(target=loop, args=(f,))
    a.start(); b.start(); a.join(); b.join()
    torch.cuda.synchronize(); return time.perf_counter() - t0
for name, fn in (("sequential", seq), ("concurrent (join per step)", par), ("concurrent (free running)", par_free), ("sequential", seq)):
    fn(2); dt = fn(8)
    print(f"{name}: {8 * N / dt:.1f} frames/s ({1e3 * dt / 8:.1f} ms per step of {N} frames)", flush=True)
