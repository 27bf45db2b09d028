"""One step of the bench workload (24 frames of 1080p: depth_anything ViT-L passes + flow_raft pairs through the clip APIs)
between cudaProfilerStart / Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum` (the launch list whose
per-kernel SHARES bench.py's roofline block is checked against)."""
import os, sys
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
def step(cont):
    da.infer_clip(clip, pass_frames=B)
    raft.infer_clip(clip, continue_clip=cont, want_flow=False, want_rgb=True)
step(False); step(True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
