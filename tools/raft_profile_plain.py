"""One full and two video passes of the RAFT engine at 1080p (for an ncu launch list: no per-step events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PRISMA_NO_GRAPH"] = "1"   # ncu serialises kernels anyway; direct launches keep kernel names per step
from prisma_b200.flow import RaftFlowEngine
from prisma_b200.seeded_weights import make_raft_weights
from prisma_b200.synthetic import synthetic_frame
eng = RaftFlowEngine(make_raft_weights(0), iterations=12)
f0, f1 = synthetic_frame(1080, 1920, 0), synthetic_frame(1080, 1920, 1)
eng.infer_pair(f0, f1)
eng.infer_pair(f0, f1, reuse_prev=True)
