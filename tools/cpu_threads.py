"""Which torch thread count gives the best CPU-oracle throughput on this host? (for bench.py's cpu_baseline)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import da as oda
from prisma_b200.seeded_weights import make_da_weights
from prisma_b200.synthetic import synthetic_frame
sd = make_da_weights("vitl", 0); f = synthetic_frame(720, 1280, 0)
for t in (8, 16, 32, 64, 128):
    if t > (os.cpu_count() or 1): break
    torch.set_num_threads(t)
    oda.da_infer(sd, f, "vitl")
    t0 = time.perf_counter(); oda.da_encode(oda.da_infer(sd, f, "vitl")); dt = time.perf_counter() - t0
    print("threads", t, "s/frame %.2f" % dt, flush=True)
