"""mask band throughput at 1080p through the public API (host frames in, union masks out), 1..4 engines in flight."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prisma_b200.mask import SoloV2Lanes
from prisma_b200.seeded_weights import make_solo_weights
from prisma_b200.synthetic import synthetic_frame
sd = make_solo_weights("r101", 0)
base = [synthetic_frame(1080, 1920, t) for t in range(4)]
frames = [np.roll(base[i % 4], 9 * (i // 4), axis=1) for i in range(48)]
for variant in ("r101", "r101-fast"):
    for lanes in (1, 2, 3, 4):
        p = SoloV2Lanes(sd, variant=variant, lanes=lanes)
        list(p.map(iter(frames[:8])))
        t0 = time.perf_counter(); n = sum(1 for _ in p.map(iter(frames))); dt = time.perf_counter() - t0
        print(f"{variant} lanes {lanes}: {n / dt:.1f} frames/s ({1e3 * dt / n:.2f} ms/frame)", flush=True)
        p.close()
