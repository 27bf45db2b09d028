#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_mask_gpu.py -m gpu -x -q 2>&1 | tail -12
PRISMA_SOLO_PROFILE=1 timeout 200 python - <<'PY' 2>&1 | grep -E "solo-profile|ms" | tail -8
import os, sys
sys.path.insert(0, os.getcwd())
from prisma_b200.mask import SoloV2Engine
from prisma_b200.seeded_weights import make_solo_weights
from prisma_b200.synthetic import synthetic_frame
eng = SoloV2Engine(make_solo_weights("r101", 0))
f = synthetic_frame(1080, 1920, 0)
eng.infer(f)
os.environ.pop("PRISMA_SOLO_PROFILE")
eng.infer(f)
print("graph ms", eng.infer(f)["ms"], eng.work(1080, 1920))
PY
