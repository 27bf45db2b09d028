#!/bin/bash
# A/B of the GEMM tail tiles (PRISMA_GEMM_TAIL=0 disables): kernel tests, engine tests, then the bench both ways.
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 400 python -m pytest tests/test_raft_gpu.py tests/test_depth_gpu.py tests/test_mask_gpu.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|raft pass|Error" | tail -8
for t in 1 0; do
  PRISMA_GEMM_TAIL=$t timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('tail',$t,round(d['value'],1),round(d['e2e']['value'],1),round(d['roofline']['achieved']),{k:round(v,2) for k,v in d['roofline']['groups_ms_per_pass'].items()}, d['extra']['flow_raft_1080p']['ms_per_pass_device'], d['extra']['mask_mmdet_1080p']['ms_per_pass_device'])"
done
PRISMA_SOLO_PROFILE=1 timeout 200 python - <<'PY' 2>&1 | grep -E "solo-profile|ms" | tail -8
import os, sys
sys.path.insert(0, os.getcwd())
from prisma_b200.mask import SoloV2Engine
from prisma_b200.seeded_weights import make_solo_weights
from oracle.frames import synthetic_frame
eng = SoloV2Engine(make_solo_weights("r101", 0))
f = synthetic_frame(1080, 1920, 0)
eng.infer(f)
os.environ.pop("PRISMA_SOLO_PROFILE")
print("graph ms", eng.infer(f)["ms"])
PY
