#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mask_gpu.py tests/test_gemm_gpu.py -m gpu -q -s > gpurun_out/r2c9_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c9_tests.log
PRISMA_SOLO_PROFILE=1 timeout 300 python - > gpurun_out/r2c9_solo_profile.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from prisma_b200.mask import SoloV2Engine
from prisma_b200.seeded_weights import make_solo_weights
from prisma_b200.synthetic import synthetic_frame
eng = SoloV2Engine(make_solo_weights("r101", 0))
f = synthetic_frame(1080, 1920, 0)
eng.infer(f); r = eng.infer(f); print("ms", r["ms"])
PY
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
grep -E "passed|failed|rc=|exact head|instances" gpurun_out/r2c9_tests.log | tail -12; tail -12 gpurun_out/r2c9_solo_profile.txt
