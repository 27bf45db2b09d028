#!/bin/bash
mkdir -p gpurun_out
for g in 4 8 16; do
  echo "== PRISMA_TF32_ACC_GROUP=$g" >> gpurun_out/r2c10_accgroup.txt
  PRISMA_TF32_ACC_GROUP=$g timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -s -k "tf32x3" 2>&1 | grep -E "tf32x3|passed|failed" >> gpurun_out/r2c10_accgroup.txt
  PRISMA_TF32_ACC_GROUP=$g PRISMA_SOLO_PROFILE=1 timeout 300 python - >> gpurun_out/r2c10_accgroup.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from prisma_b200.mask import SoloV2Engine
from prisma_b200.seeded_weights import make_solo_weights
from prisma_b200.synthetic import synthetic_frame
eng = SoloV2Engine(make_solo_weights("r101", 0))
f = synthetic_frame(1080, 1920, 0)
eng.infer(f); r = eng.infer(f); print("ms", r["ms"])
PY
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c10_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2c10_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c10_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c10_smoke.log
cat gpurun_out/r2c10_accgroup.txt | grep -E "==|tf32x3|towers|mask_feature|^ms" ; tail -4 gpurun_out/r2c10_suite.log; tail -3 gpurun_out/r2c10_smoke.log
