#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_raft_gpu.py tests/test_flow_gpu.py tests/test_depth_gpu.py "tests/test_band_surface.py::test_flow_band_sharded_with_halo_equals_single" -m gpu -q -x > gpurun_out/r2c7_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c7_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2c7_bench_pdl.json 2> gpurun_out/r2c7_bench_pdl.err
PRISMA_PDL=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2c7_bench_nopdl.json 2> gpurun_out/r2c7_bench_nopdl.err
grep -E "passed|failed|rc=" gpurun_out/r2c7_tests.log | tail -3
python - <<'PY'
import json
for n in ("pdl","nopdl"):
    try:
        d=json.load(open(f"gpurun_out/r2c7_bench_{n}.json")); print(n, d["value"], d["e2e"]["value"])
    except Exception as e: print(n, "failed", e)
PY
