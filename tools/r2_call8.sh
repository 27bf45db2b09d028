#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_mask_gpu.py tests/test_raft_gpu.py -m gpu -q -s > gpurun_out/r2c8_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c8_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
grep -E "passed|failed|rc=|exact head|instances [0-9]|mismatch|tf32x3" gpurun_out/r2c8_tests.log | tail -12
