#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the small GEMM shapes of every epilogue variant
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 30 python -m pytest tests/test_gemm_gpu.py -m gpu -q \
  -k "256-128-64-0-640 or 777-128-200-1-640 or 4100-68-128-0-640 or 700-328-200 or 300-264-200 or 128-128-64-0-128 or 300-384-200 or 300-256-200-2-512 or 300-520-64" > gpurun_out/r2c41_racecheck_gemm.log 2>&1
echo "racecheck gemm rc=$?" >> gpurun_out/r2c41_racecheck_gemm.log
tail -25 gpurun_out/r2c41_racecheck_gemm.log
