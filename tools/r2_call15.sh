#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c15_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2c15_suite.log
# compute-sanitizer memcheck over the smoke pass of every engine (tiny frames) and the small GEMM / attention unit shapes
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c15_sanitizer_smoke.log 2>&1
echo "sanitizer smoke rc=$?" >> gpurun_out/r2c15_sanitizer_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "128-128-64 or 300-384-200 or 500-48-384 or 64-64-64 or 512-256-128 or 300-256-200 or 700-256-320 or 1000-264-256 or 300-520-64 or 64-32-100 or 300-256-512" > gpurun_out/r2c15_sanitizer_gemm.log 2>&1
echo "sanitizer gemm rc=$?" >> gpurun_out/r2c15_sanitizer_gemm.log
tail -6 gpurun_out/r2c15_suite.log; tail -6 gpurun_out/r2c15_sanitizer_smoke.log; tail -6 gpurun_out/r2c15_sanitizer_gemm.log
