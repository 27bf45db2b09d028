#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mask_gpu.py tests/test_band_surface.py -m gpu -q > gpurun_out/r2c16_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c16_tests.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c16_sanitizer_smoke.log 2>&1
echo "sanitizer smoke rc=$?" >> gpurun_out/r2c16_sanitizer_smoke.log
tail -4 gpurun_out/r2c16_tests.log; tail -8 gpurun_out/r2c16_sanitizer_smoke.log
