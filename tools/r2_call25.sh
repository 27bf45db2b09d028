#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mask_gpu.py -m gpu -q -k "lanes" > gpurun_out/r2c25_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c25_tests.log
timeout 900 python tools/mask_lanes.py > gpurun_out/r2c25_lanes.txt 2>&1
timeout 600 python -m pytest tests/test_band_surface.py -m gpu -q >> gpurun_out/r2c25_tests.log 2>&1
echo "band tests rc=$?" >> gpurun_out/r2c25_tests.log
tail -4 gpurun_out/r2c25_tests.log; cat gpurun_out/r2c25_lanes.txt
