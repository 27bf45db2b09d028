#!/bin/bash
# round-2 GPU call 2: RAFT / flow / GEMM tests after the layout + epilogue changes, step profile, bench (headline only)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_flow_gpu.py tests/test_raft_gpu.py tests/test_band_surface.py -m gpu -q -s -x > gpurun_out/r2c2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c2_tests.log
timeout 300 python tools/raft_profile.py > gpurun_out/r2c2_raft_profile.txt 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?" >> gpurun_out/r2c2_bench.err
tail -4 gpurun_out/r2c2_tests.log
head -c 300 gpurun_out/r2c2_bench.json
