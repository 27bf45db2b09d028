#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_mask_gpu.py "tests/test_band_surface.py::test_flow_band_sharded_with_halo_equals_single" tests/test_raft_gpu.py tests/test_depth_gpu.py -m gpu -q -s > gpurun_out/r2c6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c6_tests.log
PRISMA_GEMM_DBG=0 timeout 300 python tools/gemm_overhead.py > gpurun_out/r2c6_gemm_overhead.txt 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
grep -E "passed|failed|rc=" gpurun_out/r2c6_tests.log | tail -3
