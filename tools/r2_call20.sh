#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_depth_gpu.py -m gpu -q > gpurun_out/r2c20_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c20_tests.log
for a in "2443 64" "2560 64" "1370 192"; do timeout 300 python tools/attn_prof.py $a; done > gpurun_out/r2c20_attn.txt 2>&1
PRISMA_B200_LIB=$PWD/prisma_b200/libprisma_b200_prof.so PRISMA_ATTN_PROF=1 timeout 300 python tools/attn_prof.py 2443 64 > gpurun_out/r2c20_attn_prof.txt 2>&1
tail -6 gpurun_out/r2c20_tests.log; cat gpurun_out/r2c20_attn.txt gpurun_out/r2c20_attn_prof.txt
