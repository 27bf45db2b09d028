#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q > gpurun_out/r2c22_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c22_tests.log
for st in 0 700 1100 1500; do echo "== stagger $st"; PRISMA_ATTN_STAGGER=$st timeout 300 python tools/attn_prof.py 2443 64 | grep -v "max abs"; PRISMA_ATTN_STAGGER=$st timeout 300 python tools/attn_prof.py 2443 192 | tail -2; done > gpurun_out/r2c22_attn.txt 2>&1
PRISMA_B200_LIB=$PWD/prisma_b200/libprisma_b200_prof.so PRISMA_ATTN_PROF=1 timeout 300 python tools/attn_prof.py 2443 64 > gpurun_out/r2c22_attn_prof.txt 2>&1
tail -3 gpurun_out/r2c22_tests.log; cat gpurun_out/r2c22_attn.txt gpurun_out/r2c22_attn_prof.txt
