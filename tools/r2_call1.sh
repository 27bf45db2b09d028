#!/bin/bash
# round-2 GPU call 1: full GPU suite, RAFT video-pass step profile, bench line (headline config)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c1_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2c1_suite.log
PRISMA_RAFT_PROFILE=1 timeout 300 python tools/raft_profile.py > gpurun_out/r2c1_raft_profile.txt 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?" >> gpurun_out/r2c1_bench.err
tail -3 gpurun_out/r2c1_suite.log
head -c 600 gpurun_out/r2c1_bench.json
