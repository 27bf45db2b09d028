#!/bin/bash
# Round profile pass (run under gpurun, 1 GPU): launch list of the bench command + full captures of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PRISMA_BENCH_BATCH=${PRISMA_BENCH_BATCH:-12}
# (1) every launch of two passes with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 440 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
# (2) full capture of the dominant kernels at ViT-L shapes / RAFT 1080p correlation
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 \
    -o gpurun_out/full_gemm python tools/prof_kernels.py gemm > gpurun_out/full_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 \
    -o gpurun_out/full_attn python tools/prof_kernels.py attn > gpurun_out/full_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 1 \
    -o gpurun_out/full_corr python tools/prof_kernels.py corr > gpurun_out/full_corr.log 2>&1
ls -la gpurun_out
