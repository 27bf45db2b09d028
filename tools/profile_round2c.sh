#!/bin/bash
# Final-state profile pass of round 2 (under gpurun, 1 GPU): launch list of one bench step + full captures of the kernels
# changed in the second session (merged coarse correlation GEMM, TMA-store fc1 epilogue, transposed tiles).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02c_launches_step.csv python tools/prof_step.py > gpurun_out/r02c_prof_step.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 \
    -o gpurun_out/r02c_full_corr_l0_and_coarse python tools/prof_kernels.py corr > gpurun_out/r02c_full_corr.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 \
    -o gpurun_out/r02c_full_gemm_fc1_tmastore python tools/prof_kernels.py gemm_tma > gpurun_out/r02c_full_gemm_tma.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 \
    -o gpurun_out/r02c_full_gemm_swap_n128 python tools/prof_kernels.py gemm_swap > gpurun_out/r02c_full_gemm_swap.log 2>&1
ls -la gpurun_out | grep r02c
