#!/bin/bash
# Round-2 profile pass (under gpurun, 1 GPU): launch list of one bench step + full captures of the kernels the roofline names.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# (1) every launch of one step (2 depth passes of 12 frames + 24 flow pairs) with its device time
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_step.csv python tools/prof_step.py > gpurun_out/r02_prof_step.log 2>&1
# (2) full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 \
    -o gpurun_out/r02_full_gemm_fc1 python tools/prof_kernels.py gemm > gpurun_out/r02_full_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 \
    -o gpurun_out/r02_full_attn python tools/prof_kernels.py attn > gpurun_out/r02_full_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 1 \
    -o gpurun_out/r02_full_corr_l0 python tools/prof_kernels.py corr > gpurun_out/r02_full_corr.log 2>&1
# three consecutive update-block convs of a RAFT pass (convc1 / convc2 / ... / gru), un-graphed launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 70 -c 4 \
    -o gpurun_out/r02_full_raft_update python tools/raft_profile_plain.py > gpurun_out/r02_full_raft.log 2>&1
ls -la gpurun_out | grep r02
