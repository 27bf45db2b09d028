#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02b_launches_step.csv python tools/prof_step.py > gpurun_out/r02b_prof_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 \
    -o gpurun_out/r02b_full_attn python tools/prof_kernels.py attn > gpurun_out/r02b_full_attn.log 2>&1
ls -la gpurun_out | grep r02b
