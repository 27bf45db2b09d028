#!/bin/bash
# full ncu capture of the ViT-L fc2 linear through the TMA reduce-add epilogue
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 \
    -o gpurun_out/r02c_full_gemm_fc2_tmareduce python tools/prof_kernels.py gemm_red > gpurun_out/r02c_full_gemm_red.log 2>&1
tail -3 gpurun_out/r02c_full_gemm_red.log; ls -la gpurun_out | grep fc2
