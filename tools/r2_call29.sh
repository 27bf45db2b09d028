#!/bin/bash
# what paces the GEMM per 64-wide K block: loads, MMA issue, or the epilogue?
mkdir -p gpurun_out
for m in 0 2 3 5 6; do echo "== PRISMA_GEMM_DBG=$m"; PRISMA_GEMM_DBG=$m timeout 300 python tools/gemm_pace.py; done > gpurun_out/r2c29_pace.txt 2>&1
cat gpurun_out/r2c29_pace.txt
