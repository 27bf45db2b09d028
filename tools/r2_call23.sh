#!/bin/bash
mkdir -p gpurun_out
for v in 0 4 3 2; do echo "== poly every $v"; PRISMA_B200_LIB=$PWD/prisma_b200/libprisma_b200_poly$v.so timeout 300 python tools/attn_prof.py 2443 64 | tail -2; PRISMA_B200_LIB=$PWD/prisma_b200/libprisma_b200_poly$v.so timeout 300 python tools/attn_prof.py 2443 192 | tail -2; done > gpurun_out/r2c23_attn.txt 2>&1
cat gpurun_out/r2c23_attn.txt
