#!/bin/bash
# where do the ~12 us of fixed cost per GEMM launch go?  PRISMA_GEMM_DBG: 0 full, 1 prologue+teardown, 2 no epilogue work, 3 loads only
for m in 0 1 2 3; do echo "== PRISMA_GEMM_DBG=$m"; PRISMA_GEMM_DBG=$m python tools/gemm_overhead.py; done
