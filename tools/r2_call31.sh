#!/bin/bash
# transposed tiles (GemmCfg SWAP) for N <= 128; r101-exact end to end
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r2c31_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/r2c31_gemm.log
tail -3 gpurun_out/r2c31_gemm.log
PRISMA_GEMM_DBG=0 timeout 120 python tools/gemm_pace.py > gpurun_out/r2c31_pace.txt 2>&1; cat gpurun_out/r2c31_pace.txt
timeout 600 python -m pytest tests/test_raft_gpu.py tests/test_flow_gpu.py -m gpu -q -s -x > gpurun_out/r2c31_raft.log 2>&1; echo "raft rc=$?" >> gpurun_out/r2c31_raft.log
grep -E "passed|failed|rc=|raft:" gpurun_out/r2c31_raft.log | tail -6
timeout 600 python -m pytest tests/test_mask_gpu.py -m gpu -q -s -k "r101_exact" > gpurun_out/r2c31_mask.log 2>&1; echo "mask rc=$?" >> gpurun_out/r2c31_mask.log
grep -E "passed|failed|rc=|exact|Error" gpurun_out/r2c31_mask.log | tail -8
PRISMA_TF32_ACC_GROUP=1 timeout 600 python -m pytest tests/test_mask_gpu.py -m gpu -q -s -k "r101_exact" > gpurun_out/r2c31_mask_g1.log 2>&1; echo "mask g1 rc=$?" >> gpurun_out/r2c31_mask_g1.log
grep -E "passed|failed|rc=|exact|Error" gpurun_out/r2c31_mask_g1.log | tail -8
timeout 400 python bench.py > gpurun_out/r2c31_bench.json 2> gpurun_out/r2c31_bench.err
PRISMA_GEMM_SWAP64=1 timeout 400 python bench.py > gpurun_out/r2c31_bench_s64.json 2> gpurun_out/r2c31_bench_s64.err
PRISMA_GEMM_SWAP=0 timeout 400 python bench.py > gpurun_out/r2c31_bench_noswap.json 2> gpurun_out/r2c31_bench_noswap.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c31_bench.json", "gpurun_out/r2c31_bench_s64.json", "gpurun_out/r2c31_bench_noswap.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
        g = d["roofline"]["groups"]
        print(g["raft_ms_per_pair"]); print(g["da_ms_per_pass"])
    except Exception as e:
        print(f, "ERR", e)
PY
