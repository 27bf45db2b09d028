#!/bin/bash
# 16 epilogue warps only for the transposed / narrow tiles (the default) vs 8 everywhere, same box
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2c36_bench_sel.json 2> gpurun_out/r2c36_bench_sel.err
PRISMA_GEMM_EW=8 timeout 400 python bench.py > gpurun_out/r2c36_bench_ew8.json 2> gpurun_out/r2c36_bench_ew8.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c36_bench_sel.json", "gpurun_out/r2c36_bench_ew8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
        g = d["roofline"]["groups"]
        print(g["da_encoder_linears"]); print(g["da_ms_per_pass"]); print(g["raft_ms_per_pair"])
    except Exception as e:
        print(f, "ERR", e)
PY
