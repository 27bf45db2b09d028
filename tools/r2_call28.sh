#!/bin/bash
# merged coarse correlation levels (2 GEMMs per direction) + the row-halo upper-bound experiment (PRISMA_GEMM_DBG=4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flow_gpu.py tests/test_raft_gpu.py -m gpu -q -s > gpurun_out/r2c28_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c28_tests.log
timeout 400 python bench.py > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28_bench.err
PRISMA_GEMM_DBG=4 timeout 400 python bench.py > gpurun_out/r2c28_bench_dbg4.json 2> gpurun_out/r2c28_bench_dbg4.err
grep -E "passed|failed|rc=" gpurun_out/r2c28_tests.log | tail -4; tail -2 gpurun_out/r2c28_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c28_bench.json", "gpurun_out/r2c28_bench_dbg4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
        g = d["roofline"]["groups"]
        print(g["raft_ms_per_pair"]); print(g["da_ms_per_pass"]); print(g["raft_corr_build_in_pass"]); print(d["extra"]["raft_corr_build"]["frac"], d["extra"]["raft_corr_build"]["ms_per_build"])
    except Exception as e:
        print(f, "ERR", e)
PY
