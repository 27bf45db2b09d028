#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_flow_gpu.py tests/test_raft_gpu.py -m gpu -q -s > gpurun_out/r2c27_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c27_tests.log
timeout 600 python bench.py > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err
grep -E "passed|failed|rc=|raft:|corr build" gpurun_out/r2c27_tests.log | tail -12; tail -2 gpurun_out/r2c27_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c27_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
g = d["roofline"]["groups"]
print(g["raft_ms_per_pair"]); print(g["raft_corr_build_in_pass"]); print(d["extra"]["raft_corr_build"]["frac"], d["extra"]["raft_corr_build"]["ms_per_build"])
PY
