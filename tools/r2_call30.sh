#!/bin/bash
# coarse correlation levels against fp16 pooled features (one K-slab); SOLOv2 fp32-class backbone ("-exact") end to end
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flow_gpu.py tests/test_raft_gpu.py tests/test_mask_gpu.py -m gpu -q -s -x > gpurun_out/r2c30_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c30_tests.log
timeout 400 python bench.py > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err
grep -E "passed|failed|rc=|exact|instances:|raft:|flow err|Error" gpurun_out/r2c30_tests.log | tail -30; tail -2 gpurun_out/r2c30_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c30_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
        g = d["roofline"]["groups"]
        print(g["raft_ms_per_pair"]); print(g["raft_corr_build_in_pass"]); print(d["extra"]["raft_corr_build"]["frac"], d["extra"]["raft_corr_build"]["ms_per_build"])
    except Exception as e:
        print(f, "ERR", e)
PY
