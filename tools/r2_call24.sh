#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c24_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2c24_suite.log
timeout 900 python bench.py > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err
tail -5 gpurun_out/r2c24_suite.log; tail -2 gpurun_out/r2c24_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c24_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
g = d["roofline"]["groups"]
print({k: (v.get("tflops") if isinstance(v, dict) else v) for k, v in g.items() if k.startswith("da_") or k.startswith("raft_c")})
print(g["da_ms_per_pass"]); print(g["raft_ms_per_pair"])
PY
