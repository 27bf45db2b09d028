#!/bin/bash
# full GPU suite + smoke + bench of the final state of round 2 (after the import clean-up)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c42_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2c42_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c42_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c42_smoke.log
timeout 900 python bench.py > gpurun_out/r2c42_bench.json 2> gpurun_out/r2c42_bench.err
tail -6 gpurun_out/r2c42_suite.log; tail -3 gpurun_out/r2c42_smoke.log; tail -2 gpurun_out/r2c42_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c42_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"], "launches", d["gpu_launches"])
g = d["roofline"]["groups"]
print({k: (v.get("tflops") if isinstance(v, dict) else v) for k, v in g.items() if k.startswith("da_") or k.startswith("raft_c")})
print(g["da_ms_per_pass"]); print(g["raft_ms_per_pair"]); print(g["raft_corr_build_in_pass"])
print(json.dumps(d["extra"]["mask_mmdet_1080p"])[:900])
PY
