#!/bin/bash
# the scale-only TMA-store path restored for the correlation volume: quick checks + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_flow_gpu.py -m gpu -q -x > gpurun_out/r2c39_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c39_tests.log
tail -3 gpurun_out/r2c39_tests.log
timeout 900 python bench.py > gpurun_out/r2c39_bench.json 2> gpurun_out/r2c39_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c39_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"], "launches", d["gpu_launches"])
g = d["roofline"]["groups"]
print({k: (v.get("tflops") if isinstance(v, dict) else v) for k, v in g.items() if k.startswith("da_") or k.startswith("raft_c")})
print(g["da_ms_per_pass"]); print(g["raft_ms_per_pair"]); print(g["raft_corr_build_in_pass"]); print(d["extra"]["raft_corr_build"]["frac"], d["extra"]["raft_corr_build"]["ms_per_build"])
PY
