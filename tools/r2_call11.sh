#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_raft_gpu.py tests/test_flow_gpu.py tests/test_band_surface.py -m gpu -q -x > gpurun_out/r2c11_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c11_tests.log
PRISMA_RAFT_PAIRS=1 timeout 600 python bench.py --no-extras > gpurun_out/r2c11_bench_np1.json 2> gpurun_out/r2c11_bench_np1.err
timeout 600 python bench.py --no-extras > gpurun_out/r2c11_bench_np2.json 2> gpurun_out/r2c11_bench_np2.err
tail -5 gpurun_out/r2c11_tests.log
for f in gpurun_out/r2c11_bench_np1 gpurun_out/r2c11_bench_np2; do tail -2 $f.err; python - $f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"].get("groups", {}).get("raft_ms_per_pair"))
except Exception as e:
    print("bad json", e)
PY
done
