#!/bin/bash
mkdir -p gpurun_out
bash tools/profile_round2.sh > gpurun_out/r2c12_profile.log 2>&1
for n in 3 4; do
  PRISMA_RAFT_PAIRS=$n timeout 600 python bench.py --no-extras > gpurun_out/r2c12_bench_np$n.json 2> gpurun_out/r2c12_bench_np$n.err
done
tail -12 gpurun_out/r2c12_profile.log
for n in 3 4; do tail -2 gpurun_out/r2c12_bench_np$n.err; python - gpurun_out/r2c12_bench_np$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], d["roofline"].get("groups", {}).get("raft_ms_per_pair"))
except Exception as e:
    print("bad json", e)
PY
done
