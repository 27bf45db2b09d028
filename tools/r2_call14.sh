#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_raft_gpu.py tests/test_flow_gpu.py -m gpu -q -x > gpurun_out/r2c14_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c14_tests.log
PRISMA_RAFT_FUSE_GRU=0 timeout 600 python bench.py --no-extras > gpurun_out/r2c14_bench_nofuse.json 2> gpurun_out/r2c14_bench_nofuse.err
timeout 600 python bench.py --no-extras > gpurun_out/r2c14_bench_fuse.json 2> gpurun_out/r2c14_bench_fuse.err
PRISMA_GEMM_PAIR128=1 timeout 600 python bench.py --no-extras > gpurun_out/r2c14_bench_pair128.json 2> gpurun_out/r2c14_bench_pair128.err
tail -5 gpurun_out/r2c14_tests.log
for n in nofuse fuse pair128; do tail -2 gpurun_out/r2c14_bench_$n.err; python - gpurun_out/r2c14_bench_$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], d["roofline"].get("groups", {}).get("raft_ms_per_pair"))
except Exception as e:
    print("bad json", e)
PY
done
