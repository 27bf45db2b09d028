#!/bin/bash
# 16 epilogue warps (GemmCfg EW) for the plain fp16 GEMM kernels: correctness + A/B on the bench
mkdir -p gpurun_out
PRISMA_GEMM_EW=16 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_raft_gpu.py tests/test_depth_gpu.py -m gpu -q -x > gpurun_out/r2c35_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c35_tests.log
tail -4 gpurun_out/r2c35_tests.log
PRISMA_GEMM_EW=16 timeout 400 python bench.py > gpurun_out/r2c35_bench_ew16.json 2> gpurun_out/r2c35_bench_ew16.err
PRISMA_GEMM_EW=8 timeout 400 python bench.py > gpurun_out/r2c35_bench_ew8.json 2> gpurun_out/r2c35_bench_ew8.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c35_bench_ew16.json", "gpurun_out/r2c35_bench_ew8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
        g = d["roofline"]["groups"]
        print(g["da_encoder_linears"]); print(g["da_ms_per_pass"]); print(g["raft_ms_per_pair"])
        print({k: (round(v.get("frames_per_s_device", 0), 1) if isinstance(v, dict) else v) for k, v in d["extra"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "ERR", e)
PY
