#!/bin/bash
# ViT proj / fc2 through the TMA reduce-add epilogue (in-place fp32 residual): correctness + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_depth_gpu.py tests/test_midas_gpu.py tests/test_zoe_gpu.py -m gpu -q -x > gpurun_out/r2c37_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c37_tests.log
tail -4 gpurun_out/r2c37_tests.log
timeout 400 python bench.py > gpurun_out/r2c37_bench.json 2> gpurun_out/r2c37_bench.err
PRISMA_DA_TMA_REDUCE=0 timeout 400 python bench.py > gpurun_out/r2c37_bench_off.json 2> gpurun_out/r2c37_bench_off.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c37_bench.json", "gpurun_out/r2c37_bench_off.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], "roofline", d["roofline"]["frac"])
        g = d["roofline"]["groups"]
        print(g["da_encoder_linears"]); print(g["da_ms_per_pass"]); print(g["raft_ms_per_pair"])
        print({k: (round(v.get("frames_per_s_device", 0), 1) if isinstance(v, dict) else v) for k, v in d["extra"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "ERR", e)
PY
