#!/bin/bash
# the scale-only TMA-store path restored for the correlation volume (quick checks) + epilogue-warp policies on the bench:
# default (16 narrow / 8 wide), 12 everywhere, 12 narrow / 8 wide, 16 narrow / 12 wide
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_flow_gpu.py -m gpu -q -x > gpurun_out/r2c39_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c39_tests.log
tail -3 gpurun_out/r2c39_tests.log
timeout 400 python bench.py --no-extras > gpurun_out/r2c39_bench_default.json 2> gpurun_out/r2c39_bench_default.err
PRISMA_GEMM_EW=12 timeout 400 python bench.py --no-extras > gpurun_out/r2c39_bench_all12.json 2> gpurun_out/r2c39_bench_all12.err
PRISMA_GEMM_EW_NARROW=12 timeout 400 python bench.py --no-extras > gpurun_out/r2c39_bench_n12w8.json 2> gpurun_out/r2c39_bench_n12w8.err
PRISMA_GEMM_EW_WIDE=12 timeout 400 python bench.py --no-extras > gpurun_out/r2c39_bench_n16w12.json 2> gpurun_out/r2c39_bench_n16w12.err
python - <<'PY'
import json
for n in ("default", "all12", "n12w8", "n16w12"):
    f = f"gpurun_out/r2c39_bench_{n}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        g = d["roofline"]["groups"]
        print(n, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "clk", d["clocks"]["sm_mhz"], "| da linear", round(g["da_ms_per_pass"]["linear"], 2), "head", round(g["da_ms_per_pass"]["head"], 2),
              "da total", round(g["da_ms_per_pass"]["total"], 2), "| raft conv", round(g["raft_ms_per_pair"]["conv_gemm"], 3), "corr", round(g["raft_ms_per_pair"]["corr_build"], 3), "raft total", round(g["raft_ms_per_pair"]["total"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
