#!/bin/bash
# where does the RAFT conv time go inside the engine?  bench.py's per-group profile under the GEMM diagnostics modes
mkdir -p gpurun_out
for m in 2 3 6 5; do
  PRISMA_GEMM_DBG=$m timeout 300 python bench.py > gpurun_out/r2c32_bench_dbg$m.json 2> gpurun_out/r2c32_bench_dbg$m.err
done
python - <<'PY'
import json
for m in (2, 3, 6, 5):
    f = f"gpurun_out/r2c32_bench_dbg{m}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        g = d["roofline"]["groups"]
        print("dbg", m, "value", round(d["value"], 1), "clocks", d["clocks"]["sm_mhz"], "raft conv_gemm", round(g["raft_ms_per_pair"]["conv_gemm"], 3), "raft total", round(g["raft_ms_per_pair"]["total"], 3),
              "da linear", round(g["da_ms_per_pass"]["linear"], 2), "da head", round(g["da_ms_per_pass"]["head"], 2), "da total", round(g["da_ms_per_pass"]["total"], 2))
    except Exception as e:
        print(f, "ERR", e)
PY
