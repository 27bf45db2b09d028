#!/bin/bash
# 2-GPU run of the final build through the driver's launch line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2c43_bench_n2.json 2> gpurun_out/r2c43_bench_n2.err
echo "rc=$?"; tail -c 1500 gpurun_out/r2c43_bench_n2.json | head -c 1500; tail -3 gpurun_out/r2c43_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c43_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"], "clocks", d["clocks"])
PY
