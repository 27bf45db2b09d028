#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_band_surface.py tests/test_raft_gpu.py tests/test_midas_gpu.py -m gpu -q -x > gpurun_out/r2c4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c4_tests.log
timeout 300 python tools/gemm_raft_shapes.py > gpurun_out/r2c4_gemm_shapes.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c4_raft_launches.csv python tools/raft_profile_plain.py > gpurun_out/r2c4_ncu.log 2>&1
python - <<'PY' > gpurun_out/r2c4_membw.txt 2>&1
import torch, time
x = torch.empty(1<<30, dtype=torch.float32, device="cuda")   # 4 GiB
for name, fn in (("fill (write only)", lambda: x.fill_(1.0)), ("copy (read+write)", lambda: x[: 1<<29].copy_(x[1<<29 :]))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nbytes = x.numel() * 4
    print(name, "%.3f ms  %.0f GB/s" % (ms, nbytes / ms / 1e6))
PY
tail -3 gpurun_out/r2c4_tests.log; cat gpurun_out/r2c4_membw.txt
