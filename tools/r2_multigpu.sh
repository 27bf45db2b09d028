#!/bin/bash
# 2-GPU checks: bench.py under torchrun (both arms), and the bands' --gpus 2 path over NCCL on two real GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2mg_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2mg_bench_n2.json 2> gpurun_out/r2mg_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r2mg_bench_n2.err
python - > gpurun_out/r2mg_bands.log 2>&1 <<'PY'
import json, os, subprocess, sys, time
import cv2, numpy as np
sys.path.insert(0, '.')
from prisma_b200.synthetic import synthetic_frame
root = "/tmp/mgclip"
for tag in ("one", "two"):
    os.makedirs(f"{root}/{tag}", exist_ok=True)
    w = cv2.VideoWriter(f"{root}/{tag}/rgba.mp4", cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (1280, 720))
    base = [synthetic_frame(720, 1280, t) for t in range(4)]
    for t in range(32):
        w.write(np.roll(base[t % 4], 5 * (t // 4), axis=1)[..., ::-1].copy())
    w.release()
    json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 1280, "height": 720, "frames": 32, "fps": 24.0}, open(f"{root}/{tag}/metadata.json", "w"))
res = {}
for band, extra in (("flow_raft", ["-b", "--iterations", "12"]), ("depth_anything", ["--encoder", "vitl"]), ("mask_mmdet", ["--sdf"])):
    for tag, g in (("one", []), ("two", ["--gpus", "2"])):
        t0 = time.time()
        rc = subprocess.call([sys.executable, f"bands/{band}.py", "-i", f"{root}/{tag}", "--seeded-weights"] + extra + g)
        res[(band, tag)] = (rc, time.time() - t0)
        print(band, tag, "rc", rc, "%.1f s" % (time.time() - t0), flush=True)
def frames(p):
    cap = cv2.VideoCapture(p); out = []
    while True:
        ok, f = cap.read()
        if not ok: break
        out.append(f)
    return out
for name in ("flow_raft.mp4", "flow_raft_bwd.mp4", "depth_anything.mp4", "mask.mp4"):
    a, b = frames(f"{root}/one/{name}"), frames(f"{root}/two/{name}")
    print(name, len(a), len(b), "equal" if len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)) else "DIFFER")
for name in ("flow_raft.csv", "depth_anything_min.csv", "depth_anything_max.csv"):
    print(name, "equal" if open(f"{root}/one/{name}").read() == open(f"{root}/two/{name}").read() else "DIFFER")
PY
head -c 400 gpurun_out/r2mg_bench_n2.json; echo; tail -3 gpurun_out/r2mg_bench_n2.err; cat gpurun_out/r2mg_bands.log | tail -14
