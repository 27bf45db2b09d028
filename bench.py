#!/usr/bin/env python
"""bench.py -- frames/sec of the Depth-Anything ViT-L band path on synthetic 720p video (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (pre-process -> ViT-L -> DPT head -> post-process/encode) over a batch of
FRAMES_PER_STEP synthetic 720p frames per GPU.  Frames shard across ranks with no data-path collective
 -> weak scaling; NCCL is used only for the barrier and the max-over-ranks of the timings.

  value      : frames/s with the frames already resident in HBM (CUDA events around K steps, max over ranks)
  e2e        : frames/s through the public Python API (prisma_b200.depth.DepthAnythingEngine.infer_clip) from
               host frames, H2D of the frame and D2H of the encoded u8 frame + (min,max) inside the timed region
  roofline   : the tcgen05 GEMM core (encoder linears, the dominant kernel): algorithmic FLOP / CUDA-event time
               of those launches, vs the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the oracle port of the reference's CPU fp32 path (oracle/da.py; the Python
               reference itself cannot travel to the GPU box) on the host cores, on a bounded sample of frames
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 720, 1280
ENCODER = os.environ.get("PRISMA_BENCH_ENCODER", "vitl")
# frames per engine pass (frames are independent).  12 frames make the attention grid (20 q-tiles x 16 heads x 12 = 3840
# CTAs on 2 x 148 slots = 12.97 rounds) and the GEMM tile counts (230 row tiles) land just under whole waves; at 4 frames
# both lose ~15-20 % to the last partial wave (measured)
BATCH = int(os.environ.get("PRISMA_BENCH_BATCH", "12"))
FRAMES_PER_STEP = int(os.environ.get("PRISMA_BENCH_FRAMES", str(4 * BATCH)))
WORKLOAD = "synthetic 256-frame 720p video, depth_anything ViT-L, frames sharded per GPU (BASELINE configs[1])"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def make_frames(n, h, w, start):
    """Frames [start, start + n) of the synthetic clip (this rank's shard, prisma_b200.shard.frame_range)."""
    from prisma_b200.synthetic import synthetic_frame  # seeded synthetic clip shared with the tests
    base = [synthetic_frame(h, w, (start + t) % 256) for t in range(min(n, 4))]
    # 4 distinct generated frames (the generator is pure numpy and slow), then cheap deterministic variants
    frames = []
    for i in range(n):
        f = base[i % len(base)]
        frames.append(np.ascontiguousarray(np.roll(f, 7 * (i // len(base)), axis=1)))
    return frames


def cpu_threads():
    """torch CPU scales to ~32 threads on the GPU box's host (measured with tools/cpu_threads.py: 8 thr 7.4 s/frame,
    16: 6.3, 32: 6.0, 64: 7.7, 128: 41 s/frame) -- use the optimum, and say so."""
    return min(32, os.cpu_count() or 1)


def cpu_baseline_frames(n_frames, threads):
    """The reference's CPU fp32 path as restated by the oracle, timed on the host cores (frames/s)."""
    import torch
    from oracle import da as oda
    from oracle.weights import make_da_weights
    torch.set_num_threads(threads)
    sd = make_da_weights(ENCODER, 0)
    frames = make_frames(n_frames + 1, H, W, 0)
    oda.da_encode(oda.da_infer(sd, frames[0], ENCODER))  # warm-up
    t0 = time.perf_counter()
    for f in frames[1:]:
        oda.da_encode(oda.da_infer(sd, f, ENCODER))
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = cpu_threads()
    per_step = 1
    import torch
    from oracle import da as oda
    from oracle.weights import make_da_weights
    torch.set_num_threads(cores)
    sd = make_da_weights(ENCODER, 0)
    frames = make_frames(2, H, W, 0)
    for _ in range(min(args.warmup, 1)):
        oda.da_encode(oda.da_infer(sd, frames[0], ENCODER))
    steps = min(args.steps, 3)  # bounded: ~10 s of CPU work per ViT-L frame
    t0 = time.perf_counter()
    for i in range(steps):
        oda.da_encode(oda.da_infer(sd, frames[i % 2], ENCODER))
    dt = time.perf_counter() - t0
    fps = steps * per_step / dt
    out = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": [H, W], "encoder": ENCODER, "frames_per_step": per_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} frames of the 720p clip through oracle/da.py (torch CPU fp32, {cores} threads = measured optimum of {os.cpu_count()} host cores)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def raft_extras(device, peaks):
    """Second workload (BASELINE configs[2]): flow_raft on synthetic 1080p pairs, 12 GRU iterations, forward+backward
    flow per pass, plus the correlation-volume build (HBM-bound) against the measured copy bandwidth."""
    import ctypes as C
    from prisma_b200._lib import check, fptr, lib
    from prisma_b200.flow import RaftFlowEngine
    from prisma_b200.seeded_weights import make_raft_weights
    from prisma_b200.synthetic import synthetic_frame
    eng = RaftFlowEngine(make_raft_weights(0), device=device, iterations=12, scale=0.75)
    f = [synthetic_frame(1080, 1920, t) for t in range(2)]
    for _ in range(3):
        eng.infer_pair(f[0], f[1])
    n = 8
    t0 = time.perf_counter()
    dev_ms = 0.0
    for i in range(n):  # a video loop: every pair's `prev` is the previous pair's `curr` (features reused, same results)
        r = eng.infer_pair(f[i % 2], f[(i + 1) % 2], want_rgb=True, reuse_prev=True)
        dev_ms += r["ms"]
    e2e_s = time.perf_counter() - t0
    w = eng.work(1080, 1920)
    eng.close()
    # correlation pyramid build alone (K13+K14), fwd+bwd, P = 18360
    l = lib()
    h = C.c_void_p()
    rng = np.random.default_rng(0)
    fm = rng.standard_normal((2, 256, 102, 180), dtype=np.float32)
    check(l.prisma_flowcorr_create(device, 2, 102, 180, C.byref(h)))
    check(l.prisma_flowcorr_set_fmaps(h, fptr(fm), fptr(np.ascontiguousarray(fm[::-1]))))
    ms = C.c_float()
    check(l.prisma_flowcorr_build(h, 10, C.byref(ms)))
    work = (C.c_double * 2)()
    check(l.prisma_flowcorr_work(h, work))
    l.prisma_engine_destroy(h)
    gbs = work[1] / (ms.value * 1e-3) / 1e9
    return {
        "flow_raft_1080p": {"workload": "synthetic 1080p pairs, flow_raft 12 GRU iterations, fwd+bwd per pass (BASELINE configs[2])",
                            "frame_steps_per_s_device": n / (dev_ms * 1e-3), "frame_steps_per_s_e2e": n / e2e_s,
                            "ms_per_pass_device": dev_ms / n, "algorithmic_gflop_per_pass": w["flop"] / 1e9,
                            "tflops": w["flop"] / (dev_ms / n * 1e-3) / 1e12, "launches_per_pass": w["launches"]},
        "raft_corr_build": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                            "ms": ms.value, "algorithmic_bytes": work[1],
                            "traffic": (measured_traffic()[1] or {}).get("corr_level0", {}).get("traffic_bytes_per_launch"),
                            "traffic_detail": (measured_traffic()[1] or {}).get("corr_level0"), "peak_source": peaks["src"],
                            "note": "4-level fp32 pyramid written once (levels 1-3 by linearity: GEMMs against pooled features)"},
    }


def measured_traffic():
    """DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        return t["traffic_bytes_per_launch"], t
    except Exception:
        return None, None


def midas_extras(device):
    """Third workload: depth_midas (MiDaS v3 DPT_Large, BASELINE north_star band) on the same synthetic 720p frames,
    12-frame passes, frames resident in HBM."""
    from prisma_b200.depth import MidasEngine
    from prisma_b200.seeded_weights import make_midas_weights
    eng = MidasEngine(make_midas_weights("dpt_large", 0), device=device)
    eng.time_resident(H, W, 2, BATCH)
    ms = eng.time_resident(H, W, 4, BATCH)
    w = eng.work(H, W, BATCH)
    eng.close()
    flop = w["linear_flop"] + w["attention_flop"] + w["head_flop"]
    return {"workload": "synthetic 720p frames, depth_midas DPT_Large (384x672 net input), frames resident",
            "frames_per_s_device": BATCH / (ms * 1e-3), "ms_per_pass": ms, "frames_per_pass": BATCH,
            "tflops": flop / (ms * 1e-3) / 1e12, "launches_per_pass": w["launches"]}


def zoe_extras(device):
    """depth_anything --metric outdoor (what the reference's process.py passes by default): ZoeDepth metric head on ViT-L,
    392x518 network input, 12-frame passes, frames resident."""
    from prisma_b200.depth import ZoeDepthEngine
    from prisma_b200.seeded_weights import make_zoe_weights
    eng = ZoeDepthEngine(make_zoe_weights("vitl", 0), device=device, encoder="vitl")
    eng.time_resident(H, W, 3, BATCH)
    ms = eng.time_resident(H, W, 8, BATCH)
    w = eng.work(H, W, BATCH)
    eng.close()
    return {"workload": "synthetic 720p frames, depth_anything --metric (ZoeDepth head, 392x518 net input), frames resident",
            "frames_per_s_device": BATCH / (ms * 1e-3), "ms_per_pass": ms, "frames_per_pass": BATCH, "launches_per_pass": w["launches"]}


def mask_extras(device):
    """Fourth workload: the mask band (SOLOv2 R-101, BASELINE north_star band) on synthetic 1080p frames: host frame in,
    union mask + instance list out (H2D / D2H inside the wall time; `ms` is the device time of the pass)."""
    from prisma_b200.mask import SoloV2Engine
    from prisma_b200.seeded_weights import make_solo_weights
    from prisma_b200.synthetic import synthetic_frame
    eng = SoloV2Engine(make_solo_weights("r101", 0), device=device)
    f = [synthetic_frame(1080, 1920, t) for t in range(2)]
    for i in range(3):
        eng.infer(f[i % 2])
    n, dev_ms = 8, 0.0
    t0 = time.perf_counter()
    for i in range(n):
        dev_ms += eng.infer(f[i % 2])["ms"]
    e2e_s = time.perf_counter() - t0
    w = eng.work(1080, 1920)
    eng.close()
    return {"workload": "synthetic 1080p frames, mask_mmdet SOLOv2 R-101 (768x1344 net input), one frame per pass",
            "frames_per_s_device": n / (dev_ms * 1e-3), "frames_per_s_e2e": n / e2e_s, "ms_per_pass_device": dev_ms / n,
            "algorithmic_gflop_per_pass": w["flop"] / 1e9, "tflops": w["flop"] / (dev_ms / n * 1e-3) / 1e12,
            "launches_per_pass": w["launches"]}


def run_b200(args, rank, local_rank, world):
    import torch
    from prisma_b200.depth import DepthAnythingEngine
    from prisma_b200.seeded_weights import make_da_weights   # the B200 arm never touches oracle/

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    from prisma_b200.shard import frame_range, max_over_ranks as _max_over_ranks

    def max_over_ranks(x):
        return _max_over_ranks(x, dist, f"cuda:{local_rank}")

    eng = DepthAnythingEngine(ENCODER, make_da_weights(ENCODER, 0), device=local_rank)
    shard_start, shard_stop, _ = frame_range(rank, world, FRAMES_PER_STEP * world)  # weak scaling: 48 frames per GPU
    frames = make_frames(shard_stop - shard_start, H, W, shard_start)
    work = eng.work(H, W, BATCH)
    assert FRAMES_PER_STEP % BATCH == 0
    passes = FRAMES_PER_STEP // BATCH
    # the step's frames in pinned host memory ("frame in pinned host memory" -> "encoded u8 frame + scalars
    # in pinned host memory"); every step uploads all of them and downloads every encoded frame + (min, max)
    from prisma_b200.depth import pinned_empty
    clip = pinned_empty((FRAMES_PER_STEP, H, W, 3), np.uint8)
    clip[...] = np.stack(frames)
    out_rgb = pinned_empty((FRAMES_PER_STEP, H, W, 3), np.uint8)

    # ---------------- e2e: public API (infer_clip = the band's video loop over one chunk), H2D + D2H inside the timed region
    for i in range(max(args.warmup, 3)):
        eng.infer_clip(clip, pass_frames=BATCH, out_rgb=out_rgb)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    for s in range(args.steps):
        rgb, mins, maxs, _ = eng.infer_clip(clip, pass_frames=BATCH, out_rgb=out_rgb)
    torch.cuda.synchronize(local_rank)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    # ---------------- value: frames resident in HBM, CUDA events on the engine stream (inside the C ABI)
    eng.time_resident(H, W, max(args.warmup, 3), BATCH)
    barrier()
    ms = eng.time_resident(H, W, args.steps * passes, BATCH)  # ms per pass of BATCH frames
    res_s = max_over_ranks(ms * 1e-3 * args.steps * passes)
    clocks = sampler.stop()
    barrier()

    prof = eng.profile(H, W, BATCH)  # per kernel-group CUDA-event times of one pass of BATCH frames (ms)
    if rank == 0:
        peaks = measured_peaks()
        total_frames = args.steps * FRAMES_PER_STEP * world
        value = total_frames / res_s
        lin_tf = work["linear_flop"] / (prof["linear"] * 1e-3) / 1e12 if prof["linear"] > 0 else 0.0
        att_tf = work["attention_flop"] / (prof["attention"] * 1e-3) / 1e12 if prof["attention"] > 0 else 0.0
        head_tf = work["head_flop"] / (prof["head"] * 1e-3) / 1e12 if prof["head"] > 0 else 0.0
        out = {
            "metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * res_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [H, W], "encoder": ENCODER, "frames_per_step": FRAMES_PER_STEP,
                       "frames_per_pass": BATCH, "parallelism": f"frame-sharded x{world}",
                       "l2": "per-frame working set (fp16 weights ~0.6 GB for ViT-L) exceeds the 126 MB L2; no flush needed"},
            "e2e": {"value": total_frames / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": FRAMES_PER_STEP * H * W * 3,
                    "d2h_bytes_per_step": FRAMES_PER_STEP * (H * W * 3 + 8)},
            "gpu_launches": work["launches"] * args.steps * passes,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (encoder linears: qkv/proj/fc1/fc2/patch-embed)",
                         "achieved": lin_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": lin_tf / peaks["tf_sustained"], "traffic": measured_traffic()[0],
                         "traffic_detail": measured_traffic()[1], "peak_source": peaks["src"],
                         "groups_ms_per_pass": prof,
                         "attention_tflops": att_tf, "head_tflops": head_tf,
                         "frame_flop": (work["linear_flop"] + work["attention_flop"] + work["head_flop"]) / BATCH},
        }
        if world == 1:
            # secondary workloads: a failure there is reported in place and never costs the headline line
            out["extra"] = {}

            def extra(name, fn):
                try:
                    r = fn()
                    out["extra"].update(r if name is None else {name: r})
                except Exception as ex:  # noqa: BLE001
                    out["extra"][name or "flow_raft_1080p"] = {"error": f"{type(ex).__name__}: {ex}"}

            extra(None, lambda: raft_extras(local_rank, peaks))

            def pipeline_1080p():
                # BASELINE metric string: "frames/sec at 1080p (depth_anything + flow_raft)": both bands over the same
                # 1080p clip on one GPU, one after the other per frame -> 1 / (1/fps_depth + 1/fps_flow)
                eng.time_resident(1080, 1920, 3, BATCH)
                ms1080 = eng.time_resident(1080, 1920, 6, BATCH)
                da1080 = BATCH / (ms1080 * 1e-3)
                fl1080 = out["extra"]["flow_raft_1080p"]["frame_steps_per_s_device"]
                return {"workload": "synthetic 1080p clip, depth_anything ViT-L then flow_raft (12 iterations, video pass) per "
                                    "frame, 1 GPU, frames resident",
                        "depth_anything_frames_per_s": da1080, "flow_raft_frames_per_s": fl1080,
                        "combined_frames_per_s": 1.0 / (1.0 / da1080 + 1.0 / fl1080)}

            extra("pipeline_1080p_depth_plus_flow", pipeline_1080p)
            extra("depth_midas_720p", lambda: midas_extras(local_rank))
            extra("mask_mmdet_1080p", lambda: mask_extras(local_rank))
            extra("depth_anything_metric_720p", lambda: zoe_extras(local_rank))
        if world == 1 and not args.no_cpu:
            cores = cpu_threads()
            fps, dt = cpu_baseline_frames(3, cores)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"3 frames of the 720p clip through oracle/da.py (torch CPU fp32, {cores} threads of "
                                             f"{os.cpu_count()} host cores: the measured optimum, 128 threads are 7x slower), {dt:.1f} s"}
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
