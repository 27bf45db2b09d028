"""The per-frame loop shared by the depth bands (reference bands/depth_anything.py:176-251 and bands/depth_midas.py:108-175
are the same loop): predict, write <sub>/%05d.npy and <sub>/%05d.png (write_depth encoding), append the heat-encoded
frame to the video, collect (min, max) into <band>_min.csv / <band>_max.csv and the metadata keys.

Frames are fed to the engine in chunks (prisma_depth_infer_stream: copies overlapped with compute); a chunk of the
reference's loop carries no state, so the results are those of the frame-by-frame loop.

Multi-GPU: frames are independent, so `--gpus N` runs N workers of the same script, worker r on
GPU r over the contiguous frame range of prisma_b200.shard.frame_range; a worker (`--frames s:e --part r`) writes its encoded
frames raw (<output>.part<r>.u8) and its (min, max) rows (<band>.part<r>.csv); the parent concatenates them into the one
video / csv pair.  No data-path collective."""
import os

import numpy as np

from .media import VideoReader, VideoWriter, create_folder, write_rgb


class _RawFrames:
    """Worker-side stand-in for the VideoWriter: encoded frames appended raw to <output>.part<r>.u8."""

    def __init__(self, path):
        self.f = open(path, "wb")

    def write(self, rgb):
        self.f.write(np.ascontiguousarray(rgb, dtype=np.uint8).tobytes())

    def close(self):
        self.f.close()


def parse_frames(spec, total):
    """'s:e' -> (s, e) clipped to the clip; '' -> the whole clip."""
    if not spec:
        return 0, total
    s, e = spec.split(":")
    return max(0, int(s)), min(total, int(e))


def process_depth_video(model, a, data, band, chunk=24, pass_frames=12, flip=True):
    reader = VideoReader(a.input)
    part = getattr(a, "part", -1)
    first, last = parse_frames(getattr(a, "frames", ""), len(reader))
    if first > 0:
        reader.seek(first)
    out = _RawFrames(a.output + ".part%d.u8" % part) if part >= 0 else VideoWriter(reader.width, reader.height, reader.get_avg_fps(), a.output)
    folder = os.path.dirname(a.output)
    sub = ""
    if a.subpath != "":
        if data:
            data["bands"][band]["folder"] = a.subpath
        sub = os.path.join(folder, a.subpath)
        create_folder(sub)
    want_pred = bool(a.npy or sub)
    mins, maxs = [], []
    index = first

    def flush(frames):
        nonlocal index, pass_frames
        if index == first:
            pass_frames = min(pass_frames, len(frames))  # short clips: one pass; the pass size then stays fixed
        rgb, mn, mx, pred = model.infer_clip(np.ascontiguousarray(np.stack(frames)), pass_frames=pass_frames,
                                             want_depth=want_pred)
        for k in range(len(frames)):
            if a.npy:
                np.save(os.path.join(sub or folder, "{:05d}.npy".format(index)), pred[k])
            if sub:  # reference :222-223 / depth_midas.py:150-151: write_depth(normalize, flip, heatmap, encode_range)
                png, _, _ = model.encode_png(pred[k], flip=flip)
                write_rgb(os.path.join(sub, "{:05d}.png".format(index)), png)
            out.write(rgb[k])
            mins.append(float(mn[k]))
            maxs.append(float(mx[k]))
            index += 1

    pending = []
    for n_read, frame in enumerate(reader):
        if first + n_read >= last:
            break
        pending.append(frame)
        if len(pending) == chunk:
            flush(pending)
            pending = []
    if pending:
        flush(pending)
    out.close()
    if part >= 0:  # worker: its rows only; the parent assembles the csv files, the video and the metadata
        with open(os.path.join(folder, band + ".part%d.csv" % part), "w") as f:
            f.writelines("{},{}\n".format(lo, hi) for lo, hi in zip(mins, maxs))
        return
    with open(os.path.join(folder, band + "_min.csv"), "w") as f:
        f.writelines("{}\n".format(v) for v in mins)
    with open(os.path.join(folder, band + "_max.csv"), "w") as f:
        f.writelines("{}\n".format(v) for v in maxs)
    if data:
        data["bands"][band]["values"] = {"min": {"type": "float", "url": band + "_min.csv"},
                                         "max": {"type": "float", "url": band + "_max.csv"}}


def strip_shard_flags(argv):
    """The parent's command line without the flags the parent itself consumes / the workers get rewritten."""
    out, skip = [], False
    for x in argv:
        if skip:
            skip = False
            continue
        if x in ("--gpus", "--device-list", "--device", "--output", "-o"):
            skip = True
            continue
        out.append(x)
    return out


def run_sharded(script, argv, a, data, band, gpus, devices=None):
    """`--gpus N`: N workers of `script` over contiguous frame ranges (worker r on GPU devices[r]), then one video + csv pair."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from prisma_b200.shard import frame_range
    reader = VideoReader(a.input)
    total = len(reader)
    devices = devices or list(range(gpus))
    folder = os.path.dirname(a.output)
    procs = []
    for r in range(gpus):
        s, e, _ = frame_range(r, gpus, total)
        if s >= e:
            continue
        cmd = [sys.executable, script] + list(argv) + ["--frames", "%d:%d" % (s, e), "--part", str(r),
                                               "--device", str(devices[r % len(devices)]), "--output", a.output]
        procs.append((r, s, e, subprocess.Popen(cmd)))
    for r, s, e, p in procs:
        if p.wait() != 0:
            raise RuntimeError("frame-range worker %d (frames %d:%d) failed" % (r, s, e))
    out = VideoWriter(reader.width, reader.height, reader.get_avg_fps(), a.output)
    mins, maxs = [], []
    fb = reader.height * reader.width * 3
    for r, s, e, _ in procs:
        raw = a.output + ".part%d.u8" % r
        frames = np.memmap(raw, np.uint8, "r").reshape(e - s, reader.height, reader.width, 3) if os.path.getsize(raw) == (e - s) * fb else None
        if frames is None:
            raise RuntimeError("worker %d wrote %d bytes, expected %d" % (r, os.path.getsize(raw), (e - s) * fb))
        for k in range(e - s):
            out.write(np.asarray(frames[k]))
        del frames
        os.remove(raw)
        csvp = os.path.join(folder, band + ".part%d.csv" % r)
        for line in open(csvp):
            lo, hi = line.strip().split(",")
            mins.append(lo)
            maxs.append(hi)
        os.remove(csvp)
    out.close()
    with open(os.path.join(folder, band + "_min.csv"), "w") as f:
        f.writelines(v + "\n" for v in mins)
    with open(os.path.join(folder, band + "_max.csv"), "w") as f:
        f.writelines(v + "\n" for v in maxs)
    if data:
        if a.subpath != "":
            data["bands"][band]["folder"] = a.subpath
        data["bands"][band]["values"] = {"min": {"type": "float", "url": band + "_min.csv"},
                                         "max": {"type": "float", "url": band + "_max.csv"}}
