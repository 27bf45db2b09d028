"""The per-frame loop shared by the depth bands (reference bands/depth_anything.py:176-251 and bands/depth_midas.py:108-175
are the same loop): predict, write <sub>/%05d.npy and <sub>/%05d.png (write_depth encoding), append the heat-encoded
frame to the video, collect (min, max) into <band>_min.csv / <band>_max.csv and the metadata keys.

Frames are fed to the engine in chunks (prisma_depth_infer_stream: copies overlapped with compute); a chunk of the
reference's loop carries no state, so the results are those of the frame-by-frame loop.

Multi-GPU: frames are independent, so `--gpus N` runs N workers of the same script
(bands/common/sharded.py: rank r on GPU r over the contiguous frame range of prisma_b200.shard.frame_range); the encoded
frames and the (min, max) rows of ranks 1.. reach the writer rank in memory over torch.distributed (NCCL on the GPUs),
which appends them to the one video / csv pair.  No data-path collective, nothing through the file system."""
import os

import numpy as np

from .media import VideoReader, VideoWriter, create_folder, write_rgb
from .sharded import OrderedStreams, ShardContext


def process_depth_video(model, a, data, band, ctx=None, chunk=24, pass_frames=12, flip=True):
    ctx = ctx or ShardContext()
    reader = VideoReader(a.input)
    first, last, _ = ctx.frames(len(reader))
    if first > 0:
        reader.seek(first)
    streams = OrderedStreams(ctx, {"depth": lambda: VideoWriter(reader.width, reader.height, reader.get_avg_fps(), a.output)})
    folder = os.path.dirname(a.output)
    sub = ""
    if a.subpath != "":
        if data:
            data["bands"][band]["folder"] = a.subpath
        sub = os.path.join(folder, a.subpath)
        create_folder(sub)
    want_pred = bool(a.npy or sub)
    state = dict(index=first, pass_frames=pass_frames)

    def flush(frames):
        if state["index"] == first:
            state["pass_frames"] = min(pass_frames, len(frames))  # short clips: one pass; the pass size then stays fixed
        rgb, mn, mx, pred = model.infer_clip(np.ascontiguousarray(np.stack(frames)), pass_frames=state["pass_frames"],
                                             want_depth=want_pred)
        for k in range(len(frames)):
            index = state["index"]
            if a.npy:
                np.save(os.path.join(sub or folder, "{:05d}.npy".format(index)), pred[k])
            if sub:  # reference :222-223 / depth_midas.py:150-151: write_depth(normalize, flip, heatmap, encode_range)
                png, _, _ = model.encode_png(pred[k], flip=flip)
                write_rgb(os.path.join(sub, "{:05d}.png".format(index)), png)
            streams.write("depth", rgb[k])
            streams.scalars(mn[k], mx[k])
            state["index"] += 1

    pending = []
    if last > first:
        for n_read, frame in enumerate(reader):
            pending.append(frame)
            if len(pending) == chunk:
                flush(pending)
                pending = []
            if first + n_read + 1 >= last:
                break
        if pending:
            flush(pending)
    table = streams.finish()
    if not ctx.is_writer():
        return
    with open(os.path.join(folder, band + "_min.csv"), "w") as f:
        f.writelines("{}\n".format(row[0]) for row in table)
    with open(os.path.join(folder, band + "_max.csv"), "w") as f:
        f.writelines("{}\n".format(row[1]) for row in table)
    if data:
        data["bands"][band]["values"] = {"min": {"type": "float", "url": band + "_min.csv"},
                                         "max": {"type": "float", "url": band + "_max.csv"}}
