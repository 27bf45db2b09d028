"""The per-frame loop shared by the depth bands (reference bands/depth_anything.py:176-251 and bands/depth_midas.py:108-175
are the same loop): predict, write <sub>/%05d.npy and <sub>/%05d.png (write_depth encoding), append the heat-encoded
frame to the video, collect (min, max) into <band>_min.csv / <band>_max.csv and the metadata keys.

Frames are fed to the engine in chunks (prisma_depth_infer_stream: copies overlapped with compute); a chunk of the
reference's loop carries no state, so the results are those of the frame-by-frame loop."""
import os

import numpy as np

from .media import VideoReader, VideoWriter, create_folder, write_rgb


def process_depth_video(model, a, data, band, chunk=24, pass_frames=12, flip=True):
    reader = VideoReader(a.input)
    out = VideoWriter(reader.width, reader.height, reader.get_avg_fps(), a.output)
    folder = os.path.dirname(a.output)
    sub = ""
    if a.subpath != "":
        if data:
            data["bands"][band]["folder"] = a.subpath
        sub = os.path.join(folder, a.subpath)
        create_folder(sub)
    want_pred = bool(a.npy or sub)
    mins, maxs = [], []
    index = 0

    def flush(frames):
        nonlocal index, pass_frames
        if index == 0:
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
    for frame in reader:
        pending.append(frame)
        if len(pending) == chunk:
            flush(pending)
            pending = []
    if pending:
        flush(pending)
    out.close()
    with open(os.path.join(folder, band + "_min.csv"), "w") as f:
        f.writelines("{}\n".format(v) for v in mins)
    with open(os.path.join(folder, band + "_max.csv"), "w") as f:
        f.writelines("{}\n".format(v) for v in maxs)
    if data:
        data["bands"][band]["values"] = {"min": {"type": "float", "url": band + "_min.csv"},
                                         "max": {"type": "float", "url": band + "_max.csv"}}
