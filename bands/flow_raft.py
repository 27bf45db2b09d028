#!/usr/bin/env python3
"""flow_raft band on the B200 engine -- drop-in for the reference's bands/flow_raft.py.

Same plugin surface: BAND / ITERATIONS / MODEL constants, init_model(args), infer(args, prev, curr),
process_video(args), the CLI flags of bands/flow_raft.py:170-189, the outputs (flow_raft.mp4, flow_raft.csv with the
per-frame max displacement, optional flow_raft_bwd.mp4) and the metadata.json keys (:143-166).
The model (RAFT, 12-20 GRU iterations), the x`scale` cubic resize and the HSV encode run in libprisma_b200.so.

Not built yet: --mask / --output_mask / --subpath_mask (fwd/bwd consistency masks) and
.flo export (--subpath); the script raises for them instead of falling back.  --small / --alternate_corr /
--mixed_precision are accepted for CLI compatibility: the engine has one (fp16-operand, fp32-accumulate) path and
always builds the correlation pyramid on the GPU.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bands.common.meta import get_target, get_url, load_metadata, write_metadata  # noqa: E402
from bands.common.media import VideoReader, VideoWriter  # noqa: E402

BAND = "flow_raft"
ITERATIONS = 20
MODEL = "models/raft-sintel.pth"

model = None
data = None


def _load_state_dict(a):
    if a.seeded_weights:
        from prisma_b200.seeded_weights import make_raft_weights
        return make_raft_weights(0)
    import torch
    return torch.load(a.model, map_location="cpu")  # keys carry the DataParallel "module." prefix (:42-44)


def init_model(args):
    global model
    from prisma_b200.flow import RaftFlowEngine
    model = RaftFlowEngine(_load_state_dict(args), device=args.device, iterations=args.iterations, scale=args.scale)
    return model


def infer(args, prev_frame, curr_frame):
    """(fwd_flow, bwd_flow, fwd_mask, bwd_mask) like the reference's infer (:51-66); frames are HxWx3 u8 RGB at the
    source resolution (the x scale resize of :100 happens inside the engine)."""
    r = model.infer_pair(prev_frame, curr_frame)
    return r["fwd"], r["bwd"], None, None


def process_video(args):
    reader = VideoReader(args.input)
    base = args.output.rsplit(".", 1)[0]
    fwd_video = VideoWriter(reader.width, reader.height, reader.get_avg_fps(), args.output)
    bwd_video = VideoWriter(reader.width, reader.height, reader.get_avg_fps(), base + "_bwd.mp4") if args.backwards else None
    max_disps = []
    prev = None
    hs = ws = 0
    for frame in reader:
        if prev is not None:
            r = model.infer_pair(prev, frame, want_rgb=True)
            fwd_video.write(r["fwd_rgb"])  # VideoWriter rescales to the source size, as the reference's does
            if bwd_video:
                bwd_video.write(r["bwd_rgb"])
            max_disps.append(r["max_fwd"])
            hs, ws = r["fwd_rgb"].shape[:2]
        prev = frame
    if prev is not None:  # last frame: zero flow (:116-126; the reference's 0/0 -> NaN -> u8 cast gives a black frame)
        if hs == 0:
            hs, ws = model.out_size(reader.height, reader.width)
        black = np.zeros((hs, ws, 3), np.uint8)
        fwd_video.write(black)
        if bwd_video:
            bwd_video.write(black)
        max_disps.append(0.0)
    fwd_video.close()
    if bwd_video:
        bwd_video.close()
    with open(base + ".csv", "w") as f:
        f.writelines("{}\n".format(v) for v in max_disps)
    if data:
        data["bands"][BAND] = {"url": BAND + ".mp4", "values": {"dist": {"type": "float", "url": BAND + ".csv"}}}
        if args.backwards:
            data["bands"][BAND + "_bwd"] = {"url": BAND + "_bwd.mp4"}


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--input", "-i", help="input", type=str, required=True)
    p.add_argument("--output", "-o", help="output", type=str, default="")
    p.add_argument("--subpath", help="path to flo files", type=str, default="")
    p.add_argument("--backwards", "-b", help="Backward video", action="store_true")
    p.add_argument("--mask", action="store_true", help="Compute mask as well")
    p.add_argument("--output_mask", help="output dense", type=str, default="")
    p.add_argument("--subpath_mask", help="path to flo files", type=str, default="")
    p.add_argument("--iterations", help="number of iterations", type=int, default=ITERATIONS)
    p.add_argument("--model", "-m", help="model path", type=str, default=MODEL)
    p.add_argument("--scale", type=float, default=0.75)
    p.add_argument("--raft_model", default="models/raft-things.pth", help="[RAFT] restore checkpoint")
    p.add_argument("--small", action="store_true", help="[RAFT] use small model")
    p.add_argument("--mixed_precision", action="store_true", help="[RAFT] use mixed precision")
    p.add_argument("--alternate_corr", action="store_true", help="[RAFT] use efficent correlation implementation")
    p.add_argument("--seeded-weights", action="store_true", help="seeded random weights (offline testing)")
    p.add_argument("--device", type=int, default=0)
    return p


def main(argv=None):
    global data
    args = build_parser().parse_args(argv)
    if args.mask or args.output_mask or args.subpath_mask or args.subpath:
        raise NotImplementedError("consistency masks / .flo export are not built yet (SURVEY.md section 8f row 2)")
    if args.small:
        raise NotImplementedError("the small RAFT variant is commented out in the reference (raft.py:28-33) and not built")
    data = load_metadata(args.input)
    if data:
        args.input = get_url(args.input, data, "rgba")
        args.output = get_target(args.input, data, band=BAND, target=args.output)
    elif args.output == "":
        args.output = os.path.join(os.path.dirname(args.input), BAND + ".mp4")
    init_model(args)
    process_video(args)
    write_metadata(args.input, data)


if __name__ == "__main__":
    main()
