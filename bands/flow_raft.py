#!/usr/bin/env python3
"""flow_raft band on the B200 engine -- drop-in for the reference's bands/flow_raft.py.

Same plugin surface: BAND / ITERATIONS / MODEL constants, init_model(args), infer(args, prev, curr),
process_video(args), the CLI flags of bands/flow_raft.py:170-189, the outputs (flow_raft.mp4, flow_raft.csv with the
per-frame max displacement, optional flow_raft_bwd.mp4, flow_raft_mask[_bwd].mp4, <subpath>_fwd|_bwd/%04d.flo,
<subpath_mask>_fwd|_bwd/%04d.png) and the metadata.json keys (:143-166).
The model (RAFT, 12-20 GRU iterations), the x`scale` cubic resize, the HSV encode, the forward/backward consistency
masks and the 16-bit flow PNG payload are computed in libprisma_b200.so.  --small raises (commented out in the
reference too); --alternate_corr /
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


def infer(args, prev_frame, curr_frame, want_rgb=False, reuse_prev=False):
    """(fwd_flow, bwd_flow, fwd_mask, bwd_mask) like the reference's infer (:51-66), plus the engine's result dict;
    frames are HxWx3 u8 RGB at the source resolution (the x scale resize of :100 happens inside the engine)."""
    from prisma_b200.flow import consistency_masks
    r = model.infer_pair(prev_frame, curr_frame, want_rgb=want_rgb, reuse_prev=reuse_prev)
    fwd_mask = bwd_mask = None
    if args.output_mask != "" or args.subpath_mask != "":
        fwd_mask, bwd_mask, r["fwd_u16"], r["bwd_u16"] = consistency_masks(r["fwd"], r["bwd"], want_u16=True,
                                                                           device=args.device)
    return r["fwd"], r["bwd"], fwd_mask, bwd_mask, r


def write_flow(args, r, fwd_video, max_disps, idx, fwd_mask=None, fwd_mask_video=None, bwd_video=None, bwd_mask=None,
               bwd_mask_video=None):
    """One frame's outputs (reference common/flow.py:64-98).  The reference's .flo branch (:90-93) calls this very
    function with (path, flow) and cannot run; the Middlebury writer it means (common/io.py:175-198) is used here."""
    from prisma_b200.flow import write_flo
    fwd_video.write(r["fwd_rgb"])  # VideoWriter rescales to the source size, as the reference's does
    max_disps.append(r["max_fwd"])
    if fwd_mask is not None and fwd_mask_video:
        fwd_mask_video.write(np.repeat((fwd_mask.astype(np.uint8) * 255)[..., None], 3, axis=-1))
    if bwd_mask is not None and bwd_mask_video:
        bwd_mask_video.write(np.repeat((bwd_mask.astype(np.uint8) * 255)[..., None], 3, axis=-1))
    if args.backwards and bwd_video:
        bwd_video.write(r["bwd_rgb"])
    if args.subpath != "":
        write_flo(os.path.join(args.subpath + "_fwd", "%04d.flo" % idx), r["fwd"])
        if args.backwards:
            write_flo(os.path.join(args.subpath + "_bwd", "%04d.flo" % idx), r["bwd"])
    if args.subpath_mask != "":
        import cv2
        cv2.imwrite(os.path.join(args.subpath_mask + "_fwd", "%04d.png" % idx), r["fwd_u16"])
        if args.backwards:
            cv2.imwrite(os.path.join(args.subpath_mask + "_bwd", "%04d.png" % idx), r["bwd_u16"])


def process_video(args):
    reader = VideoReader(args.input)
    base = args.output.rsplit(".", 1)[0]
    fps = reader.get_avg_fps()
    new_writer = lambda name: VideoWriter(reader.width, reader.height, fps, name)
    fwd_video = new_writer(args.output)
    fwd_mask_video = new_writer(args.output_mask) if args.output_mask != "" else None
    bwd_video = new_writer(base + "_bwd.mp4") if args.backwards else None
    bwd_mask_video = None
    if args.backwards and args.output_mask != "":
        bwd_mask_video = new_writer(args.output_mask.rsplit(".", 1)[0] + "_bwd.mp4")
    want_masks = args.output_mask != "" or args.subpath_mask != ""
    max_disps = []
    prev = None
    i = 0
    for i, frame in enumerate(reader):
        if prev is not None:
            # from the second pair on, `prev` is the previous call's `curr`: its encoder features are still in the engine
            _, _, fwd_mask, bwd_mask, r = infer(args, prev, frame, want_rgb=True, reuse_prev=(i > 1))
            write_flow(args, r, fwd_video, max_disps, i - 1, fwd_mask, fwd_mask_video, bwd_video, bwd_mask, bwd_mask_video)
        prev = frame
    if prev is not None:
        # last frame (:116-126): zero flow (the reference's 0/0 -> NaN -> u8 cast gives a black frame), all-false masks
        # at the SOURCE resolution, as the reference allocates them
        hs, ws = reader.height, reader.width
        zero = np.zeros((hs, ws, 2), np.float32)
        black = np.zeros((hs, ws, 3), np.uint8)
        u16 = np.zeros((hs, ws, 3), np.uint16)
        u16[..., :2] = 32768
        r = dict(fwd=zero, bwd=zero, fwd_rgb=black, bwd_rgb=black, max_fwd=0.0, fwd_u16=u16, bwd_u16=u16)
        mask = np.zeros((hs, ws), bool) if want_masks else None
        write_flow(args, r, fwd_video, max_disps, i, mask, fwd_mask_video, bwd_video, mask, bwd_mask_video)
    for v in (fwd_video, fwd_mask_video, bwd_video, bwd_mask_video):
        if v:
            v.close()
    with open(base + ".csv", "w") as f:
        f.writelines("{}\n".format(v) for v in max_disps)
    if data:
        data["bands"][BAND] = {"url": BAND + ".mp4", "values": {"dist": {"type": "float", "url": BAND + ".csv"}}}
        if args.subpath != "":
            data["bands"][BAND]["folder"] = args.subpath
        if args.backwards:
            data["bands"][BAND + "_bwd"] = {"url": BAND + "_bwd.mp4"}
            if args.subpath != "":
                data["bands"][BAND + "_bwd"]["folder"] = args.subpath + "_bwd"
        if args.output_mask != "":
            data["bands"][BAND + "_mask"] = {"url": BAND + "_mask.mp4"}
            if args.backwards:
                data["bands"][BAND + "_mask_bwd"] = {"url": BAND + "_mask_bwd.mp4"}


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
    if args.small:
        raise NotImplementedError("the small RAFT variant is commented out in the reference (raft.py:28-33) and not built")
    data = load_metadata(args.input)
    if data:
        args.input = get_url(args.input, data, "rgba")
        args.output = get_target(args.input, data, band=BAND, target=args.output)
        if args.mask:
            args.output_mask = get_target(args.input, data, band=BAND + "_mask")
    elif args.output == "":
        args.output = os.path.join(os.path.dirname(args.input), BAND + ".mp4")
    input_folder = os.path.dirname(args.input)
    for attr in ("subpath", "subpath_mask"):  # reference :208-218
        if getattr(args, attr) != "":
            setattr(args, attr, os.path.join(input_folder, getattr(args, attr)))
            os.makedirs(getattr(args, attr) + "_fwd", exist_ok=True)
            if args.backwards:
                os.makedirs(getattr(args, attr) + "_bwd", exist_ok=True)
    init_model(args)
    process_video(args)
    write_metadata(args.input, data)


if __name__ == "__main__":
    main()
