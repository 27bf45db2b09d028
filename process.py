#!/usr/bin/env python3
"""PRISMA orchestrator surface (reference process.py:60-99,163-290) for the bands this repo accelerates.

Creates the PRISMA folder + metadata.json, stores the input as the `rgba` band and launches one subprocess per
requested band (`python3 bands/<band>.py -i <folder> ...`), exactly the reference's process boundary.
Bands outside the hot-path scope are reported as 'not accelerated' and skipped.
"""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bands.common.meta import add_band, create_metadata, is_video, set_default_band, write_metadata  # noqa: E402
from bands.common.media import VideoReader, open_rgb  # noqa: E402

ACCELERATED = {"depth_anything", "depth_midas", "flow_raft", "mask_mmdet"}


def run(band, folder, extra=()):
    if band not in ACCELERATED:
        print(f"[process] band '{band}' is not accelerated by prisma_b200 yet; skipped")
        return 0
    cmd = [sys.executable, os.path.join(ROOT, "bands", band + ".py"), "-i", folder] + list(extra)
    print("[process]", " ".join(cmd))
    return subprocess.call(cmd)


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--input", "-i", type=str, required=True)
    p.add_argument("--output", "-o", type=str, default="")
    p.add_argument("--depth", "-d", type=str, default="depth_anything")
    p.add_argument("--flow", "-f", type=str, default="")
    p.add_argument("--encoder", type=str, default="vitl")
    p.add_argument("--mask", action="store_true", help="also run the mask band (the reference always does, process.py:207)")
    p.add_argument("--seeded-weights", action="store_true")
    a = p.parse_args(argv)
    base, ext = os.path.splitext(os.path.basename(a.input))
    folder = a.output or os.path.join(os.path.dirname(a.input), base)
    data = create_metadata(folder)
    rgba = "rgba" + (".mp4" if is_video(a.input) else ".png")
    if not os.path.exists(os.path.join(folder, rgba)):
        shutil.copyfile(a.input, os.path.join(folder, rgba))  # the reference transcodes (rgba.py:78-100); codec is out of scope
    add_band(data, "rgba", url=rgba)
    if is_video(a.input):
        r = VideoReader(a.input)
        data.update(width=r.width, height=r.height, fps=r.fps, frames=len(r), duration=len(r) / r.fps)
    else:
        img = open_rgb(a.input)
        data.update(width=img.shape[1], height=img.shape[0])
    write_metadata(folder, data)
    if a.mask:  # reference process.py:207: run("mask_mmdet", folder, subpath=True, ...)
        run("mask_mmdet", folder, ["--subpath", "mask"] + (["--seeded-weights"] if a.seeded_weights else []))
    extra = ["--encoder", a.encoder] + (["--seeded-weights"] if a.seeded_weights else [])
    bands = ["depth_anything", "depth_midas"] if a.depth == "all" else [a.depth]
    for b in bands:
        run(b, folder, extra if b == "depth_anything" else (["--seeded-weights"] if a.seeded_weights else []))
    if "depth_anything" in bands:
        set_default_band(folder, "depth", "depth_anything")
    if a.flow:
        flows = ["flow_gmflow", "flow_raft"] if a.flow == "all" else [a.flow]
        for fb in flows:
            run(fb, folder, ["--backwards"] + (["--seeded-weights"] if a.seeded_weights else []))
        if "flow_raft" in flows:
            set_default_band(folder, "flow", "flow_raft")
            set_default_band(folder, "flow_bwd", "flow_raft_bwd")


if __name__ == "__main__":
    main()
