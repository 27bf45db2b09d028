"""Plugin surface: metadata contract (CPU) and the depth_anything band end to end into a temp PRISMA folder (GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_metadata_contract(tmp_path):
    from bands.common import meta
    folder = str(tmp_path / "clip")
    data = meta.create_metadata(folder)
    assert data == {"bands": {}}
    meta.add_band(data, "rgba", url="rgba.mp4")
    meta.write_metadata(folder, data)
    assert meta.get_url(folder, meta.load_metadata(folder), "rgba") == os.path.join(folder, "rgba.mp4")
    tgt = meta.get_target(os.path.join(folder, "rgba.mp4"), data, band="depth_anything", force_extension="png")
    assert tgt == os.path.join(folder, "depth_anything.mp4") and data["bands"]["depth_anything"]["url"] == "depth_anything.mp4"
    tgt = meta.get_target(os.path.join(folder, "rgba.png"), data, band="depth_anything", force_extension="png")
    assert tgt.endswith("depth_anything.png")
    assert meta.is_video("a.mp4") and not meta.is_video("a.png")


def test_flow_band_cli_matches_reference_flags():
    sys.path.insert(0, ROOT)
    from bands import flow_raft as band
    a = band.build_parser().parse_args(["-i", "x", "-b", "--iterations", "12", "--scale", "0.5", "-m", "w.pth"])
    assert (a.backwards, a.iterations, a.scale, a.model) == (True, 12, 0.5, "w.pth")
    assert band.BAND == "flow_raft" and band.ITERATIONS == 20 and band.MODEL == "models/raft-sintel.pth"


def test_band_cli_matches_reference_flags():
    sys.path.insert(0, ROOT)
    from bands import depth_anything as band
    a = band.build_parser().parse_args(["-i", "x.mp4", "-o", "y.mp4", "-n", "-d", "frames", "--encoder", "vits", "--metric", "none"])
    assert (a.input, a.output, a.npy, a.subpath, a.encoder, a.metric) == ("x.mp4", "y.mp4", True, "frames", "vits", "none")
    assert band.BAND == "depth_anything"


@pytest.mark.gpu
def test_process_runs_band_into_prisma_folder(tmp_path):
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    src = str(tmp_path / "clip.mp4")
    w = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
    for t in range(3):
        w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
    w.release()
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "process.py"), "-i", src, "--encoder", "vits", "--seeded-weights",
                          "-f", "flow_raft", "-b"])
    assert rc == 0
    folder = str(tmp_path / "clip")
    meta = json.load(open(os.path.join(folder, "metadata.json")))
    band = meta["bands"]["depth_anything"]
    assert band["url"] == "depth_anything.mp4" and os.path.exists(os.path.join(folder, band["url"]))
    assert band["values"]["min"] == {"type": "float", "url": "depth_anything_min.csv"}
    mins = [float(l) for l in open(os.path.join(folder, "depth_anything_min.csv"))]
    maxs = [float(l) for l in open(os.path.join(folder, "depth_anything_max.csv"))]
    assert len(mins) == 3 and all(b > a for a, b in zip(mins, maxs))
    assert meta["bands"]["depth"] == band and meta["width"] == 320 and meta["frames"] == 3
    flow = meta["bands"]["flow_raft"]
    assert flow == {"url": "flow_raft.mp4", "values": {"dist": {"type": "float", "url": "flow_raft.csv"}}}
    dists = [float(l) for l in open(os.path.join(folder, "flow_raft.csv"))]
    assert len(dists) == 3 and dists[-1] == 0.0 and all(d > 0 for d in dists[:2])
    assert meta["bands"]["flow_raft_bwd"] == {"url": "flow_raft_bwd.mp4"} and meta["bands"]["flow"] == flow


@pytest.mark.gpu
def test_flow_band_masks_flo_and_u16_png(tmp_path):
    """--mask / --subpath / --subpath_mask outputs of bands/flow_raft.py (reference :60-66, common/flow.py:64-98)."""
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    folder = tmp_path / "clip"
    folder.mkdir()
    src = str(folder / "rgba.mp4")
    w = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
    for t in range(3):
        w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
    w.release()
    json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 320, "height": 240, "frames": 3, "fps": 24.0},
              open(folder / "metadata.json", "w"))
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "flow_raft.py"), "-i", str(folder), "-b", "--mask",
                          "--subpath", "flo", "--subpath_mask", "flow_png", "--seeded-weights", "--iterations", "6"])
    assert rc == 0
    meta = json.load(open(folder / "metadata.json"))
    assert meta["bands"]["flow_raft_mask"] == {"url": "flow_raft_mask.mp4"}
    assert meta["bands"]["flow_raft_mask_bwd"] == {"url": "flow_raft_mask_bwd.mp4"}
    assert meta["bands"]["flow_raft"]["folder"].endswith("flo") and meta["bands"]["flow_raft_bwd"]["folder"].endswith("flo_bwd")
    for name in ("flow_raft.mp4", "flow_raft_bwd.mp4", "flow_raft_mask.mp4", "flow_raft_mask_bwd.mp4"):
        assert os.path.getsize(folder / name) > 0
    for d in ("fwd", "bwd"):
        flo = np.fromfile(folder / ("flo_" + d) / "0000.flo", np.float32)
        assert flo[0] == np.float32(202021.25)
        wh = flo[1:3].view(np.int32)
        assert tuple(wh) == (240, 180) and flo.size == 3 + 240 * 180 * 2      # x0.75 working resolution
        png = cv2.imread(str(folder / ("flow_png_" + d) / "0000.png"), cv2.IMREAD_UNCHANGED)
        assert png.dtype == np.uint16 and png.shape == (180, 240, 3)
        # the PNG payload is encode_flow of that same .flo (u16 truncation of 2^15 + 256 f)
        f = flo[3:].reshape(180, 240, 2)
        assert np.array_equal(png[..., 0], (2 ** 15 + f[..., 0] * 2 ** 8).astype(np.uint16))
        assert len(os.listdir(folder / ("flo_" + d))) == 3


def test_midas_band_cli_matches_reference_flags():
    sys.path.insert(0, ROOT)
    from bands import depth_midas as band
    a = band.build_parser().parse_args(["-i", "x.mp4", "-o", "y.mp4", "-n", "-d", "frames", "--model", "midas3"])
    assert (a.input, a.output, a.npy, a.subpath, a.model) == ("x.mp4", "y.mp4", True, "frames", "midas3")
    assert band.BAND == "depth_midas" and band.MODELS_VERSIONS == ["midas2-small", "midas2", "midas3-small", "midas3"]
    with pytest.raises(NotImplementedError):
        band.init_model("midas2")


@pytest.mark.gpu
def test_midas_band_writes_video_csv_and_frames(tmp_path):
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    folder = tmp_path / "clip"
    folder.mkdir()
    w = cv2.VideoWriter(str(folder / "rgba.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
    for t in range(3):
        w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
    w.release()
    json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 320, "height": 240, "frames": 3, "fps": 24.0},
              open(folder / "metadata.json", "w"))
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "depth_midas.py"), "-i", str(folder), "-n", "-d",
                          "depth_midas_frames", "--seeded-weights"])
    assert rc == 0
    meta = json.load(open(folder / "metadata.json"))
    band = meta["bands"]["depth_midas"]
    assert band["url"] == "depth_midas.mp4" and band["folder"] == "depth_midas_frames"
    assert band["values"] == {"min": {"type": "float", "url": "depth_midas_min.csv"},
                              "max": {"type": "float", "url": "depth_midas_max.csv"}}
    mins = [float(l) for l in open(folder / "depth_midas_min.csv")]
    maxs = [float(l) for l in open(folder / "depth_midas_max.csv")]
    assert len(mins) == 3 and all(b > a for a, b in zip(mins, maxs))
    frames = sorted(os.listdir(folder / "depth_midas_frames"))
    assert frames == ["00000.npy", "00000.png", "00001.npy", "00001.png", "00002.npy", "00002.png"]
    pred = np.load(folder / "depth_midas_frames" / "00001.npy")
    assert pred.shape == (240, 320) and pred.dtype == np.float32 and np.float32(pred.min()) == np.float32(mins[1])


def test_mask_band_cli_matches_reference_flags():
    sys.path.insert(0, ROOT)
    from bands import mask_mmdet as band
    a = band.build_parser().parse_args(["-i", "x.mp4", "-o", "y.mp4", "-c", "0.6", "--subpath", "mask"])
    assert (a.input, a.output, a.confidence, a.subpath, a.sdf) == ("x.mp4", "y.mp4", 0.6, "mask", False)
    assert band.BAND == "mask" and band.CONFIDENCE_THRESHOLD == 0.5 and len(band.CLASSES) == 11
    assert band.build_parser().parse_args(["-i", "x.png", "--sdf"]).sdf is True


@pytest.mark.gpu
def test_mask_band_writes_mask_video_and_colmap_frames(tmp_path):
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    folder = tmp_path / "clip"
    folder.mkdir()
    w = cv2.VideoWriter(str(folder / "rgba.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
    for t in range(2):
        w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
    w.release()
    json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 320, "height": 240, "frames": 2, "fps": 24.0},
              open(folder / "metadata.json", "w"))
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "mask_mmdet.py"), "-i", str(folder), "--subpath", "mask",
                          "--seeded-weights", "-c", "0.5"])
    assert rc == 0
    meta = json.load(open(folder / "metadata.json"))
    assert meta["bands"]["mask"] == {"url": "mask.mp4", "ids": ['person', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant',
                                                              'bear', 'zebra', 'giraffe'], "folder": "mask"}
    assert os.path.getsize(folder / "mask.mp4") > 0
    frames = sorted(os.listdir(folder / "mask"))
    assert frames == ["00000.png", "00001.png"]
    png = cv2.imread(str(folder / "mask" / "00000.png"))
    assert png.shape == (240, 320, 3) and set(np.unique(png)) <= {0, 1, 254, 255}   # 255 - (255 * count mod 256)


@pytest.mark.gpu
def test_depth_band_metric_path(tmp_path):
    """--metric outdoor (what the reference's process.py passes by default): ZoeDepth head, no flip in the encode."""
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    folder = tmp_path / "clip"
    folder.mkdir()
    w = cv2.VideoWriter(str(folder / "rgba.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
    for t in range(3):
        w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
    w.release()
    json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 320, "height": 240, "frames": 3, "fps": 24.0},
              open(folder / "metadata.json", "w"))
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "depth_anything.py"), "-i", str(folder), "--metric", "outdoor",
                          "--encoder", "vits", "--seeded-weights", "-n", "-d", "frames"])
    assert rc == 0
    mins = [float(l) for l in open(folder / "depth_anything_min.csv")]
    maxs = [float(l) for l in open(folder / "depth_anything_max.csv")]
    assert len(mins) == 3 and all(1.0 < a < b < 10.0 for a, b in zip(mins, maxs))      # metric depths of the seeded head
    pred = np.load(folder / "frames" / "00000.npy")
    assert pred.shape == (240, 320) and np.float32(pred.min()) == np.float32(mins[0])
    assert sorted(os.listdir(folder / "frames"))[:2] == ["00000.npy", "00000.png"]


@pytest.mark.gpu
def test_depth_band_sharded_over_two_workers_equals_single(tmp_path):
    """--gpus 2 (both workers on device 0 here): frame-range workers + parent assembly == the single-process run."""
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    outs = {}
    for tag, extra in (("one", []), ("two", ["--gpus", "2", "--device-list", "0,0"])):
        folder = tmp_path / tag
        folder.mkdir()
        w = cv2.VideoWriter(str(folder / "rgba.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
        for t in range(5):
            w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
        w.release()
        json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 320, "height": 240, "frames": 5, "fps": 24.0},
                  open(folder / "metadata.json", "w"))
        rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "depth_anything.py"), "-i", str(folder), "--encoder", "vits",
                              "--seeded-weights", "-n", "-d", "frames"] + extra)
        assert rc == 0
        outs[tag] = dict(mins=open(folder / "depth_anything_min.csv").read(), maxs=open(folder / "depth_anything_max.csv").read(),
                         npy=[np.load(folder / "frames" / ("%05d.npy" % i)) for i in range(5)],
                         meta=json.load(open(folder / "metadata.json"))["bands"]["depth_anything"],
                         left=sorted(f for f in os.listdir(folder) if "part" in f))
        cap = cv2.VideoCapture(str(folder / "depth_anything.mp4"))
        outs[tag]["frames"] = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    assert outs["one"]["mins"] == outs["two"]["mins"] and outs["one"]["maxs"] == outs["two"]["maxs"]
    assert all(np.array_equal(a, b) for a, b in zip(outs["one"]["npy"], outs["two"]["npy"]))
    assert outs["one"]["meta"] == outs["two"]["meta"] and outs["two"]["frames"] == 5 and outs["two"]["left"] == []


def _clip_folder(folder, n):
    import cv2
    from prisma_b200.synthetic import synthetic_frame
    folder.mkdir()
    w = cv2.VideoWriter(str(folder / "rgba.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (320, 240))
    for t in range(n):
        w.write(synthetic_frame(240, 320, t)[..., ::-1].copy())
    w.release()
    json.dump({"bands": {"rgba": {"url": "rgba.mp4"}}, "width": 320, "height": 240, "frames": n, "fps": 24.0},
              open(folder / "metadata.json", "w"))


def _video_frames(path):
    import cv2
    cap = cv2.VideoCapture(str(path))
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    return out


@pytest.mark.gpu
def test_flow_band_sharded_with_halo_equals_single(tmp_path):
    """flow_raft --gpus 2 (1-frame halo per shard, in-memory gather to the writer rank over torch.distributed) == the
    single-process run: same max-displacement csv, same .flo files, same videos, same metadata."""
    outs = {}
    for tag, extra in (("one", []), ("two", ["--gpus", "2", "--device-list", "0,0"])):
        folder = tmp_path / tag
        _clip_folder(folder, 5)
        rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "flow_raft.py"), "-i", str(folder), "-b", "--iterations", "4",
                              "--subpath", "flo", "--seeded-weights"] + extra)
        assert rc == 0
        outs[tag] = dict(csv=open(folder / "flow_raft.csv").read(),
                         flo=[open(folder / "flo_fwd" / ("%04d.flo" % i), "rb").read() for i in range(5)],
                         flo_b=[open(folder / "flo_bwd" / ("%04d.flo" % i), "rb").read() for i in range(5)],
                         fwd=_video_frames(folder / "flow_raft.mp4"), bwd=_video_frames(folder / "flow_raft_bwd.mp4"),
                         meta=json.load(open(folder / "metadata.json"))["bands"])
    a, b = outs["one"], outs["two"]
    assert a["csv"] == b["csv"] and len(a["csv"].split()) == 5 and float(a["csv"].split()[-1]) == 0.0
    assert a["flo"] == b["flo"] and a["flo_b"] == b["flo_b"]
    assert len(b["fwd"]) == 5 and len(b["bwd"]) == 5
    assert all(np.array_equal(x, y) for x, y in zip(a["fwd"], b["fwd"])) and all(np.array_equal(x, y) for x, y in zip(a["bwd"], b["bwd"]))
    # the reference stores the --subpath folders as given joined with the input folder (flow_raft.py:208-218): compare them
    # relative to their own clip folder
    strip = lambda meta, tag: json.loads(json.dumps(meta).replace(str(tmp_path / tag), "<clip>"))
    assert strip(a["meta"], "one") == strip(b["meta"], "two") and "flow_raft_bwd" in b["meta"]


@pytest.mark.gpu
def test_mask_band_sharded_equals_single(tmp_path):
    """mask_mmdet --gpus 2 == single process (frames are independent): same mask video, same COLMAP frames, same metadata."""
    outs = {}
    for tag, extra in (("one", []), ("two", ["--gpus", "2", "--device-list", "0,0"])):
        folder = tmp_path / tag
        _clip_folder(folder, 3)
        rc = subprocess.call([sys.executable, os.path.join(ROOT, "bands", "mask_mmdet.py"), "-i", str(folder), "--subpath", "mask",
                              "--sdf", "--seeded-weights"] + extra)
        assert rc == 0
        outs[tag] = dict(video=_video_frames(folder / "mask.mp4"),
                         png=[open(folder / "mask" / ("%05d.png" % i), "rb").read() for i in range(3)],
                         meta=json.load(open(folder / "metadata.json"))["bands"]["mask"])
    a, b = outs["one"], outs["two"]
    assert len(b["video"]) == 3 and all(np.array_equal(x, y) for x, y in zip(a["video"], b["video"]))
    assert a["png"] == b["png"] and a["meta"] == b["meta"]
