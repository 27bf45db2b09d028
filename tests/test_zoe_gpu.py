"""depth_anything --metric: ZoeDepth metric head, CUDA path vs the reference fixture and the
pinned oracle.  Tolerance 1e-3 (fp16 operands, fp32 accumulate); the PIL resize and the encode are checked on their own."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import da as oda
from oracle import zoe as ozoe
from prisma_b200.synthetic import synthetic_frame
from prisma_b200.seeded_weights import make_zoe_weights


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max()), float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def zoe_s():
    from prisma_b200.depth import ZoeDepthEngine
    sd = make_zoe_weights("vits", 0)
    eng = ZoeDepthEngine(sd, encoder="vits")
    yield eng, sd
    eng.close()


@pytest.mark.gpu
def test_zoe_matches_reference_fixture(golden_dir, zoe_s):
    eng, sd = zoe_s
    g = np.load(os.path.join(golden_dir, "zoe_vits_240x320.npz"))
    rgb, dmin, dmax, pred = eng.infer_encoded(g["image"], want_depth=True)
    net = eng.read_tap("metric_net", (392, 518))
    m, l2 = rel(net, g["metric_net"])
    assert m < 1e-3 and l2 < 1e-3, ("metric_net", m, l2)
    m, l2 = rel(pred, g["prediction"])
    assert m < 1e-3 and l2 < 1e-3, ("prediction", m, l2)
    # input transform: ToTensor + bilinear(align_corners=True) + normalise in f32
    x01 = torch.from_numpy(np.ascontiguousarray(g["image"].transpose(2, 0, 1))).float().div(255).unsqueeze(0)
    xin = ((torch.nn.functional.interpolate(x01, (392, 518), mode="bilinear", align_corners=True) - ozoe.MEAN) / ozoe.STD)[0].numpy()
    assert np.abs(eng.read_tap("net_input", (3, 392, 518)) - xin).max() < 2e-6
    # PIL bicubic resize of the engine's own network output: double accumulation, f32 result
    up = np.asarray(Image.fromarray(net).resize((320, 240)))
    assert np.abs(up - pred).max() <= 1e-6 * np.abs(up).max()
    # encode (flip = False) of the engine's own prediction: bit-exact
    ref_rgb, rmin, rmax = oda.da_encode(pred, flip=False)
    assert np.array_equal(rgb, ref_rgb) and np.float32(rmin) == np.float32(dmin) and np.float32(rmax) == np.float32(dmax)


@pytest.mark.gpu
def test_zoe_720p_matches_oracle(zoe_s):
    eng, sd = zoe_s
    img = synthetic_frame(720, 1280, 2)
    pred = eng.infer(img)
    ref = ozoe.zoe_infer(sd, img, "vits")
    m, l2 = rel(pred, ref)
    assert m < 1e-3 and l2 < 1e-3, (m, l2)


@pytest.mark.gpu
def test_zoe_batch_equals_single_frames(zoe_s):
    """The metric head over a stack of frames (GEMMs over all frames' pixel rows) == frame-by-frame."""
    eng, sd = zoe_s
    frames = np.stack([synthetic_frame(240, 320, t) for t in range(5)])
    single = [eng.infer_encoded(f, want_depth=True) for f in frames]
    rgb, mins, maxs, pred = eng.infer_clip(frames, pass_frames=3, want_depth=True)   # 3 + ragged 2
    for i, (r1, mn, mx, p1) in enumerate(single):
        assert np.array_equal(pred[i], p1) and np.array_equal(rgb[i], r1), i
        assert np.float32(mn) == mins[i] and np.float32(mx) == maxs[i]


@pytest.mark.gpu
def test_zoe_vitl_matches_oracle():
    """The configuration the reference actually ships (ViT-L core, 256-channel decoder features) on a 720p frame."""
    from prisma_b200.depth import ZoeDepthEngine
    sd = make_zoe_weights("vitl", 0)
    eng = ZoeDepthEngine(sd, encoder="vitl")
    img = synthetic_frame(720, 1280, 1)
    pred = eng.infer(img)
    eng.close()
    ref = ozoe.zoe_infer(sd, img, "vitl")
    m, l2 = rel(pred, ref)
    assert m < 1e-3 and l2 < 1e-3, (m, l2)
