"""depth_midas band: CUDA path vs oracle/midas.py.

The oracle restates the published MiDaS v3 DPT graph (the hub code is not vendored: parity unpinned, see oracle/midas.py);
the test-size twin "dpt_tiny" runs the identical graph (patch 16, bilinear pos-embed resize, hooks, project readout,
RefineNet fusion, bicubic align_corners=True resize) in seconds on the CPU.  Tolerance 1e-3 (fp16 operands, fp32 accumulate).
"""
import numpy as np
import pytest
import torch

from oracle import midas as om
from prisma_b200.synthetic import synthetic_frame
from prisma_b200.seeded_weights import make_midas_weights


def test_midas_size_arithmetic_and_pos_embed():
    # hubconf default_transform ("upper_bound": fit inside 384 x 384, multiples of 32) on the reference's config-1 frame
    assert om.midas_get_size(640, 480) == (384, 288)
    assert om.midas_get_size(1280, 720) == (384, 224) and om.midas_get_size(1920, 1080) == (384, 224)
    assert om.midas_get_size(480, 640) == (288, 384) and om.midas_get_size(1000, 1000) == (384, 384)
    pos = torch.randn(1, 577, 8)
    assert torch.equal(om._resize_pos_embed(pos, 24, 24), pos)  # native grid: identity


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max()), float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def tiny():
    from prisma_b200.depth import MidasEngine
    sd = make_midas_weights("dpt_tiny", 0)
    eng = MidasEngine(sd, variant="dpt_tiny")
    yield eng, sd
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(240, 320), (360, 200)])
def test_midas_band_matches_oracle(tiny, hw):
    eng, sd = tiny
    img = synthetic_frame(hw[0], hw[1], 1)
    pred = eng.infer(img)
    wn, hn = om.midas_get_size(hw[1], hw[0])
    taps = {}
    x = torch.from_numpy(om.midas_preprocess(img)).unsqueeze(0)
    assert x.shape[-2:] == (hn, wn)
    net_in = eng.read_tap("net_input", (3, hn, wn))
    assert np.array_equal(net_in, x[0].numpy())  # cv2 f64 cubic resize + (x-0.5)/0.5, bit-exact
    with torch.no_grad():
        ref_net = om.midas_model(sd, x, "dpt_tiny", taps=taps)[0].numpy()
    tok = eng.read_tap("tokens", tuple(taps["tokens"].shape[1:]))
    m, l2 = rel(tok, taps["tokens"][0].numpy())
    assert m < 1e-3 and l2 < 1e-3, ("tokens", m, l2)      # patch embed (fp16 operands) + bilinear pos-embed resize
    net = eng.read_tap("net_depth", (hn, wn))
    m, l2 = rel(net, ref_net)
    assert m < 1e-3 and l2 < 1e-3, ("net_depth", m, l2)
    ref = om.midas_infer(sd, img, "dpt_tiny")
    m, l2 = rel(pred, ref)
    assert m < 1e-3 and l2 < 1e-3, ("prediction", m, l2)
    # bicubic(align_corners=True) resize alone, from the engine's own network output: fp32 op-order tolerance
    up = torch.nn.functional.interpolate(torch.from_numpy(net)[None, None], size=hw, mode="bicubic", align_corners=True)[0, 0].numpy()
    assert np.abs(up - pred).max() <= 2e-6 * np.abs(up).max()


@pytest.mark.gpu
def test_midas_video_encode_bit_exact(tiny):
    eng, sd = tiny
    frames = np.stack([synthetic_frame(240, 320, t) for t in range(3)])
    rgb, mins, maxs, pred = eng.infer_clip(frames, pass_frames=3, want_depth=True)
    for i in range(3):
        ref_rgb, dmin, dmax = om.midas_encode(pred[i])
        assert np.float32(dmin) == mins[i] and np.float32(dmax) == maxs[i]
        assert np.array_equal(rgb[i], ref_rgb)      # f32-hue variant of heat_to_rgb (depth_midas.py:144), bit-exact
        single = eng.infer(frames[i])
        assert np.array_equal(single, pred[i])


@pytest.mark.gpu
def test_midas_dpt_large_config1_frame_matches_oracle():
    """BASELINE configs[0]: one 640x480 frame through the full-size DPT_Large graph (384x512 input, 769 tokens)."""
    from prisma_b200.depth import MidasEngine
    sd = make_midas_weights("dpt_large", 0)
    eng = MidasEngine(sd)
    img = synthetic_frame(480, 640, 0)
    pred = eng.infer(img)
    ref = om.midas_infer(sd, img, "dpt_large")
    eng.close()
    m, l2 = rel(pred, ref)
    assert m < 1e-3 and l2 < 1e-3, (m, l2)
