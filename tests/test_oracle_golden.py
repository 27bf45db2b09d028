"""CPU: the oracle (oracle/da.py) reproduces the fixtures that oracle/tools/make_golden.py recorded from the
*reference modules* in the authoring container -- this is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import da as oda
from prisma_b200.synthetic import synthetic_frame
from prisma_b200.seeded_weights import make_da_weights


def test_sizes_match_reference(golden_dir):
    rows = np.load(os.path.join(golden_dir, "da_sizes.npz"))["rows"]
    for w, h, rw, rh in rows:
        assert oda.da_get_size(int(w), int(h)) == (int(rw), int(rh))


@pytest.mark.parametrize("tag", ["vits_160x208"])
def test_oracle_reproduces_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"da_{tag}.npz"))
    H, W = [int(v) for v in g["frame_hw"]]
    enc = str(g["encoder"])
    img = synthetic_frame(H, W, int(g["frame_index"]))
    x = oda.da_preprocess(img)
    assert np.array_equal(x[:, ::4, ::4], g["net_input_s4"])
    sd = make_da_weights(enc, int(g["seed"]))
    taps = {}
    with torch.no_grad():
        depth = oda.da_model(sd, torch.from_numpy(x)[None], enc, taps=taps)
    # same torch build => bit-identical; allow 1e-5 so a different BLAS thread count cannot flake
    np.testing.assert_allclose(taps["feats"][3][0, ::7].numpy(), g["feat3_s7"], rtol=0, atol=1e-5 * np.abs(g["feat3_s7"]).max())
    np.testing.assert_allclose(depth[0, ::2, ::2].numpy(), g["depth_s2"], rtol=0, atol=1e-5 * g["depth_s2"].max())
    pred = oda.da_upsample(depth, H, W)
    np.testing.assert_allclose(pred, g["prediction"], rtol=0, atol=1e-5 * g["prediction"].max())
    # the encoder is pinned bit-exactly on the reference's own prediction
    rgb, dmin, dmax = oda.da_encode(g["prediction"])
    assert np.array_equal(rgb, g["rgb"])
    assert np.float32(dmin) == g["dmin"] and np.float32(dmax) == g["dmax"]


def test_heat_roundtrip():
    """heat_to_rgb <-> rgb_to_heat (common/encode.py:31-33,61-64; the 1/0.65 constant)."""
    h = np.linspace(0, 1, 1001).reshape(1, -1)
    rgb = oda.heat_to_rgb(h)
    mx, mn = rgb.max(-1), rgb.min(-1)
    # hue of a fully saturated colour
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    hue = np.where(mx == r, ((g - b) / (mx - mn + 1e-12)) % 6, np.where(mx == g, (b - r) / (mx - mn + 1e-12) + 2, (r - g) / (mx - mn + 1e-12) + 4)) / 6
    back = np.clip(1.0 - hue * 1.538461538, 0, 1)
    assert np.abs(back - h).max() < 1e-6


def test_png_variant_matches_reference(golden_dir):
    """oracle write_depth restatement vs the RGB array the reference's common.io.write_depth produced."""
    g = np.load(os.path.join(golden_dir, "da_vits_480x640.npz"))
    ref = np.load(os.path.join(golden_dir, "da_png_480x640.npz"))["rgb_png"]
    rgb, dmin, dmax = oda.da_write_depth_rgb(g["prediction"], True)
    assert np.array_equal(rgb, ref)
