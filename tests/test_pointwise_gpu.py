"""GPU: HBM-bound kernels (LayerNorm, OpenCV-exact pre-process, heat encode) through the C ABI vs the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import da as oda
from prisma_b200.synthetic import synthetic_frame
from prisma_b200._lib import check, fptr, lib, u8ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D", [(2443, 1024), (1813, 384), (7, 768)])
def test_layernorm(rows, D):
    rng = np.random.default_rng(rows)
    x = rng.standard_normal((rows, D), dtype=np.float32) * 3 + 0.5
    g = 1 + 0.1 * rng.standard_normal(D, dtype=np.float32)
    b = 0.05 * rng.standard_normal(D, dtype=np.float32)
    y = np.empty((rows, D), np.float32)
    check(lib().prisma_debug_layernorm(0, fptr(x), fptr(g), fptr(b), fptr(y), rows, D))
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (D,), torch.from_numpy(g), torch.from_numpy(b), eps=1e-6).numpy()
    assert np.abs(y - ref).max() <= 1.5e-3 * np.abs(ref).max()  # fp16 output rounding


@pytest.mark.parametrize("H,W", [(480, 640), (720, 1280), (1080, 1920), (123, 321), (518, 518)])
def test_da_preprocess_matches_opencv_path(H, W):
    """K1 vs the reference transform (cv2 INTER_CUBIC on f64, normalise, f32): equal to the last f32 bit
    except where an FMA-free f64 sum lands within half an f32 ulp of a tie."""
    img = synthetic_frame(H, W, 3)
    ref = oda.da_preprocess(img)
    wn, hn = oda.da_get_size(W, H)
    out = np.empty((3, hn, wn), np.float32)
    check(lib().prisma_debug_da_preprocess(0, u8ptr(img), H, W, fptr(out), hn, wn))
    diff = np.abs(out - ref)
    assert diff.max() <= 5e-7 * max(1.0, np.abs(ref).max()), diff.max()
    assert (out != ref).mean() < 1e-3


@pytest.mark.parametrize("tag", ["vits_160x208", "vits_480x640"])
def test_depth_encode_bit_exact(golden_dir, tag):
    """K10 encode on the reference's own prediction: u8 frame and (min,max) must be bit-exact."""
    from prisma_b200.depth import DepthAnythingEngine
    g = np.load(os.path.join(golden_dir, f"da_{tag}.npz"))
    eng = DepthAnythingEngine("vits")
    rgb, dmin, dmax = eng.encode(g["prediction"], flip=True)
    assert np.float32(dmin) == g["dmin"] and np.float32(dmax) == g["dmax"]
    assert np.array_equal(rgb, g["rgb"])
    # also a large synthetic field (full 1080p size) against the oracle encoder
    rng = np.random.default_rng(0)
    p = rng.random((1080, 1920), dtype=np.float32) * 7 + 0.3
    rgb2, mn, mx = eng.encode(p)
    ref, rmin, rmax = oda.da_encode(p)
    assert np.array_equal(rgb2, ref) and np.float32(mn) == np.float32(rmin) and np.float32(mx) == np.float32(rmax)
    eng.close()


def test_depth_png_encode_bit_exact(golden_dir):
    """PNG variant (write_depth: Sobel-edge saturation + range pixels) on the reference's own prediction vs the RGB array
    the reference wrote (tests/golden/da_png_480x640.npz, recorded from common.io.write_depth) -- bit-exact."""
    from prisma_b200.depth import DepthAnythingEngine
    g = np.load(os.path.join(golden_dir, "da_vits_480x640.npz"))
    ref = np.load(os.path.join(golden_dir, "da_png_480x640.npz"))["rgb_png"]
    eng = DepthAnythingEngine("vits")
    rgb, dmin, dmax = eng.encode_png(g["prediction"], flip=True)
    assert np.array_equal(rgb, ref)
    assert np.float32(dmin) == g["dmin"] and np.float32(dmax) == g["dmax"]
    rng = np.random.default_rng(3)
    p = (rng.random((720, 1280), dtype=np.float32) * 30 + 2).astype(np.float32)
    rgb2, _, _ = eng.encode_png(p)
    ref2, _, _ = oda.da_write_depth_rgb(p, True)
    assert np.array_equal(rgb2, ref2)
    eng.close()
