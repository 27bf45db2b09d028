"""Size-independent properties of the band encodings, checked on the oracle with hypothesis (CPU): the 16-bit flow payload
and the packed depth range decode back to their inputs, a flow pair that is its own inverse passes the consistency test,
the heat / HSV encodes are bounded.  The GPU tests pin the kernels to the oracle; these pin the oracle's own invariants."""
import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra.numpy import arrays

from oracle import da as oda
from oracle import raft as oraft


@settings(max_examples=50, deadline=None)
@given(arrays(np.float32, (6, 9, 2), elements=st.floats(-120, 120, width=32)))
def test_flow_u16_payload_round_trips(flow):
    enc = oraft.encode_flow(flow.copy(), np.ones(flow.shape[:2], bool))
    dec = (enc[..., :2].astype(np.float64) - 2 ** 15) / 2 ** 8
    assert np.all(enc[..., 2] == 65535)                       # |flow| < 127.99 -> always valid
    # the reference evaluates 2^15 + 256 f in float32 (encode.py:106): 256 f is exact, the sum rounds to the f32 grid of
    # [2^14, 2^16) (ulp <= 2^-8 code units, i.e. an error <= 2^-9 / 256 = 2^-17 in flow units) and is then truncated
    eps = 2.0 ** -17
    assert np.all(dec <= flow.astype(np.float64) + eps) and np.all(flow.astype(np.float64) - dec < 1 / 256 + eps)


@settings(max_examples=50, deadline=None)
@given(st.floats(0.01, 900.0), st.floats(0.01, 90.0))
def test_depth_range_pixels_decode(dmin, span):
    pred = np.linspace(dmin, dmin + span, 64, dtype=np.float32).reshape(8, 8)
    rgb, mn, mx = oda.da_write_depth_rgb(pred, True)
    for px, val in ((rgb[0, 0], pred.min()), (rgb[0, 1], pred.max())):   # 24-bit value over [0, 1000] (encode.py:141-146)
        code = int(px[0]) + 256 * int(px[1]) + 65536 * int(px[2])
        assert abs(code / (256 ** 3 - 1) * 1000.0 - float(val)) <= 1000.0 / 2 ** 24 + 1e-3 * float(val) * 2 ** -10


def test_consistent_flow_pair_passes_the_mask_test():
    H, W = 48, 64
    fwd = np.zeros((H, W, 2), np.float32)
    fwd[..., 0], fwd[..., 1] = 3.0, -2.0          # a pure translation: the backward flow is its negation everywhere
    fm, bm = oraft.compute_fwdbwd_mask(fwd, -fwd)
    assert fm[4:-4, 4:-4].all() and bm[4:-4, 4:-4].all()      # interior consistent; the border samples zeros (remap constant 0)
    fm2, _ = oraft.compute_fwdbwd_mask(fwd, np.zeros_like(fwd))
    assert not fm2[4:-4, 4:-4].any()                           # |fwd + 0| = 3.6 > 0.05 * 3.6 + 0.5


@settings(max_examples=30, deadline=None)
@given(arrays(np.float32, (5, 7), elements=st.floats(0.0, 1.0, width=32)))
def test_heat_map_is_a_valid_colour(heat):
    rgb = oda.heat_to_rgb(heat.astype(np.float64))
    assert rgb.shape == (5, 7, 3) and rgb.min() >= 0.0 and rgb.max() <= 1.0
    assert np.all(rgb.max(axis=-1) == 1.0)                     # fully saturated hue: one channel at full scale
