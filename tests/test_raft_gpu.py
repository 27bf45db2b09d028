"""GPU: the whole RAFT band path (prisma_flow_infer through the C ABI) vs the CPU oracle (itself bit-equal to the
reference RAFT, tests/test_oracle_raft_golden.py), stage by stage and end to end."""
import os

import numpy as np
import pytest
import torch

from oracle import raft as oraft
from prisma_b200.synthetic import synthetic_frame
from prisma_b200.seeded_weights import make_raft_weights

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max()), float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def raft_engine():
    from prisma_b200.flow import RaftFlowEngine
    eng = RaftFlowEngine(make_raft_weights(0), iterations=12, scale=0.75)
    yield eng
    eng.close()


def _oracle(f0, f1, rs0, rs1, iters):
    """Oracle on OUR resized frames (the u8 cubic resize differs from cv2 by 1 LSB on <0.1% of pixels, tested in
    test_flow_gpu.py; feeding the same resized pixels isolates the model arithmetic)."""
    sd = make_raft_weights(0)
    a = torch.from_numpy(rs0).permute(2, 0, 1).float()[None]
    b = torch.from_numpy(rs1).permute(2, 0, 1).float()[None]
    i1, i2 = torch.cat([a, b]), torch.cat([b, a])
    pad = oraft.input_pad(*i1.shape[-2:])
    p1 = torch.nn.functional.pad(i1, pad, mode="replicate")
    p2 = torch.nn.functional.pad(i2, pad, mode="replicate")
    taps = {}
    with torch.no_grad():
        lo, up = oraft.raft_forward(sd, p1, p2, iters, taps=taps)
    H, W = up.shape[-2:]
    up = up[..., pad[2]:H - pad[3], pad[0]:W - pad[1]]
    return taps, lo, up


def test_raft_stages_and_flow(raft_engine):
    H, W = 240, 320
    f0, f1 = synthetic_frame(H, W, 0), synthetic_frame(H, W, 1)
    out = raft_engine.infer_pair(f0, f1, want_rgb=True)
    hs, ws = raft_engine.out_size(H, W)
    rs = raft_engine.read_tap("resized", (2, hs, ws, 3)).astype(np.uint8)
    taps, lo, up = _oracle(f0, f1, rs[0], rs[1], 12)
    h8, w8 = taps["fmap1"].shape[-2:]
    P = h8 * w8
    rep = {}
    fm = raft_engine.read_tap("fmap", (2, P, 256))
    ref_fm = taps["fmap1"].permute(0, 2, 3, 1).reshape(2, P, 256).numpy()
    rep["fmap"] = _rel(fm, ref_fm)
    cn = raft_engine.read_tap("cnet_out", (2, P, 256))
    ref_net = taps["net0"].permute(0, 2, 3, 1).reshape(2, P, 128).numpy()
    ref_inp = taps["inp"].permute(0, 2, 3, 1).reshape(2, P, 128).numpy()
    rep["cnet_net"] = _rel(np.tanh(cn[..., :128]), ref_net)
    rep["cnet_inp"] = _rel(np.maximum(cn[..., 128:], 0), ref_inp)
    c1 = raft_engine.read_tap("coords1_iter0", (2, 2, P))
    ref_c1 = (oraft.coords_grid(2, h8, w8) + taps["delta0"]).reshape(2, 2, P).numpy()
    rep["delta_iter0"] = (float(np.abs(c1 - ref_c1).max()), float(np.abs(taps["delta0"].numpy()).max()))
    fwd_ref = up[0].permute(1, 2, 0).numpy()
    bwd_ref = up[1].permute(1, 2, 0).numpy()
    rep["flow_fwd"] = _rel(out["fwd"], fwd_ref)
    rep["flow_bwd"] = _rel(out["bwd"], bwd_ref)
    rep["max_fwd"] = (out["max_fwd"], float(np.sqrt((fwd_ref ** 2).sum(-1)).max()))
    print(rep, "ms", out["ms"])
    # stage diagnostics (max element error / max |ref| after nine fp16-operand conv layers; they move by a few percent with the
    # summation order inside the tensor core, e.g. when the stem's K layout changed) -- the binding tolerance is the flow's
    assert rep["fmap"][0] <= 6e-3 and rep["fmap"][1] <= 2e-3
    assert rep["cnet_net"][0] <= 6e-3 and rep["cnet_inp"][0] <= 6e-3
    assert rep["delta_iter0"][0] <= 2e-2 * max(1.0, rep["delta_iter0"][1])
    # north_star tolerance on the flow floats: 1e-3 relative (to the largest displacement)
    assert rep["flow_fwd"][0] <= 1e-3 and rep["flow_bwd"][0] <= 1e-3, rep


def test_raft_720p_matches_oracle(raft_engine):
    """BASELINE frame size 720p (x0.75 -> 540x960, padded 544x960): CUDA vs the CPU oracle on the same resized frames."""
    H, W = 720, 1280
    f0, f1 = synthetic_frame(H, W, 0), synthetic_frame(H, W, 1)
    out = raft_engine.infer_pair(f0, f1)
    hs, ws = raft_engine.out_size(H, W)
    rs = raft_engine.read_tap("resized", (2, hs, ws, 3)).astype(np.uint8)
    _, lo, up = _oracle(f0, f1, rs[0], rs[1], 12)
    fwd_ref = up[0].permute(1, 2, 0).numpy()
    bwd_ref = up[1].permute(1, 2, 0).numpy()
    r = (_rel(out["fwd"], fwd_ref), _rel(out["bwd"], bwd_ref))
    print("720p raft: fwd", r[0], "bwd", r[1], "max|flow|", float(np.abs(fwd_ref).max()), "ms", out["ms"])
    assert r[0][0] <= 1e-3 and r[1][0] <= 1e-3


def test_raft_1080p_matches_oracle(raft_engine):
    """BASELINE configs[2] itself: 1080p x0.75 -> 810x1440 (padded 816x1440), 12 iterations, forward AND backward flow,
    CUDA vs the CPU oracle (bit-equal to the reference RAFT) on the same resized frames.  Tolerance = north_star's 1e-3
    relative to the largest displacement.  The oracle needs ~30-60 s of host CPU at this size."""
    H, W = 1080, 1920
    f0, f1 = synthetic_frame(H, W, 0), synthetic_frame(H, W, 1)
    out = raft_engine.infer_pair(f0, f1)
    hs, ws = raft_engine.out_size(H, W)
    assert (hs, ws) == (810, 1440)
    rs = raft_engine.read_tap("resized", (2, hs, ws, 3)).astype(np.uint8)
    _, lo, up = _oracle(f0, f1, rs[0], rs[1], 12)
    fwd_ref = up[0].permute(1, 2, 0).numpy()
    bwd_ref = up[1].permute(1, 2, 0).numpy()
    r = (_rel(out["fwd"], fwd_ref), _rel(out["bwd"], bwd_ref))
    print("1080p raft: fwd", r[0], "bwd", r[1], "max|flow|", float(np.abs(fwd_ref).max()), "ms", out["ms"])
    assert r[0][0] <= 1e-3 and r[1][0] <= 1e-3, r
    assert abs(out["max_fwd"] - float(np.sqrt((fwd_ref ** 2).sum(-1)).max())) <= 1e-3 * float(np.abs(fwd_ref).max())


@pytest.mark.parametrize("pairs_per_pass", [4, 2, 1])
def test_raft_streamed_clip_equals_pairwise_calls(pairs_per_pass):
    """prisma_flow_infer_stream over a clip (new clip, then a continued chunk; pinned and pageable buffers) == the
    per-pair calls of the band's loop, bit for bit (flows, HSV frames, max displacements) -- with one pair per pass and with
    two / four (n + 1 frames -> 2 n directions per pass; the chunks end on partly filled passes)."""
    from prisma_b200.depth import pinned_empty
    from prisma_b200.flow import RaftFlowEngine
    eng = RaftFlowEngine(make_raft_weights(0), iterations=3, scale=0.75)
    assert eng.pairs_per_pass == 4   # the default of the clip path
    eng.pairs_per_pass = pairs_per_pass
    assert eng.pairs_per_pass == pairs_per_pass
    frames = np.stack([synthetic_frame(240, 320, t) for t in range(6)])
    ref = [eng.infer_pair(frames[i], frames[i + 1], want_rgb=True) for i in range(5)]
    pin = pinned_empty(frames.shape, np.uint8)
    pin[...] = frames
    a = eng.infer_clip(pin[:4])                          # new clip: pairs (0,1) (1,2) (2,3)
    b = eng.infer_clip(frames[4:], continue_clip=True)   # continues: pairs (3,4) (4,5)
    assert a["pairs"] == 3 and b["pairs"] == 2
    got = [(a, j) for j in range(3)] + [(b, j) for j in range(2)]
    for i, (o, j) in enumerate(got):
        for k in ("fwd", "bwd", "fwd_rgb", "bwd_rgb"):
            assert np.array_equal(o[k][j], ref[i][k]), (i, k)
        assert o["max_fwd"][j] == np.float32(ref[i]["max_fwd"]) and o["max_bwd"][j] == np.float32(ref[i]["max_bwd"])
    c = eng.infer_clip(frames[:1])                       # a lone frame of a new clip: nothing to pair
    assert c["pairs"] == 0
    eng.close()


def test_raft_1080p_properties(raft_engine):
    """BASELINE config-3 size (1080p x0.75, 12 iterations): size-independent properties.
    (a) swapping the two frames swaps forward and backward flow bit for bit (the two directions are independent
        batch entries of the same pass); (b) the pass is deterministic; (c) the HSV frame / max displacement agree
        with the oracle's process_flow applied to OUR flow."""
    H, W = 1080, 1920
    f0, f1 = synthetic_frame(H, W, 0), synthetic_frame(H, W, 1)
    a = raft_engine.infer_pair(f0, f1, want_rgb=True)
    b = raft_engine.infer_pair(f1, f0)
    assert a["fwd"].shape == (810, 1440, 2)
    assert np.array_equal(a["fwd"], b["bwd"]) and np.array_equal(a["bwd"], b["fwd"])
    c = raft_engine.infer_pair(f0, f1)
    assert np.array_equal(a["fwd"], c["fwd"])
    ref_rgb, ref_max = oraft.process_flow(a["fwd"])
    assert np.float32(a["max_fwd"]) == np.float32(ref_max)
    d = np.abs(a["fwd_rgb"].astype(int) - ref_rgb.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    w = raft_engine.work(H, W)
    print("1080p raft pass: %.2f ms  (%.1f pairs/s, %.0f GFLOP algorithmic -> %.0f TFLOP/s), %d launches" %
          (c["ms"], 1e3 / c["ms"], w["flop"] / 1e9, w["flop"] / c["ms"] / 1e9, w["launches"]))


@pytest.mark.gpu
def test_video_pass_reuses_features_bit_identically():
    """prisma_flow_infer_video(reuse_prev): pair (f1, f2) right after pair (f0, f1), with f1's encoder features taken from the
    engine instead of being recomputed, equals a fresh full pass over (f1, f2) bit for bit; reuse_prev means "prev is the last
    call's curr" -- after (f1, f2) it makes the engine pair f2 (not the f0 handed in) with the new frame."""
    from prisma_b200.flow import RaftFlowEngine
    from prisma_b200.synthetic import synthetic_frame
    from prisma_b200.seeded_weights import make_raft_weights
    eng = RaftFlowEngine(make_raft_weights(0), iterations=4, scale=0.75)
    f = [synthetic_frame(240, 320, t) for t in range(3)]
    fresh = eng.infer_pair(f[1], f[2], want_rgb=True)                       # full pass; the cache now holds f2
    misused = eng.infer_pair(f[0], f[1], want_rgb=True, reuse_prev=True)   # -> computes the pair (f2, f1)
    a = eng.infer_pair(f[0], f[1], want_rgb=True)                          # full pass (f0, f1); the cache now holds f1
    b = eng.infer_pair(f[1], f[2], want_rgb=True, reuse_prev=True)         # video pass: only f2 is encoded
    for k in ("fwd", "bwd", "fwd_rgb", "bwd_rgb"):
        assert np.array_equal(b[k], fresh[k]), k
    assert b["max_fwd"] == fresh["max_fwd"] and b["max_bwd"] == fresh["max_bwd"]
    assert not np.array_equal(misused["fwd"], a["fwd"])
    eng.close()
