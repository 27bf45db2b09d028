"""CPU: the RAFT oracle (oracle/raft.py) reproduces the fixture recorded from the reference RAFT / CorrBlock /
process_flow in the authoring container (oracle/tools/make_golden.py raft)."""
import os

import numpy as np
import torch

from oracle import raft as oraft
from prisma_b200.synthetic import synthetic_frame


def test_raft_oracle_stages(golden_dir):
    g = np.load(os.path.join(golden_dir, "raft_240x320.npz"))
    f0 = synthetic_frame(240, 320, 0)
    assert np.array_equal(oraft.raft_preprocess(f0).permute(1, 2, 0).numpy().astype(np.uint8), g["resized0"])
    assert oraft.input_pad(180, 240) == [int(v) for v in g["pad"]]
    fm1 = torch.from_numpy(g["fmap1"].astype(np.float32))
    fm2 = torch.from_numpy(g["fmap2"].astype(np.float32))
    pyr = oraft.corr_pyramid(fm1, fm2)
    # fixture feature maps are stored as fp16 -> compare at fp16-feature accuracy
    assert np.abs(pyr[0][:64, 0].numpy() - g["corr_l0_rows"]).max() <= 2e-3 * np.abs(g["corr_l0_rows"]).max()
    assert np.abs(pyr[3][:64, 0].numpy() - g["corr_l3_rows"]).max() <= 2e-3 * np.abs(g["corr_l0_rows"]).max()
    rgb, md = oraft.process_flow(g["flow_fwd"])
    assert np.array_equal(rgb, g["flow_rgb"]) and np.float32(md) == g["flow_max"]


def test_input_pad_arithmetic():
    """InputPadder._pad known answers."""
    assert oraft.input_pad(810, 1440) == [0, 0, 3, 3]
    assert oraft.input_pad(540, 960) == [0, 0, 2, 2]
    assert oraft.input_pad(816, 1440) == [0, 0, 0, 0]
