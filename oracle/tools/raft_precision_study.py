"""Where does fp16-operand error enter RAFT's 12-iteration flow?  Emulates operand rounding per stage in the oracle
(TEST INFRASTRUCTURE).    python oracle/tools/raft_precision_study.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import torch.nn.functional as TF
from oracle import raft as oraft
from prisma_b200.seeded_weights import make_raft_weights
from prisma_b200.synthetic import synthetic_frame

torch.set_grad_enabled(False)
sd = make_raft_weights(0)
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (240, 320)
a = oraft.raft_preprocess(synthetic_frame(H, W, 0))[None]; b = oraft.raft_preprocess(synthetic_frame(H, W, 1))[None]
i1, i2 = torch.cat([a, b]), torch.cat([b, a])
pad = oraft.input_pad(*i1.shape[-2:])
p1 = TF.pad(i1, pad, mode="replicate"); p2 = TF.pad(i2, pad, mode="replicate")
h16 = lambda t: t.half().float()
STAGE = {"name": None}
ACTIVE = set()

class FProxy:
    def __getattr__(self, k):
        return getattr(TF, k)
    def conv2d(self, x, w, b=None, **kw):
        if STAGE["name"] in ACTIVE:
            return TF.conv2d(h16(x), h16(w), b, **kw)
        return TF.conv2d(x, w, b, **kw)
oraft.F = FProxy()
_enc = oraft.basic_encoder
def enc(x, sd_, p, kind):
    STAGE["name"] = p; r = _enc(x, sd_, p, kind); STAGE["name"] = None; return r
oraft.basic_encoder = enc
_ub = oraft.update_block
def ub(*a, **k):
    STAGE["name"] = "update"; r = _ub(*a, **k); STAGE["name"] = None; return r
oraft.update_block = ub
_cp = oraft.corr_pyramid
def cp(f1, f2, levels=4):
    if "corr" in ACTIVE: f1, f2 = h16(f1), h16(f2)
    return _cp(f1, f2, levels)
oraft.corr_pyramid = cp
_lk = oraft.corr_lookup
def lk(pyr, coords, r=4):
    o = _lk(pyr, coords, r)
    return h16(o) if "lookup" in ACTIVE else o
oraft.corr_lookup = lk

ref_lo, ref_up = oraft.raft_forward(sd, p1, p2, 12)
mx = float(ref_up.abs().max())
for name, act in [("fnet", {"fnet."}), ("cnet", {"cnet."}), ("corr operands", {"corr"}), ("lookup out", {"lookup"}),
                  ("update block", {"update"}), ("all", {"fnet.", "cnet.", "corr", "lookup", "update"}),
                  ("all but fnet", {"cnet.", "corr", "lookup", "update"}), ("all but update", {"fnet.", "cnet.", "corr", "lookup"})]:
    ACTIVE.clear(); ACTIVE.update(act)
    lo, up = oraft.raft_forward(sd, p1, p2, 12)
    d = (up - ref_up).abs()
    print(f"{name:16s} max|d|/max = {float(d.max())/mx:.3e}   rel-L2 = {float((up-ref_up).norm()/ref_up.norm()):.3e}")

print("---- stage errors under all-fp16 operand emulation (compare with tests/test_raft_gpu.py's report)")
ACTIVE.clear()
t_ref = {}; oraft.raft_forward(sd, p1, p2, 12, taps=t_ref)
ACTIVE.update({"fnet.", "cnet.", "corr", "lookup", "update"})
t_q = {}; oraft.raft_forward(sd, p1, p2, 12, taps=t_q)
for k in ("fmap1", "net0", "inp", "lookup0", "net1", "delta0"):
    a, b = t_q[k], t_ref[k]
    print(f"{k:8s} max-rel {float((a-b).abs().max()/b.abs().max()):.3e}  rel-L2 {float((a-b).norm()/b.norm()):.3e}  max|ref| {float(b.abs().max()):.3f}")
