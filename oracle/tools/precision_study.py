"""Which GEMM operand type keeps the DA output within 1e-3 of the fp32 CPU path?
Emulates operand rounding through the oracle's `q` hook (TEST INFRASTRUCTURE).
    python oracle/tools/precision_study.py vits|vitl
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import da as oda
from prisma_b200.seeded_weights import make_da_weights
from prisma_b200.synthetic import synthetic_frame

enc = sys.argv[1] if len(sys.argv) > 1 else "vits"
H, W = (480, 640) if enc == "vits" else (720, 1280)
torch.set_grad_enabled(False)
sd = make_da_weights(enc, 0)
x = torch.from_numpy(oda.da_preprocess(synthetic_frame(H, W, 0))).unsqueeze(0)
t = time.time(); ref = oda.da_model(sd, x, enc).numpy(); print("fp32 %.1fs" % (time.time() - t), ref.min(), ref.max(), ref.mean())
def tf32(t):
    return (t.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if t.dtype == torch.float32 else t
for name, q in [("fp16", lambda t: t.half().float()), ("bf16", lambda t: t.bfloat16().float())]:
    out = oda.da_model(sd, x, enc, q=q).numpy()
    d = np.abs(out - ref)
    print(f"{name}: max|d|/max|ref| = {d.max()/np.abs(ref).max():.3e}  rel-L2 = {np.linalg.norm(d)/np.linalg.norm(ref):.3e}  "
          f"max|d|/range = {d.max()/(ref.max()-ref.min()):.3e}  mean|d|/mean|ref| = {d.mean()/np.abs(ref).mean():.3e}")

# ---- where does the error come from?  (encoder only / head only)
h16 = lambda t: t.half().float()
def run(qe, qh):
    feats = oda.vit_features(sd, x, enc, qe)
    d = oda.dpt_head(sd, feats, x.shape[-2] // 14, x.shape[-1] // 14, qh)
    return torch.relu(d).squeeze(1).numpy(), feats
for name, qe, qh in [("enc-fp16/head-fp32", h16, oda._ident), ("enc-fp32/head-fp16", oda._ident, h16)]:
    out, feats = run(qe, qh)
    d = np.abs(out - ref)
    print(f"{name}: max|d|/range = {d.max()/(ref.max()-ref.min()):.3e} rel-L2 = {np.linalg.norm(d)/np.linalg.norm(ref):.3e}")
f32 = oda.vit_features(sd, x, enc)
f16 = oda.vit_features(sd, x, enc, h16)
for a, b in zip(f32, f16):
    print("feat rel-L2 %.3e  max %.3e" % (float((a-b).norm()/a.norm()), float((a-b).abs().max()/a.abs().max())))
