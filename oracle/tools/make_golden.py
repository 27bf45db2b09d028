"""Pin the oracle against the reference and write the golden fixtures (TEST INFRASTRUCTURE).

Runs ONLY in the authoring container (needs /root/reference).  It
  1. imports the reference modules read-only (PYTHONPATH=/root/reference/bands, cwd
     /root/reference because d_anything/dpt.py:147 uses a relative torch.hub path),
  2. loads oracle.weights' seeded state_dict into them (strict=True -> names/shapes pinned),
  3. checks every oracle stage against the reference module's output on seeded inputs,
  4. writes small fixtures to tests/golden/*.npz (inputs + reference outputs) that the
     `-m "not gpu"` tests replay against the oracle and the `-m gpu` tests against CUDA.

    python oracle/tools/make_golden.py            # from /root/repo
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REF, "bands"))
os.chdir(REF)

import numpy as np
import torch

from oracle import da as oda
from oracle.weights import make_da_weights, DA_CONFIGS
from oracle.frames import synthetic_frame

GOLD = os.path.join(REPO, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.set_grad_enabled(False)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def golden_da(encoder, H, W, tag):
    from d_anything.dpt import DPT_DINOv2
    from d_anything.util.transform import Resize, NormalizeImage, PrepareForNet
    from torchvision.transforms import Compose
    import cv2
    import common.encode as renc

    c = DA_CONFIGS[encoder]
    sd = make_da_weights(encoder, seed=0)
    ref = DPT_DINOv2(encoder, c["features"], c["out_channels"]).eval()
    missing = ref.load_state_dict(sd, strict=True)
    print(f"[{tag}] load_state_dict strict ok: {missing}")

    img = synthetic_frame(H, W, 0)
    # --- reference pipeline, exactly depth_anything.py:63-75,122-133,215-220
    transform = Compose([
        Resize(width=518, height=518, resize_target=False, keep_aspect_ratio=True, ensure_multiple_of=14,
               resize_method='lower_bound', image_interpolation_method=cv2.INTER_CUBIC),
        NormalizeImage(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
        PrepareForNet(),
    ])
    image = img / 255.0
    x_ref = transform({'image': image})['image']
    x_t = torch.from_numpy(x_ref).unsqueeze(0)
    feats_ref = ref.pretrained.get_intermediate_layers(x_t, 4, return_class_token=True)
    depth_ref = ref(x_t)
    pred_ref = torch.nn.functional.interpolate(depth_ref[None], (H, W), mode='bilinear', align_corners=False)[0, 0].numpy()
    dmin, dmax = pred_ref.min(), pred_ref.max()
    dn = 1.0 - (pred_ref - dmin) / (dmax - dmin)
    rgb_ref = (renc.heat_to_rgb(dn.astype(np.float64)) * 255).astype(np.uint8)

    # --- oracle
    x_or = oda.da_preprocess(img)
    assert x_or.shape == x_ref.shape, (x_or.shape, x_ref.shape)
    print(f"[{tag}] preprocess max abs diff {np.abs(x_or - x_ref).max():.3e}")
    assert np.array_equal(x_or, x_ref)
    taps = {}
    depth_or = oda.da_model(sd, torch.from_numpy(x_or).unsqueeze(0), encoder, taps=taps)
    for i in range(4):
        e = rel(taps["feats"][i].numpy(), feats_ref[i][0].numpy())
        print(f"[{tag}] feats[{i}] rel err {e:.3e}")
        assert e < 1e-5
    e = rel(depth_or.numpy(), depth_ref.numpy())
    print(f"[{tag}] depth rel err {e:.3e}  range [{depth_ref.min():.4f},{depth_ref.max():.4f}]")
    assert e < 1e-5
    pred_or = oda.da_upsample(depth_or, H, W)
    rgb_or, omin, omax = oda.da_encode(pred_or)
    print(f"[{tag}] pred rel err {rel(pred_or, pred_ref):.3e}; rgb max diff "
          f"{np.abs(rgb_or.astype(int) - rgb_ref.astype(int)).max()}")
    # encode pinned on the *reference* prediction so it is independent of model rounding
    rgb_or2, _, _ = oda.da_encode(pred_ref)
    assert np.array_equal(rgb_or2, rgb_ref)

    np.savez_compressed(
        os.path.join(GOLD, f"da_{tag}.npz"),
        encoder=encoder, seed=0, frame_hw=np.array([H, W]), frame_index=0,
        net_input_s4=x_ref[:, ::4, ::4].astype(np.float32),         # every 4th pixel of the net input
        feat3_s7=feats_ref[3][0].numpy()[0, ::7].astype(np.float32),  # every 7th token of the last tap
        depth_s2=depth_ref.numpy()[0, ::2, ::2].astype(np.float32),  # every 2nd pixel of the net depth
        prediction=pred_ref.astype(np.float32),
        rgb=rgb_ref, dmin=np.float32(dmin), dmax=np.float32(dmax),
    )
    print(f"[{tag}] wrote fixture")


def golden_sizes():
    from d_anything.util.transform import Resize
    import cv2
    r = Resize(width=518, height=518, resize_target=False, keep_aspect_ratio=True, ensure_multiple_of=14,
               resize_method='lower_bound', image_interpolation_method=cv2.INTER_CUBIC)
    rows = []
    for (w, h) in [(1920, 1080), (1280, 720), (640, 480), (934, 440), (518, 518), (500, 300), (300, 500),
                   (1000, 1000), (4096, 2160), (321, 123)]:
        rw, rh = r.get_size(w, h)
        assert (int(rw), int(rh)) == oda.da_get_size(w, h), (w, h)
        rows.append((w, h, int(rw), int(rh)))
    np.savez(os.path.join(GOLD, "da_sizes.npz"), rows=np.array(rows))
    print("sizes ok", rows)


def golden_raft():
    """RAFT: reference model vs oracle on a 240x320 pair (fwd+bwd), fixtures for the correlation / lookup / encode stages."""
    import argparse
    from raft.raft import RAFT
    from common import encode as renc
    from common.flow import InputPadder
    from oracle import raft as oraft
    from oracle.weights import make_raft_weights
    sd = make_raft_weights(0)
    m = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval()
    print("[raft] load_state_dict:", m.load_state_dict(sd, strict=True))
    f0, f1 = synthetic_frame(240, 320, 0), synthetic_frame(240, 320, 1)
    import cv2
    a = torch.from_numpy(np.array(cv2.resize(f0, None, fx=0.75, fy=0.75, interpolation=cv2.INTER_CUBIC))).permute(2, 0, 1).float()[None]
    b = torch.from_numpy(np.array(cv2.resize(f1, None, fx=0.75, fy=0.75, interpolation=cv2.INTER_CUBIC))).permute(2, 0, 1).float()[None]
    assert torch.equal(a[0], oraft.raft_preprocess(f0)) and torch.equal(b[0], oraft.raft_preprocess(f1))
    i1, i2 = torch.cat([a, b]), torch.cat([b, a])
    padder = InputPadder(i1.shape)
    assert padder._pad == oraft.input_pad(*i1.shape[-2:])
    p1, p2 = padder.pad(i1, i2)
    lo_r, up_r = m(p1, p2, iters=12, test_mode=True)
    taps = {}
    lo_o, up_o = oraft.raft_forward(sd, p1, p2, 12, taps=taps)
    print("[raft] flow_low err %.3e flow_up err %.3e" % (float((lo_r - lo_o).abs().max()), float((up_r - up_o).abs().max())))
    assert torch.equal(lo_r, lo_o) and torch.equal(up_r, up_o)
    fwd_r = padder.unpad(up_r[0]).permute(1, 2, 0).numpy()
    fwd_o, bwd_o = oraft.raft_infer(sd, i1, i2, 12)
    assert np.array_equal(fwd_r, fwd_o)
    bwd_r = padder.unpad(up_r[1]).permute(1, 2, 0).numpy()
    from common.flow import compute_fwdbwd_mask as ref_mask
    fm_r, bm_r = ref_mask(fwd_r, bwd_r)
    fm_o, bm_o = oraft.compute_fwdbwd_mask(fwd_r, bwd_r)
    assert np.array_equal(fm_r, fm_o) and np.array_equal(bm_r, bm_o)
    enc_r = renc.encode_flow(fwd_r.copy(), fm_r.copy())
    assert np.array_equal(enc_r, oraft.encode_flow(fwd_r.copy(), fm_r.copy()))
    rgb_r, md_r = renc.process_flow(fwd_r)
    rgb_o, md_o = oraft.process_flow(fwd_o)
    assert np.array_equal(rgb_r, rgb_o) and md_r == md_o
    # reference CorrBlock on the oracle's feature maps: pins corr_pyramid / corr_lookup separately
    from raft.corr import CorrBlock
    cb = CorrBlock(taps["fmap1"], taps["fmap2"], radius=4)
    for l in range(4):
        assert torch.equal(cb.corr_pyramid[l], oraft.corr_pyramid(taps["fmap1"], taps["fmap2"])[l])
    g = torch.Generator().manual_seed(0)
    coords = oraft.coords_grid(2, 23, 30) + 6 * torch.randn(2, 2, 23, 30, generator=g)
    lk_r = cb(coords)
    lk_o = oraft.corr_lookup(oraft.corr_pyramid(taps["fmap1"], taps["fmap2"]), coords)
    assert torch.equal(lk_r, lk_o)
    np.savez_compressed(
        os.path.join(GOLD, "raft_240x320.npz"),
        seed=0, iters=12, frame_hw=np.array([240, 320]), scale=0.75, pad=np.array(padder._pad),
        resized0=a[0].permute(1, 2, 0).numpy().astype(np.uint8),
        fmap1=taps["fmap1"].numpy().astype(np.float16), fmap2=taps["fmap2"].numpy().astype(np.float16),
        corr_l0_rows=cb.corr_pyramid[0][:64, 0].numpy().astype(np.float32),        # first 64 rows of image 0
        corr_l3_rows=cb.corr_pyramid[3][:64, 0].numpy().astype(np.float32),
        coords=coords.numpy().astype(np.float32), lookup=lk_r.numpy().astype(np.float16),
        flow_fwd=fwd_r.astype(np.float32), flow_rgb=rgb_r, flow_max=np.float32(md_r),
        flow_bwd=bwd_r.astype(np.float32), fwd_mask=fm_r, bwd_mask=bm_r, flow_u16=enc_r,
    )
    print("[raft] wrote fixture; max|flow| %.2f" % float(np.abs(fwd_r).max()))


def golden_png():
    """process_image path: the reference's common.io.write_depth on the reference prediction of da_vits_480x640."""
    import types
    import cv2
    for m in ("av", "plyfile", "decord"):  # container / point-cloud deps of common.io that are not installed
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["plyfile"].PlyData = object
    sys.modules["plyfile"].PlyElement = object
    import common.io as rio
    g = np.load(os.path.join(GOLD, "da_vits_480x640.npz"))
    out = "/tmp/prisma_write_depth.png"
    rio.write_depth(out, g["prediction"], normalize=True, heatmap=True, encode_range=True, flip=True)
    ref = cv2.cvtColor(cv2.imread(out), cv2.COLOR_BGR2RGB)
    mine, _, _ = oda.da_write_depth_rgb(g["prediction"], True)
    assert np.array_equal(ref, mine)
    np.savez_compressed(os.path.join(GOLD, "da_png_480x640.npz"), rgb_png=ref)
    print("[png] oracle == reference write_depth; wrote fixture")


def golden_zoe(encoder="vits", H=240, W=320):
    """Metric path of the depth_anything band: ZoeDepth(DepthAnythingCore(DPT_DINOv2)) built exactly as ZoeDepth.build
    does minus the checkpoint loads (zoedepth_v1.py:249-260, base_models/depth_anything.py:330-349), seeded weights
    loaded strict, band pre/post of bands/depth_anything.py:106-119 and the video-loop encode (:215-220, flip=False)."""
    import json
    import types
    from PIL import Image
    from torchvision import transforms
    sys.modules["json5"] = types.ModuleType("json5")  # utils/config.py:25 imports json5; the config file is plain JSON
    sys.modules["json5"].load = json.load
    sys.modules["json5"].loads = json.loads
    from patchfusion.zoedepth.utils.config import get_org_config
    from patchfusion.zoedepth.models.zoedepth.zoedepth_v1 import ZoeDepth
    from patchfusion.zoedepth.models.base_models.depth_anything import DepthAnythingCore
    from patchfusion.zoedepth.models.base_models.dpt_dinov2.dpt import DPT_DINOv2
    import common.encode as renc
    from oracle import zoe as ozoe
    from oracle.weights import make_zoe_weights, ZOE_CONFIG

    cfg = get_org_config("zoedepth", "eval", dataset=None)
    for k in ("n_bins", "bin_embedding_dim", "n_attractors", "attractor_alpha", "attractor_gamma", "min_temp", "max_temp"):
        assert cfg[k] == ZOE_CONFIG[k], (k, cfg[k], ZOE_CONFIG[k])
    assert tuple(cfg["img_size"]) == ZOE_CONFIG["img_size"] and cfg["bin_centers_type"] == "softplus"
    assert cfg["attractor_kind"] == "mean" and cfg["attractor_type"] == "inv"
    c = DA_CONFIGS[encoder]
    net = DPT_DINOv2(encoder, c["features"], False, c["out_channels"], use_clstoken=False)
    kw = DepthAnythingCore.parse_img_size(dict(img_size=list(cfg["img_size"])))
    core = DepthAnythingCore(net, trainable=False, fetch_features=True, freeze_bn=True, img_size=kw["img_size"],
                             keep_aspect_ratio=cfg.get("force_keep_ar", False))
    core.output_channels = [c["features"]] * 5  # set_output_channels() hard-codes ViT-L's 256
    model = ZoeDepth(core, **{k: v for k, v in cfg.items() if k not in ("midas_model_type", "pretrained_resource", "use_pretrained_midas",
                                                                         "train_midas", "freeze_midas_bn")}).eval()
    sd = make_zoe_weights(encoder, 0)
    print("[zoe] load_state_dict strict:", model.load_state_dict(sd, strict=True))
    img = synthetic_frame(H, W, 0)
    # --- reference: bands/depth_anything.py:106-119
    img_pil = Image.fromarray(img)
    image = transforms.ToTensor()(img_pil).unsqueeze(0)
    pred_dict = model(image, dataset=None)
    depth = pred_dict["metric_depth"].squeeze().detach()
    pred_ref = np.asarray(Image.fromarray(depth.numpy()).resize(img_pil.size))
    # --- oracle
    taps = {}
    pred_or = ozoe.zoe_infer(sd, img, encoder, taps)
    e_net = rel(taps["metric"].squeeze().numpy(), depth.numpy())
    e = rel(pred_or, pred_ref)
    print(f"[zoe] metric depth (392x518) err {e_net:.3e}; prediction at frame size err {e:.3e}")
    assert e_net == 0.0 and np.array_equal(pred_or, pred_ref)
    dmin, dmax = pred_ref.min(), pred_ref.max()
    dn = (pred_ref - dmin) / (dmax - dmin)  # flip = (args.metric == 'none') = False
    rgb_ref = (renc.heat_to_rgb(dn.astype(np.float64)) * 255).astype(np.uint8)
    rgb_or, mn, mx = oda.da_encode(pred_or, flip=False)
    assert np.array_equal(rgb_or, rgb_ref)
    np.savez_compressed(os.path.join(GOLD, f"zoe_{encoder}_{H}x{W}.npz"), image=img, metric_net=depth.numpy(), prediction=pred_ref,
                        rgb=rgb_ref, dmin=np.float32(dmin), dmax=np.float32(dmax))
    print("[zoe] oracle == reference; wrote fixture")


if __name__ == "__main__":
    which = sys.argv[1:] or ["sizes", "da_small", "da_vits", "raft", "png", "zoe"]
    if "zoe" in which:
        golden_zoe()
    if "sizes" in which:
        golden_sizes()
    if "da_small" in which:
        golden_da("vits", 160, 208, "vits_160x208")   # net input 518x672 (lower_bound up-scales small frames)
    if "da_vits" in which:
        golden_da("vits", 480, 640, "vits_480x640")   # BASELINE config 1 stand-in
    if "raft" in which:
        golden_raft()
    if "png" in which:
        golden_png()

