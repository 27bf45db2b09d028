"""Pin the oracle against the reference and write the golden fixtures (TEST INFRASTRUCTURE).

Runs ONLY in the authoring container (needs /root/reference).  It
  1. imports the reference modules read-only (PYTHONPATH=/root/reference/bands, cwd
     /root/reference because d_anything/dpt.py:147 uses a relative torch.hub path),
  2. loads prisma_b200.seeded_weights' seeded state_dict into them (strict=True -> names/shapes pinned),
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
from prisma_b200.seeded_weights import make_da_weights, DA_CONFIGS
from prisma_b200.synthetic import synthetic_frame

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
    from prisma_b200.seeded_weights import make_raft_weights
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
    from prisma_b200.seeded_weights import make_zoe_weights, ZOE_CONFIG

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


def _load_vendored_solo_head():
    """Import the VENDORED bands/mmdet SOLOv2 head sources with mmcv absent: mmdet/__init__.py asserts an mmcv version, so the
    package is never imported; the needed files are loaded by path into a stub package tree, and the three mmcv primitives
    they use are stubbed from their documented behaviour: ConvModule (conv -> norm -> ReLU, sub-module names conv / gn /
    activate), BaseModule (= nn.Module), auto_fp16 / force_fp32 (identity decorators).  Everything else that runs -- the head
    construction, MaskFeatModule / SOLOV2Head forward, get_results, mask_matrix_nms, generate_coordinate -- is the
    reference's own code."""
    import importlib.util
    import types
    import torch.nn as nn
    root = os.path.join(REF, "bands", "mmdet")

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def from_file(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    class ConvModule(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None,
                     act_cfg=dict(type="ReLU"), inplace=True, bias="auto"):
            super().__init__()
            assert conv_cfg is None
            if bias == "auto":
                bias = norm_cfg is None
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)
            self.with_norm = norm_cfg is not None
            if self.with_norm:
                assert norm_cfg["type"] == "GN"
                self.gn = nn.GroupNorm(norm_cfg["num_groups"], out_channels)
            self.activate = nn.ReLU(inplace=inplace) if act_cfg is not None else None

        def forward(self, x):
            x = self.conv(x)
            if self.with_norm:
                x = self.gn(x)
            return self.activate(x) if self.activate is not None else x

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    ident = lambda *a, **k: (lambda f: f)
    def build_conv_layer(cfg, *a, **k):  # mmcv.cnn.build_conv_layer(None, ...) = nn.Conv2d
        assert cfg is None
        return nn.Conv2d(*a, **k)

    def build_norm_layer(cfg, num_features, postfix=""):  # mmcv.cnn.build_norm_layer(dict(type='BN'), n, postfix) -> ('bn<postfix>', BatchNorm2d)
        assert cfg["type"] == "BN"
        return "bn" + str(postfix), nn.BatchNorm2d(num_features)

    class Sequential(BaseModule, nn.Sequential):
        def __init__(self, *args, init_cfg=None):
            BaseModule.__init__(self, init_cfg)
            nn.Sequential.__init__(self, *args)

    mmcv = module("mmcv")
    mmcv.cnn = module("mmcv.cnn", ConvModule=ConvModule, build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer,
                      build_plugin_layer=None)
    mmcv.runner = module("mmcv.runner", BaseModule=BaseModule, Sequential=Sequential, auto_fp16=ident, force_fp32=ident)

    class InstanceData:  # mmdet/core/data_structures/instance_data.py, reduced to what get_results touches
        def __init__(self, meta=None):
            object.__setattr__(self, "_fields", {})
            for k, v in (meta or {}).items():
                object.__setattr__(self, k, v)

        def __setattr__(self, k, v):
            self._fields[k] = v
            object.__setattr__(self, k, v)

        def keys(self):
            return list(self._fields)

        def __len__(self):
            return len(next(iter(self._fields.values()))) if self._fields else 0

    nms = from_file("_ref_matrix_nms", "core/post_processing/matrix_nms.py")
    src = open(os.path.join(root, "core/utils/misc.py")).read()
    ns = {"torch": torch}
    start = src.index("def generate_coordinate(")
    exec(src[start:], ns)  # generate_coordinate is the last function of the file
    pkg = module("mmdet")
    pkg.__path__ = [root]
    module("mmdet.core", InstanceData=InstanceData, mask_matrix_nms=nms.mask_matrix_nms,
           multi_apply=lambda f, *a, **k: tuple(map(list, zip(*map(lambda *x: f(*x, **k), *a)))))
    module("mmdet.core.utils", center_of_mass=None, generate_coordinate=ns["generate_coordinate"])
    models = module("mmdet.models")
    models.__path__ = [os.path.join(root, "models")]

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    module("mmdet.models.builder", HEADS=_Reg(), BACKBONES=_Reg(), NECKS=_Reg(), build_loss=lambda cfg: None)
    module("mmdet.utils")
    module("mmdet.utils.misc", floordiv=lambda a, b: torch.div(a, b, rounding_mode="floor"))
    dh = module("mmdet.models.dense_heads")
    dh.__path__ = [os.path.join(root, "models", "dense_heads")]
    from_file("mmdet.models.dense_heads.base_mask_head", "models/dense_heads/base_mask_head.py")
    from_file("mmdet.models.dense_heads.solo_head", "models/dense_heads/solo_head.py")
    ut = module("mmdet.models.utils")
    ut.ResLayer = from_file("mmdet.models.utils.res_layer", "models/utils/res_layer.py").ResLayer
    bb = module("mmdet.models.backbones")
    bb.__path__ = [os.path.join(root, "models", "backbones")]
    nk = module("mmdet.models.necks")
    nk.__path__ = [os.path.join(root, "models", "necks")]
    resnet = from_file("mmdet.models.backbones.resnet", "models/backbones/resnet.py")
    fpn = from_file("mmdet.models.necks.fpn", "models/necks/fpn.py")
    return from_file("mmdet.models.dense_heads.solov2_head", "models/dense_heads/solov2_head.py"), nms, resnet, fpn


def golden_solo():
    """SOLOv2 head + decode of oracle/solo.py against the vendored reference sources (mmcv primitives stubbed)."""
    from oracle import solo as osolo
    from prisma_b200.seeded_weights import make_solo_weights, SOLO_CONFIGS
    mod, nms, resnet_mod, fpn_mod = _load_vendored_solo_head()
    c = SOLO_CONFIGS["tiny"]
    cfgd = dict(nms_pre=500, score_thr=0.1, mask_thr=0.5, filter_thr=0.05, kernel="gaussian", sigma=2.0, max_per_img=100)

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    head = mod.SOLOV2Head(num_classes=80, in_channels=256, feat_channels=512, stacked_convs=4, strides=[8, 8, 16, 32, 32],
                          scale_ranges=((1, 96), (48, 192), (96, 384), (192, 768), (384, 2048)), pos_scale=0.2,
                          num_grids=[40, 36, 24, 16, 12], cls_down_index=0,
                          mask_feature_head=dict(feat_channels=128, start_level=0, end_level=3, out_channels=256, mask_stride=4,
                                                 norm_cfg=dict(type="GN", num_groups=32, requires_grad=True)),
                          loss_mask=None, loss_cls=None, norm_cfg=dict(type="GN", num_groups=32, requires_grad=True),
                          test_cfg=Cfg(cfgd)).eval()
    sd = make_solo_weights("tiny", 0)
    hsd = {k[len("mask_head."):]: v for k, v in sd.items() if k.startswith("mask_head.")}
    print("[solo] load_state_dict strict:", head.load_state_dict(hsd, strict=True))
    g = torch.Generator().manual_seed(0)
    sizes = [(64, 88), (32, 44), (16, 22), (8, 11), (4, 6)]
    feats = [(torch.randn(1, 256, h, w, generator=g) * 0.3).half().float() for h, w in sizes]  # fp16-representable: stored as fp16
    meta = dict(img_shape=(250, 333, 3), ori_shape=(240, 320, 3), pad_shape=(256, 352, 3))
    kernels_ref, cls_ref, mf_ref = head(tuple(feats))
    res = head.get_results([k.clone() for k in kernels_ref], [t.clone() for t in cls_ref], mf_ref, img_metas=[meta])[0]
    # --- oracle
    mf = osolo.mask_feat(sd, feats)
    kernels, cls = osolo.head(sd, feats, c["num_grids"])
    e_mf = rel(mf.numpy(), mf_ref.numpy())
    e_k = max(rel(a.numpy(), b.numpy()) for a, b in zip(kernels, kernels_ref))
    e_c = max(rel(a.numpy(), b.numpy()) for a, b in zip(cls, cls_ref))
    print(f"[solo] mask feats err {e_mf:.3e}, kernel preds err {e_k:.3e}, cls preds err {e_c:.3e}")
    assert e_mf == 0.0 and e_k == 0.0 and e_c == 0.0
    scores, labels, masks = osolo.get_results(kernels, cls, mf, dict(img_shape=(250, 333), ori_shape=(240, 320)), osolo.TEST_CFG,
                                              c["strides"], c["num_grids"])
    print(f"[solo] instances: reference {len(res.scores)}, oracle {len(scores)}")
    assert len(res.scores) == len(scores) and torch.equal(res.labels, labels)
    assert torch.allclose(res.scores, scores, rtol=0, atol=0) and torch.equal(res.masks, masks)
    # --- backbone + neck: the vendored ResNet (depth 50 machinery, one bottleneck per stage = the "tiny" twin) and FPN
    class TinyResNet(resnet_mod.ResNet):
        arch_settings = {50: (resnet_mod.Bottleneck, (1, 1, 1, 1))}

    backbone = TinyResNet(depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1, norm_cfg=dict(type="BN", requires_grad=True),
                          norm_eval=True, style="pytorch")
    backbone.eval()  # the vendored ResNet.train() does not return self (resnet.py:648-660)
    neck = fpn_mod.FPN(in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=0, num_outs=5).eval()
    print("[solo] backbone strict:", backbone.load_state_dict({k[9:]: v for k, v in sd.items() if k.startswith("backbone.")}, strict=True))
    print("[solo] neck strict:", neck.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("neck.")}, strict=True))
    x = torch.randn(1, 3, 96, 128, generator=g)
    fpn_ref = neck(backbone(x))
    fpn_or = osolo.fpn(sd, osolo.resnet(sd, x, c["layers"]))
    e_f = max(rel(a.numpy(), b.numpy()) for a, b in zip(fpn_or, fpn_ref))
    print(f"[solo] ResNet + FPN (5 levels) err {e_f:.3e}")
    assert e_f == 0.0 and len(fpn_ref) == 5
    np.savez_compressed(os.path.join(GOLD, "solo_tiny_head.npz"), **{f"feat{i}": f.numpy().astype(np.float16) for i, f in enumerate(feats)},
                        net_x=x.numpy(), **{f"fpn{i}": f.numpy() for i, f in enumerate(fpn_ref)},
                        mask_feats_sub=mf_ref.numpy()[:, ::8], cls0=cls_ref[0].numpy(), kernel4=kernels_ref[4].numpy(),
                        scores=res.scores.numpy(), labels=res.labels.numpy(),
                        masks=np.packbits(res.masks.numpy(), axis=-1), n=len(res.scores))
    print("[solo] oracle == vendored SOLOV2Head forward + get_results + mask_matrix_nms; wrote fixture")


if __name__ == "__main__":
    which = sys.argv[1:] or ["sizes", "da_small", "da_vits", "raft", "png", "zoe", "solo"]
    if "solo" in which:
        golden_solo()
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

