"""CPU oracle: depth_midas band hot path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

PARITY UNPINNED.  The arithmetic of this band is not in the reference tree: bands/depth_midas.py:35-42 loads it with
``torch.hub.load("intel-isl/MiDaS", "DPT_Large")`` (default branch, no pinned commit) on top of ``timm==0.6.12``
(environment.yml:31); neither is vendored, installed or reachable offline.  What follows restates
the *published* MiDaS v3 algorithm (upstream file names cited per function) and is anchored on the reference's own call
sites: the transform / model / bicubic-resize sequence of ``infer`` (bands/depth_midas.py:49-75) and the encode of the
video loop (:141-146).  The Resize class of the MiDaS transform is the one vendored in
bands/d_anything/util/transform.py:54-166 (same ``get_size``), which pins the size arithmetic.

The fusion blocks and residual units are the same modules as Depth-Anything's (d_anything/blocks.py copies
midas/blocks.py), so oracle.da._rcu/_fusion -- which ARE pinned against the reference -- are reused.
Independent cross-checks: the WHOLE network (patch embed, pos-embed resize, blocks, hooks, project readout, reassemble, fusion,
head) agrees to < 2e-5 with HuggingFace transformers' DPTForDepthEstimation -- the port of MiDaS v3 DPT-Large behind
`Intel/dpt-large` -- with the seeded weights mapped onto its parameter names (tests/test_oracle_midas_hf.py); _block() equals
torchvision's ViT EncoderBlock (tests/test_oracle_crosschecks.py).
"""
import math

import cv2
import numpy as np
import torch
import torch.nn.functional as F

from .da import _conv, _fusion, _ident, _linear, heat_to_rgb
from prisma_b200.seeded_weights import MIDAS_CONFIGS


def midas_get_size(width, height, target=384, multiple=32):
    """Resize.get_size with resize_method="upper_bound", keep_aspect_ratio=True (the Resize class of midas/transforms.py is
    the one vendored at d_anything/util/transform.py:111-166; constrain_to_multiple_of :98-109 with max_val = target)."""
    scale_h, scale_w = target / height, target / width
    if scale_w < scale_h:
        scale_h = scale_w
    else:
        scale_w = scale_h

    def c(x):
        y = int(np.round(x / multiple) * multiple)
        if y > target:
            y = int(np.floor(x / multiple) * multiple)
        return y
    return c(scale_w * width), c(scale_h * height)


def midas_preprocess(img_u8):
    """hubconf.transforms().default_transform -- what bands/depth_midas.py:37-40 selects for "midas3" (DPT_Large): /255 (f64),
    Resize(384, 384, keep AR, x32, "upper_bound", INTER_CUBIC), NormalizeImage(ImageNet mean / std), PrepareForNet (CHW f32).
    (Upstream's README pairs DPT models with dpt_transform ("minimal", mean = std = 0.5); the reference band does not, and the
    engine mirrors the band.  The hub code is not vendored: this is restated from the published intel-isl/MiDaS hubconf.py.)"""
    image = img_u8 / 255.0
    w, h = midas_get_size(image.shape[1], image.shape[0])
    image = cv2.resize(image, (w, h), interpolation=cv2.INTER_CUBIC)
    image = (image - np.array([0.485, 0.456, 0.406])) / np.array([0.229, 0.224, 0.225])
    return np.ascontiguousarray(np.transpose(image, (2, 0, 1))).astype(np.float32)


def _resize_pos_embed(posemb, gs_h, gs_w):
    """midas/backbones/vit.py _resize_pos_embed (start_index = 1): bilinear, align_corners=False."""
    tok, grid = posemb[:, :1], posemb[0, 1:]
    gs_old = int(math.sqrt(len(grid)))
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gs_h, gs_w), mode="bilinear")
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_h * gs_w, -1)
    return torch.cat([tok, grid], dim=1)


def _block(sd, prefix, x, heads, q):
    """timm 0.6.12 vision_transformer.Block / Attention / Mlp: pre-norm, no LayerScale (init_values=None), LN eps 1e-6,
    attn = softmax((q @ k^T) * head_dim^-0.5), exact-erf GELU."""
    B, N, C = x.shape
    hd = C // heads
    y = F.layer_norm(x, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps=1e-6)
    qkv = _linear(y, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"], q)
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    attn = ((q(qkv[0]) @ q(qkv[1]).transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
    y = (q(attn) @ q(qkv[2])).transpose(1, 2).reshape(B, N, C)
    x = x + _linear(y, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"], q)
    y = F.layer_norm(x, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps=1e-6)
    y = F.gelu(_linear(y, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"], q))
    return x + _linear(y, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"], q)


def midas_model(sd, x, variant="dpt_large", q=_ident, taps=None):
    """DPTDepthModel.forward (midas/dpt_depth.py) = forward_vit (midas/backbones/vit.py: forward_flex + hooks +
    act_postprocessN) -> scratch.layerN_rn -> refinenet4..1 -> scratch.output_conv -> squeeze.  1x3xhxw -> 1xhxw."""
    c = MIDAS_CONFIGS[variant]
    p = "pretrained.model."
    h, w = x.shape[-2:]
    ph, pw = h // 16, w // 16
    t = _conv(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], q, stride=16).flatten(2).transpose(1, 2)
    t = torch.cat((sd[p + "cls_token"].expand(t.shape[0], -1, -1), t), dim=1)
    t = t + _resize_pos_embed(sd[p + "pos_embed"], ph, pw)
    if taps is not None:
        taps["tokens"] = t
    acts = []
    for i in range(c["depth"]):
        t = _block(sd, f"{p}blocks.{i}.", t, c["heads"], q)
        if i in c["hooks"]:
            acts.append(t)  # forward hook on the block: raw output, no final norm
    layers = []
    for i, a in enumerate(acts):
        pre = f"pretrained.act_postprocess{i + 1}."
        # ProjectReadout (midas/backbones/utils.py): GELU(Linear(cat(tokens, cls expanded)))
        ro = a[:, :1].expand_as(a[:, 1:])
        f = F.gelu(_linear(torch.cat((a[:, 1:], ro), -1), sd[pre + "0.project.0.weight"], sd[pre + "0.project.0.bias"], q))
        f = f.transpose(1, 2).reshape(f.shape[0], -1, ph, pw)
        f = _conv(f, sd[pre + "3.weight"], sd[pre + "3.bias"], q)
        if i == 0:
            f = F.conv_transpose2d(q(f), q(sd[pre + "4.weight"]), sd[pre + "4.bias"], stride=4)
        elif i == 1:
            f = F.conv_transpose2d(q(f), q(sd[pre + "4.weight"]), sd[pre + "4.bias"], stride=2)
        elif i == 3:
            f = _conv(f, sd[pre + "4.weight"], sd[pre + "4.bias"], q, stride=2, padding=1)
        layers.append(f)
    if taps is not None:
        taps["layers"] = layers
    s = "scratch."
    rn = [_conv(layers[i], sd[f"{s}layer{i + 1}_rn.weight"], None, q, padding=1) for i in range(4)]
    p4 = _fusion(sd, s + "refinenet4.", q, rn[3], size=rn[2].shape[2:])
    p3 = _fusion(sd, s + "refinenet3.", q, p4, rn[2], size=rn[1].shape[2:])
    p2 = _fusion(sd, s + "refinenet2.", q, p3, rn[1], size=rn[0].shape[2:])
    p1 = _fusion(sd, s + "refinenet1.", q, p2, rn[0])
    o = _conv(p1, sd[s + "output_conv.0.weight"], sd[s + "output_conv.0.bias"], q, padding=1)
    o = F.interpolate(o, scale_factor=2, mode="bilinear", align_corners=True)
    o = F.relu(_conv(o, sd[s + "output_conv.2.weight"], sd[s + "output_conv.2.bias"], q, padding=1))
    o = F.relu(_conv(o, sd[s + "output_conv.4.weight"], sd[s + "output_conv.4.bias"], q))  # non_negative=True
    if taps is not None:
        taps["net_depth"] = o
    return o.squeeze(1)


def midas_infer(sd, img_u8, variant="dpt_large", q=_ident):
    """infer(img, normalize=False) (bands/depth_midas.py:49-75): transform -> model -> bicubic(align_corners=True) to
    the frame size -> f32 HxW."""
    x = torch.from_numpy(midas_preprocess(img_u8)).unsqueeze(0)
    with torch.no_grad():
        pred = midas_model(sd, x, variant, q)
        pred = F.interpolate(pred.unsqueeze(1), size=img_u8.shape[:2], mode="bicubic", align_corners=True).squeeze()
    return pred.numpy().astype(np.float32)


def midas_encode(prediction):
    """video loop encode (bands/depth_midas.py:141-146).  Unlike depth_anything.py:219 the f32 array goes into
    heat_to_rgb uncast, so (1-h)*0.65, *6 and +{4,2} round in f32 (common/encode.py:13-33)."""
    dmin, dmax = prediction.min(), prediction.max()
    depth = 1.0 - (prediction - dmin) / (dmax - dmin)
    return (heat_to_rgb(depth) * 255).astype(np.uint8), float(dmin), float(dmax)
