"""CPU oracle: Depth-Anything band hot path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

A functional restatement (plain torch CPU fp32 / numpy f64) of what
``bands/depth_anything.py:infer`` (:100-143, ``--metric none``) and the video loop
(:203-221) compute per frame.  Every function cites the reference lines it follows.
Pinned against the imported reference modules by oracle/tools/make_golden.py.

The optional ``q`` argument of the model functions is a hook applied to the two
operands of every contraction (GEMM / conv); it is used only by
oracle/tools/precision_study.py to emulate fp16/bf16 operand rounding.
"""
import math

import cv2
import numpy as np
import torch
import torch.nn.functional as F

from prisma_b200.seeded_weights import DA_CONFIGS

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406])  # depth_anything.py:72
IMAGENET_STD = np.array([0.229, 0.224, 0.225])


# --------------------------------------------------------------------------- pre-process
def da_get_size(width, height, target=518, multiple=14):
    """Resize.get_size, lower_bound + keep_aspect_ratio (d_anything/util/transform.py:111-166)."""
    scale_h = target / height
    scale_w = target / width
    if scale_w > scale_h:
        scale_h = scale_w
    else:
        scale_w = scale_h

    def constrain(x, min_val):
        y = int(np.round(x / multiple) * multiple)
        if y < min_val:
            y = int(np.ceil(x / multiple) * multiple)
        return y

    return constrain(scale_w * width, target), constrain(scale_h * height, target)


def da_preprocess(img_u8):
    """HxWx3 u8 RGB -> 3xhxw f32 (depth_anything.py:122-126; transform.py:168-174,219-222,232-234)."""
    image = img_u8 / 255.0  # f64
    w, h = da_get_size(image.shape[1], image.shape[0])
    image = cv2.resize(image, (w, h), interpolation=cv2.INTER_CUBIC)
    image = (image - IMAGENET_MEAN) / IMAGENET_STD
    return np.ascontiguousarray(np.transpose(image, (2, 0, 1))).astype(np.float32)


# --------------------------------------------------------------------------- ViT encoder
def _ident(t):
    return t


def _linear(x, w, b, q):
    return F.linear(q(x), q(w), b)


def _conv(x, w, b, q, **kw):
    return F.conv2d(q(x), q(w), b, **kw)


def interpolate_pos_encoding(pos_embed, ph, pw):
    """vision_transformer.py:179-210 (names w/h are swapped there; ph = H/14, pw = W/14).

    bicubic, align_corners=False, coordinates mapped with the *given* scale_factor
    ((ph+0.1)/37, (pw+0.1)/37), interpolate_offset = 0.1.
    """
    n = pos_embed.shape[1] - 1
    s = int(math.sqrt(n))
    dim = pos_embed.shape[-1]
    if ph * pw == n and ph == pw:
        return pos_embed
    cls_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:].reshape(1, s, s, dim).permute(0, 3, 1, 2)
    sx, sy = float(ph + 0.1) / math.sqrt(n), float(pw + 0.1) / math.sqrt(n)
    patch_pos = F.interpolate(patch_pos, scale_factor=(sx, sy), mode="bicubic")
    assert patch_pos.shape[-2] == ph and patch_pos.shape[-1] == pw
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pos.unsqueeze(0), patch_pos), dim=1)


def vit_tokens(sd, x, q=_ident):
    """prepare_tokens_with_masks (vision_transformer.py:212-231) + PatchEmbed (patch_embed.py:69-82)."""
    p = "pretrained."
    ph, pw = x.shape[-2] // 14, x.shape[-1] // 14
    t = _conv(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], q, stride=14)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd[p + "cls_token"].expand(t.shape[0], -1, -1), t), dim=1)
    return t + interpolate_pos_encoding(sd[p + "pos_embed"], ph, pw)


def vit_block(sd, prefix, x, heads, q=_ident):
    """Block.forward (block.py:82-107), Attention.forward (attention.py:49-62), Mlp (mlp.py:35-41)."""
    B, N, C = x.shape
    hd = C // heads
    y = F.layer_norm(x, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps=1e-6)
    qkv = _linear(y, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"], q)
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    qq, kk, vv = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = (q(qq) @ q(kk).transpose(-2, -1)).softmax(dim=-1)
    y = (q(attn) @ q(vv)).transpose(1, 2).reshape(B, N, C)
    y = _linear(y, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"], q)
    x = x + sd[prefix + "ls1.gamma"] * y
    y = F.layer_norm(x, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps=1e-6)
    y = _linear(y, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"], q)
    y = F.gelu(y)  # nn.GELU default = exact erf
    y = _linear(y, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"], q)
    return x + sd[prefix + "ls2.gamma"] * y


def vit_features(sd, x, encoder, q=_ident, taps=None):
    """get_intermediate_layers(x, 4, norm=True) patch tokens (vision_transformer.py:274-281,297-321)."""
    c = DA_CONFIGS[encoder]
    t = vit_tokens(sd, x, q)
    if taps is not None:
        taps["tokens"] = t
    feats = []
    for i in range(c["depth"]):
        t = vit_block(sd, f"pretrained.blocks.{i}.", t, c["heads"], q)
        if taps is not None and i == 0:
            taps["block0"] = t
        if i >= c["depth"] - 4:
            n = F.layer_norm(t, (c["dim"],), sd["pretrained.norm.weight"], sd["pretrained.norm.bias"], eps=1e-6)
            feats.append(n[:, 1:])
    return feats


# --------------------------------------------------------------------------- DPT head
def _rcu(sd, prefix, x, q):
    """ResidualConvUnit.forward (d_anything/blocks.py:69-92), bn=False, ReLU not in place."""
    out = F.relu(x)
    out = _conv(out, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], q, padding=1)
    out = F.relu(out)
    out = _conv(out, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], q, padding=1)
    return out + x


def _fusion(sd, prefix, q, x0, x1=None, size=None):
    """FeatureFusionBlock.forward (blocks.py:126-153): align_corners=True bilinear, then 1x1 out_conv."""
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, prefix + "resConfUnit1.", x1, q)
    out = _rcu(sd, prefix + "resConfUnit2.", out, q)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return _conv(out, sd[prefix + "out_conv.weight"], sd[prefix + "out_conv.bias"], q)


def dpt_head(sd, feats, ph, pw, q=_ident, taps=None):
    """DPTHead.forward (d_anything/dpt.py:103-136), use_clstoken=False."""
    h = "depth_head."
    out = []
    for i, x in enumerate(feats):
        x = x.permute(0, 2, 1).reshape(x.shape[0], x.shape[-1], ph, pw)
        x = _conv(x, sd[f"{h}projects.{i}.weight"], sd[f"{h}projects.{i}.bias"], q)
        if i == 0:
            x = F.conv_transpose2d(q(x), q(sd[h + "resize_layers.0.weight"]), sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(q(x), q(sd[h + "resize_layers.1.weight"]), sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = _conv(x, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], q, stride=2, padding=1)
        out.append(x)
    s = h + "scratch."
    rn = [_conv(out[i], sd[f"{s}layer{i + 1}_rn.weight"], None, q, padding=1) for i in range(4)]
    if taps is not None:
        taps["layer_rn"] = rn
    p4 = _fusion(sd, s + "refinenet4.", q, rn[3], size=rn[2].shape[2:])
    p3 = _fusion(sd, s + "refinenet3.", q, p4, rn[2], size=rn[1].shape[2:])
    p2 = _fusion(sd, s + "refinenet2.", q, p3, rn[1], size=rn[0].shape[2:])
    p1 = _fusion(sd, s + "refinenet1.", q, p2, rn[0])
    if taps is not None:
        taps["path"] = [p1, p2, p3, p4]
    o = _conv(p1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], q, padding=1)
    o = F.interpolate(o, (ph * 14, pw * 14), mode="bilinear", align_corners=True)
    o = F.relu(_conv(o, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], q, padding=1))
    if taps is not None:
        taps["out_conv_act"] = o  # output of output_conv2[1] (the hook of the metric head, base_models/depth_anything.py:302-304)
    o = F.relu(_conv(o, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"], q))
    return o


def da_model(sd, x, encoder, q=_ident, taps=None):
    """DPT_DINOv2.forward (d_anything/dpt.py:155-166): 1x3xhxw f32 -> 1xhxw f32."""
    h, w = x.shape[-2:]
    feats = vit_features(sd, x, encoder, q, taps)
    if taps is not None:
        taps["feats"] = feats
    d = dpt_head(sd, feats, h // 14, w // 14, q, taps)
    d = F.interpolate(d, size=(h, w), mode="bilinear", align_corners=True)
    return F.relu(d).squeeze(1)


# --------------------------------------------------------------------------- post-process / encode
def hue_to_rgb(hue):
    """common/encode.py:13-28 (array branch; f64 because np.zeros defaults to f64)."""
    rgb = np.zeros((hue.shape[0], hue.shape[1], 3))
    rgb[..., 0] = hue * 6.0
    rgb[..., 1] = hue * 6.0 + 4.0
    rgb[..., 2] = hue * 6.0 + 2.0
    rgb = np.abs(np.mod(rgb, 6.0) - 3.0) - 1.0
    return np.clip(rgb, 0.0, 1.0)


def heat_to_rgb(heat):
    """common/encode.py:31-33."""
    return hue_to_rgb((1.0 - heat) * 0.65)


def da_upsample(depth, h, w):
    """depth_anything.py:132: bilinear align_corners=False to the frame size; 1xhnxwn -> HxW f32 numpy."""
    return F.interpolate(depth[None], (h, w), mode="bilinear", align_corners=False)[0, 0].numpy()


def da_encode(prediction, flip=True):
    """Video-loop encode (depth_anything.py:215-221): HxW f32 -> (HxWx3 u8, min, max).

    flip = (args.metric == 'none') (depth_anything.py:188) -> True on the relative path.
    The u8 cast truncates.
    """
    dmin = prediction.min()
    dmax = prediction.max()
    depth = (prediction - dmin) / (dmax - dmin)
    if flip:
        depth = 1.0 - depth
    rgb = (heat_to_rgb(depth.astype(np.float64)) * 255).astype(np.uint8)
    return rgb, float(dmin), float(dmax)


def da_infer(sd, img_u8, encoder, q=_ident):
    """depth_anything.infer(img) (depth_anything.py:100-143, normalize=False): HxWx3 u8 -> HxW f32."""
    h, w = img_u8.shape[:2]
    x = torch.from_numpy(da_preprocess(img_u8)).unsqueeze(0)
    with torch.no_grad():
        depth = da_model(sd, x, encoder, q)
    return da_upsample(depth, h, w)


# --------------------------------------------------------------------------- PNG variant (process_image path)
def float_to_edge(channel, ksize=1):
    """common/encode.py:81-95: u8 image -> Sobel (ksize=1: [-1,0,1], BORDER_REFLECT_101) magnitude, normalised to its max."""
    img = (channel * 255).astype(np.uint8)
    sobel_x = cv2.Sobel(img, cv2.CV_64F, 1, 0, ksize=ksize)
    sobel_y = cv2.Sobel(img, cv2.CV_64F, 0, 1, ksize=ksize)
    sobel_mag = np.sqrt(np.square(sobel_x) + np.square(sobel_y))
    sobel_mag *= 255.0 / sobel_mag.max()
    return sobel_mag / 255.0


def float_to_rgb(value, min_value=0.0, max_value=1.0, base=256):
    """common/encode.py:141-146 (24-bit packing of the depth range into one pixel)."""
    L = np.clip((value - min_value) / (max_value - min_value), 0.0, 1.0) * (base * base * base - 1)
    return ((np.floor(L % base)) / (base - 1),
            (np.floor(L / base) % base) / (base - 1),
            (np.floor(L / (base * base)) % base) / (base - 1))


def da_write_depth_rgb(prediction, flip=True):
    """write_depth(..., normalize=True, heatmap=True, encode_range=True) up to the cv2.imwrite (common/io.py:138-166):
    HxW f32 -> HxWx3 u8 RGB: heat map, Sobel edges in the saturation, (min,max) packed into pixels (0,0) and (0,1)."""
    depth_min = prediction.min()
    depth_max = prediction.max()
    depth = (prediction - depth_min) / (depth_max - depth_min)
    if flip:
        depth = 1.0 - depth
    edge = float_to_edge(depth, ksize=1)
    depth = depth.astype(np.float64)
    rgb = heat_to_rgb(depth)
    sat = 1.0 - edge
    for c in range(3):  # saturation() of common/encode.py:73-78
        rgb[..., c] = rgb[..., c] * sat + (1.0 - sat)
    rgb[0, 0] = float_to_rgb(depth_min, 0.0, 1000.0)
    rgb[0, 1] = float_to_rgb(depth_max, 0.0, 1000.0)
    return (rgb * 255).astype(np.uint8), float(depth_min), float(depth_max)
