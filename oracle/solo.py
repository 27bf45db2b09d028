"""CPU oracle: mask_mmdet band hot path -- SOLOv2 (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

PARITY PINNED AGAINST THE VENDORED SOURCES, WITH mmcv's PRIMITIVES STUBBED.  ``mmdet`` is vendored under bands/mmdet but
cannot be imported as a package: bands/mmdet/__init__.py:19-27 asserts mmcv 1.3.17-1.8.0 and mmcv-full is neither vendored
nor installed; the model config (models/solov2_r101_fpn_3x_coco.py) is downloaded, not in tree (its
values are taken from upstream mmdet 2.x).  oracle/tools/make_golden.py (golden_solo) therefore loads the vendored files BY
PATH -- models/backbones/resnet.py, models/utils/res_layer.py, models/necks/fpn.py, models/dense_heads/{base_mask_head,
solo_head,solov2_head}.py, core/post_processing/matrix_nms.py, generate_coordinate of core/utils/misc.py -- into a stub package
tree in which only mmcv's primitives are restated from their documented behaviour (ConvModule = conv -> GroupNorm -> ReLU
with sub-modules conv / gn / activate, build_conv_layer = nn.Conv2d, build_norm_layer = BatchNorm2d named bn<i>, BaseModule
= nn.Module, auto_fp16 / force_fp32 = identity), loads the seeded state_dict STRICT into the vendored ResNet, FPN and
SOLOV2Head, and asserts this file equal to them bit for bit (recorded differences 0.0: FPN levels, mask features, kernel
and class predictions, get_results scores / labels / masks incl. mask_matrix_nms).  Fixture: tests/golden/solo_tiny_head.npz,
replayed by tests/test_oracle_solo_golden.py.  Still restated without a runnable counterpart: the three mmcv image
primitives of the test pipeline (imrescale = cv2 INTER_LINEAR to (int(w*s+.5), int(h*s+.5)) with s = min(long/max(h,w),
short/min(h,w)); imnormalize = f32, subtract mean, multiply by 1/std; impad to a multiple of 32) and the detector glue.
Independent cross-check (tests/test_oracle_crosschecks.py): resnet() and fpn() also equal torchvision's ResNet / FPN.
"""
import cv2
import numpy as np
import torch
import torch.nn.functional as F

from prisma_b200.seeded_weights import SOLO_CONFIGS

MEAN = np.array([123.675, 116.28, 103.53], np.float32)  # _base_/datasets/coco_instance.py:4-5 (RGB order, to_rgb=True)
STD = np.array([58.395, 57.12, 57.375], np.float32)
BAND_CLASSES = [0, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23]  # bands/mask_mmdet.py:29: person, bird ... giraffe (COCO ids)
TEST_CFG = dict(nms_pre=500, score_thr=0.1, mask_thr=0.5, filter_thr=0.05, sigma=2.0, max_per_img=100)


def solo_preprocess(img_rgb_u8, img_scale=(1333, 800)):
    """test pipeline (_base_/datasets/coco_instance.py:16-32; datasets/pipelines/transforms.py:215-243 Resize keep_ratio,
    :696-711 Normalize, :622-638 Pad size_divisor 32).  The band hands the detector a BGR image and the pipeline converts
    it back (to_rgb); resizing is per channel, so the RGB frame is used directly.  -> (1x3xHpxWp f32, meta)."""
    h, w = img_rgb_u8.shape[:2]
    s = min(max(img_scale) / max(h, w), min(img_scale) / min(h, w))
    nw, nh = int(w * s + 0.5), int(h * s + 0.5)
    img = cv2.resize(img_rgb_u8, (nw, nh), interpolation=cv2.INTER_LINEAR)
    x = (img.astype(np.float32) - MEAN) * (1.0 / STD.astype(np.float64)).astype(np.float32)
    hp, wp = (nh + 31) // 32 * 32, (nw + 31) // 32 * 32
    pad = np.zeros((hp, wp, 3), np.float32)
    pad[:nh, :nw] = x
    t = torch.from_numpy(np.ascontiguousarray(pad.transpose(2, 0, 1)))[None]
    return t, dict(img_shape=(nh, nw), ori_shape=(h, w), pad_shape=(hp, wp), resized_u8=img)


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def resnet(sd, x, layers):
    """ResNet.forward (models/backbones/resnet.py:631-646), Bottleneck style='pytorch' (:154-157: stride on the 3x3)."""
    x = F.relu(_bn(sd, "backbone.bn1", F.conv2d(x, sd["backbone.conv1.weight"], None, stride=2, padding=3)))
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    outs = []
    for li, blocks in enumerate(layers):
        for b in range(blocks):
            p = f"backbone.layer{li + 1}.{b}."
            stride = 2 if (b == 0 and li > 0) else 1
            out = F.relu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"])))
            out = F.relu(_bn(sd, p + "bn2", F.conv2d(out, sd[p + "conv2.weight"], None, stride=stride, padding=1)))
            out = _bn(sd, p + "bn3", F.conv2d(out, sd[p + "conv3.weight"]))
            if b == 0:
                x = _bn(sd, p + "downsample.1", F.conv2d(x, sd[p + "downsample.0.weight"], None, stride=stride))
            x = F.relu(out + x)
        outs.append(x)
    return outs


def fpn(sd, feats):
    """FPN.forward (models/necks/fpn.py:151-204): nearest top-down, 3x3 output convs, extra level by max_pool2d(1, 2)."""
    lat = [F.conv2d(f, sd[f"neck.lateral_convs.{i}.conv.weight"], sd[f"neck.lateral_convs.{i}.conv.bias"]) for i, f in enumerate(feats)]
    for i in range(3, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = [F.conv2d(lat[i], sd[f"neck.fpn_convs.{i}.conv.weight"], sd[f"neck.fpn_convs.{i}.conv.bias"], padding=1) for i in range(4)]
    outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    return outs


def _coord(t):
    """generate_coordinate (core/utils/misc.py:190-208): x then y in [-1, 1]."""
    x_range = torch.linspace(-1, 1, t.shape[-1])
    y_range = torch.linspace(-1, 1, t.shape[-2])
    y, x = torch.meshgrid(y_range, x_range, indexing="ij")
    return torch.cat([x.expand(t.shape[0], 1, -1, -1), y.expand(t.shape[0], 1, -1, -1)], 1)


def _cgr(sd, p, x, padding=1):
    """ConvModule(norm_cfg=GN-32): conv (no bias) -> GroupNorm(32) -> ReLU."""
    x = F.conv2d(x, sd[p + ".conv.weight"], None, padding=padding)
    return F.relu(F.group_norm(x, 32, sd[p + ".gn.weight"], sd[p + ".gn.bias"], 1e-5))


def mask_feat(sd, feats):
    """MaskFeatModule.forward (models/dense_heads/solov2_head.py:133-150), start_level 0, end_level 3."""
    m = "mask_head.mask_feature_head."
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    acc = _cgr(sd, m + "convs_all_levels.0.conv0", feats[0])
    for i in range(1, 4):
        x = feats[i]
        if i == 3:
            x = torch.cat([x, _coord(x)], 1)
        for j in range(i):
            x = up(_cgr(sd, f"{m}convs_all_levels.{i}.conv{j}", x))
        acc = acc + x
    return _cgr(sd, m + "conv_pred", acc, padding=0)


def head(sd, feats, num_grids):
    """SOLOV2Head.forward (solov2_head.py:253-292) + resize_feats (solo_head.py:133-153)."""
    rs = [F.interpolate(feats[0], size=feats[1].shape[-2:], mode="bilinear", align_corners=False), feats[1], feats[2], feats[3],
          F.interpolate(feats[4], size=feats[3].shape[-2:], mode="bilinear", align_corners=False)]
    kernels, cls = [], []
    for lvl, f in enumerate(rs):
        k = torch.cat([f, _coord(f)], 1)
        k = F.interpolate(k, size=num_grids[lvl], mode="bilinear", align_corners=False)
        c = k[:, :-2]
        for i in range(4):
            k = _cgr(sd, f"mask_head.kernel_convs.{i}", k)
        kernels.append(F.conv2d(k, sd["mask_head.conv_kernel.weight"], sd["mask_head.conv_kernel.bias"], padding=1))
        for i in range(4):
            c = _cgr(sd, f"mask_head.cls_convs.{i}", c)
        cls.append(F.conv2d(c, sd["mask_head.conv_cls.weight"], sd["mask_head.conv_cls.bias"], padding=1))
    return kernels, cls


def matrix_nms(masks, labels, scores, mask_area, nms_pre, max_num, sigma, filter_thr):
    """mask_matrix_nms (core/post_processing/matrix_nms.py:5-121), gaussian kernel.  -> scores, labels, keep_inds."""
    scores, sort_inds = torch.sort(scores, descending=True, stable=True)
    keep_inds = sort_inds
    if len(sort_inds) > nms_pre:
        sort_inds, keep_inds, scores = sort_inds[:nms_pre], keep_inds[:nms_pre], scores[:nms_pre]
    masks, mask_area, labels = masks[sort_inds], mask_area[sort_inds], labels[sort_inds]
    n = len(labels)
    flat = masks.reshape(n, -1).float()
    inter = flat @ flat.t()
    area = mask_area.expand(n, n)
    iou = (inter / (area + area.t() - inter)).triu(diagonal=1)
    lab = (labels.expand(n, n) == labels.expand(n, n).t()).triu(diagonal=1)
    decay_iou = iou * lab
    comp = decay_iou.max(0)[0].expand(n, n).t()
    coef = (torch.exp(-sigma * decay_iou ** 2) / torch.exp(-sigma * comp ** 2)).min(0)[0]
    scores = scores * coef
    keep = scores >= filter_thr
    keep_inds, scores, labels = keep_inds[keep], scores[keep], labels[keep]
    scores, si = torch.sort(scores, descending=True, stable=True)
    keep_inds, labels = keep_inds[si], labels[si]
    return scores[:max_num], labels[:max_num], keep_inds[:max_num]


def get_results(kernels, cls, mfeat, meta, cfg, strides_cfg, num_grids, taps=None):
    """SOLOV2Head.get_results/_get_results_single (solov2_head.py:582-766).  -> scores [n], labels [n], masks bool [n,H,W]."""
    empty = (torch.zeros(0), torch.zeros(0, dtype=torch.long), torch.zeros((0,) + tuple(meta["ori_shape"]), dtype=torch.bool))
    flat_cls, flat_k = [], []
    for lvl in range(5):
        s = cls[lvl].sigmoid()
        local_max = F.max_pool2d(s, 2, stride=1, padding=1)
        s = s * (local_max[:, :, :-1, :-1] == s)
        flat_cls.append(s[0].permute(1, 2, 0).reshape(-1, s.shape[1]))
        flat_k.append(kernels[lvl][0].permute(1, 2, 0).reshape(-1, kernels[lvl].shape[1]))
    cls_scores, kernel_preds = torch.cat(flat_cls), torch.cat(flat_k)
    if taps is not None:
        taps["cls_scores"], taps["kernel_preds"] = cls_scores, kernel_preds
    fh, fw = mfeat.shape[-2:]
    h, w = meta["img_shape"]
    score_mask = cls_scores > cfg["score_thr"]
    scores = cls_scores[score_mask]
    if len(scores) == 0:
        return empty
    inds = score_mask.nonzero()
    labels = inds[:, 1]
    kp = kernel_preds[inds[:, 0]]
    strides = torch.cat([torch.full((g * g,), float(s)) for g, s in zip(num_grids, strides_cfg)])[inds[:, 0]]
    mask_preds = F.conv2d(mfeat, kp[:, :, None, None]).squeeze(0).sigmoid()
    masks = mask_preds > cfg["mask_thr"]
    sum_masks = masks.sum((1, 2)).float()
    keep = sum_masks > strides
    if keep.sum() == 0:
        return empty
    masks, mask_preds, sum_masks, scores, labels = masks[keep], mask_preds[keep], sum_masks[keep], scores[keep], labels[keep]
    scores = scores * ((mask_preds * masks).sum((1, 2)) / sum_masks)
    if taps is not None:
        taps["cand_scores"], taps["cand_labels"], taps["cand_area"] = scores.clone(), labels.clone(), sum_masks.clone()
    scores, labels, keep_inds = matrix_nms(masks, labels, scores, sum_masks, cfg["nms_pre"], cfg["max_per_img"], cfg["sigma"], cfg["filter_thr"])
    if len(scores) == 0:
        return empty
    mp = F.interpolate(mask_preds[keep_inds].unsqueeze(0), size=(fh * 4, fw * 4), mode="bilinear", align_corners=False)[:, :, :h, :w]
    mp = F.interpolate(mp, size=tuple(meta["ori_shape"]), mode="bilinear", align_corners=False).squeeze(0)
    return scores, labels, mp > cfg["mask_thr"]


def solo_infer(sd, img_rgb_u8, variant="r101", taps=None):
    """inference_detector(model, img) (apis/inference.py:99-162) up to the InstanceData: (scores, labels, masks)."""
    c = SOLO_CONFIGS[variant]
    with torch.no_grad():
        x, meta = solo_preprocess(img_rgb_u8, c["img_scale"])
        feats = fpn(sd, resnet(sd, x, c["layers"]))
        mf = mask_feat(sd, feats)
        kernels, cls = head(sd, feats, c["num_grids"])
        if taps is not None:
            taps.update(net_input=x, fpn=feats, mask_feats=mf, kernels=kernels, cls=cls, meta=meta)
        return get_results(kernels, cls, mf, meta, TEST_CFG, c["strides"], c["num_grids"], taps)


def band_union(scores, labels, masks, confidence=0.5):
    """The band's frame (bands/mask_mmdet.py:43-61,134-146): sum of 255*mask over instances of the 11 classes with
    score > 0.5 (getTotalMasks) and > --confidence, as u8 (the float sum wraps modulo 256 in the cast) replicated on RGB."""
    h, w = masks.shape[-2:]
    acc = np.zeros((h, w), np.int64)
    for s, l, m in zip(scores.tolist(), labels.tolist(), masks.numpy()):
        if l in BAND_CLASSES and s > 0.5 and s > confidence:
            acc += 255 * m.astype(np.int64)
    u = (acc % 256).astype(np.uint8)
    return np.stack([u] * 3, axis=-1)


def sdf_green(union_u8):
    """getSDF + the green-channel write (bands/mask_mmdet.py:64-69,150-152).  snowy (prideout/snowy, un-vendored) defines
    generate_sdf(mask) = generate_udf(mask) - generate_udf(~mask), udf = sqrt of the exact squared Euclidean distance to
    the nearest set pixel (Felzenszwalb-Huttenlocher); scipy's exact EDT is the independent implementation used here."""
    from scipy import ndimage
    m = union_u8 != 0
    ua = ndimage.distance_transform_edt(~m) if m.any() else np.full(m.shape, 1e10)
    ub = ndimage.distance_transform_edt(m) if (~m).any() else np.full(m.shape, 1e10)
    sdf = ua - ub
    sdf = (sdf + 127.0) / 255.0
    sdf = (sdf - 0.25) * 2.0
    return ((1.0 - np.clip(sdf, 0.0, 1.0)) * 255).astype(np.uint8)
