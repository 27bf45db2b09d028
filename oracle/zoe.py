"""CPU oracle: depth_anything band, metric path (`--metric indoor|outdoor`, what process.py:53 passes by default)
(TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates ZoeDepth on the Depth-Anything core (all paths relative to bands/patchfusion/zoedepth/):
models/zoedepth/zoedepth_v1.py:127-211 (forward), models/base_models/depth_anything.py:176-189,261-277,298-320
(PrepForMidas, DepthAnythingCore.forward and its hooks), models/layers/{localbins_layers,attractor,dist_layers}.py, and the
band's pre/post (bands/depth_anything.py:106-119).  Pinned against the imported reference modules by
oracle/tools/make_golden.py (golden_zoe).
"""
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

from . import da as oda
from prisma_b200.seeded_weights import DA_CONFIGS, ZOE_CONFIG

MEAN = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)  # PrepForMidas (base_models/depth_anything.py:183-184)
STD = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)


def _mlp(sd, p, x, act_last=None):
    """nn.Sequential(Conv1x1, ReLU, Conv1x1[, act]) of Projector / SeedBinRegressorUnnormed / AttractorLayerUnnormed."""
    x = F.relu(F.conv2d(x, sd[p + "._net.0.weight"], sd[p + "._net.0.bias"]))
    x = F.conv2d(x, sd[p + "._net.2.weight"], sd[p + "._net.2.bias"])
    return act_last(x) if act_last else x


def _ac(x, size):
    return F.interpolate(x, size, mode="bilinear", align_corners=True)


def zoe_model(sd, x01, encoder="vitl", taps=None):
    """ZoeDepth.forward (zoedepth_v1.py:127-211) on a 1x3xHxW f32 image in [0, 1] -> 1x1x392x518 metric depth."""
    c = ZOE_CONFIG
    # core.prep: Resize(518, 392, keep_aspect_ratio=False, x14, "minimal") = bilinear(align_corners=True) to the fixed
    # img_size, then ImageNet normalisation (base_models/depth_anything.py:171-189)
    x = (_ac(x01, c["img_size"]) - MEAN) / STD
    core = {k[len("core.core."):]: v for k, v in sd.items() if k.startswith("core.core.")}
    t = {}
    rel_depth = oda.da_model(core, x, encoder, taps=t)  # 1 x 392 x 518
    outconv_activation, btlnck = t["out_conv_act"], t["layer_rn"][3]
    x_blocks = [t["path"][3], t["path"][2], t["path"][1], t["path"][0]]  # r4, r3, r2, r1
    x_d0 = F.conv2d(btlnck, sd["conv2.weight"], sd["conv2.bias"])
    b_prev = _mlp(sd, "seed_bin_regressor", x_d0, F.softplus)  # SeedBinRegressorUnnormed
    prev_emb = _mlp(sd, "seed_projector", x_d0)
    for i, xb in enumerate(x_blocks):
        emb = _mlp(sd, f"projectors.{i}", xb)
        # AttractorLayerUnnormed.forward (attractor.py:165-207), attractor_type "inv", kind "mean"
        A = _mlp(sd, f"attractors.{i}", emb + _ac(prev_emb, emb.shape[-2:]), F.softplus)
        b = _ac(b_prev, A.shape[-2:])
        # The layer calls dist(dx) WITHOUT its alpha / gamma (attractor.py:194-196), so inv_attractor's own defaults
        # (alpha = 300, gamma = 2, attractor.py:46) apply and config attractor_alpha = 1000 is never used; ZoeDepth does
        # not pass memory_efficient either (zoedepth_v1.py:108-112), so the non-looped torch.mean path runs.
        dx = A.unsqueeze(2) - b.unsqueeze(1)
        b_prev = b + torch.mean(dx.div(1 + 300.0 * dx.pow(2)), dim=1)
        prev_emb = emb
    b_centers, b_embedding = b_prev, prev_emb
    last = torch.cat([outconv_activation, _ac(rel_depth.unsqueeze(1), outconv_activation.shape[2:])], dim=1)
    cond = _ac(b_embedding, last.shape[-2:])
    # ConditionalLogBinomial (dist_layers.py:66-108)
    pt = F.conv2d(torch.cat((last, cond), 1), sd["conditional_log_binomial.mlp.0.weight"], sd["conditional_log_binomial.mlp.0.bias"])
    pt = F.softplus(F.conv2d(F.gelu(pt), sd["conditional_log_binomial.mlp.2.weight"], sd["conditional_log_binomial.mlp.2.bias"]))
    p, tt = pt[:, :2] + 1e-4, pt[:, 2:] + 1e-4
    p = p[:, 0] / (p[:, 0] + p[:, 1])
    tt = (tt[:, 0] / (tt[:, 0] + tt[:, 1])).unsqueeze(1)
    tt = (c["max_temp"] - c["min_temp"]) * tt + c["min_temp"]
    # LogBinomial.forward (dist_layers.py:43-63)
    K = c["n_bins"]
    k_idx = torch.arange(0, K).view(1, -1, 1, 1)
    xx = p.unsqueeze(1)
    one_minus = torch.clamp(1 - xx, 1e-4, 1)
    xx = torch.clamp(xx, 1e-4, 1)
    n, k = torch.tensor([K - 1.0]).view(1, -1, 1, 1) + 1e-7, k_idx + 1e-7
    log_binom = n * torch.log(n) - k * torch.log(k) - (n - k) * torch.log(n - k + 1e-7)
    y = log_binom + k_idx * torch.log(xx) + (K - 1 - k_idx) * torch.log(one_minus)
    prob = torch.softmax(y / tt, dim=1)
    centers = _ac(b_centers, prob.shape[-2:])
    out = torch.sum(prob * centers, dim=1, keepdim=True)
    if taps is not None:
        taps.update(net_input=x, rel_depth=rel_depth, b_centers=b_centers, b_embedding=b_embedding, prob=prob, metric=out)
    return out


def zoe_infer(sd, img_u8, encoder="vitl", taps=None):
    """infer(img) with args.metric != 'none' (bands/depth_anything.py:106-119): ToTensor, model, PIL resize of the
    prediction (mode "F", Pillow's default BICUBIC) to the frame size."""
    h, w = img_u8.shape[:2]
    x01 = torch.from_numpy(np.ascontiguousarray(img_u8.transpose(2, 0, 1))).float().div(255).unsqueeze(0)
    with torch.no_grad():
        pred = zoe_model(sd, x01, encoder, taps).squeeze().numpy()
    return np.asarray(Image.fromarray(pred).resize((w, h)))
