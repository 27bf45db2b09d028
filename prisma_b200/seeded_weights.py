"""Deterministic seeded weights for offline runs, the tests and the oracle (no arithmetic of the hot path here).

No checkpoint is reachable offline (weights come from
download_models.sh / HF hub), so parity is asserted on seeded random weights.
Tensor *names and shapes* are exactly the reference ``state_dict`` ones
(verified by ``oracle/tools/make_golden.py`` which
``load_state_dict(strict=True)``s them into the reference modules), so the
product's converter accepts real checkpoints unchanged.

Distributions follow the reference modules' own initialisers (trunc-normal
0.02 ViT linears, torch-default kaiming-uniform convs) but, unlike them, biases, LayerNorm
affine and LayerScale are non-trivial, so a swapped bias or a dropped gamma shows up in parity.  Values depend only on (seed, tensor name) through a CPU
``torch.Generator`` -> identical in the authoring container and on the GPU box.
"""
import zlib

import torch

DA_CONFIGS = {
    # dim, depth, heads: bands/d_anything/torchhub/.../vision_transformer.py:339-378
    # features/out_channels: ViT-L from patchfusion/.../base_models/depth_anything.py:339,
    # ViT-S/B from the upstream HF config.json (not in tree)
    "vits": dict(dim=384, depth=12, heads=6, features=64, out_channels=[48, 96, 192, 384]),
    "vitb": dict(dim=768, depth=12, heads=12, features=128, out_channels=[96, 192, 384, 768]),
    "vitl": dict(dim=1024, depth=24, heads=16, features=256, out_channels=[256, 512, 1024, 1024]),
}


def _gen(name, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(name, seed, shape, std, mean=0.0):
    return torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32) * std + mean


def _uniform(name, seed, shape, lo, hi):
    return torch.rand(shape, generator=_gen(name, seed), dtype=torch.float32) * (hi - lo) + lo


def make_da_weights(encoder="vits", seed=0):
    """state_dict of DPT_DINOv2(encoder) (bands/d_anything/dpt.py:139-153) with seeded values."""
    c = DA_CONFIGS[encoder]
    D, depth, F, oc = c["dim"], c["depth"], c["features"], c["out_channels"]
    sd = {}

    def lin(name, out_f, in_f, bias=True, std=None):
        std = std if std is not None else 0.02
        sd[name + ".weight"] = _normal(name + ".weight", seed, (out_f, in_f), std)
        if bias:
            sd[name + ".bias"] = _normal(name + ".bias", seed, (out_f,), 0.01)

    def conv(name, out_c, in_c, kh, kw, bias=True, transpose=False):
        # torch's default conv initialiser (what the reference modules get): kaiming_uniform(a=sqrt(5))
        # -> U(-b, b), b = 1/sqrt(fan_in); fan_in of a ConvTranspose2d weight [in,out,kh,kw] is out*kh*kw
        fan_in = (out_c if transpose else in_c) * kh * kw
        b = (1.0 / fan_in) ** 0.5
        shape = (in_c, out_c, kh, kw) if transpose else (out_c, in_c, kh, kw)
        sd[name + ".weight"] = _uniform(name + ".weight", seed, shape, -b, b)
        if bias:
            sd[name + ".bias"] = _uniform(name + ".bias", seed, (out_c,), -b, b)

    def ln(name, dim):
        sd[name + ".weight"] = _normal(name + ".weight", seed, (dim,), 0.1, 1.0)
        sd[name + ".bias"] = _normal(name + ".bias", seed, (dim,), 0.05)

    p = "pretrained."
    sd[p + "cls_token"] = _normal(p + "cls_token", seed, (1, 1, D), 0.02)
    sd[p + "pos_embed"] = _normal(p + "pos_embed", seed, (1, 1370, D), 0.02)
    sd[p + "mask_token"] = torch.zeros(1, D)
    sd[p + "patch_embed.proj.weight"] = _normal(p + "patch_embed.proj.weight", seed, (D, 3, 14, 14), 0.02)
    sd[p + "patch_embed.proj.bias"] = _normal(p + "patch_embed.proj.bias", seed, (D,), 0.02)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        ln(b + "norm1", D)
        lin(b + "attn.qkv", 3 * D, D, std=0.04)  # larger than 0.02 so attention is not uniform
        lin(b + "attn.proj", D, D)
        sd[b + "ls1.gamma"] = _uniform(b + "ls1.gamma", seed, (D,), 0.3, 1.0)
        ln(b + "norm2", D)
        lin(b + "mlp.fc1", 4 * D, D)
        lin(b + "mlp.fc2", D, 4 * D)
        sd[b + "ls2.gamma"] = _uniform(b + "ls2.gamma", seed, (D,), 0.3, 1.0)
    ln(p + "norm", D)

    h = "depth_head."
    for i, c_out in enumerate(oc):
        conv(f"{h}projects.{i}", c_out, D, 1, 1)
    conv(h + "resize_layers.0", oc[0], oc[0], 4, 4, transpose=True)
    conv(h + "resize_layers.1", oc[1], oc[1], 2, 2, transpose=True)
    conv(h + "resize_layers.3", oc[3], oc[3], 3, 3)
    for i in range(4):
        conv(f"{h}scratch.layer{i + 1}_rn", F, oc[i], 3, 3, bias=False)
    for i in range(1, 5):
        r = f"{h}scratch.refinenet{i}."
        conv(r + "out_conv", F, F, 1, 1)
        for u in (1, 2):
            conv(f"{r}resConfUnit{u}.conv1", F, F, 3, 3)
            conv(f"{r}resConfUnit{u}.conv2", F, F, 3, 3)
    conv(h + "scratch.output_conv1", F // 2, F, 3, 3)
    conv(h + "scratch.output_conv2.0", 32, F // 2, 3, 3)
    conv(h + "scratch.output_conv2.2", 1, 32, 1, 1)
    # Final bias: positive, so that (as for the trained model and for the reference's default
    # init) the depth map is strictly positive instead of being half
    # clamped to 0 by the last ReLU -- a clamped map makes the band's min/max normalisation
    # degenerate and measures cancellation noise rather than the kernels.
    sd[h + "scratch.output_conv2.2.bias"] = torch.full((1,), 0.25)
    return sd


ZOE_CONFIG = dict(  # patchfusion/zoedepth/models/zoedepth/config_zoedepth.json ("model" section) as read by get_org_config
    n_bins=64, bin_embedding_dim=128, n_attractors=[16, 8, 4, 1], attractor_alpha=1000, attractor_gamma=2, min_temp=0.0212,
    max_temp=50.0, img_size=(392, 518))


def make_zoe_weights(encoder="vitl", seed=0):
    """state_dict of ZoeDepth(DepthAnythingCore(DPT_DINOv2)) (patchfusion/zoedepth/models/zoedepth/zoedepth_v1.py:40-125):
    the relative model under core.core.* plus the metric head.  The reference only instantiates ViT-L; "vits" is the
    same head on the small encoder (btlnck / decoder features 64) for fast pinning."""
    c = DA_CONFIGS[encoder]
    F = c["features"]
    sd = {"core.core." + k: v for k, v in make_da_weights(encoder, seed).items()}

    def conv(name, out_c, in_c, bias_mean=0.0):
        b = (1.0 / in_c) ** 0.5
        sd[name + ".weight"] = _uniform(name + ".weight", seed, (out_c, in_c, 1, 1), -b, b)
        sd[name + ".bias"] = _uniform(name + ".bias", seed, (out_c,), -b, b) + bias_mean

    conv("conv2", F, F)
    conv("seed_bin_regressor._net.0", 256, F)
    conv("seed_bin_regressor._net.2", 64, 256, bias_mean=1.0)
    conv("seed_projector._net.0", 128, F)
    conv("seed_projector._net.2", 128, 128)
    for i, na in enumerate(ZOE_CONFIG["n_attractors"]):
        conv(f"projectors.{i}._net.0", 128, F)
        conv(f"projectors.{i}._net.2", 128, 128)
        conv(f"attractors.{i}._net.0", 128, 128)
        conv(f"attractors.{i}._net.2", na, 128, bias_mean=1.0)
    bott = (33 + 128) // 2
    conv("conditional_log_binomial.mlp.0", bott, 33 + 128)
    conv("conditional_log_binomial.mlp.2", 4, bott)
    sd["conditional_log_binomial.log_binomial_transform.k_idx"] = torch.arange(0, 64).view(1, -1, 1, 1)       # buffers of
    sd["conditional_log_binomial.log_binomial_transform.K_minus_1"] = torch.Tensor([63]).view(1, -1, 1, 1)  # LogBinomial
    # Shape the head like a trained one, so that parity is measured on a depth map with real dynamic range: seed bin
    # centres spread over 0.5..8, a low temperature (peaked distribution over the 64 bins) and wider output weights.
    sd["seed_bin_regressor._net.2.bias"] = sd["seed_bin_regressor._net.2.bias"] + torch.linspace(-0.5, 7.0, 64)
    sd["conditional_log_binomial.mlp.0.weight"] = sd["conditional_log_binomial.mlp.0.weight"] * 3.0
    sd["conditional_log_binomial.mlp.2.weight"] = sd["conditional_log_binomial.mlp.2.weight"] * torch.tensor([24.0, 24.0, 4.0, 4.0]).view(4, 1, 1, 1)
    sd["conditional_log_binomial.mlp.2.bias"] = sd["conditional_log_binomial.mlp.2.bias"] + torch.tensor([0.0, 0.0, -5.0, 2.0])
    for i in range(4):
        sd[f"attractors.{i}._net.2.bias"] = sd[f"attractors.{i}._net.2.bias"] + torch.linspace(0.0, 5.0, ZOE_CONFIG["n_attractors"][i])
    return sd


MIDAS_CONFIGS = {
    # MiDaS v3 DPT_Large (hub "intel-isl/MiDaS": DPTDepthModel(backbone="vitl16_384"); timm vit_large_patch16_384):
    # hooks/features/reassemble channels from midas/dpt_depth.py + midas/backbones/vit.py (un-vendored)
    "dpt_large": dict(dim=1024, depth=24, heads=16, hooks=[5, 11, 17, 23], features=256, out_channels=[256, 512, 1024, 1024]),
    # test-size twin of the same graph (CPU oracle in seconds)
    "dpt_tiny": dict(dim=384, depth=8, heads=6, hooks=[1, 3, 5, 7], features=64, out_channels=[48, 96, 192, 384]),
}


def make_midas_weights(variant="dpt_large", seed=0):
    """state_dict of MiDaS DPTDepthModel with the upstream tensor names (pretrained.model.* = timm ViT/16,
    pretrained.act_postprocessN.* = readout / 1x1 / resize, scratch.* = RefineNet fusion + output head)."""
    c = MIDAS_CONFIGS[variant]
    D, depth, F, oc = c["dim"], c["depth"], c["features"], c["out_channels"]
    sd = {}

    def lin(name, out_f, in_f, std=0.02):
        sd[name + ".weight"] = _normal(name + ".weight", seed, (out_f, in_f), std)
        sd[name + ".bias"] = _normal(name + ".bias", seed, (out_f,), 0.01)

    def conv(name, out_c, in_c, kh, kw, bias=True, transpose=False):
        fan_in = (out_c if transpose else in_c) * kh * kw
        b = (1.0 / fan_in) ** 0.5
        shape = (in_c, out_c, kh, kw) if transpose else (out_c, in_c, kh, kw)
        sd[name + ".weight"] = _uniform(name + ".weight", seed, shape, -b, b)
        if bias:
            sd[name + ".bias"] = _uniform(name + ".bias", seed, (out_c,), -b, b)

    def ln(name, dim):
        sd[name + ".weight"] = _normal(name + ".weight", seed, (dim,), 0.1, 1.0)
        sd[name + ".bias"] = _normal(name + ".bias", seed, (dim,), 0.05)

    p = "pretrained.model."
    sd[p + "cls_token"] = _normal(p + "cls_token", seed, (1, 1, D), 0.02)
    sd[p + "pos_embed"] = _normal(p + "pos_embed", seed, (1, 24 * 24 + 1, D), 0.02)
    sd[p + "patch_embed.proj.weight"] = _normal(p + "patch_embed.proj.weight", seed, (D, 3, 16, 16), 0.02)
    sd[p + "patch_embed.proj.bias"] = _normal(p + "patch_embed.proj.bias", seed, (D,), 0.02)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        ln(b + "norm1", D)
        lin(b + "attn.qkv", 3 * D, D, std=0.04)
        lin(b + "attn.proj", D, D)
        ln(b + "norm2", D)
        lin(b + "mlp.fc1", 4 * D, D)
        lin(b + "mlp.fc2", D, 4 * D)
    ln(p + "norm", D)
    for i, c_out in enumerate(oc):
        a = f"pretrained.act_postprocess{i + 1}."
        lin(a + "0.project.0", D, 2 * D)
        conv(a + "3", c_out, D, 1, 1)
    conv("pretrained.act_postprocess1.4", oc[0], oc[0], 4, 4, transpose=True)
    conv("pretrained.act_postprocess2.4", oc[1], oc[1], 2, 2, transpose=True)
    conv("pretrained.act_postprocess4.4", oc[3], oc[3], 3, 3)
    for i in range(4):
        conv(f"scratch.layer{i + 1}_rn", F, oc[i], 3, 3, bias=False)
    for i in range(1, 5):
        r = f"scratch.refinenet{i}."
        conv(r + "out_conv", F, F, 1, 1)
        for u in (1, 2):
            conv(f"{r}resConfUnit{u}.conv1", F, F, 3, 3)
            conv(f"{r}resConfUnit{u}.conv2", F, F, 3, 3)
    conv("scratch.output_conv.0", F // 2, F, 3, 3)
    conv("scratch.output_conv.2", 32, F // 2, 3, 3)
    conv("scratch.output_conv.4", 1, 32, 1, 1)
    sd["scratch.output_conv.4.bias"] = torch.full((1,), 0.25)  # positive map, see make_da_weights
    return sd


SOLO_CONFIGS = {
    # models/solov2_r101_fpn_3x_coco.py (downloaded config, not in tree; values from upstream mmdet 2.x):
    # ResNet-101 (3,4,23,3), FPN 256 x 5 levels, SOLOV2Head(feat 512, 4 stacked convs, grids 40/36/24/16/12,
    # strides 8/8/16/32/32, mask feature head 128 -> 256 at stride 4, GN-32), test pipeline scale (1333, 800)
    "r101": dict(layers=[3, 4, 23, 3], img_scale=(1333, 800), num_grids=[40, 36, 24, 16, 12], strides=[8, 8, 16, 32, 32],
                 feat=512, mask_feat=128, mask_out=256, num_classes=80),
    # test-size twin of the same graph (CPU oracle in seconds): one bottleneck per stage, small test scale
    "tiny": dict(layers=[1, 1, 1, 1], img_scale=(448, 256), num_grids=[40, 36, 24, 16, 12], strides=[8, 8, 16, 32, 32],
                 feat=512, mask_feat=128, mask_out=256, num_classes=80),
}


def make_solo_weights(variant="r101", seed=0):
    """state_dict of mmdet SOLOv2 (backbone.* ResNet, neck.* FPN, mask_head.* SOLOV2Head) with seeded values.

    Backbone convs: kaiming-normal(fan_out) as mmdet's ResNet init; BatchNorm (eval, frozen) with non-trivial affine and
    running statistics so the fold is tested.  Head: wider than mmdet's std = 0.01 / bias_prob = 0.01 init so that, with
    random features, a few hundred grid cells pass score_thr, a few dozen end above 0.5 and the dynamic-conv masks are
    structured blobs instead of sigmoid(0) noise -- the decode path (NMS, thresholds, resizes) is then exercised."""
    c = SOLO_CONFIGS[variant]
    sd = {}

    def conv(name, out_c, in_c, k, std=None, bias=None):
        std = std if std is not None else (2.0 / (out_c * k * k)) ** 0.5
        sd[name + ".weight"] = _normal(name + ".weight", seed, (out_c, in_c, k, k), std)
        if bias is not None:
            sd[name + ".bias"] = _normal(name + ".bias", seed, (out_c,), 0.02, bias)

    def bn(name, ch, gain=1.0):
        sd[name + ".weight"] = _normal(name + ".weight", seed, (ch,), 0.1, gain)
        sd[name + ".bias"] = _normal(name + ".bias", seed, (ch,), 0.05)
        sd[name + ".running_mean"] = _normal(name + ".running_mean", seed, (ch,), 0.1)
        sd[name + ".running_var"] = _uniform(name + ".running_var", seed, (ch,), 0.5, 1.5)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def gn(name, ch):
        sd[name + ".weight"] = _normal(name + ".weight", seed, (ch,), 0.1, 1.0)
        sd[name + ".bias"] = _normal(name + ".bias", seed, (ch,), 0.05)

    conv("backbone.conv1", 64, 3, 7)
    bn("backbone.bn1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip([64, 128, 256, 512], c["layers"])):
        for b in range(blocks):
            p = f"backbone.layer{li + 1}.{b}."
            conv(p + "conv1", planes, inplanes, 1)
            bn(p + "bn1", planes)
            conv(p + "conv2", planes, planes, 3)
            bn(p + "bn2", planes)
            conv(p + "conv3", planes * 4, planes, 1)
            bn(p + "bn3", planes * 4, gain=0.5)  # keeps the residual stream bounded over 33 blocks
            if b == 0:
                conv(p + "downsample.0", planes * 4, inplanes, 1)
                bn(p + "downsample.1", planes * 4, gain=0.5)
            inplanes = planes * 4
    for i, cin in enumerate([256, 512, 1024, 2048]):
        conv(f"neck.lateral_convs.{i}.conv", 256, cin, 1, std=(1.0 / cin) ** 0.5, bias=0.0)
        conv(f"neck.fpn_convs.{i}.conv", 256, 256, 3, std=(1.0 / (256 * 9)) ** 0.5, bias=0.0)
    m = "mask_head.mask_feature_head."
    F = c["mask_feat"]
    for i in range(4):
        for j in range(max(i, 1)):
            cin = (256 + (2 if i == 3 else 0)) if j == 0 else F
            conv(f"{m}convs_all_levels.{i}.conv{j}.conv", F, cin, 3, std=(2.0 / (cin * 9)) ** 0.5)
            gn(f"{m}convs_all_levels.{i}.conv{j}.gn", F)
    conv(m + "conv_pred.conv", c["mask_out"], F, 1, std=(2.0 / F) ** 0.5)
    gn(m + "conv_pred.gn", c["mask_out"])
    h = "mask_head."
    for i in range(4):
        conv(f"{h}kernel_convs.{i}.conv", c["feat"], 258 if i == 0 else c["feat"], 3, std=(2.0 / ((258 if i == 0 else c["feat"]) * 9)) ** 0.5)
        gn(f"{h}kernel_convs.{i}.gn", c["feat"])
        conv(f"{h}cls_convs.{i}.conv", c["feat"], 256 if i == 0 else c["feat"], 3, std=(2.0 / ((256 if i == 0 else c["feat"]) * 9)) ** 0.5)
        gn(f"{h}cls_convs.{i}.gn", c["feat"])
    conv(h + "conv_cls", c["num_classes"], c["feat"], 3, std=0.045, bias=-8.0)
    conv(h + "conv_kernel", c["mask_out"], c["feat"], 3, std=0.006, bias=0.0)
    return sd


def make_raft_weights(seed=0):
    """state_dict of RAFT(args) (bands/raft/raft.py:24-57), seeded.

    Convs: kaiming_normal(fan_out, relu) as BasicEncoder's own init loop (raft/extractor.py:148-150) for the
    encoders, torch-default kaiming-uniform for the update block (which has no init loop); biases torch-default
    uniform.  cnet BatchNorm gets non-trivial affine + running statistics so that the eval-mode fold is tested.
    """
    sd = {}

    def conv(name, out_c, in_c, kh, kw, mode):
        fan_in = in_c * kh * kw
        if mode == "kaiming_out":
            std = (2.0 / (out_c * kh * kw)) ** 0.5
            sd[name + ".weight"] = _normal(name + ".weight", seed, (out_c, in_c, kh, kw), std)
        else:
            b = (1.0 / fan_in) ** 0.5
            sd[name + ".weight"] = _uniform(name + ".weight", seed, (out_c, in_c, kh, kw), -b, b)
        b = (1.0 / fan_in) ** 0.5
        sd[name + ".bias"] = _uniform(name + ".bias", seed, (out_c,), -b, b)

    def bn(name, c):
        sd[name + ".weight"] = _normal(name + ".weight", seed, (c,), 0.1, 1.0)
        sd[name + ".bias"] = _normal(name + ".bias", seed, (c,), 0.05)
        sd[name + ".running_mean"] = _normal(name + ".running_mean", seed, (c,), 0.1)
        sd[name + ".running_var"] = _uniform(name + ".running_var", seed, (c,), 0.5, 1.5)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for net, out_dim, has_bn in (("fnet", 256, False), ("cnet", 256, True)):
        p = net + "."
        conv(p + "conv1", 64, 3, 7, 7, "kaiming_out")
        if has_bn:
            bn(p + "norm1", 64)
        cin = 64
        for li, dim, stride in ((1, 64, 1), (2, 96, 2), (3, 128, 2)):
            for bi in (0, 1):
                q = f"{p}layer{li}.{bi}."
                conv(q + "conv1", dim, cin if bi == 0 else dim, 3, 3, "kaiming_out")
                conv(q + "conv2", dim, dim, 3, 3, "kaiming_out")
                if has_bn:
                    bn(q + "norm1", dim)
                    bn(q + "norm2", dim)
                if bi == 0 and stride != 1:
                    conv(q + "downsample.0", dim, cin, 1, 1, "kaiming_out")
                    if has_bn:
                        bn(q + "norm3", dim)
                        for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                            sd[q + "downsample.1." + k] = sd[q + "norm3." + k]   # same module, two names
            cin = dim
        conv(p + "conv2", out_dim, 128, 1, 1, "kaiming_out")
    u = "update_block."
    conv(u + "encoder.convc1", 256, 324, 1, 1, "default")
    conv(u + "encoder.convc2", 192, 256, 3, 3, "default")
    conv(u + "encoder.convf1", 128, 2, 7, 7, "default")
    conv(u + "encoder.convf2", 64, 128, 3, 3, "default")
    conv(u + "encoder.conv", 126, 256, 3, 3, "default")
    for g in ("z", "r", "q"):
        conv(u + f"gru.conv{g}1", 128, 384, 1, 5, "default")
        conv(u + f"gru.conv{g}2", 128, 384, 5, 1, "default")
    conv(u + "flow_head.conv1", 256, 128, 3, 3, "default")
    conv(u + "flow_head.conv2", 2, 256, 3, 3, "default")
    conv(u + "mask.0", 256, 128, 3, 3, "default")
    conv(u + "mask.2", 576, 256, 1, 1, "default")
    return sd
