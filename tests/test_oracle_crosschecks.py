"""Independent cross-checks of the two UNPINNED oracles (oracle/solo.py, oracle/midas.py).

Their reference code cannot run here (mmcv / the MiDaS hub + timm are absent), so parity with the reference itself stays
unpinned.  What can be pinned is that the restated building blocks equal widely used independent implementations of the
same published architectures that ARE installed: torchvision's ResNet bottleneck stack and FPN (mmdet's ResNet
style="pytorch" and FPN are that topology, same tensor names) and torchvision's ViT encoder block (= timm's Block).
CPU only."""
import numpy as np
import torch
import torchvision

from oracle import midas as omidas
from oracle import solo as osolo
from prisma_b200.seeded_weights import make_midas_weights, make_solo_weights


def test_solo_backbone_equals_torchvision_resnet():
    from torchvision.models.resnet import Bottleneck, ResNet
    sd = make_solo_weights("tiny", 0)
    net = ResNet(Bottleneck, [1, 1, 1, 1]).eval()
    tv = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    missing = net.load_state_dict(tv, strict=False)
    assert set(missing.missing_keys) == {"fc.weight", "fc.bias"} and not missing.unexpected_keys   # same names as mmdet's
    x = torch.randn(1, 3, 96, 128)
    with torch.no_grad():
        ref, t = [], net.maxpool(net.relu(net.bn1(net.conv1(x))))
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            t = layer(t)
            ref.append(t)
        got = osolo.resnet(sd, x, [1, 1, 1, 1])
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), float((a - b).abs().max())


def test_solo_fpn_equals_torchvision_fpn():
    from torchvision.ops import FeaturePyramidNetwork
    from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
    sd = make_solo_weights("tiny", 0)
    fpn = FeaturePyramidNetwork([256, 512, 1024, 2048], 256, extra_blocks=LastLevelMaxPool()).eval()
    mapped = {}
    for i in range(4):
        mapped[f"inner_blocks.{i}.0.weight"] = sd[f"neck.lateral_convs.{i}.conv.weight"]
        mapped[f"inner_blocks.{i}.0.bias"] = sd[f"neck.lateral_convs.{i}.conv.bias"]
        mapped[f"layer_blocks.{i}.0.weight"] = sd[f"neck.fpn_convs.{i}.conv.weight"]
        mapped[f"layer_blocks.{i}.0.bias"] = sd[f"neck.fpn_convs.{i}.conv.bias"]
    fpn.load_state_dict(mapped, strict=True)
    feats = [torch.randn(1, c, s, s * 2) for c, s in ((256, 24), (512, 12), (1024, 6), (2048, 3))]
    with torch.no_grad():
        ref = list(fpn({str(i): f for i, f in enumerate(feats)}).values())
        got = osolo.fpn(sd, feats)
    assert len(ref) == len(got) == 5
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), float((a - b).abs().max())


def test_midas_vit_block_equals_torchvision_encoder_block():
    from torchvision.models.vision_transformer import EncoderBlock
    sd = make_midas_weights("dpt_tiny", 0)
    D, heads = 384, 6
    blk = EncoderBlock(heads, D, 4 * D, 0.0, 0.0, norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).eval()
    p = "pretrained.model.blocks.0."
    mapped = {"ln_1.weight": sd[p + "norm1.weight"], "ln_1.bias": sd[p + "norm1.bias"],
              "self_attention.in_proj_weight": sd[p + "attn.qkv.weight"], "self_attention.in_proj_bias": sd[p + "attn.qkv.bias"],
              "self_attention.out_proj.weight": sd[p + "attn.proj.weight"], "self_attention.out_proj.bias": sd[p + "attn.proj.bias"],
              "ln_2.weight": sd[p + "norm2.weight"], "ln_2.bias": sd[p + "norm2.bias"],
              "mlp.0.weight": sd[p + "mlp.fc1.weight"], "mlp.0.bias": sd[p + "mlp.fc1.bias"],
              "mlp.3.weight": sd[p + "mlp.fc2.weight"], "mlp.3.bias": sd[p + "mlp.fc2.bias"]}
    blk.load_state_dict(mapped, strict=True)
    x = torch.randn(2, 50, D)
    with torch.no_grad():
        ref = blk(x)
        got = omidas._block(sd, p, x, heads, lambda t: t)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), float((got - ref).abs().max())
