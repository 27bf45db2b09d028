"""Independent pin of the UNPINNED oracle/midas.py: the HuggingFace `transformers` DPT (DPTForDepthEstimation, the port of
MiDaS v3 DPT-Large that ships `Intel/dpt-large`) is installed in this image.  With the seeded MiDaS-named weights mapped onto
its parameter names (the inverse of transformers' convert_dpt_to_pytorch.py) and its LayerNorm epsilon set to timm's 1e-6,
the whole network -- patch embed, position-embedding resize, 8/24 blocks, hooks, project readout, reassemble, RefineNet
fusion, head -- must agree with the oracle to fp32 rounding.  The reference's own hub code stays un-runnable (parity with
it unpinned), but two independent restatements of the published architecture now agree.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import midas as omidas
from prisma_b200.seeded_weights import MIDAS_CONFIGS, make_midas_weights

transformers = pytest.importorskip("transformers")


def to_hf(sd, c):
    D, out = c["dim"], {}
    p = "pretrained.model."
    out["dpt.embeddings.cls_token"] = sd[p + "cls_token"]
    out["dpt.embeddings.position_embeddings"] = sd[p + "pos_embed"]
    out["dpt.embeddings.patch_embeddings.projection.weight"] = sd[p + "patch_embed.proj.weight"]
    out["dpt.embeddings.patch_embeddings.projection.bias"] = sd[p + "patch_embed.proj.bias"]
    for i in range(c["depth"]):
        b, h = f"{p}blocks.{i}.", f"dpt.encoder.layer.{i}."
        qkv_w, qkv_b = sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"]
        for j, n in enumerate(("query", "key", "value")):
            out[h + f"attention.attention.{n}.weight"] = qkv_w[j * D:(j + 1) * D]
            out[h + f"attention.attention.{n}.bias"] = qkv_b[j * D:(j + 1) * D]
        out[h + "attention.output.dense.weight"] = sd[b + "attn.proj.weight"]
        out[h + "attention.output.dense.bias"] = sd[b + "attn.proj.bias"]
        for a, z in (("layernorm_before", "norm1"), ("layernorm_after", "norm2"), ("intermediate.dense", "mlp.fc1"), ("output.dense", "mlp.fc2")):
            out[h + a + ".weight"] = sd[b + z + ".weight"]
            out[h + a + ".bias"] = sd[b + z + ".bias"]
    out["dpt.layernorm.weight"], out["dpt.layernorm.bias"] = sd[p + "norm.weight"], sd[p + "norm.bias"]
    for i in range(4):
        a = f"pretrained.act_postprocess{i + 1}."
        out[f"neck.reassemble_stage.readout_projects.{i}.0.weight"] = sd[a + "0.project.0.weight"]
        out[f"neck.reassemble_stage.readout_projects.{i}.0.bias"] = sd[a + "0.project.0.bias"]
        out[f"neck.reassemble_stage.layers.{i}.projection.weight"] = sd[a + "3.weight"]
        out[f"neck.reassemble_stage.layers.{i}.projection.bias"] = sd[a + "3.bias"]
        if i != 2:
            out[f"neck.reassemble_stage.layers.{i}.resize.weight"] = sd[a + "4.weight"]
            out[f"neck.reassemble_stage.layers.{i}.resize.bias"] = sd[a + "4.bias"]
        out[f"neck.convs.{i}.weight"] = sd[f"scratch.layer{i + 1}_rn.weight"]
        r, f = f"scratch.refinenet{4 - i}.", f"neck.fusion_stage.layers.{i}."   # fusion layer 0 is refinenet4
        out[f + "projection.weight"], out[f + "projection.bias"] = sd[r + "out_conv.weight"], sd[r + "out_conv.bias"]
        for u in (1, 2):
            for k in (1, 2):
                out[f + f"residual_layer{u}.convolution{k}.weight"] = sd[r + f"resConfUnit{u}.conv{k}.weight"]
                out[f + f"residual_layer{u}.convolution{k}.bias"] = sd[r + f"resConfUnit{u}.conv{k}.bias"]
    for j in (0, 2, 4):
        out[f"head.head.{j}.weight"] = sd[f"scratch.output_conv.{j}.weight"]
        out[f"head.head.{j}.bias"] = sd[f"scratch.output_conv.{j}.bias"]
    return out


def test_midas_oracle_equals_transformers_dpt():
    from transformers import DPTConfig, DPTForDepthEstimation
    c = MIDAS_CONFIGS["dpt_tiny"]
    cfg = DPTConfig(hidden_size=c["dim"], num_hidden_layers=c["depth"], num_attention_heads=c["heads"], intermediate_size=4 * c["dim"],
                    image_size=384, patch_size=16, backbone_out_indices=c["hooks"], neck_hidden_sizes=c["out_channels"],
                    fusion_hidden_size=c["features"], readout_type="project", reassemble_factors=[4, 2, 1, 0.5], is_hybrid=False,
                    layer_norm_eps=1e-6, qkv_bias=True, hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                    use_batch_norm_in_fusion_residual=False, add_projection=False, head_in_index=-1)
    hf = DPTForDepthEstimation(cfg).eval()
    sd = make_midas_weights("dpt_tiny", 0)
    missing = hf.load_state_dict(to_hf(sd, c), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    g = torch.Generator().manual_seed(0)
    for hw in ((256, 256), (192, 256)):   # non-native grids (16x16, 12x16 instead of 24x24): exercise the pos-embed resize
        x = torch.randn(1, 3, *hw, generator=g)
        with torch.no_grad():
            if hw[0] == hw[1]:
                ref = hf(pixel_values=x).predicted_depth
            else:  # the plain-ViT path of DPTForDepthEstimation.forward assumes a square grid; compose it with the real one
                hs = hf.dpt(x, output_hidden_states=True).hidden_states
                hs = [f for i, f in enumerate(hs[1:]) if i in cfg.backbone_out_indices]
                ref = hf.head(hf.neck(hs, hw[0] // 16, hw[1] // 16))
            got = omidas.midas_model(sd, x, "dpt_tiny")
        assert tuple(ref.shape) == tuple(got.shape) == (1,) + hw
        err = float((got - ref).abs().max() / ref.abs().max())
        assert err < 2e-5, (hw, err)
