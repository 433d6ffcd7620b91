"""CPU tests of the oracle (test infrastructure) and of the host-side mirror of the reference interface."""
import json
import os

import pytest
import torch

from oracle import cruller_ref, vit_timm

HERE = os.path.dirname(os.path.abspath(__file__))


def test_vit_restatement_matches_independent_hf_vit():
    """The timm-ViT restatement is pinned against transformers.ViTModel (same architecture, separate code)."""
    from transformers import ViTConfig, ViTModel
    torch.manual_seed(0)
    ours = vit_timm.create_model("vit_test_patch16", in_chans=1, img_size=(64, 48)).eval()
    a = ours.arch
    cfg = ViTConfig(hidden_size=a["embed_dim"], num_hidden_layers=a["depth"], num_attention_heads=a["num_heads"],
                    intermediate_size=int(a["embed_dim"] * a["mlp_ratio"]), hidden_act="gelu",
                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=a["ln_eps"],
                    image_size=(64, 48), patch_size=a["patch_size"], num_channels=1, qkv_bias=True)
    hf = ViTModel(cfg, add_pooling_layer=False).eval()
    sd = ours.state_dict()
    D = a["embed_dim"]
    new = {"embeddings.cls_token": sd["cls_token"], "embeddings.position_embeddings": sd["pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
           "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(a["depth"]):
        p, q = f"blocks.{i}.", f"encoder.layer.{i}."
        w, b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for n, name in enumerate(("query", "key", "value")):       # packed order is [q | k | v]
            new[q + f"attention.attention.{name}.weight"] = w[n * D:(n + 1) * D]
            new[q + f"attention.attention.{name}.bias"] = b[n * D:(n + 1) * D]
        new[q + "attention.output.dense.weight"] = sd[p + "attn.proj.weight"]
        new[q + "attention.output.dense.bias"] = sd[p + "attn.proj.bias"]
        new[q + "layernorm_before.weight"] = sd[p + "norm1.weight"]
        new[q + "layernorm_before.bias"] = sd[p + "norm1.bias"]
        new[q + "layernorm_after.weight"] = sd[p + "norm2.weight"]
        new[q + "layernorm_after.bias"] = sd[p + "norm2.bias"]
        new[q + "intermediate.dense.weight"] = sd[p + "mlp.fc1.weight"]
        new[q + "intermediate.dense.bias"] = sd[p + "mlp.fc1.bias"]
        new[q + "output.dense.weight"] = sd[p + "mlp.fc2.weight"]
        new[q + "output.dense.bias"] = sd[p + "mlp.fc2.bias"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    x = torch.randn(2, 1, 64, 48)
    with torch.no_grad():
        y_ours = ours(x)
        y_hf = hf(pixel_values=x).last_hidden_state
    assert y_ours.shape == (2, 13, 128)
    assert torch.allclose(y_ours, y_hf, atol=2e-5, rtol=1e-4)


def test_prenorm_vit_restatement_matches_independent_hf_clip_vision():
    """The pre-norm (CLIP-style: no patch bias, LayerNorm before the blocks) variant of the restatement -- the
    cruller_large encoder -- against transformers.CLIPVisionModel with hidden_act='gelu' (separate code, separate q / k / v
    projections). CLIP's vision tower normalises only the pooled token at the end; timm's forward_features normalises every
    token, so HF's post_layernorm is applied to all of last_hidden_state here."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    torch.manual_seed(0)
    ours = vit_timm.create_model("vit_test_patch14_clip", in_chans=3, img_size=(56, 56)).eval()
    a = ours.arch
    assert a["pre_norm"]
    D = a["embed_dim"]
    cfg = CLIPVisionConfig(hidden_size=D, intermediate_size=int(D * a["mlp_ratio"]), num_hidden_layers=a["depth"],
                           num_attention_heads=a["num_heads"], num_channels=3, image_size=56, patch_size=a["patch_size"],
                           hidden_act="gelu", layer_norm_eps=a["ln_eps"], attention_dropout=0.0)
    hf = CLIPVisionModel(cfg).eval()
    sd = ours.state_dict()
    assert "patch_embed.proj.bias" not in sd
    new = {"embeddings.class_embedding": sd["cls_token"].reshape(D), "embeddings.position_embedding.weight": sd["pos_embed"][0],
           "embeddings.patch_embedding.weight": sd["patch_embed.proj.weight"],
           "pre_layrnorm.weight": sd["norm_pre.weight"], "pre_layrnorm.bias": sd["norm_pre.bias"],
           "post_layernorm.weight": sd["norm.weight"], "post_layernorm.bias": sd["norm.bias"]}
    for i in range(a["depth"]):
        p, q = f"blocks.{i}.", f"encoder.layers.{i}."
        w, b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for n, name in enumerate(("q_proj", "k_proj", "v_proj")):
            new[q + f"self_attn.{name}.weight"] = w[n * D:(n + 1) * D]
            new[q + f"self_attn.{name}.bias"] = b[n * D:(n + 1) * D]
        new[q + "self_attn.out_proj.weight"] = sd[p + "attn.proj.weight"]
        new[q + "self_attn.out_proj.bias"] = sd[p + "attn.proj.bias"]
        new[q + "layer_norm1.weight"], new[q + "layer_norm1.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        new[q + "layer_norm2.weight"], new[q + "layer_norm2.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    vm = hf.vision_model
    missing, unexpected = vm.load_state_dict(new, strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not unexpected and not missing, (missing, unexpected)
    x = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        y_ours = ours(x)
        y_hf = vm.post_layernorm(vm(pixel_values=x).last_hidden_state)
    assert y_ours.shape == (2, 17, D)
    assert torch.allclose(y_ours, y_hf, atol=2e-5, rtol=1e-4)


def test_oracle_shapes_and_param_counts():
    m = cruller_ref.build_model("cruller_test", vocab_size=50267, seed=0)
    img = torch.randn(2, 1, 64, 48)
    ids = torch.randint(3, 50265, (2, 9))
    out = m(img, ids)
    assert out["logits"].shape == (2, 9, 50267)
    # tied embedding / lm_head (Appendix A.2)
    t = m.text_decoder.trunk
    assert t.lm_head.weight.data_ptr() == t.model.decoder.embed_tokens.weight.data_ptr()


def test_b200_modules_mirror_reference_state_dict_layout():
    """Same keys and shapes as the reference modules (Appendix A.3): checkpoints are interchangeable."""
    from pixparse_b200 import models
    for name in ("cruller_test", "cruller_test_prenorm"):
        ref = cruller_ref.build_model(name, vocab_size=50267, seed=0)
        cfg = models.get_model_config(name)
        cfg.image_encoder.pretrained = False
        cfg.text_decoder.pretrained = False
        ours = models.Cruller(cfg)
        ours.text_decoder.trunk.resize_token_embeddings(50267)
        sd_ref = ref.state_dict()
        sd_ours = ours.state_dict()
        assert list(sd_ours.keys()) == list(sd_ref.keys())
        for k in sd_ref:
            assert tuple(sd_ours[k].shape) == tuple(sd_ref[k].shape), k
        ours.load_state_dict(sd_ref, strict=True)
        assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]


def test_b200_full_size_configs_match_survey_param_counts():
    from pixparse_b200 import models
    cfg = models.get_model_config("cruller_base")
    cfg.image_encoder.pretrained = False
    cfg.text_decoder.pretrained = False
    with torch.device("meta"):
        m = models.Cruller(cfg)
    n = sum(p.numel() for p in m.parameters())
    # 86.03 M encoder + 77.20 M decoder at V=50265 (163.23 M at V=50267)
    assert abs(n - 163.23e6) < 0.02e6
    assert models.list_models()[:2] == ["cruller_base", "cruller_large"]


def test_product_path_refuses_cpu():
    from pixparse_b200 import models
    cfg = models.get_model_config("cruller_test")
    cfg.image_encoder.pretrained = False
    cfg.text_decoder.pretrained = False
    m = models.Cruller(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 1, 64, 48), torch.randint(3, 100, (1, 5)))
