"""ORACLE (test infrastructure only -- never imported by the product path).

CPU/PyTorch restatement of timm's ``VisionTransformer`` exactly as pixparse configures it in
``/root/reference/src/pixparse/models/image_encoder_timm.py:13-20``::

    timm.create_model(name, pretrained=..., in_chans=1|3, num_classes=0, global_pool='', img_size=(H, W))

timm itself is a third-party dependency that is absent from /root/reference (unpinned in
pyproject.toml:32-37) and is not installed in this image, so its published algorithm (timm 0.9.x
``vision_transformer.py``) is restated here from SURVEY.md Appendix A.1. Parameter names follow timm's
state_dict layout because that layout is pixparse's checkpoint contract.

Parity status: there are no golden vectors for this path in the reference (it has no tests). The
restatement is pinned instead against an independent implementation of the same architecture
(``transformers.ViTModel`` with remapped weights, tests/test_oracle.py).
"""
import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

# name -> architecture hyper-parameters (timm model registry entries used by models/configs/*.json)
VIT_ARCHS = {
    # models/configs/cruller_base.json
    "vit_base_patch16_224": dict(
        patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, pre_norm=False, ln_eps=1e-6,
        mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)),
    # models/configs/cruller_large.json (CLIP ViT-L/14: pre-norm, conv without bias, nn.LayerNorm default eps)
    "vit_large_patch14_clip_224": dict(
        patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0, pre_norm=True, ln_eps=1e-5,
        mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)),
    # tiny variants for fast tests (not in timm; same code path)
    "vit_test_patch16": dict(
        patch_size=16, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4.0, pre_norm=False, ln_eps=1e-6,
        mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)),
    "vit_test_patch14_clip": dict(
        patch_size=14, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4.0, pre_norm=True, ln_eps=1e-5,
        mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)),
}


def _arch(name):
    base = name.split(".")[0]  # 'vit_large_patch14_clip_224.datacompxl' -> arch name
    if base not in VIT_ARCHS:
        raise ValueError(f"unknown ViT architecture {name!r}")
    return dict(VIT_ARCHS[base])


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias):
        super().__init__()
        self.img_size = tuple(img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size[0] // patch_size, img_size[1] // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

    def forward(self, x):
        B, C, H, W = x.shape
        assert (H, W) == self.img_size, f"input size {(H, W)} != model img_size {self.img_size}"
        x = self.proj(x)                       # (B, D, gh, gw)
        return x.flatten(2).transpose(1, 2)    # (B, gh*gw, D), row-major over the patch grid


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)     # scale = head_dim ** -0.5, no mask
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()                            # exact erf GELU
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


def _trunc_normal_(t, std):
    nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)


class VisionTransformer(nn.Module):
    """num_classes=0, global_pool='' -> forward returns all tokens (B, 1 + gh*gw, D), CLS first."""

    def __init__(self, name, in_chans=1, img_size=(224, 224)):
        super().__init__()
        a = _arch(name)
        D = a["embed_dim"]
        norm_layer = partial(nn.LayerNorm, eps=a["ln_eps"])
        self.arch = a
        self.embed_dim = self.num_features = D
        self.pretrained_cfg = {"mean": a["mean"], "std": a["std"], "input_size": (3, 224, 224)}
        self.patch_embed = PatchEmbed(img_size, a["patch_size"], in_chans, D, bias=not a["pre_norm"])
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, D) * .02)
        self.norm_pre = norm_layer(D) if a["pre_norm"] else nn.Identity()
        self.blocks = nn.Sequential(*[Block(D, a["num_heads"], a["mlp_ratio"], norm_layer) for _ in range(a["depth"])])
        self.norm = norm_layer(D)
        self.init_weights()

    def init_weights(self):
        # timm default ('' mode): trunc_normal(.02) pos_embed and Linear weights, zero biases, cls ~ N(0, 1e-6)
        _trunc_normal_(self.pos_embed, .02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, .02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + self.pos_embed
        x = self.norm_pre(x)
        x = self.blocks(x)
        return self.norm(x)


def create_model(name, pretrained=False, in_chans=3, num_classes=0, global_pool='', img_size=None, **kwargs):
    """Stand-in for ``timm.create_model`` restricted to what pixparse asks for."""
    assert not pretrained, "oracle ViT: pretrained weights are unavailable offline"
    assert num_classes == 0 and global_pool == '', "pixparse always requests a headless, un-pooled trunk"
    if img_size is None:
        img_size = (224, 224)
    return VisionTransformer(name, in_chans=in_chans, img_size=tuple(img_size))
