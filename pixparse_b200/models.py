"""Host-side mirror of the reference's model interface for the Cruller hot path.

Same constructor arguments, module attribute names, ``forward`` signatures and ``state_dict`` key layout as

    pixparse.models.Cruller            /root/reference/src/pixparse/models/cruller.py:8-21
    pixparse.models.ImageEncoderTimm   /root/reference/src/pixparse/models/image_encoder_timm.py:28-42
    pixparse.models.TextDecoderHf      /root/reference/src/pixparse/models/text_decoder_hf.py:40-103
    pixparse.models.config.*           /root/reference/src/pixparse/models/config.py:15-67

but the modules below are only *parameter containers*: all arithmetic runs in the hand-written sm_100a kernels
driven by :mod:`pixparse_b200.engine` (there is no PyTorch / CPU fallback for the compute).
"""
import copy
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn

# ------------------------------------------------------------------------------------------------------------------
# configs (models/config.py:15-34 and models/configs/*.json)
# ------------------------------------------------------------------------------------------------------------------


@dataclass
class ImageEncoderCfg:
    name: str = 'vit_base_patch16_224'
    image_fmt: str = 'L'
    image_size: Optional[Tuple[int, int]] = (576, 448)
    pretrained: bool = True


@dataclass
class TextDecoderCfg:
    name: str = 'facebook/bart-base'
    pretrained: bool = True
    num_decoder_layers: Optional[int] = 4
    max_length: Optional[int] = 1024
    pad_token_id: Optional[int] = None


@dataclass
class ModelCfg:
    image_encoder: ImageEncoderCfg = field(default_factory=ImageEncoderCfg)
    text_decoder: TextDecoderCfg = field(default_factory=TextDecoderCfg)


_MODEL_CONFIGS = {
    "cruller_base": ModelCfg(),
    "cruller_large": ModelCfg(
        ImageEncoderCfg(name='vit_large_patch14_clip_224.datacompxl', image_size=(798, 616)),
        TextDecoderCfg(name='facebook/bart-large', num_decoder_layers=10)),
    # named by the reference README (README.md:53) without a json; defined explicitly here
    "cruller_large_6layers": ModelCfg(
        ImageEncoderCfg(name='vit_large_patch14_clip_224.datacompxl', image_size=(798, 616)),
        TextDecoderCfg(name='facebook/bart-large', num_decoder_layers=6)),
    # small models for tests
    "cruller_test": ModelCfg(
        ImageEncoderCfg(name='vit_test_patch16', image_size=(64, 48)),
        TextDecoderCfg(name='test/bart-tiny', num_decoder_layers=2)),
    "cruller_test_prenorm": ModelCfg(
        ImageEncoderCfg(name='vit_test_patch14_clip', image_size=(56, 42)),
        TextDecoderCfg(name='test/bart-tiny', num_decoder_layers=2)),
}


def list_models():
    return list(_MODEL_CONFIGS.keys())


def get_model_config(model_name):
    cfg = _MODEL_CONFIGS.get(model_name.replace('-', '_').lower(), None)
    return copy.deepcopy(cfg)


# timm registry entries the configs name (architecture constants: SURVEY Appendix A.1)
VIT_ARCHS = {
    "vit_base_patch16_224": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, pre_norm=False,
                                 ln_eps=1e-6, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)),
    "vit_large_patch14_clip_224": dict(patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                                       pre_norm=True, ln_eps=1e-5, mean=(0.48145466, 0.4578275, 0.40821073),
                                       std=(0.26862954, 0.26130258, 0.27577711)),
    "vit_test_patch16": dict(patch_size=16, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4.0, pre_norm=False,
                             ln_eps=1e-6, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)),
    "vit_test_patch14_clip": dict(patch_size=14, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4.0, pre_norm=True,
                                  ln_eps=1e-5, mean=(0.48145466, 0.4578275, 0.40821073),
                                  std=(0.26862954, 0.26130258, 0.27577711)),
}

# facebook/bart-* config.json constants. The hub is unreachable from here, so these are restated
# from the public files: both bart-base and bart-large ship dropout = attention_dropout = activation_dropout = 0.1 (the
# 0.0 / 0.0 attention / activation rates belong to the fine-tuned bart-large-cnn / -xsum / -mnli configs and to
# BartConfig's own defaults). When the real config.json is in the local HF cache, `bart_arch()` takes its values instead.
BART_ARCHS = {
    "facebook/bart-base": dict(vocab_size=50265, d_model=768, decoder_layers=6, decoder_attention_heads=12,
                               decoder_ffn_dim=3072, dropout=0.1, attention_dropout=0.1, activation_dropout=0.1,
                               max_position_embeddings=1024, init_std=0.02, pad_token_id=1, bos_token_id=0,
                               eos_token_id=2),
    "facebook/bart-large": dict(vocab_size=50265, d_model=1024, decoder_layers=12, decoder_attention_heads=16,
                                decoder_ffn_dim=4096, dropout=0.1, attention_dropout=0.1, activation_dropout=0.1,
                                max_position_embeddings=1024, init_std=0.02, pad_token_id=1, bos_token_id=0,
                                eos_token_id=2),
    "test/bart-tiny": dict(vocab_size=50265, d_model=128, decoder_layers=2, decoder_attention_heads=2,
                           decoder_ffn_dim=512, dropout=0.1, attention_dropout=0.1, activation_dropout=0.1,
                           max_position_embeddings=1024, init_std=0.02, pad_token_id=1, bos_token_id=0,
                           eos_token_id=2),
}


_BART_KEYS = ("vocab_size", "d_model", "decoder_layers", "decoder_attention_heads", "decoder_ffn_dim", "dropout",
              "attention_dropout", "activation_dropout", "max_position_embeddings", "init_std", "pad_token_id",
              "bos_token_id", "eos_token_id")


def bart_arch(name):
    """Architecture / dropout constants of a BART checkpoint name: the real config.json when transformers finds it in the
    local cache (AutoConfig.from_pretrained, as text_decoder_hf.py:13 does -- never the network), else the restated table."""
    arch = dict(BART_ARCHS[name]) if name in BART_ARCHS else None
    if not name.startswith("test/"):
        try:
            from transformers import AutoConfig
            hf = AutoConfig.from_pretrained(name, local_files_only=True)
            arch = {k: getattr(hf, k) for k in _BART_KEYS}
        except Exception:      # offline and not cached: keep the restated constants
            pass
    if arch is None:
        raise ValueError(f"unsupported text decoder {name!r}; supported offline: {sorted(BART_ARCHS)}")
    return arch


# ------------------------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------------------------
class ParamLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None


class ParamLayerNorm(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class ParamConv(nn.Module):
    def __init__(self, in_chans, out_chans, k, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_chans, in_chans, k, k))
        self.bias = nn.Parameter(torch.zeros(out_chans)) if bias else None


class ParamEmbedding(nn.Module):
    def __init__(self, num, dim, padding_idx=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim, self.padding_idx = num, dim, padding_idx
        self.weight = nn.Parameter(torch.empty(num, dim))


def _trunc_normal_(t, std):
    nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)


# ---- ViT (timm VisionTransformer layout) ---------------------------------------------------------------------------
class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, in_chans, dim, bias):
        super().__init__()
        self.img_size = tuple(img_size)
        self.patch_size = (patch, patch)
        self.grid_size = (img_size[0] // patch, img_size[1] // patch)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = ParamConv(in_chans, dim, patch, bias=bias)


class _VitAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = ParamLinear(dim, dim * 3)
        self.proj = ParamLinear(dim, dim)


class _VitMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = ParamLinear(dim, hidden)
        self.fc2 = ParamLinear(hidden, dim)


class _VitBlock(nn.Module):
    def __init__(self, dim, heads, hidden, eps):
        super().__init__()
        self.norm1 = ParamLayerNorm(dim, eps)
        self.attn = _VitAttention(dim, heads)
        self.norm2 = ParamLayerNorm(dim, eps)
        self.mlp = _VitMlp(dim, hidden)


class VisionTransformerB200(nn.Module):
    """Parameter container with timm's VisionTransformer key names; forward is executed by the engine."""

    def __init__(self, name, in_chans=1, img_size=(224, 224)):
        super().__init__()
        base = name.split('.')[0]
        if base not in VIT_ARCHS:
            raise ValueError(f"unsupported image encoder {name!r}; supported: {sorted(VIT_ARCHS)}")
        a = VIT_ARCHS[base]
        D = a['embed_dim']
        assert D % a['num_heads'] == 0 and D // a['num_heads'] == 64, "the sm_100a attention kernel needs head_dim 64"
        assert img_size[0] % a['patch_size'] == 0 and img_size[1] % a['patch_size'] == 0
        self.arch = dict(a)
        self.in_chans = in_chans
        self.embed_dim = self.num_features = D
        # task code reads trunk.pretrained_cfg['mean'|'std'] (task_cruller_pretrain.py:124-125)
        self.pretrained_cfg = {'mean': a['mean'], 'std': a['std'], 'input_size': (3, 224, 224)}
        self.patch_embed = _PatchEmbed(img_size, a['patch_size'], in_chans, D, bias=not a['pre_norm'])
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, D))
        self.norm_pre = ParamLayerNorm(D, a['ln_eps']) if a['pre_norm'] else nn.Identity()
        self.blocks = nn.Sequential(*[
            _VitBlock(D, a['num_heads'], int(D * a['mlp_ratio']), a['ln_eps']) for _ in range(a['depth'])])
        self.norm = ParamLayerNorm(D, a['ln_eps'])
        self.reset_parameters()

    def reset_parameters(self):
        _trunc_normal_(self.pos_embed, .02)
        nn.init.normal_(self.cls_token, std=1e-6)
        w = self.patch_embed.proj.weight
        fan_in = w.shape[1] * w.shape[2] * w.shape[3]
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))         # nn.Conv2d default
        if self.patch_embed.proj.bias is not None:
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.patch_embed.proj.bias, -bound, bound)
        for m in self.modules():
            if isinstance(m, ParamLinear):
                _trunc_normal_(m.weight, .02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        raise RuntimeError("VisionTransformerB200 holds parameters only; call ImageEncoderTimm / Cruller")


# ---- BART causal decoder (transformers BartForCausalLM layout) -----------------------------------------------------
class _BartAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.k_proj = ParamLinear(dim, dim)
        self.v_proj = ParamLinear(dim, dim)
        self.q_proj = ParamLinear(dim, dim)
        self.out_proj = ParamLinear(dim, dim)


class _BartDecoderLayer(nn.Module):
    def __init__(self, dim, heads, ffn):
        super().__init__()
        self.self_attn = _BartAttention(dim, heads)
        self.self_attn_layer_norm = ParamLayerNorm(dim, 1e-5)
        self.encoder_attn = _BartAttention(dim, heads)
        self.encoder_attn_layer_norm = ParamLayerNorm(dim, 1e-5)
        self.fc1 = ParamLinear(dim, ffn)
        self.fc2 = ParamLinear(ffn, dim)
        self.final_layer_norm = ParamLayerNorm(dim, 1e-5)


class _BartDecoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg['d_model']
        self.embed_tokens = ParamEmbedding(cfg['vocab_size'], D, padding_idx=cfg['pad_token_id'])
        # BartLearnedPositionalEmbedding: offset 2 -> max_position_embeddings + 2 rows
        self.embed_positions = ParamEmbedding(cfg['max_position_embeddings'] + 2, D)
        self.layers = nn.ModuleList([
            _BartDecoderLayer(D, cfg['decoder_attention_heads'], cfg['decoder_ffn_dim'])
            for _ in range(cfg['decoder_layers'])])
        self.layernorm_embedding = ParamLayerNorm(D, 1e-5)


class _BartDecoderWrapper(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.decoder = _BartDecoder(cfg)


class _LmHead(nn.Module):
    """lm_head.weight is the SAME Parameter object as embed_tokens.weight (tie_word_embeddings=True)."""

    def __init__(self, tied_weight):
        super().__init__()
        self.weight = tied_weight


class BartConfigLite:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.add_cross_attention = True
        self.is_decoder = True
        self.is_encoder_decoder = False
        self.tie_word_embeddings = True
        self.use_cache = True


class BartForCausalLMB200(nn.Module):
    """Parameter container with transformers' BartForCausalLM key names (+ resize_token_embeddings)."""

    def __init__(self, name, num_decoder_layers=None, max_length=None):
        super().__init__()
        cfg = bart_arch(name)
        if num_decoder_layers is not None:
            cfg['decoder_layers'] = num_decoder_layers
        if max_length is not None:
            cfg['max_position_embeddings'] = max_length
        assert cfg['d_model'] // cfg['decoder_attention_heads'] == 64, "the sm_100a attention kernel needs head_dim 64"
        self.config = BartConfigLite(**cfg)
        self.model = _BartDecoderWrapper(cfg)
        self.lm_head = _LmHead(self.model.decoder.embed_tokens.weight)
        self.reset_parameters()

    def reset_parameters(self):
        std = self.config.init_std
        for m in self.modules():
            if isinstance(m, ParamLinear):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, ParamEmbedding):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.padding_idx is not None:
                    with torch.no_grad():
                        m.weight[m.padding_idx].zero_()

    def resize_token_embeddings(self, new_num_tokens, mean_resizing=True):
        """transformers PreTrainedModel.resize_token_embeddings: keep old rows, draw the new rows from the old
        rows' mean / covariance (task_cruller_pretrain.py:116 relies on this call existing)."""
        emb = self.model.decoder.embed_tokens
        old = emb.weight.data
        old_n, D = old.shape
        if new_num_tokens == old_n:
            return emb
        new = torch.empty((new_num_tokens, D), dtype=old.dtype, device=old.device)
        n = min(old_n, new_num_tokens)
        new[:n] = old[:n]
        if new_num_tokens > old_n:
            added = new_num_tokens - old_n
            if mean_resizing:
                o32 = old.float()
                mean = o32.mean(0)
                centered = o32 - mean
                cov = centered.T @ centered / old_n
                try:
                    dist = torch.distributions.MultivariateNormal(mean, covariance_matrix=1e-9 * cov)
                    new[old_n:] = dist.sample((added,)).to(old.dtype)
                except Exception:
                    new[old_n:] = mean.to(old.dtype)
            else:
                new[old_n:].normal_(0.0, self.config.init_std)
        p = nn.Parameter(new)
        emb.weight = p
        emb.num_embeddings = new_num_tokens
        self.lm_head.weight = p
        self.config.vocab_size = new_num_tokens
        engine = getattr(self, '_b200_engine_ref', None)
        if engine is not None and engine() is not None:
            engine().invalidate()
        return emb

    def get_input_embeddings(self):
        return self.model.decoder.embed_tokens

    def set_dropout(self, dropout=0.0, attention_dropout=None, activation_dropout=None):
        """Override the dropout rates restated from the public bart config (parity runs use 0)."""
        self.config.dropout = float(dropout)
        self.config.attention_dropout = float(dropout if attention_dropout is None else attention_dropout)
        self.config.activation_dropout = float(dropout if activation_dropout is None else activation_dropout)

    def forward(self, *a, **kw):
        raise RuntimeError("BartForCausalLMB200 holds parameters only; call TextDecoderHf / Cruller")


# ------------------------------------------------------------------------------------------------------------------
# reference-facing modules
# ------------------------------------------------------------------------------------------------------------------
class CausalLMOutput(dict):
    """Mapping with attribute access, like transformers' CausalLMOutputWithCrossAttentions
    (consumers use output['logits'] -- task_cruller_pretrain.py:250 -- and output.logits -- ocr_utils.py:190)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def create_image_encoder(cfg: ImageEncoderCfg):
    assert cfg.name
    assert cfg.image_fmt in ('L', 'RGB')
    if cfg.pretrained:
        raise RuntimeError("pretrained timm weights are not available offline; load a checkpoint with load_state_dict")
    kw = {}
    if cfg.image_size is not None:
        kw['img_size'] = tuple(cfg.image_size)
    return VisionTransformerB200(cfg.name, in_chans=1 if cfg.image_fmt == 'L' else 3, **kw)


class ImageEncoderTimm(nn.Module):
    def __init__(self, cfg: ImageEncoderCfg):
        super().__init__()
        self.trunk = create_image_encoder(cfg)
        self.pool = None
        self.head = None

    def forward(self, x):
        from .engine import engine_for
        return engine_for(self).encode_images(x)


def create_text_decoder(cfg: TextDecoderCfg):
    assert cfg.name
    if cfg.pretrained:
        raise RuntimeError("pretrained HF weights are not available offline; load a checkpoint with load_state_dict")
    return BartForCausalLMB200(cfg.name, cfg.num_decoder_layers, cfg.max_length)


class TextDecoderHf(nn.Module):
    def __init__(self, cfg: TextDecoderCfg):
        super().__init__()
        self.trunk = create_text_decoder(cfg)
        self.prepare_inputs_for_generation = self.prepare_inputs_for_inference

    def prepare_inputs_for_inference(self, input_ids, encoder_outputs, pad_token_id, past_key_values=None, past=None,
                                     use_cache=None, attention_mask=None):
        if past is not None:
            past_key_values = past
        attention_mask = input_ids.ne(pad_token_id).long()
        if past_key_values is not None:
            input_ids = input_ids[:, -1:]
        return {
            "input_ids": input_ids,
            "attention_mask": attention_mask,
            "past_key_values": past_key_values,
            "use_cache": use_cache,
            "encoder_hidden_states": encoder_outputs,
        }

    def forward(self, input_ids, attention_mask=None, encoder_hidden_states=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        from .engine import engine_for
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are never materialised by the fused kernels")
        eng = engine_for(self)
        if use_cache or past_key_values is not None:
            logits, cache = eng.decode_logits(input_ids, encoder_hidden_states, attention_mask=attention_mask,
                                              past_key_values=past_key_values, use_cache=True)
            return CausalLMOutput(logits=logits, past_key_values=cache)
        logits = eng.decode_logits(input_ids, encoder_hidden_states, attention_mask=attention_mask)
        return CausalLMOutput(logits=logits, past_key_values=None)


class Cruller(nn.Module):
    def __init__(self, cfg: ModelCfg):
        super().__init__()
        self.image_encoder = ImageEncoderTimm(cfg.image_encoder)
        self.text_decoder = TextDecoderHf(cfg.text_decoder)

    def forward(self, image_input, text_input):
        from .engine import engine_for
        logits = engine_for(self).forward_logits(image_input, text_input)
        return CausalLMOutput(logits=logits, past_key_values=None)
