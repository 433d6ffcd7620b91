"""ORACLE (test infrastructure only -- never imported by the product path).

Self-contained CPU/PyTorch restatement of the reference's Cruller train step so that it can travel to
the GPU box (where /root/reference does not exist):

  * model     : /root/reference/src/pixparse/models/cruller.py:8-21
                image encoder = oracle.vit_timm (restated timm ViT, image_encoder_timm.py:13-20)
                text decoder  = the INSTALLED transformers BartForCausalLM built from a hand-written
                                BartConfig (text_decoder_hf.py:13-33; the HF hub is unreachable, so the
                                public facebook/bart-base / bart-large config constants are restated below)
  * train step: /root/reference/src/pixparse/task/task_cruller_pretrain.py:236-313 (shift, CE with
                ignore_index=-100, backward, clip_grad_norm_, AdamW, cosine schedule via oracle.timm_helpers)

Parity status: the reference has no tests / golden vectors for this path ("parity unpinned" by the
reference itself). This file is pinned by tests/golden/*.json, produced by oracle/gen_golden.py, which runs
the reference's OWN Task / Cruller / TextDecoderHf code (imported from /root/reference over oracle/ref_shims.py)
on the same seeds and must agree with this restatement bit-for-bit on CPU.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import vit_timm
from .timm_helpers import create_optimizer_v2, create_scheduler_v2, dispatch_clip_grad

# Public HF configs restated (config.json of facebook/bart-base and facebook/bart-large).
BART_CONFIGS = {
    "facebook/bart-base": dict(
        vocab_size=50265, d_model=768, encoder_layers=6, decoder_layers=6, encoder_attention_heads=12,
        decoder_attention_heads=12, encoder_ffn_dim=3072, decoder_ffn_dim=3072, activation_function="gelu",
        dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, max_position_embeddings=1024, init_std=0.02,
        scale_embedding=False, pad_token_id=1, bos_token_id=0, eos_token_id=2, decoder_start_token_id=2),
    "facebook/bart-large": dict(
        vocab_size=50265, d_model=1024, encoder_layers=12, decoder_layers=12, encoder_attention_heads=16,
        decoder_attention_heads=16, encoder_ffn_dim=4096, decoder_ffn_dim=4096, activation_function="gelu",
        dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, max_position_embeddings=1024, init_std=0.02,
        scale_embedding=False, pad_token_id=1, bos_token_id=0, eos_token_id=2, decoder_start_token_id=2),
    # tiny decoder for fast tests (same code path)
    "test/bart-tiny": dict(
        vocab_size=50265, d_model=128, encoder_layers=2, decoder_layers=2, encoder_attention_heads=2,
        decoder_attention_heads=2, encoder_ffn_dim=512, decoder_ffn_dim=512, activation_function="gelu",
        dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, max_position_embeddings=1024, init_std=0.02,
        scale_embedding=False, pad_token_id=1, bos_token_id=0, eos_token_id=2, decoder_start_token_id=2),
}


@dataclass
class ImageEncoderCfg:      # models/config.py:15-20
    name: str = 'vit_base_patch16_224'
    image_fmt: str = 'L'
    image_size: Optional[Tuple[int, int]] = (576, 448)
    pretrained: bool = False


@dataclass
class TextDecoderCfg:       # models/config.py:23-29
    name: str = 'facebook/bart-base'
    pretrained: bool = False
    num_decoder_layers: Optional[int] = 4
    max_length: Optional[int] = 1024
    pad_token_id: Optional[int] = None


@dataclass
class ModelCfg:             # models/config.py:31-34
    image_encoder: ImageEncoderCfg = field(default_factory=ImageEncoderCfg)
    text_decoder: TextDecoderCfg = field(default_factory=TextDecoderCfg)


MODEL_CONFIGS = {
    # models/configs/cruller_base.json
    "cruller_base": ModelCfg(),
    # models/configs/cruller_large.json
    "cruller_large": ModelCfg(
        ImageEncoderCfg(name='vit_large_patch14_clip_224.datacompxl', image_size=(798, 616)),
        TextDecoderCfg(name='facebook/bart-large', num_decoder_layers=10)),
    # README.md:53 names it but ships no json: cruller_large with 6 decoder layers
    "cruller_large_6layers": ModelCfg(
        ImageEncoderCfg(name='vit_large_patch14_clip_224.datacompxl', image_size=(798, 616)),
        TextDecoderCfg(name='facebook/bart-large', num_decoder_layers=6)),
    # tiny test models
    "cruller_test": ModelCfg(
        ImageEncoderCfg(name='vit_test_patch16', image_size=(64, 48)),
        TextDecoderCfg(name='test/bart-tiny', num_decoder_layers=2)),
    "cruller_test_prenorm": ModelCfg(
        ImageEncoderCfg(name='vit_test_patch14_clip', image_size=(56, 42)),
        TextDecoderCfg(name='test/bart-tiny', num_decoder_layers=2)),
}


def bart_config(name, num_decoder_layers=None, max_length=None, dropout_off=False):
    """text_decoder_hf.py:13-22 with AutoConfig.from_pretrained replaced by the restated constants."""
    import transformers
    cfg = transformers.BartConfig(**BART_CONFIGS[name])
    cfg.add_cross_attention = True
    if num_decoder_layers is not None:
        cfg.decoder_layers = num_decoder_layers
    if max_length is not None:
        cfg.max_position_embeddings = max_length
    if dropout_off:
        cfg.dropout = cfg.attention_dropout = cfg.activation_dropout = 0.0
    return cfg


class ImageEncoderTimm(nn.Module):   # models/image_encoder_timm.py:28-42
    def __init__(self, cfg: ImageEncoderCfg):
        super().__init__()
        self.trunk = vit_timm.create_model(
            cfg.name, pretrained=False, in_chans=1 if cfg.image_fmt == 'L' else 3, num_classes=0, global_pool='',
            img_size=cfg.image_size)

    def forward(self, x):
        return self.trunk(x)


class TextDecoderHf(nn.Module):      # models/text_decoder_hf.py:40-103
    def __init__(self, cfg: TextDecoderCfg, dropout_off=True):
        super().__init__()
        import transformers
        config = bart_config(cfg.name, cfg.num_decoder_layers, cfg.max_length, dropout_off=dropout_off)
        self.trunk = transformers.AutoModelForCausalLM.from_config(config)

    def forward(self, input_ids, attention_mask=None, encoder_hidden_states=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        # the reference passes use_cache=None (-> config default True) and throws the cache away; use_cache=False
        # yields identical logits without building it
        return self.trunk(input_ids=input_ids, attention_mask=attention_mask,
                          encoder_hidden_states=encoder_hidden_states, past_key_values=past_key_values,
                          use_cache=False if use_cache is None else use_cache, output_attentions=output_attentions,
                          output_hidden_states=output_hidden_states, return_dict=return_dict)


class Cruller(nn.Module):            # models/cruller.py:8-21
    def __init__(self, cfg: ModelCfg, dropout_off=True):
        super().__init__()
        self.image_encoder = ImageEncoderTimm(cfg.image_encoder)
        self.text_decoder = TextDecoderHf(cfg.text_decoder, dropout_off=dropout_off)

    def forward(self, image_input, text_input):
        encoder_output = self.image_encoder(image_input)
        return self.text_decoder(text_input, encoder_hidden_states=encoder_output, return_dict=True)


def build_model(model_name="cruller_base", vocab_size=50267, seed=0, dropout_off=True):
    """Random-init Cruller as the pretrain task builds it (task_cruller_pretrain.py:112-116)."""
    torch.manual_seed(seed)
    model = Cruller(MODEL_CONFIGS[model_name], dropout_off=dropout_off)
    if vocab_size != model.text_decoder.trunk.config.vocab_size:
        model.text_decoder.trunk.resize_token_embeddings(vocab_size)
    return model


class OracleTrainer:
    """task_cruller_pretrain.py train_setup (:155-224) + train_step (:236-313) without DDP/AMP/monitor."""

    def __init__(self, model, vocab_size, lr=3e-4, betas=(0.9, 0.98), eps=1e-6, clip_grad=1.0, clip_mode='norm',
                 layer_decay=None, num_intervals=100, num_warmup_intervals=5, steps_per_interval=1000,
                 warmup_lr=0.0, grad_accum_steps=1):
        self.model = model
        self.vocab_size = vocab_size
        self.loss = nn.CrossEntropyLoss(ignore_index=-100)
        self.clip_grad, self.clip_mode = clip_grad, clip_mode
        self.accum = grad_accum_steps
        kw = {}
        if betas is not None:
            kw['betas'] = betas
        self.optimizer = create_optimizer_v2(model, 'adamw', lr=lr, eps=eps, layer_decay=layer_decay, **kw)
        self.scheduler, _ = create_scheduler_v2(
            self.optimizer, 'cosine', warmup_lr=warmup_lr, warmup_epochs=num_warmup_intervals,
            num_epochs=num_intervals, step_on_epochs=False, updates_per_epoch=steps_per_interval // grad_accum_steps)
        self.scheduler.step_update(0)
        self.step = 0
        self.interval_batch_idx = 0
        self.optimizer.zero_grad()

    def forward_loss(self, image, text, target):
        text_input = text[:, :-1]
        text_target = target[:, 1:]
        output = self.model(image, text_input)
        logits = output['logits']
        loss = self.loss(logits.reshape(-1, self.vocab_size), text_target.reshape(-1))
        return loss, logits

    def train_step(self, sample, keep_grads=False):
        image, text, target = sample
        need_update = (self.interval_batch_idx + 1) % self.accum == 0
        loss, logits = self.forward_loss(image, text, target)
        if self.accum > 1:
            loss = loss / self.accum
        loss.backward()
        result = {"loss": float(loss.detach()), "logits": logits.detach()}
        self.interval_batch_idx += 1
        if not need_update:
            return result
        params = [p for p in self.model.parameters() if p.grad is not None]
        result["grad_norm"] = float(torch.linalg.vector_norm(
            torch.stack([torch.linalg.vector_norm(p.grad, 2.0) for p in params]), 2.0))
        if keep_grads:
            result["grads"] = {n: p.grad.detach().clone() for n, p in self.model.named_parameters()
                               if p.grad is not None}
        if self.clip_grad is not None:
            dispatch_clip_grad(self.model.parameters(), self.clip_grad, self.clip_mode)
        self.optimizer.step()
        self.step += 1
        self.scheduler.step_update(self.step)
        self.optimizer.zero_grad()
        return result


def greedy_decode_uncached(model, encoder_outputs, start_id, pad_id, eos_id, max_steps):
    """utils/ocr_utils.py:165-197 (get_generated_tokens with use_sample=False): the whole prefix is re-fed
    each step with past_key_values=None; finished rows keep generating; stop when all rows emitted EOS."""
    B = encoder_outputs.shape[0]
    input_ids = torch.full((B, 1), start_id, dtype=torch.long, device=encoder_outputs.device)
    finished = torch.zeros(B, dtype=torch.bool, device=encoder_outputs.device)
    with torch.inference_mode():
        for _ in range(max_steps):
            attention_mask = input_ids.ne(pad_id).long()     # text_decoder_hf.py:68
            out = model.text_decoder(input_ids, attention_mask=attention_mask,
                                     encoder_hidden_states=encoder_outputs, return_dict=True)
            nxt = out.logits[:, -1, :].argmax(1).unsqueeze(-1)
            finished |= nxt.squeeze(-1) == eos_id
            if finished.all():
                break
            input_ids = torch.cat([input_ids, nxt], dim=-1)
    return input_ids
