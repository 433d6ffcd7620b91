"""Cruller pre-training task on the B200-native path.

Drop-in for ``pixparse.task.TaskCrullerPretrain`` (/root/reference/src/pixparse/task/task_cruller_pretrain.py:50-391):
same constructor signature ``(cfg, device_env, monitor)``, same ``train_setup / train_interval_start / train_step /
train_interval_end / state_dict / get_current_lr`` sequence, same attributes the app loop reads. What changes is what
runs underneath ``train_step``:

    reference                                            here
    -------------------------------------------------   ----------------------------------------------------------
    autocast + Cruller.forward (timm / transformers)     engine.forward_backward: tcgen05 GEMM / attention kernels
    nn.CrossEntropyLoss on materialised fp32 logits      one-pass CE kernel writing dlogits in place (bf16)
    scaler.scale(loss).backward() (autograd)             hand-sequenced backward kernels, fp32 grads in a flat arena
    DDP reducer (25 MiB buckets)                         reducer.py: copy-engine reduce-scatter / all-gather of arena ranges under backward
    GradScaler.unscale_ + clip_grad_norm_ + AdamW        grad_norm + fused clip/AdamW/bf16-refresh/zero-grad kernels

Numerics are bf16 operands with fp32 accumulation and fp32 master weights, i.e. what the reference gets from
``--task.dtype bfloat16``; no GradScaler is needed for bf16 (a non-finite gradient norm still skips the update,
which is the observable effect the reference's scaler has).
"""
import logging
from dataclasses import dataclass, field
from functools import partial
from typing import Optional

import torch

from .engine import engine_for
from .framework import DeviceEnv, OptimizationCfg, TaskTrain, TaskTrainCfg
from .models import Cruller, ModelCfg, get_model_config
from .optim import FusedAdamW
from .reducer import make_grad_reducer
from .schedule import create_scheduler
from . import synthetic

_logger = logging.getLogger(__name__)


@dataclass
class TokenizerCfg:
    name: str = 'facebook/bart-large'
    pretrained: bool = True


@dataclass
class TaskCrullerPretrainCfg(TaskTrainCfg):
    model_name: Optional[str] = None
    model: ModelCfg = field(default_factory=ModelCfg)
    tokenizer: TokenizerCfg = field(default_factory=TokenizerCfg)

    def __post_init__(self):
        if self.model_name:
            model = get_model_config(self.model_name)
            if model is None:
                _logger.warning(f'Model config for {self.model_name} was not found, using defaults.')
            else:
                self.model = model
        else:
            self.model_name = 'custom'


def load_tokenizer(cfg: TokenizerCfg):
    """HF AutoTokenizer when its files are reachable (tokenizers/tokenizer_hf.py:6-13); otherwise a tokenizer with
    bart's id layout that can only map the special tokens (enough for synthetic-token training and benchmarks)."""
    try:
        from transformers import AutoTokenizer
        return AutoTokenizer.from_pretrained(cfg.name)
    except Exception as e:   # offline hub
        _logger.warning(f"tokenizer {cfg.name} unavailable ({type(e).__name__}); using the synthetic bart-layout tokenizer")
        return synthetic.SyntheticBartTokenizer()


class _TokenizerHolder:
    """Stand-in for pixparse.tokenizers.TokenizerHF: exposes ``.trunk``."""

    def __init__(self, trunk):
        self.trunk = trunk


class TaskCrullerPretrain(TaskTrain):
    def __init__(self, cfg: TaskCrullerPretrainCfg, device_env: DeviceEnv, monitor=None, tokenizer=None):
        super().__init__(cfg=cfg, device_env=device_env, monitor=monitor)
        self.cfg = cfg
        self.amp_dtype = None
        if cfg.dtype is not None:
            self.amp_dtype = torch.bfloat16 if cfg.dtype in ('bfloat16', 'bf16') else torch.float16
        if cfg.amp and self.amp_dtype is torch.float16:
            raise ValueError("the B200 path computes in bf16; fp16 autocast + loss scaling is not implemented")

        self.task_start_token = '<s_pretrain>'
        self.prompt_end_token = self.task_start_token
        self.max_position_embeddings = cfg.model.text_decoder.max_length
        self.text_anno_fn = False
        self.tokenizer = _TokenizerHolder(tokenizer if tokenizer is not None else load_tokenizer(cfg.tokenizer))

        special_tokens = ["<sep/>", self.task_start_token, self.prompt_end_token]
        newly_added_num = self.tokenizer.trunk.add_special_tokens(
            {"additional_special_tokens": sorted(set(special_tokens))})
        self.vocab_size = len(self.tokenizer.trunk)

        # text_anno_fn=False (pre-training on OCR annotations) selects the page-sampling OCR preprocessor, exactly as the
        # reference wires it (task_cruller_pretrain.py:102-109)
        self.anno_preprocess_train = partial(
            preprocess_text_tokens if self.text_anno_fn else preprocess_ocr_anno, tokenizer=self.tokenizer.trunk,
            max_position_embeddings=self.max_position_embeddings, task_start_token=self.task_start_token,
            prompt_end_token=self.prompt_end_token)

        cfg.model.image_encoder.pretrained = False     # no hub access: weights come from load_state_dict
        cfg.model.text_decoder.pretrained = False
        self.model = Cruller(cfg.model)
        if newly_added_num > 0:
            self.model.text_decoder.trunk.resize_token_embeddings(len(self.tokenizer.trunk))

        self.has_no_sync = False
        self.num_image_chs = 1 if cfg.model.image_encoder.image_fmt == 'L' else 3
        img_mean = self.model.image_encoder.trunk.pretrained_cfg['mean']
        img_std = self.model.image_encoder.trunk.pretrained_cfg['std']
        self.img_mean = sum(img_mean) / len(img_mean) if cfg.model.image_encoder.image_fmt == 'L' else img_mean
        self.img_std = sum(img_std) / len(img_std) if cfg.model.image_encoder.image_fmt == 'L' else img_std
        self.image_preprocess_train = build_image_preprocess(
            cfg.model.image_encoder.image_size, self.img_mean, self.img_std)
        self.image_preprocess_eval = None
        self.train_metrics = {}
        self.eval_metrics = {}
        self.max_recursion_length = 1000
        self._init_step_state()

    # ------------------------------------------------------------------------------------------------------------
    def train_setup(self, num_batches_per_interval: int):
        device = self.device_env.device
        self.model.to(device)
        self.engine = engine_for(self.model)
        arena = self.engine.ensure_bound()

        if self.device_env.world_size > 1:
            # all ranks start from rank 0's weights (DDP broadcasts parameters at construction)
            torch.distributed.broadcast(arena.p32, src=0)
            self.reducer = make_grad_reducer(arena)
            self.has_no_sync = True

            def _ready(first_key, last_key, _ar=arena, _r=self.reducer):
                lo = _ar.index[first_key][0]
                o, n, _ = _ar.index[last_key]
                _r.range_ready(lo, o + (n + 63) // 64 * 64)
            if self.reducer is not None:
                self.engine._grad_ready_hook = _ready

        opt = self.cfg.opt
        if opt.optimizer != 'adamw':
            raise ValueError("only 'adamw' is on the Cruller hot path (framework/config.py:8)")
        kw = {}
        if opt.betas is not None:
            kw['betas'] = tuple(opt.betas)
        self.optimizer = FusedAdamW(self.model, self.engine, lr=opt.learning_rate, eps=opt.eps,
                                    layer_decay=opt.layer_decay, **kw)
        self.scaler = None          # bf16: no loss scaling
        self.autocast = None
        if opt.clip_grad_value is not None and (opt.clip_grad_mode or 'norm') != 'norm':
            raise ValueError("only clip_grad_mode='norm' is fused on the B200 path")

        self.num_steps_per_interval = num_batches_per_interval // opt.grad_accum_steps
        self.scheduler, _ = create_scheduler(
            self.optimizer, opt.scheduler, warmup_lr=opt.warmup_learning_rate,
            warmup_intervals=self.num_warmup_intervals, num_intervals=self.num_intervals,
            updates_per_interval=self.num_steps_per_interval)
        self.scheduler.step_update(0)

    def train_interval_start(self):
        self.engine.zero_grads()
        self.interval_batch_idx = 0

    def train_interval_end(self):
        if self.monitor is not None:
            self.monitor.log_phase('train', self.interval_idx)
        self.interval_idx += 1

    def _init_step_state(self):
        """Per-task state of the host -> device copy and loss read-back machinery (shared with the finetune task)."""
        self.engine = None
        self.reducer = None
        self.last_loss = None      # device tensor [n_valid, mean_loss] of the latest micro-step
        self._copy_stream = None   # host batch -> device copies (train_step)
        self._loss_stream = None   # asynchronous loss read-back (last_loss_value)
        self._loss_host = None
        self._loss_event = None

    def _to_device(self, *tensors):
        """Batch tensors -> device. Pinned host tensors are copied on a side stream, so the transfer runs under the tail
        of the previous step (the host thread is ahead of the GPU) and the compute stream only waits for its event;
        batches staged ahead by ``data.DevicePrefetcher`` arrive as device tensors and pass through."""
        device = self.device_env.device
        if device.type == 'cuda' and all(t.device.type == 'cpu' and t.is_pinned() for t in tensors):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=device)
            main = torch.cuda.current_stream(device)
            with torch.cuda.stream(self._copy_stream):
                out = [t.to(device, non_blocking=True) for t in tensors]
            main.wait_stream(self._copy_stream)
            for t in out:
                t.record_stream(main)
            return out
        return [t.to(device, non_blocking=True) for t in tensors]

    def train_step(self, sample):
        image_input, text_input, text_target = sample
        result = {}
        device = self.device_env.device
        image_input, text_input, text_target = self._to_device(image_input, text_input, text_target)
        text_input = text_input[:, :-1].contiguous()
        text_target = text_target[:, 1:].contiguous()

        accum_steps = self.cfg.opt.grad_accum_steps
        need_update = (self.interval_batch_idx + 1) % accum_steps == 0

        if self.reducer is not None:
            self.reducer.enabled = need_update      # == model.no_sync() on accumulation micro-steps
            self.reducer.begin()
        if device.type == 'cuda' and self.engine.on_loss_ready is None:
            self.engine.on_loss_ready = self._stage_loss_readback
        self.last_loss = self.engine.forward_backward(image_input, text_input, text_target,
                                                      grad_scale=1.0 / accum_steps)
        if self.reducer is not None:
            self.reducer.finish()

        self.batch_idx += 1
        self.interval_batch_idx += 1
        if not need_update:
            return result

        self.optimizer.step(clip_grad_norm=self.cfg.opt.clip_grad_value)
        self.step += 1
        self.scheduler.step_update(self.step)
        self.optimizer.zero_grad()

        if self.step % self.eval_frequency == 0 and self.monitor is not None:
            self.monitor.log_step(
                'train', step_idx=self.step, step_end_idx=self.num_intervals * self.num_steps_per_interval,
                interval=self.interval_idx, loss=self.last_loss_value() / accum_steps,
                lr=self.get_current_lr(), metrics=self.train_metrics, eval_data=None)
        return result

    # ---- loss read-back without draining the step -------------------------------------------------------------
    # The CE kernel produces [n_valid, mean_loss] after the forward pass; backward and the optimizer (two thirds of the
    # step) do not change it. A side stream copies it to pinned memory as soon as it exists, so a host read of the
    # step's loss (logging every step) waits for the forward pass only and the GPU keeps running.
    def _stage_loss_readback(self, stats):
        dev = stats.device
        if self._loss_stream is None:
            self._loss_stream = torch.cuda.Stream(device=dev)
            self._loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
            self._loss_event = torch.cuda.Event()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._loss_stream):
            self._loss_stream.wait_event(ready)
            self._loss_host.copy_(stats, non_blocking=True)
            self._loss_event.record(self._loss_stream)
        stats.record_stream(self._loss_stream)

    def last_loss_value(self):
        """Mean loss of the last train_step as a Python float (device -> host read; waits for that step's forward pass,
        not for its backward / optimizer)."""
        if self._loss_event is None:
            return float(self.last_loss[1].item())
        self._loss_event.synchronize()
        return float(self._loss_host[1])

    def eval_step(self, sample):
        pass

    def state_dict(self):
        state_dicts = {'model': self.model.state_dict(), 'optimizer': self.optimizer.state_dict()}
        if hasattr(self.scheduler, 'state_dict'):
            state_dicts['scheduler'] = self.scheduler.state_dict()
        return state_dicts

    def load_state_dict(self, state_dict):
        sd = state_dict.get('model', state_dict)
        sd = {k.replace('module.', '', 1) if k.startswith('module.') else k: v for k, v in sd.items()}
        self.model.load_state_dict(sd)
        if 'optimizer' in state_dict and self.optimizer is not None:
            self.optimizer.load_state_dict(state_dict['optimizer'])
        if 'scheduler' in state_dict and self.scheduler is not None:
            self.scheduler.load_state_dict(state_dict['scheduler'])

    def __repr__(self):
        return '\n'.join([f'model: {repr(self.model)}', f'opt: {repr(self.optimizer)}', f'sched: {repr(self.scheduler)}'])


def build_image_preprocess(image_size, mean, std):
    """The reference's CPU page preprocessing (task_cruller_pretrain.py:132-143): ToTensor -> bicubic antialiased
    Resize (no aspect preservation) -> scalar Normalize. Runs in DataLoader workers, outside the GPU step."""
    import torchvision.transforms as transforms
    return transforms.Compose([
        transforms.ToTensor(),
        transforms.Resize(tuple(image_size), interpolation=transforms.InterpolationMode.BICUBIC, antialias=True),
        transforms.Normalize(mean=mean, std=std),
    ])


def preprocess_text_tokens(anno, tokenizer, max_position_embeddings, task_start_token, prompt_end_token,
                           ignore_id=-100, generator=None):
    """Annotation -> (text ids, target ids), as data/preprocess.py:9-40 does for raw-text annotations:
    pad to max length, PAD -> ignore_id, everything up to and including the prompt-end token -> ignore_id."""
    text = task_start_token + anno + tokenizer.eos_token
    ids = tokenizer(text, add_special_tokens=False, return_tensors='pt', max_length=max_position_embeddings,
                    padding='max_length', truncation=True).input_ids[0]
    target = ids.clone()
    target[target == tokenizer.pad_token_id] = ignore_id
    prompt_end_token_id = tokenizer.convert_tokens_to_ids(prompt_end_token)
    target[:torch.nonzero(target == prompt_end_token_id).sum() + 1] = ignore_id
    return dict(text=[ids], target=[target])


def _tokenize_padded(tokenizer, text, max_len):
    return tokenizer(text, add_special_tokens=False, return_tensors='pt', max_length=max_len, padding='max_length',
                     truncation=True).input_ids[0]


def _prompt_masked_target(ids, pad_id, prompt_end_id, ignore_id):
    target = ids.clone()
    target[target == pad_id] = ignore_id
    target[:torch.nonzero(target == prompt_end_id).sum() + 1] = ignore_id
    return target


def next_page_with_text(index, num_pages, anno, retries=10):
    """data/preprocess.py:112-131: the next page (cyclically) whose 'text' is non-empty, within `retries` tries."""
    for _ in range(retries):
        index = (index + 1) % num_pages
        if anno['pages'][index]['text']:
            return index
    raise RuntimeError(f"No non-empty page found after {retries} attempts")


def preprocess_ocr_anno(anno, tokenizer, max_position_embeddings, task_start_token, prompt_end_token, ignore_id=-100,
                        generator=None):
    """pixparse OCR annotation ``{'pages': [{'text': [line, ...], ...}, ...]}`` -> one randomly sampled page with text,
    its lines joined by newlines, tokenised / padded / masked like the raw-text path (data/preprocess.py:43-110).
    ``generator`` is a ``random.Random``-like object (``randint`` inclusive on both ends), as chug passes it.
    Returns ``(dict(text=[ids], target=[ids]), dict(page_indices, num_pages, orig_text))``."""
    if isinstance(anno, list):          # legacy [id, {...}] form
        _logger.warning("Old [id, {}] annotation form found, correcting...")
        anno = anno[1]
    if not isinstance(anno, dict) or 'pages' not in anno:
        raise TypeError("preprocess_ocr_anno expects a pixparse OCR annotation dict with a 'pages' list; raw-text "
                        "annotations go through preprocess_text_tokens (task.text_anno_fn = True)")
    num_pages = len(anno['pages'])
    if not num_pages:
        raise RuntimeError("Empty annotation. Skipping...")
    if generator is None:
        import random
        generator = random
    index = generator.randint(0, num_pages - 1)
    if not anno['pages'][index]['text']:
        index = next_page_with_text(index, num_pages, anno)
    pad_id = tokenizer.pad_token_id
    prompt_end_id = tokenizer.convert_tokens_to_ids(prompt_end_token)
    texts, targets, indices = [], [], []
    orig_text = None
    wanted = min(1, num_pages)          # single-page mode, as in the reference
    while len(texts) < wanted:
        page = anno['pages'][index]
        if not page['text']:
            raise RuntimeError("No text on page, skipping...")
        orig_text = '\n'.join(page['text'])
        ids = _tokenize_padded(tokenizer, task_start_token + orig_text + tokenizer.eos_token, max_position_embeddings)
        texts.append(ids)
        targets.append(_prompt_masked_target(ids, pad_id, prompt_end_id, ignore_id))
        indices.append(index)
        index = next_page_with_text(index, num_pages, anno)
    return dict(text=texts, target=targets), dict(page_indices=indices, num_pages=num_pages, orig_text=orig_text)
