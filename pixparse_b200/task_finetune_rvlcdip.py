"""RVL-CDIP fine-tuning task on the B200 path: mirror of ``pixparse.task.TaskCrullerFinetuneRVLCDIP``
(/root/reference/src/pixparse/task/task_cruller_finetune_RVLCDIP.py:51-403).

Same step as pre-training on very short json-completion targets: ``<s_rvlcdip><class/></s>`` padded to 5 tokens
(T = 4 after the shift), 21 extra tokens on top of the pre-training vocabulary (V = 50286), dict samples
``{'image', 'label', 'text_target'}``, optional layer-decay. Differences kept from the reference: the pre-training
checkpoint is loaded (keys with ``module.`` stripped) BEFORE the vocabulary grows (:222-234); logging happens before
the update gate (:386-396). Difference fixed: an absent checkpoint no longer raises.
"""
import logging
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import torch

from .framework import DeviceEnv
from .models import Cruller, ModelCfg, get_model_config
from .task_pretrain import (TaskCrullerPretrain, TaskCrullerPretrainCfg, TokenizerCfg, _TokenizerHolder,
                            build_image_preprocess, load_tokenizer)

_logger = logging.getLogger(__name__)


@dataclass
class TaskCrullerFinetuneRVLCDIPCfg(TaskCrullerPretrainCfg):
    pass


class TaskCrullerFinetuneRVLCDIP(TaskCrullerPretrain):
    int2str = {0: "letter", 1: "form", 2: "email", 3: "handwritten", 4: "advertisement", 5: "scientific_report",
               6: "scientific_publication", 7: "specification", 8: "file_folder", 9: "news_article", 10: "budget",
               11: "invoice", 12: "presentation", 13: "questionnaire", 14: "resume", 15: "memo"}

    def __init__(self, cfg: TaskCrullerFinetuneRVLCDIPCfg, device_env: DeviceEnv, monitor=None, tokenizer=None):
        # TaskTrain.__init__ only (the pre-training constructor adds different tokens)
        super(TaskCrullerPretrain, self).__init__(cfg=cfg, device_env=device_env, monitor=monitor)
        self.cfg = cfg
        self.amp_dtype = None
        if cfg.dtype is not None:
            self.amp_dtype = torch.bfloat16 if cfg.dtype in ("bfloat16", "bf16") else torch.float16
        if cfg.amp and self.amp_dtype is torch.float16:
            raise ValueError("the B200 path computes in bf16; fp16 autocast + loss scaling is not implemented")
        self.task_start_token = "<s_rvlcdip>"
        self.prompt_end_token = self.task_start_token
        self.max_position_embeddings = cfg.model.text_decoder.max_length
        self.text_anno_fn = True
        self.tokenizer = _TokenizerHolder(tokenizer if tokenizer is not None else load_tokenizer(cfg.tokenizer))
        self.state_dict_to_load = OrderedDict()     # the reference stashes the checkpoint on task.state_dict
        self.resume = False
        self.special_tokens_finetune = [
            "<sep/>", self.task_start_token, self.prompt_end_token, "<s_class>", "</s_class>", "<advertisement/>",
            "<budget/>", "<email/>", "<file_folder/>", "<form/>", "<handwritten/>", "<invoice/>", "<letter/>",
            "<memo/>", "<news_article/>", "<presentation/>", "<questionnaire/>", "<resume/>",
            "<scientific_publication/>", "<scientific_report/>", "<specification/>"]
        cfg.model.image_encoder.pretrained = False
        cfg.model.text_decoder.pretrained = False
        self.model = Cruller(cfg.model)
        n_pre = self.tokenizer.trunk.add_special_tokens(
            {"additional_special_tokens": sorted(set(["<sep/>", "<s_pretrain>"]))})
        if n_pre > 0:
            self.model.text_decoder.trunk.resize_token_embeddings(len(self.tokenizer.trunk))
        self.vocab_size = len(self.tokenizer.trunk)
        self.has_no_sync = False
        self.num_image_chs = 1 if cfg.model.image_encoder.image_fmt == "L" else 3
        img_mean = self.model.image_encoder.trunk.pretrained_cfg["mean"]
        img_std = self.model.image_encoder.trunk.pretrained_cfg["std"]
        gray = cfg.model.image_encoder.image_fmt == "L"
        self.img_mean = sum(img_mean) / len(img_mean) if gray else img_mean
        self.img_std = sum(img_std) / len(img_std) if gray else img_std
        self.image_preprocess_train = build_image_preprocess(cfg.model.image_encoder.image_size, self.img_mean,
                                                             self.img_std)
        self.image_preprocess_eval = None
        self.train_metrics, self.eval_metrics = {}, {}
        self.max_recursion_length = 1000
        self._init_step_state()

    def train_setup(self, num_batches_per_interval: int):
        if self.state_dict_to_load:
            _logger.info("Resuming from existing checkpoint.")
            sd = {k.replace("module.", ""): v for k, v in self.state_dict_to_load.items()}
            self.model.load_state_dict(sd)
        self.newly_added_num = self.tokenizer.trunk.add_special_tokens(
            {"additional_special_tokens": sorted(set(self.special_tokens_finetune))})
        self.vocab_size = len(self.tokenizer.trunk)
        if self.newly_added_num > 0:
            self.model.text_decoder.trunk.resize_token_embeddings(len(self.tokenizer.trunk))
        super().train_setup(num_batches_per_interval)

    def train_interval_start(self):      # the reference defines no interval hooks for this task
        pass

    def train_interval_end(self):
        pass

    def text_input_to_target(self, text_input, ignore_id=-100):
        target = text_input.clone()
        target[target == self.tokenizer.trunk.pad_token_id] = ignore_id
        prompt_end_token_id = self.tokenizer.trunk.convert_tokens_to_ids(self.prompt_end_token)
        target[: torch.nonzero(target == prompt_end_token_id).sum() + 1] = ignore_id
        return target

    def label_tokens(self, label: int):
        """``<s_rvlcdip><class/></s>`` padded to 5 ids (collate_fn :309-321); needs only special-token lookups."""
        t = self.tokenizer.trunk
        ids = [t.convert_tokens_to_ids(self.task_start_token), t.convert_tokens_to_ids("<" + self.int2str[label] + "/>"),
               t.eos_token_id]
        ids = ids[:5] + [t.pad_token_id] * (5 - len(ids))
        return torch.tensor(ids, dtype=torch.long)

    def collate_fn(self, batch):
        images = torch.stack([self.image_preprocess_train(item["image"]) for item in batch])
        labels = torch.stack([self.label_tokens(int(item["label"])) for item in batch])
        targets = torch.stack([self.text_input_to_target(text) for text in labels])
        return {"image": images, "label": labels[:, :-1], "text_target": targets[:, 1:]}

    def train_step(self, sample: Dict[str, Any]) -> Dict[str, Any]:
        device = self.device_env.device
        image_input, label, text_target = self._to_device(sample["image"], sample["label"], sample["text_target"])
        label, text_target = label.contiguous(), text_target.contiguous()
        result = {}
        accum_steps = self.cfg.opt.grad_accum_steps
        need_update = (self.interval_batch_idx + 1) % accum_steps == 0
        if self.reducer is not None:
            self.reducer.enabled = need_update
            self.reducer.begin()
        if device.type == 'cuda' and self.engine.on_loss_ready is None:
            self.engine.on_loss_ready = self._stage_loss_readback
        self.last_loss = self.engine.forward_backward(image_input, label, text_target, grad_scale=1.0 / accum_steps)
        if self.reducer is not None:
            self.reducer.finish()
        self.batch_idx += 1
        self.interval_batch_idx += 1
        if self.step % self.eval_frequency == 0 and self.monitor is not None:
            self.monitor.log_step(
                "finetune", step_idx=self.step, step_end_idx=self.num_intervals * self.num_steps_per_interval,
                interval=self.interval_idx, loss=self.last_loss_value() / accum_steps, lr=self.get_current_lr(),
                metrics=None, eval_data=None)
        if not need_update:
            return result
        self.optimizer.step(clip_grad_norm=self.cfg.opt.clip_grad_value)
        self.step += 1
        self.scheduler.step_update(self.step)
        self.optimizer.zero_grad()
        return result
