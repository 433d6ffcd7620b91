"""OCR evaluation task on the B200 path: mirror of ``pixparse.task.TaskCrullerEvalOCR``
(/root/reference/src/pixparse/task/task_cruller_eval_ocr.py:49-250): ``setup`` loads ``resume_state_dict`` and
switches to eval mode, ``step`` runs the encoder once and the uncached greedy decode, ``average_metrics`` averages
CER / WER over batches."""
import logging
import time
from dataclasses import dataclass, field
from typing import Optional

import torch

from .framework import DeviceEnv, TaskEval, TaskEvalCfg
from .models import Cruller, ModelCfg, get_model_config
from .ocr_utils import get_ocr_metrics
from .task_pretrain import TokenizerCfg, _TokenizerHolder, build_image_preprocess, load_tokenizer

_logger = logging.getLogger(__name__)


@dataclass
class TaskCrullerEvalOCRCfg(TaskEvalCfg):
    model_name: Optional[str] = None
    model: ModelCfg = field(default_factory=ModelCfg)
    tokenizer: TokenizerCfg = field(default_factory=TokenizerCfg)

    def __post_init__(self):
        if self.model_name:
            model = get_model_config(self.model_name)
            if model is None:
                _logger.warning(f"Model config for {self.model_name} was not found, using defaults.")
            else:
                self.model = model
        else:
            self.model_name = "custom"


class TaskCrullerEvalOCR(TaskEval):
    def __init__(self, cfg: TaskCrullerEvalOCRCfg, device_env: DeviceEnv, monitor=None, tokenizer=None):
        super().__init__(cfg=cfg, device_env=device_env, monitor=monitor)
        self.cfg = cfg
        self.task_start_token = "<s_pretrain>"
        self.prompt_end_token = self.task_start_token
        self.max_position_embeddings = cfg.model.text_decoder.max_length
        self.tokenizer = _TokenizerHolder(tokenizer if tokenizer is not None else load_tokenizer(cfg.tokenizer))
        special_tokens = ["<sep/>", self.task_start_token, self.prompt_end_token]
        newly_added_num = self.tokenizer.trunk.add_special_tokens(
            {"additional_special_tokens": sorted(set(special_tokens))})
        self.vocab_size = len(self.tokenizer.trunk)
        cfg.model.image_encoder.pretrained = False
        cfg.model.text_decoder.pretrained = False
        self.model = Cruller(cfg.model)
        if newly_added_num > 0:
            self.model.text_decoder.trunk.resize_token_embeddings(len(self.tokenizer.trunk))
        img_mean = self.model.image_encoder.trunk.pretrained_cfg["mean"]
        img_std = self.model.image_encoder.trunk.pretrained_cfg["std"]
        gray = cfg.model.image_encoder.image_fmt == "L"
        self.img_mean = sum(img_mean) / len(img_mean) if gray else img_mean
        self.img_std = sum(img_std) / len(img_std) if gray else img_std
        self.image_preprocess_eval = build_image_preprocess(cfg.model.image_encoder.image_size, self.img_mean,
                                                            self.img_std)
        self.anno_preprocess_eval = None
        self.resume_state_dict = None
        self.eval_metrics = {}
        self.max_recursion_length = 1000

    def setup(self):
        device = self.device_env.device
        if self.resume_state_dict:
            sd = {k.replace("module.", "", 1) if k.startswith("module.") else k: v
                  for k, v in self.resume_state_dict.items()}
            self.model.load_state_dict(sd)
        self.model.eval()
        self.model.to(device)

    def prepare_for_evaluation(self, loaders):
        return {k: v for k, v in loaders.items() if k in ["eval", "eval_FUNSD"]}

    def step(self, sample):
        t0 = time.time()
        metrics = {}
        image_input, text_input, text_target = sample
        if isinstance(text_target, (list, tuple)):
            text_target = torch.stack([item[0] for item in text_target], dim=0)
        text_target = text_target.to(self.device_env.device, non_blocking=True)
        image_input = image_input.to(self.device_env.device, non_blocking=True)
        ocr_metrics, _ = get_ocr_metrics(
            model=self.model, tokenizer=self.tokenizer, image_input=image_input, text_input=text_target,
            device_env=self.device_env, max_recursion_length=self.max_recursion_length,
            prompt_token=self.task_start_token)
        metrics["ocr_reconstruction"] = ocr_metrics
        _logger.info(f"Executed method step in {time.time() - t0:.2f} seconds")
        return metrics

    def average_metrics(self, metrics: dict):
        wer_sum = sum(m["ocr_reconstruction"]["wer"] for m in metrics.values())
        cer_sum = sum(m["ocr_reconstruction"]["cer"] for m in metrics.values())
        n = len(metrics)
        return {"ocr_reconstruction": {"wer": wer_sum / n, "cer": cer_sum / n}}

    def end(self):
        pass

    def state_dict(self):
        return {"model": self.model.state_dict()}
