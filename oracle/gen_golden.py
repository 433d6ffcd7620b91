"""ORACLE tooling: generate tests/golden/*.json by running the REFERENCE's own train step.

Runs only in the build container (needs /root/reference). What executes here is pixparse's unmodified
``TaskCrullerPretrain.__init__ / train_setup / train_step`` (task/task_cruller_pretrain.py:63-313), ``Cruller`` and
``TextDecoderHf`` over the shim layer in oracle/ref_shims.py, on CPU fp32 with amp=False and dropout 0, fed by
pixparse_b200.synthetic batches. The numbers it records pin oracle/cruller_ref.py (tests/test_golden.py) and are the
fixtures the GPU parity tests compare against.

    python -m oracle.gen_golden            # writes tests/golden/pretrain_*.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from pixparse_b200 import synthetic  # noqa: E402

CASES = {
    # name: (image encoder, image size, text decoder, decoder layers, batch, text_len, steps)
    "pretrain_tiny": dict(enc="vit_test_patch16", size=(64, 48), dec="test/bart-tiny", layers=2, B=2, Lt=17, steps=3),
    "pretrain_tiny_prenorm": dict(enc="vit_test_patch14_clip", size=(56, 42), dec="test/bart-tiny", layers=2, B=3,
                                  Lt=9, steps=2),
    # BASELINE.json configs[0]
    "pretrain_cruller_base_b2": dict(enc="vit_base_patch16_224", size=(576, 448), dec="facebook/bart-base", layers=4,
                                     B=2, Lt=513, steps=1),
}


class _HostBatchTensor(torch.Tensor):
    """On a GPU the reference's ``x[:, 1:].to(device)`` (task_cruller_pretrain.py:240-242) is a host-to-device COPY
    and therefore contiguous, which its later ``.view(-1)`` relies on; on CPU ``.to`` is a no-op. This subclass
    restores the copy semantics so the unmodified train_step runs on CPU. Values are untouched."""

    def to(self, *a, **k):
        return super().to(*a, **k).contiguous().as_subclass(torch.Tensor)


class _Monitor:
    def log_step(self, *a, **k):
        pass

    def log_phase(self, *a, **k):
        pass


def run_case(name, c, seed=0):
    ref_shims.install()
    from pixparse.framework.config import OptimizationCfg
    from pixparse.models.config import ImageEncoderCfg, ModelCfg, TextDecoderCfg
    from pixparse.task.task_cruller_pretrain import TaskCrullerPretrain, TaskCrullerPretrainCfg

    model_cfg = ModelCfg(
        image_encoder=ImageEncoderCfg(name=c["enc"], image_fmt='L', image_size=c["size"], pretrained=False),
        text_decoder=TextDecoderCfg(name=c["dec"], pretrained=False, num_decoder_layers=c["layers"], max_length=1024))
    opt = OptimizationCfg(optimizer='adamw', scheduler='cosine', learning_rate=3e-4, warmup_learning_rate=0.0,
                          eps=1e-6, clip_grad_value=1.0, clip_grad_mode='norm', grad_accum_steps=1,
                          betas=(0.9, 0.98))
    cfg = TaskCrullerPretrainCfg(num_intervals=100, num_warmup_intervals=5, eval_frequency=10 ** 9, opt=opt,
                                 dtype=None, amp=False, model=model_cfg)
    torch.manual_seed(seed)
    task = TaskCrullerPretrain(cfg, ref_shims.CpuDeviceEnv("cpu"), monitor=_Monitor())
    assert task.vocab_size == synthetic.PRETRAIN_VOCAB
    assert task.tokenizer.trunk.convert_tokens_to_ids('<s_pretrain>') == synthetic.S_PRETRAIN_ID
    task.train_setup(num_batches_per_interval=1000)
    task.train_interval_start()

    rec = {"case": name, "config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()},
           "seed": seed, "vocab": task.vocab_size, "steps": [],
           "optimizer": dict(lr=3e-4, betas=[0.9, 0.98], eps=1e-6, clip_grad=1.0, weight_decay=0.0,
                             num_intervals=100, num_warmup_intervals=5, steps_per_interval=1000)}
    losses = []
    orig_loss = task.loss

    def spy_loss(logits, target):
        out = orig_loss(logits, target)
        losses.append((float(out.detach()), logits.detach().clone()))
        return out

    task.loss = spy_loss
    gnorms = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def spy_clip(params, max_norm, norm_type=2.0, **kw):
        total = orig_clip(params, max_norm, norm_type=norm_type, **kw)
        gnorms.append(float(total))
        return total

    torch.nn.utils.clip_grad_norm_ = spy_clip
    try:
        for step in range(c["steps"]):
            sample = tuple(t.as_subclass(_HostBatchTensor)
                           for t in synthetic.synthetic_batch(c["B"], c["size"], c["Lt"], seed=seed + step))
            task.train_step(sample)
            loss, logits = losses[-1]
            probe = logits.reshape(-1, logits.shape[-1])
            idx = torch.arange(0, probe.shape[0], max(1, probe.shape[0] // 8))[:8]
            rec["steps"].append({
                "loss": loss,
                "grad_norm": gnorms[-1],
                "lr_after": task.get_current_lr(),
                "logits_abs_mean": float(logits.abs().mean()),
                "logits_probe": probe[idx, :4].tolist(),
                "logits_probe_rows": idx.tolist(),
            })
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    model = task.model
    rec["param_checksums_after"] = {
        n: [float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
        for n, p in list(model.named_parameters())[:6] + list(model.named_parameters())[-6:]}
    rec["num_params"] = sum(p.numel() for p in model.parameters())
    return rec


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    which = sys.argv[1:] or list(CASES)
    for name in which:
        rec = run_case(name, CASES[name])
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(rec, f, indent=1)
        print(name, [s["loss"] for s in rec["steps"]], [s["grad_norm"] for s in rec["steps"]])


if __name__ == "__main__":
    main()
