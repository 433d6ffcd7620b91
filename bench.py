#!/usr/bin/env python
"""Benchmark of the Cruller train step (BASELINE.json metric: cruller_base train pages/sec, MFU vs bf16 peak).

    python bench.py --gpus 1 --steps K --warmup W              # this repo's sm_100a path
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # data parallel, one rank per GPU, NCCL
    python bench.py --impl reference --steps K --warmup W      # the reference path (oracle) on the host CPU cores

A step = one full optimizer update on a fresh batch: forward, token cross-entropy, backward, gradient all-reduce
(N > 1), global-norm clip, AdamW, grad zeroing. Workload = BASELINE.json configs[1]: cruller_base pretrain, bf16,
32 synthetic grayscale 576x448 pages + 512-token targets per GPU (weak scaling).

One JSON line on stdout (rank 0). `value` is timed with the batches already in HBM; `e2e` goes through the
public Task API (TaskCrullerPretrain.train_step) from pinned host buffers and reads the loss back every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "cruller_base"
BATCH_PER_GPU = 32
TEXT_LEN = 513           # 512 decoder positions after the shift
GFLOP_PER_PAGE = 905.3   # fwd + bwd, SURVEY.md Appendix D / BASELINE.md section 3

# BASELINE.json configs[1..4]. The headline line is always "base" (configs[1]); the others are reported as extra keys of
# the same JSON line (`extra_configs`) or on their own with --config.
CONFIGS = {
    # configs[1]: cruller_base pretrain bf16, 32 pages / GPU, T = 512, lr 3e-4, betas (0.9, 0.98) (README.md:26-30)
    "base": dict(model="cruller_base", task="pretrain", batch=32, text_len=513, gflop_per_page=905.3, lr=3e-4,
                 betas=(0.9, 0.98), layer_decay=None,
                 workload="cruller_base pretrain step (fwd+CE+bwd+allreduce+clip+AdamW), bf16, {B} synthetic 576x448 "
                          "grayscale pages + 512-token targets per GPU"),
    # configs[2]: cruller_large (ViT-L/14 CLIP 798x616 -> S = 2509, bart-large x10), 8 pages / GPU (README.md:22)
    "large": dict(model="cruller_large", task="pretrain", batch=8, text_len=513, gflop_per_page=7520.5, lr=3e-4,
                  betas=(0.9, 0.98), layer_decay=None,
                  workload="cruller_large pretrain step, bf16, {B} synthetic 798x616 grayscale pages + 512-token targets "
                           "per GPU"),
    # configs[3]: RVL-CDIP json completion: T = 4, V = 50286, layer_decay 0.75, lr 1e-4, betas (0.9, 0.99) (README.md:65-89,127)
    "finetune": dict(model="cruller_base", task="finetune", batch=32, text_len=5, gflop_per_page=658.4, lr=1e-4,
                     betas=(0.9, 0.99), layer_decay=0.75,
                     workload="cruller_base RVL-CDIP finetune step (T=4 json-completion targets, V=50286, layer_decay "
                              "0.75), bf16, {B} synthetic pages + 16-class labels per GPU"),
    # configs[4]: cruller_large_6layers greedy decode, 16 pages, 512 new tokens, KV-cached (README.md:53, ocr_utils.py:165-197)
    "eval_ocr": dict(model="cruller_large_6layers", task="eval_ocr", batch=16, new_tokens=512,
                     gflop_per_page_encoder=2135.2,
                     workload="cruller_large_6layers eval_ocr: encoder + greedy autoregressive decode of {T} tokens with "
                              "KV-cached self/cross attention, {B} synthetic 798x616 pages"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '', 1).isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '', 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        pw = [float(s[2]) for s in self.samples if s[2].replace('.', '', 1).isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (CPU restatement of the reference train step) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def time_cpu_reference(steps, warmup, batch, budget_s=None):
    """Times oracle.cruller_ref.OracleTrainer.train_step (fp32, all host threads). Returns (pages/s, s/step, info)."""
    from oracle import cruller_ref
    from pixparse_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = cruller_ref.build_model(MODEL, vocab_size=synthetic.PRETRAIN_VOCAB, seed=0)
    trainer = cruller_ref.OracleTrainer(model, synthetic.PRETRAIN_VOCAB, lr=3e-4, betas=(0.9, 0.98), eps=1e-6,
                                        clip_grad=1.0)
    size = tuple(cruller_ref.MODEL_CONFIGS[MODEL].image_encoder.image_size)
    t_start = time.perf_counter()
    for i in range(warmup):
        trainer.train_step(synthetic.synthetic_batch(batch, size, TEXT_LEN, seed=1000 + i))
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    times = []
    for i in range(steps):
        sample = synthetic.synthetic_batch(batch, size, TEXT_LEN, seed=i)
        t0 = time.perf_counter()
        trainer.train_step(sample)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    sec = sum(times) / len(times)
    return batch / sec, sec, {"cores": cores, "batch": batch, "steps": len(times)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: 1 page per step keeps K + W steps within a few minutes on the host cores
    batch = 1
    pps, sec, info = time_cpu_reference(args.steps, args.warmup, batch)
    line = {
        "impl": "reference", "metric": "cruller_base train pages/sec", "value": pps, "unit": "pages/s",
        "n_gpus": args.gpus, "n_gpus_used": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cruller_base pretrain fwd+bwd+AdamW step, 576x448 grayscale pages, 512-token targets",
                   "pages_per_step": batch, "device": "host CPU, one process (rank 0) whatever --gpus says",
                   "note": "bounded sample: 1 page per step instead of the GPU arm's 32 per GPU (same model, sequence "
                           "length and optimizer); pages/s is per-page work either way"},
        "cpu_baseline": {"value": pps, "unit": "pages/s", "cores": info["cores"], "kind": "port",
                         "sample": f"{args.steps} train steps of {batch} page(s) (T=512, fp32, all host threads); "
                                   "oracle = restated timm ViT + installed transformers BART + torch AdamW"},
        "e2e": {"value": pps, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# GPU reference: the SAME oracle modules (restated timm ViT + installed transformers BART) on the B200 under
# torch.autocast(bf16) with cuBLAS / SDPA / foreach AdamW / DDP -- what the reference's train_step runs on a GPU box
# (task_cruller_pretrain.py:205-210, 247-278). "Does the hand-written path beat stock PyTorch on the same box?"
# ---------------------------------------------------------------------------------------------------------------------
def time_gpu_reference(cfgname, steps, warmup, dev, world=1, rank=0, batch=None):
    import torch.distributed as dist
    import torch.nn.functional as F
    from oracle import cruller_ref
    from oracle.timm_helpers import create_optimizer_v2, dispatch_clip_grad
    from pixparse_b200 import synthetic
    c = CONFIGS[cfgname]
    B = batch or c["batch"]
    vocab = synthetic.PRETRAIN_VOCAB if c["task"] == "pretrain" else synthetic.PRETRAIN_VOCAB + 19
    model = cruller_ref.build_model(c["model"], vocab_size=vocab, seed=0, dropout_off=False).to(dev)
    model.train()
    size = tuple(cruller_ref.MODEL_CONFIGS[c["model"]].image_encoder.image_size)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev], static_graph=True)
    opt = create_optimizer_v2(model, 'adamw', lr=c["lr"], eps=1e-6, layer_decay=c["layer_decay"], betas=c["betas"])
    batches = []
    for i in range(4):
        img, txt, tgt = synthetic.synthetic_batch(B, size, c["text_len"], vocab=vocab, seed=100 * rank + i)
        batches.append((img.to(dev), txt[:, :-1].contiguous().to(dev), tgt[:, 1:].contiguous().to(dev)))

    def step(i):
        img, txt, tgt = batches[i % 4]
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            logits = net(img, txt)["logits"]
            loss = F.cross_entropy(logits.view(-1, vocab), tgt.view(-1), ignore_index=-100)
        loss.backward()
        dispatch_clip_grad(model.parameters(), 1.0, 'norm')
        opt.step()
        opt.zero_grad()
        return loss

    for i in range(max(warmup, 3)):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = {"value": world * B / (ms * 1e-3), "unit": "pages/s", "ms_per_step": ms, "batch_per_gpu": B,
           "loss": float(loss.item()), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
           "what": "oracle modules (restated timm ViT + transformers BartForCausalLM) under torch.autocast(bf16): cuBLAS "
                   "GEMMs, F.scaled_dot_product_attention, nn.CrossEntropyLoss, clip_grad_norm_, torch.optim.AdamW"
                   + (", DistributedDataParallel(static_graph=True)" if world > 1 else "") + "; dropout live",
           "torch": torch.__version__}
    del model, net, opt, batches
    torch.cuda.empty_cache()
    return out


def run_reference_gpu_arm(args):
    from pixparse_b200.framework import DeviceEnv
    env = DeviceEnv()
    dev = env.device if env.device.index is not None else torch.device("cuda", torch.cuda.current_device())
    torch.cuda.set_device(dev)
    c = CONFIGS[args.config]
    sampler = ClockSampler(dev.index)
    sampler.start()
    r = time_gpu_reference(args.config, args.steps, args.warmup, dev, env.world_size, env.global_rank, args.batch)
    sampler.stop()
    if env.global_rank == 0:
        peaks = load_peaks()
        line = {"impl": "reference_gpu", "metric": f"{c['model']} train pages/sec", "value": r["value"], "unit": "pages/s",
                "n_gpus": env.world_size, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": c["workload"].format(B=r["batch_per_gpu"]), "implementation": r["what"]},
                "mfu": {"vs_measured_burst": r["value"] * c["gflop_per_page"] / 1e3 / (env.world_size * peaks["bf16_burst"])},
                "clocks": sampler.summary(), "loss": r["loss"], "peak_mem_gb": r["peak_mem_gb"]}
        emit_json(line)
    if env.world_size > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def build_train_task(cfgname, env, batch=None, dropout_off=False):
    """Task + a rotation of pinned host batches + the matching device-resident batches for a train config."""
    from pixparse_b200 import synthetic
    from pixparse_b200.framework import OptimizationCfg
    c = CONFIGS[cfgname]
    B = batch or c["batch"]
    rank = env.global_rank
    dev = env.device
    opt = OptimizationCfg(learning_rate=c["lr"], betas=c["betas"], eps=1e-6, clip_grad_value=1.0, clip_grad_mode='norm',
                          grad_accum_steps=1, layer_decay=c["layer_decay"])
    torch.manual_seed(0)
    if c["task"] == "pretrain":
        from pixparse_b200.task_pretrain import TaskCrullerPretrain, TaskCrullerPretrainCfg
        cfg = TaskCrullerPretrainCfg(model_name=c["model"], opt=opt, dtype='bfloat16', amp=True, eval_frequency=10 ** 9,
                                     num_intervals=100, num_warmup_intervals=5)
        task = TaskCrullerPretrain(cfg, env, monitor=None, tokenizer=synthetic.SyntheticBartTokenizer())
        assert task.vocab_size == synthetic.PRETRAIN_VOCAB
    else:
        from pixparse_b200.task_finetune_rvlcdip import TaskCrullerFinetuneRVLCDIP, TaskCrullerFinetuneRVLCDIPCfg
        cfg = TaskCrullerFinetuneRVLCDIPCfg(model_name=c["model"], opt=opt, dtype='bfloat16', amp=True,
                                            eval_frequency=10 ** 9, num_intervals=100, num_warmup_intervals=5)
        task = TaskCrullerFinetuneRVLCDIP(cfg, env, monitor=None, tokenizer=synthetic.SyntheticBartTokenizer())
    if dropout_off:
        task.model.text_decoder.trunk.set_dropout(0.0)
    task.train_setup(num_batches_per_interval=1000)
    task.train_interval_start()
    size = tuple(cfg.model.image_encoder.image_size)
    n_rot = 4
    host, resident = [], []
    for i in range(n_rot):
        if c["task"] == "pretrain":
            img, txt, tgt = synthetic.synthetic_batch(B, size, c["text_len"], seed=100 * rank + i)
            host.append((img.pin_memory(), txt.pin_memory(), tgt.pin_memory()))
            resident.append((img.to(dev), txt[:, :-1].contiguous().to(dev), tgt[:, 1:].contiguous().to(dev)))
        else:
            g = torch.Generator().manual_seed(100 * rank + i)
            img = (torch.rand((B, 1) + size, generator=g) - 0.5) / 0.5
            labels = torch.randint(0, 16, (B,), generator=g)
            ids = torch.stack([task.label_tokens(int(l)) for l in labels])              # collate_fn :309-321
            tgt = torch.stack([task.text_input_to_target(t) for t in ids])
            sample = {"image": img.pin_memory(), "label": ids[:, :-1].contiguous().pin_memory(),
                      "text_target": tgt[:, 1:].contiguous().pin_memory()}
            host.append(sample)
            resident.append((img.to(dev), sample["label"].to(dev), sample["text_target"].to(dev)))
    if c["task"] == "pretrain":
        h2d = sum(t.numel() * t.element_size() for t in host[0])
    else:
        h2d = sum(t.numel() * t.element_size() for t in host[0].values())
    return task, host, resident, h2d, B


def time_train_config(cfgname, env, steps, warmup, batch=None, dropout_off=False, want_profile=False):
    """Device-resident and end-to-end (Task.train_step from pinned host batches, loss read back) timing of one train
    config. Returns a dict; the caller assembles the JSON line."""
    import torch.distributed as dist
    from pixparse_b200 import _lib
    world, dev = env.world_size, env.device
    task, host, resident, h2d_bytes, B = build_train_task(cfgname, env, batch, dropout_off)
    n_rot = len(host)
    clip = task.cfg.opt.clip_grad_value

    def device_step(i):
        img, txt, tgt = resident[i % n_rot]
        if task.reducer is not None:
            task.reducer.enabled = True
            task.reducer.begin()
        stats = task.engine.forward_backward(img, txt, tgt)
        if task.reducer is not None:
            task.reducer.finish()
        task.optimizer.step(clip_grad_norm=clip)
        task.step += 1
        task.scheduler.step_update(task.step)
        return stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(warmup, 3)):
        device_step(i)
    barrier()
    sampler = ClockSampler(dev.index if dev.index is not None else 0)
    sampler.start()
    _lib.reset_launch_count()
    if task.reducer is not None:
        task.reducer.measure_exposed = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(steps):
        last = device_step(i)
    e1.record()
    barrier()
    if task.reducer is not None:
        task.reducer.measure_exposed = False
    launches = _lib.launch_count()
    exposed = task.reducer.exposed_ms() if (task.reducer is not None and hasattr(task.reducer, "exposed_ms")) else None
    own_ms = e0.elapsed_time(e1) / steps
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / steps
    per_rank = [own_ms]
    if world > 1:      # every rank's own device time (diagnostic: box-internal GPU-to-GPU spread vs exchange cost)
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[env.global_rank] = own_ms
        dist.all_reduce(t)
        per_rank = [round(float(x), 3) for x in t.tolist()]
    loss_val = float(last[1].item())

    # ---- end to end through the Task API: pinned host batch -> train_step -> loss read back
    for i in range(2):
        task.train_step(host[i % n_rot])
    barrier()
    e0.record()
    for i in range(steps):
        task.train_step(host[i % n_rot])
        _ = task.last_loss_value()      # device -> host read of the step's loss (pinned copy made right after the CE kernel)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / steps
    sampler.stop()
    res = {"task": task, "device_step": device_step, "barrier": barrier, "B": B, "ms_step": ms_step, "ms_e2e": ms_e2e,
           "loss": loss_val, "launches": launches, "per_rank_ms": per_rank, "exchange_exposed_ms": exposed, "h2d_bytes": h2d_bytes, "clocks": sampler.summary(),
           "pages_per_s": world * B / (ms_step * 1e-3), "e2e_pages_per_s": world * B / (ms_e2e * 1e-3)}
    return res


def release(res):
    """Drop a finished config's model / optimizer / activations before the next one is built."""
    from pixparse_b200 import ops
    res.pop("task", None)
    res.pop("device_step", None)
    res.pop("barrier", None)
    ops.release_workspaces()
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def extra_train_line(cfgname, env, steps, warmup, peaks):
    c = CONFIGS[cfgname]
    res = time_train_config(cfgname, env, steps, warmup)
    world = env.world_size
    out = {"metric": f"{c['model']} {'finetune' if c['task'] == 'finetune' else 'train'} pages/sec", "value": res["pages_per_s"],
           "unit": "pages/s", "n_gpus": world, "steps": steps, "ms_per_step": res["ms_step"],
           "config": {"workload": c["workload"].format(B=res["B"]), "global_batch": world * res["B"],
                      "gflop_per_page": c["gflop_per_page"]},
           "mfu": {"vs_measured_burst": res["pages_per_s"] * c["gflop_per_page"] / 1e3 / (world * peaks["bf16_burst"]),
                   "vs_measured_sustained": res["pages_per_s"] * c["gflop_per_page"] / 1e3 / (world * peaks["bf16_sustained"])},
           "e2e": {"value": res["e2e_pages_per_s"], "unit": "pages/s", "ms_per_step": res["ms_e2e"],
                   "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": 4},
           "loss": res["loss"], "gpu_launches": res["launches"], "clocks": res["clocks"],
           "peak_mem_gb": torch.cuda.max_memory_allocated(env.device) / 2 ** 30}
    release(res)
    return out


def eval_ocr_line(env, batch=None, new_tokens=None, uncached_tokens=24, stock_reference=True):
    """BASELINE configs[4]: encoder once, then greedy decode of a FIXED number of tokens (random weights do not emit EOS
    reliably) through TaskCrullerEvalOCR's model + ocr_utils.get_generated_tokens(use_cache=True). The
    reference's own loop (whole prefix re-fed every step, ocr_utils.py:182-196) is timed beside it on a short prefix."""
    from pixparse_b200 import _lib, synthetic
    from pixparse_b200.engine import engine_for
    from pixparse_b200.ocr_utils import get_generated_tokens
    from pixparse_b200.task_eval_ocr import TaskCrullerEvalOCR, TaskCrullerEvalOCRCfg
    c = CONFIGS["eval_ocr"]
    B = batch or c["batch"]
    T = new_tokens or c["new_tokens"]
    dev = env.device
    torch.manual_seed(0)
    cfg = TaskCrullerEvalOCRCfg(model_name=c["model"])
    task = TaskCrullerEvalOCR(cfg, env, tokenizer=synthetic.SyntheticBartTokenizer())
    task.setup()
    size = tuple(cfg.model.image_encoder.image_size)
    g = torch.Generator().manual_seed(7)
    host_img = ((torch.rand((B, 1) + size, generator=g) - 0.5) / 0.5).pin_memory()

    def run(tokens, use_cache):
        img = host_img.to(dev, non_blocking=True)
        with torch.inference_mode():
            enc = task.model.image_encoder(img)
            ids = get_generated_tokens(task.model, task.tokenizer, enc, env, tokens, task.task_start_token,
                                       use_cache=use_cache, stop_on_eos=False)
        return ids

    def timed(tokens, use_cache, reps):
        run(min(tokens, 8), use_cache)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ms_enc = ms_all = 0.0
        for _ in range(reps):
            e[0].record()
            img = host_img.to(dev, non_blocking=True)
            with torch.inference_mode():
                enc = task.model.image_encoder(img)
                e[1].record()
                ids = get_generated_tokens(task.model, task.tokenizer, enc, env, tokens, task.task_start_token,
                                           use_cache=use_cache, stop_on_eos=False)
            ids_host = ids.cpu()            # the result leaves the device inside the timed region
            e[2].record()
            torch.cuda.synchronize()
            ms_enc += e[0].elapsed_time(e[1]) / reps
            ms_all += e[0].elapsed_time(e[2]) / reps
        assert ids_host.shape == (B, tokens + 1), ids_host.shape
        return ms_enc, ms_all

    sampler = ClockSampler(dev.index if dev.index is not None else 0)
    sampler.start()
    _lib.reset_launch_count()
    ms_enc, ms_all = timed(T, True, reps=2)
    launches = _lib.launch_count() // 2
    sampler.stop()
    ms_dec = ms_all - ms_enc
    _, ms_unc = timed(uncached_tokens, False, reps=1)
    _, ms_c_short = timed(uncached_tokens, True, reps=1)
    # HBM roofline of one token step: every decoder weight once (bf16), the cross-attention K | V of the image tokens,
    # the self-attention K | V of the prefix (average length T / 2); activations for 16 pages are noise next to these
    ar = engine_for(task.model.text_decoder).arena
    dcfg = task.model.text_decoder.trunk.config
    D, nl, S = dcfg.d_model, dcfg.decoder_layers, enc_tokens(size, task.model)
    w_bytes = 2 * (ar.index["dec.tok"][1] + sum(ar.index[k][1] for k in ar.keys if k.startswith("dec.") and k.endswith(".w")
                                                and ".ca.k." not in k and ".ca.v." not in k and "ln" not in k))
    kv_bytes = nl * B * (S + T / 2) * 2 * D * 2
    step_bytes = w_bytes + kv_bytes
    peaks = load_peaks()
    ref = None
    if stock_reference:
        try:
            ref = time_gpu_reference_decode(c["model"], B, uncached_tokens, dev)
        except Exception as e:      # the reference arm must never take the repo's line down
            ref = {"unavailable": repr(e)[:200]}
    out = {"metric": "cruller_large_6layers eval_ocr greedy decode tokens/sec", "value": B * T / (ms_dec * 1e-3),
           "unit": "tokens/s", "pages_per_s": B / (ms_all * 1e-3), "ms_encoder": ms_enc, "ms_decode": ms_dec,
           "ms_per_token_step": ms_dec / T, "n_gpus": 1,
           "config": {"workload": c["workload"].format(B=B, T=T), "batch": B, "new_tokens": T, "kv_cache": True,
                      "note": "timed end to end: pinned host pages -> H2D -> encoder -> decode loop -> ids back on the host; "
                              "the decode loop is one CUDA graph of single-token kernels replayed per token"},
           "e2e": {"value": B * T / (ms_all * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": host_img.numel() * 4,
                   "d2h_bytes_per_step": B * (T + 1) * 8},
           "roofline": {"bound": "hbm", "kernel": "decode step (decode_linear / decode_attention / layernorm, 77 kernels)",
                        "achieved": step_bytes / (ms_dec / T * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": step_bytes / (ms_dec / T * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                        "bytes_per_token_step": int(step_bytes), "weight_bytes": int(w_bytes), "kv_bytes": int(kv_bytes),
                        "per_kernel": "profiles/r02_ncu_hbm_kernels.txt (decode_attention 4.9 TB/s, LM head 3.1 TB/s)"},
           "uncached_reference_loop": {"new_tokens": uncached_tokens, "ms": ms_unc - ms_enc,
                                       "kv_cached_same_tokens_ms": ms_c_short - ms_enc,
                                       "what": "ocr_utils.py:182-196 as written (whole prefix re-fed each step) on this repo's "
                                               "kernels"},
           "gpu_reference": ref, "clocks": sampler.summary(),
           "gpu_launches": launches, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    if ref and "ms" in ref:
        # per generated token: the reference loop's cost grows with the prefix, so this ratio (taken over its first
        # `uncached_tokens` tokens, its cheapest) is a lower bound for longer outputs
        out["gpu_reference"]["ratio_b200_over_stock_torch"] = (ref["ms"] / max(ref["new_tokens"], 1)) / (ms_dec / T)
    del task
    release({})
    return out


def enc_tokens(size, model):
    p = model.image_encoder.trunk.arch['patch_size']
    return (size[0] // p) * (size[1] // p) + 1


def time_gpu_reference_decode(model_name, B, tokens, dev):
    """The reference's own greedy loop (utils/ocr_utils.py:165-197: whole prefix re-fed each step, no KV cache) on stock
    PyTorch: oracle modules under torch.autocast(bf16) on the same GPU. Only this reference leg touches oracle/."""
    from oracle import cruller_ref
    from pixparse_b200 import synthetic
    model = cruller_ref.build_model(model_name, vocab_size=synthetic.PRETRAIN_VOCAB, seed=0).to(dev).eval()
    size = tuple(cruller_ref.MODEL_CONFIGS[model_name].image_encoder.image_size)
    img = torch.rand((B, 1) + size, device=dev)
    with torch.inference_mode(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        enc = torch.cat([model.image_encoder(img[i:i + 4]) for i in range(0, B, 4)])
        cruller_ref.greedy_decode_uncached(model, enc, synthetic.S_PRETRAIN_ID, 1, 2, 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids = cruller_ref.greedy_decode_uncached(model, enc, synthetic.S_PRETRAIN_ID, 1, 2, tokens)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tokens = ids.shape[1] - 1
    del model
    torch.cuda.empty_cache()
    return {"ms": ms, "new_tokens": tokens, "tokens_per_s": B * tokens / (ms * 1e-3),
            "what": "oracle modules (transformers BartForCausalLM) under torch.autocast(bf16), the reference's uncached loop"}


def hbm_roofline(task, B, peaks):
    """Second roofline entry: the HBM-bound kernels of the step at the bench shapes, timed alone with CUDA events (burst
    copy peak). Algorithmic bytes per launch follow DESIGN.md section 4."""
    from pixparse_b200 import ops
    eng, opt = task.engine, task.optimizer
    ar = eng.arena
    dev = ar.device
    n = ar.total

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    peak = peaks["hbm_gbs"]
    # AdamW: read p, g, m, v (16 B) + write p, m, v, bf16 shadow, zeroed g (18 B) per parameter; scratch arenas of the
    # model's size and the optimizer's own segment table
    opt._segments()
    g0 = opt.param_groups[0]
    sp, sg, sm, sv = (torch.zeros(n, device=dev) for _ in range(4))
    s16 = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.adamw_step(sp, sg, sm, sv, s16, opt._segs, opt._nseg, lr=1e-3, beta1=g0['betas'][0],
                                       beta2=g0['betas'][1], eps=g0['eps'], step=1, norm_stats=None, zero_grad=True))
    out["adamw_kernel"] = {"bytes": 34.0 * n, "ms": ms}
    ms = timeit(lambda: ops.grad_norm(sg, max_norm=1.0))
    del sp, sg, sm, sv, s16
    out["sumsq_partial_kernel"] = {"bytes": 4.0 * n, "ms": ms}
    # cross-entropy at [B*512, V]: one bf16 read + one bf16 write per logit
    V = ar.index["dec.tok"][2][0]
    rows = B * 512
    ldv = (V + 7) // 8 * 8
    logits = torch.randn((rows, ldv), device=dev).bfloat16()
    tgt = torch.randint(3, V, (rows,), device=dev)
    dl = torch.empty_like(logits)
    ms = timeit(lambda: ops.cross_entropy(logits, tgt, V, dlogits=dl))
    out["ce_fwd_bwd_kernel"] = {"bytes": 4.0 * rows * V, "ms": ms}
    del logits, dl
    # LayerNorm at the encoder shape [B*1009, 768]: fwd reads fp32, writes bf16; bwd reads x fp32, dy bf16, residual-path
    # fp32 gradient, writes dx fp32 + bf16
    M, D = B * 1009, 768
    x = torch.randn((M, D), device=dev)
    w = torch.ones(D, device=dev)
    b = torch.zeros(D, device=dev)
    ms = timeit(lambda: ops.layernorm_fwd(x, w, b, 1e-6))
    out["layernorm_fwd_kernel"] = {"bytes": 6.0 * M * D, "ms": ms}
    _, _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
    dy = torch.randn((M, D), device=dev).bfloat16()
    dres = torch.randn((M, D), device=dev)
    dx32, dx16 = torch.empty_like(x), torch.empty_like(dy)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    ms = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, w, dg, db, dy16=dy, dres32=dres, dx32=dx32, dx16=dx16))
    out["layernorm_bwd_kernel"] = {"bytes": (4 + 2 + 4 + 4 + 2.0) * M * D, "ms": ms}
    for k, v in out.items():
        v["achieved"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        v["frac"] = v["achieved"] / peak
        v["bytes"] = int(v["bytes"])
    return {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peaks["source"] + ", copy bandwidth (burst: kernels timed alone)",
            "kernels": out, "how": "each kernel alone at the bench shapes, 5 launches between CUDA events; algorithmic bytes per "
                                   "DESIGN.md section 4; ncu dram__bytes for the same launches: profiles/r02_ncu_hbm_kernels.txt"}


def run_b200_arm(args):
    import torch.distributed as dist
    from pixparse_b200 import _lib
    from pixparse_b200.framework import DeviceEnv

    env = DeviceEnv()
    rank, world = env.global_rank, env.world_size
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE is {world} (launch N > 1 with torchrun)"
    dev = env.device
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
        env.device = dev
    torch.cuda.set_device(dev)
    peaks = load_peaks()

    if args.config == "eval_ocr":
        line = eval_ocr_line(env, args.batch)
        line.update({"steps": 2, "warmup": 1, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                     "dtype": "bf16", "data": "synthetic"})
        if rank == 0:
            emit_json(line)
        return 0

    c = CONFIGS[args.config]
    gflop = c["gflop_per_page"]
    res = time_train_config(args.config, env, args.steps, args.warmup, args.batch, args.dropout_off)
    task, device_step, barrier = res["task"], res["device_step"], res["barrier"]
    B, ms_step, pages_per_s = res["B"], res["ms_step"], res["pages_per_s"]

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), CUDA events around every launch -----------------
    prof = _lib.profile_ops(lambda: device_step(0), names=("b200_gemm_bf16", "b200_attention_fwd",
                                                           "b200_attention_bwd"), repeats=2)
    barrier()
    if args.profile_all and rank == 0:
        allp = _lib.profile_ops(lambda: device_step(0), names=tuple(_lib.SIGNATURES.keys()), repeats=2)
        rows = sorted(((v["ms"], n, v["calls"]) for n, v in allp.items() if v["calls"]), reverse=True)
        with open(args.profile_all, "w") as fh:
            fh.write(f"per-step device time by entry point (CUDA events around each call; step = {ms_step:.2f} ms)\n")
            for ms, n, calls in rows:
                fh.write(f"{ms:9.3f} ms  {100 * ms / ms_step:5.1f}%  calls={calls:4d}  avg={1e3 * ms / calls:8.1f} us  {n}\n")
            fh.write(f"{sum(r[0] for r in rows):9.3f} ms  total of the above\n")
            for n in ("b200_gemm_bf16", "b200_attention_fwd", "b200_attention_bwd"):
                fh.write(f"\n{n} by variant (ms per step, TFLOP/s, calls):\n")
                for tag, v in sorted(allp[n]["detail"].items(), key=lambda kv: -kv[1]["ms"]):
                    fh.write(f"{v['ms']:9.3f} ms  {str(v['tflops']):>7} TF/s  calls={v['calls']:3d}  {tag}\n")
    gemm_ms, gemm_flops, gemm_n = prof["b200_gemm_bf16"]["ms"], prof["b200_gemm_bf16"]["flops"], prof["b200_gemm_bf16"]["calls"]
    att_ms = prof["b200_attention_fwd"]["ms"] + prof["b200_attention_bwd"]["ms"]
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM traffic per launch from the committed ncu --set full captures (profiles/), weighted by this step's call mix
    traffic, traffic_detail = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_gemm_dram_traffic.json")) as fh:
            cap = json.load(fh)["variants"]
        tot_b, tot_c = 0.0, 0
        traffic_detail = {}
        for tag, det in prof["b200_gemm_bf16"]["detail"].items():
            if tag in cap and args.config == "base":
                b = cap[tag]["dram_read_bytes"] + cap[tag]["dram_write_bytes"]
                traffic_detail[tag] = {"dram_bytes_per_launch": b, "algorithmic_bytes": cap[tag]["algorithmic_bytes"],
                                       "shape": cap[tag]["shape"]}
                tot_b += b * det["calls"]
                tot_c += det["calls"]
        if tot_c:
            traffic = tot_b / tot_c
            traffic_detail["note"] = ("mean DRAM bytes per launch over the captured variants (each at its encoder shape), "
                                      "weighted by calls per step; source profiles/r02_gemm_dram_traffic.json")
    except (OSError, KeyError, ValueError):
        pass
    # attention: algorithmic FLOPs of the step's attention calls (full S x S / T x T / T x S, bwd = 2.5 x fwd)
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05, all layouts/epilogues)", "achieved": achieved,
                "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "launches_per_step": gemm_n, "gemm_ms_per_step": gemm_ms, "attention_ms_per_step": att_ms,
                "share_of_step": gemm_ms / ms_step,
                "by_variant": prof["b200_gemm_bf16"]["detail"],
                "attention_fwd_ms": prof["b200_attention_fwd"]["ms"], "attention_bwd_ms": prof["b200_attention_bwd"]["ms"],
                "how": "2 extra steps right after the timed region with CUDA events around each launch on the launch stream"}
    roofline_hbm = None
    if rank == 0 and world == 1 and args.config == "base":
        roofline_hbm = hbm_roofline(task, B, peaks)

    dcfg = task.model.text_decoder.trunk.config
    mfu_burst = pages_per_s * gflop / 1e3 / (world * peaks["bf16_burst"])
    line = {
        "metric": f"{c['model']} train pages/sec", "value": pages_per_s, "unit": "pages/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": c["workload"].format(B=B),
                   "global_batch": world * B, "seq_len": c["text_len"] - 1, "parallelism": f"dp{world}",
                   "l2": "working set (activations >> 126 MB) is larger than L2; 4 distinct batches rotate",
                   "dropout": {"decoder": dcfg.dropout, "attention": dcfg.attention_dropout,
                               "activation": dcfg.activation_dropout, "encoder": 0.0,
                               "note": "live in train_step as in the reference (public bart config values)"},
                   "gflop_per_page": gflop},
        "mfu": {"vs_measured_burst": mfu_burst,
                "vs_measured_sustained": pages_per_s * gflop / 1e3 / (world * peaks["bf16_sustained"]),
                "vs_nominal_2250": pages_per_s * gflop / 1e3 / (world * 2250.0)},
        "loss": res["loss"],
        "clocks": res["clocks"],
        "per_rank_ms": res["per_rank_ms"],
        "exchange_exposed_ms": res["exchange_exposed_ms"],      # compute stream waiting for the gradient exchange after backward
        "e2e": {"value": res["e2e_pages_per_s"], "unit": "pages/s", "ms_per_step": res["ms_e2e"],
                "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": 4},
        "gpu_launches": res["launches"],
        "roofline": roofline,
    }
    if roofline_hbm is not None:
        line["roofline_hbm"] = roofline_hbm
    release(res)
    del task, device_step

    # ---- the other BASELINE configs as extra keys of the same line (the headline stays configs[1]) ----------------
    if args.config == "base" and not args.no_extras:
        extras = {}
        ksteps = max(3, min(args.steps, 6))
        try:
            extras["large"] = extra_train_line("large", env, ksteps, 3, peaks)            # configs[2], every N (DDP)
        except Exception as e:      # an extra must never take the headline line down with it
            extras["large"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        if world == 1:
            for name, fn in (("finetune", lambda: extra_train_line("finetune", env, ksteps, 3, peaks)),
                             ("eval_ocr", lambda: eval_ocr_line(env))):
                try:
                    extras[name] = fn()
                except Exception as e:
                    extras[name] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        line["extra_configs"] = extras
    if rank == 0 and world == 1 and args.config == "base" and not args.no_gpu_reference:
        try:
            ref = time_gpu_reference("base", steps=max(3, min(args.steps, 8)), warmup=3, dev=dev)
            ref["ratio_b200_over_stock_torch"] = pages_per_s / ref["value"]
            line["gpu_reference"] = ref
        except Exception as e:
            line["gpu_reference"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        pps, sec, info = time_cpu_reference(steps=6, warmup=1, batch=2, budget_s=25.0)
        line["cpu_baseline"] = {"value": pps, "unit": "pages/s", "cores": info["cores"], "kind": "port",
                                "sample": f"{info['steps']} train steps of 2 pages (of {B}) after 1 warm-up, T=512, fp32, all "
                                          f"host threads: {sec:.1f} s/step (setup+steps {time.perf_counter() - t0:.0f} s)"}
    if rank == 0:
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit_json(line):
    """The one JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # keep stdout to the single JSON line: NCCL prints its version banner to stdout at communicator creation on some
    # boxes (NCCL_DEBUG exported by the environment), and libraries may do the same. Everything that writes to fd 1 is
    # sent to stderr for the life of the process; the JSON line goes to the saved original stdout. (NCCL's own logging
    # settings are left alone.)
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference_gpu"])
    ap.add_argument("--config", default="base", choices=sorted(CONFIGS),
                    help="workload: base = BASELINE configs[1] (the headline); large / finetune / eval_ocr = configs[2..4]")
    ap.add_argument("--batch", type=int, default=None, help="pages per GPU (default: the config's own, base = 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_configs keys (configs[2..4])")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the stock-PyTorch GPU reference key")
    ap.add_argument("--dropout-off", action="store_true", help="diagnostic only: the reference trains with dropout 0.1")
    ap.add_argument("--profile-all", default=None, metavar="FILE",
                    help="diagnostic: also time every C-ABI entry point of one step (CUDA events) and write the table to FILE")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "reference_gpu":
        return run_reference_gpu_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
