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
    sec = sum(times) / len(times)
    return batch / sec, sec, {"cores": cores, "batch": batch}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: 1 page per step keeps K + W steps within a few minutes on the host cores
    batch = 1
    pps, sec, info = time_cpu_reference(args.steps, args.warmup, batch)
    line = {
        "impl": "reference", "metric": "cruller_base train pages/sec", "value": pps, "unit": "pages/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cruller_base pretrain fwd+bwd+AdamW step, 576x448 grayscale pages, 512-token targets",
                   "pages_per_step": batch, "device": "host CPU"},
        "cpu_baseline": {"value": pps, "unit": "pages/s", "cores": info["cores"], "kind": "port",
                         "sample": f"{args.steps} train steps of {batch} page(s) (T=512, fp32, all host threads); "
                                   "oracle = restated timm ViT + installed transformers BART + torch AdamW"},
        "e2e": {"value": pps, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch.distributed as dist
    from pixparse_b200 import _lib, synthetic
    from pixparse_b200.framework import DeviceEnv, OptimizationCfg
    from pixparse_b200.task_pretrain import TaskCrullerPretrain, TaskCrullerPretrainCfg

    env = DeviceEnv()
    rank, world = env.global_rank, env.world_size
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE is {world} (launch N > 1 with torchrun)"
    dev = env.device
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    torch.cuda.set_device(dev)

    opt = OptimizationCfg(learning_rate=3e-4, betas=(0.9, 0.98), eps=1e-6, clip_grad_value=1.0,
                          clip_grad_mode='norm', grad_accum_steps=1)
    cfg = TaskCrullerPretrainCfg(model_name=MODEL, opt=opt, dtype='bfloat16', amp=True, eval_frequency=10 ** 9,
                                 num_intervals=100, num_warmup_intervals=5)
    torch.manual_seed(0)
    task = TaskCrullerPretrain(cfg, env, monitor=None, tokenizer=synthetic.SyntheticBartTokenizer())
    assert task.vocab_size == synthetic.PRETRAIN_VOCAB
    if args.dropout_off:
        task.model.text_decoder.trunk.set_dropout(0.0)
    task.train_setup(num_batches_per_interval=1000)
    task.train_interval_start()
    B = args.batch
    size = tuple(cfg.model.image_encoder.image_size)

    # a small rotation of distinct host batches (seed + rank as framework/random.py:8-11 does)
    n_rot = 4
    host = []
    for i in range(n_rot):
        img, txt, tgt = synthetic.synthetic_batch(B, size, TEXT_LEN, seed=100 * rank + i)
        host.append((img.pin_memory(), txt.pin_memory(), tgt.pin_memory()))
    resident = [(h[0].to(dev), h[1][:, :-1].contiguous().to(dev), h[2][:, 1:].contiguous().to(dev)) for h in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def device_step(i):
        img, txt, tgt = resident[i % n_rot]
        if task.reducer is not None:
            task.reducer.enabled = True
            task.reducer.begin()
        stats = task.engine.forward_backward(img, txt, tgt)
        if task.reducer is not None:
            task.reducer.finish()
        task.optimizer.step(clip_grad_norm=opt.clip_grad_value)
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

    # ---- device-resident timing -------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        device_step(i)
    barrier()
    sampler = ClockSampler(dev.index if dev.index is not None else 0)
    sampler.start()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(args.steps):
        last = device_step(i)
    e1.record()
    barrier()
    launches = _lib.launch_count()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    loss_val = float(last[1].item())
    pages_per_s = world * B / (ms_step * 1e-3)

    # ---- end to end through the Task API: pinned host batch -> train_step -> loss read back ---------------------
    for i in range(2):
        task.train_step(host[i % n_rot])
    barrier()
    e0.record()
    for i in range(args.steps):
        task.train_step(host[i % n_rot])
        _ = task.last_loss_value()            # device -> host read of the step's loss (pinned copy made right after the CE kernel)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    sampler.stop()
    e2e_pps = world * B / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), CUDA events around every launch -----------------
    peaks = load_peaks()
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
    gemm_ms, gemm_flops, gemm_n = prof["b200_gemm_bf16"]["ms"], prof["b200_gemm_bf16"]["flops"], prof["b200_gemm_bf16"]["calls"]
    att_ms = prof["b200_attention_fwd"]["ms"] + prof["b200_attention_bwd"]["ms"]
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM traffic per launch from the committed ncu --set full captures (profiles/), weighted by this step's call mix
    traffic, traffic_detail = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_gemm_dram_traffic.json")) as fh:
            cap = json.load(fh)["variants"]
        tot_b, tot_c = 0.0, 0
        traffic_detail = {}
        for tag, det in prof["b200_gemm_bf16"]["detail"].items():
            if tag in cap:
                b = cap[tag]["dram_read_bytes"] + cap[tag]["dram_write_bytes"]
                traffic_detail[tag] = {"dram_bytes_per_launch": b, "algorithmic_bytes": cap[tag]["algorithmic_bytes"],
                                       "shape": cap[tag]["shape"]}
                tot_b += b * det["calls"]
                tot_c += det["calls"]
        if tot_c:
            traffic = tot_b / tot_c
            traffic_detail["note"] = ("mean DRAM bytes per launch over the captured variants (each at its encoder shape), "
                                      "weighted by calls per step; source profiles/r01_gemm_dram_traffic.json")
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05, all layouts/epilogues)", "achieved": achieved,
                "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "launches_per_step": gemm_n, "gemm_ms_per_step": gemm_ms, "attention_ms_per_step": att_ms,
                "share_of_step": gemm_ms / ms_step,
                "by_variant": prof["b200_gemm_bf16"]["detail"],
                "attention_fwd_ms": prof["b200_attention_fwd"]["ms"], "attention_bwd_ms": prof["b200_attention_bwd"]["ms"],
                "how": "2 extra steps right after the timed region with CUDA events around each launch on the launch stream"}

    mfu_burst = pages_per_s * GFLOP_PER_PAGE / 1e3 / (world * peaks["bf16_burst"])
    line = {
        "metric": "cruller_base train pages/sec", "value": pages_per_s, "unit": "pages/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cruller_base pretrain step (fwd+CE+bwd+allreduce+clip+AdamW), bf16, "
                               f"{B} synthetic 576x448 grayscale pages + 512-token targets per GPU",
                   "global_batch": world * B, "seq_len": 512, "parallelism": f"dp{world}",
                   "l2": "working set (activations >> 126 MB) is larger than L2; 4 distinct batches rotate",
                   "dropout": {"decoder": task.model.text_decoder.trunk.config.dropout,
                               "attention": task.model.text_decoder.trunk.config.attention_dropout,
                               "activation": task.model.text_decoder.trunk.config.activation_dropout,
                               "encoder": 0.0, "note": "live in train_step as in the reference (bart-base config)"},
                   "gflop_per_page": GFLOP_PER_PAGE},
        "mfu": {"vs_measured_burst": mfu_burst,
                "vs_measured_sustained": pages_per_s * GFLOP_PER_PAGE / 1e3 / (world * peaks["bf16_sustained"]),
                "vs_nominal_2250": pages_per_s * GFLOP_PER_PAGE / 1e3 / (world * 2250.0)},
        "loss": loss_val,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_pps, "unit": "pages/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        pps, sec, info = time_cpu_reference(steps=1, warmup=0, batch=2)
        line["cpu_baseline"] = {"value": pps, "unit": "pages/s", "cores": info["cores"], "kind": "port",
                                "sample": f"1 train step of 2 pages (of {B}), T=512, fp32, all host threads: "
                                          f"{sec:.1f} s (setup+step {time.perf_counter() - t0:.0f} s)"}
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
    # sent to stderr for the life of the process; the JSON line goes to the saved original stdout.
    global _REAL_STDOUT
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="pages per GPU (BASELINE config: 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout-off", action="store_true", help="diagnostic only: the reference trains with dropout 0.1")
    ap.add_argument("--profile-all", default=None, metavar="FILE",
                    help="diagnostic: also time every C-ABI entry point of one step (CUDA events) and write the table to FILE")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
