"""cruller_large_6layers graph decode of N tokens for 16 pages (for `ncu -k regex:decode_` launch lists / full captures).
Prints the event-timed ms per token step as well."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import synthetic
from pixparse_b200.framework import DeviceEnv
from pixparse_b200.ocr_utils import get_generated_tokens
from pixparse_b200.task_eval_ocr import TaskCrullerEvalOCR, TaskCrullerEvalOCRCfg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
env = DeviceEnv()
task = TaskCrullerEvalOCR(TaskCrullerEvalOCRCfg(model_name="cruller_large_6layers"), env,
                          tokenizer=synthetic.SyntheticBartTokenizer())
task.setup()
size = tuple(task.cfg.model.image_encoder.image_size)
img = torch.rand((B, 1) + size, device=env.device)
with torch.inference_mode():
    enc = task.model.image_encoder(img)
    run = lambda n: get_generated_tokens(task.model, task.tokenizer, enc, env, n, "<s_pretrain>", use_cache=True,
                                         stop_on_eos=False)
    run(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids = run(N)
    e1.record()
    torch.cuda.synchronize()
print(f"{N} tokens x {B} pages: {e0.elapsed_time(e1):.3f} ms = {e0.elapsed_time(e1) / N * 1e3:.1f} us per token step "
      f"({B * N / e0.elapsed_time(e1) * 1e3:.0f} tokens/s)")
