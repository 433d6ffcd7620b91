"""Run the attention kernels at the encoder shape of BASELINE config 2 (for ncu captures) and time them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops
B, H, S = 32, 12, 1009
D = H * 64
torch.manual_seed(0)
qkv = (torch.randn((B * S, 3 * D), device="cuda") * 0.5).bfloat16()
dout = torch.randn((B * S, D), device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
def fwd():
    return ops.attention_fwd(qkv, qkv, qkv, B=B, H=H, Sq=S, Sk=S, q_col0=0, k_col0=D, v_col0=2 * D)
out, lse = fwd()
def bwd():
    ops.attention_bwd(qkv, qkv, qkv, out, dout, lse, dqkv, dqkv, dqkv, B=B, H=H, Sq=S, Sk=S, q_col0=0, k_col0=D,
                      v_col0=2 * D, dq_col0=0, dk_col0=D, dv_col0=2 * D)
iters = int(os.environ.get("ITERS", "10"))
from pixparse_b200 import _lib
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, fn, flops in (("fwd", fwd, 4.0 * B * H * S * S * 64), ("bwd", bwd, 10.0 * B * H * S * S * 64)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"attention {name}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
