"""Time the attention kernels at the three shapes of BASELINE config 2 (encoder, decoder self, decoder cross)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops, _lib
if os.environ.get("QUERY_MAJOR"):
    _lib.lib().b200_debug_attention_bwd_query_major(1)
B, H = 32, 12
D = H * 64
iters = int(os.environ.get("ITERS", "10"))
torch.manual_seed(0)

def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [("encoder", 1009, 1009, False, None), ("dec-self", 512, 512, True, (0.1, 1234)), ("dec-cross", 512, 1009, False, (0.1, 1234))]
if os.environ.get("NODROP"):
    shapes = [(n, a, b_, c, None) for n, a, b_, c, d in shapes]
if os.environ.get("ONLY"):
    shapes = [s for s in shapes if s[0] == os.environ["ONLY"]]
for name, Sq, Sk, causal, drop in shapes:
    q = (torch.randn((B * Sq, D), device="cuda") * 0.5).bfloat16()
    kv = (torch.randn((B * Sk, 2 * D), device="cuda") * 0.5).bfloat16()
    dout = torch.randn((B * Sq, D), device="cuda").bfloat16()
    dq = torch.empty_like(q); dkv = torch.empty_like(kv)
    kw = dict(B=B, H=H, Sq=Sq, Sk=Sk, q_col0=0, k_col0=0, v_col0=D, causal=causal, drop=drop)
    fwd = lambda: ops.attention_fwd(q, kv, kv, **kw)
    out, lse = fwd()
    bwd = lambda: ops.attention_bwd(q, kv, kv, out, dout, lse, dq, dkv, dkv, dq_col0=0, dk_col0=0, dv_col0=D, **kw)
    frac = 0.5 if causal else 1.0
    for kind, fn, fl in (("fwd", fwd, 4.0), ("bwd", bwd, 10.0)):
        ms = timeit(fn)
        print(f"{name:9s} {kind}: {ms:.3f} ms  {fl * frac * B * H * Sq * Sk * 64 / ms / 1e9:.0f} TFLOP/s", flush=True)
