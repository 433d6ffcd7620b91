"""LM head (decode_linear, 50267 x 1024) and cross-attention (decode_attention, 2509 keys) alone, L2 flushed between launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops
dev = "cuda"
torch.manual_seed(0)
xd = torch.randn((16, 1024), device=dev).bfloat16()
wv = torch.randn((50267, 1024), device=dev).bfloat16()
part = torch.zeros((16, ops.decode_linear_ctas(50267)), device=dev, dtype=torch.int64)
kv = torch.randn((16 * 2509, 2048), device=dev).bfloat16()
o = torch.empty((16, 1024), device=dev, dtype=torch.bfloat16)
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
t = timeit(lambda: ops.decode_linear(xd, wv, M=16, argmax_partial=part))
print(f"G={os.environ.get('PIXPARSE_B200_DECODE_G', 'auto')} LM head: {t:.1f} us = {50267 * 1024 * 2 / t / 1e3:.0f} GB/s")
t = timeit(lambda: ops.decode_attention(xd, kv, kv, o, B=16, H=16, ld_kv=2048, kv_bstride=2509 * 2048, v_col0=1024, sk=2509))
print(f"splits={os.environ.get('PIXPARSE_B200_DECODE_SPLITS', 'auto')} cross-attention: {t:.1f} us = {16 * 2509 * 2048 * 2 / t / 1e3:.0f} GB/s")
