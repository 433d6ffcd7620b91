"""Bandwidth of the memory-bound helper kernels at the encoder shapes of BASELINE config 2 (bytes moved / time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops
M, D = 32 * 1009, 768
torch.manual_seed(0)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for N in (768, 2304, 3072):
    dy = torch.randn((M, N), device="cuda").bfloat16(); out = torch.zeros(N, device="cuda")
    ms = timeit(lambda: ops.colsum(dy, out))
    print(f"colsum N={N}: {ms * 1e3:.1f} us  {M * N * 2 / ms / 1e6:.0f} GB/s", flush=True)
x = torch.randn((M, D), device="cuda"); g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
y16, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
y16o = torch.empty_like(y16)
ms = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-6))
print(f"layernorm_fwd: {ms * 1e3:.1f} us  {(M * D * 6) / ms / 1e6:.0f} GB/s (x fp32 in, bf16 out; includes 2 torch.empty)", flush=True)
dy16 = torch.randn((M, D), device="cuda").bfloat16(); dres = torch.randn((M, D), device="cuda")
dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
dx32 = torch.empty((M, D), device="cuda"); dx16 = torch.empty((M, D), device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, g, dg, db, dy16=dy16, dres32=dres, dx32=dx32, want_bf16=False))
print(f"layernorm_bwd (dy16 + dres32 -> dx32): {ms * 1e3:.1f} us  {(M * D * 14) / ms / 1e6:.0f} GB/s", flush=True)
ms = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, g, dg, db, dy16=dy16, dres32=dres, dx32=dx32, dx16=dx16))
print(f"layernorm_bwd (dy16 + dres32 -> dx32 + dx16): {ms * 1e3:.1f} us  {(M * D * 16) / ms / 1e6:.0f} GB/s", flush=True)
ms = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, g, dg, db, dy16=dy16, dx32=dx32, want_bf16=False))
print(f"layernorm_bwd (dy16 -> dx32): {ms * 1e3:.1f} us  {(M * D * 10) / ms / 1e6:.0f} GB/s", flush=True)
# cross-entropy at the LM-head shape (16384 rows, V = 50267)
V, R = 50267, 16384
ldv = (V + 7) // 8 * 8
logits = torch.randn((R, ldv), device="cuda").bfloat16()
tgt = torch.randint(0, V, (R,), device="cuda")
dl = torch.empty_like(logits)
ms = timeit(lambda: ops.cross_entropy(logits, tgt, V, dlogits=dl), iters=10)
print(f"cross_entropy fwd+bwd: {ms * 1e3:.1f} us  {(R * ldv * 4) / ms / 1e6:.0f} GB/s", flush=True)
