"""One launch of every HBM-bound kernel of the step at the BASELINE configs[1] shapes (for `ncu --set full`), plus the
decode-step kernels of configs[4] (LM head fused with argmax, cross-attention over 2509 image tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops

dev = "cuda"
B = 32
n = 163_229_184                       # cruller_base arena (parameters, 64-element aligned)
torch.manual_seed(0)
# AdamW (one segment: lr_scale 1, wd 0 like the pretrain task) + grad norm
sp, sg, sm, sv = (torch.zeros(n, device=dev) for _ in range(4))
s16 = torch.zeros(n, device=dev, dtype=torch.bfloat16)
import numpy as np
from pixparse_b200.optim import _SEG_DTYPE
arr = np.zeros(1, dtype=_SEG_DTYPE)
arr[0] = (n, 1.0, 0.0)
seg_t, nseg = torch.from_numpy(arr.view(np.uint8).copy()).to(dev), 1
ops.adamw_step(sp, sg, sm, sv, s16, seg_t, nseg, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, step=1, norm_stats=None, zero_grad=True)
ops.grad_norm(sg, max_norm=1.0)
del sp, sg, sm, sv, s16
# cross-entropy
V = 50267; rows = B * 512; ldv = (V + 7) // 8 * 8
logits = torch.randn((rows, ldv), device=dev).bfloat16()
tgt = torch.randint(3, V, (rows,), device=dev)
dl = torch.empty_like(logits)
ops.cross_entropy(logits, tgt, V, dlogits=dl)
del logits, dl
# LayerNorm
M, D = B * 1009, 768
x = torch.randn((M, D), device=dev); w = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
_, _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
dy = torch.randn((M, D), device=dev).bfloat16(); dres = torch.randn((M, D), device=dev)
dx32, dx16 = torch.empty_like(x), torch.empty_like(dy)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
ops.layernorm_bwd(x, mean, rstd, w, dg, db, dy16=dy, dres32=dres, dx32=dx32, dx16=dx16)
# page preprocessing: 32 uint8 letter-size scans -> 576 x 448
pages = torch.randint(0, 255, (B, 1100, 850), device=dev, dtype=torch.uint8)
ops.preprocess_pages(pages, (576, 448), 0.5, 0.5)
# gradient-exchange reduction: 8 ranks, one 64 MB bucket -> 8 MB share
c = 2 * 1024 * 1024
ops.reduce_shards(torch.zeros(c, device=dev), torch.zeros(8 * c, device=dev), c, 8, 3, 0.125)
# decode step (cruller_large_6layers, 16 pages): LM head + argmax, cross-attention, a 1024 x 1024 linear
xd = torch.randn((16, 1024), device=dev).bfloat16()
wv = torch.randn((50267, 1024), device=dev).bfloat16()
part = torch.zeros((16, ops.decode_linear_ctas(50267)), device=dev, dtype=torch.int64)
ops.decode_linear(xd, wv, M=16, argmax_partial=part)
kv = torch.randn((16 * 2509, 2048), device=dev).bfloat16()
o = torch.empty((16, 1024), device=dev, dtype=torch.bfloat16)
ops.decode_attention(xd, kv, kv, o, B=16, H=16, ld_kv=2048, kv_bstride=2509 * 2048, v_col0=1024, sk=2509)
w1 = torch.randn((1024, 1024), device=dev).bfloat16()
ops.decode_linear(xd, w1, M=16, out16=o)
torch.cuda.synchronize()
print("done")
