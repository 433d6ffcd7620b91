"""Run one GEMM variant at a train-step shape (for ncu captures). usage: gpu_gemm_one.py <gelu|dgelu|resid|store|wgrad>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops, _lib
if os.environ.get("SINGLE_CTA"):      # 1 = never CTA pairs, -1 = CTA pairs for every epilogue
    _lib.lib().b200_debug_gemm_single_cta(int(os.environ["SINGLE_CTA"]))
which = sys.argv[1] if len(sys.argv) > 1 else "gelu"
M, D, F = 32 * 1009, 768, 3072
torch.manual_seed(0)
if which == "gelu":
    A = torch.randn((M, D), device="cuda").bfloat16(); B = torch.randn((F, D), device="cuda").bfloat16() * 0.05
    bias = torch.randn(F, device="cuda"); h = torch.empty((M, F), device="cuda", dtype=torch.bfloat16); g = torch.empty_like(h)
    fn = lambda: ops.gemm(A, B, epi=ops.EPI_GELU_BF16, bias=bias, out=g, out2=h)
elif which == "dgelu":
    A = torch.randn((M, D), device="cuda").bfloat16(); B = torch.randn((D, F), device="cuda").bfloat16() * 0.05
    h = torch.randn((M, F), device="cuda").bfloat16(); o = torch.empty_like(h)
    fn = lambda: ops.gemm(A, B, b_mn=True, epi=ops.EPI_DGELU_BF16, aux=h, out=o)
elif which == "resid":
    A = torch.randn((M, D), device="cuda").bfloat16(); B = torch.randn((D, D), device="cuda").bfloat16() * 0.05
    x = torch.randn((M, D), device="cuda"); o = torch.empty_like(x); bias = torch.randn(D, device="cuda")
    fn = lambda: ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x, out=o)
elif which == "resid3072":      # fc2 forward: K = 3072, fp32 residual in / out
    A = torch.randn((M, F), device="cuda").bfloat16(); B = torch.randn((D, F), device="cuda").bfloat16() * 0.05
    x = torch.randn((M, D), device="cuda"); o = torch.empty_like(x); bias = torch.randn(D, device="cuda")
    fn = lambda: ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x, out=o)
elif which == "dgrad3072":      # fc1 dgrad: same M, N, K as resid3072 with a plain bf16 store
    A = torch.randn((M, F), device="cuda").bfloat16(); B = torch.randn((F, D), device="cuda").bfloat16() * 0.05
    o = torch.empty((M, D), device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(A, B, b_mn=True, out=o)
elif which == "store3072":      # the GELU GEMM's shape with a plain bf16 store: isolates the cost of the activation epilogue
    A = torch.randn((M, D), device="cuda").bfloat16(); B = torch.randn((F, D), device="cuda").bfloat16() * 0.05
    bias = torch.randn(F, device="cuda"); o = torch.empty((M, F), device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias, out=o)
elif which == "store":
    A = torch.randn((M, D), device="cuda").bfloat16(); B = torch.randn((3 * D, D), device="cuda").bfloat16() * 0.05
    bias = torch.randn(3 * D, device="cuda"); o = torch.empty((M, 3 * D), device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias, out=o)
else:
    NO = {"wgrad": F, "wgrad_qkv": 3 * D, "wgrad_proj": D}[which]       # dW [NO, D] = dy[M, NO]^T x[M, D]
    A = torch.randn((M, NO), device="cuda").bfloat16(); B = torch.randn((M, D), device="cuda").bfloat16()
    o = torch.zeros((NO, D), device="cuda")
    fn = lambda: ops.gemm(A, B, a_mn=True, b_mn=True, epi=ops.EPI_REDUCE_F32, out=o)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
print(which, e0.elapsed_time(e1) / 5, "ms")
