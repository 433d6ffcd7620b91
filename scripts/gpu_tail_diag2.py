"""Which term of the residual epilogue is wrong in a split tail tile?"""
import sys
import torch
sys.path.insert(0, ".")
from pixparse_b200 import ops, _lib

DEV = "cuda"
M, N, K = 6656, 768, 2048
torch.manual_seed(4)
A = torch.randn((M, K), device=DEV).bfloat16()
B = torch.randn((N, K), device=DEV).bfloat16()
bias = torch.randn(N, device=DEV)
acc = A.float() @ B.float().t()
x0 = torch.randn((M, N), device=DEV)
zero = torch.zeros_like(x0)
# integer-coded aux: value = row * 1000 + col  (exact in fp32 up to 2^24)
rows = torch.arange(M, device=DEV, dtype=torch.float32)[:, None]
cols = torch.arange(N, device=DEV, dtype=torch.float32)[None, :]
coded = rows * 1000 + cols


def run(name, Ause, bias_use, aux):
    y = torch.empty_like(x0)
    ops.gemm(Ause, B, epi=ops.EPI_RESID_F32, bias=bias_use, aux=aux, out=y)
    torch.cuda.synchronize()
    ref = aux + (Ause.float() @ B.float().t()) + (bias_use if bias_use is not None else 0)
    d = (y - ref).abs()
    bad = d > 1e-2
    nb = int(bad.sum().item())
    print(f"{name}: bad {nb} max {d.max().item():.4g}", flush=True)
    if nb:
        idx = bad.nonzero()[:8]
        for r, c in idx.tolist():
            print(f"   [{r},{c}] got {y[r, c].item():.6g} want {ref[r, c].item():.6g} aux {aux[r, c].item():.6g}")


A0 = torch.zeros_like(A)
run("acc=0 bias=None aux=coded", A0, None, coded)
run("acc=0 bias aux=0", A0, bias, zero)
run("acc bias=None aux=0", A, None, zero)
run("full", A, bias, x0)
