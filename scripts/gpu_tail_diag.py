"""Diagnostic for the GEMM tail split: per variant, where the output differs from the fp32 reference and whether the
workspace is left zeroed."""
import sys
import torch
sys.path.insert(0, ".")
from pixparse_b200 import ops, _lib

DEV = "cuda"
M, N, K = 6656, 768, 2048
torch.manual_seed(4)
A = torch.randn((M, K), device=DEV).bfloat16()
B = torch.randn((N, K), device=DEV).bfloat16()
Bt = B.t().contiguous()
bias = torch.randn(N, device=DEV)
acc = A.float() @ B.float().t()
x0 = torch.randn((M, N), device=DEV)


def report(name, got, ref, tol):
    torch.cuda.synchronize()
    d = (got.float() - ref).abs()
    bad = d > tol
    nb = int(bad.sum().item())
    ws, _ = ops._tail_workspace(A.device)
    nz = int(ws.count_nonzero().item())
    msg = f"{name}: max err {d.max().item():.4g}, bad {nb}, ws nonzero {nz}"
    if nb:
        idx = bad.nonzero()
        rows, cols = idx[:, 0], idx[:, 1]
        msg += f" rows [{rows.min().item()}, {rows.max().item()}] cols [{cols.min().item()}, {cols.max().item()}]"
        msg += f" distinct rows {rows.unique().numel()} distinct cols {cols.unique().numel()}; first {idx[:6].tolist()}"
        msg += f" col%32 set {sorted(set((cols % 32).tolist()))[:40]} row%128 set size {len(set((rows % 128).tolist()))}"
    print(msg, flush=True)
    if nz:
        ws.zero_()


for rep in range(2):
    out = ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias)
    report(f"rep{rep} store", out, acc + bias, 0.5)
    out = ops.gemm(A, Bt, b_mn=True, epi=ops.EPI_STORE_BF16)
    report(f"rep{rep} store b_mn", out, acc, 0.5)
    y = torch.empty_like(x0)
    ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x0, out=y)
    report(f"rep{rep} resid out-of-place", y, x0 + acc + bias, 1e-3)
    x = x0.clone()
    ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x, out=x)
    report(f"rep{rep} resid in-place", x, x0 + acc + bias, 1e-3)
    out = ops.gemm(A, B, epi=ops.EPI_STORE_F32, bias=bias)
    report(f"rep{rep} store_f32 (no tail split)", out, acc + bias, 1e-3)
_lib.lib().b200_debug_gemm_tail_split(0)
y = torch.empty_like(x0)
ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x0, out=y)
report("no split resid", y, x0 + acc + bias, 1e-3)
