"""Thin tensor-level wrappers over the C-ABI (one Python function per entry point).

Every function enqueues on torch's current CUDA stream and returns immediately. Tensors must be CUDA,
contiguous in their last dimension; outputs are allocated by the caller or here through torch's
caching allocator (the kernels never allocate).
"""
import torch

from . import _lib
from ._lib import (EPI_DGELU_BF16, EPI_GELU_BF16, EPI_REDUCE_F32, EPI_RESID_F32, EPI_STORE_BF16, EPI_STORE_F32,
                   call, ptr, stream)

BF16 = torch.bfloat16
F32 = torch.float32


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a 2-D tensor with unit inner stride"
    return t.stride(0)


def gemm(a, b, *, a_mn=False, b_mn=False, epi=EPI_STORE_BF16, out=None, out2=None, bias=None, aux=None,
         splits=0, block_n=0, M=None, N=None, K=None):
    """D[M,N] = sum_k A(m,k) B(n,k) on the tcgen05 path.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True) bf16;  b: [N,K] (b_mn=False) or [K,N] (b_mn=True) bf16.
    """
    assert a.dtype == BF16 and b.dtype == BF16
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    kb = b.shape[0] if b_mn else b.shape[1]
    assert kb == K, f"reduction dims differ: {K} vs {kb}"
    if out is None:
        if epi in (EPI_RESID_F32, EPI_STORE_F32):
            out = torch.empty((M, N), device=a.device, dtype=F32)
        elif epi == EPI_REDUCE_F32:
            out = torch.zeros((M, N), device=a.device, dtype=F32)
        else:
            out = torch.empty((M, N), device=a.device, dtype=BF16)
    call("b200_gemm_bf16", ptr(a), _ld(a), int(a_mn), ptr(b), _ld(b), int(b_mn), M, N, K, epi,
         ptr(out), _ld(out), ptr(out2), _ld(out2) if out2 is not None else 0, ptr(bias),
         ptr(aux), _ld(aux) if aux is not None else 0, splits, block_n, stream())
    return out
