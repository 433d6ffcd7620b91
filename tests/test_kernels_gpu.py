"""GPU parity tests of the individual sm_100a kernels, called through the C-ABI, against plain PyTorch fp32
references of the same op (tolerances written per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


# ------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (1009, 1000, 200, 256), (515, 776, 1032, 128),
                                      (3, 16, 8, 128), (1, 776, 768, 256), (130, 264, 24, 256)])
def test_gemm_layouts(cuda_lib, a_mn, b_mn, M, N, K, bn):
    from pixparse_b200 import ops
    torch.manual_seed(1)
    pad8 = lambda n: (n + 7) // 8 * 8
    # MN-major operands keep M / N contiguous: the row stride must be a multiple of 8 elements
    A = torch.randn((K, pad8(M)) if a_mn else (M, K), device=DEV).bfloat16()
    B = torch.randn((K, pad8(N)) if b_mn else (N, K), device=DEV).bfloat16()
    Af = A.float()[:, :M].t() if a_mn else A.float()
    Bf = B.float()[:, :N] if b_mn else B.float().t()
    ref = Af @ Bf
    out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, epi=ops.EPI_STORE_F32, block_n=bn, M=M, N=N, K=K)
    assert rel_err(out, ref) < 1e-5      # bf16 products are exact in fp32; only accumulation order differs


# (1009, 776, 320): ragged M / N / K inside single tiles; (129, 512, 64) and (300, 256, 128): CTA pairs whose second CTA is
# (almost) empty; (9000, 1536, 192): 216 pair tiles on 74 pairs -> every CTA loops over several tiles (accumulator-stage
# phases, aux prefetch and bias restaging across tiles); (100, 768, 256): single-CTA 256-wide tiles (M <= 128)
# (600, 520, 2112): K >= 2048 selects the residual epilogue's deep-ring variant (one aux tile per group, five operand stages)
@pytest.mark.parametrize("M,N,K", [(1009, 776, 320), (129, 512, 64), (300, 256, 128), (9000, 1536, 192), (100, 768, 256),
                                   (600, 520, 2112)])
@pytest.mark.parametrize("single_cta", [0, 1])
def test_gemm_epilogues(cuda_lib, M, N, K, single_cta):
    from pixparse_b200 import ops, _lib
    _lib.lib().b200_debug_gemm_single_cta(single_cta)
    try:
        _gemm_epilogues_case(M, N, K)
    finally:
        _lib.lib().b200_debug_gemm_single_cta(0)


def _gemm_epilogues_case(M, N, K):
    from pixparse_b200 import ops
    torch.manual_seed(2)
    A = torch.randn((M, K), device=DEV).bfloat16()
    B = torch.randn((N, K), device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    acc = A.float() @ B.float().t()
    out = ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias)
    assert rel_err(out, acc + bias) < 4e-3            # one bf16 rounding of the output
    h = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    g = ops.gemm(A, B, epi=ops.EPI_GELU_BF16, bias=bias, out2=h)
    href = (acc + bias).bfloat16()
    assert rel_err(h, href) < 1e-3
    assert rel_err(g, F.gelu(href.float())) < 4e-3
    x = torch.randn((M, N), device=DEV)
    xref = x + acc + bias
    ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x, out=x)
    assert rel_err(x, xref) < 1e-5
    hh = torch.randn((M, N), device=DEV).bfloat16()
    hf = hh.float().requires_grad_(True)
    F.gelu(hf).backward(acc)
    out = ops.gemm(A, B, epi=ops.EPI_DGELU_BF16, aux=hh)
    assert rel_err(out, hf.grad) < 4e-3


# Tail split (B200GemmArgs.tail_workspace): tiles that leave the persistent grid's last wave at most half full are cut along
# K and summed through the workspace. On 74 CTA pairs: (6656, 768, 2048) = 78 pair tiles -> 4 tail tiles x 4 pieces;
# (6700, 776, 2304) = ragged M / N, 108 tiles -> 34 tail tiles x 2 pieces; single-CTA tiles change the counts (156 / 212
# tiles on 148 CTAs). Each case runs twice: the kernel must leave workspace and counters zeroed.
@pytest.mark.parametrize("M,N,K", [(6656, 768, 2048), (6700, 776, 2304)])
@pytest.mark.parametrize("single_cta", [0, 1])
def test_gemm_tail_split(cuda_lib, M, N, K, single_cta):
    from pixparse_b200 import ops, _lib
    _lib.lib().b200_debug_gemm_single_cta(single_cta)
    prev_split = ops.set_tail_split(True)       # opt-in feature
    try:
        torch.manual_seed(4)
        A = torch.randn((M, K), device=DEV).bfloat16()
        B = torch.randn((N, K), device=DEV).bfloat16()
        Bt = B.t().contiguous() if N % 8 == 0 else None      # MN-major B (dgrad layout) needs a 16-byte row pitch
        bias = torch.randn(N, device=DEV)
        acc = A.float() @ B.float().t()
        x0 = torch.randn((M, N), device=DEV)
        for rep in range(2):
            out = ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias)
            assert rel_err(out, acc + bias) < 4e-3, rep
            if Bt is not None:
                out = ops.gemm(A, Bt, b_mn=True, epi=ops.EPI_STORE_BF16)
                assert rel_err(out, acc) < 4e-3, rep
            x = x0.clone()
            ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x, out=x)      # K >= 2048: the deep-ring variant
            assert rel_err(x, x0 + acc + bias) < 1e-5, rep
            y = torch.empty_like(x0)
            ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x0, out=y, drop=(0.25, 1234 + rep))
            kept = (y != x0)
            frac = kept.float().mean().item()
            assert abs(frac - 0.75) < 5e-3, frac
            assert rel_err(y[kept], (x0 + (acc + bias) / 0.75)[kept]) < 1e-5
        # the fix-up's ordering (pieces' reduce-adds -> ticket -> read-back by the last arriver; the aux tile released only
        # after it was consumed) under repetition: a partial sum that is read too early, or lands after the workspace was
        # zeroed, is off by a whole K piece (~20) against a tolerance of 1e-2
        ref = x0 + acc + bias
        y = torch.empty_like(x0)
        worst = torch.zeros((), device=DEV)
        for _ in range(40):
            ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=x0, out=y)
            worst = torch.maximum(worst, (y - ref).abs().max())
        assert worst.item() < 1e-2, worst.item()
        ws, _ = ops._tail_workspace(A.device)
        torch.cuda.synchronize()
        assert int(ws.count_nonzero().item()) == 0, "tail workspace / counters not left zeroed"
        # and the switch: without the split the same call gives the same answer up to fp32 summation order
        ops.set_tail_split(False)
        ref = ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias)
        ops.set_tail_split(True)
        out = ops.gemm(A, B, epi=ops.EPI_STORE_BF16, bias=bias)
        assert (out.float() - ref.float()).abs().max().item() <= 2 ** -6 * ref.float().abs().max().item()
    finally:
        _lib.lib().b200_debug_gemm_single_cta(0)
        ops.set_tail_split(prev_split)


def test_gemm_splitk_reduce_accumulates(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(3)
    Mw, Nw, Kw = 776, 520, 4036
    A = torch.randn((Kw, Mw), device=DEV).bfloat16()
    B = torch.randn((Kw, Nw), device=DEV).bfloat16()
    base = torch.randn((Mw, Nw), device=DEV)
    ref = base + A.float().t() @ B.float()
    for sp in (1, 4, 0):
        o = base.clone()
        ops.gemm(A, B, a_mn=True, b_mn=True, epi=ops.EPI_REDUCE_F32, out=o, splits=sp)
        assert rel_err(o, ref) < 1e-5


def test_gemm_rejects_bad_alignment(cuda_lib):
    from pixparse_b200 import ops, _lib
    A = torch.randn((128, 64), device=DEV).bfloat16()
    B = torch.randn((130, 64), device=DEV).bfloat16()
    out = torch.empty((128, 130), device=DEV, dtype=torch.bfloat16)   # ldo = 130: not a multiple of 4
    with pytest.raises(_lib.B200Error):
        ops.gemm(A, B, out=out)


# ------------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, causal, scale):
    # q: [B,H,Sq,64] float
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        Sq, Sk = s.shape[-2:]
        mask = torch.ones(Sq, Sk, device=s.device, dtype=torch.bool).tril(Sk - Sq)
        s = s.masked_fill(~mask, float("-inf"))
    p = s.softmax(-1)
    return p @ v, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,H,Sq,Sk,causal", [
    (2, 2, 128, 128, False), (1, 3, 1009, 1009, False), (2, 2, 512, 512, True), (2, 2, 200, 1009, False),
    (1, 2, 4, 300, False), (1, 2, 4, 4, True), (1, 1, 77, 77, True)])
def test_attention_fwd(cuda_lib, B, H, Sq, Sk, causal):
    from pixparse_b200 import ops
    torch.manual_seed(4)
    D = H * 64
    # packed projections like the model produces them: q from one tensor, k|v from another
    qbuf = torch.randn((B * Sq, D), device=DEV).bfloat16()
    kvbuf = torch.randn((B * Sk, 2 * D), device=DEV).bfloat16()
    out, lse = ops.attention_fwd(qbuf, kvbuf, kvbuf, B=B, H=H, Sq=Sq, Sk=Sk, k_col0=0, v_col0=D, causal=causal)
    q = qbuf.float().view(B, Sq, H, 64).transpose(1, 2)
    k = kvbuf[:, :D].float().reshape(B, Sk, H, 64).transpose(1, 2)
    v = kvbuf[:, D:].float().reshape(B, Sk, H, 64).transpose(1, 2)
    oref, lref = _attn_ref(q, k, v, causal, 0.125)
    o = out.float().view(B, Sq, H, 64).transpose(1, 2)
    assert rel_err(o, oref) < 1e-2           # P and O are rounded to bf16
    assert (lse - lref).abs().max().item() < 2e-3


@pytest.mark.parametrize("causal", [False, True])
def test_attention_fwd_growing_scores(cuda_lib, causal):
    """Later key tiles carry scores tens of nats above the first tile's maximum: the single-read softmax must move its
    reference by exact powers of two (rows renormalise several times) and still match the fp32 softmax and logsumexp;
    the backward kernels then consume that logsumexp."""
    from pixparse_b200 import ops
    torch.manual_seed(9)
    B, H, Sq, Sk = 2, 2, 256, 448
    D = H * 64
    qbuf = torch.randn((B * Sq, D), device=DEV).bfloat16()
    kv = torch.randn((B, Sk, 2 * D), device=DEV)
    ramp = torch.linspace(0.05, 40.0, Sk, device=DEV).view(1, Sk, 1)      # key norms grow 800x along the sequence
    kv[:, :, :D] *= ramp
    kvbuf = kv.reshape(B * Sk, 2 * D).bfloat16()
    out, lse = ops.attention_fwd(qbuf, kvbuf, kvbuf, B=B, H=H, Sq=Sq, Sk=Sk, k_col0=0, v_col0=D, causal=causal)
    q = qbuf.float().view(B, Sq, H, 64).transpose(1, 2)
    k = kvbuf[:, :D].float().reshape(B, Sk, H, 64).transpose(1, 2)
    v = kvbuf[:, D:].float().reshape(B, Sk, H, 64).transpose(1, 2)
    oref, lref = _attn_ref(q, k, v, causal, 0.125)
    assert (lref.max() - lref.min()).item() > 50          # the case really spans many renormalisations
    o = out.float().view(B, Sq, H, 64).transpose(1, 2)
    assert torch.isfinite(o).all() and torch.isfinite(lse).all()
    assert rel_err(o, oref) < 1e-2
    assert ((lse - lref).abs() / lref.abs().clamp_min(1.0)).max().item() < 2e-3
    # backward on the same inputs (both kernels)
    dout = torch.randn((B * Sq, D), device=DEV).bfloat16()
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    _attn_ref(qr, kr, vr, causal, 0.125)[0].backward(dout.float().view(B, Sq, H, 64).transpose(1, 2))
    from pixparse_b200 import _lib
    for query_major in (1, 0):
        _lib.lib().b200_debug_attention_bwd_query_major(query_major)
        try:
            dq = torch.empty((B * Sq, D), device=DEV, dtype=torch.bfloat16)
            dkv = torch.empty((B * Sk, 2 * D), device=DEV, dtype=torch.bfloat16)
            ops.attention_bwd(qbuf, kvbuf, kvbuf, out, dout, lse, dq, dkv, dkv, B=B, H=H, Sq=Sq, Sk=Sk, k_col0=0,
                              v_col0=D, dk_col0=0, dv_col0=D, causal=causal)
        finally:
            _lib.lib().b200_debug_attention_bwd_query_major(1)
        g = lambda t, S: t.float().reshape(B, S, H, 64).transpose(1, 2)
        assert rel_err(g(dq, Sq), qr.grad) < 3e-2
        assert rel_err(g(dkv[:, :D], Sk), kr.grad) < 3e-2
        assert rel_err(g(dkv[:, D:], Sk), vr.grad) < 3e-2


@pytest.mark.parametrize("B,H,Sq,Sk,causal", [
    (2, 2, 128, 128, False), (1, 3, 1009, 1009, False), (2, 2, 512, 512, True), (2, 2, 200, 1009, False),
    (1, 2, 4, 300, False), (1, 2, 4, 4, True), (1, 1, 300, 300, True)])
@pytest.mark.parametrize("query_major", [1, 0])
def test_attention_bwd(cuda_lib, B, H, Sq, Sk, causal, query_major):
    """Both no-dropout kernels: query-major (default) and key-major (P^T / dS^T as tensor-memory A operands)."""
    from pixparse_b200 import ops, _lib
    _lib.lib().b200_debug_attention_bwd_query_major(query_major)
    try:
        _attention_bwd_case(B, H, Sq, Sk, causal)
    finally:
        _lib.lib().b200_debug_attention_bwd_query_major(1)


def _attention_bwd_case(B, H, Sq, Sk, causal):
    from pixparse_b200 import ops
    torch.manual_seed(11)
    D = H * 64
    qbuf = torch.randn((B * Sq, D), device=DEV).bfloat16()
    kvbuf = torch.randn((B * Sk, 2 * D), device=DEV).bfloat16()
    dout = torch.randn((B * Sq, D), device=DEV).bfloat16()
    out, lse = ops.attention_fwd(qbuf, kvbuf, kvbuf, B=B, H=H, Sq=Sq, Sk=Sk, k_col0=0, v_col0=D, causal=causal)
    dq = torch.full((B * Sq, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    dkv = torch.full((B * Sk, 2 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention_bwd(qbuf, kvbuf, kvbuf, out, dout, lse, dq, dkv, dkv, B=B, H=H, Sq=Sq, Sk=Sk, k_col0=0, v_col0=D,
                      dk_col0=0, dv_col0=D, causal=causal)
    q = qbuf.float().view(B, Sq, H, 64).transpose(1, 2).requires_grad_(True)
    k = kvbuf[:, :D].float().reshape(B, Sk, H, 64).transpose(1, 2).requires_grad_(True)
    v = kvbuf[:, D:].float().reshape(B, Sk, H, 64).transpose(1, 2).requires_grad_(True)
    oref, _ = _attn_ref(q, k, v, causal, 0.125)
    oref.backward(dout.float().view(B, Sq, H, 64).transpose(1, 2))
    g = lambda t, S: t.float().reshape(B, S, H, 64).transpose(1, 2)
    # P, dS and the outputs are rounded to bf16; D uses the bf16 forward output
    assert rel_err(g(dq, Sq), q.grad) < 2e-2
    assert rel_err(g(dkv[:, :D], Sk), k.grad) < 2e-2
    assert rel_err(g(dkv[:, D:], Sk), v.grad) < 2e-2


# ------------------------------------------------------------------------------------------------- layernorm
@pytest.mark.parametrize("rows,dim,eps", [(1009, 768, 1e-6), (77, 1024, 1e-5), (5, 128, 1e-5)])
def test_layernorm_fwd_bwd(cuda_lib, rows, dim, eps):
    from pixparse_b200 import ops
    torch.manual_seed(5)
    x = (torch.randn((rows, dim), device=DEV) * 2 + 0.5)
    g = torch.randn(dim, device=DEV)
    b = torch.randn(dim, device=DEV)
    y16, y32, mean, rstd = ops.layernorm_fwd(x, g, b, eps, want_f32=True)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yref = F.layer_norm(xr, (dim,), gr, br, eps)
    assert rel_err(y32, yref) < 1e-5
    assert rel_err(y16, yref) < 4e-3
    dy16 = torch.randn((rows, dim), device=DEV).bfloat16()
    dy32 = torch.randn((rows, dim), device=DEV)
    dres = torch.randn((rows, dim), device=DEV)
    yref.backward(dy16.float() + dy32)
    dgamma = torch.zeros(dim, device=DEV)
    dbeta = torch.zeros(dim, device=DEV)
    dx32, dx16 = ops.layernorm_bwd(x, mean, rstd, g, dgamma, dbeta, dy16=dy16, dy32=dy32, dres32=dres)
    assert rel_err(dx32, xr.grad + dres) < 1e-4
    assert rel_err(dx16, xr.grad + dres) < 4e-3
    assert rel_err(dgamma, gr.grad) < 1e-4
    assert rel_err(dbeta, br.grad) < 1e-4


# ------------------------------------------------------------------------------------------------- helpers
def test_colsum(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(6)
    dy = torch.randn((1009, 776), device=DEV).bfloat16()
    out = torch.ones(776, device=DEV)
    ops.colsum(dy, out)
    assert rel_err(out, 1 + dy.float().sum(0)) < 1e-5


def test_patch_unfold_and_assemble(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(7)
    B, C, H, W, P, D = 2, 1, 64, 48, 16, 128
    img = torch.randn((B, C, H, W), device=DEV)
    patches = ops.patch_unfold(img, P)
    ref = F.unfold(img, P, stride=P).transpose(1, 2).reshape(-1, C * P * P)
    assert torch.equal(patches, ref.bfloat16())
    S = (H // P) * (W // P) + 1
    proj = torch.randn((B * (S - 1), D), device=DEV).bfloat16()
    cls = torch.randn(D, device=DEV)
    pos = torch.randn((S, D), device=DEV)
    x = ops.tokens_assemble(proj, cls, pos, B, S, D)
    xref = torch.cat([cls.expand(B, 1, D), proj.float().view(B, S - 1, D)], 1) + pos
    assert torch.allclose(x.view(B, S, D), xref, atol=1e-6)
    dx = torch.randn((B * S, D), device=DEV)
    dcls = torch.zeros(D, device=DEV)
    dpos = torch.zeros((S, D), device=DEV)
    dproj = ops.tokens_assemble_bwd(dx, dcls, dpos, B, S, D)
    d3 = dx.view(B, S, D)
    assert torch.allclose(dpos, d3.sum(0), atol=1e-5)
    assert torch.allclose(dcls, d3[:, 0].sum(0), atol=1e-5)
    assert torch.equal(dproj.view(B, S - 1, D), d3[:, 1:].bfloat16())


def test_embedding_fwd_bwd(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(8)
    B, T, D, V = 3, 17, 128, 1000
    ids = torch.randint(0, V, (B, T), device=DEV)
    ids[0, -3:] = 1   # padding
    tok = torch.randn((V, D), device=DEV)
    pos = torch.randn((T + 2, D), device=DEV)
    x = ops.embed_fwd(ids, tok, pos)
    ref = tok[ids] + pos[torch.arange(T, device=DEV) + 2]
    assert torch.allclose(x.view(B, T, D), ref, atol=1e-6)
    dx = torch.randn((B * T, D), device=DEV)
    dtok = torch.zeros_like(tok)
    dpos = torch.zeros_like(pos)
    ops.embed_bwd(ids, dx, dtok, dpos, padding_idx=1)
    emb = torch.nn.Embedding(V, D, padding_idx=1).to(DEV)
    emb.weight.data.copy_(tok)
    emb(ids).backward(dx.view(B, T, D))
    assert torch.allclose(dtok, emb.weight.grad, atol=1e-5)
    assert torch.allclose(dpos[2:], dx.view(B, T, D).sum(0), atol=1e-5)


@pytest.mark.parametrize("rows,V", [(64, 50267), (33, 1003), (16, 50286)])
def test_cross_entropy(cuda_lib, rows, V):
    from pixparse_b200 import ops
    torch.manual_seed(9)
    ld = (V + 7) // 8 * 8
    logits = torch.zeros((rows, ld), device=DEV, dtype=torch.bfloat16)
    logits[:, :V] = (torch.randn((rows, V), device=DEV) * 3).bfloat16()
    logits[:, V:] = 1000.0      # garbage in the padding must be ignored
    tgt = torch.randint(0, V, (rows,), device=DEV)
    tgt[::5] = -100
    lf = logits[:, :V].float().requires_grad_(True)
    ref = F.cross_entropy(lf, tgt, ignore_index=-100)
    ref.backward()
    dl = torch.empty_like(logits)
    stats = ops.cross_entropy(logits, tgt, V, dlogits=dl)
    assert stats[0].item() == (tgt != -100).sum().item()
    assert abs(stats[1].item() - ref.item()) < 1e-4 * abs(ref.item())     # fp32 statistics over bf16 logits
    assert rel_err(dl[:, :V], lf.grad) < 4e-3                             # bf16 rounding of the gradient


def _segments(rows):
    import numpy as np
    seg = np.zeros(len(rows), dtype=[("end", "<i8"), ("lr_scale", "<f4"), ("wd", "<f4")])
    for i, (end, scale, wd) in enumerate(rows):
        seg[i] = (end, scale, wd)
    return torch.from_numpy(seg.view(np.uint8).copy()).to(DEV)


def test_grad_norm_and_adamw(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(10)
    n = 1 << 20
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 0.01
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    p16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=3e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.0)
    segs = _segments([(n, 1.0, 0.0)])
    stats = torch.zeros(4, device=DEV)       # [sumsq, norm, clip coefficient, applied updates]: kept across steps
    for step in (1, 2, 3):
        g = torch.randn(n, device=DEV) * 0.01
        pr.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        ops.grad_norm(g, max_norm=1.0, out=stats)
        assert abs(stats[1].item() - total.item()) < 1e-5 * total.item()
        assert stats[3].item() == step         # the device-side update counter drives the bias corrections
        ops.adamw_step(p, g, m, v, p16, segs, 1, lr=3e-4, beta1=0.9, beta2=0.98, eps=1e-6, norm_stats=stats)
        assert (p - pr.detach()).abs().max().item() < 2e-6
        assert torch.equal(p16, p.bfloat16())
        assert g.abs().max().item() == 0.0      # zero_grad fused


def test_adamw_segments_lr_scale_and_weight_decay(cuda_lib):
    """Per-tensor {end, lr_scale, weight_decay} table (timm param_groups_layer_decay -> torch.optim.AdamW param groups,
    task_cruller_pretrain.py:196-203): five segments with distinct lr scales and weight decays, boundaries that are not
    multiples of the kernel's grid stride, three steps, no clipping (host-supplied step number)."""
    from pixparse_b200 import ops
    torch.manual_seed(11)
    sizes = [64 * 13, 64 * 1001, 64 * 3, 64 * 257, 64 * 64]
    scales = [0.75 ** 4, 0.75 ** 3, 0.75 ** 2, 0.75, 1.0]
    wds = [0.0, 0.05, 0.0, 0.02, 0.1]
    n = sum(sizes)
    p = torch.randn(n, device=DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ends, refs, groups, o = [], [], [], 0
    for sz, sc, wd in zip(sizes, scales, wds):
        t = p[o:o + sz].clone().requires_grad_(True)
        refs.append(t)
        groups.append({"params": [t], "lr": 1e-3 * sc, "weight_decay": wd})
        o += sz
        ends.append(o)
    opt = torch.optim.AdamW(groups, lr=1e-3, betas=(0.9, 0.99), eps=1e-6)
    segs = _segments(list(zip(ends, scales, wds)))
    for step in (1, 2, 3):
        g = torch.randn(n, device=DEV) * 0.1
        o = 0
        for t, sz in zip(refs, sizes):
            t.grad = g[o:o + sz].clone()
            o += sz
        opt.step()
        ops.adamw_step(p, g, m, v, p16, segs, len(sizes), lr=1e-3, beta1=0.9, beta2=0.99, eps=1e-6, step=step)
        ref = torch.cat([t.detach() for t in refs])
        assert (p - ref).abs().max().item() < 3e-6, step
        # every segment really got its own scale: undoing the update with a neighbour's scale must not match
        assert torch.equal(p16, p.bfloat16())
    o = 0
    for t, sz, sc in zip(refs, sizes, scales):      # per-segment check with a tolerance far below the scale differences
        assert (p[o:o + sz] - t.detach()).abs().max().item() < 3e-6
        o += sz


def test_adamw_skips_non_finite_gradients_and_recovers(cuda_lib):
    """GradScaler semantics (timm NativeScaler, task_cruller_pretrain.py:259-268): a step with inf / nan gradients leaves
    parameters and moments untouched, does NOT count as an update, still clears the gradients, and the next clean step
    behaves exactly like torch.optim.AdamW's first step."""
    from pixparse_b200 import ops
    torch.manual_seed(12)
    n = 64 * 1024
    p = torch.randn(n, device=DEV)
    p0 = p.clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16 = p.bfloat16()
    segs = _segments([(n, 1.0, 0.0)])
    stats = torch.zeros(4, device=DEV)
    for bad in (float("inf"), float("nan")):
        g = torch.randn(n, device=DEV) * 0.01
        g[12345] = bad
        ops.grad_norm(g, max_norm=1.0, out=stats)
        ops.adamw_step(p, g, m, v, p16, segs, 1, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, norm_stats=stats)
        assert torch.equal(p, p0) and m.abs().max().item() == 0.0 and v.abs().max().item() == 0.0
        assert g.abs().max().item() == 0.0          # cleared: nothing non-finite survives into the next accumulation
        assert stats[3].item() == 0.0               # not an applied update
    g = torch.randn(n, device=DEV) * 0.01
    pr = p0.clone().requires_grad_(True)
    pr.grad = g.clone()
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.0)
    torch.nn.utils.clip_grad_norm_([pr], 1.0)
    opt.step()
    ops.grad_norm(g, max_norm=1.0, out=stats)
    ops.adamw_step(p, g, m, v, p16, segs, 1, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, norm_stats=stats)
    assert stats[3].item() == 1.0
    assert not torch.equal(p, p0)
    assert (p - pr.detach()).abs().max().item() < 2e-6      # bias correction of step ONE, not three


def test_wgrad_gemm_fused_bias_gradient(cuda_lib):
    """dW = dY^T X with the nn.Linear bias gradient colsum(dY) summed from the staged A tiles (B200GemmArgs.bias_grad):
    ragged token / feature counts, CTA pairs and single CTAs, split-K, accumulation onto existing values."""
    from pixparse_b200 import ops
    torch.manual_seed(13)
    # (tokens, out features, in features): encoder fc1 / proj like shapes, ragged everything, tiny
    for tokens, n_out, n_in in [(1009, 3072, 768), (2000, 768, 768), (515, 200, 264), (64, 16, 8), (4100, 1536, 768)]:
        pad8 = lambda k: (k + 7) // 8 * 8
        dy = torch.randn((tokens, pad8(n_out)), device=DEV).bfloat16()
        x = torch.randn((tokens, pad8(n_in)), device=DEV).bfloat16()
        dw = torch.zeros((n_out, n_in), device=DEV)
        db = torch.full((n_out,), 0.5, device=DEV)
        ops.gemm(dy, x, a_mn=True, b_mn=True, epi=ops.EPI_REDUCE_F32, out=dw, M=n_out, N=n_in, K=tokens, bias_grad=db)
        ref_w = dy.float()[:, :n_out].t() @ x.float()[:, :n_in]
        ref_b = dy.float()[:, :n_out].sum(0) + 0.5
        assert rel_err(dw, ref_w) < 1e-5, (tokens, n_out, n_in)
        assert (db - ref_b).abs().max().item() < 1e-3 * max(1.0, ref_b.abs().max().item()), (tokens, n_out, n_in)


@pytest.mark.parametrize("Hin,Win,Hout,Wout", [(1100, 850, 576, 448), (330, 250, 576, 448), (64, 48, 64, 48),
                                                 (1754, 1240, 798, 616)])
def test_page_preprocess_matches_torchvision(cuda_lib, Hin, Win, Hout, Wout):
    """ToTensor -> Resize(BICUBIC, antialias=True) -> Normalize as task_cruller_pretrain.py:132-143 builds it."""
    import torchvision.transforms as T
    from PIL import Image
    from pixparse_b200 import ops, synthetic
    pages = synthetic.synthetic_pages_u8(2, Hin, Win, seed=3)
    mean, std = 0.449164, 0.268570
    tf = T.Compose([T.ToTensor(), T.Resize((Hout, Wout), interpolation=T.InterpolationMode.BICUBIC, antialias=True),
                    T.Normalize(mean=mean, std=std)])
    ref = torch.stack([tf(Image.fromarray(p.numpy(), mode="L")) for p in pages])
    out = ops.preprocess_pages(pages.cuda(), (Hout, Wout), mean, std)
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max().item() < 2e-5


# ------------------------------------------------------------------------------------------------- single-token decode
@pytest.mark.parametrize("M,N,K", [(16, 1024, 1024), (5, 4096, 1024), (16, 1000, 4096), (1, 50267, 768), (16, 50286, 1024)])
def test_decode_linear(cuda_lib, M, N, K):
    """y = x W^T (+ bias, GELU, fp32 residual) for <= 16 rows against fp32 matmul; the LM-head mode returns the argmax of
    the bf16-rounded logits (first index on ties) without storing them."""
    from pixparse_b200 import ops
    torch.manual_seed(2)
    x = torch.randn((16, K), device=DEV).bfloat16()
    w = (torch.randn((N, K), device=DEV) * K ** -0.5).bfloat16()
    bias = torch.randn((N,), device=DEV)
    resid = torch.randn((16, N), device=DEV)
    ref = x[:M].float() @ w.float().t()
    out32 = torch.full((16, N), 7.0, device=DEV)
    ops.decode_linear(x, w, M=M, out32=out32, bias=bias, resid=resid)
    assert rel_err(out32[:M], ref + bias + resid[:M]) < 1e-5
    assert (out32[M:] == 7.0).all()
    out16 = torch.zeros((16, N), device=DEV, dtype=torch.bfloat16)
    ops.decode_linear(x, w, M=M, out16=out16, bias=bias, act=1)
    want = F.gelu((ref + bias).bfloat16().float())
    assert rel_err(out16[:M], want) < 4e-3
    # fused argmax (LM head): partial keys -> finalize appends the token at ids[:, pos + 1]
    n_cta = ops.decode_linear_ctas(N)
    partial = torch.zeros((16, n_cta), device=DEV, dtype=torch.int64)
    ops.decode_linear(x, w, M=M, argmax_partial=partial)
    ids = torch.zeros((M, 8), device=DEV, dtype=torch.int64)
    state = torch.tensor([2, -1, 0, 0], device=DEV, dtype=torch.int32)
    fin = torch.zeros(16, device=DEV, dtype=torch.int32)
    logits16 = ref.bfloat16().float()
    want_tok = logits16.argmax(1)
    ops.decode_finalize(partial, n_cta, ids, state, fin, eos_id=int(want_tok[0]))
    got_tok = ids[:, 3]
    # the kernel's fp32 sums differ from torch's in the last bits: accept any column whose bf16 logit equals the row maximum
    picked = logits16.gather(1, got_tok[:, None])[:, 0]
    assert (picked >= logits16.max(1).values - 1e-2 * logits16.abs().max()).all()
    assert (got_tok == want_tok).float().mean().item() >= 0.75
    eos = int(want_tok[0])
    assert (fin[:M] == (got_tok == eos).int()).all() and (fin[M:] == 0).all()
    st = state.tolist()
    assert st[0] == 3 and st[2] == 1 and st[1] == (2 if bool((got_tok == eos).all()) else -1)


@pytest.mark.parametrize("M,N,K", [(16, 1024, 1024), (3, 768, 3072)])
def test_decode_linear_split_output(cuda_lib, M, N, K):
    """The packed q | k | v form: columns below n_split dense and unshifted, the rest appended to a [B, T_max, 2D] cache at
    the device-side position."""
    from pixparse_b200 import ops
    torch.manual_seed(8)
    x = torch.randn((16, K), device=DEV).bfloat16()
    w = (torch.randn((N, K), device=DEV) * K ** -0.5).bfloat16()
    bias = torch.randn((N,), device=DEV)
    D = N // 4
    t_max = 7
    q = torch.zeros((16, D), device=DEV, dtype=torch.bfloat16)
    cache = torch.zeros((16, t_max, N - D), device=DEV, dtype=torch.bfloat16)
    pos = torch.tensor([5], device=DEV, dtype=torch.int32)
    ops.decode_linear(x, w, M=M, out16=q, bias=bias, pos=pos, out_pos_stride=N - D, split=(D, cache, t_max * (N - D)))
    full = (x[:M].float() @ w.float().t() + bias)
    assert rel_err(q[:M], full[:, :D]) < 4e-3 and (q[M:] == 0).all()
    assert rel_err(cache[:M, 5], full[:, D:]) < 4e-3
    cache[:M, 5] = 0
    assert (cache == 0).all()


@pytest.mark.parametrize("B,H,Sk,split_cache", [(16, 16, 2509, False), (3, 12, 77, False), (16, 16, 37, True), (2, 4, 1, True)])
def test_decode_attention(cuda_lib, B, H, Sk, split_cache):
    """One query per (page, head) against strided K | V: cross-attention form (fixed key count, cluster key splits) and
    self-attention form (key count = device position + 1, pad keys hidden through the generated ids)."""
    from pixparse_b200 import ops
    torch.manual_seed(4)
    D = H * 64
    t_max = 64 if split_cache else Sk
    q = torch.randn((B, D), device=DEV).bfloat16()
    kv = torch.randn((B, t_max, 2 * D), device=DEV).bfloat16()
    out = torch.zeros((B, D), device=DEV, dtype=torch.bfloat16)
    qf = q.float().view(B, H, 1, 64)
    kf = kv[:, :Sk, :D].float().reshape(B, Sk, H, 64).transpose(1, 2)
    vf = kv[:, :Sk, D:].float().reshape(B, Sk, H, 64).transpose(1, 2)
    if split_cache:
        pos = torch.tensor([Sk - 1], device=DEV, dtype=torch.int32)
        ids = torch.randint(3, 100, (B, t_max + 1), device=DEV)
        if Sk > 4:
            ids[:, 2] = 1
            ids[0, 4] = 1
        ops.decode_attention(q, kv, kv, out, B=B, H=H, ld_kv=2 * D, kv_bstride=t_max * 2 * D, v_col0=D, pos=pos,
                             key_ids=ids, pad_id=1)
        mask = ids[:, :Sk].ne(1)[:, None, None, :]
    else:
        ops.decode_attention(q, kv.view(B * t_max, 2 * D), kv.view(B * t_max, 2 * D), out, B=B, H=H, ld_kv=2 * D,
                             kv_bstride=t_max * 2 * D, v_col0=D, sk=Sk)
        mask = None
    ref = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask).reshape(B, D)
    assert rel_err(out, ref) < 8e-3      # bf16 probabilities and output


def test_decode_embed(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(6)
    B, D, V = 5, 256, 300
    tok, pe = torch.randn((V, D), device=DEV), torch.randn((40, D), device=DEV)
    ids = torch.randint(0, V, (B, 20), device=DEV)
    pos = torch.tensor([7], device=DEV, dtype=torch.int32)
    x = torch.zeros((B, D), device=DEV)
    ops.decode_embed(ids, pos, tok, pe, x, pos_offset=2, scale=1.5)
    assert torch.allclose(x, tok[ids[:, 7]] * 1.5 + pe[9], atol=1e-6)
