"""Dropout fused into the sm_100a kernels: masks are stateless (seed, element index), so forward and backward must
regenerate the same mask; kept values are scaled by 1/(1-p). Each kernel is checked against PyTorch fp32 math that is
fed the mask extracted from the kernel's own forward output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
P, SEED = 0.1, 1234
SCALE = 65536.0 / (65536.0 - round(P * 65536))


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _resid_mask(rows, cols, seed=SEED):
    """Mask of the RESID epilogue for a [rows, cols] output: run it with acc = 1 (A = ones/K, B = ones), aux = 0."""
    from pixparse_b200 import ops
    K = 64
    A = torch.full((rows, K), 1.0 / K, device=DEV).bfloat16()
    B = torch.ones((cols, K), device=DEV).bfloat16()
    aux = torch.zeros((rows, cols), device=DEV)
    out = ops.gemm(A, B, epi=ops.EPI_RESID_F32, aux=aux, out=torch.empty_like(aux), drop=(P, seed))
    return out != 0


def test_gemm_resid_dropout_statistics_and_determinism(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(0)
    M, N, K = 1009, 768, 256
    A = torch.randn((M, K), device=DEV).bfloat16()
    B = torch.randn((N, K), device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    aux = torch.randn((M, N), device=DEV)
    acc = A.float() @ B.float().t() + bias
    o1 = ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=aux, out=torch.empty_like(aux), drop=(P, SEED))
    o2 = ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=aux, out=torch.empty_like(aux), drop=(P, SEED))
    o3 = ops.gemm(A, B, epi=ops.EPI_RESID_F32, bias=bias, aux=aux, out=torch.empty_like(aux), drop=(P, SEED + 1))
    assert torch.equal(o1, o2)                       # same seed -> same mask
    mask = _resid_mask(M, N)
    assert abs(mask.float().mean().item() - (1 - P)) < 5e-3
    assert rel_err(o1, aux + acc * mask * SCALE) < 1e-5
    assert not torch.equal(o1, o3)
    # masks are uncorrelated between neighbouring elements / rows
    m = mask.float() - (1 - P)
    assert abs((m[:, :-1] * m[:, 1:]).mean().item()) < 2e-3
    assert abs((m[:-1] * m[1:]).mean().item()) < 2e-3


def test_gelu_and_dgelu_share_the_activation_dropout_mask(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(1)
    M, N, K = 515, 512, 128
    A = torch.randn((M, K), device=DEV).bfloat16()
    B = torch.randn((N, K), device=DEV).bfloat16()
    h = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    g = ops.gemm(A, B, epi=ops.EPI_GELU_BF16, out2=h, drop=(P, SEED))
    href = (A.float() @ B.float().t()).bfloat16()
    assert rel_err(h, href) < 1e-3                   # the saved pre-activation is NOT dropped
    mask = _resid_mask(M, N)                         # same (seed, index) -> same mask in every epilogue
    assert rel_err(g, F.gelu(href.float()) * mask * SCALE) < 5e-3
    # backward: d_h = mask * scale * d_g * gelu'(h)
    dg = torch.randn((M, K), device=DEV).bfloat16()
    W = torch.randn((K, N), device=DEV).bfloat16()
    d_h = ops.gemm(dg, W, b_mn=True, epi=ops.EPI_DGELU_BF16, aux=h, drop=(P, SEED))
    hf = h.float().requires_grad_(True)
    (F.gelu(hf) * mask * SCALE).backward(dg.float() @ W.float())
    assert rel_err(d_h, hf.grad) < 5e-3


def test_layernorm_dropout_forward_backward(cuda_lib):
    from pixparse_b200 import ops
    torch.manual_seed(2)
    rows, dim = 301, 768
    x = torch.randn((rows, dim), device=DEV) * 2 + 0.3
    gam = torch.randn(dim, device=DEV)
    bet = torch.randn(dim, device=DEV)
    y16, y32, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-5, want_f32=True, drop=(P, SEED))
    mask = _resid_mask(rows, dim)                    # LayerNorm uses the same index space [rows, dim]
    xr, gr, br = x.clone().requires_grad_(True), gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    yref = F.layer_norm(xr, (dim,), gr, br, 1e-5) * mask * SCALE
    assert rel_err(y32, yref) < 1e-5
    dy16 = torch.randn((rows, dim), device=DEV).bfloat16()
    dy32 = torch.randn((rows, dim), device=DEV)
    yref.backward(dy16.float() + dy32)
    dgamma, dbeta = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    dx32, _ = ops.layernorm_bwd(x, mean, rstd, gam, dgamma, dbeta, dy16=dy16, dy32=dy32, in_drop=(P, SEED))
    assert rel_err(dx32, xr.grad) < 1e-4
    assert rel_err(dgamma, gr.grad) < 1e-4 and rel_err(dbeta, br.grad) < 1e-4
    # out_drop: only the bf16 copy that feeds the dropped sub-layer is masked
    mask2 = _resid_mask(rows, dim, seed=SEED + 5)
    dg2, db2 = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    dx32b, dx16b = ops.layernorm_bwd(x, mean, rstd, gam, dg2, db2, dy32=dy32, out_drop=(P, SEED + 5))
    ref = F.layer_norm(xr, (dim,), gr, br, 1e-5)
    gx, = torch.autograd.grad(ref, xr, dy32)
    assert rel_err(dx32b, gx) < 1e-4
    assert rel_err(dx16b, gx * mask2 * SCALE) < 5e-3


@pytest.mark.parametrize("causal,Sq,Sk", [(False, 100, 64), (True, 64, 64), (False, 130, 200)])
def test_attention_dropout_forward_backward(cuda_lib, causal, Sq, Sk):
    from pixparse_b200 import ops
    torch.manual_seed(3)
    B, H = 2, 2
    D = H * 64
    q = torch.randn((B * Sq, D), device=DEV).bfloat16()
    kv = torch.randn((B * Sk, 2 * D), device=DEV).bfloat16()
    # extract the probability mask with an identity-like V in blocks of 64 keys: O[:, j] = sum_k P'[k] [k % 64 == j]
    masks = torch.zeros((B, H, Sq, Sk), device=DEV, dtype=torch.bool)
    for blk in range((Sk + 63) // 64):
        kv_probe = kv.clone()
        vprobe = torch.zeros((B, Sk, H, 64), device=DEV)
        idx = torch.arange(blk * 64, min(Sk, blk * 64 + 64), device=DEV)
        vprobe[:, idx, :, idx - blk * 64] = 1.0
        kv_probe[:, D:] = vprobe.reshape(B * Sk, D).bfloat16()
        o, _ = ops.attention_fwd(q, kv_probe, kv_probe, B=B, H=H, Sq=Sq, Sk=Sk, v_col0=D, causal=causal, drop=(P, SEED))
        o = o.float().view(B, Sq, H, 64).transpose(1, 2)
        masks[:, :, :, idx] = o[:, :, :, : idx.numel()] != 0
    qf = q.float().view(B, Sq, H, 64).transpose(1, 2).requires_grad_(True)
    kf = kv[:, :D].float().reshape(B, Sk, H, 64).transpose(1, 2).requires_grad_(True)
    vf = kv[:, D:].float().reshape(B, Sk, H, 64).transpose(1, 2).requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if causal:
        s = s.masked_fill(~torch.ones(Sq, Sk, device=DEV, dtype=torch.bool).tril(Sk - Sq), float("-inf"))
    visible = torch.isfinite(s)
    keep_rate = masks[visible.expand_as(masks)].float().mean().item()
    assert abs(keep_rate - (1 - P)) < 2e-2
    oref = (s.softmax(-1) * masks * SCALE) @ vf
    out, lse = ops.attention_fwd(q, kv, kv, B=B, H=H, Sq=Sq, Sk=Sk, v_col0=D, causal=causal, drop=(P, SEED))
    assert rel_err(out.float().view(B, Sq, H, 64).transpose(1, 2), oref) < 1.5e-2
    dout = torch.randn((B * Sq, D), device=DEV).bfloat16()
    oref.backward(dout.float().view(B, Sq, H, 64).transpose(1, 2))
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ops.attention_bwd(q, kv, kv, out, dout, lse, dq, dkv, dkv, B=B, H=H, Sq=Sq, Sk=Sk, v_col0=D, dk_col0=0, dv_col0=D,
                      causal=causal, drop=(P, SEED))
    g = lambda t, S: t.float().reshape(B, S, H, 64).transpose(1, 2)
    assert rel_err(g(dq, Sq), qf.grad) < 2.5e-2
    assert rel_err(g(dkv[:, :D], Sk), kf.grad) < 2.5e-2
    assert rel_err(g(dkv[:, D:], Sk), vf.grad) < 2.5e-2


def test_model_trains_with_dropout_live(cuda_lib):
    """Reference semantics: train_step never calls .eval(), so BART dropout (p = 0.1) is live. Check that the fused
    step with dropout is deterministic for a fixed engine seed, differs from the p = 0 step, and still optimises."""
    from oracle import cruller_ref
    from pixparse_b200 import models, synthetic
    from pixparse_b200.engine import engine_for
    from pixparse_b200.optim import FusedAdamW

    def make():
        ref = cruller_ref.build_model("cruller_test", vocab_size=50267, seed=0)
        cfg = models.get_model_config("cruller_test")
        cfg.image_encoder.pretrained = cfg.text_decoder.pretrained = False
        m = models.Cruller(cfg)
        m.text_decoder.trunk.resize_token_embeddings(50267)
        m.load_state_dict(ref.state_dict(), strict=True)
        return m.to(DEV)
    image, text, target = synthetic.synthetic_batch(4, (64, 48), 33, seed=0)
    image, ti, tt = image.to(DEV), text[:, :-1].contiguous().to(DEV), target[:, 1:].contiguous().to(DEV)
    losses = {}
    for tag, p in (("drop_a", 0.1), ("drop_b", 0.1), ("nodrop", 0.0)):
        m = make()
        m.text_decoder.trunk.set_dropout(p)
        eng = engine_for(m)
        eng.zero_grads()
        stats = eng.forward_backward(image, ti, tt)
        losses[tag] = (stats[1].item(), eng.arena.g32.clone())
    # same engine seed -> same masks; fp32 atomics (loss / split-K / LayerNorm reductions) only reorder the sums
    assert abs(losses["drop_a"][0] - losses["drop_b"][0]) < 1e-5 * losses["drop_a"][0]
    assert rel_err(losses["drop_a"][1], losses["drop_b"][1]) < 1e-4
    assert abs(losses["drop_a"][0] - losses["nodrop"][0]) > 1e-4 * losses["nodrop"][0]
    assert abs(losses["drop_a"][0] - losses["nodrop"][0]) < 0.5
    assert rel_err(losses["drop_a"][1], losses["nodrop"][1]) > 1e-2
    m = make()                                   # default config: dropout 0.1 live, as in the reference
    eng = engine_for(m)
    opt = FusedAdamW(m, eng, lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    eng.zero_grads()
    first = last = None
    for it in range(30):
        stats = eng.forward_backward(image, ti, tt)
        opt.step(clip_grad_norm=1.0)
        if it == 0:
            first = stats[1].item()
        last = stats[1].item()
    assert last < first - 1.0, (first, last)
