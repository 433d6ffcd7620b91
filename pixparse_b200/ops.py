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
         splits=0, block_n=0, M=None, N=None, K=None, drop=None):
    """D[M,N] = sum_k A(m,k) B(n,k) on the tcgen05 path.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True) bf16;  b: [N,K] (b_mn=False) or [K,N] (b_mn=True) bf16.
    """
    assert a.dtype == BF16 and b.dtype == BF16
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
        kb = b.shape[0] if b_mn else b.shape[1]
        assert kb == K, f"reduction dims differ: {K} vs {kb}"
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        if epi in (EPI_RESID_F32, EPI_STORE_F32):
            out = torch.empty((M, N), device=a.device, dtype=F32)
        elif epi == EPI_REDUCE_F32:
            out = torch.zeros((M, N), device=a.device, dtype=F32)
        else:
            out = torch.empty((M, N), device=a.device, dtype=BF16)
    if drop is not None and drop[0] > 0.0:     # drop = (p, seed): dropout fused into the epilogue
        call("b200_gemm_bf16_dropout", ptr(a), _ld(a), int(a_mn), ptr(b), _ld(b), int(b_mn), M, N, K, epi,
             ptr(out), _ld(out), ptr(out2), _ld(out2) if out2 is not None else 0, ptr(bias),
             ptr(aux), _ld(aux) if aux is not None else 0, splits, block_n, float(drop[0]), int(drop[1]), stream())
        return out
    call("b200_gemm_bf16", ptr(a), _ld(a), int(a_mn), ptr(b), _ld(b), int(b_mn), M, N, K, epi,
         ptr(out), _ld(out), ptr(out2), _ld(out2) if out2 is not None else 0, ptr(bias),
         ptr(aux), _ld(aux) if aux is not None else 0, splits, block_n, stream())
    return out


def attention_fwd(q, k, v, *, B, H, Sq, Sk, q_col0=0, k_col0=0, v_col0=0, causal=False, scale=None, out=None,
                  lse=None, drop=None, q_bs=0, kv_bs=0, out_bs=0, want_lse=True):
    """q/k/v: 2-D bf16 views [B*S, row_width]; head h lives at columns [col0 + 64h, col0 + 64h + 64)."""
    dh = 64
    if scale is None:
        scale = dh ** -0.5
    if out is None:
        out = torch.empty((B * Sq, H * dh), device=q.device, dtype=BF16)
    if lse is None and want_lse:
        lse = torch.empty((B, H, Sq), device=q.device, dtype=F32)
    if q_bs or kv_bs or out_bs:      # explicit batch strides (KV cache): token stride = row stride of the 2-D view
        call("b200_attention_fwd_strided", ptr(q), _ld(q), q_bs, q_col0, ptr(k), _ld(k), kv_bs, k_col0, ptr(v), _ld(v),
             kv_bs, v_col0, ptr(out), _ld(out), out_bs, ptr(lse), B, H, Sq, Sk, dh, int(causal), float(scale), stream())
        return out, lse
    if drop is not None and drop[0] > 0.0:
        call("b200_attention_fwd_dropout", ptr(q), _ld(q), q_col0, ptr(k), _ld(k), k_col0, ptr(v), _ld(v), v_col0,
             ptr(out), _ld(out), ptr(lse), B, H, Sq, Sk, dh, int(causal), float(scale), float(drop[0]), int(drop[1]),
             stream())
        return out, lse
    call("b200_attention_fwd", ptr(q), _ld(q), q_col0, ptr(k), _ld(k), k_col0, ptr(v), _ld(v), v_col0,
         ptr(out), _ld(out), ptr(lse), B, H, Sq, Sk, dh, int(causal), float(scale), stream())
    return out, lse


_att_ws = {}


def attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, *, B, H, Sq, Sk, q_col0=0, k_col0=0, v_col0=0, do_col0=0,
                  dq_col0=0, dk_col0=0, dv_col0=0, causal=False, scale=None, drop=None):
    """Gradients of attention_fwd. dq/dk/dv are caller-provided bf16 2-D buffers laid out like q/k/v."""
    dh = 64
    if scale is None:
        scale = dh ** -0.5
    need = _lib.lib().b200_attention_bwd_workspace_bytes(B, H, Sq)
    key = q.device
    ws = _att_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty((need,), device=q.device, dtype=torch.uint8)
        _att_ws[key] = ws
    if drop is not None and drop[0] > 0.0:
        call("b200_attention_bwd_dropout", ptr(q), _ld(q), q_col0, ptr(k), _ld(k), k_col0, ptr(v), _ld(v), v_col0,
             ptr(o), _ld(o), ptr(d_o), _ld(d_o), do_col0, ptr(lse), ptr(dq), _ld(dq), dq_col0, ptr(dk), _ld(dk),
             dk_col0, ptr(dv), _ld(dv), dv_col0, ptr(ws), B, H, Sq, Sk, dh, int(causal), float(scale), float(drop[0]),
             int(drop[1]), stream())
        return dq, dk, dv
    call("b200_attention_bwd", ptr(q), _ld(q), q_col0, ptr(k), _ld(k), k_col0, ptr(v), _ld(v), v_col0,
         ptr(o), _ld(o), ptr(d_o), _ld(d_o), do_col0, ptr(lse), ptr(dq), _ld(dq), dq_col0, ptr(dk), _ld(dk), dk_col0,
         ptr(dv), _ld(dv), dv_col0, ptr(ws), B, H, Sq, Sk, dh, int(causal), float(scale), stream())
    return dq, dk, dv


def layernorm_fwd(x, gamma, beta, eps, *, want_bf16=True, want_f32=False, drop=None):
    rows, dim = x.shape
    assert x.dtype == F32 and x.is_contiguous()
    y16 = torch.empty((rows, dim), device=x.device, dtype=BF16) if want_bf16 else None
    y32 = torch.empty((rows, dim), device=x.device, dtype=F32) if want_f32 else None
    mean = torch.empty((rows,), device=x.device, dtype=F32)
    rstd = torch.empty((rows,), device=x.device, dtype=F32)
    if drop is not None and drop[0] > 0.0:
        call("b200_layernorm_fwd_dropout", ptr(x), ptr(gamma), ptr(beta), ptr(y16), ptr(y32), ptr(mean), ptr(rstd),
             rows, dim, float(eps), float(drop[0]), int(drop[1]), stream())
        return y16, y32, mean, rstd
    call("b200_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(y16), ptr(y32), ptr(mean), ptr(rstd), rows, dim,
         float(eps), stream())
    return y16, y32, mean, rstd


def layernorm_bwd(x, mean, rstd, gamma, dgamma, dbeta, *, dy16=None, dy32=None, dres32=None, dx32=None, dx16=None,
                  want_f32=True, want_bf16=True, in_drop=None, out_drop=None):
    rows, dim = x.shape
    if dx32 is None and want_f32:
        dx32 = torch.empty((rows, dim), device=x.device, dtype=F32)
    if dx16 is None and want_bf16:
        dx16 = torch.empty((rows, dim), device=x.device, dtype=BF16)
    i_on = in_drop is not None and in_drop[0] > 0.0
    o_on = out_drop is not None and out_drop[0] > 0.0
    if i_on or o_on:
        ip, iseed = in_drop if i_on else (0.0, 0)
        op_, oseed = out_drop if o_on else (0.0, 0)
        call("b200_layernorm_bwd_dropout", ptr(dy16), ptr(dy32), ptr(dres32), ptr(x), ptr(mean), ptr(rstd), ptr(gamma),
             ptr(dx32), ptr(dx16), ptr(dgamma), ptr(dbeta), rows, dim, float(ip), int(iseed), float(op_), int(oseed),
             stream())
        return dx32, dx16
    call("b200_layernorm_bwd", ptr(dy16), ptr(dy32), ptr(dres32), ptr(x), ptr(mean), ptr(rstd), ptr(gamma),
         ptr(dx32), ptr(dx16), ptr(dgamma), ptr(dbeta), rows, dim, stream())
    return dx32, dx16


def colsum(dy, out, rows=None, cols=None):
    """out[n] += sum_m dy[m, n] (dy bf16)."""
    rows = dy.shape[0] if rows is None else rows
    cols = dy.shape[1] if cols is None else cols
    call("b200_colsum_bf16", ptr(dy), _ld(dy), rows, cols, ptr(out), stream())


def patch_unfold(image, P, ld=None):
    B, C, H, W = image.shape
    assert image.dtype == F32 and image.is_contiguous()
    n = B * (H // P) * (W // P)
    K = C * P * P
    # ld > K pads each row (TMA needs 16-byte row strides); the GEMM is told K, so the padding is never read
    patches = torch.empty((n, ld or K), device=image.device, dtype=BF16)
    call("b200_patch_unfold", ptr(image), ptr(patches), B, C, H, W, P, patches.stride(0), stream())
    return patches


def tokens_assemble(proj, cls, pos, B, S, D):
    x = torch.empty((B * S, D), device=proj.device, dtype=F32)
    call("b200_tokens_assemble", ptr(proj), ptr(cls), ptr(pos), ptr(x), B, S, D, stream())
    return x


def tokens_assemble_bwd(dx, dcls, dpos, B, S, D):
    dproj = torch.empty((B * (S - 1), D), device=dx.device, dtype=BF16)
    call("b200_tokens_assemble_bwd", ptr(dx), ptr(dproj), ptr(dcls), ptr(dpos), B, S, D, stream())
    return dproj


def embed_fwd(ids, tok_emb, pos_emb, pos_offset=2, scale=1.0):
    B, T = ids.shape
    D = tok_emb.shape[1]
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    x = torch.empty((B * T, D), device=ids.device, dtype=F32)
    call("b200_embed_fwd", ptr(ids), ptr(tok_emb), ptr(pos_emb), ptr(x), B, T, D, pos_offset, float(scale), stream())
    return x


def embed_bwd(ids, dx, d_tok, d_pos, pos_offset=2, scale=1.0, padding_idx=1):
    B, T = ids.shape
    D = d_tok.shape[1]
    call("b200_embed_bwd", ptr(ids), ptr(dx), ptr(d_tok), ptr(d_pos), B, T, D, pos_offset, float(scale),
         int(padding_idx), stream())


def cast_bf16(src, dst=None):
    n = src.numel()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=BF16)
    call("b200_cast_f32_bf16", ptr(src), ptr(dst), n, stream())
    return dst


def cross_entropy(logits, targets, vocab, *, dlogits=None, grad_scale=1.0, ignore_index=-100, row_loss=None,
                  stats=None):
    """logits: [rows, ld] bf16 (ld >= vocab rounded up to 8). Returns stats tensor: [n_valid, mean_loss].
    dlogits (may alias logits) receives d(mean loss * grad_scale)/dlogits."""
    rows = logits.shape[0]
    assert targets.dtype == torch.int64 and targets.is_contiguous() and targets.numel() == rows
    if stats is None:
        stats = torch.empty((2,), device=logits.device, dtype=F32)
    call("b200_ce_prepare", ptr(targets), rows, ignore_index, ptr(stats), stream())
    call("b200_ce_fwd_bwd", ptr(logits), _ld(logits), ptr(targets), ptr(dlogits),
         _ld(dlogits) if dlogits is not None else 0, ptr(row_loss), ptr(stats), rows, vocab, ignore_index,
         float(grad_scale), stream())
    return stats


_norm_ws = {}


def grad_norm(grads, max_norm=0.0, pre_scale=1.0, out=None):
    """Deterministic global L2 norm of a flat fp32 arena. out = [sumsq, norm, clip_coef]."""
    dev = grads.device
    if dev not in _norm_ws:
        _norm_ws[dev] = torch.empty((_lib.lib().b200_grad_norm_workspace_floats(),), device=dev, dtype=F32)
    if out is None:
        out = torch.empty((3,), device=dev, dtype=F32)
    call("b200_grad_norm", ptr(grads), grads.numel(), ptr(_norm_ws[dev]), ptr(out), float(max_norm or 0.0),
         float(pre_scale), stream())
    return out


def adamw_step(params, grads, exp_avg, exp_avg_sq, params_bf16, segments, num_segments, *, lr, beta1, beta2, eps,
               step, norm_stats=None, grad_scale=1.0, zero_grad=True):
    call("b200_adamw_step", ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), ptr(params_bf16),
         params.numel(), ptr(segments), num_segments, ptr(norm_stats), float(grad_scale), float(lr), float(beta1),
         float(beta2), float(eps), int(step), int(zero_grad), stream())


def release_workspaces():
    """Drop the cached attention-backward / grad-norm workspaces (they are sized by the largest call seen)."""
    _att_ws.clear()
    _norm_ws.clear()


def preprocess_pages(pages_u8, out_size, mean, std, out=None):
    """uint8 grayscale pages [B, Hin, Win] on the device -> normalised fp32 [B, 1, Hout, Wout] (antialiased bicubic)."""
    assert pages_u8.dtype == torch.uint8 and pages_u8.dim() == 3 and pages_u8.is_contiguous()
    B, Hin, Win = pages_u8.shape
    Hout, Wout = out_size
    if out is None:
        out = torch.empty((B, 1, Hout, Wout), device=pages_u8.device, dtype=F32)
    ws = torch.empty((_lib.lib().b200_preprocess_workspace_bytes(Hout, Wout),), device=pages_u8.device,
                     dtype=torch.uint8)
    call("b200_preprocess_pages", ptr(pages_u8), B, Hin, Win, Hin * Win, ptr(out), Hout, Wout, float(mean), float(std),
         ptr(ws), stream())
    return out
