"""Thin tensor-level wrappers over the C-ABI (one Python function per entry point).

Every function enqueues on torch's current CUDA stream and returns immediately. Tensors must be CUDA,
contiguous in their last dimension; outputs are allocated by the caller or here through torch's
caching allocator (the kernels never allocate).
"""
import os

import torch

from . import _lib
from ._lib import (EPI_DGELU_BF16, EPI_GELU_BF16, EPI_REDUCE_F32, EPI_RESID_F32, EPI_STORE_BF16, EPI_STORE_F32,
                   call, ptr, stream)

BF16 = torch.bfloat16
F32 = torch.float32


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a 2-D tensor with unit inner stride"
    return t.stride(0)


# Dynamic tile scheduling of the persistent GEMM grids (B200GemmArgs.tile_counter): every launch gets the next of a pool of
# zero-initialised device counters (a counter is back at zero when its kernel ends, and 2048 launches later nothing of
# that kernel is in flight), so kernels on different streams share the SMs work-conservingly. Opt-in
# (PIXPARSE_B200_DYN_SCHED=1, with PIXPARSE_B200_SIDE_WGRAD=1 for the second stream): results are identical either way and,
# measured on the headline step, so is the speed (profiles/r02_experiments_no_gain.txt).
_DYN_SCHED = os.environ.get("PIXPARSE_B200_DYN_SCHED", "0") == "1"
_SCHED_POOL = 2048
_sched_pools = {}


def set_dynamic_tiles(on):
    """Switch the GEMM's dynamic tile scheduler on / off; returns the previous setting."""
    global _DYN_SCHED
    prev, _DYN_SCHED = _DYN_SCHED, bool(on)
    return prev


def _tile_counter(device):
    if not _DYN_SCHED or torch.cuda.is_current_stream_capturing():
        return None
    st = _sched_pools.get(device)
    if st is None:
        st = _sched_pools[device] = [torch.zeros(_SCHED_POOL, device=device, dtype=torch.int32), 0]
    i = st[1]
    st[1] = (i + 1) % _SCHED_POOL
    return st[0].data_ptr() + 4 * i


# Workspace of the GEMM's tail split (B200GemmArgs.tail_workspace): one zero-initialised buffer per device and stream -- the
# kernel leaves it zeroed, and launches that share it must not overlap in time, which stream order guarantees. Opt-in
# (PIXPARSE_B200_GEMM_TAIL_SPLIT=1 / set_tail_split): parity-green but slower on the train step's shapes.
_tail_ws = {}
_TAIL_SPLIT = None


def set_tail_split(on):
    """Switch the GEMM's tail split on / off (library switch + workspace hand-over); returns the previous setting."""
    global _TAIL_SPLIT
    prev, _TAIL_SPLIT = bool(_TAIL_SPLIT), bool(on)
    _lib.lib().b200_debug_gemm_tail_split(int(_TAIL_SPLIT))
    return prev


def _tail_workspace(device):
    if _TAIL_SPLIT is None:
        set_tail_split(os.environ.get("PIXPARSE_B200_GEMM_TAIL_SPLIT", "0") == "1")
    if not _TAIL_SPLIT:
        return None, 0
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _tail_ws.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            return None, 0
        nbytes = _lib.lib().b200_gemm_tail_workspace_bytes()
        ws = _tail_ws[key] = torch.zeros((nbytes + 3) // 4, device=device, dtype=torch.int32)
    return ws, ws.numel() * 4


def gemm(a, b, *, a_mn=False, b_mn=False, epi=EPI_STORE_BF16, out=None, out2=None, bias=None, aux=None,
         splits=0, block_n=0, M=None, N=None, K=None, drop=None, bias_grad=None):
    """D[M,N] = sum_k A(m,k) B(n,k) on the tcgen05 path.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True) bf16;  b: [N,K] (b_mn=False) or [K,N] (b_mn=True) bf16.
    drop = (p, seed): dropout fused into the RESID / GELU / DGELU epilogue.
    bias_grad (weight gradients only: epi=EPI_REDUCE_F32, a_mn=True): fp32 [M] vector that receives += the column sums
    of dY (the A operand), i.e. the nn.Linear bias gradient, without another pass over dY.
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
    if bias_grad is not None:
        assert bias_grad.dtype == F32 and bias_grad.numel() >= M
    p, seed = drop if (drop is not None and drop[0] > 0.0) else (0.0, 0)
    tail_ws, tail_bytes = _tail_workspace(a.device) if epi in (EPI_STORE_BF16, EPI_RESID_F32) else (None, 0)
    args = _lib.GemmArgs(epilogue=epi, a=ptr(a), lda=_ld(a), a_mn_major=int(a_mn), b=ptr(b), ldb=_ld(b),
                         b_mn_major=int(b_mn), m=M, n=N, k=K, out=ptr(out), ldo=_ld(out), out2=ptr(out2),
                         ldo2=_ld(out2) if out2 is not None else 0, bias=ptr(bias), aux=ptr(aux),
                         ld_aux=_ld(aux) if aux is not None else 0, bias_grad=ptr(bias_grad), splits=splits,
                         block_n=block_n, drop_p=float(p), drop_seed=int(seed), tile_counter=_tile_counter(a.device),
                         tail_workspace=ptr(tail_ws), tail_workspace_bytes=tail_bytes)
    call("b200_gemm_bf16", args, stream())
    return out


def attention_fwd(q, k, v, *, B, H, Sq, Sk, q_col0=0, k_col0=0, v_col0=0, causal=False, scale=None, out=None,
                  lse=None, drop=None, q_bs=0, kv_bs=0, out_bs=0, want_lse=True, key_mask=None):
    """q/k/v: 2-D bf16 views [B*S, row_width]; head h lives at columns [col0 + 64h, col0 + 64h + 64).
    q_bs / kv_bs / out_bs: explicit batch strides in elements (KV cache), 0 = dense [B, S, ld].
    key_mask: optional uint8 / bool [B, >= Sk] (row stride = its stride(0)), 0 = key hidden (decoder attention_mask)."""
    dh = 64
    if scale is None:
        scale = dh ** -0.5
    if out is None:
        out = torch.empty((B * Sq, H * dh), device=q.device, dtype=BF16)
    if lse is None and want_lse:
        lse = torch.empty((B, H, Sq), device=q.device, dtype=F32)
    p, seed = drop if (drop is not None and drop[0] > 0.0) else (0.0, 0)
    km_bs = 0
    if key_mask is not None:
        assert key_mask.dim() == 2 and key_mask.stride(1) == 1 and key_mask.shape[1] >= Sk
        assert key_mask.dtype in (torch.uint8, torch.bool)
        km_bs = key_mask.stride(0)
    args = _lib.AttentionFwdArgs(batch=B, q=ptr(q), ldq=_ld(q), q_bstride=q_bs, q_col0=q_col0, k_col0=k_col0, k=ptr(k),
                                 ldk=_ld(k), k_bstride=kv_bs, v=ptr(v), ldv=_ld(v), v_bstride=kv_bs, v_col0=v_col0,
                                 heads=H, out=ptr(out), ld_out=_ld(out), out_bstride=out_bs, lse=ptr(lse),
                                 key_mask=ptr(key_mask), key_mask_bstride=km_bs, sq=Sq, sk=Sk, head_dim=dh,
                                 causal=int(causal), scale=float(scale), drop_p=float(p), drop_seed=int(seed))
    call("b200_attention_fwd", args, stream())
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
    p, seed = drop if (drop is not None and drop[0] > 0.0) else (0.0, 0)
    args = _lib.AttentionBwdArgs(batch=B, q=ptr(q), ldq=_ld(q), k=ptr(k), ldk=_ld(k), v=ptr(v), ldv=_ld(v),
                                 q_col0=q_col0, k_col0=k_col0, v_col0=v_col0, do_col0=do_col0, o=ptr(o), ld_o=_ld(o),
                                 d_o=ptr(d_o), ld_do=_ld(d_o), lse=ptr(lse), dq=ptr(dq), ld_dq=_ld(dq), dk=ptr(dk),
                                 ld_dk=_ld(dk), dv=ptr(dv), ld_dv=_ld(dv), dq_col0=dq_col0, dk_col0=dk_col0,
                                 dv_col0=dv_col0, heads=H, workspace=ptr(ws), sq=Sq, sk=Sk, head_dim=dh,
                                 causal=int(causal), scale=float(scale), drop_p=float(p), drop_seed=int(seed))
    call("b200_attention_bwd", args, stream())
    return dq, dk, dv


def layernorm_fwd(x, gamma, beta, eps, *, want_bf16=True, want_f32=False, drop=None, out=None):
    """out = (y16, y32, mean, rstd) preallocated (entries may be None): nothing is allocated (CUDA-graph capture)."""
    rows, dim = x.shape
    assert x.dtype == F32 and x.is_contiguous()
    if out is not None:
        y16, y32, mean, rstd = out
    else:
        y16 = torch.empty((rows, dim), device=x.device, dtype=BF16) if want_bf16 else None
        y32 = torch.empty((rows, dim), device=x.device, dtype=F32) if want_f32 else None
        mean = torch.empty((rows,), device=x.device, dtype=F32)
        rstd = torch.empty((rows,), device=x.device, dtype=F32)
    p, seed = drop if (drop is not None and drop[0] > 0.0) else (0.0, 0)
    args = _lib.LayerNormFwdArgs(rows=rows, x=ptr(x), gamma=ptr(gamma), beta=ptr(beta), y_bf16=ptr(y16), y_f32=ptr(y32),
                                 mean=ptr(mean), rstd=ptr(rstd), dim=dim, eps=float(eps), drop_p=float(p),
                                 drop_seed=int(seed))
    call("b200_layernorm_fwd", args, stream())
    return y16, y32, mean, rstd


def layernorm_bwd(x, mean, rstd, gamma, dgamma, dbeta, *, dy16=None, dy32=None, dres32=None, dx32=None, dx16=None,
                  want_f32=True, want_bf16=True, in_drop=None, out_drop=None):
    rows, dim = x.shape
    if dx32 is None and want_f32:
        dx32 = torch.empty((rows, dim), device=x.device, dtype=F32)
    if dx16 is None and want_bf16:
        dx16 = torch.empty((rows, dim), device=x.device, dtype=BF16)
    ip, iseed = in_drop if (in_drop is not None and in_drop[0] > 0.0) else (0.0, 0)
    op_, oseed = out_drop if (out_drop is not None and out_drop[0] > 0.0) else (0.0, 0)
    args = _lib.LayerNormBwdArgs(rows=rows, dy_bf16=ptr(dy16), dy_f32=ptr(dy32), dres_f32=ptr(dres32), x=ptr(x),
                                 mean=ptr(mean), rstd=ptr(rstd), gamma=ptr(gamma), dx_f32=ptr(dx32), dx_bf16=ptr(dx16),
                                 dgamma=ptr(dgamma), dbeta=ptr(dbeta), dim=dim, in_p=float(ip), in_seed=int(iseed),
                                 out_p=float(op_), out_seed=int(oseed))
    call("b200_layernorm_bwd", args, stream())
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


def reduce_shards(dst, stage, stride, nsrc, skip, scale):
    """dst (fp32, flat) = (dst + sum over the nsrc slots of `stage` except `skip`, `stride` elements apart) * scale."""
    assert dst.dtype == F32 and stage.dtype == F32 and dst.is_contiguous()
    call("b200_reduce_shards", ptr(dst), ptr(stage), int(stride), int(nsrc), int(skip), float(scale), dst.numel(), stream())
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
    """Deterministic global L2 norm of a flat fp32 arena. out = [sumsq, norm, clip_coef, updates]: out[3] is a counter
    the caller keeps across steps, incremented when the norm is finite (the number of updates really applied)."""
    dev = grads.device
    if dev not in _norm_ws:
        _norm_ws[dev] = torch.empty((_lib.lib().b200_grad_norm_workspace_floats(),), device=dev, dtype=F32)
    if out is None:
        out = torch.zeros((4,), device=dev, dtype=F32)
    assert out.numel() >= 4
    call("b200_grad_norm", ptr(grads), grads.numel(), ptr(_norm_ws[dev]), ptr(out), float(max_norm or 0.0),
         float(pre_scale), stream())
    return out


def adamw_step(params, grads, exp_avg, exp_avg_sq, params_bf16, segments, num_segments, *, lr, beta1, beta2, eps,
               step=0, norm_stats=None, grad_scale=1.0, zero_grad=True):
    """Fused clip + AdamW + bf16 shadow + grad zeroing. With norm_stats (grad_norm's output) the clip coefficient, the
    step number (norm_stats[3]) and the skip-on-non-finite decision all come from the device; otherwise `step` (>= 1)."""
    args = _lib.AdamWArgs(num_segments=num_segments, params=ptr(params), grads=ptr(grads), exp_avg=ptr(exp_avg),
                          exp_avg_sq=ptr(exp_avg_sq), params_bf16=ptr(params_bf16), n=params.numel(),
                          segments=ptr(segments), norm_stats=ptr(norm_stats), grad_scale=float(grad_scale), lr=float(lr),
                          beta1=float(beta1), beta2=float(beta2), eps=float(eps), step=int(step),
                          zero_grad=int(zero_grad))
    call("b200_adamw_step", args, stream())


# ---- single-token greedy decode (csrc/decode.cu) --------------------------------------------------------------------
def decode_linear_ctas(n):
    return _lib.lib().b200_decode_linear_ctas(int(n))


def decode_linear(x, w, *, M, out16=None, out32=None, bias=None, resid=None, act=0, pos=None, out_pos_stride=0,
                  argmax_partial=None, ldo=None, split=None):
    """y[M, N] = x[M, K] w[N, K]^T (+ bias) (act=1: GELU) (+ resid fp32), M <= 16. Outputs bf16 (out16) and / or fp32 (out32)
    with row pitch ldo (default: their stride(0)), shifted by pos[0] * out_pos_stride elements when pos (device int32) is
    given. argmax_partial: nothing is stored, the per-CTA (max, argmax) keys of the bf16-rounded outputs are.
    split = (n_split, out2, ldo2): columns >= n_split go to out2 (bf16) and only they take the position shift."""
    assert x.dtype == BF16 and w.dtype == BF16 and x.dim() == 2 and w.dim() == 2
    N, K = w.shape
    assert x.shape[1] == K and x.shape[0] >= M
    if ldo is None:
        ref = out16 if out16 is not None else out32
        ldo = ref.stride(0) if ref is not None else 0
    n_split, out2, ldo2 = split if split is not None else (0, None, 0)
    args = _lib.DecodeLinearArgs(m=M, x=ptr(x), ldx=_ld(x), w=ptr(w), ldw=_ld(w), bias=ptr(bias), resid=ptr(resid),
                                 ld_resid=_ld(resid) if resid is not None else 0, out_bf16=ptr(out16), out_f32=ptr(out32),
                                 ldo=ldo, pos=ptr(pos), out_pos_stride=out_pos_stride, argmax_partial=ptr(argmax_partial),
                                 n=N, k=K, act=act, n_split=int(n_split), out2_bf16=ptr(out2), ldo2=int(ldo2))
    call("b200_decode_linear", args, stream())


def decode_attention(q, k, v, out, *, B, H, ld_kv, kv_bstride, q_col0=0, k_col0=0, v_col0=0, sk=0, pos=None,
                     key_ids=None, pad_id=0, scale=None):
    """One query per (page, head) against K / V rows `ld_kv` elements apart (pages `kv_bstride` apart). Keys: sk, or
    pos[0] + 1 when pos (device int32) is given; keys whose id in key_ids [B, ld] equals pad_id are hidden."""
    if scale is None:
        scale = 64 ** -0.5
    args = _lib.DecodeAttentionArgs(batch=B, q=ptr(q), ldq=_ld(q), k=ptr(k), v=ptr(v), ld_kv=ld_kv, kv_bstride=kv_bstride,
                                    out=ptr(out), ld_out=_ld(out), pos=ptr(pos), key_ids=ptr(key_ids),
                                    ld_ids=key_ids.stride(0) if key_ids is not None else 0, pad_id=int(pad_id),
                                    q_col0=q_col0, k_col0=k_col0, v_col0=v_col0, heads=H, head_dim=64, sk=int(sk),
                                    scale=float(scale))
    call("b200_decode_attention", args, stream())


def decode_embed(ids, pos, tok_emb, pos_emb, x, pos_offset=2, scale=1.0):
    B, D = x.shape
    assert ids.dtype == torch.int64 and pos.dtype == torch.int32
    call("b200_decode_embed", ptr(ids), ids.stride(0), ptr(pos), ptr(tok_emb), ptr(pos_emb), ptr(x), B, D, pos_offset,
         float(scale), stream())


def decode_finalize(partial, n_cta, ids, state, finished, eos_id):
    """One token per row of ids [B, >= pos + 2]: argmax over the partial keys -> ids[:, pos + 1]; finished[:B] |= == eos."""
    assert state.dtype == torch.int32 and finished.dtype == torch.int32 and ids.dtype == torch.int64
    assert finished.numel() >= ids.shape[0]
    call("b200_decode_finalize", ptr(partial), int(n_cta), ptr(ids), ids.stride(0), ptr(state), ptr(finished),
         ids.shape[0], int(eos_id), stream())


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
