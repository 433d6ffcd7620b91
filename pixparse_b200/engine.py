"""Execution engine of the Cruller train step on sm_100a.

Owns the flat parameter / gradient / bf16-shadow arenas and sequences the hand-written kernels
(:mod:`pixparse_b200.ops`) for

    Cruller.forward                       /root/reference/src/pixparse/models/cruller.py:14-21
      ImageEncoderTimm.forward            models/image_encoder_timm.py:35-42   (timm ViT)
      TextDecoderHf.forward               models/text_decoder_hf.py:80-103     (HF BartForCausalLM, Appendix A.2)
    CrossEntropyLoss + backward           task/task_cruller_pretrain.py:247-278

Precision policy (mirrors torch autocast(bf16) as the reference runs it): fp32 master parameters, fp32 residual
stream and LayerNorm statistics, bf16 GEMM / attention operands with fp32 accumulation, fp32 gradients.
There is no CPU / PyTorch fallback: tensors must live on a B200.
"""
import weakref

import numpy as np
import torch

from . import ops
from .ops import (EPI_DGELU_BF16, EPI_GELU_BF16, EPI_REDUCE_F32, EPI_RESID_F32, EPI_STORE_BF16, EPI_STORE_F32)

ALIGN = 64  # arena alignment in elements (256 B fp32 / 128 B bf16): keeps every tensor TMA-addressable


def _round_up(n, m):
    return (n + m - 1) // m * m


class ParamArena:
    """Flat fp32 parameters + fp32 gradients + bf16 shadow weights; nn.Parameters become views into it."""

    def __init__(self, ordered, device):
        # ordered: list of (key, nn.Parameter) in physical order, no duplicates
        self.device = device
        self.index = {}
        off = 0
        for key, p in ordered:
            n = p.numel()
            self.index[key] = (off, n, tuple(p.shape))
            off += _round_up(n, ALIGN)
        self.total = off
        self.p32 = torch.zeros(self.total, device=device, dtype=torch.float32)
        self.g32 = torch.zeros(self.total, device=device, dtype=torch.float32)
        self.p16 = torch.zeros(self.total, device=device, dtype=torch.bfloat16)
        self.params = [p for _, p in ordered]
        self.keys = [k for k, _ in ordered]
        with torch.no_grad():
            for key, p in ordered:
                o, n, shape = self.index[key]
                view = self.p32[o:o + n].view(shape)
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = self.g32[o:o + n].view(shape)
        self._ptrs = [p.data_ptr() for p in self.params]

    def intact(self):
        """False when a parameter was re-allocated behind our back (module.to(), resize_token_embeddings...)."""
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    def w32(self, key, shape=None):
        o, n, s = self.index[key]
        return self.p32[o:o + n].view(shape or s)

    def w16(self, key, shape=None):
        o, n, s = self.index[key]
        return self.p16[o:o + n].view(shape or s)

    def grad(self, key, shape=None):
        o, n, s = self.index[key]
        return self.g32[o:o + n].view(shape or s)

    def span(self, first_key, last_key, which):
        """Contiguous view covering [first_key .. last_key] (tensors must be adjacent and unpadded)."""
        o0, _, _ = self.index[first_key]
        o1, n1, _ = self.index[last_key]
        buf = {"w32": self.p32, "w16": self.p16, "grad": self.g32}[which]
        return buf[o0:o1 + n1]

    def attach_grads(self):
        for key, p in zip(self.keys, self.params):
            if p.grad is None or p.grad.data_ptr() != self.g32.data_ptr() + 4 * self.index[key][0]:
                o, n, shape = self.index[key]
                p.grad = self.g32[o:o + n].view(shape)

    def sync_shadow(self):
        ops.cast_bf16(self.p32, self.p16)

    def replace_grad_buffer(self, new):
        """Move the gradient arena into `new` (same size; e.g. a peer-addressable symmetric-memory allocation)."""
        assert new.numel() == self.total and new.dtype == torch.float32 and new.device == self.g32.device
        new.copy_(self.g32)
        self.g32 = new
        with torch.no_grad():
            for key, p in zip(self.keys, self.params):
                o, n, shape = self.index[key]
                p.grad = self.g32[o:o + n].view(shape)


def engine_for(module):
    """Engine of the Cruller that owns `module` (created lazily; sub-modules share their parent's engine)."""
    eng = module.__dict__.get('_b200_engine')
    if eng is None:
        root_ref = module.__dict__.get('_b200_root')
        root = root_ref() if root_ref is not None else None
        if root is not None and root is not module:
            return engine_for(root)
        eng = CrullerEngine(module)
        module.__dict__['_b200_engine'] = eng
    return eng


class _Saved:
    pass


class _SideLane:
    """Work with no consumer before the optimizer (weight / bias gradients) on a second stream. Every kernel of the
    step owns whole SMs, so the two streams do not run side by side on an SM; what the side stream gets are the SMs
    the main stream's kernel frees in its partly empty last wave and the gaps between dependent kernels. That only
    pays with the GEMM's dynamic tile scheduler (ops.gemm, B200GemmArgs.tile_counter): with a static deal a CTA that
    starts late still holds a full share of its kernel. ``side`` = None runs everything inline on the current stream."""

    def __init__(self, side, device):
        self.side = side
        self.main = torch.cuda.current_stream(device) if side is not None else None

    def mark(self):
        """Event at the current point of the main stream (None when inline)."""
        if self.side is None:
            return None
        ev = torch.cuda.Event()
        ev.record(self.main)
        return ev

    def run(self, fn, keep, after=None):
        """Enqueue fn() on the side stream behind the main-stream point `after` (default: everything issued so far);
        returns an event recorded after it. `keep`: tensors the side stream reads (they must outlive its kernels)."""
        if self.side is None:
            fn()
            return None
        self.side.wait_event(after if after is not None else self.mark())
        with torch.cuda.stream(self.side):
            fn()
        done = torch.cuda.Event()
        done.record(self.side)
        for t in keep:
            t.record_stream(self.side)
        return done

    def wait(self, ev):
        """The main stream waits for a side-stream event (before it overwrites something that work reads)."""
        if ev is not None:
            self.main.wait_event(ev)

    def join(self):
        if self.side is not None:
            self.main.wait_stream(self.side)


class DecodeCache:
    """KV cache of one greedy-decoding session: per decoder layer the self-attention keys | values of
    every generated position ([B, T_max, 2D], appended in place) and the cross-attention K | V projection of the image
    tokens (computed at the first step). It is what TextDecoderHf returns as ``past_key_values``."""

    def __init__(self, num_layers, B, t_max, D, device):
        self.t_max = t_max
        self.length = 0
        self.self_kv = [torch.empty((B, t_max, 2 * D), device=device, dtype=torch.bfloat16) for _ in range(num_layers)]
        self.cross_kv = [None] * num_layers

    def get_seq_length(self):
        return self.length


class CrullerEngine:
    def __init__(self, module):
        from . import models
        self.module_ref = weakref.ref(module)
        if isinstance(module, models.Cruller):
            self.vit = module.image_encoder.trunk
            self.bart = module.text_decoder.trunk
            module.image_encoder.__dict__['_b200_root'] = weakref.ref(module)
            module.text_decoder.__dict__['_b200_root'] = weakref.ref(module)
        elif isinstance(module, models.ImageEncoderTimm):
            self.vit, self.bart = module.trunk, None
        elif isinstance(module, models.TextDecoderHf):
            self.vit, self.bart = None, module.trunk
        else:
            raise TypeError(f"no B200 engine for {type(module).__name__}")
        if self.bart is not None:
            self.bart.__dict__['_b200_engine_ref'] = weakref.ref(self)
        self.arena = None
        self.saved = None
        self._shadow_fresh = False
        self._param_versions = None
        # dropout: live in training mode exactly where BartDecoder applies it (the ViT has all drop rates 0);
        # masks are regenerated from (seed, site) in backward, the seed advances every forward. The per-engine seed is
        # drawn from torch's default generator (torch.manual_seed controls it; framework/random.py seeds every rank with
        # seed + rank, so DDP ranks draw different masks) the first time a training forward needs it.
        self.dropout_seed = None
        self._dropout_calls = 0
        self._decode_sessions = {}

    # ------------------------------------------------------------------------------------------------ binding
    def invalidate(self):
        self.arena = None
        self._decode_sessions = {}

    def _ordered_params(self):
        """Physical arena order: q|k|v (and cross k|v) weights adjacent so they act as one packed GEMM operand."""
        out = []
        if self.vit is not None:
            v = self.vit
            pre = "vit."
            out += [(pre + "cls_token", v.cls_token), (pre + "pos_embed", v.pos_embed),
                    (pre + "patch.w", v.patch_embed.proj.weight)]
            if v.patch_embed.proj.bias is not None:
                out.append((pre + "patch.b", v.patch_embed.proj.bias))
            if v.arch['pre_norm']:
                out += [(pre + "norm_pre.w", v.norm_pre.weight), (pre + "norm_pre.b", v.norm_pre.bias)]
            for i, blk in enumerate(v.blocks):
                b = f"{pre}{i}."
                out += [(b + "n1.w", blk.norm1.weight), (b + "n1.b", blk.norm1.bias),
                        (b + "qkv.w", blk.attn.qkv.weight), (b + "qkv.b", blk.attn.qkv.bias),
                        (b + "proj.w", blk.attn.proj.weight), (b + "proj.b", blk.attn.proj.bias),
                        (b + "n2.w", blk.norm2.weight), (b + "n2.b", blk.norm2.bias),
                        (b + "fc1.w", blk.mlp.fc1.weight), (b + "fc1.b", blk.mlp.fc1.bias),
                        (b + "fc2.w", blk.mlp.fc2.weight), (b + "fc2.b", blk.mlp.fc2.bias)]
            out += [(pre + "norm.w", v.norm.weight), (pre + "norm.b", v.norm.bias)]
        if self.bart is not None:
            d = self.bart.model.decoder
            pre = "dec."
            out += [(pre + "tok", d.embed_tokens.weight), (pre + "pos", d.embed_positions.weight),
                    (pre + "ln_emb.w", d.layernorm_embedding.weight), (pre + "ln_emb.b", d.layernorm_embedding.bias)]
            for j, L in enumerate(d.layers):
                b = f"{pre}{j}."
                sa, ca = L.self_attn, L.encoder_attn
                out += [(b + "sa.q.w", sa.q_proj.weight), (b + "sa.k.w", sa.k_proj.weight),
                        (b + "sa.v.w", sa.v_proj.weight),
                        (b + "sa.q.b", sa.q_proj.bias), (b + "sa.k.b", sa.k_proj.bias), (b + "sa.v.b", sa.v_proj.bias),
                        (b + "sa.o.w", sa.out_proj.weight), (b + "sa.o.b", sa.out_proj.bias),
                        (b + "sa_ln.w", L.self_attn_layer_norm.weight), (b + "sa_ln.b", L.self_attn_layer_norm.bias),
                        (b + "ca.q.w", ca.q_proj.weight), (b + "ca.q.b", ca.q_proj.bias),
                        (b + "ca.k.w", ca.k_proj.weight), (b + "ca.v.w", ca.v_proj.weight),
                        (b + "ca.k.b", ca.k_proj.bias), (b + "ca.v.b", ca.v_proj.bias),
                        (b + "ca.o.w", ca.out_proj.weight), (b + "ca.o.b", ca.out_proj.bias),
                        (b + "ca_ln.w", L.encoder_attn_layer_norm.weight),
                        (b + "ca_ln.b", L.encoder_attn_layer_norm.bias),
                        (b + "fc1.w", L.fc1.weight), (b + "fc1.b", L.fc1.bias),
                        (b + "fc2.w", L.fc2.weight), (b + "fc2.b", L.fc2.bias),
                        (b + "f_ln.w", L.final_layer_norm.weight), (b + "f_ln.b", L.final_layer_norm.bias)]
        return out

    def ensure_bound(self):
        if self.arena is not None and self.arena.intact():
            return self.arena
        ordered = self._ordered_params()
        dev = ordered[0][1].device
        if dev.type != 'cuda':
            raise RuntimeError("pixparse_b200: parameters are on %s; move the model to a B200 (model.to('cuda')) -- "
                               "there is no CPU fallback for the Cruller hot path" % dev)
        from . import _lib
        _lib.check(_lib.lib().b200_device_check(), "b200_device_check")
        self.arena = ParamArena(ordered, dev)
        self._shadow_fresh = False
        self._param_versions = None
        if self.vit is not None:
            a = self.vit.arch
            K = self.vit.in_chans * a['patch_size'] ** 2
            self._patch_k = K
            self._patch_kpad = _round_up(K, 8)
            self._patch_w16 = (torch.zeros((a['embed_dim'], self._patch_kpad), device=dev, dtype=torch.bfloat16)
                               if self._patch_kpad != K else None)
        return self.arena

    def _versions(self):
        return [p._version for p in self.arena.params]

    def refresh_shadow(self, force=False):
        """bf16 shadow weights follow the fp32 masters (one cast kernel over the arena).

        The cast runs when a Parameter's version counter moved (optimizer.step of a torch optimizer, load_state_dict,
        p.copy_ / p.add_ under no_grad ...) or after the fused optimizer reported an update. Writers that bypass the
        version counter -- ``p.data.copy_()``, EMA swaps through ``.data``, a collective writing ``arena.p32`` -- must call
        ``refresh_shadow(force=True)`` (or ``invalidate_shadow()``) themselves; there is no way to observe them."""
        ar = self.ensure_bound()
        ver = self._versions()
        if force or not self._shadow_fresh or ver != self._param_versions:
            ar.sync_shadow()
            if self.vit is not None and self._patch_w16 is not None:
                D = self.vit.arch['embed_dim']
                self._patch_w16[:, :self._patch_k].copy_(ar.w16("vit.patch.w", (D, self._patch_k)))
            self._shadow_fresh = True
            self._param_versions = ver

    def invalidate_shadow(self):
        """Mark the bf16 shadow stale: the next forward re-casts the fp32 masters (for writers that go through ``.data``)."""
        self._shadow_fresh = False

    def mark_params_updated_by_kernel(self, shadow_written):
        """Called by the fused optimizer: masters changed through raw pointers (no torch version bump)."""
        self._shadow_fresh = bool(shadow_written)
        if shadow_written and self.vit is not None and self._patch_w16 is not None:
            D = self.vit.arch['embed_dim']
            self._patch_w16[:, :self._patch_k].copy_(self.arena.w16("vit.patch.w", (D, self._patch_k)))
        self._param_versions = self._versions()

    # ------------------------------------------------------------------------------------------------ encoder
    def _patch_weight16(self):
        if self._patch_w16 is not None:
            return self._patch_w16
        return self.arena.w16("vit.patch.w", (self.vit.arch['embed_dim'], self._patch_k))

    def encoder_forward(self, image, save):
        ar, v = self.arena, self.vit
        a = v.arch
        D, Hh, P, eps = a['embed_dim'], a['num_heads'], a['patch_size'], a['ln_eps']
        B, C, H, W = image.shape
        assert (H, W) == v.patch_embed.img_size, f"input size {(H, W)} != model img_size {v.patch_embed.img_size}"
        assert C == v.in_chans
        if image.dtype != torch.float32 or not image.is_contiguous():
            image = image.float().contiguous()
        S = v.patch_embed.num_patches + 1
        M = B * S
        st = _Saved()
        st.B, st.S = B, S
        patches = ops.patch_unfold(image, P, ld=self._patch_kpad)
        proj = ops.gemm(patches, self._patch_weight16(), epi=EPI_STORE_BF16,
                        bias=ar.w32("vit.patch.b") if "vit.patch.b" in ar.index else None, K=self._patch_k)
        x = ops.tokens_assemble(proj, ar.w32("vit.cls_token"), ar.w32("vit.pos_embed"), B, S, D)
        st.patches = patches
        if a['pre_norm']:
            _, x2, mean, rstd = ops.layernorm_fwd(x, ar.w32("vit.norm_pre.w"), ar.w32("vit.norm_pre.b"), eps,
                                                  want_bf16=False, want_f32=True)
            st.pre = (x, mean, rstd)
            x = x2
        st.blocks = []
        for i in range(a['depth']):
            k = f"vit.{i}."
            ln1, _, mean1, rstd1 = ops.layernorm_fwd(x, ar.w32(k + "n1.w"), ar.w32(k + "n1.b"), eps)
            qkv = ops.gemm(ln1, ar.w16(k + "qkv.w"), bias=ar.w32(k + "qkv.b"))
            attn, lse = ops.attention_fwd(qkv, qkv, qkv, B=B, H=Hh, Sq=S, Sk=S, q_col0=0, k_col0=D, v_col0=2 * D)
            x1 = torch.empty_like(x)
            ops.gemm(attn, ar.w16(k + "proj.w"), bias=ar.w32(k + "proj.b"), epi=EPI_RESID_F32, aux=x, out=x1)
            ln2, _, mean2, rstd2 = ops.layernorm_fwd(x1, ar.w32(k + "n2.w"), ar.w32(k + "n2.b"), eps)
            hpre = torch.empty((M, ar.index[k + "fc1.w"][2][0]), device=x.device, dtype=torch.bfloat16)
            g = ops.gemm(ln2, ar.w16(k + "fc1.w"), bias=ar.w32(k + "fc1.b"), epi=EPI_GELU_BF16, out2=hpre)
            x2 = torch.empty_like(x)
            ops.gemm(g, ar.w16(k + "fc2.w"), bias=ar.w32(k + "fc2.b"), epi=EPI_RESID_F32, aux=x1, out=x2)
            if save:
                st.blocks.append((x, mean1, rstd1, ln1, qkv, attn, lse, x1, mean2, rstd2, ln2, hpre, g))
            x = x2
        enc16, _, meanf, rstdf = ops.layernorm_fwd(x, ar.w32("vit.norm.w"), ar.w32("vit.norm.b"), eps)
        st.final = (x, meanf, rstdf)
        return enc16, (st if save else None)

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    def encoder_backward(self, st, d_enc32):
        ar, v = self.arena, self.vit
        a = v.arch
        D, Hh = a['embed_dim'], a['num_heads']
        B, S = st.B, st.S
        x, meanf, rstdf = st.final
        dx32, dx16 = ops.layernorm_bwd(x, meanf, rstdf, ar.w32("vit.norm.w"), ar.grad("vit.norm.w"),
                                       ar.grad("vit.norm.b"), dy32=d_enc32)
        # Bias gradients ride on the weight-gradient GEMMs (their A operand is dY: B200GemmArgs.bias_grad), no colsum pass.
        # Weight / bias gradients have no consumer before the optimizer: they go to a side stream (_SideLane) right behind
        # the kernel that produces their dY, the critical-path kernel of the main stream always enqueued first.
        lane = _SideLane(self._side_stream() if self.side_wgrad and dx16.is_cuda else None, dx16.device)
        for i in reversed(range(a['depth'])):
            k = f"vit.{i}."
            (x0, mean1, rstd1, ln1, qkv, attn, lse, x1, mean2, rstd2, ln2, hpre, g) = st.blocks[i]
            # --- MLP
            def w_fc2(dx16=dx16, g=g, k=k):
                ops.gemm(dx16, g, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "fc2.w"),
                         bias_grad=ar.grad(k + "fc2.b"))
            at_start = lane.mark()      # dx16 of this block is complete here
            d_h = ops.gemm(dx16, ar.w16(k + "fc2.w"), b_mn=True, epi=EPI_DGELU_BF16, aux=hpre)
            ev_fc2 = lane.run(w_fc2, (dx16, g), after=at_start)

            def w_fc1(d_h=d_h, ln2=ln2, k=k):
                ops.gemm(d_h, ln2, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "fc1.w"),
                         bias_grad=ar.grad(k + "fc1.b"))
            at_dh = lane.mark()
            d_ln2 = ops.gemm(d_h, ar.w16(k + "fc1.w"), b_mn=True)
            lane.run(w_fc1, (d_h, ln2), after=at_dh)
            lane.wait(ev_fc2)           # the LayerNorm backward below overwrites dx16
            ops.layernorm_bwd(x1, mean2, rstd2, ar.w32(k + "n2.w"), ar.grad(k + "n2.w"), ar.grad(k + "n2.b"),
                              dy16=d_ln2, dres32=dx32, dx32=dx32, dx16=dx16)
            # --- attention
            def w_proj(dx16=dx16, attn=attn, k=k):
                ops.gemm(dx16, attn, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "proj.w"),
                         bias_grad=ar.grad(k + "proj.b"))
            at_ln2 = lane.mark()
            d_attn = ops.gemm(dx16, ar.w16(k + "proj.w"), b_mn=True)
            ev_proj = lane.run(w_proj, (dx16, attn), after=at_ln2)
            dqkv = torch.empty_like(qkv)
            ops.attention_bwd(qkv, qkv, qkv, attn, d_attn, lse, dqkv, dqkv, dqkv, B=B, H=Hh, Sq=S, Sk=S,
                              q_col0=0, k_col0=D, v_col0=2 * D, dq_col0=0, dk_col0=D, dv_col0=2 * D)

            def w_qkv(dqkv=dqkv, ln1=ln1, k=k):
                ops.gemm(dqkv, ln1, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "qkv.w"),
                         bias_grad=ar.grad(k + "qkv.b"))
            at_dqkv = lane.mark()
            d_ln1 = ops.gemm(dqkv, ar.w16(k + "qkv.w"), b_mn=True)
            lane.run(w_qkv, (dqkv, ln1), after=at_dqkv)
            lane.wait(ev_proj)          # dx16 is overwritten again
            ops.layernorm_bwd(x0, mean1, rstd1, ar.w32(k + "n1.w"), ar.grad(k + "n1.w"), ar.grad(k + "n1.b"),
                              dy16=d_ln1, dres32=dx32, dx32=dx32, dx16=dx16)
            st.blocks[i] = None
            if self._grad_ready_hook is not None:
                lane.join()             # the exchange of this block's range must see its weight gradients
                self._grad_ready_hook(k + "n1.w", k + "fc2.b")
        lane.join()
        if a['pre_norm']:
            xp, mean, rstd = st.pre
            ops.layernorm_bwd(xp, mean, rstd, ar.w32("vit.norm_pre.w"), ar.grad("vit.norm_pre.w"),
                              ar.grad("vit.norm_pre.b"), dy32=dx32, dx32=dx32, want_bf16=False)
        dproj = ops.tokens_assemble_bwd(dx32, ar.grad("vit.cls_token"), ar.grad("vit.pos_embed"), B, S, D)
        ops.gemm(dproj, st.patches, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32,
                 out=ar.grad("vit.patch.w", (D, self._patch_k)), N=self._patch_k,
                 bias_grad=ar.grad("vit.patch.b") if "vit.patch.b" in ar.index else None)

    # ------------------------------------------------------------------------------------------------ decoder
    class _Drop:
        """(p, seed) pairs for every dropout site of one decoder forward/backward."""

        def __init__(self, cfg, base_seed, training):
            on = bool(training)
            self.p = float(cfg.dropout) if on else 0.0
            self.pa = float(cfg.attention_dropout) if on else 0.0
            self.pact = float(cfg.activation_dropout) if on else 0.0
            self.base = int(base_seed) & 0xFFFFFFFF

        def _seed(self, site):
            # murmur3 finaliser over (forward counter, site): the kernels combine the seed LINEARLY with the element
            # counter (x = pair * 0x9E3779B1 + seed, common.cuh dropout_hash), so seeds that are affine in the step or
            # the site would make one mask a shifted copy of another; a full-avalanche mix removes that structure
            x = (self.base ^ (site * 0x9E3779B1)) & 0xFFFFFFFF
            x ^= x >> 16
            x = (x * 0x85EBCA6B) & 0xFFFFFFFF
            x ^= x >> 13
            x = (x * 0xC2B2AE35) & 0xFFFFFFFF
            x ^= x >> 16
            x = (x + 0x27D4EB2F * (site + 1)) & 0xFFFFFFFF
            x ^= x >> 15
            x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
            x ^= x >> 12
            return x

        def emb(self):
            return (self.p, self._seed(0))

        def site(self, layer, k):
            # k: 0 self-attn probs, 1 self-attn out, 2 cross-attn probs, 3 cross-attn out, 4 activation, 5 ffn out
            p = (self.pa, self.p, self.pa, self.p, self.pact, self.p)[k]
            return (p, self._seed(1 + 8 * layer + k))

    def decoder_forward(self, ids, enc16, B, S, save, cache=None, key_mask=None):
        """cache: DecodeCache for incremental decoding (inference only): `ids` are the NEW tokens, positions continue
        at cache.length, self-attention keys / values are appended to the cache, the cross-attention K / V projections
        of the image tokens are computed once and reused.
        key_mask: uint8 [B, past + T], 0 = padding key hidden from the decoder self-attention (the reference's
        attention_mask = input_ids.ne(pad), text_decoder_hf.py:68); inference only."""
        ar, bart = self.arena, self.bart
        cfg = bart.config
        D, Hh, nl = cfg.d_model, cfg.decoder_attention_heads, cfg.decoder_layers
        eps = 1e-5
        T = ids.shape[1]
        assert ids.shape[0] == B
        past = cache.length if cache is not None else 0
        assert past + T <= cfg.max_position_embeddings, "sequence longer than max_position_embeddings"
        assert not (save and cache is not None), "the KV cache is an inference-only path"
        assert not (save and key_mask is not None), "attention_mask is honoured on the inference path only"
        M = B * T
        V = ar.index["dec.tok"][2][0]
        if ids.dtype != torch.int64 or not ids.is_contiguous():
            ids = ids.long().contiguous()
        st = _Saved()
        st.B, st.T, st.S, st.V, st.ids = B, T, S, V, ids
        self._dropout_calls += 1
        training = bool(save and bart.training)
        if training and self.dropout_seed is None:
            rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
            self.dropout_seed = (int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) + 0x632BE5AB * rank) & 0xFFFFFFFF
        dr = CrullerEngine._Drop(cfg, ((self.dropout_seed or 0) * 0x9E3779B1 + self._dropout_calls * 0x85EBCA77) & 0xFFFFFFFF,
                                 training)
        st.drop = dr
        x_emb = ops.embed_fwd(ids, ar.w32("dec.tok"), ar.w32("dec.pos"), pos_offset=2 + past, scale=1.0)
        h16, h32, me, re_ = ops.layernorm_fwd(x_emb, ar.w32("dec.ln_emb.w"), ar.w32("dec.ln_emb.b"), eps, want_f32=True,
                                              drop=dr.emb())
        st.emb = (x_emb, me, re_)
        st.layers = []
        for j in range(nl):
            k = f"dec.{j}."
            # causal self-attention (q|k|v packed)
            wqkv = ar.span(k + "sa.q.w", k + "sa.v.w", "w16").view(3 * D, D)
            bqkv = ar.span(k + "sa.q.b", k + "sa.v.b", "w32")
            qkv = ops.gemm(h16, wqkv, bias=bqkv)
            if cache is not None:
                # append this step's keys / values, attend over the whole cached prefix (causal inside the new block)
                kvbuf = cache.self_kv[j]                                      # [B, T_max, 2D]
                kvbuf[:, past:past + T, :].copy_(qkv.view(B, T, 3 * D)[:, :, D:])
                kv2d = kvbuf.view(B * cache.t_max, 2 * D)
                a_s, lse_s = ops.attention_fwd(qkv, kv2d, kv2d, B=B, H=Hh, Sq=T, Sk=past + T, q_col0=0, k_col0=0,
                                               v_col0=D, causal=True, q_bs=T * 3 * D, kv_bs=cache.t_max * 2 * D,
                                               out_bs=T * D, want_lse=False, key_mask=key_mask)
            else:
                a_s, lse_s = ops.attention_fwd(qkv, qkv, qkv, B=B, H=Hh, Sq=T, Sk=T, q_col0=0, k_col0=D,
                                               v_col0=2 * D, causal=True, drop=dr.site(j, 0), key_mask=key_mask)
            u1 = torch.empty_like(h32)
            ops.gemm(a_s, ar.w16(k + "sa.o.w"), bias=ar.w32(k + "sa.o.b"), epi=EPI_RESID_F32, aux=h32, out=u1,
                     drop=dr.site(j, 1))
            h1_16, h1_32, m1, r1 = ops.layernorm_fwd(u1, ar.w32(k + "sa_ln.w"), ar.w32(k + "sa_ln.b"), eps,
                                                     want_f32=True)
            # cross-attention over the image tokens (k|v packed)
            qc = ops.gemm(h1_16, ar.w16(k + "ca.q.w"), bias=ar.w32(k + "ca.q.b"))
            wkv = ar.span(k + "ca.k.w", k + "ca.v.w", "w16").view(2 * D, D)
            bkv = ar.span(k + "ca.k.b", k + "ca.v.b", "w32")
            if cache is not None and cache.cross_kv[j] is not None:
                kvc = cache.cross_kv[j]
            else:
                kvc = ops.gemm(enc16, wkv, bias=bkv)
                if cache is not None:
                    cache.cross_kv[j] = kvc
            a_c, lse_c = ops.attention_fwd(qc, kvc, kvc, B=B, H=Hh, Sq=T, Sk=S, q_col0=0, k_col0=0, v_col0=D,
                                           drop=dr.site(j, 2))
            u2 = torch.empty_like(h32)
            ops.gemm(a_c, ar.w16(k + "ca.o.w"), bias=ar.w32(k + "ca.o.b"), epi=EPI_RESID_F32, aux=h1_32, out=u2,
                     drop=dr.site(j, 3))
            h2_16, h2_32, m2, r2 = ops.layernorm_fwd(u2, ar.w32(k + "ca_ln.w"), ar.w32(k + "ca_ln.b"), eps,
                                                     want_f32=True)
            # feed-forward
            F_ = ar.index[k + "fc1.w"][2][0]
            hpre = torch.empty((M, F_), device=h16.device, dtype=torch.bfloat16)
            g = ops.gemm(h2_16, ar.w16(k + "fc1.w"), bias=ar.w32(k + "fc1.b"), epi=EPI_GELU_BF16, out2=hpre,
                         drop=dr.site(j, 4))
            u3 = torch.empty_like(h32)
            ops.gemm(g, ar.w16(k + "fc2.w"), bias=ar.w32(k + "fc2.b"), epi=EPI_RESID_F32, aux=h2_32, out=u3,
                     drop=dr.site(j, 5))
            h3_16, h3_32, m3, r3 = ops.layernorm_fwd(u3, ar.w32(k + "f_ln.w"), ar.w32(k + "f_ln.b"), eps,
                                                     want_f32=True)
            if save:
                st.layers.append((h16, qkv, a_s, lse_s, u1, m1, r1, h1_16, qc, kvc, a_c, lse_c, u2, m2, r2, h2_16,
                                  hpre, g, u3, m3, r3))
            h16, h32 = h3_16, h3_32
        if cache is not None:
            cache.length = past + T
        ldv = _round_up(V, 8)
        logits = torch.empty((M, ldv), device=h16.device, dtype=torch.bfloat16)
        ops.gemm(h16, ar.w16("dec.tok"), epi=EPI_STORE_BF16, out=logits, N=V)
        st.h_last16 = h16
        st.enc16 = enc16
        return logits, (st if save else None)

    def decoder_backward(self, st, dlogits):
        """dlogits: [B*T, ldv] bf16. Returns d_enc32 [B*S, D] (gradient w.r.t. the encoder output)."""
        ar, bart = self.arena, self.bart
        cfg = bart.config
        D, Hh, nl = cfg.d_model, cfg.decoder_attention_heads, cfg.decoder_layers
        B, T, S, V = st.B, st.T, st.S, st.V
        enc16 = st.enc16
        dr = st.drop
        # lm_head (tied to embed_tokens): dgrad + wgrad
        # dgrad of the LM head: K = V = 50 267 makes each of the 192 output tiles 786 k-blocks long, i.e. 2.6 waves of
        # 0.45 ms on 74 CTA pairs; as a split-K reduce-add into fp32 (which the LayerNorm backward below takes as dy32)
        # the same work is ~13 even waves and the gradient skips one bf16 rounding
        if self.lmhead_dgrad_splitk:
            dy32 = torch.zeros((dlogits.shape[0], D), device=dlogits.device, dtype=torch.float32)
            ops.gemm(dlogits, ar.w16("dec.tok"), b_mn=True, K=V, epi=EPI_REDUCE_F32, out=dy32)
            dy16 = None
        else:
            dy16 = ops.gemm(dlogits, ar.w16("dec.tok"), b_mn=True, K=V)
            dy32 = None
        ops.gemm(dlogits, st.h_last16, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad("dec.tok"), M=V)
        d_enc32 = None
        # weight / bias gradients behind the kernel that produces their dY, on the side stream (see encoder_backward)
        lane = _SideLane(self._side_stream() if self.side_wgrad and dlogits.is_cuda else None, dlogits.device)
        for j in reversed(range(nl)):
            k = f"dec.{j}."
            (h0_16, qkv, a_s, lse_s, u1, m1, r1, h1_16, qc, kvc, a_c, lse_c, u2, m2, r2, h2_16, hpre, g, u3, m3,
             r3) = st.layers[j]
            # final LN (post-LN): du3 = LNbwd(dy)
            # du16 carries the mask of the sub-layer output that was dropped before the residual add (du32 does not)
            du32, du16 = ops.layernorm_bwd(u3, m3, r3, ar.w32(k + "f_ln.w"), ar.grad(k + "f_ln.w"),
                                           ar.grad(k + "f_ln.b"), dy16=dy16, dy32=dy32, out_drop=dr.site(j, 5))
            at = lane.mark()
            d_h = ops.gemm(du16, ar.w16(k + "fc2.w"), b_mn=True, epi=EPI_DGELU_BF16, aux=hpre, drop=dr.site(j, 4))
            ev_du = lane.run(lambda du16=du16, g=g, k=k: ops.gemm(
                du16, g, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "fc2.w"),
                bias_grad=ar.grad(k + "fc2.b")), (du16, g), after=at)
            at = lane.mark()
            d_h2 = ops.gemm(d_h, ar.w16(k + "fc1.w"), b_mn=True)
            lane.run(lambda d_h=d_h, h2_16=h2_16, k=k: ops.gemm(
                d_h, h2_16, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "fc1.w"),
                bias_grad=ar.grad(k + "fc1.b")), (d_h, h2_16), after=at)
            lane.wait(ev_du)            # the LayerNorm backward below overwrites du16
            # cross-attention LN
            du32, du16 = ops.layernorm_bwd(u2, m2, r2, ar.w32(k + "ca_ln.w"), ar.grad(k + "ca_ln.w"),
                                           ar.grad(k + "ca_ln.b"), dy16=d_h2, dy32=du32, dx32=du32, dx16=du16,
                                           out_drop=dr.site(j, 3))
            at = lane.mark()
            d_ac = ops.gemm(du16, ar.w16(k + "ca.o.w"), b_mn=True)
            ev_du = lane.run(lambda du16=du16, a_c=a_c, k=k: ops.gemm(
                du16, a_c, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "ca.o.w"),
                bias_grad=ar.grad(k + "ca.o.b")), (du16, a_c), after=at)
            dqc = torch.empty_like(qc)
            dkvc = torch.empty_like(kvc)
            ops.attention_bwd(qc, kvc, kvc, a_c, d_ac, lse_c, dqc, dkvc, dkvc, B=B, H=Hh, Sq=T, Sk=S, q_col0=0,
                              k_col0=0, v_col0=D, dq_col0=0, dk_col0=0, dv_col0=D, drop=dr.site(j, 2))
            at = lane.mark()
            d_h1 = ops.gemm(dqc, ar.w16(k + "ca.q.w"), b_mn=True)
            lane.run(lambda dqc=dqc, h1_16=h1_16, k=k: ops.gemm(
                dqc, h1_16, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "ca.q.w"),
                bias_grad=ar.grad(k + "ca.q.b")), (dqc, h1_16), after=at)
            wkv = ar.span(k + "ca.k.w", k + "ca.v.w", "w16").view(2 * D, D)
            if d_enc32 is None:
                d_enc32 = ops.gemm(dkvc, wkv, b_mn=True, epi=EPI_STORE_F32)
            else:
                ops.gemm(dkvc, wkv, b_mn=True, epi=EPI_RESID_F32, aux=d_enc32, out=d_enc32)
            lane.run(lambda dkvc=dkvc, k=k: ops.gemm(
                dkvc, enc16, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32,
                out=ar.span(k + "ca.k.w", k + "ca.v.w", "grad").view(2 * D, D),
                bias_grad=ar.span(k + "ca.k.b", k + "ca.v.b", "grad")), (dkvc, enc16), after=at)
            lane.wait(ev_du)            # du16 is overwritten again
            # self-attention LN
            du32, du16 = ops.layernorm_bwd(u1, m1, r1, ar.w32(k + "sa_ln.w"), ar.grad(k + "sa_ln.w"),
                                           ar.grad(k + "sa_ln.b"), dy16=d_h1, dy32=du32, dx32=du32, dx16=du16,
                                           out_drop=dr.site(j, 1))
            at = lane.mark()
            d_as = ops.gemm(du16, ar.w16(k + "sa.o.w"), b_mn=True)
            lane.run(lambda du16=du16, a_s=a_s, k=k: ops.gemm(
                du16, a_s, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32, out=ar.grad(k + "sa.o.w"),
                bias_grad=ar.grad(k + "sa.o.b")), (du16, a_s), after=at)
            dqkv = torch.empty_like(qkv)
            ops.attention_bwd(qkv, qkv, qkv, a_s, d_as, lse_s, dqkv, dqkv, dqkv, B=B, H=Hh, Sq=T, Sk=T, q_col0=0,
                              k_col0=D, v_col0=2 * D, dq_col0=0, dk_col0=D, dv_col0=2 * D, causal=True,
                              drop=dr.site(j, 0))
            wqkv = ar.span(k + "sa.q.w", k + "sa.v.w", "w16").view(3 * D, D)
            at = lane.mark()
            dy16 = ops.gemm(dqkv, wqkv, b_mn=True)
            lane.run(lambda dqkv=dqkv, h0_16=h0_16, k=k: ops.gemm(
                dqkv, h0_16, a_mn=True, b_mn=True, epi=EPI_REDUCE_F32,
                out=ar.span(k + "sa.q.w", k + "sa.v.w", "grad").view(3 * D, D),
                bias_grad=ar.span(k + "sa.q.b", k + "sa.v.b", "grad")), (dqkv, h0_16), after=at)
            dy32 = du32
            st.layers[j] = None
        x_emb, me, re_ = st.emb
        dx_emb, _ = ops.layernorm_bwd(x_emb, me, re_, ar.w32("dec.ln_emb.w"), ar.grad("dec.ln_emb.w"),
                                      ar.grad("dec.ln_emb.b"), dy16=dy16, dy32=dy32, want_bf16=False,
                                      in_drop=dr.emb())
        ops.embed_bwd(st.ids, dx_emb, ar.grad("dec.tok"), ar.grad("dec.pos"), pos_offset=2, scale=1.0,
                      padding_idx=cfg.pad_token_id)
        lane.join()
        if self._grad_ready_hook is not None:
            self._grad_ready_hook("dec.tok", f"dec.{nl - 1}.f_ln.b")
        return d_enc32

    _grad_ready_hook = None
    side_wgrad = bool(int(__import__('os').environ.get('PIXPARSE_B200_SIDE_WGRAD', '0')))   # see encoder_backward
    _side = None
    lmhead_dgrad_splitk = bool(int(__import__('os').environ.get('PIXPARSE_B200_LMHEAD_DGRAD_SPLITK', '1')))
    on_loss_ready = None        # optional callable(stats): invoked right after the CE kernel is enqueued (async loss read-back)

    # ------------------------------------------------------------------------------------------------ public API
    def encode_images(self, image):
        """ImageEncoderTimm.forward: (B, C, H, W) -> (B, S, D) bf16 tokens, CLS first (inference path, no grad)."""
        self.refresh_shadow()
        enc16, _ = self.encoder_forward(image, save=False)
        B = image.shape[0]
        return enc16.view(B, -1, enc16.shape[-1])

    def decode_logits(self, input_ids, encoder_hidden_states, attention_mask=None, past_key_values=None,
                      use_cache=False):
        """TextDecoderHf.forward (teacher-forced / greedy step): logits (B, T, V) bf16 [, DecodeCache].

        attention_mask: (B, past + T), 0 = padding; the reference builds it as input_ids.ne(pad)
        (text_decoder_hf.py:68) and HF combines it with the causal mask of the decoder SELF-attention: a pad token inside
        the prefix (greedy decoding can emit id 1) is hidden from every later query. Cross-attention is unmasked.
        use_cache / past_key_values: the incremental path of prepare_inputs_for_inference (text_decoder_hf.py:69-70):
        only the new tokens are passed, keys / values of the prefix come from the cache."""
        self.refresh_shadow()
        B, S, D = encoder_hidden_states.shape
        enc16 = encoder_hidden_states.reshape(B * S, D)
        if enc16.dtype != torch.bfloat16:
            enc16 = enc16.to(torch.bfloat16)
        enc16 = enc16.contiguous()
        cache = past_key_values
        if cache is None and use_cache:
            cfg = self.bart.config
            cache = DecodeCache(cfg.decoder_layers, B, cfg.max_position_embeddings, cfg.d_model, enc16.device)
        key_mask = None
        if attention_mask is not None:
            past = cache.length if cache is not None else 0
            assert attention_mask.shape == (B, past + input_ids.shape[1]), \
                f"attention_mask {tuple(attention_mask.shape)} must cover the {past} cached + {input_ids.shape[1]} new tokens"
            key_mask = attention_mask.to(device=enc16.device, dtype=torch.uint8).contiguous()
        logits, _ = self.decoder_forward(input_ids, enc16, B, S, save=False, cache=cache, key_mask=key_mask)
        V = self.arena.index["dec.tok"][2][0]
        logits = logits.view(B, input_ids.shape[1], -1)[:, :, :V]
        return (logits, cache) if (use_cache or past_key_values is not None) else logits

    def greedy_decode(self, encoder_hidden_states, prompt_id, max_new_tokens, eos_id, pad_id, stop_on_eos=True):
        """The greedy loop of utils/ocr_utils.py:165-197 for up to 16 pages as a replayed CUDA graph of single-token
        kernels (pixparse_b200/decode.py). Returns ids [B, 1 + n]."""
        from .decode import MAX_PAGES, GreedyDecodeSession
        self.refresh_shadow()
        B, S, D = encoder_hidden_states.shape
        assert B <= MAX_PAGES, f"graph decode handles at most {MAX_PAGES} pages per call"
        enc16 = encoder_hidden_states.reshape(B * S, D)
        if enc16.dtype != torch.bfloat16:
            enc16 = enc16.to(torch.bfloat16)
        enc16 = enc16.contiguous()
        key = (B, S, id(self.arena))
        sess = self._decode_sessions.get(key)
        if sess is None:
            self._decode_sessions.clear()      # one live session: its caches are sized for max_position_embeddings
            sess = self._decode_sessions[key] = GreedyDecodeSession(self, B, S)
        return sess.run(enc16, prompt_id, max_new_tokens, eos_id, pad_id, stop_on_eos=stop_on_eos)

    def forward_logits(self, image, text_ids):
        """Cruller.forward. Under autograd the returned logits carry a backward that runs the fused kernels."""
        self.refresh_shadow()
        if torch.is_grad_enabled():
            return _CrullerLogitsFn.apply(self._anchor(), self, image, text_ids)
        B = image.shape[0]
        enc16, _ = self.encoder_forward(image, save=False)
        logits, _ = self.decoder_forward(text_ids, enc16, B, enc16.shape[0] // B, save=False)
        V = self.arena.index["dec.tok"][2][0]
        return logits.view(B, text_ids.shape[1], -1)[:, :, :V]

    def _anchor(self):
        if getattr(self, '_anchor_t', None) is None or self._anchor_t.device != self.arena.device:
            self._anchor_t = torch.zeros((), device=self.arena.device, requires_grad=True)
        return self._anchor_t

    def zero_grads(self):
        self.ensure_bound().g32.zero_()

    def forward_backward(self, image, text_ids, targets, grad_scale=1.0, stats=None):
        """Fused hot path: fwd -> CE (loss + dlogits in place) -> bwd. Gradients ACCUMULATE into the arena
        (zeroed by the fused optimizer or zero_grads()). Returns a device tensor [n_valid, mean_loss]."""
        self.refresh_shadow()
        B = image.shape[0]
        enc16, st_e = self.encoder_forward(image, save=True)
        S = st_e.S
        logits, st_d = self.decoder_forward(text_ids, enc16, B, S, save=True)
        tflat = targets.reshape(-1)
        if tflat.dtype != torch.int64 or not tflat.is_contiguous():
            tflat = tflat.long().contiguous()
        stats = ops.cross_entropy(logits, tflat, st_d.V, dlogits=logits, grad_scale=grad_scale, stats=stats)
        if self.on_loss_ready is not None:      # the loss exists here, two thirds of the step before its end
            self.on_loss_ready(stats)
        d_enc32 = self.decoder_backward(st_d, logits)
        self.encoder_backward(st_e, d_enc32)
        return stats


class _CrullerLogitsFn(torch.autograd.Function):
    """Compatibility path for callers that follow the reference literally (logits -> nn.CrossEntropyLoss ->
    loss.backward()): forward/backward still run on the fused kernels; gradients land in param.grad."""

    @staticmethod
    def forward(ctx, anchor, engine, image, text_ids):
        B = image.shape[0]
        enc16, st_e = engine.encoder_forward(image, save=True)
        logits, st_d = engine.decoder_forward(text_ids, enc16, B, st_e.S, save=True)
        ctx.engine, ctx.st_e, ctx.st_d = engine, st_e, st_d
        ctx.ld = logits.shape[1]
        return logits.view(B, text_ids.shape[1], -1)[:, :, :st_d.V]

    @staticmethod
    def backward(ctx, dlogits):
        engine, st_e, st_d = ctx.engine, ctx.st_e, ctx.st_d
        ar = engine.arena
        if any(p.grad is None for p in ar.params[:1]):
            ar.g32.zero_()        # the caller dropped the gradients (optimizer.zero_grad(set_to_none=True))
        M = st_d.B * st_d.T
        buf = torch.zeros((M, ctx.ld), device=dlogits.device, dtype=torch.bfloat16)
        buf[:, :st_d.V].copy_(dlogits.reshape(M, st_d.V))
        d_enc32 = engine.decoder_backward(st_d, buf)
        engine.encoder_backward(st_e, d_enc32)
        ar.attach_grads()
        return None, None, None, None
