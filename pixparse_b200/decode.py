"""KV-cached greedy decoding as ONE replayed CUDA graph per token (BASELINE configs[4]).

The reference loop (/root/reference/src/pixparse/utils/ocr_utils.py:165-197) re-feeds the whole prefix through
``TextDecoderHf.forward`` every step and takes ``logits[:, -1].argmax``. Here a step feeds the ONE new token per page
through the single-token kernels of ``csrc/decode.cu`` (weight-streaming linears for <= 16 rows, one-query attention
over the caches, LM head fused with the argmax); the position and the generated ids live in device memory, so the whole
step is a fixed kernel sequence that is captured once and replayed -- no host work, no logits in HBM, no host
synchronisation until the caller wants the ids (or, with ``stop_on_eos``, one 4-byte read every ``check_every`` steps).

Same arithmetic contract as the teacher-forced path it replaces (bf16 operands, fp32 accumulation, bf16 q / k / v /
attention / GELU outputs, fp32 residual stream, logits rounded to bf16 before the argmax); decoder self-attention hides
keys whose token is the pad id (attention_mask = input_ids.ne(pad), models/text_decoder_hf.py:68).
"""
import os

import torch

from . import _lib, ops

MAX_PAGES = 16      # activation rows per decode_linear launch (csrc/decode.cu DL_M); larger batches take TextDecoderHf.forward(past_key_values=...)


class GreedyDecodeSession:
    """Buffers + captured graph for greedy decoding of `B` pages against `S` image tokens with this engine's decoder."""

    def __init__(self, engine, B, S):
        assert 1 <= B <= MAX_PAGES
        self.engine = engine
        ar = engine.ensure_bound()
        self.arena = ar
        cfg = engine.bart.config
        self.cfg = cfg
        self.B, self.S = B, S
        D, nl = cfg.d_model, cfg.decoder_layers
        self.D, self.nl, self.H = D, nl, cfg.decoder_attention_heads
        self.t_max = cfg.max_position_embeddings
        self.F = ar.index["dec.0.fc1.w"][2][0]
        self.V = ar.index["dec.tok"][2][0]
        dev = ar.device
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt=bf: torch.zeros(s, device=dev, dtype=dt)
        self.ids = torch.zeros((B, self.t_max + 1), device=dev, dtype=torch.int64)
        self.state = torch.zeros(4, device=dev, dtype=torch.int32)      # pos, done_step, steps run, (pad)
        self.pos = self.state[0:1]
        self.finished = torch.zeros(MAX_PAGES, device=dev, dtype=torch.int32)
        self.self_kv = [z(B, self.t_max, 2 * D) for _ in range(nl)]
        self.cross_kv = [z(B * S, 2 * D) for _ in range(nl)]
        self.x_emb = z(B, D, dt=f32)
        self.h16 = [z(B, D) for _ in range(2)]
        self.h32 = [z(B, D, dt=f32) for _ in range(2)]
        self.q16, self.a16, self.g16 = z(B, D), z(B, D), z(B, self.F)
        self.u32 = z(B, D, dt=f32)
        self.mean, self.rstd = z(B, dt=f32), z(B, dt=f32)
        self.n_cta = ops.decode_linear_ctas(self.V)
        self.partial = torch.zeros((MAX_PAGES, self.n_cta), device=dev, dtype=torch.int64)
        self.graph = None
        self.kernels_per_step = 0
        self.eos_id = None
        self.pad_id = None
        self.pdl = os.environ.get("PIXPARSE_B200_DECODE_PDL", "1") != "0"
        # one launch for the packed q | k | v projection (q dense, k | v appended to the cache): -1.5 % against two launches.
        # (LayerNorm run by the last CTA of the producing linear was tried and measured 10 % SLOWER than its own 4-CTA
        # launch under PDL -- a grid-wide fence + ticket in every CTA and a one-CTA tail cost more than a kernel boundary --
        # and cost every linear 30 registers: removed, profiles/r02_decode_ab.txt)
        self.fuse_qkv = os.environ.get("PIXPARSE_B200_DECODE_FUSE", "1") != "0"

    # ---- one decode step (enqueue only; every per-step quantity is read from device memory) -------------------------
    def _step(self):
        # programmatic dependent launch for the whole step: each kernel's grid becomes resident while its predecessor
        # drains, and the linears issue their (predecessor-independent) weight loads before the grid-dependency wait
        prev = _lib.lib().b200_set_pdl(1 if self.pdl else 0)
        try:
            return self._enqueue_step()
        finally:
            _lib.lib().b200_set_pdl(prev)

    def _enqueue_step(self):
        ar, B, D, H, S = self.arena, self.B, self.D, self.H, self.S
        eps = 1e-5
        n0 = _lib.launch_count()
        ops.decode_embed(self.ids, self.pos, ar.w32("dec.tok"), ar.w32("dec.pos"), self.x_emb, pos_offset=2, scale=1.0)
        cur = 0
        ops.layernorm_fwd(self.x_emb, ar.w32("dec.ln_emb.w"), ar.w32("dec.ln_emb.b"), eps,
                          out=(self.h16[cur], self.h32[cur], self.mean, self.rstd))
        for j in range(self.nl):
            k = f"dec.{j}."
            h16, h32 = self.h16[cur], self.h32[cur]
            n16, n32 = self.h16[cur ^ 1], self.h32[cur ^ 1]

            def ln_after(name, o16, o32):
                ops.layernorm_fwd(self.u32, ar.w32(k + name + ".w"), ar.w32(k + name + ".b"), eps,
                                  out=(o16, o32, self.mean, self.rstd))
            # self-attention: packed q | k | v projection; q dense, k | v appended to the cache at the device-side position
            wqkv = ar.span(k + "sa.q.w", k + "sa.v.w", "w16").view(3 * D, D)
            bqkv = ar.span(k + "sa.q.b", k + "sa.v.b", "w32")
            if self.fuse_qkv:
                ops.decode_linear(h16, wqkv, M=B, out16=self.q16, bias=bqkv, pos=self.pos, out_pos_stride=2 * D,
                                  split=(D, self.self_kv[j], self.t_max * 2 * D))
            else:
                ops.decode_linear(h16, wqkv[:D], M=B, out16=self.q16, bias=bqkv[:D])
                ops.decode_linear(h16, wqkv[D:], M=B, out16=self.self_kv[j], ldo=self.t_max * 2 * D, bias=bqkv[D:],
                                  pos=self.pos, out_pos_stride=2 * D)
            ops.decode_attention(self.q16, self.self_kv[j], self.self_kv[j], self.a16, B=B, H=H, ld_kv=2 * D,
                                 kv_bstride=self.t_max * 2 * D, k_col0=0, v_col0=D, pos=self.pos, key_ids=self.ids,
                                 pad_id=self.pad_id)
            ops.decode_linear(self.a16, ar.w16(k + "sa.o.w"), M=B, out32=self.u32, bias=ar.w32(k + "sa.o.b"), resid=h32)
            ln_after("sa_ln", n16, n32)
            # cross-attention over the cached projection of the image tokens
            ops.decode_linear(n16, ar.w16(k + "ca.q.w"), M=B, out16=self.q16, bias=ar.w32(k + "ca.q.b"))
            ops.decode_attention(self.q16, self.cross_kv[j], self.cross_kv[j], self.a16, B=B, H=H, ld_kv=2 * D,
                                 kv_bstride=S * 2 * D, k_col0=0, v_col0=D, sk=S)
            ops.decode_linear(self.a16, ar.w16(k + "ca.o.w"), M=B, out32=self.u32, bias=ar.w32(k + "ca.o.b"), resid=n32)
            ln_after("ca_ln", h16, h32)
            # feed-forward
            ops.decode_linear(h16, ar.w16(k + "fc1.w"), M=B, out16=self.g16, bias=ar.w32(k + "fc1.b"), act=1)
            ops.decode_linear(self.g16, ar.w16(k + "fc2.w"), M=B, out32=self.u32, bias=ar.w32(k + "fc2.b"), resid=h32)
            ln_after("f_ln", n16, n32)
            cur ^= 1
        # LM head (tied embedding) fused with the argmax, then: append the token, EOS bookkeeping, pos += 1
        ops.decode_linear(self.h16[cur], ar.w16("dec.tok"), M=B, argmax_partial=self.partial)
        ops.decode_finalize(self.partial, self.n_cta, self.ids, self.state, self.finished, self.eos_id)
        return _lib.launch_count() - n0

    def _capture(self):
        # warm-up on a side stream (lazy module loading must not happen inside the capture), then capture one step
        state0 = (self.ids.clone(), self.state.clone(), self.finished.clone())
        side = torch.cuda.Stream(device=self.arena.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.kernels_per_step = self._step()
        _lib.add_launches(-self.kernels_per_step)      # recorded, not launched
        self.ids.copy_(state0[0])
        self.state.copy_(state0[1])
        self.finished.copy_(state0[2])
        self.graph = g

    # ---- public -----------------------------------------------------------------------------------------------------
    def run(self, enc16, prompt_id, max_new_tokens, eos_id, pad_id, stop_on_eos=True, check_every=16):
        """enc16 [B*S, D] bf16 image tokens. Returns ids [B, 1 + n] (prompt first) exactly as the reference loop builds
        them: with stop_on_eos the step at which every page has emitted EOS is NOT appended (ocr_utils.py:192-195)."""
        B, D, S = self.B, self.D, self.S
        ar = self.arena
        assert enc16.shape == (B * S, D) and enc16.dtype == torch.bfloat16
        assert max_new_tokens <= self.t_max, "sequence longer than max_position_embeddings"
        if self.graph is not None and (eos_id, pad_id) != (self.eos_id, self.pad_id):
            self.graph = None      # these two are baked into the captured launches
        self.eos_id, self.pad_id = int(eos_id), int(pad_id)
        for j in range(self.nl):      # image-token K | V of every layer, once per batch of pages (tcgen05 GEMM)
            k = f"dec.{j}."
            ops.gemm(enc16, ar.span(k + "ca.k.w", k + "ca.v.w", "w16").view(2 * D, D),
                     bias=ar.span(k + "ca.k.b", k + "ca.v.b", "w32"), out=self.cross_kv[j])
        self.ids.zero_()
        self.ids[:, 0] = int(prompt_id)
        self.state.copy_(torch.tensor([0, -1, 0, 0], dtype=torch.int32), non_blocking=False)
        self.finished.zero_()
        if self.graph is None:
            self._capture()
        done = -1
        t = 0
        while t < max_new_tokens:
            n = min(check_every if stop_on_eos else max_new_tokens, max_new_tokens - t)
            for _ in range(n):
                self.graph.replay()
            _lib.add_launches(n * self.kernels_per_step)
            t += n
            if stop_on_eos:
                done = int(self.state[1].item())
                if done >= 0:
                    break
        n_tok = done if done >= 0 else t
        return self.ids[:, :n_tok + 1].clone()
