/* pixparse_b200 C-ABI  --  the drop-in boundary of the B200-native Cruller train step.
 *
 * The reference (huggingface/pixparse) has NO FFI layer of its own: every FLOP of its hot path is reached
 * through Python modules (timm VisionTransformer, transformers BartForCausalLM, torch.optim.AdamW, DDP). Each
 * entry point below therefore cites the reference call site whose library kernel(s) it replaces
 * (paths relative to /root/reference/src/pixparse).
 *
 * Contract:
 *   - plain pointers + sizes only; every buffer is owned by the caller (torch caching allocator);
 *     kernels never allocate device memory and keep no global state besides cached function attributes;
 *   - the large operators take ONE pointer to a POD argument struct whose first field, `struct_size`, must equal
 *     sizeof(the struct) as this header declares it: a caller built against another layout is rejected (-1) instead of
 *     being misread. Optional pointers are NULL, optional features are 0. Small helpers keep positional arguments;
 *   - all work is enqueued on the cudaStream_t passed as `stream` (void*), no implicit synchronisation;
 *   - return 0 on success, negative for argument/shape/arch errors, positive = cudaError_t;
 *     b200_last_error() returns a thread-local description; nothing throws across the boundary;
 *   - there is no CPU or alternative-backend fallback: a non-sm_100 device is an error.
 */
#ifndef PIXPARSE_B200_H_
#define PIXPARSE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 3

/* ---- runtime ---------------------------------------------------------------------------------- */
const char* b200_last_error(void);
int b200_abi_version(void);
/* 0 when the current device is sm_100 (B200); replaces framework/device.py:112 `assert torch.cuda.device_count()` */
int b200_device_check(void);
/* Programmatic dependent launch for the kernels that support it: 1 = on, 0 = off, -1 = follow the environment
 * (PIXPARSE_B200_PDL, default off). Returns the previous mode. Process-wide; the decode session brackets its step with it. */
int b200_set_pdl(int mode);

/* ---- GEMM (tcgen05 / TMEM / TMA) ---------------------------------------------------------------
 * D[M,N] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation in tensor memory.
 *   a_mn_major = 0: A is row-major [M][K] (lda elements);  1: A is row-major [K][M]
 *   b_mn_major = 0: B is row-major [N][K] (ldb elements);  1: B is row-major [K][N]
 * Replaces the cuBLASLt calls behind nn.Linear fwd/bwd in timm Attention.qkv/proj, Mlp.fc1/fc2
 * (models/image_encoder_timm.py:13-20), BartAttention q/k/v/out_proj, BartDecoderLayer.fc1/fc2 and lm_head
 * (models/text_decoder_hf.py:13-33), and the patch-embed conv (16x16/16 conv == GEMM over unfolded patches).
 *
 * Dropout (BART trains with it live: bart-base dropout = attention_dropout = activation_dropout = 0.1; the reference
 * never calls model.eval() in train_step) is fused into the RESID / GELU / DGELU epilogues when drop_p > 0.
 * Masks are stateless: keep(seed, element index) from a counter-based hash, regenerated identically in backward; kept
 * values are scaled by 1 / (1 - p).
 */
enum {
  B200_EPI_STORE_BF16 = 0, /* out(bf16)  = acc + bias                                                        */
  B200_EPI_GELU_BF16  = 1, /* out2(bf16) = h = acc + bias ; out(bf16) = drop(gelu_erf(h))   (timm Mlp / BART fc1) */
  B200_EPI_RESID_F32  = 2, /* out(f32)   = aux(f32) + drop(acc + bias)   (residual stream; out may alias aux) */
  B200_EPI_DGELU_BF16 = 3, /* out(bf16)  = mask * acc * gelu_erf'(aux(bf16))   (fc2 dgrad fused with GELU backward) */
  B200_EPI_REDUCE_F32 = 4, /* out(f32)  += acc   via TMA reduce-add, split-K capable (weight gradients)       */
  B200_EPI_STORE_F32  = 5  /* out(f32)   = acc + bias                                                         */
};
typedef struct B200GemmArgs {
  unsigned int struct_size;
  int epilogue;
  const void* a;
  long long lda;
  int a_mn_major;
  const void* b;
  long long ldb;
  int b_mn_major;
  int m;
  int n;
  int k;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const float* bias;
  const void* aux;
  long long ld_aux;
  /* REDUCE_F32 with A MN-major only (weight gradient dW = dY^T X): bias_grad[m] += sum_k A(m, k), i.e. the column sums
   * of dY = the nn.Linear bias gradient, accumulated from the A tiles already staged in shared memory (no extra pass
   * over dY in HBM). NULL = off. */
  float* bias_grad;
  int splits;            /* 0 = choose (wave-aware split-K for REDUCE_F32) */
  int block_n;           /* 0 = choose; 128 or 256 */
  float drop_p;
  unsigned int drop_seed;
  /* Dynamic tile scheduling. NULL: the persistent grid deals its work items statically (CTA i takes items i, i + grid, ...).
   * Otherwise a caller-owned device counter that is ZERO when the kernel starts; CTAs draw items from it with one
   * atomicAdd each and the kernel leaves it at zero again, so one counter serves any number of launches that do not
   * overlap in time (kernels that may run concurrently, e.g. on two streams, need different counters). Results are
   * identical; what changes is that kernels sharing the GPU across streams split its SMs work-conservingly (a weight
   * gradient on a side stream fills the partly empty last wave of the main stream's kernel). */
  unsigned int* tile_counter;
  /* Tail split (STORE_BF16 / RESID_F32, static tile deal). When the tiles do not fill the persistent grid's last wave
   * (N = 768 outputs of the train step: 381 tiles on 74 CTA pairs = 5.15 waves) the tiles of that wave are cut along K
   * into one piece per CTA that would otherwise idle; the pieces sum their fp32 accumulators in this workspace and the
   * last piece to arrive runs the epilogue. Caller-owned, 128-byte aligned, ZERO on entry (the kernel leaves it zero), at
   * least b200_gemm_tail_workspace_bytes() for every shape; launches sharing it must not overlap in time. NULL = off;
   * also off unless b200_debug_gemm_tail_split(1) was called (opt-in, see there). */
  void* tail_workspace;
  long long tail_workspace_bytes;
} B200GemmArgs;
int b200_gemm_bf16(const B200GemmArgs* args, void* stream);
long long b200_gemm_tail_workspace_bytes(void);

/* ---- attention (tcgen05, flash-style, head_dim 64) ---------------------------------------------
 * q/k/v are [B, S, row] bf16 activations with `ld*` elements between tokens and `*_bstride` elements between batches
 * (0 = densely packed [B, S, ld]); head h occupies columns [*_col0 + 64 h, *_col0 + 64 h + 64). out is [B, Sq, ld_out]
 * (head h at column 64 h); lse is [B, H, Sq] fp32 (may be NULL).
 * Replaces F.scaled_dot_product_attention in timm Attention (encoder, non-causal) and BartAttention
 * (decoder causal self-attention and cross-attention; modeling_bart.py:185-258). Explicit batch strides serve the
 * KV-cached greedy decode, whose self-attention keys/values live in a pre-allocated [B, T_max, 2D] cache
 * (models/text_decoder_hf.py:69-70 is the past_key_values branch; utils/ocr_utils.py:165-197 the loop).
 * key_mask: optional [B, Sk] bytes, 0 = this key is masked for every query of the batch row -- the decoder
 * attention_mask the reference builds as input_ids.ne(pad) (models/text_decoder_hf.py:68). Inference path only.
 * drop_p > 0: dropout on the softmax probabilities (normaliser computed before dropping, as F.dropout(softmax)).
 */
typedef struct B200AttentionFwdArgs {
  unsigned int struct_size;
  int batch;
  const void* q;
  long long ldq;
  long long q_bstride;
  int q_col0;
  int k_col0;
  const void* k;
  long long ldk;
  long long k_bstride;
  const void* v;
  long long ldv;
  long long v_bstride;
  int v_col0;
  int heads;
  void* out;
  long long ld_out;
  long long out_bstride;
  float* lse;
  const unsigned char* key_mask;
  long long key_mask_bstride;
  int sq;
  int sk;
  int head_dim;
  int causal;
  float scale;
  float drop_p;
  unsigned int drop_seed;
  int reserved;
} B200AttentionFwdArgs;
int b200_attention_fwd(const B200AttentionFwdArgs* args, void* stream);

/* backward: dq/dk/dv (bf16, strided like q/k/v) from d_o; `o` and `lse` are the forward outputs.
 * workspace: b200_attention_bwd_workspace_bytes(B, H, Sq) bytes of device memory (fp32 dQ accumulator + row sums). */
long long b200_attention_bwd_workspace_bytes(int B, int H, int Sq);
typedef struct B200AttentionBwdArgs {
  unsigned int struct_size;
  int batch;
  const void* q;
  long long ldq;
  const void* k;
  long long ldk;
  const void* v;
  long long ldv;
  int q_col0;
  int k_col0;
  int v_col0;
  int do_col0;
  const void* o;
  long long ld_o;
  const void* d_o;
  long long ld_do;
  const float* lse;
  void* dq;
  long long ld_dq;
  void* dk;
  long long ld_dk;
  void* dv;
  long long ld_dv;
  int dq_col0;
  int dk_col0;
  int dv_col0;
  int heads;
  void* workspace;
  int sq;
  int sk;
  int head_dim;
  int causal;
  float scale;
  float drop_p;
  unsigned int drop_seed;
  int reserved;
} B200AttentionBwdArgs;
int b200_attention_bwd(const B200AttentionBwdArgs* args, void* stream);

/* ---- LayerNorm (timm Block.norm1/norm2/norm; BART layernorm_embedding / *_layer_norm) --------------
 * fwd: y = (x - mean) * rstd * gamma + beta; x fp32 [rows, dim]; optional bf16 and/or fp32 outputs; mean/rstd saved;
 *      drop_p > 0 drops the normalised OUTPUT (dropout(layernorm_embedding(x))).
 * bwd: dy = dy_bf16 + dy_f32 (either may be NULL); dx = LNbwd(dy) + dres_f32 (optional residual-path gradient);
 *      dgamma/dbeta are accumulated (+=). in_p/in_seed: the forward's output mask applied to the incoming gradient;
 *      out_p/out_seed: mask of the sub-layer output feeding this LayerNorm, applied to dx_bf16 only.
 */
typedef struct B200LayerNormFwdArgs {
  unsigned int struct_size;
  int rows;
  const float* x;
  const float* gamma;
  const float* beta;
  void* y_bf16;
  float* y_f32;
  float* mean;
  float* rstd;
  int dim;
  float eps;
  float drop_p;
  unsigned int drop_seed;
} B200LayerNormFwdArgs;
int b200_layernorm_fwd(const B200LayerNormFwdArgs* args, void* stream);

typedef struct B200LayerNormBwdArgs {
  unsigned int struct_size;
  int rows;
  const void* dy_bf16;
  const float* dy_f32;
  const float* dres_f32;
  const float* x;
  const float* mean;
  const float* rstd;
  const float* gamma;
  float* dx_f32;
  void* dx_bf16;
  float* dgamma;
  float* dbeta;
  int dim;
  float in_p;
  unsigned int in_seed;
  float out_p;
  unsigned int out_seed;
  int reserved;
} B200LayerNormBwdArgs;
int b200_layernorm_bwd(const B200LayerNormBwdArgs* args, void* stream);

/* ---- HBM-bound helpers ---------------------------------------------------------------------------- */
/* out[n] += sum_m dy[m, n]  (stand-alone column sums; the train step gets its bias gradients from B200GemmArgs.bias_grad) */
int b200_colsum_bf16(const void* dy, long long ld, int rows, int cols, float* out, void* stream);
/* timm PatchEmbed: unfold (B, C, H, W) fp32 pixels into [B*gh*gw, C*P*P] bf16 rows for the patch GEMM */
int b200_patch_unfold(const float* image, void* patches_bf16, int B, int C, int H, int W, int P,
                      long long ld_patches, void* stream);
/* timm _pos_embed: x[b,0] = cls + pos[0]; x[b,1+p] = proj[b,p] + pos[1+p]; and its backward */
int b200_tokens_assemble(const void* proj_bf16, const float* cls, const float* pos, float* x, int B, int S, int D,
                         void* stream);
int b200_tokens_assemble_bwd(const float* dx, void* dproj_bf16, float* dcls, float* dpos, int B, int S, int D,
                             void* stream);
/* BartDecoder embedding: x[b,t] = embed_tokens[ids[b,t]] * scale + embed_positions[t + pos_offset]; and backward
 * (scatter-add, skipping padding_idx like nn.Embedding(padding_idx)) */
int b200_embed_fwd(const long long* ids, const float* tok_emb, const float* pos_emb, float* x, int B, int T, int D,
                   int pos_offset, float scale, void* stream);
int b200_embed_bwd(const long long* ids, const float* dx, float* d_tok_emb, float* d_pos_emb, int B, int T, int D,
                   int pos_offset, float scale, long long padding_idx, void* stream);
int b200_cast_f32_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* dst[i] = (dst[i] + sum_{s != skip} stage[s * stride + i]) * scale for i < n: the local reduction of the copy-engine
 * gradient exchange that stands in for DistributedDataParallel's all-reduce (task/task_cruller_pretrain.py:181-189):
 * dst = this rank's share of a gradient bucket, stage = the copies of that share its peers pushed over NVLink */
int b200_reduce_shards(float* dst, const float* stage, long long stride, int nsrc, int skip, float scale, long long n,
                       void* stream);

/* ---- loss + optimizer ------------------------------------------------------------------------------
 * nn.CrossEntropyLoss(ignore_index) fwd+bwd in one pass (task_cruller_pretrain.py:118,251-254):
 *   ce_prepare: stats[0] = #valid targets, stats[1] = 0
 *   ce_fwd_bwd: stats[1] += mean loss; dlogits = (softmax - onehot) * grad_scale / #valid (may alias logits)
 * grad_norm: deterministic global L2 norm of the flat gradient arena; out4 = {sumsq, norm, clip coefficient, updates}
 *   (timm dispatch_clip_grad 'norm' -> clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)); max_norm <= 0: coef 1).
 *   out4[3] is a counter the CALLER keeps across steps: it is incremented when the norm is finite, i.e. it counts the
 *   optimizer updates that were really applied (torch.optim.AdamW's state['step'] under a GradScaler that skips).
 * adamw_step: torch.optim.AdamW semantics over the flat arena, per-tensor lr_scale / weight_decay segments
 *   {int64 end; float lr_scale; float weight_decay}; also refreshes the bf16 shadow weights and zeroes grads.
 *   With norm_stats (the out4 of grad_norm): gradients are scaled by norm_stats[2], the bias corrections use
 *   norm_stats[3] as the step number, and a non-finite norm SKIPS the parameter / moment update (what timm's
 *   NativeScaler / GradScaler.step does, task_cruller_pretrain.py:259-268) while the gradients are still zeroed.
 *   Without norm_stats: `step` is the 1-based update number and nothing is ever skipped.
 */
int b200_ce_prepare(const long long* targets, int n, long long ignore_index, float* stats, void* stream);
int b200_ce_fwd_bwd(const void* logits_bf16, long long ld, const long long* targets, void* dlogits_bf16,
                    long long ldd, float* row_loss, float* stats, int rows, int vocab, long long ignore_index,
                    float grad_scale, void* stream);
int b200_grad_norm(const float* grads, long long n, float* workspace, float* out4, float max_norm, float pre_scale,
                   void* stream);
int b200_grad_norm_workspace_floats(void);
typedef struct B200AdamWArgs {
  unsigned int struct_size;
  int num_segments;
  float* params;
  float* grads;
  float* exp_avg;
  float* exp_avg_sq;
  void* params_bf16;
  long long n;
  const void* segments;
  const float* norm_stats;
  float grad_scale;
  float lr;
  float beta1;
  float beta2;
  float eps;
  int step;
  int zero_grad;
  int reserved;
} B200AdamWArgs;
int b200_adamw_step(const B200AdamWArgs* args, void* stream);

/* ---- single-token greedy decode (BASELINE configs[4]) -----------------------------------
 * One decode step feeds one new token per page: every linear has M = pages <= 16 rows and is bound by streaming its
 * weights once. Per-step quantities (the position, the generated ids) are read from DEVICE memory, so a whole step is a
 * fixed kernel sequence with fixed arguments that the host captures once in a CUDA graph and replays per token.
 * Replaces the uncached loop of utils/ocr_utils.py:165-197 / the past_key_values branch of
 * models/text_decoder_hf.py:69-70 for batch <= 16.
 *
 * decode_linear: y[M, N] = x[M, K] W[N, K]^T (+ bias) (act = 1: GELU on the bf16-rounded pre-activation) (+ fp32 resid);
 *   outputs bf16 and/or fp32 with row pitch ldo, shifted by (*pos) * out_pos_stride elements when pos != NULL (appending
 *   K | V to a [B, T_max, 2D] cache). With out2_bf16 the output is split at column n_split: [0, n_split) -> out_bf16 / out_f32
 *   unshifted, [n_split, N) -> out2_bf16 shifted (the packed q | k | v projection: q dense, k | v appended to the cache).
 *   With argmax_partial != NULL nothing is stored: the per-CTA (max, argmax) of the
 *   bf16-rounded outputs is written as [16][b200_decode_linear_ctas(N)] packed 64-bit keys (LM head fused with argmax).
 * decode_attention: one query per (page, head): q [B, ldq], K / V rows ld_kv apart, pages kv_bstride apart; the key
 *   count is sk, or (*pos) + 1 when pos != NULL; key j of page b is hidden when key_ids[b * ld_ids + j] == pad_id
 *   (attention_mask = input_ids.ne(pad), models/text_decoder_hf.py:68).
 * decode_embed: x[b] = tok_emb[ids[b * ld_ids + *pos]] * scale + pos_emb[*pos + pos_offset]   (BartDecoder embedding).
 * decode_finalize: token[b] = argmax over the partial keys (first index on ties); ids[b * ld_ids + *pos + 1] = token;
 *   finished[b] |= token == eos; state = {pos, done_step, steps}: done_step (-1 until set) = first step at which every
 *   page had emitted EOS -- the reference loop breaks there without appending; pos += 1.
 */
typedef struct B200DecodeLinearArgs {
  unsigned int struct_size;
  int m;
  const void* x;
  long long ldx;
  const void* w;
  long long ldw;
  const float* bias;
  const float* resid;
  long long ld_resid;
  void* out_bf16;
  float* out_f32;
  long long ldo;
  const int* pos;
  long long out_pos_stride;
  void* argmax_partial;
  int n;
  int k;
  int act;
  int n_split;             /* with out2_bf16: columns >= n_split go to out2 (column n - n_split) and only they are shifted */
  void* out2_bf16;
  long long ldo2;
} B200DecodeLinearArgs;
int b200_decode_linear(const B200DecodeLinearArgs* args, void* stream);
int b200_decode_linear_ctas(int n);

typedef struct B200DecodeAttentionArgs {
  unsigned int struct_size;
  int batch;
  const void* q;
  long long ldq;
  const void* k;
  const void* v;
  long long ld_kv;
  long long kv_bstride;
  void* out;
  long long ld_out;
  const int* pos;
  const long long* key_ids;
  long long ld_ids;
  long long pad_id;
  int q_col0;
  int k_col0;
  int v_col0;
  int heads;
  int head_dim;
  int sk;
  float scale;
  int reserved;
} B200DecodeAttentionArgs;
int b200_decode_attention(const B200DecodeAttentionArgs* args, void* stream);

int b200_decode_embed(const long long* ids, long long ld_ids, const int* pos, const float* tok_emb, const float* pos_emb,
                      float* x, int B, int D, int pos_offset, float scale, void* stream);
int b200_decode_finalize(const void* argmax_partial, int n_cta, long long* ids, long long ld_ids, int* state,
                         int* finished, int B, long long eos_id, void* stream);

/* ---- on-device page preprocessing ------------------------------------------------------
 * uint8 'L' pages [B, Hin, Win] (page_stride bytes apart) -> fp32 [B, 1, Hout, Wout]:
 * ToTensor -> Resize(BICUBIC, antialias=True) -> Normalize(mean, std), i.e. the transforms.Compose built in
 * task/task_cruller_pretrain.py:132-143 (ATen _upsample_bicubic2d_aa). workspace: b200_preprocess_workspace_bytes().
 */
long long b200_preprocess_workspace_bytes(int Hout, int Wout);
int b200_preprocess_pages(const void* pages_u8, int B, int Hin, int Win, long long page_stride, float* out, int Hout,
                          int Wout, float mean, float std_, void* workspace, void* stream);

/* bring-up aid: override the UMMA shared-memory descriptor fields (-1 keeps the default) */
int b200_debug_gemm_desc(int a_lbo, int a_sbo, int a_kadv, int b_lbo, int b_sbo, int b_kadv);
/* Bring-up / A-B switch: on = 1 makes b200_gemm_bf16 use one CTA per 128-row tile everywhere instead of CTA pairs
 * (tcgen05 cta_group::2, 256-row tiles) for the 256-column tile width. Results are identical either way. */
int b200_debug_gemm_single_cta(int on);
/* Opt-in switch of the tail split: on = 1 lets b200_gemm_bf16 use B200GemmArgs.tail_workspace. Default 0 (measured
 * slower on the train step's shapes: the fix-up costs more than the idle last wave it removes). */
int b200_debug_gemm_tail_split(int on);
/* Kernel selection for b200_attention_bwd without dropout: on = 1 (default) query-major accumulators (the kernel the
 * dropout variant always uses), on = 0 key-major accumulators (P^T / dS^T fed to the dV / dK MMAs from tensor memory).
 * Results agree to bf16 rounding; the two are on par on B200 (profiles/), the switch exists for A-B runs. */
int b200_debug_attention_bwd_query_major(int on);

#ifdef __cplusplus
}
#endif
#endif /* PIXPARSE_B200_H_ */
