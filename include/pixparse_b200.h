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

#define B200_ABI_VERSION 1

/* ---- runtime ---------------------------------------------------------------------------------- */
const char* b200_last_error(void);
int b200_abi_version(void);
/* 0 when the current device is sm_100 (B200); replaces framework/device.py:112 `assert torch.cuda.device_count()` */
int b200_device_check(void);

/* ---- GEMM (tcgen05 / TMEM / TMA) ---------------------------------------------------------------
 * D[M,N] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation in tensor memory.
 *   a_mn_major = 0: A is row-major [M][K] (lda elements);  1: A is row-major [K][M]
 *   b_mn_major = 0: B is row-major [N][K] (ldb elements);  1: B is row-major [K][N]
 * Replaces the cuBLASLt calls behind nn.Linear fwd/bwd in timm Attention.qkv/proj, Mlp.fc1/fc2
 * (models/image_encoder_timm.py:13-20), BartAttention q/k/v/out_proj, BartDecoderLayer.fc1/fc2 and lm_head
 * (models/text_decoder_hf.py:13-33), and the patch-embed conv (16x16/16 conv == GEMM over unfolded patches).
 */
enum {
  B200_EPI_STORE_BF16 = 0, /* out(bf16)  = acc + bias                                                        */
  B200_EPI_GELU_BF16  = 1, /* out2(bf16) = h = acc + bias ; out(bf16) = gelu_erf(h)   (timm Mlp / BART fc1)   */
  B200_EPI_RESID_F32  = 2, /* out(f32)   = aux(f32) + acc + bias   (residual stream; out may alias aux)      */
  B200_EPI_DGELU_BF16 = 3, /* out(bf16)  = acc * gelu_erf'(aux(bf16))   (fc2 dgrad fused with GELU backward)  */
  B200_EPI_REDUCE_F32 = 4, /* out(f32)  += acc   via TMA reduce-add, split-K capable (weight gradients)       */
  B200_EPI_STORE_F32  = 5  /* out(f32)   = acc + bias                                                         */
};
int b200_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                   int M, int N, int K, int epilogue, void* out, long long ldo, void* out2, long long ldo2,
                   const float* bias, const void* aux, long long ld_aux, int splits, int block_n, void* stream);

/* bring-up aid: override the UMMA shared-memory descriptor fields (-1 keeps the default) */
int b200_debug_gemm_desc(int a_lbo, int a_sbo, int a_kadv, int b_lbo, int b_sbo, int b_kadv);

#ifdef __cplusplus
}
#endif
#endif /* PIXPARSE_B200_H_ */
