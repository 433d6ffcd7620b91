// Declarations shared by the two attention-backward kernels (query-major with dropout support: attention_bwd_sm100.cu,
// key-major: attention_bwd_kt_sm100.cu). Separate translation units on purpose: with both kernels in one file the
// query-major dropout variant picked up 80 bytes of spills (it sits exactly at the 128-register limit).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int AB_T = 128;                 // tile edge (queries and keys)
constexpr int AB_D = 64;
constexpr int AB_COMPUTE_THREADS = 256;   // warps 0-7: warp w owns query rows 32*(w%4).., key columns 64*(w/4)..
constexpr int AB_DRAIN_WARP0 = 8;         // warps 8-11: dQ_t TMEM -> shared -> TMA reduce-add, off the compute warps' path
constexpr int AB_TMA_WARP = 12, AB_MMA_WARP = 13;
constexpr int AB_THREADS = (AB_MMA_WARP + 1) * 32;    // 448 threads -> 128 registers per thread
constexpr int AB_TILE = AB_T * AB_D * 2;  // 16 KB bf16 tile
constexpr int AB_SMEM = 12 * AB_TILE + 256 + 1024;   // K, V, Q[2], dO[2], P(2), dS(2), dQ staging(2) = 192 KB

constexpr uint32_t TB_S = 0, TB_DP = 128, TB_DV = 256, TB_DK = 320, TB_DQ = 384, TB_DS = 448;   // TB_DS: dS as bf16 pairs

#ifdef AB_TRACE
#define AB_STAMP(slot)                                                                     \
  do {                                                                                     \
    if (trace_on && t < 8) p.trace[(t * 16 + (slot))] = clock64();                         \
  } while (0)
#else
#define AB_STAMP(slot) do { } while (0)
#endif

struct AttBwdParams {
  int B, H, Sq, Sk, causal;
  float scale, scale_log2;
  const float* lse;      // [B, H, Sq]
  const float* dsum;     // [B, H, Sq]  rowsum(dO * O)
  bf16* dk; long long ld_dk; int dk_col0;
  bf16* dv; long long ld_dv; int dv_col0;
  int q_col0, k_col0, v_col0, do_col0;
  uint32_t drop_threshold16, drop_seed;   // attention-probability dropout of the forward pass (0 = off)
  long long* trace;                       // bring-up builds only (-DAB_TRACE): clock64 stamps of CTA (3,0,0)
};
// key-major kernel only (a separate argument: the query-major dropout kernel sits at the 128-register limit and its
// code generation changed - 80 bytes of spills - when these fields lived in AttBwdParams)
struct AttBwdPadded {
  const float* lse2_pad; // [B, H, Sq_pad] LSE * log2(e), +inf beyond Sq   (Sq_pad = 128 * ceil(Sq / 128))
  const float* dsum_pad; // [B, H, Sq_pad] rowsum(dO * O), 0 beyond Sq
  int Sq_pad;
};


int launch_attention_bwd_kt(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                            const CUtensorMap& tdq, const AttBwdParams& p, const AttBwdPadded& pp, dim3 grid, cudaStream_t s);

}  // namespace b200
