// Flash-style attention forward for sm_100a, head_dim 64, bf16 in / bf16 out, fp32 softmax statistics.
//
//   S = Q K^T        tcgen05.mma  (A = Q tile, B = K tile, both K-major in shared memory via TMA)  -> TMEM
//   P = softmax(S)   one thread per query row: tcgen05.ld, online max/sum in registers, exp2, bf16 pack,
//                    tcgen05.st back into TMEM (P never touches shared memory)
//   O += P V         tcgen05.mma  (A = P from TMEM, B = V tile MN-major in shared memory)         -> TMEM
//
// One CTA per (128-query tile, head, batch) walks the keys 64 at a time; P aliases the first half of S's TMEM columns,
// so a CTA needs only 128 TMEM columns and 48 KB of shared memory: FOUR CTAs are co-resident per SM (16 softmax warps)
// and the MMA / TMA / barrier latencies of one CTA hide under the exp2 work of the others (measured: 4 CTAs/SM with two
// TMEM passes over S beats 3 CTAs/SM with S held in registers, 0.20 ms vs 0.26 ms per encoder layer).
//
// Serves the three attention shapes of the Cruller step:
//   encoder self-attention (non-causal, Sq = Sk = 1009 / 2509), decoder causal self-attention (Sq = Sk = T),
//   decoder cross-attention over image tokens (Sq = T, Sk = S, no mask).
// Replaces torch SDPA reached from timm Attention (fused_attn) and BartAttention (sdpa).
#include <type_traits>

#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 64;    // keys per inner tile
constexpr int ATT_D = 64;     // head dim
constexpr int ATT_THREADS = 192;
constexpr int ATT_Q_BYTES = ATT_BM * ATT_D * 2;    // 16 KB
constexpr int ATT_KV_BYTES = ATT_BN * ATT_D * 2;   // 8 KB
constexpr int ATT_SMEM = ATT_Q_BYTES + 4 * ATT_KV_BYTES + 256 + 1024;
constexpr int ATT_TMEM_COLS = 128;

constexpr uint32_t TM_S = 0;      // 64 columns: S (fp32)
constexpr uint32_t TM_P = 0;      // 32 columns: P (bf16 pairs) -- aliases S: chunk c of P only overwrites S columns
                                  //             that were already consumed; S_{j+1} is issued behind P_j V in order
constexpr uint32_t TM_O = 64;     // 64 columns: O (fp32)

struct AttFwdParams {
  int B, H, Sq, Sk;
  int causal;
  float scale_log2;        // softmax scale * log2(e)
  bf16* out;               // token (b, q) at out + b * out_bstride + q * ld_out, head h at column h*64
  long long ld_out, out_bstride;
  float* lse;              // [B, H, Sq] natural-log logsumexp of the scaled scores
  int q_col0, k_col0, v_col0;   // column offsets of head 0 inside the Q / K / V row
  uint32_t drop_threshold16, drop_seed;   // attention-probability dropout (BART attention_dropout); 0 = off
  const uint8_t* key_mask;                // [B, Sk] bytes (key_mask_bstride apart), 0 = key hidden from every query; KMASK kernels only
  long long key_mask_bstride;
};

// KMASK: key-padding mask (decoder attention_mask = input_ids.ne(pad), text_decoder_hf.py:68); inference path, no dropout
template <bool DROP, bool KMASK>
__global__ void __launch_bounds__(ATT_THREADS, DROP ? 4 : 3)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_Q_BYTES;                       // 2 stages
  uint8_t* sV = smem + ATT_Q_BYTES + 2 * ATT_KV_BYTES;    // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_Q_BYTES + 4 * ATT_KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [2]
  uint64_t* kv_empty = bars + 5;   // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = gridDim.x - 1 - blockIdx.x;   // heavy (late, causal) tiles first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = q_tile * ATT_BM;

  int n_tiles = (p.Sk + ATT_BN - 1) / ATT_BN;
  if (p.causal) {
    const int last_key = q0 + ATT_BM - 1 + (p.Sk - p.Sq);   // largest key index any row of this tile may see
    const int lim = last_key / ATT_BN + 1;
    if (lim < n_tiles) n_tiles = lim;
  }

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 4) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_3d(sQ, &tmap_q, q_full, p.q_col0 + h * ATT_D, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[s], ATT_KV_BYTES);
        tma_load_3d(sK + s * ATT_KV_BYTES, &tmap_k, &k_full[s], p.k_col0 + h * ATT_D, j * ATT_BN, b);
        mbar_expect_tx(&v_full[s], ATT_KV_BYTES);
        tma_load_3d(sV + s * ATT_KV_BYTES, &tmap_v, &v_full[s], p.v_col0 + h * ATT_D, j * ATT_BN, b);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, false, false);   // S = Q K^T
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, false, true);     // O = P V (V is MN-major)
      mbar_wait(q_full, 0);
      const uint64_t dq = make_smem_desc(smem_u32(sQ), 16, 1024);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[s], ph);
        tc_fence_after();
        const uint64_t dk = make_smem_desc(smem_u32(sK + s * ATT_KV_BYTES), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_ss(tmem_base + TM_S, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
        // P_j ready (and O rescaled) -> O += P V
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
        const uint64_t dv = make_smem_desc(smem_u32(sV + s * ATT_KV_BYTES), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k)
            umma_ts(tmem_base + TM_O, tmem_base + TM_P + 8 * k, dv + (uint64_t)(128 * k), idesc_o,
                    (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&kv_empty[s]);
          umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax / correction / output (128 threads, one query row each) =====================
    const int row = warp * 32 + lane;
    const int qidx = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int causal_shift = p.Sk - p.Sq;
    // dropout pair index of (row, key k) = drop_row + k / 2  (rows padded to an even number of keys)
    const uint32_t drop_row = (uint32_t)((((long long)b * p.H + h) * p.Sq + qidx) * ((p.Sk + 1) >> 1));
    const float drop_sc = dropout_scale(p.drop_threshold16);
    float m_run = -INFINITY;   // running max in the scaled log2 domain
    float l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int k0 = j * ATT_BN;
      const bool need_mask = KMASK || (k0 + ATT_BN > p.Sk) || (p.causal && (k0 + ATT_BN - 1 > q0 + causal_shift));
      int kmax = p.Sk - 1;                                   // largest visible key index for this row
      if (p.causal) kmax = min(kmax, qidx + causal_shift);
      uint64_t kbits = ~0ull;                                // bit i: key k0 + i is not padding
      if (KMASK) {
        const uint8_t* mrow = p.key_mask + (long long)b * p.key_mask_bstride + k0;
        kbits = 0ull;
#pragma unroll 8
        for (int i = 0; i < ATT_BN; ++i)
          if (k0 + i < p.Sk && mrow[i] != 0) kbits |= 1ull << i;
      }
      auto vis = [&](int i) { return (k0 + i <= kmax) && (!KMASK || ((kbits >> i) & 1ull) != 0ull); };
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // Tensor-memory reads bound this kernel (two 32 KB passes over S per tile = ~1000 cycles at the measured ~64 B/clk
      // against 512 MMA / 512 MUFU cycles), so only the first tile takes the exact two-pass route (row max, then exp).
      // Later tiles read S ONCE: P = exp2(S*c - m_run) against the running reference, and a row whose P outgrows 2^16
      // is renormalised by an exact power of two (P, its sum, the running O and l), which moves the reference instead
      // of re-reading S. The result O / l and the logsumexp do not depend on the reference chosen.
      // (the dropout variant keeps the two-pass route throughout: its mask code needs the registers, 4 CTAs per SM)
      bool exact_pass = DROP || (j == 0) || __any_sync(0xffffffffu, m_run == -INFINITY);     // warp-uniform
      if (!DROP && !exact_pass) {
        uint32_t pk[ATT_BN / 2];
        float l_tile = 0.f, pmax = 0.f;
        auto single = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
          for (int c = 0; c < ATT_BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(lane_addr + TM_S + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * e]), p.scale_log2, -m_run));
              float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * e + 1]), p.scale_log2, -m_run));
              if (MASKED) {
                if (!vis(c * 32 + 2 * e)) p0 = 0.f;
                if (!vis(c * 32 + 2 * e + 1)) p1 = 0.f;
              }
              pmax = fmaxf(pmax, fmaxf(p0, p1));
              l_tile += p0 + p1;        // the softmax normaliser uses the un-dropped probabilities
              if (DROP)
                dropout_pair(p.drop_seed, drop_row + (uint32_t)((k0 + c * 32 + 2 * e) >> 1), p.drop_threshold16, drop_sc,
                             p0, p1);
              pk[c * 16 + e] = pack_bf16(p0, p1);
            }
          }
        };
        if (need_mask) single(std::true_type{}); else single(std::false_type{});
        if (__any_sync(0xffffffffu, !(pmax <= 1.8446744e19f))) {
          // some row jumped by more than 2^64 against its reference (or overflowed to inf): nothing has been written
          // yet, S is intact in TMEM, so this tile is simply redone on the exact two-pass route
          exact_pass = true;
        } else {
          // exact power-of-two renormalisation of this row when the reference is more than 2^16 too small
          float alpha = 1.0f;
          if (pmax > 65536.0f) {
            const int kexp = (int)((__float_as_uint(pmax) >> 23) & 0xffu) - 127;      // floor(log2(pmax))
            alpha = __uint_as_float((uint32_t)(127 - kexp) << 23);                   // 2^-kexp
  #pragma unroll
            for (int e = 0; e < ATT_BN / 2; ++e) pk[e] = pack_bf16(bf16_lo(pk[e]) * alpha, bf16_hi(pk[e]) * alpha);
            l_tile *= alpha;
            m_run += (float)kexp;
          }
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
  #pragma unroll 1
            for (int c = 0; c < ATT_D / 32; ++c) {
              uint32_t r[32];
              tmem_ld_32x32(lane_addr + TM_O + c * 32, r);
              tmem_ld_wait();
  #pragma unroll
              for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
              uint32_t lo[16], hi[16];
  #pragma unroll
              for (int e = 0; e < 16; ++e) {
                lo[e] = r[e];
                hi[e] = r[16 + e];
              }
              tmem_st_32x16(lane_addr + TM_O + c * 32, lo);
              tmem_st_32x16(lane_addr + TM_O + c * 32 + 16, hi);
            }
          }
  #pragma unroll
          for (int c = 0; c < ATT_BN / 32; ++c) {
            uint32_t w16[16];
  #pragma unroll
            for (int e = 0; e < 16; ++e) w16[e] = pk[c * 16 + e];
            tmem_st_32x16(lane_addr + TM_P + c * 16, w16);
          }
          tmem_st_wait();
          l_run = l_run * alpha + l_tile;
      
        }
      }
      if (exact_pass) {
        // ---- pass 1: row max
        float mx = -INFINITY;
  #pragma unroll 1
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + TM_S + c * 32, r);
          tmem_ld_wait();
          if (need_mask) {
  #pragma unroll
            for (int e = 0; e < 32; ++e)
              if (vis(c * 32 + e)) mx = fmaxf(mx, __uint_as_float(r[e]));
          } else {
  #pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
          }
        }
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;   // fully masked so far: keep exp2 finite
        const float alpha = ex2_approx(m_run - m_use);                  // 0 on the first tile (m_run = -inf)
        // ---- rescale the running O (TMEM) once the previous P V has retired
        if (j > 0) {
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
  #pragma unroll 1
            for (int c = 0; c < ATT_D / 32; ++c) {
              uint32_t r[32];
              tmem_ld_32x32(lane_addr + TM_O + c * 32, r);
              tmem_ld_wait();
  #pragma unroll
              for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
              uint32_t lo[16], hi[16];
  #pragma unroll
              for (int e = 0; e < 16; ++e) {
                lo[e] = r[e];
                hi[e] = r[16 + e];
              }
              tmem_st_32x16(lane_addr + TM_O + c * 32, lo);
              tmem_st_32x16(lane_addr + TM_O + c * 32 + 16, hi);
            }
          }
        }
        // ---- pass 2: P = exp2(S * scale - m), row sum, bf16 pack -> TMEM
        // (two instantiations: interior tiles carry no masking instructions at all)
        float l_tile = 0.f;
        auto pass2 = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
  #pragma unroll 1
          for (int c = 0; c < ATT_BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(lane_addr + TM_S + c * 32, r);
            tmem_ld_wait();
            uint32_t pk[16];
  #pragma unroll
            for (int e = 0; e < 16; ++e) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * e]), p.scale_log2, -m_use));
              float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * e + 1]), p.scale_log2, -m_use));
              if (MASKED) {
                if (!vis(c * 32 + 2 * e)) p0 = 0.f;
                if (!vis(c * 32 + 2 * e + 1)) p1 = 0.f;
              }
              l_tile += p0 + p1;        // the softmax normaliser uses the un-dropped probabilities
              if (DROP)
                dropout_pair(p.drop_seed, drop_row + (uint32_t)((k0 + c * 32 + 2 * e) >> 1), p.drop_threshold16, drop_sc,
                             p0, p1);
              pk[e] = pack_bf16(p0, p1);
            }
            tmem_st_32x16(lane_addr + TM_P + c * 16, pk);
          }
        };
        if (need_mask) pass2(std::true_type{}); else pass2(std::false_type{});
        tmem_st_wait();
        l_run = l_run * alpha + l_tile;
        m_run = m_new;
      }
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16, logsumexp
    mbar_wait(o_full, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    // tcgen05.ld is warp-collective: every lane issues the loads, only the stores are predicated
    const bool row_ok = qidx < p.Sq;
    bf16* orow = p.out + (long long)b * p.out_bstride + (long long)(row_ok ? qidx : 0) * p.ld_out + h * ATT_D;
#pragma unroll 1
    for (int c = 0; c < ATT_D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + TM_O + c * 32, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[8 * v4 + 0]) * inv_l, __uint_as_float(r[8 * v4 + 1]) * inv_l);
          o.y = pack_bf16(__uint_as_float(r[8 * v4 + 2]) * inv_l, __uint_as_float(r[8 * v4 + 3]) * inv_l);
          o.z = pack_bf16(__uint_as_float(r[8 * v4 + 4]) * inv_l, __uint_as_float(r[8 * v4 + 5]) * inv_l);
          o.w = pack_bf16(__uint_as_float(r[8 * v4 + 6]) * inv_l, __uint_as_float(r[8 * v4 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + v4 * 8) = o;
        }
      }
    }
    if (row_ok && p.lse)
      p.lse[((long long)b * p.H + h) * p.Sq + qidx] = (m_run + log2f(l_run)) * 0.6931471805599453f;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

// 3-D tensor map over a [B, S, width] bf16 activation whose rows are `ld` elements apart; box = 64 x box_rows x 1
int make_att_tmap(CUtensorMap* m, const void* base, int B, int S, long long width, long long ld, int box_rows,
                  long long batch_stride) {
  uint64_t dims[3] = {(uint64_t)width, (uint64_t)S, (uint64_t)B};
  if (batch_stride <= 0) batch_stride = ld * S;      // densely packed [B, S, ld]
  if (B == 1) batch_stride = ld * S;                 // irrelevant for a single batch; keep it a valid multiple of 16 B
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)batch_stride * 2};
  uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tmap(m, base, TMA_BF16, 3, dims, strides, box, TMA_SWIZZLE_128B);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention_fwd(const B200AttentionFwdArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200AttentionFwdArgs, "b200_attention_fwd");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = a->batch, H = a->heads, Sq = a->sq, Sk = a->sk;
  B200_CHECK_ARG(a->head_dim == ATT_D, "b200_attention_fwd: head_dim %d unsupported (only 64)", a->head_dim);
  B200_CHECK_ARG(a->q && a->k && a->v && a->out && B > 0 && H > 0 && Sq > 0 && Sk > 0, "b200_attention_fwd: bad arguments");
  B200_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ld_out % 8 == 0,
                 "b200_attention_fwd: row strides must be multiples of 8 elements");
  B200_CHECK_ARG(a->q_col0 % 8 == 0 && a->k_col0 % 8 == 0 && a->v_col0 % 8 == 0, "b200_attention_fwd: column offsets % 8");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "b200_attention_fwd: out must be 16-byte aligned");
  B200_CHECK_ARG(a->q_bstride % 8 == 0 && a->k_bstride % 8 == 0 && a->v_bstride % 8 == 0 && a->out_bstride % 8 == 0,
                 "b200_attention_fwd: batch strides must be multiples of 8 elements");
  if (a->causal) B200_CHECK_ARG(Sk >= Sq, "b200_attention_fwd: causal attention needs Sk >= Sq");
  B200_CHECK_ARG(a->drop_p >= 0.f && a->drop_p < 1.f, "b200_attention_fwd: dropout p must be in [0, 1)");
  B200_CHECK_ARG(!(a->key_mask != nullptr && a->drop_p > 0.f),
                 "b200_attention_fwd: key_mask is an inference-path feature (no dropout variant)");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_att_tmap(&tq, a->q, B, Sq, a->q_col0 + (long long)H * ATT_D, a->ldq, ATT_BM, a->q_bstride))) return rc;
  if ((rc = make_att_tmap(&tk, a->k, B, Sk, a->k_col0 + (long long)H * ATT_D, a->ldk, ATT_BN, a->k_bstride))) return rc;
  if ((rc = make_att_tmap(&tv, a->v, B, Sk, a->v_col0 + (long long)H * ATT_D, a->ldv, ATT_BN, a->v_bstride))) return rc;
  AttFwdParams p;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out); p.ld_out = a->ld_out; p.lse = a->lse;
  p.out_bstride = a->out_bstride > 0 ? a->out_bstride : (long long)Sq * a->ld_out;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.drop_threshold16 = a->drop_p > 0.f ? (uint32_t)(a->drop_p * 65536.0f + 0.5f) : 0u;
  p.drop_seed = a->drop_seed;
  p.key_mask = a->key_mask;
  p.key_mask_bstride = a->key_mask_bstride > 0 ? a->key_mask_bstride : (long long)Sk;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attention_fwd)");
    configured = true;
  }
  dim3 grid((Sq + ATT_BM - 1) / ATT_BM, H, B);
  cudaError_t le;
  if (p.drop_threshold16 != 0u) le = launch_kernel(attention_fwd_kernel<true, false>, grid, dim3(ATT_THREADS), ATT_SMEM, s, 1, tq, tk, tv, p);
  else if (p.key_mask != nullptr) le = launch_kernel(attention_fwd_kernel<false, true>, grid, dim3(ATT_THREADS), ATT_SMEM, s, 1, tq, tk, tv, p);
  else le = launch_kernel(attention_fwd_kernel<false, false>, grid, dim3(ATT_THREADS), ATT_SMEM, s, 1, tq, tk, tv, p);
  if (le != cudaSuccess) return check_cuda(le, "attention_fwd launch");
  B200_CHECK_LAUNCH("attention_fwd");
  return 0;
}
