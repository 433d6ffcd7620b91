// Flash-style attention backward for sm_100a, head_dim 64, key-major accumulators (no dropout).
#include "attention_bwd_common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------------------------
// Key-major variant (no dropout): the CTA's 128 keys are the TMEM LANES of every accumulator,
//     S^T = K Q^T, dP^T = V dO^T   (lanes = keys, columns = queries)
// so P^T and dS^T, written back to TMEM as bf16 pairs by the thread that owns the key row, ARE the A operands of
//     dV += P^T dO   and   dK += dS^T Q      (tcgen05.mma with A in TMEM: 52 instead of 76 cycles per step)
// and P never touches shared memory; only dS^T is also stored there, as the MN-major A operand of dQ = dS K.
// The query-major kernel above moves ~350 KB through shared memory per 128 x 128 tile pair (operand reads of five
// MMAs + P / dS / dQ staging) against ~2700 cycles of tensor work, i.e. it runs at the shared-memory port's
// 128 B / clk; this layout moves ~270 KB. The freed 32 KB hold a third Q / dO stage.
// TMEM: S^T 0..127 | dP^T 128..255 (dS^T pairs overwrite the first 32 columns of each 64-query half once that half is
// consumed) | dV 256..319 | dK 320..383 | dQ 384..447 | P^T pairs 448..511.
// ------------------------------------------------------------------------------------------------------------------
constexpr int KT_STAGES = 3;
constexpr int KT_SMEM = AB_SMEM + KT_STAGES * 1024;      // + per-stage row statistics: LSE*log2e[128], D[128]
constexpr uint32_t TK_P = 448;
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kt_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_dq, const AttBwdParams p, const AttBwdPadded pp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;                      // KT_STAGES stages
  uint8_t* sdO = smem + (2 + KT_STAGES) * AB_TILE;       // KT_STAGES stages
  uint8_t* sdS = smem + (2 + 2 * KT_STAGES) * AB_TILE;   // 2 x [128 keys][64 queries]  (dS^T)
  uint8_t* sStage = smem + (4 + 2 * KT_STAGES) * AB_TILE;   // 2 x [128 q][32 fp32]
  float* sStat = reinterpret_cast<float*>(smem + (6 + 2 * KT_STAGES) * AB_TILE);   // [stage][lse2 128 | dsum 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (6 + 2 * KT_STAGES) * AB_TILE + KT_STAGES * 1024);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;                 // [KT_STAGES]
  uint64_t* q_empty = q_full + KT_STAGES;      // [KT_STAGES]
  uint64_t* s_full = q_empty + KT_STAGES;
  uint64_t* dp_full = s_full + 1;
  uint64_t* p_ready = s_full + 2;    // 256 arrivals: P^T_t in TMEM, S^T_t consumed
  uint64_t* ds_ready = s_full + 3;   // 256 arrivals: dS^T_t in TMEM and shared memory, dP^T_t consumed
  uint64_t* p_free = s_full + 4;     // dV_t retired: the P^T columns may be overwritten
  uint64_t* dq_full = s_full + 5;    // dK_t, dQ_t retired: dS buffers reusable, dQ_t readable
  uint64_t* dkv_full = s_full + 6;
  uint64_t* dq_drained = s_full + 7; // 128 arrivals (drain warps): dQ_t has left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv_tile = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int k0 = kv_tile * AB_T;
  const int shift = p.Sk - p.Sq;
  const int q_tiles = (p.Sq + AB_T - 1) / AB_T;
  int i_begin = 0;
  if (p.causal) {
    const int first_q = k0 - shift;     // first query index that may see key k0
    i_begin = first_q > 0 ? first_q / AB_T : 0;
  }
  const int n_iter = q_tiles - i_begin;

  if (warp == AB_TMA_WARP && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_do);
    prefetch_tmap(&tmap_dq);
  }
  if (warp == AB_MMA_WARP && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < KT_STAGES; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, AB_COMPUTE_THREADS);
    mbar_init(ds_ready, AB_COMPUTE_THREADS);
    mbar_init(p_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    mbar_init(dq_drained, 128);
    fence_mbar_init();
  }
  if (warp == AB_MMA_WARP) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AB_TMA_WARP) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_3d(sK, &tmap_k, kv_full, p.k_col0 + h * AB_D, k0, b);
      tma_load_3d(sV, &tmap_v, kv_full, p.v_col0 + h * AB_D, k0, b);
    }
    __syncwarp();
    for (int t = 0; t < n_iter; ++t) {
      const int s = t % KT_STAGES;
      const uint32_t ph = (t / KT_STAGES) & 1;
      const int q0 = (i_begin + t) * AB_T;
      mbar_wait(&q_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[s], 2 * AB_TILE + 1024);
        tma_load_3d(sQ + s * AB_TILE, &tmap_q, &q_full[s], p.q_col0 + h * AB_D, q0, b);
        tma_load_3d(sdO + s * AB_TILE, &tmap_do, &q_full[s], p.do_col0 + h * AB_D, q0, b);
        // row statistics of these 128 queries (padded arrays: always 512 bytes, in bounds)
        const long long so = ((long long)b * p.H + h) * pp.Sq_pad + q0;
        bulk_load_1d(sStat + s * 256, pp.lse2_pad + so, 512, &q_full[s]);
        bulk_load_1d(sStat + s * 256 + 128, pp.dsum_pad + so, 512, &q_full[s]);
      }
      __syncwarp();
    }
  } else if (warp == AB_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (n_iter > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);   // S^T, dP^T: A = K / V, B = Q / dO, all K-major
      constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, false, true);     // dV, dK: A in TMEM, B = dO / Q MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, true, true);      // dQ: A = dS^T (MN-major), B = K (MN-major)
      mbar_wait(kv_full, 0);
      tc_fence_after();
      const uint64_t dK_k = make_smem_desc(smem_u32(sK), 16, 1024);
      const uint64_t dV_k = make_smem_desc(smem_u32(sV), 16, 1024);
      const uint64_t dK_mn = make_smem_desc(smem_u32(sK), 16384, 1024);
      const uint64_t dS_mn = make_smem_desc(smem_u32(sdS), 16384, 1024);     // rows = keys (K index), 64-query chunks 16 KB apart

      auto issue_s = [&](int t) {
        const uint64_t dQ_k = make_smem_desc(smem_u32(sQ + (t % KT_STAGES) * AB_TILE), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + TB_S, dK_k + (uint64_t)(2 * k), dQ_k + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int t) {
        const uint64_t dO_k = make_smem_desc(smem_u32(sdO + (t % KT_STAGES) * AB_TILE), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + TB_DP, dV_k + (uint64_t)(2 * k), dO_k + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
      };

      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t % KT_STAGES;
        const uint64_t dO_mn = make_smem_desc(smem_u32(sdO + s * AB_TILE), 16384, 1024);
        const uint64_t dQ_mn = make_smem_desc(smem_u32(sQ + s * AB_TILE), 16384, 1024);
        // ---- P^T_t ready: S^T_{t+1}, then dV += P^T_t dO_t
        mbar_wait(p_ready, t & 1);
        tc_fence_after();
        if (t + 1 < n_iter) {
          mbar_wait(&q_full[(t + 1) % KT_STAGES], ((t + 1) / KT_STAGES) & 1);
          tc_fence_after();
          issue_s(t + 1);
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts(tmem_base + TB_DV, tmem_base + TK_P + 8 * k, dO_mn + (uint64_t)(128 * k), idesc_t,
                    (t > 0 || k > 0) ? 1u : 0u);
          umma_commit(p_free);
        }
        __syncwarp();
        // ---- dS^T_t ready: dK += dS^T_t Q_t, dQ_t = dS_t K, then dP^T_{t+1} (which overwrites dS^T_t: same queue, in order)
        mbar_wait(ds_ready, t & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts(tmem_base + TB_DK, tmem_base + TB_DP + (k < 4 ? 8 * k : 64 + 8 * (k - 4)), dQ_mn + (uint64_t)(128 * k),
                    idesc_t, (t > 0 || k > 0) ? 1u : 0u);
          umma_commit(&q_empty[s]);      // Q_t / dO_t are dead once dK_t retires
        }
        __syncwarp();
        if (t > 0) {                     // dQ_{t-1} must have left its TMEM columns
          mbar_wait(dq_drained, (t - 1) & 1);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ss(tmem_base + TB_DQ, dS_mn + (uint64_t)(128 * k), dK_mn + (uint64_t)(128 * k), idesc_q, k > 0 ? 1u : 0u);
          umma_commit(dq_full);
        }
        __syncwarp();
        if (t + 1 < n_iter) issue_dp(t + 1);
      }
      if (elect_one()) umma_commit(dkv_full);
      __syncwarp();
    }
  } else if (warp >= AB_DRAIN_WARP0) {
    // ===================== dQ drain warps (as in the query-major kernel) =====================
    const int dt = threadIdx.x - AB_DRAIN_WARP0 * 32;      // 0..127 = query row
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int sw = dt & 7;
    for (int t = 0; t < n_iter; ++t) {
      mbar_wait(dq_full, t & 1);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld_32x32_at<0>(lane_addr + TB_DQ, r);
      tmem_ld_32x32_at<32>(lane_addr + TB_DQ + 32, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_drained);
      if (dt == 0) tma_wait_group_read<0>();
      named_bar_sync(1, 128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t rowp = smem_u32(sStage) + c * AB_TILE + dt * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(rowp + ((j ^ sw) << 4), r[32 * c + 4 * j], r[32 * c + 4 * j + 1], r[32 * c + 4 * j + 2], r[32 * c + 4 * j + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (dt == 0) {
        tma_reduce_add_3d(&tmap_dq, sStage, h * AB_D, (i_begin + t) * AB_T, b);
        tma_reduce_add_3d(&tmap_dq, sStage + AB_TILE, h * AB_D + 32, (i_begin + t) * AB_T, b);
        tma_commit_group();
      }
    }
    if (dt == 0) tma_wait_group<0>();
  } else {
    // ===================== compute warps: thread = (key row, 64-query half) =====================
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;     // TMEM lane = key
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int sw = row & 7;
    const int kidx = k0 + row;
    const bool key_ok = kidx < p.Sk;
    const uint32_t stat_s = smem_u32(sStat) + half * 64 * 4;
    const uint32_t drow = smem_u32(sdS) + half * AB_TILE + row * 128;
    const f32x2 scale2 = f2_splat(p.scale);
    for (int t = 0; t < n_iter; ++t) {
      const int q0 = (i_begin + t) * AB_T;
      const int qb = q0 + half * 64;            // first query of this thread's columns
      const bool need_mask = (k0 + AB_T > p.Sk) || (p.causal && (k0 + AB_T - 1 > q0 + shift));
      // ---------------- stage A: P^T_t = exp2(S^T_t * c - LSE[q]) ----------------
      uint32_t pk[32];                          // P^T as bf16 pairs (queries 2j, 2j+1)
      mbar_wait(s_full, t & 1);            // (S^T_t complete implies Q_t and the statistics of this stage have landed;
      mbar_wait(&q_full[t % KT_STAGES], (t / KT_STAGES) & 1);     //  the explicit wait makes the TMA writes visible here)
      tc_fence_after();
      const uint32_t lse_s = stat_s + (t % KT_STAGES) * 1024, dsum_s = lse_s + 512;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32];
        tmem_ld_32x32(lane_addr + TB_S + half * 64 + c * 32, rs);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 lu = lds128(lse_s + (c * 32 + g * 4) * 4);      // same address in all lanes: broadcast
          const float4 l4 = make_float4(__uint_as_float(lu.x), __uint_as_float(lu.y), __uint_as_float(lu.z), __uint_as_float(lu.w));
          float v0 = ex2_approx(fmaf(__uint_as_float(rs[4 * g]), p.scale_log2, -l4.x));
          float v1 = ex2_approx(fmaf(__uint_as_float(rs[4 * g + 1]), p.scale_log2, -l4.y));
          float v2 = ex2_approx(fmaf(__uint_as_float(rs[4 * g + 2]), p.scale_log2, -l4.z));
          float v3 = ex2_approx(fmaf(__uint_as_float(rs[4 * g + 3]), p.scale_log2, -l4.w));
          if (need_mask) {
            const int qi = qb + c * 32 + g * 4;         // key visible to query qi  <=>  kidx <= qi + shift (causal), kidx < Sk
            const int lim = p.causal ? kidx - shift : -0x40000000;     // masked  <=>  qi < lim
            if (!key_ok || qi < lim) v0 = 0.f;
            if (!key_ok || qi + 1 < lim) v1 = 0.f;
            if (!key_ok || qi + 2 < lim) v2 = 0.f;
            if (!key_ok || qi + 3 < lim) v3 = 0.f;
          }
          pk[c * 16 + 2 * g] = pack_bf16(v0, v1);
          pk[c * 16 + 2 * g + 1] = pack_bf16(v2, v3);
        }
      }
      if (t > 0) mbar_wait(p_free, (t - 1) & 1);      // dV_{t-1} no longer reads the P^T columns
      tc_fence_after();
      {
        uint32_t w16[16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int e = 0; e < 16; ++e) w16[e] = pk[c * 16 + e];
          tmem_st_32x16(lane_addr + TK_P + half * 32 + c * 16, w16);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);

      // ---------------- stage B: dS^T_t = P^T_t * (dP^T_t - D[q]) * scale ----------------
      if (t > 0) mbar_wait(dq_full, (t - 1) & 1);     // dQ_{t-1} no longer reads the dS^T tile in shared memory
      mbar_wait(dp_full, t & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rp[32];
        tmem_ld_32x32(lane_addr + TB_DP + half * 64 + c * 32, rp);
        tmem_ld_wait();
        uint32_t dw[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 du = lds128(dsum_s + (c * 32 + g * 4) * 4);
          const float4 d4 = make_float4(__uint_as_float(du.x), __uint_as_float(du.y), __uint_as_float(du.z), __uint_as_float(du.w));
          const uint32_t w0 = pk[c * 16 + 2 * g], w1 = pk[c * 16 + 2 * g + 1];
          float a0, a1, a2, a3;
          f2_unpack(f2_mul(f2_add(f2_pack(__uint_as_float(rp[4 * g]), __uint_as_float(rp[4 * g + 1])), f2_pack(-d4.x, -d4.y)),
                           f2_mul(f2_pack(bf16_lo(w0), bf16_hi(w0)), scale2)), a0, a1);
          f2_unpack(f2_mul(f2_add(f2_pack(__uint_as_float(rp[4 * g + 2]), __uint_as_float(rp[4 * g + 3])), f2_pack(-d4.z, -d4.w)),
                           f2_mul(f2_pack(bf16_lo(w1), bf16_hi(w1)), scale2)), a2, a3);
          dw[2 * g] = pack_bf16(a0, a1);
          dw[2 * g + 1] = pack_bf16(a2, a3);
        }
        // TMEM: A operand of dK += dS^T Q (overwrites the first half of the dP^T columns this thread has already read);
        // shared memory: [key row][64 queries], the MN-major A operand of dQ = dS K
        tmem_st_32x16(lane_addr + TB_DP + half * 64 + c * 16, dw);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) sts128(drow + (((c * 4 + jj) ^ sw) << 4), dw[4 * jj], dw[4 * jj + 1], dw[4 * jj + 2], dw[4 * jj + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
    }
    if (n_iter > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    // ---- final: dK, dV (lanes = keys; this thread owns d columns [32*half, 32*half+32)) -> bf16 -> global
    const bool k_ok = key_ok;
    bf16* dkrow = p.dk + ((long long)b * p.Sk + (k_ok ? kidx : 0)) * p.ld_dk + p.dk_col0 + h * AB_D + half * 32;
    bf16* dvrow = p.dv + ((long long)b * p.Sk + (k_ok ? kidx : 0)) * p.ld_dv + p.dv_col0 + h * AB_D + half * 32;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bf16* orow = which == 0 ? dvrow : dkrow;
      uint32_t r[32];
      if (n_iter > 0) {
        tmem_ld_32x32(lane_addr + (which == 0 ? TB_DV : TB_DK) + half * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0u;
      }
      if (k_ok) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[8 * v4 + 0]), __uint_as_float(r[8 * v4 + 1]));
          o.y = pack_bf16(__uint_as_float(r[8 * v4 + 2]), __uint_as_float(r[8 * v4 + 3]));
          o.z = pack_bf16(__uint_as_float(r[8 * v4 + 4]), __uint_as_float(r[8 * v4 + 5]));
          o.w = pack_bf16(__uint_as_float(r[8 * v4 + 6]), __uint_as_float(r[8 * v4 + 7]));
          *reinterpret_cast<uint4*>(orow + v4 * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AB_MMA_WARP) tmem_dealloc<512>(tmem_base);
}


int launch_attention_bwd_kt(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                            const CUtensorMap& tdq, const AttBwdParams& p, const AttBwdPadded& pp, dim3 grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attention_bwd_kt)");
    configured = true;
  }
  attention_bwd_kt_kernel<<<grid, AB_THREADS, KT_SMEM, s>>>(tq, tk, tv, tdo, tdq, p, pp);
  B200_CHECK_LAUNCH("attention_bwd_kt");
  return 0;
}

}  // namespace b200
