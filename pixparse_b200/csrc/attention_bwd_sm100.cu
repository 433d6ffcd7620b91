// Flash-style attention backward for sm_100a, head_dim 64 (recomputes P from Q, K and the saved logsumexp).
//
// One CTA per (128-key tile, head, batch) loops over the 128-query tiles that can see it:
//     S  = Q K^T            tcgen05.mma -> TMEM           dP = dO V^T          tcgen05.mma -> TMEM
//     P  = exp(S*scale - LSE) ; dS = P * (dP - D) * scale   (one thread per query row, bf16 -> shared memory)
//     dV += P^T dO          dK += dS^T Q                  (accumulate in TMEM over the whole loop)
//     dQ  = dS K            -> TMEM -> shared -> TMA reduce-add into an fp32 dQ accumulator in HBM
// P / dS are written once to shared memory in the 128-byte-swizzled layout that serves BOTH as the MN-major A
// operand of the dV / dK products and as the K-major A operand of the dQ product; Q, K, V, dO tiles are used
// exactly as TMA delivers them (K-major for S / dP, MN-major for dV / dK / dQ).
//
// Replaces the SDPA backward kernels autograd reaches from timm Attention and BartAttention.
#include <type_traits>

#include "attention_bwd_common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

int make_att_tmap(CUtensorMap* m, const void* base, int B, int S, long long width, long long ld, int box_rows,
                  long long batch_stride);


// Pipeline per query tile t (tensor pipe on the left, the 8 compute warps on the right run concurrently):
//     S_t, dP_t ready ............ stage A: P_t = exp2(S_t*c - LSE)  -> smem, keeps P_t in registers
//     dV += P_t^T dO_t ; S_{t+1} .. (drain dQ_{t-1}: TMEM -> smem -> TMA reduce-add) ; stage B: dS_t = P_t*(dP_t - D)*scale -> smem
//     dK += dS_t^T Q_t ; dQ_t = dS_t K ; dP_{t+1} ....... stage A of tile t+1 ...
// so the MMAs of one stage always run under the exp / multiply work of the other stage.
template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_dq, const AttBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;      // 2 stages
  uint8_t* sdO = smem + 4 * AB_TILE;     // 2 stages
  uint8_t* sP = smem + 6 * AB_TILE;      // 2 x [128 q][64 keys]
  uint8_t* sdS = smem + 8 * AB_TILE;     // 2 x [128 q][64 keys]
  uint8_t* sStage = smem + 10 * AB_TILE; // 2 x [128 q][32 fp32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * AB_TILE);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;    // [2]
  uint64_t* q_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* dp_full = bars + 6;
  uint64_t* p_ready = bars + 7;   // 256 arrivals: P_t in smem, S_t consumed
  uint64_t* ds_ready = bars + 8;  // 256 arrivals: dS_t in smem, dP_t consumed
  uint64_t* p_free = bars + 9;    // dV_t retired: P buffer reusable
  uint64_t* dq_full = bars + 10;  // dK_t, dQ_t retired: dS buffer reusable, dQ_t readable
  uint64_t* dkv_full = bars + 11;
  uint64_t* dq_drained = bars + 12;   // 128 arrivals (drain warps): dQ_t has left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  pdl_launch_dependents();
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
    for (int i = 0; i < 2; ++i) {
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
  pdl_wait();

  if (warp == AB_TMA_WARP) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_3d(sK, &tmap_k, kv_full, p.k_col0 + h * AB_D, k0, b);
      tma_load_3d(sV, &tmap_v, kv_full, p.v_col0 + h * AB_D, k0, b);
    }
    __syncwarp();
    for (int t = 0; t < n_iter; ++t) {
      const int s = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      const int q0 = (i_begin + t) * AB_T;
      mbar_wait(&q_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[s], 2 * AB_TILE);
        tma_load_3d(sQ + s * AB_TILE, &tmap_q, &q_full[s], p.q_col0 + h * AB_D, q0, b);
        tma_load_3d(sdO + s * AB_TILE, &tmap_do, &q_full[s], p.do_col0 + h * AB_D, q0, b);
      }
      __syncwarp();
    }
  } else if (warp == AB_MMA_WARP) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    if (n_iter > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);   // S, dP
      constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, true, true);      // dV, dK (A = P^T / dS^T)
      constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);     // dQ
      mbar_wait(kv_full, 0);
      tc_fence_after();
      const uint64_t dK_k = make_smem_desc(smem_u32(sK), 16, 1024);        // K as K-major B of S
      const uint64_t dV_k = make_smem_desc(smem_u32(sV), 16, 1024);        // V as K-major B of dP
      const uint64_t dK_mn = make_smem_desc(smem_u32(sK), 16384, 1024);    // K as MN-major B of dQ
      const uint64_t dP_mn = make_smem_desc(smem_u32(sP), 16384, 1024);    // P  as MN-major A (M = keys)
      const uint64_t dS_mn = make_smem_desc(smem_u32(sdS), 16384, 1024);   // dS as MN-major A (M = keys)

      auto issue_s = [&](int t) {
        const uint64_t dQ_k = make_smem_desc(smem_u32(sQ + (t & 1) * AB_TILE), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + TB_S, dQ_k + (uint64_t)(2 * k), dK_k + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int t) {
        const uint64_t dO_k = make_smem_desc(smem_u32(sdO + (t & 1) * AB_TILE), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + TB_DP, dO_k + (uint64_t)(2 * k), dV_k + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
      };

      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        const uint64_t dO_mn = make_smem_desc(smem_u32(sdO + s * AB_TILE), 16384, 1024);
        const uint64_t dQ_mn = make_smem_desc(smem_u32(sQ + s * AB_TILE), 16384, 1024);
        // ---- P_t ready: S_{t+1}, then dV += P_t^T dO_t
#ifdef AB_TRACE
        const bool trace_on = p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif
        AB_STAMP(8);
        mbar_wait(p_ready, t & 1);
        AB_STAMP(9);
        tc_fence_after();
        if (t + 1 < n_iter) {      // S_{t+1} first: the compute warps start the next tile with it
          mbar_wait(&q_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(t + 1);
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ss(tmem_base + TB_DV, dP_mn + (uint64_t)(128 * k), dO_mn + (uint64_t)(128 * k), idesc_t,
                    (t > 0 || k > 0) ? 1u : 0u);
          umma_commit(p_free);
        }
        __syncwarp();
        // ---- dS_t ready: dK += dS_t^T Q_t, dQ_t = dS_t K, then dP_{t+1}
        AB_STAMP(10);
        mbar_wait(ds_ready, t & 1);
        AB_STAMP(11);
        tc_fence_after();
        // dS read K-major for dQ: 64-key chunk = k / 4, 32 bytes per step inside the 128-byte row
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ss(tmem_base + TB_DK, dS_mn + (uint64_t)(128 * k), dQ_mn + (uint64_t)(128 * k), idesc_t,
                    (t > 0 || k > 0) ? 1u : 0u);
          umma_commit(&q_empty[s]);      // Q_t / dO_t are dead once dK_t retires: the reload for tile t+2 starts now
        }
        __syncwarp();
        if (t > 0) {                     // dQ_{t-1} must have left its TMEM columns
          mbar_wait(dq_drained, (t - 1) & 1);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts(tmem_base + TB_DQ, tmem_base + TB_DS + 8 * k, dK_mn + (uint64_t)(128 * k), idesc_q, k > 0 ? 1u : 0u);
          umma_commit(dq_full);
        }
        __syncwarp();
        if (t + 1 < n_iter) issue_dp(t + 1);
        AB_STAMP(12);
      }
      if (elect_one()) umma_commit(dkv_full);
      __syncwarp();
    }
  } else if (warp >= AB_DRAIN_WARP0) {
    // ===================== dQ drain warps =====================
    // dQ_t (128 queries x 64): TMEM -> two swizzled fp32 staging tiles -> TMA reduce-add into the fp32 accumulator in HBM.
    // Separate warps because this ~750-cycle copy used to sit between the two stages of the compute warps.
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
#pragma unroll
      for (int e = 0; e < 64; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * p.scale);      // dQ = (dS / scale) K * scale
      if (dt == 0) tma_wait_group_read<0>();     // the previous reduce has finished reading the staging tiles
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
    // ===================== compute warps =====================
    const int half = warp >> 2;                 // key-column half of S / dP, d-column half of dQ / dK / dV
    const int row = (warp & 3) * 32 + lane;     // TMEM lane = query row (S, dP, dQ) or key row (dK, dV)
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int sw = row & 7;
    const float kLog2e = 1.4426950408889634f;

    // per-row statistics of the NEXT query tile are fetched one iteration ahead: a dependent global load at the loop
    // top costs ~900 cycles per tile (measured with the AB_TRACE build), a fifth of the whole iteration
    auto load_stats = [&](int t, float& lse_raw, float& dsum_raw) {
      const int qi = (i_begin + t) * AB_T + row;
      const bool ok = t < n_iter && qi < p.Sq;
      const long long idx = ((long long)b * p.H + h) * p.Sq + (ok ? qi : 0);
      lse_raw = ok ? __ldg(p.lse + idx) : INFINITY;      // +inf -> P = 0 for padded query rows
      dsum_raw = ok ? __ldg(p.dsum + idx) : 0.f;
    };
    float lse_next, dsum_next;
    load_stats(0, lse_next, dsum_next);
    for (int t = 0; t < n_iter; ++t) {
      const int q0 = (i_begin + t) * AB_T;
      const int qidx = q0 + row;
      const float lse2 = lse_next * kLog2e;
      const float dsum = dsum_next;
      load_stats(t + 1, lse_next, dsum_next);
      int kmax = p.Sk - 1;
      if (p.causal) kmax = min(kmax, qidx + shift);
      const bool need_mask = (k0 + AB_T > p.Sk) || (p.causal && (k0 + AB_T - 1 > q0 + shift));
      const int kbase = k0 + half * 64;

      // ---------------- stage A: P_t ----------------
      // P is kept as packed bf16 pairs (what the dV MMA sees); under dropout the sign bit of each half carries
      // "dropped in the forward pass" (P >= 0, so the bit is free) and pd holds the dropped, rescaled copy for dV.
      uint32_t pk[32];
      uint32_t pd[DROP ? 32 : 1];
      // dropout pair index of (query row, key k) = drop_row + k / 2, exactly as in the forward kernel
      const uint32_t drop_row = (uint32_t)((((long long)b * p.H + h) * p.Sq + qidx) * ((p.Sk + 1) >> 1)) +
                                (uint32_t)(kbase >> 1);
      const float drop_sc = dropout_scale(p.drop_threshold16);
#ifdef AB_TRACE
      const bool trace_on = p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
#endif
      AB_STAMP(0);
      mbar_wait(s_full, t & 1);
      AB_STAMP(1);
      tc_fence_after();
      // (two instantiations: interior tiles -- all but the ragged last key tile and the causal diagonal -- carry no masking
      //  instructions at all; as a runtime flag the mask cost every tile 64 ISETP + 64 FSEL, a fifth of this loop's
      //  instructions, on warps that are issue / latency bound at two per scheduler)
      auto stage_a = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t rs[32];
          tmem_ld_32x32(lane_addr + TB_S + half * 64 + c * 32, rs);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            // All eight compute warps run this stage at the same time (they wait for the same S tile), two per scheduler:
            // 2 x 64 MUFU.EX2 x 8 cycles = 1024 cycles of XU pipe per tile while the FMA pipe idles (ncu: this stage was
            // 41 % of the compute warps' time). Every other element pair goes through the FMA-pipe polynomial instead.
            const float a0 = fmaf(__uint_as_float(rs[e]), p.scale_log2, -lse2);
            const float a1 = fmaf(__uint_as_float(rs[e + 1]), p.scale_log2, -lse2);
            float v0, v1;
            // (not under dropout: that variant is issue-bound on its mask hashing and sits at the 128-register limit)
            if (!DROP && ((e >> 1) & 1)) {
              ex2_poly2(a0, a1, v0, v1);
            } else {
              v0 = ex2_approx(a0);
              v1 = ex2_approx(a1);
            }
            if (MASKED) {
              if (kbase + c * 32 + e > kmax) v0 = 0.f;
              if (kbase + c * 32 + e + 1 > kmax) v1 = 0.f;
            }
            uint32_t w = pack_bf16(v0, v1);
            if (DROP) {
              const uint32_t hsh = dropout_hash(p.drop_seed, drop_row + (uint32_t)((c * 32 + e) >> 1));
              const bool keep0 = (hsh & 0xFFFFu) >= p.drop_threshold16, keep1 = (hsh >> 16) >= p.drop_threshold16;
              pd[DROP ? c * 16 + (e >> 1) : 0] = pack_bf16(keep0 ? v0 * drop_sc : 0.f, keep1 ? v1 * drop_sc : 0.f);
              w |= (keep0 ? 0u : 0x8000u) | (keep1 ? 0u : 0x80000000u);
            }
            pk[c * 16 + (e >> 1)] = w;
          }
        }
      };
      if (need_mask) stage_a(std::true_type{}); else stage_a(std::false_type{});
      AB_STAMP(2);
      if (t > 0) mbar_wait(p_free, (t - 1) & 1);      // dV_{t-1} no longer reads the P buffer
      {
        const uint32_t prow = smem_u32(sP) + half * AB_TILE + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (DROP)
            sts128(prow + ((j ^ sw) << 4), pd[DROP ? 4 * j : 0], pd[DROP ? 4 * j + 1 : 0], pd[DROP ? 4 * j + 2 : 0],
                   pd[DROP ? 4 * j + 3 : 0]);
          else
            sts128(prow + ((j ^ sw) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
      fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_ready);
      AB_STAMP(3);

      if (t > 0) mbar_wait(dq_full, (t - 1) & 1);     // dK_{t-1} / dQ_{t-1} no longer read the dS buffer
      AB_STAMP(7);

      // ---------------- stage B: dS_t ----------------
      // dS is produced WITHOUT the softmax scale: dQ = (dS K) * scale is scaled by the (mostly idle) drain warps and
      // dK = (dS^T Q) * scale once in the epilogue -- 32 FMUL2 fewer per thread and tile here
      const f32x2 ndsum2 = f2_splat(-dsum);
      AB_STAMP(4);
      mbar_wait(dp_full, t & 1);
      AB_STAMP(5);
      tc_fence_after();
      {
        const uint32_t drow = smem_u32(sdS) + half * AB_TILE + row * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t rp[32];
          tmem_ld_32x32(lane_addr + TB_DP + half * 64 + c * 32, rp);
          tmem_ld_wait();
          float dsv[32];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {      // dS / scale = P * (mask * dP - D), on packed fp32 pairs
            uint32_t w = pk[c * 16 + (e >> 1)];
            f32x2 g = f2_pack(__uint_as_float(rp[e]), __uint_as_float(rp[e + 1]));
            if (DROP) {
              g = f2_mul(g, f2_pack((w & 0x8000u) ? 0.f : drop_sc, (w & 0x80000000u) ? 0.f : drop_sc));
              w &= 0x7FFF7FFFu;
            }
            f2_unpack(f2_mul(f2_add(g, ndsum2), f2_pack(bf16_lo(w), bf16_hi(w))), dsv[e], dsv[e + 1]);
          }
          uint32_t dw[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) dw[e] = pack_bf16(dsv[2 * e], dsv[2 * e + 1]);
          // dS goes to shared memory (MN-major A operand of dK += dS^T Q) AND to TMEM: there it is already laid out as the
          // A operand of dQ = dS K (lane = query, columns = keys), and an MMA with A in TMEM costs 52 instead of 76 cycles
          tmem_st_32x16(lane_addr + TB_DS + half * 32 + c * 16, dw);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int piece = (c * 4 + jj) ^ sw;
            sts128(drow + (piece << 4), dw[4 * jj], dw[4 * jj + 1], dw[4 * jj + 2], dw[4 * jj + 3]);
          }
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      AB_STAMP(6);
    }
    if (n_iter > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    // ---- final: dK, dV (rows = keys of this tile; this thread owns d columns [32*half, 32*half+32)) -> bf16 -> global
    const int kidx = k0 + row;
    const bool k_ok = kidx < p.Sk;
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
        const float osc = which == 0 ? 1.0f : p.scale;      // dK accumulated dS^T Q without the softmax scale
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[8 * v4 + 0]) * osc, __uint_as_float(r[8 * v4 + 1]) * osc);
          o.y = pack_bf16(__uint_as_float(r[8 * v4 + 2]) * osc, __uint_as_float(r[8 * v4 + 3]) * osc);
          o.z = pack_bf16(__uint_as_float(r[8 * v4 + 4]) * osc, __uint_as_float(r[8 * v4 + 5]) * osc);
          o.w = pack_bf16(__uint_as_float(r[8 * v4 + 6]) * osc, __uint_as_float(r[8 * v4 + 7]) * osc);
          *reinterpret_cast<uint4*>(orow + v4 * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AB_MMA_WARP) tmem_dealloc<512>(tmem_base);
}

// D[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (8 threads x 16 bytes per (b,q,h) row of 64 elements; coalesced)
// Also writes the padded per-(b, h) copies the key-major kernel reads with unpredicated 16-byte loads:
// lse2_pad = LSE * log2(e) (+inf beyond Sq -> P = 0 for padded queries), dsum_pad (0 beyond Sq).
__global__ void attention_bwd_pad_kernel(float* __restrict__ lse2_pad, float* __restrict__ dsum_pad, int BH, int Sq,
                                         int Sq_pad) {
  pdl_launch_dependents();
  pdl_wait();
  const int npad = Sq_pad - Sq;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BH * npad; i += gridDim.x * blockDim.x) {
    const int bh = i / npad, q = Sq + i % npad;
    lse2_pad[(long long)bh * Sq_pad + q] = INFINITY;
    dsum_pad[(long long)bh * Sq_pad + q] = 0.f;
  }
}
__global__ void attention_bwd_prep_kernel(const bf16* __restrict__ o, long long ld_o, const bf16* __restrict__ d_o,
                                          long long ld_do, int do_col0, float* __restrict__ dsum, int B, int H,
                                          int Sq, const float* __restrict__ lse, float* __restrict__ lse2_pad,
                                          float* __restrict__ dsum_pad, int Sq_pad, float* __restrict__ dq32) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * Sq * H * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {      // total and the stride are multiples of 8: groups stay intact
    const int part = (int)(idx & 7);
    const long long w = idx >> 3;
    const int hh = (int)(w % H);
    const long long bq = w / H;     // b * Sq + q
    // zero this (row, head)'s 64 floats of the fp32 dQ accumulator on the way (replaces a separate 50-100 MB memset)
    {
      float4* z = reinterpret_cast<float4*>(dq32 + (bq * H + hh) * 64 + part * 8);
      z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint4 a = *reinterpret_cast<const uint4*>(o + bq * ld_o + hh * 64 + part * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + bq * ld_do + do_col0 + hh * 64 + part * 8);
    float s = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
              bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
              bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    const uint32_t gmask = 0xFFu << (threadIdx.x & 24);   // the 8 lanes of this row (a warp's tail may be idle)
    s += __shfl_xor_sync(gmask, s, 1);
    s += __shfl_xor_sync(gmask, s, 2);
    s += __shfl_xor_sync(gmask, s, 4);
    if (part == 0) {
      const int bb = (int)(bq / Sq), q = (int)(bq % Sq);
      dsum[((long long)bb * H + hh) * Sq + q] = s;
      const long long ip = ((long long)bb * H + hh) * Sq_pad + q;
      dsum_pad[ip] = s;
      lse2_pad[ip] = lse[((long long)bb * H + hh) * Sq + q] * 1.4426950408889634f;
    }
  }
}

// dq bf16 [rows, ld_dq] (columns dq_col0 ..) = bf16(dq32 [rows, width])
__global__ void attention_dq_convert_kernel(const float* __restrict__ dq32, bf16* __restrict__ dq, long long ld_dq,
                                            int dq_col0, long long rows, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const int w4 = width / 4;
  const long long total = rows * w4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / w4;
    const int c = (int)(i % w4);
    const float4 v = reinterpret_cast<const float4*>(dq32)[i];
    *reinterpret_cast<uint2*>(dq + r * ld_dq + dq_col0 + c * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

}  // namespace b200

using namespace b200;

#ifdef AB_TRACE
long long* g_ab_trace = nullptr;
extern "C" int b200_debug_set_trace(long long* ptr) { g_ab_trace = ptr; return 0; }
#endif

// kernel selection for the no-dropout case: 1 = query-major (default), 0 = key-major (attention_bwd_kt_sm100.cu)
static int g_ab_query_major = 1;      // measured in the full step: query-major 0.3 % ahead of key-major, standalone on par
extern "C" int b200_debug_attention_bwd_query_major(int on) {
  g_ab_query_major = on;
  return 0;
}

extern "C" long long b200_attention_bwd_workspace_bytes(int B, int H, int Sq) {
  // fp32 dQ accumulator [B*Sq, H*64], D [B, H, Sq], then the padded LSE*log2e and D [B, H, Sq_pad] (each 16-byte aligned)
  const long long sq_pad = (Sq + 127) / 128 * 128;
  const long long head = (((long long)B * Sq * H * 64 + (long long)B * H * Sq) + 3) / 4 * 4;
  return (head + 2 * (long long)B * H * sq_pad) * 4;
}

extern "C" int b200_attention_bwd(const B200AttentionBwdArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200AttentionBwdArgs, "b200_attention_bwd");
  const void *q = a->q, *k = a->k, *v = a->v, *o = a->o, *d_o = a->d_o;
  const long long ldq = a->ldq, ldk = a->ldk, ldv = a->ldv, ld_o = a->ld_o, ld_do = a->ld_do;
  const long long ld_dq = a->ld_dq, ld_dk = a->ld_dk, ld_dv = a->ld_dv;
  const int q_col0 = a->q_col0, k_col0 = a->k_col0, v_col0 = a->v_col0, do_col0 = a->do_col0;
  const int dq_col0 = a->dq_col0, dk_col0 = a->dk_col0, dv_col0 = a->dv_col0;
  const float* lse = a->lse;
  void *dq = a->dq, *dk = a->dk, *dv = a->dv, *workspace = a->workspace;
  const int B = a->batch, H = a->heads, Sq = a->sq, Sk = a->sk, head_dim = a->head_dim, causal = a->causal;
  const float scale = a->scale, drop_p = a->drop_p;
  const unsigned int drop_seed = a->drop_seed;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(head_dim == AB_D, "b200_attention_bwd: head_dim %d unsupported (only 64)", head_dim);
  B200_CHECK_ARG(q && k && v && o && d_o && lse && dq && dk && dv && workspace, "b200_attention_bwd: null argument");
  B200_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ld_o % 8 == 0 && ld_do % 8 == 0 && ld_dq % 8 == 0 &&
                     ld_dk % 8 == 0 && ld_dv % 8 == 0,
                 "b200_attention_bwd: row strides must be multiples of 8 elements");
  B200_CHECK_ARG(q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0 && do_col0 % 8 == 0 && dq_col0 % 8 == 0 &&
                     dk_col0 % 8 == 0 && dv_col0 % 8 == 0,
                 "b200_attention_bwd: column offsets must be multiples of 8");
  if (causal) B200_CHECK_ARG(Sk >= Sq, "b200_attention_bwd: causal attention needs Sk >= Sq");
  const int W = H * AB_D;
  float* dq32 = reinterpret_cast<float*>(workspace);
  float* dsum = dq32 + (long long)B * Sq * W;
  const int Sq_pad = (Sq + AB_T - 1) / AB_T * AB_T;
  float* lse2_pad = dq32 + (((long long)B * Sq * W + (long long)B * H * Sq) + 3) / 4 * 4;
  float* dsum_pad = lse2_pad + (long long)B * H * Sq_pad;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "b200_attention_bwd: workspace must be 16-byte aligned");
  cudaError_t e = cudaSuccess;      // (dq32 is zeroed by the prep kernel)

  {
    const long long items = (long long)B * Sq * H * 8;
    long long blocks = (items + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (Sq_pad > Sq) {
      e = launch_kernel(attention_bwd_pad_kernel, dim3((B * H * (Sq_pad - Sq) + 255) / 256), dim3(256), 0, s, 1, lse2_pad,
                        dsum_pad, B * H, Sq, Sq_pad);
      if (e != cudaSuccess) return check_cuda(e, "attention_bwd_pad launch");
    }
    e = launch_kernel(attention_bwd_prep_kernel, dim3((unsigned)blocks), dim3(256), 0, s, 1, reinterpret_cast<const bf16*>(o), ld_o,
                      reinterpret_cast<const bf16*>(d_o), ld_do, do_col0, dsum, B, H, Sq, lse, lse2_pad, dsum_pad, Sq_pad, dq32);
    if (e != cudaSuccess) return check_cuda(e, "attention_bwd_prep launch");
    B200_CHECK_LAUNCH("attention_bwd_prep");
  }

  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = make_att_tmap(&tq, q, B, Sq, q_col0 + (long long)W, ldq, 128, 0))) return rc;
  if ((rc = make_att_tmap(&tk, k, B, Sk, k_col0 + (long long)W, ldk, 128, 0))) return rc;
  if ((rc = make_att_tmap(&tv, v, B, Sk, v_col0 + (long long)W, ldv, 128, 0))) return rc;
  if ((rc = make_att_tmap(&tdo, d_o, B, Sq, do_col0 + (long long)W, ld_do, 128, 0))) return rc;
  {
    // fp32 dQ accumulator viewed as [B][Sq][W]; box = 32 floats x 128 rows (rows beyond Sq are clipped)
    uint64_t dims[3] = {(uint64_t)W, (uint64_t)Sq, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)W * 4, (uint64_t)W * 4 * (uint64_t)Sq};
    uint32_t box[3] = {32, 128, 1};
    if ((rc = make_tmap(&tdq, dq32, TMA_F32, 3, dims, strides, box, TMA_SWIZZLE_128B))) return rc;
  }
  AttBwdParams p;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse; p.dsum = dsum;
  AttBwdPadded pp;
  pp.lse2_pad = lse2_pad; pp.dsum_pad = dsum_pad; pp.Sq_pad = Sq_pad;
  p.dk = reinterpret_cast<bf16*>(dk); p.ld_dk = ld_dk; p.dk_col0 = dk_col0;
  p.dv = reinterpret_cast<bf16*>(dv); p.ld_dv = ld_dv; p.dv_col0 = dv_col0;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.do_col0 = do_col0;
  B200_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "b200_attention_bwd: dropout p must be in [0, 1)");
  p.drop_threshold16 = drop_p > 0.f ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;
  p.drop_seed = drop_seed;
  p.trace = nullptr;
#ifdef AB_TRACE
  { extern long long* g_ab_trace; p.trace = g_ab_trace; }
#endif
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);

    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attention_bwd)");
    configured = true;
  }
  dim3 grid((Sk + AB_T - 1) / AB_T, H, B);
  if (p.drop_threshold16 != 0u) e = launch_kernel(attention_bwd_kernel<true>, grid, dim3(AB_THREADS), AB_SMEM, s, 1, tq, tk, tv, tdo, tdq, p);
  else if (g_ab_query_major) e = launch_kernel(attention_bwd_kernel<false>, grid, dim3(AB_THREADS), AB_SMEM, s, 1, tq, tk, tv, tdo, tdq, p);
  else if ((rc = launch_attention_bwd_kt(tq, tk, tv, tdo, tdq, p, pp, grid, s))) return rc;
  if (e != cudaSuccess) return check_cuda(e, "attention_bwd launch");
  B200_CHECK_LAUNCH("attention_bwd");

  {
    const long long total = (long long)B * Sq * W / 4;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    e = launch_kernel(attention_dq_convert_kernel, dim3((unsigned)blocks), dim3(256), 0, s, 1, dq32, reinterpret_cast<bf16*>(dq),
                      ld_dq, dq_col0, (long long)B * Sq, W);
    if (e != cudaSuccess) return check_cuda(e, "attention_dq_convert launch");
    B200_CHECK_LAUNCH("attention_dq_convert");
  }
  return 0;
}
