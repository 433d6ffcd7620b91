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
#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

int make_att_tmap(CUtensorMap* m, const void* base, int B, int S, long long width, long long ld);

constexpr int AB_T = 128;                 // tile edge (queries and keys)
constexpr int AB_D = 64;
constexpr int AB_THREADS = 192;
constexpr int AB_TILE = AB_T * AB_D * 2;  // 16 KB bf16 tile
constexpr int AB_SMEM = 12 * AB_TILE + 256 + 1024;   // K, V, Q[2], dO[2], P(2), dS(2), dQ staging(2) = 192 KB

constexpr uint32_t TB_S = 0, TB_DP = 128, TB_DV = 256, TB_DK = 320, TB_DQ = 384;

struct AttBwdParams {
  int B, H, Sq, Sk, causal;
  float scale, scale_log2;
  const float* lse;      // [B, H, Sq]
  const float* dsum;     // [B, H, Sq]  rowsum(dO * O)
  bf16* dk; long long ld_dk; int dk_col0;
  bf16* dv; long long ld_dv; int dv_col0;
  int q_col0, k_col0, v_col0, do_col0;
};

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
  uint64_t* sdp_full = bars + 5;
  uint64_t* ps_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dkv_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

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

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_do);
    prefetch_tmap(&tmap_dq);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(ps_full, 128);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_3d(sK, &tmap_k, kv_full, p.k_col0 + h * AB_D, k0, b);
      tma_load_3d(sV, &tmap_v, kv_full, p.v_col0 + h * AB_D, k0, b);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int q0 = (i_begin + t) * AB_T;
        mbar_wait(&q_empty[s], ph ^ 1);
        mbar_expect_tx(&q_full[s], 2 * AB_TILE);
        tma_load_3d(sQ + s * AB_TILE, &tmap_q, &q_full[s], p.q_col0 + h * AB_D, q0, b);
        tma_load_3d(sdO + s * AB_TILE, &tmap_do, &q_full[s], p.do_col0 + h * AB_D, q0, b);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_iter > 0) {
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

      auto issue_s_dp = [&](int t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(&q_full[s], ph);
        tc_fence_after();
        const uint64_t dQ_k = make_smem_desc(smem_u32(sQ + s * AB_TILE), 16, 1024);
        const uint64_t dO_k = make_smem_desc(smem_u32(sdO + s * AB_TILE), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + TB_S, dQ_k + (uint64_t)(2 * k), dK_k + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + TB_DP, dO_k + (uint64_t)(2 * k), dV_k + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
      };

      issue_s_dp(0);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        mbar_wait(ps_full, t & 1);
        tc_fence_after();
        const uint64_t dO_mn = make_smem_desc(smem_u32(sdO + s * AB_TILE), 16384, 1024);
        const uint64_t dQ_mn = make_smem_desc(smem_u32(sQ + s * AB_TILE), 16384, 1024);
        // dV += P^T dO ; dK += dS^T Q    (reduction over the 128 queries, 16 per UMMA)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem_base + TB_DV, dP_mn + (uint64_t)(128 * k), dO_mn + (uint64_t)(128 * k), idesc_t,
                  (t > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem_base + TB_DK, dS_mn + (uint64_t)(128 * k), dQ_mn + (uint64_t)(128 * k), idesc_t,
                  (t > 0 || k > 0) ? 1u : 0u);
        // dQ = dS K    (reduction over the 128 keys; dS read K-major: 64-key chunk = k / 4, 32 B per step inside)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dS_k = make_smem_desc(smem_u32(sdS + (k >> 2) * AB_TILE + (k & 3) * 32), 16, 1024);
          umma_ss(tmem_base + TB_DQ, dS_k, dK_mn + (uint64_t)(128 * k), idesc_q, k > 0 ? 1u : 0u);
        }
        umma_commit(&q_empty[s]);
        umma_commit(dq_full);
        if (t + 1 < n_iter) issue_s_dp(t + 1);
      }
      umma_commit(dkv_full);
    }
  } else {
    // ===================== compute warps: one query row per thread =====================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int sw = row & 7;
    const float kLog2e = 1.4426950408889634f;
    for (int t = 0; t < n_iter; ++t) {
      const int q0 = (i_begin + t) * AB_T;
      const int qidx = q0 + row;
      const bool q_ok = qidx < p.Sq;
      const long long stat_idx = ((long long)b * p.H + h) * p.Sq + (q_ok ? qidx : 0);
      const float lse2 = q_ok ? p.lse[stat_idx] * kLog2e : INFINITY;   // +inf -> P = 0 for padded query rows
      const float dsum = q_ok ? p.dsum[stat_idx] : 0.f;
      int kmax = p.Sk - 1;
      if (p.causal) kmax = min(kmax, qidx + shift);
      const bool need_mask = (k0 + AB_T > p.Sk) || (p.causal && (k0 + AB_T - 1 > q0 + shift));
      mbar_wait(sdp_full, t & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t rs[32], rp[32];
        tmem_ld_32x32(lane_addr + TB_S + c * 32, rs);
        tmem_ld_32x32(lane_addr + TB_DP + c * 32, rp);
        tmem_ld_wait();
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float p0 = exp2f(fmaf(__uint_as_float(rs[2 * e]), p.scale_log2, -lse2));
          float p1 = exp2f(fmaf(__uint_as_float(rs[2 * e + 1]), p.scale_log2, -lse2));
          if (need_mask) {
            if (k0 + c * 32 + 2 * e > kmax) p0 = 0.f;
            if (k0 + c * 32 + 2 * e + 1 > kmax) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(rp[2 * e]) - dsum) * p.scale;
          const float d1 = p1 * (__uint_as_float(rp[2 * e + 1]) - dsum) * p.scale;
          pk[e] = pack_bf16(p0, p1);
          dk[e] = pack_bf16(d0, d1);
        }
        // keys [c*32, c*32+32) of this row: 64-key chunk c/2, 16-byte pieces (c%2)*4 .. +3, 128B-swizzled
        uint8_t* prow = sP + (c >> 1) * AB_TILE + row * 128;
        uint8_t* drow = sdS + (c >> 1) * AB_TILE + row * 128;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int piece = ((c & 1) * 4 + jj) ^ sw;
          *reinterpret_cast<uint4*>(prow + (piece << 4)) = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
          *reinterpret_cast<uint4*>(drow + (piece << 4)) = make_uint4(dk[4 * jj], dk[4 * jj + 1], dk[4 * jj + 2], dk[4 * jj + 3]);
        }
      }
      fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(ps_full);

      // ---- drain dQ_t: TMEM -> swizzled fp32 staging -> TMA reduce-add into the fp32 dQ accumulator
      mbar_wait(dq_full, t & 1);
      tc_fence_after();
      if (threadIdx.x == 0) tma_wait_group_read<0>();   // staging buffers no longer read by the previous reduce
      named_bar_sync(1, 128);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(lane_addr + TB_DQ + c * 32, r);
        tmem_ld_wait();
        uint8_t* rowp = sStage + c * AB_TILE + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        tma_reduce_add_3d(&tmap_dq, sStage, h * AB_D, q0, b);
        tma_reduce_add_3d(&tmap_dq, sStage + AB_TILE, h * AB_D + 32, q0, b);
        tma_commit_group();
      }
    }
    // ---- final: dK, dV (rows = keys of this tile) -> bf16 -> global
    if (n_iter > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    const int kidx = k0 + row;
    const bool k_ok = kidx < p.Sk;
    bf16* dkrow = p.dk + ((long long)b * p.Sk + (k_ok ? kidx : 0)) * p.ld_dk + p.dk_col0 + h * AB_D;
    bf16* dvrow = p.dv + ((long long)b * p.Sk + (k_ok ? kidx : 0)) * p.ld_dv + p.dv_col0 + h * AB_D;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bf16* orow = which == 0 ? dvrow : dkrow;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        if (n_iter > 0) {
          tmem_ld_32x32(lane_addr + (which == 0 ? TB_DV : TB_DK) + c * 32, r);
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
            *reinterpret_cast<uint4*>(orow + c * 32 + v4 * 8) = o;
          }
        }
      }
    }
    if (threadIdx.x == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem_base);
}

// D[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (one warp per (b,q,h) row of 64 elements)
__global__ void attention_bwd_prep_kernel(const bf16* __restrict__ o, long long ld_o, const bf16* __restrict__ d_o,
                                          long long ld_do, int do_col0, float* __restrict__ dsum, int B, int H,
                                          int Sq) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * Sq * H;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
       w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int hh = (int)(w % H);
    const long long bq = w / H;     // b * Sq + q
    const uint32_t a = *reinterpret_cast<const uint32_t*>(o + bq * ld_o + hh * 64 + lane * 2);
    const uint32_t g = *reinterpret_cast<const uint32_t*>(d_o + bq * ld_do + do_col0 + hh * 64 + lane * 2);
    float s = bf16_lo(a) * bf16_lo(g) + bf16_hi(a) * bf16_hi(g);
    s = warp_sum(s);
    if (lane == 0) {
      const int bb = (int)(bq / Sq), q = (int)(bq % Sq);
      dsum[((long long)bb * H + hh) * Sq + q] = s;
    }
  }
}

// dq bf16 [rows, ld_dq] (columns dq_col0 ..) = bf16(dq32 [rows, width])
__global__ void attention_dq_convert_kernel(const float* __restrict__ dq32, bf16* __restrict__ dq, long long ld_dq,
                                            int dq_col0, long long rows, int width) {
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

extern "C" long long b200_attention_bwd_workspace_bytes(int B, int H, int Sq) {
  // fp32 dQ accumulator [B*Sq, H*64] followed by D [B, H, Sq]
  return ((long long)B * Sq * H * 64 + (long long)B * H * Sq) * 4;
}

extern "C" int b200_attention_bwd(const void* q, long long ldq, int q_col0, const void* k, long long ldk, int k_col0,
                                  const void* v, long long ldv, int v_col0, const void* o, long long ld_o,
                                  const void* d_o, long long ld_do, int do_col0, const float* lse, void* dq,
                                  long long ld_dq, int dq_col0, void* dk, long long ld_dk, int dk_col0, void* dv,
                                  long long ld_dv, int dv_col0, void* workspace, int B, int H, int Sq, int Sk,
                                  int head_dim, int causal, float scale, void* stream) {
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
  cudaError_t e = cudaMemsetAsync(dq32, 0, (size_t)B * Sq * W * 4, s);
  if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(dq32)");

  {
    const long long warps = (long long)B * Sq * H;
    long long blocks = (warps * 32 + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    attention_bwd_prep_kernel<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<const bf16*>(o), ld_o,
                                                         reinterpret_cast<const bf16*>(d_o), ld_do, do_col0, dsum, B,
                                                         H, Sq);
    B200_CHECK_LAUNCH("attention_bwd_prep");
  }

  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = make_att_tmap(&tq, q, B, Sq, q_col0 + (long long)W, ldq))) return rc;
  if ((rc = make_att_tmap(&tk, k, B, Sk, k_col0 + (long long)W, ldk))) return rc;
  if ((rc = make_att_tmap(&tv, v, B, Sk, v_col0 + (long long)W, ldv))) return rc;
  if ((rc = make_att_tmap(&tdo, d_o, B, Sq, do_col0 + (long long)W, ld_do))) return rc;
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
  p.dk = reinterpret_cast<bf16*>(dk); p.ld_dk = ld_dk; p.dk_col0 = dk_col0;
  p.dv = reinterpret_cast<bf16*>(dv); p.ld_dv = ld_dv; p.dv_col0 = dv_col0;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.do_col0 = do_col0;
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attention_bwd)");
    configured = true;
  }
  dim3 grid((Sk + AB_T - 1) / AB_T, H, B);
  attention_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, s>>>(tq, tk, tv, tdo, tdq, p);
  B200_CHECK_LAUNCH("attention_bwd");

  {
    const long long total = (long long)B * Sq * W / 4;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    attention_dq_convert_kernel<<<(int)blocks, 256, 0, s>>>(dq32, reinterpret_cast<bf16*>(dq), ld_dq, dq_col0,
                                                           (long long)B * Sq, W);
    B200_CHECK_LAUNCH("attention_dq_convert");
  }
  return 0;
}
