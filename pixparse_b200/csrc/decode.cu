// Single-token greedy decode kernels (BASELINE configs[4]: cruller_large_6layers eval_ocr, 16 pages).
//
// One decode step feeds ONE new token per page through the BART decoder: every linear layer has M = batch <= 16 rows,
// so the work is streaming the weights once (HBM-bound, 2 bytes per weight) -- a 128-row tcgen05 tile would use an
// eighth of the tensor core and, worse, N / 256 CTAs (4 for N = 1024) to pull the weights. These kernels instead give
// every warp two weight rows and all M activations:
//   decode_linear      y[M, N] = x[M, K] W[N, K]^T (+ bias, GELU, fp32 residual)  or, for the LM head, the per-CTA
//                      (max, argmax) of the bf16-rounded logits -- the logits never reach HBM
//   decode_attention   one query per (page, head) against a strided K / V store (self-attention cache or the cached
//                      cross-attention projection of the image tokens), keys split over the warps of a CTA, online
//                      softmax, optional pad-key mask read straight from the generated ids (attention_mask =
//                      input_ids.ne(pad), models/text_decoder_hf.py:68)
//   decode_embed       token + learned position embedding of the current position
//   decode_finalize    argmax over the CTAs' partial results, append the token, EOS bookkeeping, advance the position
// Every per-step quantity (position, ids) lives in device memory, so one step is a fixed kernel sequence with fixed
// arguments: the host captures it ONCE in a CUDA graph and replays it per token (pixparse_b200/engine.py greedy_decode).
// Replaces the uncached HF generate-style loop of utils/ocr_utils.py:165-197 (which re-feeds the whole prefix and
// re-projects the image tokens every step).
#include "common.cuh"
#include "../../include/pixparse_b200.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace b200 {

constexpr int DL_WARPS = 8;        // warps per CTA: they split K in 32-element chunks (chunk c goes to warp c % 8)
constexpr int DL_M = 16;           // activation rows per launch = the M of one mma.m16n8k16

__device__ __forceinline__ unsigned long long argmax_pack(float v, int n) {
  // monotonic map of the float's bits, then the column index inverted: larger key = larger value, ties -> smaller index
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)n);
}

struct DecodeLinearParams {
  const bf16* x; long long ldx;
  const bf16* w; long long ldw;
  const float* bias;
  const float* resid; long long ld_resid;
  bf16* out16; float* out32; long long ldo;
  const int* pos; long long out_pos_stride;      // outputs are shifted by *pos * out_pos_stride elements (KV-cache append)
  unsigned long long* argmax_partial;            // [DL_M][gridDim.x] when mode == argmax
  int M, N, K, act;
  // split output: columns >= n_split go to out2 (bf16, pitch ldo2, column n - n_split) and only THEY take the position
  // shift -- the packed q | k | v projection of a decode step writes q to a dense buffer and k | v into the cache
  int n_split; bf16* out2; long long ldo2;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// read-once data (weights): non-coherent load without L1 allocation
__device__ __forceinline__ uint4 ld_nc_na(const uint4* ptr) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}

// y[16, N] = x[16, K] W[N, K]^T for the <= 16 pages of a decode step. The work is reading W once, so the kernel is built
// around bytes in flight, not math: a CTA owns 8 G weight rows, its 8 warps split K, and every lane fetches 16-byte pieces of
// W with plain coalesced loads (4 lanes = 64 contiguous bytes of one row) that are ALL issued before the first use. The
// 16 x 8 x 16 warp-level MMA does the arithmetic straight from those registers -- the reduction index of a dot product may be
// permuted freely, so the lane's 8 consecutive weights ARE two B fragments as loaded (k-steps {.x,.y} and {.z,.w}) provided the
// activation fragments use the same permutation (two 16-byte loads of rows g and g + 8 at the same k). No shared-memory
// staging, no conversions, no shuffles in the main loop; partial sums of the 8 warps meet in shared memory.
// (tcgen05 needs M = 128 and operands staged in shared memory -- for 16 rows bound by HBM it would only add latency.)
template <int G>
__global__ void __launch_bounds__(DL_WARPS * 32)
decode_linear_kernel(const DecodeLinearParams p) {
  constexpr int U = G == 4 ? 2 : 4;      // chunks in flight per warp and loop trip
  constexpr int NC = 8 * G;              // output columns per CTA
  __shared__ float s_red[DL_WARPS][DL_M][NC + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, j = lane & 3;
  const int n_base = blockIdx.x * NC;
  float acc[G][4];
#pragma unroll
  for (int gi = 0; gi < G; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[gi][i] = 0.f;
  const int nchunks = (p.K + 31) >> 5;
  // every load is unconditional (addresses clamped into the operands): rows >= M and columns >= N compute garbage that the
  // epilogue never stores, chunks past K are cancelled by zeroing the activation fragment. Unpredicated loads are what the
  // compiler hoists to the top of the trip, so a warp's U * (G + 2) 16-byte loads are in flight together.
  const bf16* xa_row = p.x + (long long)min(g8, p.M - 1) * p.ldx + j * 8;
  const bf16* xb_row = p.x + (long long)min(g8 + 8, p.M - 1) * p.ldx + j * 8;
  const bf16* w_row[G];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) w_row[gi] = p.w + (long long)min(n_base + gi * 8 + g8, p.N - 1) * p.ldw + j * 8;
  const int last_chunk = nchunks - 1;
  // Programmatic dependent launch: this grid may be resident while its predecessor is still running. The weights do not
  // depend on it, so the first trip's weight loads go out BEFORE the grid-dependency wait; the activations come after.
  pdl_launch_dependents();
  uint4 wv[U][G];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int k = min(warp + DL_WARPS * u, last_chunk) * 32;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) wv[u][gi] = ld_nc_na(reinterpret_cast<const uint4*>(w_row[gi] + k));
  }
  pdl_wait();
  for (int c0 = warp; c0 < nchunks; c0 += DL_WARPS * U) {
    uint4 xa[U], xb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = min(c0 + DL_WARPS * u, last_chunk) * 32;
      if (c0 != warp) {      // (warp-uniform) later trips fetch their weights here
#pragma unroll
        for (int gi = 0; gi < G; ++gi) wv[u][gi] = ld_nc_na(reinterpret_cast<const uint4*>(w_row[gi] + k));
      }
      xa[u] = __ldg(reinterpret_cast<const uint4*>(xa_row + k));
      xb[u] = __ldg(reinterpret_cast<const uint4*>(xb_row + k));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + DL_WARPS * u > last_chunk) xa[u] = xb[u] = make_uint4(0u, 0u, 0u, 0u);      // (warp-uniform)
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        mma_bf16_16816(acc[gi], xa[u].x, xb[u].x, xa[u].y, xb[u].y, wv[u][gi].x, wv[u][gi].y);
        mma_bf16_16816(acc[gi], xa[u].z, xb[u].z, xa[u].w, xb[u].w, wv[u][gi].z, wv[u][gi].w);
      }
    }
  }
  // accumulator fragment: c0, c1 = (row g8, cols 2j, 2j + 1), c2, c3 = (row g8 + 8, same cols)
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    s_red[warp][g8][gi * 8 + 2 * j] = acc[gi][0];
    s_red[warp][g8][gi * 8 + 2 * j + 1] = acc[gi][1];
    s_red[warp][g8 + 8][gi * 8 + 2 * j] = acc[gi][2];
    s_red[warp][g8 + 8][gi * 8 + 2 * j + 1] = acc[gi][3];
  }
  __syncthreads();
  const long long shift = (p.pos != nullptr && p.argmax_partial == nullptr) ? (long long)(*p.pos) * p.out_pos_stride : 0;
  for (int o = threadIdx.x; o < DL_M * NC; o += DL_WARPS * 32) {      // (whole warps enter or skip: NC divides 32)
    const int m = o / NC, nl = o % NC, n = n_base + nl;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < DL_WARPS; ++w) v += s_red[w][m][nl];
    if (p.argmax_partial != nullptr) {
      // LM head: the logit the teacher-forced path would have stored (bf16), compared without leaving the chip
      unsigned long long key = (n < p.N) ? argmax_pack(round_bf16(v), n) : 0ull;
#pragma unroll
      for (int off = NC / 2; off > 0; off >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, off);
        key = other > key ? other : key;
      }
      if (nl == 0) p.argmax_partial[(long long)m * gridDim.x + blockIdx.x] = key;
      continue;
    }
    if (m >= p.M || n >= p.N) continue;
    if (p.bias != nullptr) v += __ldg(p.bias + n);
    if (p.act == 1) v = gelu_erf(round_bf16(v));      // nn.GELU on the bf16 pre-activation, as the GEMM epilogue does
    if (p.resid != nullptr) v += p.resid[(long long)m * p.ld_resid + n];
    if (p.out2 != nullptr && n >= p.n_split) {
      p.out2[shift + (long long)m * p.ldo2 + (n - p.n_split)] = __float2bfloat16_rn(v);
      continue;
    }
    const long long sh = p.out2 != nullptr ? 0 : shift;
    if (p.out32 != nullptr) p.out32[sh + (long long)m * p.ldo + n] = v;
    if (p.out16 != nullptr) p.out16[sh + (long long)m * p.ldo + n] = __float2bfloat16_rn(v);
  }
}

// groups of 8 weight rows per CTA: 2 when that still leaves >= 2 CTAs per SM (the LM head: 3142 CTAs, three resident per SM;
// measured 34 us against 39 us at G = 4 and 44 us at G = 1), else 1 (128 CTAs for N = 1024). PIXPARSE_B200_DECODE_G overrides.
static int decode_linear_groups(int n) {
  static int forced = -1;
  if (forced < 0) {
    const char* env = getenv("PIXPARSE_B200_DECODE_G");
    forced = env != nullptr ? atoi(env) : 0;
  }
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  return (n + 15) / 16 >= 2 * num_sms() ? 2 : 1;
}

// ------------------------------------------------------------------------------------------------
struct DecodeAttnParams {
  const bf16* q; long long ldq; int q_col0;
  const bf16* k; const bf16* v; long long ld_kv, kv_bstride; int k_col0, v_col0;
  bf16* out; long long ld_out;
  const int* pos; int sk;                 // keys = pos ? *pos + 1 : sk
  const long long* key_ids; long long ld_ids; long long pad_id;      // key j of page b is hidden when key_ids[b][j] == pad_id
  int H; float scale_log2;
};
constexpr int DA_WARPS = 4;
constexpr int DA_MAX_SPLITS = 8;      // portable cluster size

// Grid (heads, pages, key splits); the splits of one (page, head) form a thread-block cluster and merge their partial
// (max, sum, output) through distributed shared memory -- no workspace, no atomics, graph-capturable as is.
// A warp takes 32 keys per iteration as 8 fully coalesced 16-byte loads of K and 8 of V: lane = 8 g + c holds the 8-dim
// chunk c of key 4 i + g in iteration i (one 128-byte row per 8 lanes), so all 16 loads of a block are independent and
// in flight together (8 KB per warp). Scores are reduced over the 8 lanes of a key group; lane 8 g + c keeps the score of
// key 4 c + g, which is exactly where the V phase of group g looks for it (shuffle inside the group).
__global__ void __launch_bounds__(DA_WARPS * 32)
decode_attention_kernel(const DecodeAttnParams p) {
  __shared__ float s_acc[DA_WARPS][64];
  __shared__ float s_m[DA_WARPS], s_l[DA_WARPS];
  __shared__ float s_out[64];
  __shared__ float s_ML[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, c = lane & 7;
  const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z, nsplit = gridDim.z;
  pdl_launch_dependents();      // the next linear may start pulling its weights
  pdl_wait();                   // q (and the position) come from the predecessors
  const int Sk = p.pos != nullptr ? *p.pos + 1 : p.sk;
  const int nblk = (Sk + 31) >> 5;      // balanced partition of the 32-key blocks over the splits
  const int blk0 = (int)(((long long)nblk * split) / nsplit), blk1 = (int)(((long long)nblk * (split + 1)) / nsplit);
  // this lane's 8 dims of the query, pre-multiplied by scale * log2(e)
  f32x2 q0, q1, q2, q3;
  {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p.q + (long long)b * p.ldq + p.q_col0 + h * 64 + 8 * c));
    const f32x2 sc = f2_splat(p.scale_log2);
    q0 = f2_mul(f2_pack(bf16_lo(t.x), bf16_hi(t.x)), sc); q1 = f2_mul(f2_pack(bf16_lo(t.y), bf16_hi(t.y)), sc);
    q2 = f2_mul(f2_pack(bf16_lo(t.z), bf16_hi(t.z)), sc); q3 = f2_mul(f2_pack(bf16_lo(t.w), bf16_hi(t.w)), sc);
  }
  const bf16* kbase = p.k + (long long)b * p.kv_bstride + p.k_col0 + h * 64 + 8 * c;
  const bf16* vbase = p.v + (long long)b * p.kv_bstride + p.v_col0 + h * 64 + 8 * c;
  float m_run = -INFINITY, l_run = 0.f;
  f32x2 a0 = f2_splat(0.f), a1 = f2_splat(0.f), a2 = f2_splat(0.f), a3 = f2_splat(0.f);      // dims 8 c .. 8 c + 7 (this group's keys)
  // Software pipeline: the K rows of the NEXT block are requested as soon as this block's scores are done (their registers are
  // free), its V rows as soon as this block's P V is done -- every warp keeps 8 + 8 16-byte loads in flight while it computes,
  // instead of alternating between a load phase and a compute phase.
  uint4 kk[8], vv[8];
  auto load_k = [&](int kb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // unconditional loads (keys past the end re-read the last key: its score is masked below, so its V row, which is
      // finite, is weighted by 0): unpredicated loads are issued together
      const int key = min(kb + 4 * i + g, Sk - 1);
      kk[i] = ld_nc_na(reinterpret_cast<const uint4*>(kbase + (long long)key * p.ld_kv));
    }
  };
  auto load_v = [&](int kb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = min(kb + 4 * i + g, Sk - 1);
      vv[i] = ld_nc_na(reinterpret_cast<const uint4*>(vbase + (long long)key * p.ld_kv));
    }
  };
  if (blk0 + warp < blk1) {
    load_k((blk0 + warp) << 5);
    load_v((blk0 + warp) << 5);
  }
  for (int blk = blk0 + warp; blk < blk1; blk += DA_WARPS) {
    const int kb = blk << 5;
    const bool has_next = blk + DA_WARPS < blk1;      // (warp-uniform)
    const int my_key = kb + 4 * c + g;
    bool valid = my_key < Sk;
    if (valid && p.key_ids != nullptr) valid = p.key_ids[(long long)b * p.ld_ids + my_key] != p.pad_id;
    float s = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f32x2 d = f2_mul(f2_pack(bf16_lo(kk[i].x), bf16_hi(kk[i].x)), q0);
      d = f2_fma(f2_pack(bf16_lo(kk[i].y), bf16_hi(kk[i].y)), q1, d);
      d = f2_fma(f2_pack(bf16_lo(kk[i].z), bf16_hi(kk[i].z)), q2, d);
      d = f2_fma(f2_pack(bf16_lo(kk[i].w), bf16_hi(kk[i].w)), q3, d);
      float lo, hi;
      f2_unpack(d, lo, hi);
      float r = lo + hi;
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      r += __shfl_xor_sync(0xffffffffu, r, 4);
      if (i == c && valid) s = r;
    }
    if (has_next) load_k((blk + DA_WARPS) << 5);
    // (no early-out for a fully hidden block: everything below is branch-free arithmetic)
    const float m_new = fmaxf(m_run, warp_max(s));
    const float pr = valid ? ex2_approx(s - m_new) : 0.f;
    const float alpha = m_new == -INFINITY ? 1.f : ex2_approx(m_run - m_new);      // 0 on the first visible block (m_run = -inf)
    l_run = l_run * alpha + warp_sum(pr);
    const f32x2 al = f2_splat(alpha);
    a0 = f2_mul(a0, al); a1 = f2_mul(a1, al); a2 = f2_mul(a2, al); a3 = f2_mul(a3, al);
    const float pb = round_bf16(pr);      // the flash kernel feeds bf16 probabilities to its P V product
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const f32x2 pj = f2_splat(__shfl_sync(0xffffffffu, pb, i, 8));      // key 4 i + g lives in lane i of this group
      a0 = f2_fma(pj, f2_pack(bf16_lo(vv[i].x), bf16_hi(vv[i].x)), a0);
      a1 = f2_fma(pj, f2_pack(bf16_lo(vv[i].y), bf16_hi(vv[i].y)), a1);
      a2 = f2_fma(pj, f2_pack(bf16_lo(vv[i].z), bf16_hi(vv[i].z)), a2);
      a3 = f2_fma(pj, f2_pack(bf16_lo(vv[i].w), bf16_hi(vv[i].w)), a3);
    }
    if (has_next) load_v((blk + DA_WARPS) << 5);
    m_run = m_new;
  }
  // sum the four key groups of the warp, then the warps of the CTA, then the CTAs of the cluster
  float o[8];
  f2_unpack(a0, o[0], o[1]); f2_unpack(a1, o[2], o[3]); f2_unpack(a2, o[4], o[5]); f2_unpack(a3, o[6], o[7]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[warp][8 * c + i] = o[i];
  }
  if (lane == 0) {
    s_m[warp] = m_run;
    s_l[warp] = l_run;
  }
  __syncthreads();
  float M = -INFINITY, L = 0.f, acc = 0.f;
  if (threadIdx.x < 64) {
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) M = fmaxf(M, s_m[w]);
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float sc = s_m[w] == -INFINITY ? 0.f : ex2_approx(s_m[w] - M);
      L += s_l[w] * sc;
      acc += s_acc[w][threadIdx.x] * sc;
    }
  }
  if (nsplit == 1) {
    if (threadIdx.x < 64)
      p.out[(long long)b * p.ld_out + h * 64 + threadIdx.x] = __float2bfloat16_rn(L > 0.f ? acc / L : 0.f);
    return;
  }
  cg::cluster_group cluster = cg::this_cluster();
  if (threadIdx.x < 64) s_out[threadIdx.x] = acc;
  if (threadIdx.x == 0) {
    s_ML[0] = M;
    s_ML[1] = L;
  }
  cluster.sync();
  if (split == 0 && threadIdx.x < 64) {
    float Mg = -INFINITY;
    for (int r = 0; r < nsplit; ++r) Mg = fmaxf(Mg, cluster.map_shared_rank(s_ML, r)[0]);
    float Lg = 0.f, og = 0.f;
    for (int r = 0; r < nsplit; ++r) {
      const float* ml = cluster.map_shared_rank(s_ML, r);
      const float sc = ml[0] == -INFINITY ? 0.f : ex2_approx(ml[0] - Mg);
      Lg += ml[1] * sc;
      og += cluster.map_shared_rank(s_out, r)[threadIdx.x] * sc;
    }
    p.out[(long long)b * p.ld_out + h * 64 + threadIdx.x] = __float2bfloat16_rn(Lg > 0.f ? og / Lg : 0.f);
  }
  cluster.sync();      // nobody leaves while its shared memory is still being read
}

// ------------------------------------------------------------------------------------------------
__global__ void decode_embed_kernel(const long long* __restrict__ ids, long long ld_ids, const int* __restrict__ pos,
                                    const float* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                    float* __restrict__ x, int B, int D, int pos_offset, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = *pos;
  const int d4 = D / 4;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * d4; idx += gridDim.x * blockDim.x) {
    const int c = idx % d4, b = idx / d4;
    const long long id = ids[(long long)b * ld_ids + t];
    const float4 e = reinterpret_cast<const float4*>(tok_emb)[id * d4 + c];
    const float4 q = reinterpret_cast<const float4*>(pos_emb)[(long long)(t + pos_offset) * d4 + c];
    reinterpret_cast<float4*>(x)[idx] = make_float4(e.x * scale + q.x, e.y * scale + q.y, e.z * scale + q.z, e.w * scale + q.w);
  }
}

// state = {pos, done_step (-1 = not yet), steps_run}; finished[b] = row b has emitted EOS
__global__ void __launch_bounds__(1024)
decode_finalize_kernel(const unsigned long long* __restrict__ partial, int n_cta, long long* __restrict__ ids,
                       long long ld_ids, int* __restrict__ state, int* __restrict__ finished, int B, long long eos_id) {
  __shared__ int s_fin[DL_M];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  const int t = state[0];
  if (warp < B) {
    unsigned long long best = 0ull;
    for (int c = lane; c < n_cta; c += 32) {
      const unsigned long long v = partial[(long long)warp * n_cta + c];
      best = v > best ? v : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o);
      best = v > best ? v : best;
    }
    if (lane == 0) {
      const long long tok = (long long)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
      ids[(long long)warp * ld_ids + t + 1] = tok;
      const int f = finished[warp] | (tok == eos_id ? 1 : 0);
      finished[warp] = f;
      s_fin[warp] = f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int all = 1;
    for (int b = 0; b < B; ++b) all &= s_fin[b];
    if (all && state[1] < 0) state[1] = t;      // the reference loop breaks here WITHOUT appending this step's tokens
    state[0] = t + 1;
    state[2] += 1;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_decode_linear(const B200DecodeLinearArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200DecodeLinearArgs, "b200_decode_linear");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(a->x && a->w && a->m > 0 && a->n > 0 && a->k > 0, "b200_decode_linear: bad arguments");
  B200_CHECK_ARG(a->m <= DL_M, "b200_decode_linear: at most %d activation rows per call (got %d)", DL_M, a->m);
  B200_CHECK_ARG(a->k % 32 == 0 && a->ldx % 8 == 0 && a->ldw % 8 == 0,
                 "b200_decode_linear: K must be a multiple of 32, ldx and ldw multiples of 8");
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w)) & 15) == 0,
                 "b200_decode_linear: x and w must be 16-byte aligned");
  B200_CHECK_ARG(a->argmax_partial != nullptr || a->out_bf16 != nullptr || a->out_f32 != nullptr,
                 "b200_decode_linear: no output");
  B200_CHECK_ARG(a->act == 0 || a->act == 1, "b200_decode_linear: act must be 0 (none) or 1 (GELU)");
  DecodeLinearParams p;
  p.x = reinterpret_cast<const bf16*>(a->x); p.ldx = a->ldx;
  p.w = reinterpret_cast<const bf16*>(a->w); p.ldw = a->ldw;
  p.bias = a->bias; p.resid = a->resid; p.ld_resid = a->ld_resid;
  p.out16 = reinterpret_cast<bf16*>(a->out_bf16); p.out32 = a->out_f32; p.ldo = a->ldo;
  p.pos = a->pos; p.out_pos_stride = a->out_pos_stride;
  p.argmax_partial = reinterpret_cast<unsigned long long*>(a->argmax_partial);
  p.M = a->m; p.N = a->n; p.K = a->k; p.act = a->act;
  p.n_split = a->n_split; p.out2 = reinterpret_cast<bf16*>(a->out2_bf16); p.ldo2 = a->ldo2;
  B200_CHECK_ARG(a->out2_bf16 == nullptr || (a->n_split > 0 && a->n_split < a->n && a->argmax_partial == nullptr),
                 "b200_decode_linear: split output needs 0 < n_split < n");
  const int G = decode_linear_groups(a->n);
  const int grid = (a->n + 8 * G - 1) / (8 * G);
  cudaError_t err;
  if (G == 4) err = launch_kernel(decode_linear_kernel<4>, dim3(grid), dim3(DL_WARPS * 32), 0, s, 1, p);
  else if (G == 2) err = launch_kernel(decode_linear_kernel<2>, dim3(grid), dim3(DL_WARPS * 32), 0, s, 1, p);
  else err = launch_kernel(decode_linear_kernel<1>, dim3(grid), dim3(DL_WARPS * 32), 0, s, 1, p);
  B200_CHECK_ARG(err == cudaSuccess, "b200_decode_linear: launch failed: %s", cudaGetErrorString(err));
  B200_CHECK_LAUNCH("decode_linear");
  return 0;
}

extern "C" int b200_decode_linear_ctas(int n) {
  const int G = decode_linear_groups(n);
  return (n + 8 * G - 1) / (8 * G);
}

extern "C" int b200_decode_attention(const B200DecodeAttentionArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200DecodeAttentionArgs, "b200_decode_attention");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(a->q && a->k && a->v && a->out && a->batch > 0 && a->heads > 0, "b200_decode_attention: bad arguments");
  B200_CHECK_ARG(a->head_dim == 64, "b200_decode_attention: head_dim %d unsupported (only 64)", a->head_dim);
  B200_CHECK_ARG(a->pos != nullptr || a->sk > 0, "b200_decode_attention: need sk > 0 or a device position");
  B200_CHECK_ARG(a->ldq % 8 == 0 && a->ld_kv % 8 == 0 && a->kv_bstride % 8 == 0 && a->q_col0 % 8 == 0 && a->k_col0 % 8 == 0 &&
                     a->v_col0 % 8 == 0,
                 "b200_decode_attention: strides and column offsets must be multiples of 8 elements");
  DecodeAttnParams p;
  p.q = reinterpret_cast<const bf16*>(a->q); p.ldq = a->ldq; p.q_col0 = a->q_col0;
  p.k = reinterpret_cast<const bf16*>(a->k); p.v = reinterpret_cast<const bf16*>(a->v);
  p.ld_kv = a->ld_kv; p.kv_bstride = a->kv_bstride; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.out = reinterpret_cast<bf16*>(a->out); p.ld_out = a->ld_out;
  p.pos = a->pos; p.sk = a->sk;
  p.key_ids = a->key_ids; p.ld_ids = a->ld_ids; p.pad_id = a->pad_id;
  p.H = a->heads; p.scale_log2 = a->scale * 1.4426950408889634f;
  // key splits: ~10 blocks of 32 keys per warp (cross-attention over the image tokens: 2509 keys -> 2 CTAs per (page, head),
  // 512 CTAs all resident at 16 pages x 16 heads: 38 us against 40 us with 4 splits, 44-48 us with 3 or 5-8);
  // the growing self-attention prefix (device-side length) stays in one CTA
  int nsplit = 1;
  if (a->pos == nullptr) {
    const int nblk = (a->sk + 31) / 32;
    nsplit = (nblk + DA_WARPS * 10 - 1) / (DA_WARPS * 10);
    static int forced = -1;      // PIXPARSE_B200_DECODE_SPLITS=<n>: experiments
    if (forced < 0) {
      const char* env = getenv("PIXPARSE_B200_DECODE_SPLITS");
      forced = env != nullptr ? atoi(env) : 0;
    }
    if (forced > 0) nsplit = forced;
    nsplit = nsplit < 1 ? 1 : (nsplit > DA_MAX_SPLITS ? DA_MAX_SPLITS : nsplit);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a->heads, a->batch, nsplit);
  cfg.blockDim = dim3(DA_WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = nsplit;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  cudaError_t err = cudaLaunchKernelEx(&cfg, decode_attention_kernel, p);
  B200_CHECK_ARG(err == cudaSuccess, "b200_decode_attention: launch failed: %s", cudaGetErrorString(err));
  B200_CHECK_LAUNCH("decode_attention");
  return 0;
}

extern "C" int b200_decode_embed(const long long* ids, long long ld_ids, const int* pos, const float* tok_emb,
                                 const float* pos_emb, float* x, int B, int D, int pos_offset, float scale, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(ids && pos && tok_emb && pos_emb && x && B > 0 && D % 4 == 0, "b200_decode_embed: bad arguments");
  cudaError_t err = launch_kernel(decode_embed_kernel, dim3((B * D / 4 + 255) / 256), dim3(256), 0, s, 1, ids, ld_ids, pos,
                                  tok_emb, pos_emb, x, B, D, pos_offset, scale);
  B200_CHECK_ARG(err == cudaSuccess, "b200_decode_embed: launch failed: %s", cudaGetErrorString(err));
  B200_CHECK_LAUNCH("decode_embed");
  return 0;
}

extern "C" int b200_decode_finalize(const void* argmax_partial, int n_cta, long long* ids, long long ld_ids, int* state,
                                    int* finished, int B, long long eos_id, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(argmax_partial && ids && state && finished && B > 0 && B <= DL_M && n_cta > 0,
                 "b200_decode_finalize: bad arguments");
  cudaError_t err = launch_kernel(decode_finalize_kernel, dim3(1), dim3(1024), 0, s, 1,
                                  reinterpret_cast<const unsigned long long*>(argmax_partial), n_cta, ids, ld_ids, state,
                                  finished, B, eos_id);
  B200_CHECK_ARG(err == cudaSuccess, "b200_decode_finalize: launch failed: %s", cudaGetErrorString(err));
  B200_CHECK_LAUNCH("decode_finalize");
  return 0;
}
