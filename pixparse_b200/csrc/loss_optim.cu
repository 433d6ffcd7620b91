// Softmax-cross-entropy over the vocabulary (fwd + bwd in one pass over HBM) and the optimizer side of the
// train step: deterministic global grad-norm, fused clip + AdamW + bf16 weight refresh + grad zeroing.
//
// Replaces: nn.CrossEntropyLoss(ignore_index=-100) (task/task_cruller_pretrain.py:118,251-254; K13),
// timm dispatch_clip_grad -> clip_grad_norm_ (:264-277; K15), torch.optim.AdamW foreach step via
// create_optimizer_v2 (:196-203,278; K16), optimizer.zero_grad (:295; K17).
#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// cross entropy
// ------------------------------------------------------------------------------------------------
constexpr int CE_THREADS = 512;

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < (int)(blockDim.x >> 5)) ? red[lane] : -INFINITY;
  r = warp_max(r);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// stats[0] = number of valid (!= ignore_index) targets, stats[1] = 0 (mean-loss accumulator)
__global__ void ce_prepare_kernel(const long long* __restrict__ targets, int n, long long ignore_index,
                                  float* __restrict__ stats) {
  __shared__ float red[32];
  float c = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += (targets[i] != ignore_index) ? 1.f : 0.f;
  c = block_reduce_sum(c, red);
  if (threadIdx.x == 0) {
    stats[0] = c;
    stats[1] = 0.f;
  }
}

// One CTA per row; the row is staged once in shared memory (bf16), so HBM sees one read and one write.
//   loss   : stats[1] += (logsumexp(row) - row[target]) / n_valid
//   dlogits: (softmax(row) - onehot(target)) * grad_scale / n_valid     (zeros for ignored rows)
__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_bwd_kernel(const bf16* __restrict__ logits, long long ld, const long long* __restrict__ targets,
                  bf16* __restrict__ dlogits, long long ldd, float* __restrict__ row_loss,
                  float* __restrict__ stats, int V, long long ignore_index, float grad_scale) {
  extern __shared__ uint4 ce_smem[];
  __shared__ float red[32];
  const int row = blockIdx.x;
  const long long tgt = targets[row];
  const int nvec = (V + 7) / 8;
  uint4* out = dlogits ? reinterpret_cast<uint4*>(dlogits + (long long)row * ldd) : nullptr;
  if (tgt == ignore_index) {
    if (out)
      for (int i = threadIdx.x; i < nvec; i += CE_THREADS) out[i] = make_uint4(0u, 0u, 0u, 0u);
    if (row_loss && threadIdx.x == 0) row_loss[row] = 0.f;
    return;
  }
  const uint4* in = reinterpret_cast<const uint4*>(logits + (long long)row * ld);
  float mx = -INFINITY;
  // bulk of the row: four independent 16-byte loads in flight per thread (one per iteration left the HBM pipe a third full)
  const int nbulk = (nvec - 1) / (4 * CE_THREADS) * (4 * CE_THREADS);      // never includes the (possibly padded) last vector
  for (int i0 = threadIdx.x; i0 < nbulk; i0 += 4 * CE_THREADS) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(in + i0 + u * CE_THREADS);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ce_smem[i0 + u * CE_THREADS] = v[u];
      mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(bf16_lo(v[u].x), bf16_hi(v[u].x)), fmaxf(bf16_lo(v[u].y), bf16_hi(v[u].y))),
                           fmaxf(fmaxf(bf16_lo(v[u].z), bf16_hi(v[u].z)), fmaxf(bf16_lo(v[u].w), bf16_hi(v[u].w)))));
    }
  }
  for (int i = nbulk + threadIdx.x; i < nvec; i += CE_THREADS) {
    uint4 v = in[i];
    if (i == nvec - 1 && (V & 7)) {
      // mask the padding columns of the last vector with -inf (bf16 0xff80)
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
      for (int e = (V & 7); e < 8; ++e) {
        const int wi = e >> 1;
        w[wi] = (e & 1) ? ((w[wi] & 0x0000ffffu) | 0xff800000u) : ((w[wi] & 0xffff0000u) | 0x0000ff80u);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    ce_smem[i] = v;
    mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(bf16_lo(v.x), bf16_hi(v.x)), fmaxf(bf16_lo(v.y), bf16_hi(v.y))),
                         fmaxf(fmaxf(bf16_lo(v.z), bf16_hi(v.z)), fmaxf(bf16_lo(v.w), bf16_hi(v.w)))));
  }
  mx = block_reduce_max(mx, red);   // contains __syncthreads: smem row visible to all
  const float kLog2e = 1.4426950408889634f;
  const float mneg = -mx * kLog2e;
  const int tvec = (int)(tgt >> 3), te = (int)(tgt & 7);
  __shared__ float s_xt;
  // second pass over the staged row: e = exp(x - max) once per element (MUFU work is the larger half of this kernel),
  // summed in fp32 and written back over the logit as bf16 for the gradient pass; the target logit is rescued first
  float sum = 0.f;
  for (int i = threadIdx.x; i < nvec; i += CE_THREADS) {
    const uint4 v = ce_smem[i];
    if (i == tvec) s_xt = __bfloat162float(reinterpret_cast<const bf16*>(&v)[te]);
    const float e0 = ex2_approx(fmaf(bf16_lo(v.x), kLog2e, mneg)), e1 = ex2_approx(fmaf(bf16_hi(v.x), kLog2e, mneg));
    const float e2 = ex2_approx(fmaf(bf16_lo(v.y), kLog2e, mneg)), e3 = ex2_approx(fmaf(bf16_hi(v.y), kLog2e, mneg));
    const float e4 = ex2_approx(fmaf(bf16_lo(v.z), kLog2e, mneg)), e5 = ex2_approx(fmaf(bf16_hi(v.z), kLog2e, mneg));
    const float e6 = ex2_approx(fmaf(bf16_lo(v.w), kLog2e, mneg)), e7 = ex2_approx(fmaf(bf16_hi(v.w), kLog2e, mneg));
    sum += ((e0 + e1) + (e2 + e3)) + ((e4 + e5) + (e6 + e7));
    if (out) ce_smem[i] = make_uint4(pack_bf16(e0, e1), pack_bf16(e2, e3), pack_bf16(e4, e5), pack_bf16(e6, e7));
  }
  sum = block_reduce_sum(sum, red);     // contains __syncthreads: s_xt and the rewritten row are visible
  const float n_valid = stats[0];
  const float inv_n = n_valid > 0.f ? 1.0f / n_valid : 0.f;
  if (threadIdx.x == 0) {
    const float l = logf(sum) + mx - s_xt;
    if (row_loss) row_loss[row] = l;
    atomicAdd(stats + 1, l * inv_n);
  }
  if (out) {
    const float gs = grad_scale * inv_n;
    const float c = gs / sum;
#pragma unroll 4
    for (int i = threadIdx.x; i < nvec; i += CE_THREADS) {
      const uint4 v = ce_smem[i];
      float p[8] = {bf16_lo(v.x) * c, bf16_hi(v.x) * c, bf16_lo(v.y) * c, bf16_hi(v.y) * c,
                    bf16_lo(v.z) * c, bf16_hi(v.z) * c, bf16_lo(v.w) * c, bf16_hi(v.w) * c};
      if (i == tvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e == te) p[e] -= gs;
      }
      out[i] = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
    }
  }
}

// Persistent, software-pipelined form of the kernel above (used whenever two rows fit in shared memory, i.e. always for
// the BART vocabularies): one CTA per SM loops over rows; a producer thread moves whole rows with 1-D bulk copies --
// global -> shared on an mbarrier, shared -> global as bulk-group stores -- so the load of row i + 1 and the store of row
// i - 1 are in flight while the 31 compute warps make their three passes over row i. The one-CTA-per-row kernel keeps
// its load, compute and store phases apart (two co-resident CTAs only partly overlap them: 4.4 TB/s, ncu
// profiles/r02_ncu_hbm_kernels.txt); here HBM sees a continuous read and a continuous write stream.
constexpr int CEP_WARPS = 31;                  // compute warps; warp 31 is the producer
constexpr int CEP_COMPUTE = CEP_WARPS * 32;

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ float cep_reduce(float v, float* red, bool is_max) {      // over the 992 compute threads
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  named_bar_sync(1, CEP_COMPUTE);
  float r = lane < CEP_WARPS ? red[lane] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  named_bar_sync(1, CEP_COMPUTE);
  return r;
}

__global__ void __launch_bounds__(1024, 1)
ce_fwd_bwd_pipe_kernel(const bf16* __restrict__ logits, long long ld, const long long* __restrict__ targets,
                       bf16* __restrict__ dlogits, long long ldd, float* __restrict__ row_loss, float* __restrict__ stats,
                       int rows, int V, long long ignore_index, float grad_scale) {
  extern __shared__ uint4 ce_smem[];
  __shared__ uint64_t full_bar[2], done_bar[2];
  __shared__ float red[32];
  __shared__ float s_xt;
  const int nvec = (V + 7) / 8;
  const uint32_t row_bytes = (uint32_t)nvec * 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
    mbar_init(&done_bar[0], CEP_WARPS); mbar_init(&done_bar[1], CEP_WARPS);
    fence_mbar_init();
  }
  __syncthreads();
  const int stride = gridDim.x;
  if (warp == CEP_WARPS) {
    if (lane != 0) return;
    for (int it = 0; it < 2; ++it) {
      const long long row = blockIdx.x + (long long)it * stride;
      if (row < rows) {
        mbar_expect_tx(&full_bar[it], row_bytes);
        bulk_load_1d(ce_smem + (size_t)it * nvec, logits + row * ld, row_bytes, &full_bar[it]);
      }
    }
    for (int it = 0;; ++it) {
      const long long row = blockIdx.x + (long long)it * stride;
      if (row >= rows) break;
      const int b = it & 1;
      mbar_wait(&done_bar[b], (it >> 1) & 1);      // the gradient of row `it` is in buffer b (writers fenced to the async proxy)
      if (dlogits != nullptr) {
        bulk_store_1d(dlogits + row * ldd, ce_smem + (size_t)b * nvec, row_bytes);
        tma_commit_group();
      }
      const long long next = row + 2LL * stride;
      if (next < rows) {
        if (dlogits != nullptr) tma_wait_group_read<0>();      // the store has read buffer b
        mbar_expect_tx(&full_bar[b], row_bytes);
        bulk_load_1d(ce_smem + (size_t)b * nvec, logits + next * ld, row_bytes, &full_bar[b]);
      }
    }
    if (dlogits != nullptr) tma_wait_group<0>();
    return;
  }
  const float kLog2e = 1.4426950408889634f;
  const float n_valid = stats[0];
  const float inv_n = n_valid > 0.f ? 1.0f / n_valid : 0.f;
  const float gs = grad_scale * inv_n;
  for (int it = 0;; ++it) {
    const long long row = blockIdx.x + (long long)it * stride;
    if (row >= rows) break;
    const int b = it & 1;
    uint4* buf = ce_smem + (size_t)b * nvec;
    const long long tgt = targets[row];
    mbar_wait(&full_bar[b], (it >> 1) & 1);
    if (tgt == ignore_index) {
      if (dlogits != nullptr)
        for (int i = threadIdx.x; i < nvec; i += CEP_COMPUTE) buf[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_loss && threadIdx.x == 0) row_loss[row] = 0.f;
    } else {
      // The kernel is bound by instruction issue (ncu: 62 % of the issue slots, eligible warps waiting to be selected), so
      // the three passes over the staged row are written for instruction count: four vectors per trip (predicated, four
      // shared-memory loads in flight), the row maximum with packed bf16 max, exp / scale arithmetic on packed fp32 pairs.
      const int tvec = (int)(tgt >> 3), te = (int)(tgt & 7);
      if (threadIdx.x == 0) s_xt = __bfloat162float(reinterpret_cast<const bf16*>(buf + tvec)[te]);      // before anything is rewritten
      if ((V & 7) && threadIdx.x == 32) {      // padding columns of the last vector -> -inf (bf16 0xff80)
        uint4 v = buf[nvec - 1];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        for (int e = (V & 7); e < 8; ++e) {
          const int wi = e >> 1;
          w[wi] = (e & 1) ? ((w[wi] & 0x0000ffffu) | 0xff800000u) : ((w[wi] & 0xffff0000u) | 0x0000ff80u);
        }
        buf[nvec - 1] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      named_bar_sync(1, CEP_COMPUTE);
      __nv_bfloat162 m2 = __float2bfloat162_rn(-INFINITY);
      for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * CEP_COMPUTE) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * CEP_COMPUTE;
          if (i < nvec) {
            const uint4 v = buf[i];
            m2 = __hmax2(m2, __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&v.x), *reinterpret_cast<const __nv_bfloat162*>(&v.y)),
                                     __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&v.z), *reinterpret_cast<const __nv_bfloat162*>(&v.w))));
          }
        }
      }
      float mx = fmaxf(__bfloat162float(m2.x), __bfloat162float(m2.y));
      mx = cep_reduce(mx, red, true);
      const f32x2 k2 = f2_splat(kLog2e), mneg2 = f2_splat(-mx * kLog2e);
      f32x2 sum2 = f2_splat(0.f);
      for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * CEP_COMPUTE) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * CEP_COMPUTE;
          if (i < nvec) {
            const uint4 v = buf[i];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a0, a1;
              f2_unpack(f2_fma(f2_pack(bf16_lo(w[q]), bf16_hi(w[q])), k2, mneg2), a0, a1);
              const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
              sum2 = f2_add(sum2, f2_pack(e0, e1));
              o[q] = pack_bf16(e0, e1);
            }
            if (dlogits != nullptr) buf[i] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      float s_lo, s_hi;
      f2_unpack(sum2, s_lo, s_hi);
      const float sum = cep_reduce(s_lo + s_hi, red, false);      // (its barriers also publish the rewritten row)
      if (threadIdx.x == 0) {
        const float l = logf(sum) + mx - s_xt;
        if (row_loss) row_loss[row] = l;
        atomicAdd(stats + 1, l * inv_n);
      }
      if (dlogits != nullptr) {
        const float c = gs / sum;
        const f32x2 c2 = f2_splat(c);
        for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * CEP_COMPUTE) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * CEP_COMPUTE;
            if (i < nvec) {
              const uint4 v = buf[i];
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
              float p[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) f2_unpack(f2_mul(f2_pack(bf16_lo(w[q]), bf16_hi(w[q])), c2), p[2 * q], p[2 * q + 1]);
              if (i == tvec) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (e == te) p[e] -= gs;
              }
              buf[i] = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
            }
          }
        }
      }
    }
    fence_proxy_async_smem();      // this thread's writes to buffer b -> visible to the bulk store
    __syncwarp();
    if (lane == 0) mbar_arrive(&done_bar[b]);
  }
}

// ------------------------------------------------------------------------------------------------
// global grad norm (deterministic two-stage) + fused AdamW
// ------------------------------------------------------------------------------------------------
constexpr int SUMSQ_BLOCKS = 1184;   // 148 SMs x 8, fixed so that the summation order never changes
constexpr int SUMSQ_THREADS = 256;

__global__ void __launch_bounds__(SUMSQ_THREADS)
sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  __shared__ float red[32];
  const long long n4 = n / 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long i = (long long)blockIdx.x * SUMSQ_THREADS + threadIdx.x; i < n4;
       i += (long long)SUMSQ_BLOCKS * SUMSQ_THREADS) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    a0 = fmaf(v, v, a0);
  }
  const float s = block_reduce_sum((a0 + a1) + (a2 + a3), red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = sum of squares, out[1] = norm, out[2] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 if max_norm <= 0),
// out[3] += 1 when the norm is finite: the number of optimizer updates really applied (kept by the caller across steps)
__global__ void sumsq_final_kernel(const float* __restrict__ partial, float* __restrict__ out, float max_norm,
                                   float pre_scale) {
  __shared__ double red[SUMSQ_THREADS];
  double a = 0.0;
  for (int i = threadIdx.x; i < SUMSQ_BLOCKS; i += SUMSQ_THREADS) a += (double)partial[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = SUMSQ_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double ss = red[0] * (double)pre_scale * (double)pre_scale;
    const float norm = (float)sqrt(ss);
    out[0] = (float)ss;
    out[1] = norm;
    out[2] = (max_norm > 0.f) ? fminf(1.0f, max_norm / (norm + 1e-6f)) : 1.0f;
    if (isfinite(out[0])) out[3] += 1.0f;
  }
}

struct AdamSeg {          // per parameter tensor (arena offsets are multiples of 4 elements)
  long long end;          // one past the last arena element of the tensor
  float lr_scale;
  float weight_decay;
};

// p, g, m, v: fp32 arenas. Writes the bf16 shadow of p and (optionally) zeroes g: 4 reads + 4.5 writes per element.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, long long n, const AdamSeg* __restrict__ segs, int nseg,
             const float* __restrict__ norm_stats, float grad_scale, float lr, float beta1, float beta2, float eps,
             float bc1, float bc2, int zero_grad) {
  float coef = grad_scale;
  const long long n4 = n / 4;
  if (norm_stats) {
    if (!isfinite(norm_stats[0])) {
      // GradScaler semantics (timm NativeScaler, task_cruller_pretrain.py:259-268): a non-finite gradient skips the
      // parameter / moment update. The gradients are still cleared -- the reference calls optimizer.zero_grad() right
      // after (:295) -- otherwise the NaN / Inf would stay in the accumulate-only arena and poison every later step.
      if (zero_grad)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
          reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      return;
    }
    coef *= norm_stats[2];
    // bias corrections from the number of updates really applied (skipped steps do not count, as in torch.optim.AdamW
    // whose state['step'] only advances inside step())
    const double st = (double)norm_stats[3];
    bc1 = (float)(1.0 - pow((double)beta1, st));
    bc2 = (float)(1.0 - pow((double)beta2, st));
  }
  const float inv_bc1 = 1.0f / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e0 = i * 4;
    // binary search the segment (tensor) containing element e0
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (e0 < segs[mid].end) hi = mid; else lo = mid + 1;
    }
    const float lr_t = lr * segs[lo].lr_scale;
    const float decay = 1.0f - lr_t * segs[lo].weight_decay;
    const float step = lr_t * inv_bc1;
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = ga[e] * coef;
      pa[e] *= decay;
      ma[e] = beta1 * ma[e] + (1.0f - beta1) * gr;
      va[e] = beta2 * va[e] + (1.0f - beta2) * gr * gr;
      const float denom = sqrtf(va[e]) * inv_sqrt_bc2 + eps;
      pa[e] -= step * (ma[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (p16) reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_bf16(pa[0], pa[1]), pack_bf16(pa[2], pa[3]));
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ce_prepare(const long long* targets, int n, long long ignore_index, float* stats, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(targets && stats && n > 0, "b200_ce_prepare: bad arguments");
  ce_prepare_kernel<<<1, 1024, 0, s>>>(targets, n, ignore_index, stats);
  B200_CHECK_LAUNCH("ce_prepare");
  return 0;
}

extern "C" int b200_ce_fwd_bwd(const void* logits_bf16, long long ld, const long long* targets, void* dlogits_bf16,
                               long long ldd, float* row_loss, float* stats, int rows, int vocab,
                               long long ignore_index, float grad_scale, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(logits_bf16 && targets && stats && rows > 0 && vocab > 0, "b200_ce_fwd_bwd: bad arguments");
  B200_CHECK_ARG(ld % 8 == 0 && ld >= (vocab + 7) / 8 * 8, "b200_ce_fwd_bwd: ld must be a multiple of 8 and >= vocab rounded up to 8");
  if (dlogits_bf16) B200_CHECK_ARG(ldd % 8 == 0 && ldd >= (vocab + 7) / 8 * 8, "b200_ce_fwd_bwd: bad ldd");
  const int smem = (vocab + 7) / 8 * 16;
  B200_CHECK_ARG(smem <= 200 * 1024, "b200_ce_fwd_bwd: vocab %d too large for the single-pass kernel", vocab);
  static int pipe_mode = -1;      // PIXPARSE_B200_CE_PIPE=0: the one-CTA-per-row kernel (A/B)
  if (pipe_mode < 0) {
    const char* env = getenv("PIXPARSE_B200_CE_PIPE");
    pipe_mode = (env != nullptr && env[0] == '0') ? 0 : 1;
  }
  if (pipe_mode == 1 && 2 * smem <= 220 * 1024) {
    static int configured_pipe = 0;
    if (2 * smem > configured_pipe) {
      cudaError_t e = cudaFuncSetAttribute(ce_fwd_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * smem);
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(ce pipe)");
      configured_pipe = 2 * smem;
    }
    const int grid = rows < num_sms() ? rows : num_sms();
    ce_fwd_bwd_pipe_kernel<<<grid, 1024, 2 * smem, s>>>(reinterpret_cast<const bf16*>(logits_bf16), ld, targets,
                                                        reinterpret_cast<bf16*>(dlogits_bf16), ldd, row_loss, stats, rows,
                                                        vocab, ignore_index, grad_scale);
    B200_CHECK_LAUNCH("ce_fwd_bwd_pipe");
    return 0;
  }
  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(ce_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(ce)");
    configured_smem = smem;
  }
  ce_fwd_bwd_kernel<<<rows, CE_THREADS, smem, s>>>(reinterpret_cast<const bf16*>(logits_bf16), ld, targets,
                                                  reinterpret_cast<bf16*>(dlogits_bf16), ldd, row_loss, stats, vocab,
                                                  ignore_index, grad_scale);
  B200_CHECK_LAUNCH("ce_fwd_bwd");
  return 0;
}

extern "C" int b200_grad_norm(const float* grads, long long n, float* workspace, float* out4, float max_norm,
                              float pre_scale, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(grads && workspace && out4 && n > 0, "b200_grad_norm: bad arguments");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "b200_grad_norm: grads must be 16-byte aligned");
  sumsq_partial_kernel<<<SUMSQ_BLOCKS, SUMSQ_THREADS, 0, s>>>(grads, n, workspace);
  B200_CHECK_LAUNCH("sumsq_partial");
  sumsq_final_kernel<<<1, SUMSQ_THREADS, 0, s>>>(workspace, out4, max_norm, pre_scale);
  B200_CHECK_LAUNCH("sumsq_final");
  return 0;
}

extern "C" int b200_grad_norm_workspace_floats(void) { return SUMSQ_BLOCKS; }

extern "C" int b200_adamw_step(const B200AdamWArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200AdamWArgs, "b200_adamw_step");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->segments && a->num_segments > 0 && a->n > 0,
                 "b200_adamw_step: bad arguments");
  B200_CHECK_ARG(a->norm_stats != nullptr || a->step > 0, "b200_adamw_step: step must be >= 1 when norm_stats is NULL");
  B200_CHECK_ARG(a->n % 4 == 0, "b200_adamw_step: arena length must be a multiple of 4");
  const int step = a->step > 0 ? a->step : 1;      // (ignored when norm_stats carries the device-side update counter)
  const float bc1 = (float)(1.0 - pow((double)a->beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)a->beta2, (double)step));
  int grid = num_sms() * 8;
  adamw_kernel<<<grid, 256, 0, s>>>(a->params, a->grads, a->exp_avg, a->exp_avg_sq, reinterpret_cast<bf16*>(a->params_bf16),
                                    a->n, reinterpret_cast<const AdamSeg*>(a->segments), a->num_segments, a->norm_stats,
                                    a->grad_scale, a->lr, a->beta1, a->beta2, a->eps, bc1, bc2, a->zero_grad);
  B200_CHECK_LAUNCH("adamw_step");
  return 0;
}
