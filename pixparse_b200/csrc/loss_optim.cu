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
