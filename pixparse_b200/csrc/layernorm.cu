// LayerNorm forward / backward, HBM-bound: one warp per row, the row lives in registers, float4 / 8-byte
// vector accesses, warp-shuffle reductions. Statistics are fp32 (torch autocast keeps layer_norm in fp32).
//
// Replaces at::native layer_norm fwd/bwd reached from timm Block.norm1/norm2/norm (models/image_encoder_timm.py:13-20)
// and BART layernorm_embedding / *_layer_norm (models/text_decoder_hf.py:13-33)  -- SURVEY.md 2.3 K3.
#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

constexpr int LN_WARPS = 4;

// NV = number of float4 per lane: D = NV * 128
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     bf16* __restrict__ y16, float* __restrict__ y32, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int rows, float eps, uint32_t drop_thr, uint32_t drop_seed) {
  constexpr int D = NV * 128;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i);
    const float4 b = __ldg(b4 + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (drop_thr != 0u) {      // dropout on the normalised output (BART: dropout(layernorm_embedding(...)))
      const uint32_t pair = (uint32_t)(((size_t)row * D + (size_t)(lane + 32 * i) * 4) >> 1);
      const float sc = dropout_scale(drop_thr);
      dropout_pair(drop_seed, pair, drop_thr, sc, o.x, o.y);
      dropout_pair(drop_seed, pair + 1, drop_thr, sc, o.z, o.w);
    }
    if (y32) reinterpret_cast<float4*>(y32 + (size_t)row * D)[lane + 32 * i] = o;
    if (y16)
      reinterpret_cast<uint2*>(y16 + (size_t)row * D)[lane + 32 * i] =
          make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
  }
}

// dy = (dy16 ? dy16 : 0) + (dy32 ? dy32 : 0);   dx = LNbwd(dy) + (dres32 ? dres32 : 0)
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (fp32 atomics, one set per block)
template <int NV, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const bf16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ dres32,
                     const float* __restrict__ x, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, const float* __restrict__ gamma, float* __restrict__ dx32,
                     bf16* __restrict__ dx16, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
                     uint32_t in_thr, uint32_t in_seed, uint32_t out_thr, uint32_t out_seed) {
  constexpr int D = NV * 128;
  __shared__ float red[LN_WARPS][D];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);

  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const float mean = mean_in[row];
    const float rstd = rstd_in[row];
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = reinterpret_cast<const float4*>(x + (size_t)row * D)[lane + 32 * i];
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dy16) {
        const uint2 r = reinterpret_cast<const uint2*>(dy16 + (size_t)row * D)[lane + 32 * i];
        d = make_float4(bf16_lo(r.x), bf16_hi(r.x), bf16_lo(r.y), bf16_hi(r.y));
      }
      if (dy32) {
        const float4 r = reinterpret_cast<const float4*>(dy32 + (size_t)row * D)[lane + 32 * i];
        d.x += r.x; d.y += r.y; d.z += r.z; d.w += r.w;
      }
      if (DROP && in_thr != 0u) {      // the forward dropped this LayerNorm's OUTPUT: mask the incoming gradient the same way
        const uint32_t pair = (uint32_t)(((size_t)row * D + (size_t)(lane + 32 * i) * 4) >> 1);
        const float sc = dropout_scale(in_thr);
        dropout_pair(in_seed, pair, in_thr, sc, d.x, d.y);
        dropout_pair(in_seed, pair + 1, in_thr, sc, d.z, d.w);
      }
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    const float m1 = warp_sum(s1) * (1.0f / D);
    const float m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (g[i].x - m1 - xh[i].x * m2);
      o.y = rstd * (g[i].y - m1 - xh[i].y * m2);
      o.z = rstd * (g[i].z - m1 - xh[i].z * m2);
      o.w = rstd * (g[i].w - m1 - xh[i].w * m2);
      if (dres32) {
        const float4 r = reinterpret_cast<const float4*>(dres32 + (size_t)row * D)[lane + 32 * i];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (dx32) reinterpret_cast<float4*>(dx32 + (size_t)row * D)[lane + 32 * i] = o;
      if (dx16) {
        if (DROP && out_thr != 0u) {   // dx16 feeds the sub-layer whose output was dropped before the residual add
          const uint32_t pair = (uint32_t)(((size_t)row * D + (size_t)(lane + 32 * i) * 4) >> 1);
          const float sc = dropout_scale(out_thr);
          dropout_pair(out_seed, pair, out_thr, sc, o.x, o.y);
          dropout_pair(out_seed, pair + 1, out_thr, sc, o.z, o.w);
        }
        reinterpret_cast<uint2*>(dx16 + (size_t)row * D)[lane + 32 * i] =
            make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
  // block reduction of the parameter gradients, then one atomic per column per block
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i)
      reinterpret_cast<float4*>(&red[warp][0])[lane + 32 * i] = pass == 0 ? dg[i] : db[i];
    __syncthreads();
    float* dst = pass == 0 ? dgamma : dbeta;
    for (int c = threadIdx.x; c < D; c += LN_WARPS * 32) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) a += red[w][c];
      atomicAdd(dst + c, a);
    }
  }
}

template <int NV>
static int ln_fwd_launch(const float* x, const float* gamma, const float* beta, bf16* y16, float* y32, float* mean,
                         float* rstd, int rows, float eps, uint32_t thr, uint32_t seed, cudaStream_t s) {
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  cudaError_t e = launch_kernel(layernorm_fwd_kernel<NV>, dim3(grid), dim3(LN_WARPS * 32), 0, s, 1, x, gamma, beta, y16, y32,
                                mean, rstd, rows, eps, thr, seed);
  if (e != cudaSuccess) return check_cuda(e, "layernorm_fwd launch");
  B200_CHECK_LAUNCH("layernorm_fwd");
  return 0;
}

template <int NV>
static int ln_bwd_launch(const bf16* dy16, const float* dy32, const float* dres32, const float* x, const float* mean,
                         const float* rstd, const float* gamma, float* dx32, bf16* dx16, float* dgamma, float* dbeta,
                         int rows, uint32_t in_thr, uint32_t in_seed, uint32_t out_thr, uint32_t out_seed,
                         cudaStream_t s) {
  // exactly one wave of resident CTAs: every CTA strides over the rows, so a partial second wave (the former fixed
  // 4 CTAs per SM against 3 that fit) ran at a quarter of the memory parallelism for a third of the rows
  static int per_sm[2] = {0, 0};
  const bool drop = in_thr != 0u || out_thr != 0u;
  if (per_sm[drop] == 0) {
    int n = 0;
    cudaError_t e = drop ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, layernorm_bwd_kernel<NV, true>, LN_WARPS * 32, 0)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, layernorm_bwd_kernel<NV, false>, LN_WARPS * 32, 0);
    if (e != cudaSuccess) return check_cuda(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(layernorm_bwd)");
    per_sm[drop] = n > 0 ? n : 1;
  }
  int grid = num_sms() * per_sm[drop];
  const int need = (rows + LN_WARPS - 1) / LN_WARPS;
  if (grid > need) grid = need;
  cudaError_t le;
  if (in_thr != 0u || out_thr != 0u)
    le = launch_kernel(layernorm_bwd_kernel<NV, true>, dim3(grid), dim3(LN_WARPS * 32), 0, s, 1, dy16, dy32, dres32, x, mean,
                       rstd, gamma, dx32, dx16, dgamma, dbeta, rows, in_thr, in_seed, out_thr, out_seed);
  else      // the common (encoder) case carries no dropout code at all
    le = launch_kernel(layernorm_bwd_kernel<NV, false>, dim3(grid), dim3(LN_WARPS * 32), 0, s, 1, dy16, dy32, dres32, x, mean,
                       rstd, gamma, dx32, dx16, dgamma, dbeta, rows, 0u, 0u, 0u, 0u);
  if (le != cudaSuccess) return check_cuda(le, "layernorm_bwd launch");
  B200_CHECK_LAUNCH("layernorm_bwd");
  return 0;
}

}  // namespace b200

using namespace b200;

static inline uint32_t thr16(float p) { return p > 0.f ? (uint32_t)(p * 65536.0f + 0.5f) : 0u; }

static int ln_fwd_dispatch(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                           float* mean, float* rstd, int rows, int dim, float eps, float drop_p, unsigned int drop_seed,
                           void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(rows > 0 && x && gamma && beta, "b200_layernorm_fwd: bad arguments");
  B200_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "b200_layernorm_fwd: dropout p must be in [0, 1)");
  bf16* y16 = reinterpret_cast<bf16*>(y_bf16);
  const uint32_t t = thr16(drop_p);
  switch (dim) {
    case 128: return ln_fwd_launch<1>(x, gamma, beta, y16, y_f32, mean, rstd, rows, eps, t, drop_seed, s);
    case 256: return ln_fwd_launch<2>(x, gamma, beta, y16, y_f32, mean, rstd, rows, eps, t, drop_seed, s);
    case 512: return ln_fwd_launch<4>(x, gamma, beta, y16, y_f32, mean, rstd, rows, eps, t, drop_seed, s);
    case 768: return ln_fwd_launch<6>(x, gamma, beta, y16, y_f32, mean, rstd, rows, eps, t, drop_seed, s);
    case 1024: return ln_fwd_launch<8>(x, gamma, beta, y16, y_f32, mean, rstd, rows, eps, t, drop_seed, s);
  }
  set_last_error("b200_layernorm_fwd: unsupported dim %d (supported: 128, 256, 512, 768, 1024)", dim);
  return -1;
}

static int ln_bwd_dispatch(const void* dy_bf16, const float* dy_f32, const float* dres_f32, const float* x,
                           const float* mean, const float* rstd, const float* gamma, float* dx_f32, void* dx_bf16,
                           float* dgamma, float* dbeta, int rows, int dim, float in_p, unsigned int in_seed,
                           float out_p, unsigned int out_seed, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(rows > 0 && x && mean && rstd && gamma && dgamma && dbeta, "b200_layernorm_bwd: bad arguments");
  B200_CHECK_ARG(dy_bf16 || dy_f32, "b200_layernorm_bwd: need dy_bf16 and/or dy_f32");
  const bf16* dy16 = reinterpret_cast<const bf16*>(dy_bf16);
  bf16* dx16 = reinterpret_cast<bf16*>(dx_bf16);
  const uint32_t ti = thr16(in_p), to = thr16(out_p);
  switch (dim) {
    case 128: return ln_bwd_launch<1>(dy16, dy_f32, dres_f32, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, rows, ti, in_seed, to, out_seed, s);
    case 256: return ln_bwd_launch<2>(dy16, dy_f32, dres_f32, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, rows, ti, in_seed, to, out_seed, s);
    case 512: return ln_bwd_launch<4>(dy16, dy_f32, dres_f32, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, rows, ti, in_seed, to, out_seed, s);
    case 768: return ln_bwd_launch<6>(dy16, dy_f32, dres_f32, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, rows, ti, in_seed, to, out_seed, s);
    case 1024: return ln_bwd_launch<8>(dy16, dy_f32, dres_f32, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, rows, ti, in_seed, to, out_seed, s);
  }
  set_last_error("b200_layernorm_bwd: unsupported dim %d (supported: 128, 256, 512, 768, 1024)", dim);
  return -1;
}

extern "C" int b200_layernorm_fwd(const B200LayerNormFwdArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200LayerNormFwdArgs, "b200_layernorm_fwd");
  return ln_fwd_dispatch(a->x, a->gamma, a->beta, a->y_bf16, a->y_f32, a->mean, a->rstd, a->rows, a->dim, a->eps, a->drop_p,
                         a->drop_seed, stream);
}

extern "C" int b200_layernorm_bwd(const B200LayerNormBwdArgs* a, void* stream) {
  B200_CHECK_STRUCT(a, B200LayerNormBwdArgs, "b200_layernorm_bwd");
  return ln_bwd_dispatch(a->dy_bf16, a->dy_f32, a->dres_f32, a->x, a->mean, a->rstd, a->gamma, a->dx_f32, a->dx_bf16,
                         a->dgamma, a->dbeta, a->rows, a->dim, a->in_p, a->in_seed, a->out_p, a->out_seed, stream);
}
