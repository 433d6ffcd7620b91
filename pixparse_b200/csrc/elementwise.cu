// HBM-bound helper kernels of the Cruller train step (coalesced, vectorised, warp-level reductions):
//   column sums (bias gradients), patch unfold / token assembly (ViT stem), token+position embedding
//   gather and its scatter-add backward, fp32 -> bf16 casts.
#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// column sums: out[n] += sum_m dy[m, n]      (bias gradients of every Linear: SURVEY 2.3 K4-K11 backward)
// block = 256 threads = 32 column-groups (8 bf16 = 16 B each) x 8 row lanes -> 256 columns per block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ dy, long long ld, int rows, int cols, float* __restrict__ out,
                   int rows_per_block) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (col < cols) {   // cols is a multiple of 8
    const bf16* src = dy + col;
    int r = r0 + rl;
    // four independent 16-byte loads in flight per thread (the rolled loop ran at ~4.5 of ~7 TB/s)
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(src + (long long)(r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16_lo(v[u].x); acc[1] += bf16_hi(v[u].x); acc[2] += bf16_lo(v[u].y); acc[3] += bf16_hi(v[u].y);
        acc[4] += bf16_lo(v[u].z); acc[5] += bf16_hi(v[u].z); acc[6] += bf16_lo(v[u].w); acc[7] += bf16_hi(v[u].w);
      }
    }
    for (; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (long long)r * ld));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cg * 8 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < cols) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][c];
    atomicAdd(out + blockIdx.x * 256 + c, a);
  }
}

// ------------------------------------------------------------------------------------------------
// ViT stem. Conv2d(C, D, P, stride P) == GEMM over unfolded patches (timm PatchEmbed, K1).
//   patches[(b*gh + py)*gw + px][c*P*P + i*P + j] = image[b][c][py*P + i][px*P + j]      (bf16)
// ------------------------------------------------------------------------------------------------
__global__ void patch_unfold_kernel(const float* __restrict__ img, bf16* __restrict__ patches, int B, int C, int H,
                                    int W, int P, int ldp) {
  const int gh = H / P, gw = W / P;
  const int kdim = C * P * P;
  const long long total = (long long)B * gh * gw * kdim;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % kdim);
    const long long prow = idx / kdim;
    const int px = (int)(prow % gw);
    const int py = (int)((prow / gw) % gh);
    const int b = (int)(prow / ((long long)gw * gh));
    const int j = k % P, i = (k / P) % P, c = k / (P * P);
    const float v = img[(((long long)b * C + c) * H + (py * P + i)) * W + (px * P + j)];
    patches[prow * ldp + k] = __float2bfloat16_rn(v);
  }
}

// x[b, 0, :] = cls + pos[0];  x[b, 1 + p, :] = proj[b*np + p, :] + pos[1 + p]     (timm _pos_embed, K2)
__global__ void tokens_assemble_kernel(const bf16* __restrict__ proj, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int S, int D) {
  const int d4 = D / 4;
  const long long total = (long long)B * S * d4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % d4);
    const int s = (int)((idx / d4) % S);
    const int b = (int)(idx / ((long long)d4 * S));
    const float4 pe = reinterpret_cast<const float4*>(pos)[(long long)s * d4 + c];
    float4 v;
    if (s == 0) {
      v = reinterpret_cast<const float4*>(cls)[c];
    } else {
      const uint2 r = reinterpret_cast<const uint2*>(proj)[((long long)b * (S - 1) + (s - 1)) * d4 + c];
      v = make_float4(bf16_lo(r.x), bf16_hi(r.x), bf16_lo(r.y), bf16_hi(r.y));
    }
    reinterpret_cast<float4*>(x)[idx] = make_float4(v.x + pe.x, v.y + pe.y, v.z + pe.z, v.w + pe.w);
  }
}

// backward of the assembly: dproj = bf16(dx[:, 1:, :]); dpos[s] += sum_b dx[b, s]; dcls += sum_b dx[b, 0]
// one thread per (s, 4 columns), loops over the batch -> no atomics needed
__global__ void tokens_assemble_bwd_kernel(const float* __restrict__ dx, bf16* __restrict__ dproj,
                                           float* __restrict__ dcls, float* __restrict__ dpos, int B, int S, int D) {
  const int d4 = D / 4;
  const long long total = (long long)S * d4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % d4);
    const int s = (int)(idx / d4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 g = reinterpret_cast<const float4*>(dx)[((long long)b * S + s) * d4 + c];
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      if (s > 0)
        reinterpret_cast<uint2*>(dproj)[((long long)b * (S - 1) + (s - 1)) * d4 + c] =
            make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
    }
    float4* dp = reinterpret_cast<float4*>(dpos) + (long long)s * d4 + c;
    float4 o = *dp;
    *dp = make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w);
    if (s == 0) {
      float4* dc = reinterpret_cast<float4*>(dcls) + c;
      float4 oc = *dc;
      *dc = make_float4(oc.x + acc.x, oc.y + acc.y, oc.z + acc.z, oc.w + acc.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BART decoder embedding (K8): x[b, t, :] = embed_tokens[ids[b, t]] * embed_scale + embed_positions[t + offset]
// ------------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ tok_emb,
                                 const float* __restrict__ pos_emb, float* __restrict__ x, int B, int T, int D,
                                 int pos_offset, float scale) {
  const int d4 = D / 4;
  const long long total = (long long)B * T * d4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % d4);
    const long long bt = idx / d4;
    const int t = (int)(bt % T);
    const long long id = ids[bt];
    const float4 e = reinterpret_cast<const float4*>(tok_emb)[id * d4 + c];
    const float4 p = reinterpret_cast<const float4*>(pos_emb)[(long long)(t + pos_offset) * d4 + c];
    reinterpret_cast<float4*>(x)[idx] =
        make_float4(e.x * scale + p.x, e.y * scale + p.y, e.z * scale + p.z, e.w * scale + p.w);
  }
}

// d_tok_emb[ids[b,t]] += dx[b,t] * scale (skipped for padding_idx, as nn.Embedding(padding_idx=1) does);
// d_pos_emb[t + offset] += sum_b dx[b, t]
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dx,
                                 float* __restrict__ d_tok, float* __restrict__ d_pos, int B, int T, int D,
                                 int pos_offset, float scale, long long padding_idx) {
  const long long total = (long long)B * T * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long bt = idx / D;
    const int t = (int)(bt % T);
    const long long id = ids[bt];
    const float g = dx[idx];
    if (id != padding_idx) atomicAdd(d_tok + id * D + c, g * scale);
    atomicAdd(d_pos + (long long)(t + pos_offset) * D + c, g);
  }
}

// ------------------------------------------------------------------------------------------------
// casts
// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

static inline int grid_for(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_colsum_bf16(const void* dy, long long ld, int rows, int cols, float* out, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(dy && out && rows > 0 && cols > 0, "b200_colsum_bf16: bad arguments");
  B200_CHECK_ARG(cols % 8 == 0 && ld % 8 == 0, "b200_colsum_bf16: cols and ld must be multiples of 8");
  const int col_blocks = (cols + 255) / 256;
  int row_blocks = (num_sms() * 8 + col_blocks - 1) / col_blocks;
  int rows_per_block = (rows + row_blocks - 1) / row_blocks;
  rows_per_block = (rows_per_block + 7) / 8 * 8;
  row_blocks = (rows + rows_per_block - 1) / rows_per_block;
  colsum_bf16_kernel<<<dim3(col_blocks, row_blocks), 256, 0, s>>>(reinterpret_cast<const bf16*>(dy), ld, rows, cols,
                                                                  out, rows_per_block);
  B200_CHECK_LAUNCH("colsum_bf16");
  return 0;
}

extern "C" int b200_patch_unfold(const float* image, void* patches_bf16, int B, int C, int H, int W, int P,
                                 long long ld_patches, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(image && patches_bf16 && H % P == 0 && W % P == 0, "b200_patch_unfold: H, W must be multiples of P");
  const long long total = (long long)B * (H / P) * (W / P) * C * P * P;
  patch_unfold_kernel<<<grid_for(total, 256), 256, 0, s>>>(image, reinterpret_cast<bf16*>(patches_bf16), B, C, H, W,
                                                           P, (int)ld_patches);
  B200_CHECK_LAUNCH("patch_unfold");
  return 0;
}

extern "C" int b200_tokens_assemble(const void* proj_bf16, const float* cls, const float* pos, float* x, int B, int S,
                                    int D, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(D % 4 == 0, "b200_tokens_assemble: D must be a multiple of 4");
  tokens_assemble_kernel<<<grid_for((long long)B * S * D / 4, 256), 256, 0, s>>>(
      reinterpret_cast<const bf16*>(proj_bf16), cls, pos, x, B, S, D);
  B200_CHECK_LAUNCH("tokens_assemble");
  return 0;
}

extern "C" int b200_tokens_assemble_bwd(const float* dx, void* dproj_bf16, float* dcls, float* dpos, int B, int S,
                                        int D, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(D % 4 == 0, "b200_tokens_assemble_bwd: D must be a multiple of 4");
  tokens_assemble_bwd_kernel<<<grid_for((long long)S * D / 4, 128), 128, 0, s>>>(
      dx, reinterpret_cast<bf16*>(dproj_bf16), dcls, dpos, B, S, D);
  B200_CHECK_LAUNCH("tokens_assemble_bwd");
  return 0;
}

extern "C" int b200_embed_fwd(const long long* ids, const float* tok_emb, const float* pos_emb, float* x, int B,
                              int T, int D, int pos_offset, float scale, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(D % 4 == 0, "b200_embed_fwd: D must be a multiple of 4");
  embed_fwd_kernel<<<grid_for((long long)B * T * D / 4, 256), 256, 0, s>>>(ids, tok_emb, pos_emb, x, B, T, D,
                                                                           pos_offset, scale);
  B200_CHECK_LAUNCH("embed_fwd");
  return 0;
}

extern "C" int b200_embed_bwd(const long long* ids, const float* dx, float* d_tok_emb, float* d_pos_emb, int B, int T,
                              int D, int pos_offset, float scale, long long padding_idx, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  embed_bwd_kernel<<<grid_for((long long)B * T * D, 256), 256, 0, s>>>(ids, dx, d_tok_emb, d_pos_emb, B, T, D,
                                                                       pos_offset, scale, padding_idx);
  B200_CHECK_LAUNCH("embed_bwd");
  return 0;
}

// dst[i] = (dst[i] + sum_{s != skip} stage[s * stride + i]) * scale: the reduction step of the copy-engine gradient exchange
// (pixparse_b200/reducer.py): this rank's share of a gradient bucket plus the copies its peers pushed into the staging area
namespace b200 {
__global__ void __launch_bounds__(256)
reduce_shards_kernel(float* __restrict__ dst, const float* __restrict__ stage, long long stride, int nsrc, int skip,
                     float scale, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(dst)[i];
    for (int s = 0; s < nsrc; ++s) {
      if (s == skip) continue;
      const float4 b = __ldcs(reinterpret_cast<const float4*>(stage + s * stride) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float a = dst[i];
    for (int s = 0; s < nsrc; ++s)
      if (s != skip) a += stage[s * stride + i];
    dst[i] = a * scale;
  }
}
}  // namespace b200

extern "C" int b200_reduce_shards(float* dst, const float* stage, long long stride, int nsrc, int skip, float scale,
                                  long long n, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(n >= 0 && nsrc >= 1 && stride >= 0, "b200_reduce_shards: bad arguments");
  if (n == 0) return 0;
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(stage)) & 15) == 0 && stride % 4 == 0,
                 "b200_reduce_shards: pointers must be 16-byte aligned and the stride a multiple of 4 elements");
  // a small grid: the kernel runs between persistent GEMM / attention kernels that own every SM
  const long long want = (n / 4 + 255) / 256;
  const int grid = (int)(want < 1 ? 1 : (want > 2LL * b200::num_sms() ? 2LL * b200::num_sms() : want));
  b200::reduce_shards_kernel<<<grid, 256, 0, s>>>(dst, stage, stride, nsrc, skip, scale, n);
  B200_CHECK_LAUNCH("reduce_shards");
  return 0;
}

extern "C" int b200_cast_f32_bf16(const float* src, void* dst_bf16, long long n, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(n >= 0, "b200_cast_f32_bf16: negative size");
  if (n == 0) return 0;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 7) == 0,
                 "b200_cast_f32_bf16: misaligned pointers");
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, s>>>(src, reinterpret_cast<bf16*>(dst_bf16), n);
  B200_CHECK_LAUNCH("cast_f32_bf16");
  return 0;
}
