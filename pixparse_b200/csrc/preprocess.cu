// On-device page preprocessing: uint8 grayscale page -> ToTensor (x / 255) -> antialiased separable
// bicubic resize to the model's image size -> (x - mean) / std, streamed in one pass over HBM.
//
// Replaces the CPU pipeline the reference builds in task/task_cruller_pretrain.py:132-143
//   transforms.Compose([ToTensor(), Resize(size, BICUBIC, antialias=True), Normalize(mean, std)])
// i.e. ATen's _upsample_bicubic2d_aa (cubic a = -0.5, support scaled by the down-sampling ratio, weights
// normalised per output pixel). The filter taps are built on the device with the same fp32 formulas ATen uses.
#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

constexpr int PP_MAX_TAPS = 24;     // supports down-scaling ratios up to ~5.5x
constexpr int PP_TILE_W = 64;       // output tile per CTA
constexpr int PP_TILE_H = 16;

__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}

// One thread per output index: first tap, tap count and normalised weights (ATen HelperInterpBase, align_corners=False)
__global__ void resize_taps_kernel(int in_size, int out_size, int* __restrict__ first, int* __restrict__ count,
                                   float* __restrict__ weights) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  const float scale = (float)in_size / (float)out_size;
  const float support = (scale >= 1.0f) ? 2.0f * scale : 2.0f;
  const float invscale = (scale >= 1.0f) ? 1.0f / scale : 1.0f;
  const float center = scale * ((float)i + 0.5f);
  int xmin = (int)(center - support + 0.5f);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5f);
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > PP_MAX_TAPS) n = PP_MAX_TAPS;
  float total = 0.f;
  float w[PP_MAX_TAPS];
  for (int j = 0; j < n; ++j) {
    w[j] = cubic_aa(((float)(j + xmin) - center + 0.5f) * invscale);
    total += w[j];
  }
  const float inv = total != 0.f ? 1.0f / total : 0.f;
  for (int j = 0; j < PP_MAX_TAPS; ++j) weights[(size_t)i * PP_MAX_TAPS + j] = j < n ? w[j] * inv : 0.f;
  first[i] = xmin;
  count[i] = n;
}

// One CTA per (16 x 64 output tile, page): stage the input footprint (uint8) in shared memory, horizontal pass into a
// float strip, vertical pass + normalisation to the output. 256 threads.
__global__ void __launch_bounds__(256)
preprocess_pages_kernel(const uint8_t* __restrict__ pages, int B, int Hin, int Win, long long page_stride,
                        float* __restrict__ out, int Hout, int Wout, const int* __restrict__ yfirst,
                        const int* __restrict__ ycount, const float* __restrict__ yw, const int* __restrict__ xfirst,
                        const int* __restrict__ xcount, const float* __restrict__ xw, float mean, float inv_std,
                        int max_rows, int max_cols) {
  extern __shared__ uint8_t pp_smem[];
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * PP_TILE_W;
  const int oy0 = blockIdx.y * PP_TILE_H;
  const int ox1 = min(Wout, ox0 + PP_TILE_W);
  const int oy1 = min(Hout, oy0 + PP_TILE_H);
  // input footprint of the tile
  const int ix0 = xfirst[ox0];
  const int ix1 = xfirst[ox1 - 1] + xcount[ox1 - 1];
  const int iy0 = yfirst[oy0];
  const int iy1 = yfirst[oy1 - 1] + ycount[oy1 - 1];
  const int cols = ix1 - ix0, rows = iy1 - iy0;
  uint8_t* s_in = pp_smem;                                                  // [rows][max_cols] uint8
  float* s_h = reinterpret_cast<float*>(pp_smem + (((size_t)max_rows * max_cols + 15) & ~(size_t)15));   // [rows][64]
  const uint8_t* src = pages + (long long)b * page_stride;
  for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
    const int r = idx / cols, c = idx - r * cols;
    s_in[r * max_cols + c] = src[(long long)(iy0 + r) * Win + (ix0 + c)];
  }
  __syncthreads();
  // horizontal pass: s_h[r][ox] = sum_j w[ox][j] * in[r][first[ox] + j] / 255
  const int tw = ox1 - ox0;
  for (int idx = threadIdx.x; idx < rows * tw; idx += blockDim.x) {
    const int r = idx / tw, t = idx - r * tw;
    const int ox = ox0 + t;
    const int f = xfirst[ox] - ix0, n = xcount[ox];
    const float* w = xw + (size_t)ox * PP_MAX_TAPS;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += w[j] * ((float)s_in[r * max_cols + f + j] * (1.0f / 255.0f));
    s_h[r * PP_TILE_W + t] = acc;
  }
  __syncthreads();
  // vertical pass + Normalize
  for (int idx = threadIdx.x; idx < (oy1 - oy0) * tw; idx += blockDim.x) {
    const int ty = idx / tw, t = idx - ty * tw;
    const int oy = oy0 + ty;
    const int f = yfirst[oy] - iy0, n = ycount[oy];
    const float* w = yw + (size_t)oy * PP_MAX_TAPS;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += w[j] * s_h[(f + j) * PP_TILE_W + t];
    out[((long long)b * Hout + oy) * Wout + ox0 + t] = (acc - mean) * inv_std;
  }
}

}  // namespace b200

using namespace b200;

extern "C" long long b200_preprocess_workspace_bytes(int Hout, int Wout) {
  // per axis: first[int], count[int], weights[float * PP_MAX_TAPS]
  return (long long)(Hout + Wout) * (2 * 4 + PP_MAX_TAPS * 4);
}

extern "C" int b200_preprocess_pages(const void* pages_u8, int B, int Hin, int Win, long long page_stride,
                                     float* out, int Hout, int Wout, float mean, float std_, void* workspace,
                                     void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(pages_u8 && out && workspace && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && std_ != 0.f,
                 "b200_preprocess_pages: bad arguments");
  const float sy = (float)Hin / Hout, sx = (float)Win / Wout;
  const int taps_y = (int)(2 * (sy >= 1.f ? 2.f * sy : 2.f)) + 2, taps_x = (int)(2 * (sx >= 1.f ? 2.f * sx : 2.f)) + 2;
  B200_CHECK_ARG(taps_y <= PP_MAX_TAPS && taps_x <= PP_MAX_TAPS,
                 "b200_preprocess_pages: down-scaling ratio too large (%dx%d -> %dx%d)", Hin, Win, Hout, Wout);
  char* ws = reinterpret_cast<char*>(workspace);
  int* yfirst = reinterpret_cast<int*>(ws);
  int* ycount = yfirst + Hout;
  int* xfirst = ycount + Hout;
  int* xcount = xfirst + Wout;
  float* yw = reinterpret_cast<float*>(xcount + Wout);
  float* xw = yw + (size_t)Hout * PP_MAX_TAPS;
  resize_taps_kernel<<<(Hout + 127) / 128, 128, 0, s>>>(Hin, Hout, yfirst, ycount, yw);
  B200_CHECK_LAUNCH("resize_taps(y)");
  resize_taps_kernel<<<(Wout + 127) / 128, 128, 0, s>>>(Win, Wout, xfirst, xcount, xw);
  B200_CHECK_LAUNCH("resize_taps(x)");
  // shared-memory footprint bounds of one output tile
  const int max_rows = (int)(PP_TILE_H * (sy > 1.f ? sy : 1.f)) + taps_y + 2;
  const int max_cols = ((int)(PP_TILE_W * (sx > 1.f ? sx : 1.f)) + taps_x + 2 + 15) & ~15;
  const size_t smem = (((size_t)max_rows * max_cols + 15) & ~(size_t)15) + (size_t)max_rows * PP_TILE_W * 4;
  B200_CHECK_ARG(smem <= 200 * 1024, "b200_preprocess_pages: tile footprint too large");
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(preprocess_pages_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(preprocess)");
    configured = smem;
  }
  dim3 grid((Wout + PP_TILE_W - 1) / PP_TILE_W, (Hout + PP_TILE_H - 1) / PP_TILE_H, B);
  preprocess_pages_kernel<<<grid, 256, smem, s>>>(reinterpret_cast<const uint8_t*>(pages_u8), B, Hin, Win, page_stride,
                                                  out, Hout, Wout, yfirst, ycount, yw, xfirst, xcount, xw, mean,
                                                  1.0f / std_, max_rows, max_cols);
  B200_CHECK_LAUNCH("preprocess_pages");
  return 0;
}
