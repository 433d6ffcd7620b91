// Microbenchmark: cycles per tcgen05.mma for the shapes / operand layouts the attention kernels use.
#include <cstdio>
#include "common.cuh"
using namespace b200;

struct Cfg { int N; int a_mn; int b_mn; int a_tmem; int n_acc; int kadv_a; int kadv_b; int lbo; };

__global__ void __launch_bounds__(128, 1) bench_kernel(Cfg c, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32 && (c.lbo & 1) == 0) {
    // whole-warp uniform loop; a single elected lane issues (CUTLASS / DeepGEMM pattern)
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    const uint64_t da = make_smem_desc(smem_u32(smem), c.a_mn ? c.lbo : 16, 1024);
    const uint64_t db = make_smem_desc(smem_u32(smem + 32768), c.b_mn ? c.lbo : 16, 1024);
    if (elect_one()) {
      for (int i = 0; i < 8; ++i) {
        if (c.a_tmem) umma_ts(tm + 256, tm + 448, db, idesc, 1); else umma_ss(tm + 256, da, db, idesc, 1);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i0 = 0; i0 < iters; i0 += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tm + (uint32_t)(((i0 + k) & (c.n_acc - 1)) * c.N);
          if (c.a_tmem) umma_ts(d, tm + 448 + 8 * k, db + (uint64_t)(c.kadv_b * k), idesc, 1);
          else umma_ss(d, da + (uint64_t)(c.kadv_a * k), db + (uint64_t)(c.kadv_b * k), idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 1);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  if (threadIdx.x == 0 && (c.lbo & 1) == 1) {
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    const uint64_t da = make_smem_desc(smem_u32(smem), c.a_mn ? c.lbo : 16, 1024);
    const uint64_t db = make_smem_desc(smem_u32(smem + 32768), c.b_mn ? c.lbo : 16, 1024);
    // warm-up
    for (int i = 0; i < 8; ++i) {
      if (c.a_tmem) umma_ts(tm + 256, tm + 448, db, idesc, 1); else umma_ss(tm + 256, da, db, idesc, 1);
    }
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 7;
      const uint32_t d = tm + (uint32_t)((i % c.n_acc) * c.N);
      if (c.a_tmem) umma_ts(d, tm + 448 + 8 * (k & 3), db + (uint64_t)(c.kadv_b * (k & 3)), idesc, 1);
      else umma_ss(d, da + (uint64_t)(c.kadv_a * (k & 3)), db + (uint64_t)(c.kadv_b * (k & 3)), idesc, 1);
    }
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct { const char* name; Cfg c; } tests[] = {
    {"N256 KxK  1acc (gemm fwd)      ", {256, 0, 0, 0, 1, 2, 2, 8192}},
    {"N256 KxMN 1acc (gemm dgrad)    ", {256, 0, 1, 0, 1, 2, 128, 8192}},
    {"N256 MNxMN 1acc (gemm wgrad)   ", {256, 1, 1, 0, 1, 128, 128, 8192}},
    {"N128 KxK  1acc (S, dP)         ", {128, 0, 0, 0, 1, 2, 2, 16384}},
    {"N128 KxK  2acc                 ", {128, 0, 0, 0, 2, 2, 2, 16384}},
    {"N64  KxK  1acc (fwd S, BN=64)  ", {64, 0, 0, 0, 1, 2, 2, 16384}},
    {"N64  KxMN 1acc (dQ)            ", {64, 0, 1, 0, 1, 2, 128, 16384}},
    {"N64  MNxMN 1acc (dV, dK)       ", {64, 1, 1, 0, 1, 128, 128, 16384}},
    {"N64  MNxMN 2acc                ", {64, 1, 1, 0, 2, 128, 128, 16384}},
    {"N64  MNxMN 4acc                ", {64, 1, 1, 0, 4, 128, 128, 16384}},
    {"N64  TMEM-A x MN 1acc (fwd PV) ", {64, 0, 1, 1, 1, 0, 128, 16384}},
    {"N64  TMEM-A x MN 2acc          ", {64, 0, 1, 1, 2, 0, 128, 16384}},
    {"N128 MNxMN 1acc                ", {128, 1, 1, 0, 1, 128, 128, 16384}},
    {"N128 KxMN 1acc                 ", {128, 0, 1, 0, 1, 2, 128, 16384}},
  };
  const int iters = 256;
  for (auto& t : tests) {
    bench_kernel<<<1, 128, 100 * 1024>>>(t.c, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double ideal = 128.0 * t.c.N / 256.0;
    printf("%s : %7.1f cycles/MMA (ideal %5.1f)  %s\n", t.name, (double)h / iters, ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  // all SMs busy variant: same kernel on 148 CTAs to expose shared limits
  for (auto& t : tests) {
    bench_kernel<<<148, 128, 100 * 1024>>>(t.c, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("[148 CTAs] %s : %7.1f cycles/MMA  %s\n", t.name, (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
