// Microbenchmark: tensor-memory read bandwidth of tcgen05.ld per SM, by warp count and vector width.
#include <cstdio>
#include "common.cuh"
using namespace b200;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, "
      "%46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// other tcgen05.ld shapes at the same 4 KB per instruction (32 registers per thread): 16 lanes x 256 bit x 8 and
// 16 lanes x 128 bit x 16 -- does the ~60 B/clk per SM depend on the access shape?
#define LD32_OUT                                                                                                       \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),        \
      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define LD32_REGS                                                                                            \
  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                  \
  "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " LD32_REGS : LD32_OUT : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x128b_x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " LD32_REGS : LD32_OUT : "r"(taddr) : "memory");
}

// mode 4: 16x256b.x8 + wait each; mode 5: 16x128b.x16 + wait each
// mode 0: x32 ld + wait each; mode 1: two x32 then wait; mode 2: x64 + wait; mode 3: x32 ld, wait deferred by one (pipelined)
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld_32x32(tm + ((i * 32) & 255) + (warp >> 2) * 0, r);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 31]);
    }
  } else if (mode == 1) {
    for (int i = 0; i < iters; i += 2) {
      uint32_t a[32], b[32];
      tmem_ld_32x32(tm + ((i * 32) & 255), a);
      tmem_ld_32x32(tm + ((i * 32 + 32) & 255), b);
      tmem_ld_wait();
      acc += __uint_as_float(a[i & 31]) + __uint_as_float(b[i & 31]);
    }
  } else if (mode == 2) {
    for (int i = 0; i < iters; i += 2) {
      uint32_t r[64];
      ld_x64(tm + ((i * 32) & 255), r);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 63]);
    }
  } else if (mode == 4) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      ld_16x256b_x8(tm + ((i * 64) & 255), r);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 31]);
    }
  } else if (mode == 5) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      ld_16x128b_x16(tm + ((i * 64) & 255), r);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 31]);
    }
  } else {
    uint32_t a[32], b[32];
    tmem_ld_32x32(tm, a);
    for (int i = 0; i < iters; i += 2) {
      tmem_ld_wait();
      tmem_ld_32x32(tm + ((i * 32 + 32) & 255), b);
      acc += __uint_as_float(a[i & 31]);
      tmem_ld_wait();
      tmem_ld_32x32(tm + ((i * 32 + 64) & 255), a);
      acc += __uint_as_float(b[i & 31]);
    }
    tmem_ld_wait();
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 2048;
  const char* names[] = {"32x32b.x32+wait", "2 32x32b.x32 then wait", "32x32b.x64+wait", "32x32b.x32 software-pipelined",
                         "16x256b.x8+wait", "16x128b.x16+wait"};
  for (int mode = 0; mode < 6; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      tmem_read_kernel<<<1, warps * 32, 0>>>(mode, iters, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * 32 * 32 * 4;
      printf("%-30s warps=%2d : %8lld cycles, %7.1f B/clk/SM, %6.1f cycles per 4 KB load per warp  %s\n", names[mode],
             warps, h, bytes / h, (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
