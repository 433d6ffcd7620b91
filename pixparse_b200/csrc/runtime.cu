// Host-side runtime glue shared by every entry point of the C-ABI library: last-error string,
// device properties, TMA descriptor encoding through the driver entry point (no link-time libcuda).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <cstdlib>
#include "common.cuh"
#include "../../include/pixparse_b200.h"

namespace b200 {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_last_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    // PIXPARSE_B200_RESERVE_SMS=<k>: size the persistent grids for k fewer SMs (even, < half the part), e.g. to leave
    // room for NCCL kernels that otherwise only run in the gaps between kernels owning every SM
    const char* env = getenv("PIXPARSE_B200_RESERVE_SMS");
    if (env != nullptr) {
      int k = atoi(env) & ~1;
      if (k > 0 && k < n / 2) n -= k;
    }
    cached = n;
  }
  return cached;
}

static int g_pdl_override = -1;      // -1: follow PIXPARSE_B200_PDL; 0 / 1: forced by b200_set_pdl()

bool pdl_enabled() {
  if (g_pdl_override >= 0) return g_pdl_override != 0;
  static int cached = -1;
  if (cached < 0) {
    const char* env = getenv("PIXPARSE_B200_PDL");
    // opt-in: measured on B200 (profiles/r02_experiments_no_gain.txt) the step is 0.5-1 % SLOWER with it -- the kernels here own whole
    // SMs, so a dependent CTA cannot become resident before its predecessor's CTA has left, and the front end already
    // hides the launch latency of a queued kernel
    cached = (env != nullptr && env[0] == '1') ? 1 : 0;
  }
  return cached != 0;
}

}  // namespace b200

// The greedy-decode step is the opposite regime (pixparse_b200/decode.py): ~70 dependent kernels of a few microseconds each,
// most of them streaming weights that do NOT depend on the predecessor -- there a dependent grid that is resident early
// has its weight loads in flight while the predecessor drains. The host switches PDL on around the capture of that step.
extern "C" int b200_set_pdl(int mode) {
  const int prev = b200::g_pdl_override;
  b200::g_pdl_override = mode < 0 ? -1 : (mode != 0);
  return prev;
}

namespace b200 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_last_error("cuTensorMapEncodeTiled entry point unavailable (err %d)", (int)e);
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, TmaDtype dt, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -2;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("TMA base pointer %p not 16-byte aligned", base);
    return -1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (gstr[i] % 16 != 0) {
      set_last_error("TMA global stride %llu (dim %d) not a multiple of 16 bytes", (unsigned long long)gstr[i], i + 1);
      return -1;
    }
  }
  CUresult r = fn(out, dt == TMA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz == TMA_SWIZZLE_128B ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (swz == TMA_SWIZZLE_64B ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                   rank > 1 ? box[1] : 0);
    return -3;
  }
  return 0;
}

}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::g_last_error; }

extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }

extern "C" int b200_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b200::check_cuda(e, "cudaGetDevice");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    b200::set_last_error("device compute capability %d.%d is not sm_100 (B200); no fallback path exists", major, minor);
    return -4;
  }
  return 0;
}
