// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// UMMA shared-memory / instruction descriptors, small math utilities.
//
// Everything here is written directly against the PTX ISA for sm_100a (tcgen05.*, cp.async.bulk.tensor.*);
// there is no dependency on CUTLASS/CuTe.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace b200 {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define B200_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      b200::set_last_error(__VA_ARGS__);     \
      return -1;                             \
    }                                        \
  } while (0)

// argument structs of the C-ABI start with their own size: a caller built against another layout is rejected
#define B200_CHECK_STRUCT(args, T, name)                                                                      \
  do {                                                                                                       \
    B200_CHECK_ARG((args) != nullptr, name ": null argument struct");                                        \
    B200_CHECK_ARG((args)->struct_size == sizeof(T), name ": struct_size %u != %zu (caller built against a "  \
                   "different include/pixparse_b200.h)", (args)->struct_size, sizeof(T));                   \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                   \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return b200::check_cuda(e__, name);   \
  } while (0)

int num_sms();
// Programmatic dependent launch (opt-in: PIXPARSE_B200_PDL=1): see launch_kernel() below
bool pdl_enabled();

// TMA descriptor encode (host). dims/strides innermost-first; strides in bytes for dims 1..rank-1.
enum TmaSwizzle { TMA_SWIZZLE_NONE = 0, TMA_SWIZZLE_64B = 2, TMA_SWIZZLE_128B = 3 };
enum TmaDtype { TMA_BF16 = 0, TMA_F32 = 1 };
int make_tmap(CUtensorMap* out, const void* base, TmaDtype dt, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swz);

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// ---- programmatic dependent launch --------------------------------------------------------------
// The step is ~480 dependent kernels; launched back to back each one pays grid-launch latency plus its prologue
// (barrier init, tensor-memory allocation, descriptor prefetch) AFTER its predecessor has drained. With the
// programmatic-stream-serialization launch attribute a kernel's CTAs may become resident as soon as every CTA of the
// predecessor has called pdl_launch_dependents() (first thing in every kernel here) and an SM has room, run their
// prologue, and block in pdl_wait() until the predecessor grid has completed and its writes are visible. Rule: no
// global-memory access of any kind before pdl_wait(). Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- proxies / fences -------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// orders accesses made through the async proxy (TMA, bulk reduce-adds) with generic-proxy accesses, all state spaces
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA --------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1-D bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a 2-cluster on one TPC execute one 256-row MMA -----------------
// Each CTA holds its own 128 rows of A and HALF of the B columns in shared memory and receives its 128 accumulator
// rows in its own TMEM; the even (leader) CTA issues the MMAs. Halving the B bytes each SM pulls from L2 is the point:
// at 128 x 256 tiles every K = 768 GEMM of the step ran into the L2 -> SM bandwidth (~14 TB/s measured).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  // default semantics (release at CTA scope): the cluster-scope form costs every arriving thread a MEMBAR.GPU + ERRBAR
  // (9 % of the epilogue warps' samples); the TMEM reads this arrive publishes are ordered by tcgen05.fence
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// TMA load into this CTA's shared memory whose transaction bytes are credited to a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {      // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base_lane + i), columns [c, c+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// same, into elements [OFF, OFF+32) of a larger register array (keeps the array in registers: no pointer casts)
template <int OFF, int N>
__device__ __forceinline__ void tmem_ld_32x32_at(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(OFF + 32 <= N, "out of range");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15]), "=r"(r[OFF + 16]),
        "=r"(r[OFF + 17]), "=r"(r[OFF + 18]), "=r"(r[OFF + 19]), "=r"(r[OFF + 20]), "=r"(r[OFF + 21]),
        "=r"(r[OFF + 22]), "=r"(r[OFF + 23]), "=r"(r[OFF + 24]), "=r"(r[OFF + 25]), "=r"(r[OFF + 26]),
        "=r"(r[OFF + 27]), "=r"(r[OFF + 28]), "=r"(r[OFF + 29]), "=r"(r[OFF + 30]), "=r"(r[OFF + 31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bits.
//   K-major operand  (rows = M/N index, 64 bf16 = 128 B of K per row, 8-row groups of 1024 B):
//       SBO = byte distance between 8-row groups, LBO unused (1).
//   MN-major operand (rows = K index, 64 bf16 = 128 B of M/N per row, 8-k groups of 1024 B):
//       SBO = byte distance between 8-k groups, LBO = byte distance between 64-element M/N chunks.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                              // D format: f32
         | (1u << 7)                            // A format: bf16
         | (1u << 10)                           // B format: bf16
         | ((a_mn_major ? 1u : 0u) << 15)       // A major
         | ((b_mn_major ? 1u : 0u) << 16)       // B major
         | ((uint32_t)(N >> 3) << 17)           // N / 8
         | ((uint32_t)(M >> 4) << 24);          // M / 16
}

// ---- math -------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf-GELU (timm nn.GELU / BART "gelu"): x * Phi(x). Phi is evaluated as 0.5 * (1 + tanh(u(x))) with an odd minimax
// polynomial u(x) = x (c0 + c1 x^2 + c2 x^4) fitted to atanh(erf(x / sqrt 2)) on [-6, 6]: max |gelu error| 3.4e-5 (plus
// MUFU.TANH's 2^-11 relative error), two orders of magnitude below the bf16 resolution of the stored activation, and
// a 6-deep dependency chain (1 MUFU) instead of erff()'s ~25 instructions. This matters: the GELU epilogue is
// latency-bound (ncu: issue slots 38 % busy, 'wait' + scoreboard stalls dominate).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_arg(float x) {
  x = fminf(fmaxf(x, -6.0f), 6.0f);      // the fit holds on [-6, 6]; beyond it Phi is 0 / 1 to 1e-9 and tanh(u(6)) = 1 - 4e-9
  const float x2 = x * x;
  float p = fmaf(x2, -3.47437544e-04f, 3.69593885e-02f);
  p = fmaf(x2, p, 7.97600733e-01f);
  return x * p;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(gelu_tanh_arg(x)), hx);
}
// d/dx [x Phi(x)] = Phi(x) + x phi(x), phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = fmaf(0.5f, tanh_approx(gelu_tanh_arg(x)), 0.5f);
  const float E = ex2_approx(x * x * -0.72134752044448170f);
  return fmaf(x * 0.3989422804014327f, E, cdf);
}

// ---- explicit shared-space 16-byte accesses -----------------------------------------------------------
// Pointers derived from the dynamic shared-memory base are generic to the compiler: it emits LD.E / ST.E, which hold
// their registers on the long scoreboard (ncu: the GELU epilogue spent 40 % of its stall samples there). These
// take the 32-bit shared address (smem_u32) instead.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// polled shared-memory words (the tile ring of the GEMM's dynamic scheduler): volatile so that a spin loop re-reads them;
// the second form stores into the same offset of another CTA of the cluster (address from mapa_u32)
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32_volatile(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts32_cluster(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.volatile.shared::cluster.b32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// fire-and-forget fp32 vector add into global memory (performed at L2), 16-byte aligned
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) ------------------------------------------------
// A 3-register FFMA issues every second cycle per SM sub-partition; the f32x2 forms do two lanes' worth per issue, so
// the fp32-heavy epilogues (bias, GELU, GELU', dropout scaling) cost half the fma-pipe slots.
// (the float2 intrinsics of sm_100_rt.h rather than inline PTX on .b64 registers: with the asm form ptxas spent a
//  quarter of the GELU epilogue's instructions on IMAD.MOV register-pair shuffles, with the intrinsics none)
typedef float2 f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { return make_float2(lo, hi); }
__device__ __forceinline__ f32x2 f2_splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  lo = v.x;
  hi = v.y;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { return __fadd2_rn(a, b); }
// pair forms of gelu_erf / gelu_erf_grad (same polynomial; the clamp moves to x^2 <= 36, one FMNMX per lane:
// beyond |x| = 6 the tanh argument keeps growing linearly with the positive slope p(36), so tanh -> +-1 as it must)
__device__ __forceinline__ f32x2 gelu_tanh2(f32x2 x, f32x2& x2c) {
  float a, b;
  f2_unpack(f2_mul(x, x), a, b);
  x2c = f2_pack(fminf(a, 36.0f), fminf(b, 36.0f));
  f32x2 p = f2_fma(x2c, f2_splat(-3.47437544e-04f), f2_splat(3.69593885e-02f));
  p = f2_fma(x2c, p, f2_splat(7.97600733e-01f));
  f2_unpack(f2_mul(x, p), a, b);
  return f2_pack(tanh_approx(a), tanh_approx(b));
}
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  f32x2 x2c;
  const f32x2 t = gelu_tanh2(x, x2c);
  const f32x2 hx = f2_mul(x, f2_splat(0.5f));
  return f2_fma(hx, t, hx);
}
__device__ __forceinline__ f32x2 gelu_erf_grad2(f32x2 x) {
  f32x2 x2c;
  const f32x2 t = gelu_tanh2(x, x2c);
  const f32x2 cdf = f2_fma(t, f2_splat(0.5f), f2_splat(0.5f));
  float a, b;
  f2_unpack(f2_mul(x2c, f2_splat(-0.72134752044448170f)), a, b);      // exp(-x^2 / 2); 1.5e-8 at the clamp
  const f32x2 E = f2_pack(ex2_approx(a), ex2_approx(b));
  return f2_fma(f2_mul(x, f2_splat(0.3989422804014327f)), E, cdf);
}

// 2^x for x <= 0 on the FMA pipe, two values at a time (softmax probabilities; FA4-style MUFU offload). The attention
// kernels are bound by the 16 ex2 / clk / SM of the XU pipe while their FMA pipe idles, so a share of the exponentials is
// evaluated here instead: Cody-Waite reduction x = n + f with n = round(x) taken from the low mantissa bits of
// x + 1.5 * 2^23, a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, fifty times below
// bf16 resolution -- every consumer rounds the result to bf16), and n added straight into the exponent field.
// Arguments below -126 are clamped (2^-126 ~ 1e-38 stands in for 0; -inf marks padded rows).
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& y0, float& y1) {
  const f32x2 x = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const f32x2 t = f2_add(x, f2_splat(12582912.0f));                 // 1.5 * 2^23: ulp(t) = 1
  const f32x2 n = f2_add(t, f2_splat(-12582912.0f));
  const f32x2 f = f2_fma(n, f2_splat(-1.0f), x);
  f32x2 p = f2_fma(f, f2_splat(5.517166853e-02f), f2_splat(2.426111251e-01f));
  p = f2_fma(p, f, f2_splat(6.932609677e-01f));
  p = f2_fma(p, f, f2_splat(9.999280572e-01f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---- dropout: stateless counter-based masks ---------------------------------------------------------
// One 32-bit hash of (seed, pair index) yields two 16-bit uniforms -> keep decisions for two adjacent elements.
// keep <=> u16 >= threshold16, threshold16 = round(p * 65536); kept values are scaled by 65536 / (65536 - threshold16).
// Forward and backward regenerate identical masks from (seed, index); nothing is stored.
__device__ __forceinline__ uint32_t dropout_hash(uint32_t seed, uint32_t pair_idx) {
  uint32_t x = pair_idx * 0x9E3779B1u + seed;
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float dropout_scale(uint32_t threshold16) {
  return 65536.0f / (65536.0f - (float)threshold16);
}
// multiplies the element pair (a, b) whose first element has linear index 2 * pair_idx
__device__ __forceinline__ void dropout_pair(uint32_t seed, uint32_t pair_idx, uint32_t threshold16, float scale,
                                             float& a, float& b) {
  const uint32_t h = dropout_hash(seed, pair_idx);
  a = ((h & 0xFFFFu) >= threshold16) ? a * scale : 0.f;
  b = ((h >> 16) >= threshold16) ? b * scale : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Launch `kern` on `s` (optionally as clusters of cluster_x CTAs) with programmatic dependent launch when enabled.
// Only kernels that follow the pdl_wait() rule above may go through here.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#endif  // __CUDACC__

}  // namespace b200
