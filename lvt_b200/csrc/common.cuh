// lvt_b200 :: common device/host helpers for the sm_100a kernels.
//
// Everything here is hand-written inline PTX for Blackwell (tcgen05 / TMEM / TMA / mbarrier).
// No CUTLASS/CuTe dependency: the bit layouts of the UMMA shared-memory descriptor and the
// instruction descriptor follow the PTX ISA tables (the same fields CuTe's
// cute/arch/mma_sm100_desc.hpp names) and are spelled out below.
#pragma once
#include <string.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef LVT_DEVICE_INLINE
#define LVT_DEVICE_INLINE __device__ __forceinline__
#endif

// ----------------------------------------------------------------------------------------
// Host-side error plumbing (C-ABI: every entry point returns 0 or a negative code and
// leaves a message retrievable with lvt_last_error()).
// ----------------------------------------------------------------------------------------
enum LvtStatus : int {
  LVT_OK = 0,
  LVT_ERR_INVALID = -1,   // bad argument / unsupported shape
  LVT_ERR_CUDA = -2,      // CUDA runtime / driver error
  LVT_ERR_NO_DEVICE = -3, // no sm_100 device
};

void lvt_set_error(const char* fmt, ...);

#define LVT_CHECK_ARG(cond, ...)                                                          \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      lvt_set_error(__VA_ARGS__);                                                         \
      return LVT_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)

#define LVT_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      lvt_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                    __LINE__);                                                            \
      return LVT_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define LVT_CHECK_LAUNCH()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      lvt_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                            \
      return LVT_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// ----------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): consecutive kernels of one stream (or of a captured CUDA graph)
// overlap the next grid's launch + prologue with the previous grid's tail.  Every kernel launched through
// lvt_launch() calls pdl_wait() before its first access to global memory; the persistent GEMM kernels call
// pdl_launch_dependents() when they start their LAST tile, so only the next grid's launch latency and prologue overlap the tail
// (griddepcontrol.wait returns when the prerequisite grid has completed and flushed), so data dependencies
// between neighbours are unchanged.  The launch attribute is only set when LVT_PDL=1 (see lvt_pdl_enabled()).
// ----------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  // elementwise kernels only wait: an early launch_dependents would let the next grid's CTAs sit on thread / register
  // slots of the waves of THIS grid that have not started yet (measured: the step got slower)
  pdl_wait();
}
bool lvt_pdl_enabled();
template <typename... P, typename... A>
static inline cudaError_t lvt_launch_cluster(int cluster, void (*kern)(P...), dim3 grid, dim3 block, size_t smem,
                                             cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (lvt_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {  // thread-block cluster along x (CTA pairs of the cta_group::2 kernels)
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}
template <typename... P, typename... A>
static inline cudaError_t lvt_launch(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     A&&... args) {
  return lvt_launch_cluster(1, kern, grid, block, smem, stream, static_cast<A&&>(args)...);
}
#endif

static inline int lvt_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------
// Device: shared-memory address helpers
// ----------------------------------------------------------------------------------------
LVT_DEVICE_INLINE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

LVT_DEVICE_INLINE bool elect_one() {
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

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
LVT_DEVICE_INLINE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
LVT_DEVICE_INLINE void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
LVT_DEVICE_INLINE void fence_proxy_async() {
  // make generic-proxy smem writes visible to the async proxy (TMA / tcgen05.mma reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
LVT_DEVICE_INLINE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
LVT_DEVICE_INLINE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LVT_DEVICE_INLINE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
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
// Bounded wait: a mis-programmed pipeline traps (→ a CUDA error the host reports) instead
// of hanging the GPU box. ~2^28 polls of a HW-sleeping try_wait is many seconds.
LVT_DEVICE_INLINE void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("lvt_b200: mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             threadIdx.x);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, global -> shared, completion on an mbarrier
// ----------------------------------------------------------------------------------------
LVT_DEVICE_INLINE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
LVT_DEVICE_INLINE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                   int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
LVT_DEVICE_INLINE void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                   int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// pulls one box into L2 ahead of the TMA load that will want it
LVT_DEVICE_INLINE void tma_prefetch_5d(const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// TMA store, shared -> global (bulk async group completion)
LVT_DEVICE_INLINE void tma_store_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                    int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}
LVT_DEVICE_INLINE void tma_reduce_add_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                         int c3, int c4) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}
LVT_DEVICE_INLINE void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING their smem source
template <int N>
LVT_DEVICE_INLINE void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
LVT_DEVICE_INLINE void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
LVT_DEVICE_INLINE void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
LVT_DEVICE_INLINE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
LVT_DEVICE_INLINE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
LVT_DEVICE_INLINE void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
LVT_DEVICE_INLINE void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
LVT_DEVICE_INLINE void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tf32 inputs (fp32 bits in smem), fp32 accumulate.
LVT_DEVICE_INLINE void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
LVT_DEVICE_INLINE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// TMEM -> registers: 32 lanes (this warp's quarter) x 32 consecutive fp32 columns.
LVT_DEVICE_INLINE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns.
LVT_DEVICE_INLINE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
LVT_DEVICE_INLINE void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2): one 256-row MMA tile spans the two SMs of a
// TPC; each CTA stages its own 128 rows of A and HALF of the B tile, so the shared-memory fill traffic per
// SM drops by a third against two independent 128 x 256 tiles.  The leader (cluster rank 0) owns the
// "operands landed" barriers and issues the MMAs; completion is multicast to both CTAs' barriers.
// (PTX forms as in the ISA's tcgen05 / cp.async.bulk.tensor .cta_group::2 sections.)
// ----------------------------------------------------------------------------------------
LVT_DEVICE_INLINE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
LVT_DEVICE_INLINE void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) inside CTA `rank` of the cluster
LVT_DEVICE_INLINE uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
LVT_DEVICE_INLINE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (same offset, cluster rank 0:
// bit 24 of a shared::cluster address selects the CTA of a pair); issued by both CTAs of the pair.
LVT_DEVICE_INLINE void tma_load_5d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                       int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
LVT_DEVICE_INLINE void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
LVT_DEVICE_INLINE void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
LVT_DEVICE_INLINE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows per CTA] * B[N x 16: N/2 rows per CTA]; issued by ONE thread of the leader
LVT_DEVICE_INLINE void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs of this thread have completed
LVT_DEVICE_INLINE void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B (PTX ISA "Shared memory descriptor" table):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset>>4 [46,48) version = 1 (sm_100)
//   [49,52) base offset = 0           [61,64) layout type: 2 = SWIZZLE_128B
// K-major tile  (rows x 64 bf16, one 128 B swizzle row per matrix row):
//   SBO = 1024 B (8 rows), LBO unused (set to 16 B).
// MN-major tile (64 bf16 of M/N contiguous per 128 B row, one row per k):
//   LBO = byte distance between consecutive 64-element M/N atoms, SBO = 1024 B (8 k rows).
LVT_DEVICE_INLINE uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                          uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 / kind::tf32 (PTX ISA "Instruction descriptor"):
//   [4,6) D fmt (1 = f32)  [7,10) A fmt  [10,13) B fmt (f16: 0 = f16, 1 = bf16; tf32: 2)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
LVT_DEVICE_INLINE constexpr uint32_t umma_idesc(int m, int n, int ab_fmt, bool a_mn, bool b_mn) {
  return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) |
         ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// small math helpers
// ----------------------------------------------------------------------------------------
LVT_DEVICE_INLINE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
LVT_DEVICE_INLINE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// one 16-byte vector reduction instead of four scalar atomics (the address must be 16-byte aligned)
LVT_DEVICE_INLINE void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
LVT_DEVICE_INLINE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
