// countr_b200 — sm_100a device primitives (mbarrier, TMA, tcgen05/TMEM) as inline PTX.
//
// Everything in here is hand-written for Blackwell (compute_100a). There is no
// fallback path: the translation units that include this header only build for
// sm_100a and the host side refuses to run on anything else.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace countr {

// ---------------------------------------------------------------------------
// error plumbing shared by all translation units (defined in api.cu)
// ---------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);

enum : int {
  COUNTR_OK = 0,
  COUNTR_ERR_ARG = -1,      // bad shape / alignment / null pointer
  COUNTR_ERR_CUDA = -2,     // a CUDA runtime/driver call failed
  COUNTR_ERR_WORKSPACE = -3,// workspace too small
  COUNTR_ERR_UNSUPPORTED = -4,
};

#define COUNTR_CHECK_CUDA(expr)                                                        \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return ::countr::set_error(::countr::COUNTR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);          \
  } while (0)

#define COUNTR_REQUIRE(cond, ...)                                                      \
  do {                                                                                 \
    if (!(cond)) return ::countr::set_error(::countr::COUNTR_ERR_ARG, __VA_ARGS__);    \
  } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor drains; pdl_wait() blocks until the predecessor grid has
// completed and its memory is visible, pdl_trigger() lets the successor begin its own prologue early.
// Both are no-ops for kernels launched without the attribute.  This hides ~1-2 us of launch latency and
// prologue (barrier init, TMEM alloc, descriptor prefetch) per kernel; the step has ~260 kernels.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // api.cu: COUNTR_PDL environment switch (default on)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: a call site remembers which devices it has
// configured, so a second GPU driven from the same process gets its own call.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
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

// ---------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug becomes a trap (reported as a launch failure) instead of a
// hung GPU box.  2^24 polls with the HW-suspending try_wait is seconds; no legitimate wait here exceeds ~1 ms.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1   // left alone, nvcc unrolls the poll 64x at every wait site (hundreds of KB of dead code per kernel)
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// ---------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — tile mode, mbarrier completion
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA stores (shared::cta -> global, bulk async-group completion).  The issuing thread commits the group and later waits
// for it: wait_read<N> = at most N groups still READING their shared-memory source (the buffer may be overwritten),
// wait_all<N> = at most N groups not yet complete (global writes performed).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// element-wise fp32 add into global memory (the tensor map's data type selects the arithmetic): split-K accumulation
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------
// thread-block clusters: rank, cluster-wide barrier, TMA multicast, multicast commit
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One L2 read, delivered to the same CTA-relative smem offset (and mbarrier) of every CTA in cta_mask.
__device__ __forceinline__ void tma_load_4d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this CTA-relative offset in every CTA of cta_mask.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM cols: pow2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]  (A operand read from tensor memory)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make an mbarrier track completion of all prior tcgen05 async ops of this thread.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (base_lane+t),
// columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---------------------------------------------------------------------------
// CTA pairs (tcgen05 cta_group::2): one MMA spans two SMs — each CTA supplies its 128 rows of A and half of
// the B tile from its own shared memory and receives its 128 accumulator rows in its own TMEM.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
// TMA load into THIS CTA's smem whose completion bytes are credited to an mbarrier of either CTA of the pair
// (bar_cluster_addr is a shared::cluster address, e.g. mapa(local, 0) for the leader's barrier).
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// One 64-wide k-block of the GEMM main loop as ONE instruction sequence: a non-blocking probe of the NEXT stage's `full`
// barrier, four tcgen05.mma (k-steps of 16) issued by the elected lane, the commit that frees the stage and (last k-block)
// the commit that publishes the accumulator.  The probe's result is only read after the MMAs have been issued, so the
// barrier round trip (~150 clk for test_wait) hides behind the ~60 clk each UTCHMMA takes to issue instead of preceding it
// (measured with scripts/trace_gemm.py: the issue loop, not the tensor pipe or TMA, paced the 128 x 64..192 tiles at
// 410 + bn/2 clk per k-block against a pipe time of 2*bn).  Executed by the WHOLE warp; returns 1 when the probed barrier
// phase has completed.  kGroup = 1: single CTA; 2: CTA pair (commits are multicast to both CTAs, mask 3).
template <int kGroup>
__device__ __forceinline__ uint32_t umma_kblock(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t a_step, uint32_t b_step,
                                                uint32_t idesc, uint32_t accumulate, uint32_t empty_bar, uint32_t acc_bar,
                                                uint32_t is_last, uint32_t probe_bar, uint32_t probe_parity) {
  uint32_t ready;
  const uint64_t as = a_step, bs = b_step;
  if (kGroup == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, e, l;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q, [%11], %12;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.ne.b32 l, %10, 0;\n\t"
        "and.pred l, l, e;\n\t"
        "add.u64 a1, %2, %4;\n\t"
        "add.u64 b1, %3, %5;\n\t"
        "add.u64 a2, a1, %4;\n\t"
        "add.u64 b2, b1, %5;\n\t"
        "add.u64 a3, a2, %4;\n\t"
        "add.u64 b3, b2, %5;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a1, b1, %6, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a2, b2, %6, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a3, b3, %6, 1;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "@l tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t"
        "}\n"
        : "=r"(ready)
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "l"(as), "l"(bs), "r"(idesc), "r"(accumulate), "r"(empty_bar), "r"(acc_bar),
          "r"(is_last), "r"(probe_bar), "r"(probe_parity)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, e, l;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b16 m;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q, [%11], %12;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.ne.b32 l, %10, 0;\n\t"
        "and.pred l, l, e;\n\t"
        "mov.b16 m, 3;\n\t"
        "add.u64 a1, %2, %4;\n\t"
        "add.u64 b1, %3, %5;\n\t"
        "add.u64 a2, a1, %4;\n\t"
        "add.u64 b2, b1, %5;\n\t"
        "add.u64 a3, a2, %4;\n\t"
        "add.u64 b3, b2, %5;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], a1, b1, %6, 1;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], a2, b2, %6, 1;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], a3, b3, %6, 1;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], m;\n\t"
        "@l tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], m;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t"
        "}\n"
        : "=r"(ready)
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "l"(as), "l"(bs), "r"(idesc), "r"(accumulate), "r"(empty_bar), "r"(acc_bar),
          "r"(is_last), "r"(probe_bar), "r"(probe_parity)
        : "memory");
  }
  return ready;
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---------------------------------------------------------------------------
// UMMA descriptors (bit layout as published in CUTLASS cute/arch/mma_sm100_desc.hpp)
// ---------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, 16-bit elements.
//   K-major  operand: rows of 64 elements (128 B); 8-row core groups 1024 B apart (SBO).
//                     LBO is ignored for swizzled K-major layouts.
//   MN-major operand: 64 MN-elements (128 B) x 8 K-rows = 1024 B atoms; LBO = byte distance
//                     between 64-element MN blocks, SBO = byte distance between 8-row K groups.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // [0,14)  start address
  d |= static_cast<uint64_t>((lbo_bytes & 0x3FFFFu) >> 4) << 16;  // [16,30) leading byte offset
  d |= static_cast<uint64_t>((sbo_bytes & 0x3FFFFu) >> 4) << 32;  // [32,46) stride byte offset
  d |= 1ull << 46;                                                // [46,48) descriptor version (sm_100)
  d |= 2ull << 61;                                                // [61,64) SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16, fp16 (or bf16) operands, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, bool a_mn_major, bool b_mn_major,
                                                      bool bf16 = false) {
  return (1u << 4)                          // c_format = F32
         | ((bf16 ? 1u : 0u) << 7)          // a_format
         | ((bf16 ? 1u : 0u) << 10)         // b_format
         | ((a_mn_major ? 1u : 0u) << 15)   // a_major
         | ((b_mn_major ? 1u : 0u) << 16)   // b_major
         | (static_cast<uint32_t>(n >> 3) << 17)   // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);  // m_dim
}

// ---------------------------------------------------------------------------
// numerics shared by epilogues
// ---------------------------------------------------------------------------
// erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. at fp32 rounding level):
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),  t = 1 / (1 + p z),  z >= 0.
// ~12 instructions (1 MUFU.RCP + 1 MUFU.EX2) instead of erff's ~40 with branches: the GELU
// epilogues of fc1 (M x 4D elements per layer) were issue-bound on erff.
__device__ __forceinline__ void erf_parts(float x, float& erf_abs, float& gauss) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  gauss = __expf(-z * z);  // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  erf_abs = fmaf(-poly * t, gauss, 1.0f);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e, g;
  erf_parts(x, e, g);
  return 0.5f * x * (1.0f + copysignf(e, x));
}
// d/dx [ 0.5 x (1 + erf(x/sqrt2)) ] = 0.5 (1 + erf(x/sqrt2)) + x * exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e, g;
  erf_parts(x, e, g);
  return fmaf(x * 0.39894228040143267794f, g, 0.5f * (1.0f + copysignf(e, x)));
}


// ---------------------------------------------------------------------------
// packed fp32x2 math (sm_100 FFMA2 / FMUL2 / FADD2): two lanes per instruction.  The GELU epilogues are
// issue-bound (M x 4D erf evaluations per MLP), so halving the FMA instruction count matters.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f2_pack(float2 v) {
  return (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32) | __float_as_uint(v.x);
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long u) {
  return make_float2(__uint_as_float(static_cast<uint32_t>(u)), __uint_as_float(static_cast<uint32_t>(u >> 32)));
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two-lane version of erf_parts (same A&S 7.1.26 polynomial): e = erf(x/sqrt2) with the sign of x, g = exp(-x^2/2)
__device__ __forceinline__ void erf_parts2(float2 x, float2& e, float2& g) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = mul2(ax, splat2(0.70710678118654752440f));
  const float2 den = fma2(z, splat2(0.3275911f), splat2(1.0f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  const float2 e2 = mul2(mul2(z, z), splat2(-1.44269504088896340736f));
  g = make_float2(ex2_approx_f(e2.x), ex2_approx_f(e2.y));
  float2 np = fma2(splat2(-1.061405429f), t, splat2(1.453152027f));   // negated polynomial
  np = fma2(np, t, splat2(-1.421413741f));
  np = fma2(np, t, splat2(0.284496736f));
  np = fma2(np, t, splat2(-0.254829592f));
  const float2 ea = fma2(mul2(np, t), g, splat2(1.0f));                // erf(|x|/sqrt2) >= 0
  e = make_float2(__uint_as_float(__float_as_uint(ea.x) | (__float_as_uint(x.x) & 0x80000000u)),
                  __uint_as_float(__float_as_uint(ea.y) | (__float_as_uint(x.y) & 0x80000000u)));
}
// Forward GELU only needs erf, not the Gaussian: Abramowitz & Stegun 7.1.28,
//   erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16,  z >= 0   (|abs err| <= 3e-7; ~1.8e-6 in fp32 after the four squarings),
// costs ONE MUFU (the reciprocal) per element instead of the two (rcp + ex2) of 7.1.26 — the fc1 epilogue evaluates
// M x 4D of these per MLP and was MUFU-bound (2 x 128 x 256 per tile at 16/clk/SM = 4.1 k clk against a 6 k clk main loop).
// A huge |x| overflows the 16th power to +inf, whose reciprocal is +0: erf -> 1, no NaN.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = mul2(ax, splat2(0.70710678118654752440f));
  float2 pl = fma2(splat2(0.0000430638f), z, splat2(0.0002765672f));
  pl = fma2(pl, z, splat2(0.0001520143f));
  pl = fma2(pl, z, splat2(0.0092705272f));
  pl = fma2(pl, z, splat2(0.0422820123f));
  pl = fma2(pl, z, splat2(0.0705230784f));
  pl = fma2(pl, z, splat2(1.0f));
  pl = mul2(pl, pl);
  pl = mul2(pl, pl);
  pl = mul2(pl, pl);
  pl = mul2(pl, pl);
  const float2 r = make_float2(rcp_approx(pl.x), rcp_approx(pl.y));
  const float2 ea = fma2(r, splat2(-1.0f), splat2(1.0f));               // erf(|x|/sqrt2) >= 0
  const float2 e = make_float2(__uint_as_float(__float_as_uint(ea.x) | (__float_as_uint(x.x) & 0x80000000u)),
                               __uint_as_float(__float_as_uint(ea.y) | (__float_as_uint(x.y) & 0x80000000u)));
  const float2 hx = mul2(x, splat2(0.5f));
  return fma2(hx, e, hx);
}
__device__ __forceinline__ float2 gelu_erf_grad2(float2 x) {
  float2 e, g;
  erf_parts2(x, e, g);
  return fma2(mul2(x, splat2(0.39894228040143267794f)), g, fma2(e, splat2(0.5f), splat2(0.5f)));
}

}  // namespace countr
