// countr_b200 — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution.
//
//   warp 0      : TMA producer (one elected lane) — fills a 4-deep ring of {A 128x64, B bn x64}
//                 fp16 tiles in SWIZZLE_128B shared memory
//   warp 1      : TMEM allocator + MMA issuer (one lane) — tcgen05.mma kind::f16, M=128, N=bn,
//                 fp32 accumulators double-buffered in tensor memory (2 x 256 columns)
//   warps 2..9  : epilogue — tcgen05.ld the accumulator (thread == row), transpose each 32x32 chunk through
//                 swizzled shared memory so that 8 lanes cover 128 contiguous bytes of ONE output row, then
//                 alpha/bias/GELU/GELU'/residual/GroupNorm statistics and coalesced fp16 / fp32 stores
//
// Both single-thread roles run their loops WARP-UNIFORMLY and only predicate the issue itself with elect.sync:
// inside an `if (lane == 0)` region ptxas cannot keep descriptors / coordinates in uniform registers and wraps
// every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST loop (~130-190 clk per instruction, measured with
// scripts/trace_gemm.py: the MMA thread then needs 730 clk per 64-wide k-block, twice the tensor-pipe time).
//
// Conv mode re-uses the whole pipeline: the A tile of k-block (tap, cin-block) is the TMA box of
// the NHWC activation shifted by the tap offset; the TMA unit zero-fills the halo, so there is no
// im2col buffer and no padding copy.
//
// Reference arithmetic replaced: nn.Linear / Conv2d(3x3, s1, p1) call sites listed in
// include/countr_b200.h (models_crossvit.py:55-92,104-127; models_mae_cross.py:39,53-100,152).
#include "../../include/countr_b200.h"
#include "common.cuh"
#include "tma.h"

#include <algorithm>

namespace countr {

// Optional in-kernel timeline (clock64 stamps of one CTA's producer / MMA / epilogue roles) for scripts/trace_gemm.py.
// Compiled only into the -DCOUNTR_TRACE build (make trace); the product library carries none of it.
#ifdef COUNTR_TRACE
__device__ long long* g_trace = nullptr;
__device__ int g_trace_cta = 0;
__device__ int g_dbg = 0;   // experiment knobs: 1 = skip the C stores, 2 = skip the residual loads, 4 = no TMA operand loads
                            // (fast-path producer), 8 = no MMAs (single-CTA path)
#define DBG_INIT const int dbg_ = g_dbg
#define DBG(bit) (dbg_ & (bit))
#define TR_INIT long long* const tr_ = (g_trace != nullptr && static_cast<int>(blockIdx.x) == g_trace_cta) ? g_trace : nullptr
#define TR(slot)                                                     \
  do {                                                               \
    if (tr_ != nullptr && (threadIdx.x & 31) == 0 && (slot) < 2048) tr_[(slot)] = clock64();    \
  } while (0)
#else
#define DBG_INIT do {} while (0)
#define DBG(bit) 0
#define TR_INIT do {} while (0)
#define TR(slot) do {} while (0)
#endif

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kMaxStages = 8;
constexpr int kMaxBN = 256;
constexpr uint32_t kABytes = BM * BK * 2;       // 16 KB
constexpr int kEpiWarps = 8;                     // 2 per TMEM lane quarter: they split the column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kEpiBufBytes = 32 * 128;      // one epilogue staging buffer: 32 rows x 128 bytes (XOR / SWIZZLE_128B layout)
// Dynamic shared memory (the 227 KB opt-in maximum), laid out by the host per launch (GemmArgs):
//   [0, nstages * stage_bytes)                    operand ring: A 16 KB + B bn x 128 B per stage (pair mode: half of B)
//   [epi_off, + 8 warps * nbuf * 4 KB)            epilogue staging, 1024-byte aligned (TMA SWIZZLE_128B source / destination)
//   [kBarOff, + 512)                              mbarriers + TMEM base address
constexpr uint32_t kSmemBytes = 232448;
constexpr uint32_t kBarOff = kSmemBytes - 1024 /*alignment slack*/ - 512;
constexpr uint32_t kPairStageBytes = 32768;

struct GemmArgs {
  int M, N, K;
  int nb2;            // inner batch count (batch = b1 * nb2 + b2)
  int bn;             // N tile
  int a_mn, b_mn;
  int bf16;
  int split_k, k_per_split;
  int m_tiles, n_tiles, total_tiles;   // total_tiles counts CLUSTER tiles (cs consecutive m tiles x one n tile)
  uint32_t mg_magic, nt_magic, sk_magic, nb2_magic;   // fast_div constants of m_groups, n_tiles, split_k, nb2
  int epi64;                           // 1: 16-bit outputs drain 64 columns per round (COUNTR_EPI64=0 turns it off for A/B runs)
  int nstages, stage_bytes;            // operand ring
  int epi_off, nbuf;                   // epilogue staging: offset, buffers per epilogue warp (1 or 2)
  int epi_mode;                        // 0: LSU stores (all modes); 1: 16-bit tile through TMA stores; 2: fp32 tile through TMA
                                       // stores / reduce-adds, residual through TMA loads
  int box_x;                           // conv: x extent of one warp's 32-pixel store box
  int pair;                            // 1: CTA pair (cta_group::2): 256 x bn tile over two SMs, 6 stages of 32 KB
  int cs, m_groups;                    // cluster size (CTAs sharing the B tile via TMA multicast), ceil(m_tiles / cs)
  // conv mode
  int conv, H, W, cin_blocks, bx, by, tiles_x;   // conv: 0 none, 1 forward/dX implicit GEMM, 2 dW (pixels are K)
  int tiles_per_img;
  // epilogue
  void* C;
  long long ldc, sc1, sc2;
  int out_f32, atomic;
  float alpha;
  const float* bias;
  int act;
  void* aux;
  long long ldaux;
  const float* residual;
  long long ldr;
  int res_mod;
  double* gn_stats;
  // LayerNorm folded into the neighbouring GEMMs (see countr_gemm_desc)
  uint16_t* ln_x16;
  long long ld_x16;
  float* ln_stats;
  const float* ln_colsum;
  float ln_inv_dim, ln_eps;
};

struct TileCoord {
  int m, n, s, b1, b2;
};

// idx / d for idx * d < 2^32 with the host-computed magic = floor(2^32 / d) + 1 (d > 1): a multiply-high instead of the
// ~100-clk emulated integer division — five of them per tile decode were ~0.5 us of every kernel's prologue.
__device__ __forceinline__ int fast_div(int idx, int d, uint32_t magic) {
  return d == 1 ? idx : static_cast<int>(__umulhi(static_cast<uint32_t>(idx), magic));
}

__device__ __forceinline__ TileCoord decode_tile(const GemmArgs& p, int idx, int rank) {
  TileCoord t;
  int q = fast_div(idx, p.m_groups, p.mg_magic);
  t.m = (idx - q * p.m_groups) * p.cs + rank;
  idx = q;
  q = fast_div(idx, p.n_tiles, p.nt_magic);
  t.n = idx - q * p.n_tiles;
  idx = q;
  q = fast_div(idx, p.split_k, p.sk_magic);
  t.s = idx - q * p.split_k;
  idx = q;
  q = fast_div(idx, p.nb2, p.nb2_magic);
  t.b2 = idx - q * p.nb2;
  t.b1 = q;
  return t;
}

__device__ __forceinline__ uint32_t pack_16b(float a, float b, int bf16) {
  uint32_t r;
  if (bf16)
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_16b(uint32_t v, int bf16) {
  if (bf16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}

// kPair instantiates the cta_group::2 code path; kernels that contain cta_group::2 instructions must be launched
// as clusters of two, so the single-CTA variant is a separate instantiation.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nstages = p.nstages;
  const uint32_t stage_bytes = p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* full = bars;                 // [8]
  uint64_t* empty = bars + 8;            // [8]
  uint64_t* tmem_full = bars + 16;       // [2]
  uint64_t* tmem_empty = bars + 18;      // [2]
  uint64_t* res_full = bars + 20;        // [8 warps][3 buffers]: residual chunk landed (epi_mode 2)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 44);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  pdl_trigger();   // the next kernel may start its prologue as soon as SMs drain
  TR_INIT;
  DBG_INIT;
  if (threadIdx.x == 0) TR(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < nstages; ++i) {
      mbar_init(&full[i], 1);
      // multicast clusters: every CTA must have consumed the stage (one commit each); pair: ONE multicast commit
      mbar_init(&empty[i], kPair ? 1 : p.cs);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kPair ? 2 * kEpiWarps : kEpiWarps);   // pair: the leader waits for both CTAs' epilogues
    }
    for (int i = 0; i < 3 * kEpiWarps; ++i) mbar_init(&res_full[i], 1);
    if (p.epi_mode != 0) tma_prefetch_desc(&tma_c);
    if ((p.epi_mode == 2 && p.residual != nullptr) || (p.epi_mode == 1 && p.aux != nullptr)) tma_prefetch_desc(&tma_r);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_2cta<512>(tmem_ptr); else tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  if (p.cs > 1) cluster_sync_all(); else __syncthreads();   // peers' barriers must be initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) TR(1);
  const int rank = p.cs > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / p.cs, num_clusters = gridDim.x / p.cs;
  const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cs) - 1u);
  // pdl_wait() (griddepcontrol.wait: the predecessor grid has completed, its writes are visible) is issued per role,
  // right before the role's first access to global memory, so tile decoding / row bookkeeping (a handful of integer
  // divisions, ~0.5 us) also overlaps the predecessor's tail.  The MMA warp never touches global memory.
  if (threadIdx.x == 0) TR(2);

  const uint32_t b_bytes = static_cast<uint32_t>(kPair ? p.bn / 2 : p.bn) * BK * 2;   // B bytes landing in THIS CTA per stage

  if (warp == 0) {
    // ------------------------------- TMA producer (whole warp in the loop, one elected lane issues) -------------------------------
    int stage = 0;
    uint32_t phase = 0;
    int trk = 0;
    if (!kPair && p.conv == 0 && p.cs == 1 && p.a_mn == p.b_mn) {
      // ---- fast paths: plain Linear (both operands K-major) and dW = dY^T X (both MN-major), single CTA.  The general
      // loop below evaluates every mode per k-block (~100 instructions, ~390 clk per k-block measured): fine under a
      // 128 x 256 tile (512 clk of MMA per k-block) but it starves the tensor pipe at N tiles of 128 / 64 (256 / 128 clk).
      const bool mn = p.a_mn != 0;
      const int nboxb = p.bn >> 6;
      const uint32_t tx = kABytes + b_bytes;
      for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
        const TileCoord t = decode_tile(p, tile, rank);
        const int k_begin = t.s * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + BK - 1) / BK;
        const int m0 = t.m * BM, n0 = t.n * p.bn;
        if (tile == cluster_id) pdl_wait();   // first tile: the operands come from earlier kernels
        int k = k_begin;
        for (int kb = 0; kb < nkb; ++kb, k += BK) {
          mbar_wait(&empty[stage], phase ^ 1);
          TR(16 + trk); ++trk;
          uint8_t* sa = smem + stage * stage_bytes;
          if (DBG(4)) {
            if (elect_one()) mbar_arrive(&full[stage]);
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], tx);
            if (!mn) {
              tma_load_4d(sa, &tma_a, &full[stage], k, m0, t.b2, t.b1);
              tma_load_4d(sa + kABytes, &tma_b, &full[stage], k, n0, t.b2, t.b1);
            } else {
              tma_load_4d(sa, &tma_a, &full[stage], m0, k, t.b2, t.b1);
              tma_load_4d(sa + 8192, &tma_a, &full[stage], m0 + 64, k, t.b2, t.b1);
              for (int i = 0; i < nboxb; ++i)
                tma_load_4d(sa + kABytes + i * 8192, &tma_b, &full[stage], n0 + i * 64, k, t.b2, t.b1);
            }
          }
          __syncwarp();
          TR(256 + trk - 1);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (kPair && p.conv == 0 && !p.a_mn && !p.b_mn) {
      // ---- fast path: plain Linear on a CTA pair (both operands K-major): this CTA's 128 rows of A and its half of the B tile;
      // every byte is credited to the LEADER's full barrier
      const uint32_t tx2 = 2 * (kABytes + b_bytes);
      const int nhalf = p.bn / 2;
      for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
        const TileCoord t = decode_tile(p, tile, rank);
        const int k_begin = t.s * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + BK - 1) / BK;
        const int m0 = t.m * BM, n0 = t.n * p.bn + rank * nhalf;
        if (tile == cluster_id) pdl_wait();
        int k = k_begin;
        for (int kb = 0; kb < nkb; ++kb, k += BK) {
          mbar_wait(&empty[stage], phase ^ 1);
          TR(16 + trk); ++trk;
          uint8_t* sa = smem + stage * stage_bytes;
          const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], tx2);
            tma_load_4d_2sm(sa, &tma_a, full_leader, k, m0, t.b2, t.b1);
            tma_load_4d_2sm(sa + kABytes, &tma_b, full_leader, k, n0, t.b2, t.b1);
          }
          __syncwarp();
          TR(256 + trk - 1);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else
    for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile, rank);
      const int k_begin = t.s * p.k_per_split;
      const int k_end = min(p.K, k_begin + p.k_per_split);
      const int nkb = (k_end - k_begin + BK - 1) / BK;
      int x0 = 0, y0 = 0;
      if (p.conv == 1) {
        x0 = (t.m % p.tiles_x) * p.bx;
        y0 = (t.m / p.tiles_x) * p.by;
      }
      if (tile == cluster_id) pdl_wait();   // first tile: the operands come from earlier kernels
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        TR(16 + trk); ++trk;
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + kABytes;
        const int k = k_begin + kb * BK;
        if (elect_one()) {
          if (kPair) {
            // both CTAs fill their own stage; all bytes are credited to the LEADER's full barrier, which the
            // leader's MMA thread waits on before issuing the cta_group::2 MMAs over both shared memories
            const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (kABytes + b_bytes));
            const int nbox = p.bn / 128;   // 64-wide MN boxes of B held by EACH CTA (MN-major B only)
            if (p.conv == 2) {
              // weight gradient: A = this CTA's 128 output channels of dY, B = its half of the input channels
              const int kbg = k / BK;
              const int bimg = kbg / p.tiles_per_img;
              const int rr = kbg - bimg * p.tiles_per_img;
              const int px0 = (rr % p.tiles_x) * p.bx, py0 = (rr / p.tiles_x) * p.by;
              const int ky = t.b2 / 3, kx = t.b2 - ky * 3;
              tma_load_4d_2sm(sa, &tma_a, full_leader, t.m * BM, px0, py0, bimg);
              tma_load_4d_2sm(sa + 8192, &tma_a, full_leader, t.m * BM + 64, px0, py0, bimg);
              for (int i = 0; i < nbox; ++i)
                tma_load_4d_2sm(sb + i * 8192, &tma_b, full_leader, t.n * p.bn + (rank * nbox + i) * 64, px0 + kx - 1, py0 + ky - 1, bimg);
            } else {
              if (p.conv) {
                const int tap = kb / p.cin_blocks;
                const int cb = kb - tap * p.cin_blocks;
                const int ky = tap / 3, kx = tap - ky * 3;
                tma_load_4d_2sm(sa, &tma_a, full_leader, cb * BK, x0 + kx - 1, y0 + ky - 1, t.b1);
              } else if (!p.a_mn) {
                tma_load_4d_2sm(sa, &tma_a, full_leader, k, t.m * BM, t.b2, t.b1);
              } else {
                tma_load_4d_2sm(sa, &tma_a, full_leader, t.m * BM, k, t.b2, t.b1);
                tma_load_4d_2sm(sa + 8192, &tma_a, full_leader, t.m * BM + 64, k, t.b2, t.b1);
              }
              if (!p.b_mn) {
                tma_load_4d_2sm(sb, &tma_b, full_leader, k, t.n * p.bn + rank * (p.bn / 2), p.conv ? 0 : t.b2, p.conv ? 0 : t.b1);
              } else {
                for (int i = 0; i < nbox; ++i)
                  tma_load_4d_2sm(sb + i * 8192, &tma_b, full_leader, t.n * p.bn + (rank * nbox + i) * 64, k, t.b2, t.b1);
              }
            }
          } else {
            mbar_arrive_expect_tx(&full[stage], kABytes + b_bytes);
            if (p.conv == 2) {
              // dW: k-block = one bx x by (= 64) pixel tile of image b; A = dY (M = Cout), B = X shifted by the tap
              const int kbg = k / BK;
              const int b = kbg / p.tiles_per_img;
              const int r = kbg - b * p.tiles_per_img;
              const int px0 = (r % p.tiles_x) * p.bx, py0 = (r / p.tiles_x) * p.by;
              const int ky = t.b2 / 3, kx = t.b2 - ky * 3;
              tma_load_4d(sa, &tma_a, &full[stage], t.m * BM, px0, py0, b);
              tma_load_4d(sa + 8192, &tma_a, &full[stage], t.m * BM + 64, px0, py0, b);
              if (p.cs == 1) {
                for (int i = 0; i < p.bn / 64; ++i)
                  tma_load_4d(sb + i * 8192, &tma_b, &full[stage], t.n * p.bn + i * 64, px0 + kx - 1, py0 + ky - 1, b);
              } else {
                const int per = p.bn / 64 / p.cs;
                for (int i = rank * per; i < (rank + 1) * per; ++i)
                  tma_load_4d_mc(sb + i * 8192, &tma_b, &full[stage], t.n * p.bn + i * 64, px0 + kx - 1, py0 + ky - 1, b, mc_mask);
              }
            } else {
              if (p.conv) {
                const int tap = kb / p.cin_blocks;
                const int cb = kb - tap * p.cin_blocks;
                const int ky = tap / 3, kx = tap - ky * 3;
                tma_load_4d(sa, &tma_a, &full[stage], cb * BK, x0 + kx - 1, y0 + ky - 1, t.b1);
              } else if (!p.a_mn) {
                tma_load_4d(sa, &tma_a, &full[stage], k, t.m * BM, t.b2, t.b1);
              } else {
                tma_load_4d(sa, &tma_a, &full[stage], t.m * BM, k, t.b2, t.b1);
                tma_load_4d(sa + 8192, &tma_a, &full[stage], t.m * BM + 64, k, t.b2, t.b1);
              }
              if (!p.b_mn) {
                if (p.cs == 1) {
                  tma_load_4d(sb, &tma_b, &full[stage], k, t.n * p.bn, p.conv ? 0 : t.b2, p.conv ? 0 : t.b1);
                } else {
                  // this CTA fetches rows [rank*bn/cs, (rank+1)*bn/cs) of the shared B tile once and multicasts them
                  const int rows = p.bn / p.cs;
                  tma_load_4d_mc(sb + rank * rows * 128, &tma_b, &full[stage], k, t.n * p.bn + rank * rows, p.conv ? 0 : t.b2,
                                 p.conv ? 0 : t.b1, mc_mask);
                }
              } else if (p.cs == 1) {
                for (int i = 0; i < p.bn / 64; ++i)
                  tma_load_4d(sb + i * 8192, &tma_b, &full[stage], t.n * p.bn + i * 64, k, t.b2, t.b1);
              } else {
                const int per = p.bn / 64 / p.cs;
                for (int i = rank * per; i < (rank + 1) * per; ++i)
                  tma_load_4d_mc(sb + i * 8192, &tma_b, &full[stage], t.n * p.bn + i * 64, k, t.b2, t.b1, mc_mask);
              }
            }
          }
        }
        __syncwarp();
        TR(256 + trk - 1);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (whole warp in the loop, one elected lane issues) -------------------------------
    if (!kPair || rank == 0) {   // pair mode: only the leader CTA issues MMAs
      const uint32_t idesc = make_idesc_f16(kPair ? 2 * BM : BM, p.bn, p.a_mn != 0, p.b_mn != 0, p.bf16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int trk = 0;
      // K-major : 8-row groups 1024 B apart; +32 B per 16-element k-step inside the swizzle atom
      // MN-major: 64-wide MN blocks 8192 B apart (LBO), 8-row k groups 1024 B apart (SBO); +2048 B per 16-row k-step
      const uint32_t a_step = p.a_mn ? (2048u >> 4) : (32u >> 4);
      const uint32_t b_step = p.b_mn ? (2048u >> 4) : (32u >> 4);
      uint32_t ready = 0;       // the `full` barrier of the CURRENT stage was already seen complete by the previous k-block's probe
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
      const bool fused_issue = p.cs == 1 || kPair;     // multicast clusters (experiments only) keep the plain sequence
      for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
        const TileCoord t = decode_tile(p, tile, rank);
        const int k_begin = t.s * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + BK - 1) / BK;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxBN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (!ready) mbar_wait(&full[stage], phase);
          TR(512 + trk); ++trk;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t a_desc = p.a_mn ? make_smem_desc_sw128(sa, 8192, 1024) : make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t b_desc = p.b_mn ? make_smem_desc_sw128(sb, 8192, 1024) : make_smem_desc_sw128(sb, 16, 1024);
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == nstages) {
            nstage = 0;
            nphase ^= 1;
          }
          if (fused_issue && !DBG(8)) {
            // probe of the next stage + 4 MMAs + commits in one sequence (common.cuh: umma_kblock)
            ready = umma_kblock<kPair ? 2 : 1>(d_tmem, a_desc, b_desc, a_step, b_step, idesc, kb != 0 ? 1u : 0u,
                                                empty0 + stage * 8, smem_u32(&tmem_full[acc]), kb == nkb - 1 ? 1u : 0u,
                                                full0 + nstage * 8, nphase);
          } else {
            ready = 0;
            if (elect_one()) {
              if (!DBG(8)) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_f16_ss(d_tmem, a_desc + static_cast<uint64_t>(a_step * k), b_desc + static_cast<uint64_t>(b_step * k),
                              idesc, (kb | k) != 0);
              }
              if (kPair) {
                umma_commit_2cta_mc(&empty[stage], 3);
                if (kb == nkb - 1) umma_commit_2cta_mc(&tmem_full[acc], 3);
              } else {
                if (p.cs > 1) umma_commit_mc(&empty[stage], mc_mask); else umma_commit(&empty[stage]);
                if (kb == nkb - 1) umma_commit(&tmem_full[acc]);
              }
            }
          }
          __syncwarp();
          TR(768 + trk - 1);
          stage = nstage;
          phase = nphase;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------- epilogue -------------------------------
    // TMEM hands every thread one accumulator ROW (lane == row).  Storing from that mapping makes each warp store
    // touch 32 different cache lines (32 LSU wavefronts per instruction; measured 8-18 k clk per 128 x bn tile), so every
    // 32 x 32 chunk is transposed through shared memory: written row-per-lane with the 16-byte column chunk XOR-swizzled
    // by (row & 7), read back as {row = 4*it + lane/8, 4 columns = lane%8} — 8 lanes cover 128 contiguous bytes of one
    // output row, a warp instruction touches 4 lines.  All elementwise work happens in the coalesced mapping.
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const int egroup = (warp - 2) >> 2;  // which share of the column chunks this warp drains
    const uint32_t stg = smem_u32(smem + p.epi_off) + (warp - 2) * p.nbuf * kEpiBufBytes;   // shared-space address (1024-aligned)
    const int rr = lane >> 3, c4 = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    int trt = 0;
    uint64_t* const my_res = res_full + 3 * (warp - 2);
    uint32_t nst = 0;          // staging rounds of this warp so far (buffer = nst % nbuf); warp-uniform
    uint32_t res_uses[3] = {0, 0, 0};   // residual loads consumed per staging buffer (mbarrier phase bookkeeping)
    for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile, rank);
      if (p.epi_mode != 0) {
        // =========================== TMA epilogue ===========================
        // Everything happens in the TMEM mapping (lane == accumulator row): bias / GELU / GroupNorm statistics / residual
        // add on the registers tcgen05.ld delivers, the row is written into a SWIZZLE_128B staging buffer (32 rows x 128 B per
        // warp, conflict-free: 16-byte chunk q of row r sits at chunk q ^ (r & 7)) and ONE elected lane hands the 4 KB box to
        // the TMA unit (UTMASTG), which also clips row / column tails.  No transpose read-back, no per-row address
        // arithmetic, no LSU global stores; the next round's TMEM read overlaps the store.
        const int n0 = t.n * p.bn;
        const bool first_split = (t.s == 0);
        const int r_in_tile = quarter * 32 + lane;
        int c1, c2v, c3;       // tensor-map coordinates 1..3 of this warp's 32-row box
        bool row_ok;
        if (p.conv == 1) {
          const int x0 = (t.m % p.tiles_x) * p.bx, y0 = (t.m / p.tiles_x) * p.by;
          c1 = x0 + (quarter * 32) % p.bx;
          c2v = y0 + (quarter * 32) / p.bx;
          c3 = t.b1;
          row_ok = (x0 + r_in_tile % p.bx < p.W) && (y0 + r_in_tile / p.bx < p.H);
        } else {
          c1 = t.m * BM + quarter * 32;
          c2v = t.b2;
          c3 = t.b1;
          row_ok = c1 + lane < p.M;
        }
        const bool use_bias = p.bias != nullptr && first_split;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const uint32_t rowoff = static_cast<uint32_t>(lane) * 128u;
        if (p.epi_mode == 1) {
          if (tile == cluster_id) pdl_wait();   // bias comes from earlier kernels; C may still be read by them
          // bias of this warp's column chunks, one column per lane, fetched BEFORE the accumulator is waited for (the L2 round
          // trip hides behind the main loop); the row-per-lane arithmetic below reads column j's value with a shuffle
          // LayerNorm consumer: this lane's row statistics (eight fixed-order partials written by the producer GEMM) and the
          // column sums of W * diag(gamma), one column per lane like the bias
          const bool ln_in = p.ln_colsum != nullptr;
          float ln_rstd = 1.f, ln_nmr = 0.f;          // rstd and -rstd * mean of this lane's row
          float slane[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
          if (ln_in) {
            float s1 = 0.f, s2 = 0.f;
            if (row_ok) {
              const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + static_cast<long long>(c1 + lane) * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 q4 = sp[i];
                s1 += q4.x; s2 += q4.y; s1 += q4.z; s2 += q4.w;
              }
            }
            const float mean = s1 * p.ln_inv_dim;
            const float var = fmaxf(s2 * p.ln_inv_dim - mean * mean, 0.f);
            ln_rstd = rsqrtf(var + p.ln_eps);
            ln_nmr = -ln_rstd * mean;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int col0 = n0 + (egroup + 2 * i) * 64;
              if (egroup + 2 * i < p.bn / 64 && col0 < p.N) {
                slane[i][0] = __ldg(p.ln_colsum + col0 + lane);
                slane[i][1] = __ldg(p.ln_colsum + col0 + 32 + lane);
              }
            }
          }
          float blane[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
          if (use_bias) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int col0 = n0 + (egroup + 2 * i) * 64;
              if (egroup + 2 * i < p.bn / 64 && col0 < p.N) {
                blane[i][0] = __ldg(p.bias + col0 + lane);
                blane[i][1] = __ldg(p.bias + col0 + 32 + lane);
              }
            }
          }
          // act == 2 (GELU' of the saved pre-activation, fc2's dX): this warp's chunks of `aux` are TMA-loaded into its staging
          // buffers ahead of the accumulator (host: nbuf covers the chunks per warp) and the product is formed in place
          const bool aux_in = p.act == 2, aux_out = p.act == 1 && p.aux != nullptr;
          if (aux_in) {
            if (lane == 0) {
              bulk_wait_read<0>();            // the previous tile's stores have read the buffers
              for (int i = 0; i < 2; ++i) {
                const int cc = egroup + 2 * i;
                if (cc < p.bn / 64 && n0 + cc * 64 < p.N) {
                  mbar_arrive_expect_tx(&my_res[i], kEpiBufBytes);
                  tma_load_4d(smem + p.epi_off + ((warp - 2) * p.nbuf + i) * kEpiBufBytes, &tma_r, &my_res[i], n0 + cc * 64, c1, c2v, c3);
                }
              }
            }
            __syncwarp();
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          if (warp == 2) TR(1024 + 2 * trt);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kMaxBN;
          for (int cc = egroup; cc < p.bn / 64; cc += kEpiWarps / 4) {
            const int col0 = n0 + cc * 64;
            if (col0 >= p.N) break;
            const bool second = cc != egroup;
            const float bl0 = second ? blane[1][0] : blane[0][0], bl1 = second ? blane[1][1] : blane[0][1];
            const float sl0 = second ? slane[1][0] : slane[0][0], sl1 = second ? slane[1][1] : slane[0][1];
            uint32_t r[2][32];
            tmem_ld_32x32b_x32(t_row + cc * 64, r[0]);
            tmem_ld_32x32b_x32(t_row + cc * 64 + 32, r[1]);
            tmem_ld_wait();
            // plain: rotate the staging buffers; aux_in: chunk i lives in buffer i; aux_out: buffer 0 = C chunk, buffer 1 = aux chunk
            const uint32_t buf = (aux_in ? (second ? 1u : 0u) : aux_out ? 0u : (nst % p.nbuf)) * kEpiBufBytes + stg;
            const uint32_t buf2 = stg + kEpiBufBytes;
            if (aux_in) {
              const int bi = second ? 1 : 0;
              mbar_wait(&my_res[bi], res_uses[bi] & 1u);
              ++res_uses[bi];
            } else if (aux_out) {
              if (nst > 0) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
              }
            } else if (nst >= static_cast<uint32_t>(p.nbuf)) {      // the store that last used this buffer has read it
              if (lane == 0) { if (p.nbuf == 3) bulk_wait_read<2>(); else if (p.nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
              __syncwarp();
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float v[32];
              const float blh = hf == 0 ? bl0 : bl1;
              if (ln_in) {
                const float slh = hf == 0 ? sl0 : sl1;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  v[j] = fmaf(__uint_as_float(r[hf][j]), ln_rstd, fmaf(ln_nmr, __shfl_sync(0xffffffffu, slh, j), __shfl_sync(0xffffffffu, blh, j)));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[hf][j]), p.alpha, __shfl_sync(0xffffffffu, blh, j));
              }
              if (p.gn_stats != nullptr) {
                float s1 = 0.f, s2 = 0.f;
                if (row_ok) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) {
                    s1 += v[j];
                    s2 = fmaf(v[j], v[j], s2);
                  }
                }
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                if (lane == 0) {
                  double* st = p.gn_stats + (static_cast<long long>(t.b1) * (p.N / 32) + (col0 + hf * 32) / 32) * 2;
                  atomicAdd(st, static_cast<double>(s1));
                  atomicAdd(st + 1, static_cast<double>(s2));
                }
              }
              if (aux_out) {
                // the pre-activation (saved for the backward) goes out through the second staging buffer
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t w0 = pack_16b(v[8 * q], v[8 * q + 1], p.bf16), w1 = pack_16b(v[8 * q + 2], v[8 * q + 3], p.bf16);
                  const uint32_t w2 = pack_16b(v[8 * q + 4], v[8 * q + 5], p.bf16), w3 = pack_16b(v[8 * q + 6], v[8 * q + 7], p.bf16);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf2 + rowoff + (((hf * 4 + q) ^ sw) << 4)),
                               "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                               : "memory");
                }
              }
              if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 gq = gelu_erf2(make_float2(v[j], v[j + 1]));
                  v[j] = gq.x;
                  v[j + 1] = gq.y;
                }
              } else if (aux_in) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t a0, a1, a2, a3;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                               : "r"(buf + rowoff + (((hf * 4 + q) ^ sw) << 4))
                               : "memory");
                  const float2 g0 = gelu_erf_grad2(unpack_16b(a0, p.bf16)), g1 = gelu_erf_grad2(unpack_16b(a1, p.bf16));
                  const float2 g2 = gelu_erf_grad2(unpack_16b(a2, p.bf16)), g3 = gelu_erf_grad2(unpack_16b(a3, p.bf16));
                  v[8 * q] *= g0.x; v[8 * q + 1] *= g0.y; v[8 * q + 2] *= g1.x; v[8 * q + 3] *= g1.y;
                  v[8 * q + 4] *= g2.x; v[8 * q + 5] *= g2.y; v[8 * q + 6] *= g3.x; v[8 * q + 7] *= g3.y;
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t w0, w1, w2, w3;
                if (p.bf16) {
                  w0 = pack_16b(v[8 * q], v[8 * q + 1], 1); w1 = pack_16b(v[8 * q + 2], v[8 * q + 3], 1);
                  w2 = pack_16b(v[8 * q + 4], v[8 * q + 5], 1); w3 = pack_16b(v[8 * q + 6], v[8 * q + 7], 1);
                } else {
                  w0 = pack_16b(v[8 * q], v[8 * q + 1], 0); w1 = pack_16b(v[8 * q + 2], v[8 * q + 3], 0);
                  w2 = pack_16b(v[8 * q + 4], v[8 * q + 5], 0); w3 = pack_16b(v[8 * q + 6], v[8 * q + 7], 0);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + rowoff + (((hf * 4 + q) ^ sw) << 4)),
                             "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
              }
            }
            fence_proxy_async_smem();      // generic-proxy writes of this lane -> visible to the TMA unit
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tma_c, buf, col0, c1, c2v, c3);
              if (aux_out) tma_store_4d(&tma_r, buf2, col0, c1, c2v, c3);
              bulk_commit();
            }
            ++nst;
          }
        } else {
          // ---------------- fp32 tile: 32 columns (128 bytes per row) per round; optional residual (TMA-loaded into the
          // staging buffer ahead of time, added in place) and split-K accumulation (TMA reduce-add)
          const bool use_res = p.residual != nullptr && first_split;
          const int nchunks = min(p.bn, p.N - n0) / 32;
          const int mine = nchunks > egroup ? (nchunks - egroup + 1) / 2 : 0;       // chunks egroup, egroup + 2, ...
          if (tile == cluster_id) pdl_wait();   // residual / bias come from earlier kernels; C may still be read by them
          const uint32_t nst0 = nst;           // round counter at the start of this tile: chunk k uses buffer (nst0 + k) % nbuf
          auto load_res = [&](int k) {          // lane 0: residual chunk k of this warp -> its staging buffer
            const uint32_t b = (nst0 + static_cast<uint32_t>(k)) % p.nbuf;
            mbar_arrive_expect_tx(&my_res[b], kEpiBufBytes);
            tma_load_4d(smem + p.epi_off + ((warp - 2) * p.nbuf + b) * kEpiBufBytes, &tma_r, &my_res[b],
                        n0 + (egroup + 2 * k) * 32, c1, c2v, c3);
          };
          if (use_res && mine > 0) {            // issued BEFORE the accumulator is waited for: lands during the main loop
            if (lane == 0) {
              bulk_wait_read<0>();              // the previous tile's stores have read every buffer
              load_res(0);
              if (p.nbuf >= 2 && mine > 1) load_res(1);
              if (p.nbuf >= 3 && mine > 2) load_res(2);
            }
            __syncwarp();
          }
          // bias of this warp's chunks, one column per lane (see the 16-bit path)
          float blane[4] = {0.f, 0.f, 0.f, 0.f};
          if (use_bias) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < mine) blane[i] = __ldg(p.bias + n0 + (egroup + 2 * i) * 32 + lane);
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          if (warp == 2) TR(1024 + 2 * trt);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kMaxBN;
          float ln_s1 = 0.f, ln_s2 = 0.f;       // LayerNorm producer: sum / sum of squares of this lane's row over this warp's chunks
          for (int k = 0; k < mine; ++k) {
            const int col0 = n0 + (egroup + 2 * k) * 32;
            const uint32_t bsel = (nst0 + static_cast<uint32_t>(k)) % p.nbuf;
            const uint32_t buf = stg + bsel * kEpiBufBytes;
            const float blk_ = k == 0 ? blane[0] : k == 1 ? blane[1] : k == 2 ? blane[2] : blane[3];
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + (egroup + 2 * k) * 32, r);
            if (use_res) {
              // the store of round k-1 has read its buffer -> refill it with the residual of chunk k-1+nbuf
              if (k >= 1 && k - 1 + p.nbuf < mine) {
                if (lane == 0) {
                  bulk_wait_read<0>();
                  load_res(k - 1 + p.nbuf);
                }
                __syncwarp();
              }
              mbar_wait(&my_res[bsel], res_uses[bsel] & 1u);
              ++res_uses[bsel];
            } else if (nst0 + k >= static_cast<uint32_t>(p.nbuf)) {
              if (lane == 0) { if (p.nbuf == 3) bulk_wait_read<2>(); else if (p.nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
              __syncwarp();
            }
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, __shfl_sync(0xffffffffu, blk_, j));
            if (use_res) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 rv;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(rv.x), "=f"(rv.y), "=f"(rv.z), "=f"(rv.w)
                             : "r"(buf + rowoff + ((static_cast<uint32_t>(q) ^ sw) << 4))
                             : "memory");
                v[4 * q] += rv.x; v[4 * q + 1] += rv.y; v[4 * q + 2] += rv.z; v[4 * q + 3] += rv.w;
              }
            }
            if (p.ln_x16 != nullptr) {
              // the consumer GEMM reads these rows as its 16-bit A operand and normalises in its epilogue
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                ln_s1 += v[j];
                ln_s2 = fmaf(v[j], v[j], ln_s2);
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + rowoff + ((static_cast<uint32_t>(q) ^ sw) << 4)),
                           "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                           : "memory");
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.atomic) tma_reduce_add_4d(&tma_c, buf, col0, c1, c2v, c3);
              else tma_store_4d(&tma_c, buf, col0, c1, c2v, c3);
              bulk_commit();
            }
            if (p.ln_x16 != nullptr) {
              // 16-bit copy of the chunk, read back from the staging buffer (the TMA store only reads it too) in the transposed
              // mapping — 8 lanes cover 64 contiguous bytes of one row, 4 rows per instruction — instead of 32 scattered 16-byte
              // row pieces per instruction from the lane == row mapping
#pragma unroll 1
              for (int it = 0; it < 8; ++it) {
                const int rw = it * 4 + rr;
                float4 f;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                             : "r"(buf + rw * 128 + ((c4 ^ (rw & 7)) << 4))
                             : "memory");
                if (c1 + rw < p.M)
                  *reinterpret_cast<uint2*>(p.ln_x16 + static_cast<long long>(c1 + rw) * p.ld_x16 + col0 + c4 * 4) =
                      make_uint2(pack_16b(f.x, f.y, p.bf16), pack_16b(f.z, f.w, p.bf16));
              }
            }
          }
          nst = nst0 + static_cast<uint32_t>(mine);
          if (p.ln_x16 != nullptr && row_ok)
            *reinterpret_cast<float2*>(p.ln_stats + static_cast<long long>(c1 + lane) * 16 + (t.n * 2 + egroup) * 2) = make_float2(ln_s1, ln_s2);
        }
        // accumulator drained (it lives in registers / shared memory now): hand the TMEM stage back to the MMA warp
        if (warp == 2) TR(1025 + 2 * trt);
        ++trt;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
          else mbar_arrive(&tmem_empty[acc]);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      // =========================== LSU epilogue (every mode) ===========================
      // rows handled by this lane in the coalesced mapping: r_in_tile = quarter*32 + 4*it + rr, it = 0..7
      int rowi[8];            // logical row (residual / aux addressing); c_off = cbase + rowi * ldc
      uint32_t vmask = 0;
      const long long cbase = p.conv == 1 ? t.b1 * p.sc1 : t.b1 * p.sc1 + t.b2 * p.sc2;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r_in_tile = quarter * 32 + it * 4 + rr;
        bool ok;
        if (p.conv == 1) {
          const int x = (t.m % p.tiles_x) * p.bx + r_in_tile % p.bx;
          const int y = (t.m / p.tiles_x) * p.by + r_in_tile / p.bx;
          ok = (x < p.W) && (y < p.H);
          rowi[it] = y * p.W + x;
        } else {
          rowi[it] = t.m * BM + r_in_tile;
          ok = rowi[it] < p.M;
        }
        if (ok) vmask |= 1u << it;
      }
      const bool all_valid = __all_sync(0xffffffffu, vmask == 0xffu);
      const int n0 = t.n * p.bn;
      const bool first_split = (t.s == 0);
      const int nchunks = p.bn / 32;

      // The residual does not depend on the MMA: fetch this warp's first chunk of it BEFORE waiting for the
      // accumulator, and each following chunk one iteration ahead, so the global-load latency hides behind the
      // main loop / the staging round trip instead of serialising the epilogue.
      const bool use_res = p.residual != nullptr && first_split;
      float4 rpre[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) rpre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      auto fetch_res = [&](int c) {
        const int col = n0 + c * 32 + c4 * 4;
        if (use_res && c < nchunks && col + 4 <= p.N && !DBG(2)) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if ((vmask >> it) & 1u) {
              const long long rrow = p.res_mod > 0 ? (rowi[it] % p.res_mod) : rowi[it];
              rpre[it] = *reinterpret_cast<const float4*>(p.residual + rrow * p.ldr + col);
            }
        }
      };
      if (tile == cluster_id) pdl_wait();   // first tile: residual / bias / aux come from earlier kernels, C may still be read by them
      fetch_res(egroup);

      mbar_wait(&tmem_full[acc], acc_phase);
      if (warp == 2) TR(1024 + 2 * trt);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kMaxBN;

      // ---------------- 16-bit output, whole tile valid: 64 columns per round.  Bias / GELU / GroupNorm statistics are
      // applied in the TMEM mapping (lane == row), the row is packed to 16 bit BEFORE the transpose — 128 bytes per
      // row: half the staging traffic of the fp32 path, and each read-back instruction (8 lanes x 16 B per row, 4 rows)
      // stores four complete 128-byte lines.
      if (p.epi64 && !p.out_f32 && p.aux == nullptr && p.residual == nullptr && !p.atomic && (p.bn & 63) == 0 && all_valid &&
          n0 + p.bn <= p.N) {
        const bool use_bias = p.bias != nullptr && first_split;
        uint16_t* cb = reinterpret_cast<uint16_t*>(p.C) + cbase + c4 * 8;
        int trc = 0;
        for (int c2 = egroup; c2 < p.bn / 64; c2 += kEpiWarps / 4) {
          const int col0 = n0 + c2 * 64;
          if (warp == 2 && trt == 0) TR(1200 + 4 * trc);
          // bias of the first 32 columns: issued before the TMEM read (same address in every lane); the second half is
          // fetched while the first is being processed
          float4 bq[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (use_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
          }
          uint32_t r[2][32];
          tmem_ld_32x32b_x32(t_row + c2 * 64, r[0]);
          tmem_ld_32x32b_x32(t_row + c2 * 64 + 32, r[1]);
          tmem_ld_wait();
          if (warp == 2 && trt == 0) TR(1201 + 4 * trc);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] = fmaf(__uint_as_float(r[hf][4 * j]), p.alpha, bq[j].x);
              v[4 * j + 1] = fmaf(__uint_as_float(r[hf][4 * j + 1]), p.alpha, bq[j].y);
              v[4 * j + 2] = fmaf(__uint_as_float(r[hf][4 * j + 2]), p.alpha, bq[j].z);
              v[4 * j + 3] = fmaf(__uint_as_float(r[hf][4 * j + 3]), p.alpha, bq[j].w);
            }
            if (hf == 0 && use_bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 32 + 4 * j));
            }
            if (p.gn_stats != nullptr) {
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                s1 += v[j];
                s2 = fmaf(v[j], v[j], s2);
              }
              s1 = warp_sum(s1);
              s2 = warp_sum(s2);
              if (lane == 0) {
                double* st = p.gn_stats + (static_cast<long long>(t.b1) * (p.N / 32) + (col0 + hf * 32) / 32) * 2;
                atomicAdd(st, static_cast<double>(s1));
                atomicAdd(st + 1, static_cast<double>(s2));
              }
            }
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float2 gq = gelu_erf2(make_float2(v[j], v[j + 1]));
                v[j] = gq.x;
                v[j + 1] = gq.y;
              }
            }
            // this lane's row: 16-byte chunk q of the 128-byte row goes to chunk q ^ (lane & 7)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t w0, w1, w2, w3;
              if (p.bf16) {
                w0 = pack_16b(v[8 * q], v[8 * q + 1], 1); w1 = pack_16b(v[8 * q + 2], v[8 * q + 3], 1);
                w2 = pack_16b(v[8 * q + 4], v[8 * q + 5], 1); w3 = pack_16b(v[8 * q + 6], v[8 * q + 7], 1);
              } else {
                w0 = pack_16b(v[8 * q], v[8 * q + 1], 0); w1 = pack_16b(v[8 * q + 2], v[8 * q + 3], 0);
                w2 = pack_16b(v[8 * q + 4], v[8 * q + 5], 0); w3 = pack_16b(v[8 * q + 6], v[8 * q + 7], 0);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + (((hf * 4 + q) ^ (lane & 7)) << 4)),
                           "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                           : "memory");
            }
          }
          __syncwarp();
          if (warp == 2 && trt == 0) TR(1202 + 4 * trc);
          uint4 o[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rw = it * 4 + rr;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o[it].x), "=r"(o[it].y), "=r"(o[it].z), "=r"(o[it].w)
                         : "r"(stg + rw * 128 + ((c4 ^ (rw & 7)) << 4))
                         : "memory");
          }
          __syncwarp();   // the next round overwrites the staging buffer
          if (!DBG(1)) {
#pragma unroll
            for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(cb + static_cast<long long>(rowi[it]) * p.ldc + col0) = o[it];
          }
          if (warp == 2 && trt == 0) TR(1203 + 4 * trc);
          ++trc;
        }
      } else
      {
      int trc = 0;
      for (int c = egroup; c < nchunks; c += kEpiWarps / 4) {
        const int col0 = n0 + c * 32;           // first column of the chunk
        const int col = col0 + c4 * 4;          // this lane's 4 columns
        const int ncols = min(32, p.N - col0);  // valid columns of the chunk (<= 0: nothing)
        const int nmine = min(4, p.N - col);    // valid columns of this lane
        const bool full4 = nmine == 4;
        const bool fast = ncols == 32 && all_valid;
        const bool use_bias = p.bias != nullptr && first_split;
        // issued before the TMEM read so its L2 latency hides behind the read and the transpose
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fast && use_bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        uint32_t r[32];
        if (warp == 2 && trt == 0) TR(1200 + 4 * trc);
        tmem_ld_32x32b_x32(t_row + c * 32, r);
        tmem_ld_wait();
        if (warp == 2 && trt == 0) TR(1201 + 4 * trc);
        // transpose: lane == row  ->  8 lanes per row
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((j ^ (lane & 7)) << 4)), "r"(r[4 * j]),
                       "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        __syncwarp();
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rw = it * 4 + rr;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[it].x), "=f"(v[it].y), "=f"(v[it].z), "=f"(v[it].w)
                       : "r"(stg + rw * 128 + ((c4 ^ (rw & 7)) << 4))
                       : "memory");
        }
        __syncwarp();   // the next chunk overwrites the staging buffer
        if (warp == 2 && trt == 0) TR(1202 + 4 * trc);


        if (fast) {
          // ---------------- fast path: full 32 x 32 chunk.  Every mode test is hoisted out of the row loops
          // (the modes are kernel-uniform), so each row costs one address IMAD plus the loads / stores themselves.
          if (p.alpha != 1.0f) {
#pragma unroll
            for (int it = 0; it < 8; ++it) { v[it].x *= p.alpha; v[it].y *= p.alpha; v[it].z *= p.alpha; v[it].w *= p.alpha; }
          }
          if (use_bias) {
#pragma unroll
            for (int it = 0; it < 8; ++it) { v[it].x += bb.x; v[it].y += bb.y; v[it].z += bb.z; v[it].w += bb.w; }
          }
          if (p.gn_stats != nullptr) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              s1 += (v[it].x + v[it].y) + (v[it].z + v[it].w);
              s2 += (v[it].x * v[it].x + v[it].y * v[it].y) + (v[it].z * v[it].z + v[it].w * v[it].w);
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) {
              double* st = p.gn_stats + (static_cast<long long>(t.b1) * (p.N / 32) + col0 / 32) * 2;
              atomicAdd(st, static_cast<double>(s1));
              atomicAdd(st + 1, static_cast<double>(s2));
            }
          }
          if (p.act == 1) {
            if (p.aux != nullptr) {
              uint16_t* ab = reinterpret_cast<uint16_t*>(p.aux) + col;
              if (p.bf16) {
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  *reinterpret_cast<uint2*>(ab + static_cast<long long>(rowi[it]) * p.ldaux) =
                      make_uint2(pack_16b(v[it].x, v[it].y, 1), pack_16b(v[it].z, v[it].w, 1));
              } else {
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  *reinterpret_cast<uint2*>(ab + static_cast<long long>(rowi[it]) * p.ldaux) =
                      make_uint2(pack_16b(v[it].x, v[it].y, 0), pack_16b(v[it].z, v[it].w, 0));
              }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const float2 g0 = gelu_erf2(make_float2(v[it].x, v[it].y)), g1 = gelu_erf2(make_float2(v[it].z, v[it].w));
              v[it] = make_float4(g0.x, g0.y, g1.x, g1.y);
            }
          } else if (p.act == 2) {
            const uint16_t* ab = reinterpret_cast<const uint16_t*>(p.aux) + col;
            uint2 a[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) a[it] = *reinterpret_cast<const uint2*>(ab + static_cast<long long>(rowi[it]) * p.ldaux);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const float2 g0 = gelu_erf_grad2(unpack_16b(a[it].x, p.bf16)), g1 = gelu_erf_grad2(unpack_16b(a[it].y, p.bf16));
              v[it].x *= g0.x; v[it].y *= g0.y; v[it].z *= g1.x; v[it].w *= g1.y;
            }
          }
          if (use_res) {
#pragma unroll
            for (int it = 0; it < 8; ++it) { v[it].x += rpre[it].x; v[it].y += rpre[it].y; v[it].z += rpre[it].z; v[it].w += rpre[it].w; }
          }
          // next chunk's residual: issued BEFORE this chunk's stores (its registers are free now; behind the stores the
          // loads would first wait for the store queue to drain) — it lands during the stores, the next TMEM read and transpose
          fetch_res(c + kEpiWarps / 4);
          if (DBG(1)) {
          } else if (p.out_f32) {
            float* cb = reinterpret_cast<float*>(p.C) + cbase + col;
            if (p.atomic) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cb + static_cast<long long>(rowi[it]) * p.ldc),
                             "f"(v[it].x), "f"(v[it].y), "f"(v[it].z), "f"(v[it].w)
                             : "memory");
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(cb + static_cast<long long>(rowi[it]) * p.ldc) = v[it];
            }
          } else {
            uint16_t* cb = reinterpret_cast<uint16_t*>(p.C) + cbase + col;
            if (p.bf16) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                *reinterpret_cast<uint2*>(cb + static_cast<long long>(rowi[it]) * p.ldc) =
                    make_uint2(pack_16b(v[it].x, v[it].y, 1), pack_16b(v[it].z, v[it].w, 1));
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                *reinterpret_cast<uint2*>(cb + static_cast<long long>(rowi[it]) * p.ldc) =
                    make_uint2(pack_16b(v[it].x, v[it].y, 0), pack_16b(v[it].z, v[it].w, 0));
            }
          }
          if (warp == 2 && trt == 0) TR(1203 + 4 * trc);
          ++trc;
          continue;
        }
        // ---------------- generic path: row / column tails
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          v[it].x *= p.alpha; v[it].y *= p.alpha; v[it].z *= p.alpha; v[it].w *= p.alpha;
        }
        if (p.bias != nullptr && first_split && nmine > 0) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (full4) {
            bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          } else {
            bb.x = __ldg(p.bias + col);
            if (nmine > 1) bb.y = __ldg(p.bias + col + 1);
            if (nmine > 2) bb.z = __ldg(p.bias + col + 2);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            v[it].x += bb.x; v[it].y += bb.y; v[it].z += bb.z; v[it].w += bb.w;
          }
        }

        if (p.gn_stats != nullptr) {
          // GroupNorm statistics of this 32-channel group over the 32 pixels of this warp (N % 32 == 0 in this mode).
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if ((vmask >> it) & 1u) {
              s1 += (v[it].x + v[it].y) + (v[it].z + v[it].w);
              s2 += (v[it].x * v[it].x + v[it].y * v[it].y) + (v[it].z * v[it].z + v[it].w * v[it].w);
            }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          if (lane == 0 && ncols > 0) {
            double* st = p.gn_stats + (static_cast<long long>(t.b1) * (p.N / 32) + col0 / 32) * 2;
            atomicAdd(st, static_cast<double>(s1));
            atomicAdd(st + 1, static_cast<double>(s2));
          }
        }

        if (nmine > 0) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (!((vmask >> it) & 1u)) continue;
            float4 o = v[it];
            if (p.act == 1) {
              if (p.aux != nullptr) {
                uint16_t* ap = reinterpret_cast<uint16_t*>(p.aux) + static_cast<long long>(rowi[it]) * p.ldaux + col;
                if (full4) {
                  *reinterpret_cast<uint2*>(ap) = make_uint2(pack_16b(o.x, o.y, p.bf16), pack_16b(o.z, o.w, p.bf16));
                } else {
                  const float e[4] = {o.x, o.y, o.z, o.w};
                  for (int j = 0; j < nmine; ++j) ap[j] = static_cast<uint16_t>(pack_16b(e[j], 0.f, p.bf16) & 0xffffu);
                }
              }
              const float2 g0 = gelu_erf2(make_float2(o.x, o.y)), g1 = gelu_erf2(make_float2(o.z, o.w));
              o = make_float4(g0.x, g0.y, g1.x, g1.y);
            } else if (p.act == 2) {
              const uint16_t* ap = reinterpret_cast<const uint16_t*>(p.aux) + static_cast<long long>(rowi[it]) * p.ldaux + col;
              if (full4) {
                const uint2 a = *reinterpret_cast<const uint2*>(ap);
                const float2 g0 = gelu_erf_grad2(unpack_16b(a.x, p.bf16)), g1 = gelu_erf_grad2(unpack_16b(a.y, p.bf16));
                o.x *= g0.x; o.y *= g0.y; o.z *= g1.x; o.w *= g1.y;
              } else {
                float e[4] = {o.x, o.y, o.z, o.w};
                for (int j = 0; j < nmine; ++j) e[j] *= gelu_erf_grad(unpack_16b(ap[j], p.bf16).x);
                o = make_float4(e[0], e[1], e[2], e[3]);
              }
            }
            if (use_res) {
              if (full4) {
                o.x += rpre[it].x; o.y += rpre[it].y; o.z += rpre[it].z; o.w += rpre[it].w;
              } else {
                const long long rrow = p.res_mod > 0 ? (rowi[it] % p.res_mod) : rowi[it];
                const float* rp = p.residual + rrow * p.ldr + col;
                o.x += rp[0];
                if (nmine > 1) o.y += rp[1];
                if (nmine > 2) o.z += rp[2];
              }
            }
            const long long c_off = cbase + static_cast<long long>(rowi[it]) * p.ldc + col;
            if (p.out_f32) {
              float* cp = reinterpret_cast<float*>(p.C) + c_off;
              if (p.atomic) {
                if (full4) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                               : "memory");
                } else {
                  const float e[4] = {o.x, o.y, o.z, o.w};
                  for (int j = 0; j < nmine; ++j) atomicAdd(cp + j, e[j]);
                }
              } else if (full4) {
                *reinterpret_cast<float4*>(cp) = o;
              } else {
                const float e[4] = {o.x, o.y, o.z, o.w};
                for (int j = 0; j < nmine; ++j) cp[j] = e[j];
              }
            } else {
              uint16_t* cp = reinterpret_cast<uint16_t*>(p.C) + c_off;
              if (full4) {
                *reinterpret_cast<uint2*>(cp) = make_uint2(pack_16b(o.x, o.y, p.bf16), pack_16b(o.z, o.w, p.bf16));
              } else {
                const float e[4] = {o.x, o.y, o.z, o.w};
                for (int j = 0; j < nmine; ++j) cp[j] = static_cast<uint16_t>(pack_16b(e[j], 0.f, p.bf16) & 0xffffu);
              }
            }
          }
        }
        if (warp == 2 && trt == 0) TR(1203 + 4 * trc);
        ++trc;
        fetch_res(c + kEpiWarps / 4);   // lands while the next chunk is read from TMEM and transposed
      }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      if (warp == 2) TR(1025 + 2 * trt);
      ++trt;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));   // the leader owns the MMA
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // the staging buffers must outlive the TMA unit's READS of them; the global writes themselves complete asynchronously and
  // are ordered before grid completion like any other store (same contract as CUTLASS's tma_store_wait)
  if (p.epi_mode != 0 && warp >= 2 && lane == 0) bulk_wait_read<0>();
  tc_fence_before();
  if (p.cs > 1) cluster_sync_all(); else __syncthreads();   // no peer may still signal this CTA's barriers after it exits
  if (threadIdx.x == 0) TR(3);
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2cta<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
  }
}

int pick_bn(int m_tiles, int N, int other, int b_mn, int sms) {
  // Choose the N tile that minimises (waves * tile cost); larger tiles win ties (more operand reuse).
  int best = 0;
  double best_cost = 1e30;
  const int cands[4] = {256, 192, 128, 64};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (b_mn && (bn % 64)) continue;
    if (bn > 64 && bn >= 2 * N && N > 0) continue;  // grossly oversized
    const long long n_tiles = (N + bn - 1) / bn;
    const long long tiles = n_tiles * m_tiles * other;
    const long long waves = (tiles + sms - 1) / sms;
    // per-tile cost ~ MMA time (bn) plus a fixed per-tile overhead
    const double cost = static_cast<double>(waves) * (bn + 24);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace

}  // namespace countr

#ifdef COUNTR_TRACE
extern "C" int countr_debug_set_knobs(int knobs) {
  COUNTR_CHECK_CUDA(cudaMemcpyToSymbol(countr::g_dbg, &knobs, sizeof(knobs)));
  return 0;
}
extern "C" int countr_debug_set_trace(void* buf, int cta) {
  long long* pbuf = reinterpret_cast<long long*>(buf);
  COUNTR_CHECK_CUDA(cudaMemcpyToSymbol(countr::g_trace, &pbuf, sizeof(pbuf)));
  COUNTR_CHECK_CUDA(cudaMemcpyToSymbol(countr::g_trace_cta, &cta, sizeof(cta)));
  return 0;
}
#endif

extern "C" int countr_gemm(const countr_gemm_desc* d, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(d != nullptr, "null descriptor");
  COUNTR_REQUIRE(d->a && d->b && d->c, "null operand pointer");
  COUNTR_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "bad GEMM shape M=%d N=%d K=%d", d->M, d->N, d->K);
  const int nb1 = d->nb1 > 0 ? d->nb1 : 1, nb2 = d->nb2 > 0 ? d->nb2 : 1;
  const bool conv_dw = d->conv_h > 0 && d->conv_dw != 0;
  const bool conv = d->conv_h > 0 && !conv_dw;
  const int sms = num_sms();
  COUNTR_REQUIRE(sms > 0, "no CUDA device");

  GemmArgs p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.nb2 = nb2;
  p.a_mn = d->a_mn; p.b_mn = d->b_mn; p.bf16 = d->bf16;
  p.conv = conv ? 1 : 0;
  if (conv) {
    COUNTR_REQUIRE(!d->a_mn && !d->b_mn, "conv mode needs K-major operands");
    COUNTR_REQUIRE(d->conv_cin % BK == 0, "conv Cin=%d must be a multiple of %d", d->conv_cin, BK);
    COUNTR_REQUIRE(d->conv_bx * d->conv_by == BM, "conv tile %dx%d must cover %d pixels", d->conv_bx, d->conv_by, BM);
    COUNTR_REQUIRE(d->K == 9 * d->conv_cin, "conv K=%d must equal 9*Cin", d->K);
    p.H = d->conv_h; p.W = d->conv_w; p.bx = d->conv_bx; p.by = d->conv_by;
    p.cin_blocks = d->conv_cin / BK;
    p.tiles_x = (p.W + p.bx - 1) / p.bx;
    p.m_tiles = p.tiles_x * ((p.H + p.by - 1) / p.by);
  } else {
    p.m_tiles = (d->M + BM - 1) / BM;
  }
  if (conv_dw) {
    // M = Cout, N = Cin, K = (#pixel tiles) * 64, nb2 = 9 taps; both operands MN-major NHWC
    COUNTR_REQUIRE(d->a_mn && d->b_mn && nb2 == 9 && nb1 == 1, "conv dW mode needs MN-major operands and nb2 == 9");
    COUNTR_REQUIRE(d->conv_bx * d->conv_by == BK, "conv dW pixel tile %dx%d must cover %d pixels", d->conv_bx, d->conv_by, BK);
    COUNTR_REQUIRE(d->M % 64 == 0 && d->N % 64 == 0, "conv dW needs Cout, Cin multiples of 64");
    p.conv = 2;
    p.H = d->conv_h; p.W = d->conv_w; p.bx = d->conv_bx; p.by = d->conv_by;
    p.tiles_x = (p.W + p.bx - 1) / p.bx;
    p.tiles_per_img = p.tiles_x * ((p.H + p.by - 1) / p.by);
    COUNTR_REQUIRE(d->K == d->conv_batch * p.tiles_per_img * BK, "conv dW K=%d must equal batch*tiles*64", d->K);
  }
  const int split_k = d->split_k > 1 ? d->split_k : 1;
  COUNTR_REQUIRE(split_k == 1 || (d->atomic && d->out_f32), "split_k > 1 needs atomic fp32 output");
  COUNTR_REQUIRE(!d->atomic || (d->out_f32 && d->act == 0), "atomic output must be fp32 without activation");
  int kps = ((d->K + split_k - 1) / split_k + BK - 1) / BK * BK;
  p.k_per_split = kps;
  p.split_k = (d->K + kps - 1) / kps;  // drop empty splits
  int bn = d->bn;
  if (bn <= 0) bn = pick_bn(p.m_tiles, d->N, p.split_k * nb1 * nb2, d->b_mn, sms);
  COUNTR_REQUIRE(bn >= 32 && bn <= kMaxBN && bn % 32 == 0 && (!d->b_mn || bn % 64 == 0), "bad N tile %d", bn);
  p.bn = bn;
  p.n_tiles = (d->N + bn - 1) / bn;
  // Optional cluster of `cs` CTAs on consecutive m tiles of the same n tile: the B tile is fetched from L2 once
  // per cluster (each CTA loads 1/cs of it and multicasts).  Measured on B200 (profiles/r1_bench_gemm_v4.log):
  // no gain at cs=2 and a loss at cs=4 — the kernel is NOT L2->SM bound (the limiter at 128 x 256 tiles is
  // shared-memory bandwidth: TMA fill + UMMA operand reads = 192 B/clk against 128 B/clk), so the default is
  // cs = 1 and the path is kept for experiments (cta_group::2 is the real fix).
  // CTA pair (cta_group::2): K-major B only; the pair computes a 256 x bn tile with half of B per SM.
  const bool pair_ok = p.m_tiles >= 2 && ((d->b_mn || conv_dw) ? (bn % 128 == 0) : (bn % 32 == 0));
  // auto policy (cta_pair == 0): pairs for the big implicit-GEMM convolutions, where the larger operand reuse
  // (32 KB instead of 48 KB of smem fill per k-block and SM) is worth +13 % (1.10 -> 1.24 PF/s on decode_head3);
  // neutral on the mid-size Linear layers, which are latency- not throughput-bound (profiles/r1_bench_gemm_pair.log).
  // ... and for long-K Linear layers whose tile list is a single wave of CTA pairs (the encoder's fc2: K = 3072, 72 pairs of
  // 256 x 192): the pair's main loop runs at the tensor floor (384 clk per k-block against 456 for one CTA), which outweighs
  // its longer prologue once there are ~48 k-blocks (fp32 + residual: 22.0 -> 20.0 us; no gain at K <= 2048)
  const bool pair_auto = (bn == 256 && ((conv && static_cast<long long>(p.m_tiles) * nb1 >= 128) ||
                                        (conv_dw && p.m_tiles % 2 == 0 && d->K >= 64 * 64))) ||
                         (!conv && !conv_dw && !d->a_mn && !d->b_mn && bn == 192 && d->K >= 3072 && p.m_tiles % 2 == 0 && split_k == 1 &&
                          nb1 * nb2 == 1 && (p.m_tiles / 2) * ((d->N + bn - 1) / bn) <= sms / 2);
  // large-M Linear layers run as CTA pairs (less shared-memory fill and L2 traffic per flop; these long kernels are power-limited);
  // COUNTR_PAIR_LINEAR=<min m_tiles> moves the threshold (0 = never)
  static int pair_linear = -1;
  if (pair_linear < 0) {
    const char* e = getenv("COUNTR_PAIR_LINEAR");
    pair_linear = e != nullptr ? atoi(e) : 64;     // default: from 64 m tiles on (M >= 8192: B = 128 inference 33.6 -> 32.8 ms, pre-train step 18.27 -> 17.91 ms)
  }
  const bool pair_forced = pair_linear > 0 && !conv && !conv_dw && !d->a_mn && !d->b_mn && p.m_tiles >= pair_linear && p.m_tiles % 2 == 0 &&
                           split_k == 1 && nb1 * nb2 == 1;
  p.pair = (pair_ok && (d->cta_pair > 0 || (d->cta_pair == 0 && (pair_auto || pair_forced)))) ? 1 : 0;
  int cs = p.pair ? 2 : (d->cluster > 0 ? d->cluster : 1);
  if (cs != 1 && cs != 2 && cs != 4) cs = 1;
  while (cs > 1 && (p.m_tiles < cs || (d->b_mn || conv_dw ? (bn / 64) % cs != 0 : (bn % (8 * cs)) != 0))) cs >>= 1;
  p.cs = cs;
  p.m_groups = (p.m_tiles + cs - 1) / cs;
  p.total_tiles = p.m_groups * p.n_tiles * p.split_k * nb1 * nb2;
  {
    auto magic = [](int d) { return d > 1 ? static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(d)) + 1u : 0u; };
    p.mg_magic = magic(p.m_groups); p.nt_magic = magic(p.n_tiles); p.sk_magic = magic(p.split_k); p.nb2_magic = magic(nb2);
    COUNTR_REQUIRE(static_cast<long long>(p.total_tiles) * std::max(std::max(p.m_groups, p.n_tiles), std::max(p.split_k, nb2)) < (1ll << 32),
                   "tile count %d too large for the fast tile decode", p.total_tiles);
  }

  p.C = d->c; p.ldc = d->ldc; p.sc1 = d->sc1; p.sc2 = d->sc2;
  p.out_f32 = d->out_f32; p.atomic = d->atomic;
  p.alpha = d->alpha;
  p.bias = d->bias;
  p.act = d->act; p.aux = d->aux; p.ldaux = d->ldaux;
  p.residual = d->residual; p.ldr = d->ldr; p.res_mod = d->res_mod;
  p.gn_stats = d->gn_stats;
  p.ln_x16 = reinterpret_cast<uint16_t*>(d->ln_x16); p.ld_x16 = d->ld_x16; p.ln_stats = d->ln_stats; p.ln_colsum = d->ln_colsum;
  p.ln_inv_dim = d->ln_dim > 0 ? 1.0f / static_cast<float>(d->ln_dim) : 0.f; p.ln_eps = d->ln_eps;
  {
    static int epi64 = -1;
    if (epi64 < 0) {
      const char* e = getenv("COUNTR_EPI64");
      epi64 = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    p.epi64 = epi64;
  }
  COUNTR_REQUIRE(d->act != 2 || d->aux != nullptr, "act=2 needs aux");
  COUNTR_REQUIRE(d->ldc > 0, "ldc must be positive");
  COUNTR_REQUIRE(d->bias == nullptr || (reinterpret_cast<uintptr_t>(d->bias) & 15u) == 0, "bias must be 16-byte aligned");
  const int c_align = d->out_f32 ? 4 : 8;
  COUNTR_REQUIRE(d->ldc % c_align == 0 && (reinterpret_cast<uintptr_t>(d->c) & 15u) == 0,
                 "C must be 16-byte aligned with ldc %% %d == 0", c_align);
  COUNTR_REQUIRE(d->gn_stats == nullptr || (conv && !conv_dw && d->N % 32 == 0), "gn_stats needs conv mode with N %% 32 == 0");

  CUtensorMap ta, tb;
  int rc;
  if (conv_dw) {
    const uint64_t dims[4] = {(uint64_t)d->M, (uint64_t)d->conv_w, (uint64_t)d->conv_h, (uint64_t)d->conv_batch};
    const uint64_t str[4] = {1, (uint64_t)d->lda, (uint64_t)d->lda * d->conv_w, (uint64_t)d->lda * d->conv_w * d->conv_h};
    const uint32_t box[4] = {64, (uint32_t)d->conv_bx, (uint32_t)d->conv_by, 1};
    rc = make_tmap_4d_16b(&ta, d->a, dims, str, box, TMAP_SW_128);
  } else if (conv) {
    const uint64_t dims[4] = {(uint64_t)d->conv_cin, (uint64_t)d->conv_w, (uint64_t)d->conv_h, (uint64_t)nb1};
    const uint64_t str[4] = {1, (uint64_t)d->lda, (uint64_t)d->lda * d->conv_w, (uint64_t)d->sa1};
    const uint32_t box[4] = {BK, (uint32_t)d->conv_bx, (uint32_t)d->conv_by, 1};
    rc = make_tmap_4d_16b(&ta, d->a, dims, str, box, TMAP_SW_128);
  } else if (!d->a_mn) {
    const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->M, (uint64_t)nb2, (uint64_t)nb1};
    const uint64_t str[4] = {1, (uint64_t)d->lda, (uint64_t)(nb2 > 1 ? d->sa2 : d->lda), (uint64_t)(nb1 > 1 ? d->sa1 : d->lda)};
    const uint32_t box[4] = {BK, BM, 1, 1};
    rc = make_tmap_4d_16b(&ta, d->a, dims, str, box, TMAP_SW_128);
  } else {
    const uint64_t dims[4] = {(uint64_t)d->M, (uint64_t)d->K, (uint64_t)nb2, (uint64_t)nb1};
    const uint64_t str[4] = {1, (uint64_t)d->lda, (uint64_t)(nb2 > 1 ? d->sa2 : d->lda), (uint64_t)(nb1 > 1 ? d->sa1 : d->lda)};
    const uint32_t box[4] = {64, BK, 1, 1};
    rc = make_tmap_4d_16b(&ta, d->a, dims, str, box, TMAP_SW_128);
  }
  if (rc) return rc;
  const int bnb1 = conv ? 1 : nb1, bnb2 = conv ? 1 : nb2;
  if (conv_dw) {
    const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->conv_w, (uint64_t)d->conv_h, (uint64_t)d->conv_batch};
    const uint64_t str[4] = {1, (uint64_t)d->ldb, (uint64_t)d->ldb * d->conv_w, (uint64_t)d->ldb * d->conv_w * d->conv_h};
    const uint32_t box[4] = {64, (uint32_t)d->conv_bx, (uint32_t)d->conv_by, 1};
    rc = make_tmap_4d_16b(&tb, d->b, dims, str, box, TMAP_SW_128);
  } else if (!d->b_mn) {
    const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->N, (uint64_t)bnb2, (uint64_t)bnb1};
    const uint64_t str[4] = {1, (uint64_t)d->ldb, (uint64_t)(bnb2 > 1 ? d->sb2 : d->ldb), (uint64_t)(bnb1 > 1 ? d->sb1 : d->ldb)};
    const uint32_t box[4] = {BK, (uint32_t)(bn / p.cs), 1, 1};
    rc = make_tmap_4d_16b(&tb, d->b, dims, str, box, TMAP_SW_128);
  } else {
    const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->K, (uint64_t)bnb2, (uint64_t)bnb1};
    const uint64_t str[4] = {1, (uint64_t)d->ldb, (uint64_t)(bnb2 > 1 ? d->sb2 : d->ldb), (uint64_t)(bnb1 > 1 ? d->sb1 : d->ldb)};
    const uint32_t box[4] = {64, BK, 1, 1};
    rc = make_tmap_4d_16b(&tb, d->b, dims, str, box, TMAP_SW_128);
  }
  if (rc) return rc;

  // ---- epilogue mode: TMA stores wherever the tile is plain (see the kernel), LSU stores for the rest
  static int epi_tma = -1;
  if (epi_tma < 0) {
    const char* e = getenv("COUNTR_EPI_TMA");
    epi_tma = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const int c_elt = d->out_f32 ? 4 : 2;
  const long long sc1e = d->sc1, sc2e = d->sc2;
  const bool c_strides_ok = (d->ldc * c_elt) % 16 == 0 && (nb1 <= 1 || conv || (sc1e * c_elt) % 16 == 0) &&
                            (nb2 <= 1 || conv || (sc2e * c_elt) % 16 == 0) && (!conv || (sc1e * c_elt) % 16 == 0);
  p.epi_mode = 0;
  // aux (fc1's saved pre-activation: an extra 16-bit output with act 1, an extra 16-bit input with act 2) also goes through TMA when
  // it is a plain [M, N] matrix
  static int epi_aux = -1;
  if (epi_aux < 0) {
    const char* e = getenv("COUNTR_EPI_AUX");
    epi_aux = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool aux_tma = d->aux == nullptr || (epi_aux && !d->out_f32 && !conv && nb1 * nb2 == 1 && (d->ldaux * 2) % 16 == 0 &&
                                              (reinterpret_cast<uintptr_t>(d->aux) & 15u) == 0 && (d->act == 1 || d->act == 2));
  if (epi_tma && c_strides_ok && aux_tma) {
    if (!d->out_f32 && d->residual == nullptr && !d->atomic && bn % 64 == 0 && d->N % 64 == 0 && (d->act == 0 || d->act == 1 || d->act == 2)) p.epi_mode = 1;
    else if (d->aux == nullptr && d->out_f32 && d->act == 0 && d->gn_stats == nullptr && d->res_mod == 0 && d->N % 32 == 0 &&
             (d->residual == nullptr || ((d->ldr * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15u) == 0 && !d->atomic &&
                                         nb1 * nb2 == 1 && !conv)))
      p.epi_mode = 2;
  }
  if (d->ln_x16 != nullptr || d->ln_colsum != nullptr) {
    COUNTR_REQUIRE(d->ln_stats != nullptr && (reinterpret_cast<uintptr_t>(d->ln_stats) & 15u) == 0, "LayerNorm folding needs 16-byte aligned ln_stats");
    COUNTR_REQUIRE(!conv && !conv_dw && nb1 * nb2 == 1 && split_k == 1, "LayerNorm folding: plain GEMM only");
    if (d->ln_x16 != nullptr) {
      COUNTR_REQUIRE(p.epi_mode == 2 && p.n_tiles * 2 <= 8 && bn % 64 == 0 && d->N % 64 == 0 && d->ld_x16 % 8 == 0 &&
                         (reinterpret_cast<uintptr_t>(d->ln_x16) & 15u) == 0 && d->ln_colsum == nullptr,
                     "LayerNorm producer needs the fp32 TMA epilogue, at most four N tiles of an even number of 32-column chunks per warp");
    } else {
      COUNTR_REQUIRE(p.epi_mode == 1 && d->ln_dim == d->K && d->alpha == 1.0f && d->aux == nullptr,
                     "LayerNorm consumer needs the 16-bit TMA epilogue, ln_dim == K, alpha == 1 and no aux");
    }
  }
  // ---- shared-memory layout
  p.stage_bytes = p.pair ? static_cast<int>(kPairStageBytes) : static_cast<int>(kABytes) + bn * 128;
  p.nbuf = 2;
  int room = static_cast<int>(kBarOff) - kEpiWarps * p.nbuf * static_cast<int>(kEpiBufBytes);
  const bool aux_needs_two = p.epi_mode == 1 && d->aux != nullptr;   // the aux epilogues address staging buffers 0 and 1
  if (room / p.stage_bytes < 4 && !aux_needs_two) {          // keep at least four operand stages: one staging buffer per warp instead
    p.nbuf = 1;
    room = static_cast<int>(kBarOff) - kEpiWarps * static_cast<int>(kEpiBufBytes);
  }
  // fp32 + residual tiles with three 32-column chunks per epilogue warp (bn = 192: proj / fc2 of the encoder): a third staging
  // buffer lets the WHOLE residual tile land during the main loop, so the exposed epilogue of these one-tile-per-CTA GEMMs is
  // TMEM read + add + store per chunk instead of a residual round trip per chunk — as long as three operand stages remain
  {
    static int nbuf3 = -1;
    if (nbuf3 < 0) {
      const char* e = getenv("COUNTR_EPI_NBUF3");
      nbuf3 = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int room3 = static_cast<int>(kBarOff) - kEpiWarps * 3 * static_cast<int>(kEpiBufBytes);
    if (nbuf3 && p.epi_mode == 2 && d->residual != nullptr && bn == 192 && p.nbuf == 2 && room3 / p.stage_bytes >= 3) {
      p.nbuf = 3;
      room = room3;
    }
  }
  p.nstages = room / p.stage_bytes;
  if (p.nstages > kMaxStages) p.nstages = kMaxStages;
  p.epi_off = p.nstages * p.stage_bytes;    // stage_bytes is a multiple of 4096: the staging area stays 1024-byte aligned
  p.box_x = conv ? (p.bx < 32 ? p.bx : 32) : 32;

  CUtensorMap tc = ta, tr = ta;             // placeholders when unused (never dereferenced)
  if (p.epi_mode != 0) {
    const uint32_t bcols = d->out_f32 ? 32u : 64u;
    if (conv) {
      const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->conv_w, (uint64_t)d->conv_h, (uint64_t)nb1};
      const uint64_t str[4] = {1, (uint64_t)d->ldc, (uint64_t)d->ldc * d->conv_w, (uint64_t)d->sc1};
      const uint32_t box[4] = {bcols, (uint32_t)p.box_x, (uint32_t)(32 / p.box_x), 1};
      rc = make_tmap_4d(&tc, d->c, c_elt, dims, str, box, TMAP_SW_128);
    } else {
      const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->M, (uint64_t)nb2, (uint64_t)nb1};
      const uint64_t str[4] = {1, (uint64_t)d->ldc, (uint64_t)(nb2 > 1 ? d->sc2 : d->ldc), (uint64_t)(nb1 > 1 ? d->sc1 : d->ldc)};
      const uint32_t box[4] = {bcols, 32, 1, 1};
      rc = make_tmap_4d(&tc, d->c, c_elt, dims, str, box, TMAP_SW_128);
    }
    if (rc) return rc;
    if (p.epi_mode == 1 && d->aux != nullptr) {
      const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->M, 1, 1};
      const uint64_t str[4] = {1, (uint64_t)d->ldaux, (uint64_t)d->ldaux, (uint64_t)d->ldaux};
      const uint32_t box[4] = {64, 32, 1, 1};
      rc = make_tmap_4d(&tr, d->aux, 2, dims, str, box, TMAP_SW_128);
      if (rc) return rc;
    }
    if (p.epi_mode == 2 && d->residual != nullptr) {
      const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->M, 1, 1};
      const uint64_t str[4] = {1, (uint64_t)d->ldr, (uint64_t)d->ldr, (uint64_t)d->ldr};
      const uint32_t box[4] = {32, 32, 1, 1};
      rc = make_tmap_4d(&tr, d->residual, 4, dims, str, box, TMAP_SW_128);
      if (rc) return rc;
    }
  }

  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  const int max_clusters = sms / p.cs;
  const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * p.cs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = p.cs;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.cs > 1 ? 2 : 1;   // plain (non-cluster) launch path for single-CTA tiles
  if (p.pair) COUNTR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<true>, ta, tb, tc, tr, p));
  else COUNTR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<false>, ta, tb, tc, tr, p));
  return COUNTR_OK;
}
