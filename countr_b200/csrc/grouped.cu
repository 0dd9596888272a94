// countr_b200 — grouped weight-gradient GEMM: every dW = dY^T X of a backward pass in ONE persistent launch.
//
// The decoder backward of the fine-tune step (and the whole backward of the MAE pre-training step) produces one weight
// gradient per Linear layer: dW[N_out, K_in] = dY[tokens, N_out]^T X[tokens, K_in] — a dozen mid-size GEMMs whose
// reduction dimension is the token count.  Launched one by one between the dX GEMMs of the critical path each of them
// pays a launch + prologue + pipeline fill + exposed epilogue (3-5 us against 5-10 us of tensor work) and fills the
// machine badly (8-32 output tiles).  They are leaves of the backward graph, so the host defers them: the dX chain runs
// first, then this kernel walks the concatenated tile list of all problems — split over the token dimension so that
// the list is several waves long — with one prologue and one tail.
//
//   warp 0      TMA producer: {dY 64 tokens x 128 outputs, X 64 tokens x 256 inputs} per stage, both MN-major boxes
//   warp 1      tcgen05.mma issue (kind::f16, M = 128, N = 256), fp32 accumulators double-buffered in TMEM
//   warps 2..9  epilogue: TMEM -> registers -> SWIZZLE_128B staging -> TMA reduce-add (fp32) into the gradient arena
//
// replaces: the `mm`(dY^T, X) half of every addmm backward autograd runs for nn.Linear (models_crossvit.py:55-57,77,
// 80,104-108; models_mae_cross.py:39; models_mae_noct.py under autograd) — see include/countr_b200.h.
#include "../../include/countr_b200.h"
#include "common.cuh"
#include "tma.h"

#include <algorithm>

namespace countr {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kMaxProb = COUNTR_MAX_GROUP;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK * 2;            // 16 KB
constexpr uint32_t kBBytes = BN * BK * 2;            // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;  // 48 KB
constexpr int kStages = 4;
constexpr uint32_t kEpiBufBytes = 32 * 128;
constexpr uint32_t kEpiOff = kStages * kStageBytes;                                   // 192 KB
constexpr uint32_t kBarOff = kEpiOff + kEpiWarps * kEpiBufBytes;                      // + 32 KB (one buffer per warp)
constexpr uint32_t kSmemBytes = kBarOff + 256 + 1024 /*alignment slack*/;

struct Prob {
  int m_tiles, n_tiles, mn_tiles;   // output tiles
  int split_k, k_per_split, K;      // token dimension: split_k pieces of k_per_split (multiple of 64) tokens
  int M, N;                         // n_out, k_in
  int tile_begin;                   // first tile of this problem in the concatenated list
};

struct GroupArgs {
  int nprob, total_tiles, bf16;
  Prob pr[kMaxProb];
};

struct GroupMaps {
  CUtensorMap a[kMaxProb], b[kMaxProb], c[kMaxProb];
};

struct Tile {
  int p, m, n, k_begin, nkb;
};

__device__ __forceinline__ Tile decode(const GroupArgs& g, int tile) {
  int p = 0;
  while (p + 1 < g.nprob && tile >= g.pr[p + 1].tile_begin) ++p;
  const Prob& q = g.pr[p];
  int idx = tile - q.tile_begin;
  const int s = idx / q.mn_tiles;
  idx -= s * q.mn_tiles;
  Tile t;
  t.p = p;
  t.n = idx / q.m_tiles;
  t.m = idx - t.n * q.m_tiles;
  t.k_begin = s * q.k_per_split;
  t.nkb = (min(q.K, t.k_begin + q.k_per_split) - t.k_begin + BK - 1) / BK;
  return t;
}

__global__ void __launch_bounds__(kThreads, 1)
grouped_dw_kernel(const __grid_constant__ GroupMaps maps, const __grid_constant__ GroupArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* full = bars;                // [kStages]
  uint64_t* empty = bars + 4;           // [kStages]
  uint64_t* tmem_full = bars + 8;       // [2]
  uint64_t* tmem_empty = bars + 10;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const Tile t = decode(g, tile);
      const CUtensorMap* ma = &maps.a[t.p];
      const CUtensorMap* mb = &maps.b[t.p];
      if (first) {
        if (lane == 0) {
          tma_prefetch_desc(ma);
          tma_prefetch_desc(mb);
        }
        pdl_wait();          // the operands come from earlier kernels
        first = false;
      }
      const int m0 = t.m * BM, n0 = t.n * BN;
      int k = t.k_begin;
      for (int kb = 0; kb < t.nkb; ++kb, k += BK) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tma_load_4d(sa, ma, &full[stage], m0, k, 0, 0);
          tma_load_4d(sa + 8192, ma, &full[stage], m0 + 64, k, 0, 0);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) tma_load_4d(sa + kABytes + i * 8192, mb, &full[stage], n0 + i * 64, k, 0, 0);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc = make_idesc_f16(BM, BN, true, true, g.bf16 != 0);
    const uint32_t step = 2048u >> 4;     // MN-major: +2048 B per 16-row k-step
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, ready = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const Tile t = decode(g, tile);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < t.nkb; ++kb) {
        if (!ready) mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint64_t a_desc = make_smem_desc_sw128(sa, 8192, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(sa + kABytes, 8192, 1024);
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == kStages) {
          nstage = 0;
          nphase ^= 1;
        }
        ready = umma_kblock<1>(d_tmem, a_desc, b_desc, step, step, idesc, kb != 0 ? 1u : 0u, empty0 + stage * 8,
                               smem_u32(&tmem_full[acc]), kb == t.nkb - 1 ? 1u : 0u, full0 + nstage * 8, nphase);
        __syncwarp();
        stage = nstage;
        phase = nphase;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue: fp32 reduce-add through TMA -------------------------------
    const int quarter = warp & 3, egroup = (warp - 2) >> 2;
    const uint32_t buf = smem_u32(smem + kEpiOff) + (warp - 2) * kEpiBufBytes;
    const uint32_t sw = static_cast<uint32_t>(lane & 7), rowoff = static_cast<uint32_t>(lane) * 128u;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const Tile t = decode(g, tile);
      const CUtensorMap* mc = &maps.c[t.p];
      if (first) {
        if (lane == 0) tma_prefetch_desc(mc);
        pdl_wait();          // the gradient arena is zeroed / still read by earlier kernels
        first = false;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      const int row0 = t.m * BM + quarter * 32;
      const Prob& q = g.pr[t.p];
      // (warp-uniform) nothing to store when this warp's 32 rows or a chunk's 32 columns are past the matrix
      for (int c = egroup; c < BN / 32 && row0 < q.M; c += 2) {
        if (t.n * BN + c * 32 >= q.N) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c * 32, r);
        if (lane == 0) bulk_wait_read<0>();     // the previous reduce-add has read the staging buffer
        __syncwarp();
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + rowoff + ((static_cast<uint32_t>(q) ^ sw) << 4)),
                       "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_4d(mc, buf, t.n * BN + c * 32, row0, 0, 0);     // rows / columns past the matrix are clipped
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------
// grouped column sums (bias gradients of the deferred Linear layers): out_i[n] += sum_r x_i[r][n]
// ------------------------------------------------------------------------------------------
struct ColsumProb {
  const void* x;
  float* out;
  long long R, ld;
  int N, dtype, rows_per_block, col_blocks, block_begin, vec8;
};
struct ColsumArgs {
  int nprob;
  ColsumProb pr[kMaxProb];
};

__global__ void __launch_bounds__(256) grouped_colsum_kernel(const __grid_constant__ ColsumArgs g) {
  int pi = 0;
  while (pi + 1 < g.nprob && static_cast<int>(blockIdx.x) >= g.pr[pi + 1].block_begin) ++pi;
  const ColsumProb& q = g.pr[pi];
  const int local = blockIdx.x - q.block_begin;
  const int cb = local % q.col_blocks, rb = local / q.col_blocks;
  pdl_trigger();
  pdl_wait();
  const long long r0 = static_cast<long long>(rb) * q.rows_per_block;
  const long long r1 = min(q.R, r0 + q.rows_per_block);
  __shared__ float red[8][257];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 column groups of 8 x 8 row lanes: 16-byte loads
  const int c = (cb * 32 + tx) * 8;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (q.vec8) {
    if (c < q.N) {
#pragma unroll 4
      for (long long r = r0 + ty; r < r1; r += 8) {
        if (q.dtype == 0) {
          const float4 v0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q.x) + r * q.ld + c);
          const float4 v1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q.x) + r * q.ld + c + 4);
          a[0] += v0.x; a[1] += v0.y; a[2] += v0.z; a[3] += v0.w; a[4] += v1.x; a[5] += v1.y; a[6] += v1.z; a[7] += v1.w;
        } else {
          const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(q.x) + r * q.ld + c);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (q.dtype == 2) { a[2 * j] += __uint_as_float(w[j] << 16); a[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u); }
            else { const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[j])); a[2 * j] += v.x; a[2 * j + 1] += v.y; }
          }
        }
      }
    }
  } else {
    // narrow / unaligned matrices: scalar loads, same thread layout
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c + j < q.N)
        for (long long r = r0 + ty; r < r1; r += 8) {
          if (q.dtype == 0) a[j] += reinterpret_cast<const float*>(q.x)[r * q.ld + c + j];
          else {
            const uint16_t hv = reinterpret_cast<const uint16_t*>(q.x)[r * q.ld + c + j];
            a[j] += q.dtype == 2 ? __uint_as_float(static_cast<uint32_t>(hv) << 16) : __half2float(*reinterpret_cast<const __half*>(&hv));
          }
        }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = a[j];
  __syncthreads();
  {
    const int col = threadIdx.x;          // 256 columns of the block, one per thread
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][col];
    if (cb * 256 + col < q.N) atomicAdd(q.out + cb * 256 + col, t);
  }
}

}  // namespace
}  // namespace countr

extern "C" int countr_grouped_colsum(const countr_colsum_problem* probs, int n, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(probs != nullptr && n >= 1 && n <= kMaxProb, "need 1..%d problems (got %d)", kMaxProb, n);
  ColsumArgs g{};
  g.nprob = n;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    const countr_colsum_problem& d = probs[i];
    COUNTR_REQUIRE(d.x && d.out && d.rows > 0 && d.cols > 0 && d.ld >= d.cols && d.dtype >= 0 && d.dtype <= 2,
                   "problem %d: bad arguments", i);
    ColsumProb& q = g.pr[i];
    q.x = d.x; q.out = d.out; q.R = d.rows; q.ld = d.ld; q.N = d.cols; q.dtype = d.dtype;
    q.col_blocks = (d.cols + 255) / 256;
    q.vec8 = (d.cols % 8 == 0 && d.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d.x) & 15u) == 0) ? 1 : 0;
    q.rows_per_block = 128;           // 256 columns x 128 rows per block
    const long long row_blocks = (d.rows + q.rows_per_block - 1) / q.rows_per_block;
    q.block_begin = blocks;
    blocks += q.col_blocks * static_cast<int>(row_blocks);
  }
  COUNTR_CHECK_CUDA(launch_pdl(grouped_colsum_kernel, dim3(blocks), dim3(256), 0, stream, g));
  return COUNTR_OK;
}

extern "C" int countr_grouped_dw(const countr_dw_problem* probs, int n, int bf16, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(probs != nullptr && n >= 1 && n <= kMaxProb, "need 1..%d problems (got %d)", kMaxProb, n);
  const int sms = num_sms();
  COUNTR_REQUIRE(sms > 0, "no CUDA device");
  GroupMaps maps;
  GroupArgs g{};
  g.nprob = n;
  g.bf16 = bf16;
  for (int i = 0; i < n; ++i) {
    const countr_dw_problem& d = probs[i];
    COUNTR_REQUIRE(d.dy && d.x && d.dw && d.tokens > 0 && d.n_out > 0 && d.k_in > 0, "problem %d: bad arguments", i);
    COUNTR_REQUIRE(d.ld_dy % 8 == 0 && d.ld_x % 8 == 0 && d.ld_dw % 4 == 0 && d.n_out % 8 == 0 && d.k_in % 8 == 0,
                   "problem %d: leading dimensions / widths must keep 16-byte alignment", i);
    COUNTR_REQUIRE(((reinterpret_cast<uintptr_t>(d.dy) | reinterpret_cast<uintptr_t>(d.x) | reinterpret_cast<uintptr_t>(d.dw)) & 15u) == 0,
                   "problem %d: operands must be 16-byte aligned", i);
    {
      const uint64_t dims[4] = {(uint64_t)d.n_out, (uint64_t)d.tokens, 1, 1};
      const uint64_t str[4] = {1, (uint64_t)d.ld_dy, (uint64_t)d.ld_dy, (uint64_t)d.ld_dy};
      const uint32_t box[4] = {64, BK, 1, 1};
      int rc = make_tmap_4d_16b(&maps.a[i], d.dy, dims, str, box, TMAP_SW_128);
      if (rc) return rc;
    }
    {
      const uint64_t dims[4] = {(uint64_t)d.k_in, (uint64_t)d.tokens, 1, 1};
      const uint64_t str[4] = {1, (uint64_t)d.ld_x, (uint64_t)d.ld_x, (uint64_t)d.ld_x};
      const uint32_t box[4] = {64, BK, 1, 1};
      int rc = make_tmap_4d_16b(&maps.b[i], d.x, dims, str, box, TMAP_SW_128);
      if (rc) return rc;
    }
    {
      const uint64_t dims[4] = {(uint64_t)d.k_in, (uint64_t)d.n_out, 1, 1};
      const uint64_t str[4] = {1, (uint64_t)d.ld_dw, (uint64_t)d.ld_dw, (uint64_t)d.ld_dw};
      const uint32_t box[4] = {32, 32, 1, 1};
      int rc = make_tmap_4d(&maps.c[i], d.dw, 4, dims, str, box, TMAP_SW_128);
      if (rc) return rc;
    }
    Prob& q = g.pr[i];
    q.m_tiles = (d.n_out + BM - 1) / BM;
    q.n_tiles = (d.k_in + BN - 1) / BN;
    q.mn_tiles = q.m_tiles * q.n_tiles;
    q.K = d.tokens;
    q.M = d.n_out;
    q.N = d.k_in;
  }
  // Split the token dimension so that the concatenated tile list is a whole number of waves of near-equal tiles: try every
  // tile length (in 64-token k-blocks) and keep the cheapest  waves x (k-blocks x 540 clk + 2500 clk per-tile epilogue / hand-off).
  int best_c = 1;
  double best_cost = 1e300;
  for (int c = 4; c <= 96; ++c) {
    long long tiles = 0;
    int longest = 0;
    for (int i = 0; i < n; ++i) {
      const int kb = (g.pr[i].K + BK - 1) / BK;
      const int split = (kb + c - 1) / c;
      const int per = (kb + split - 1) / split;
      tiles += static_cast<long long>(split) * g.pr[i].mn_tiles;
      longest = std::max(longest, per);
    }
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = static_cast<double>(waves) * (longest * 540.0 + 2500.0);
    if (cost < best_cost) {
      best_cost = cost;
      best_c = c;
    }
  }
  int total = 0;
  for (int i = 0; i < n; ++i) {
    Prob& q = g.pr[i];
    const int kb = (q.K + BK - 1) / BK;
    const int split = (kb + best_c - 1) / best_c;
    q.k_per_split = ((kb + split - 1) / split) * BK;
    q.split_k = (q.K + q.k_per_split - 1) / q.k_per_split;
    q.tile_begin = total;
    total += q.split_k * q.mn_tiles;
  }
  g.total_tiles = total;
  static PerDeviceOnce attr_once;
  if (attr_once.need())
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(grouped_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int grid = std::min(total, sms);
  COUNTR_CHECK_CUDA(launch_pdl(grouped_dw_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, maps, g));
  return COUNTR_OK;
}
