// countr_b200 — fused multi-head self-attention forward (flash-style, tcgen05 + TMEM).
//
//   out[b, q, h*dh:(h+1)*dh] = softmax_k( scale * Q[b,h,q,:] . K[b,h,k,:] ) @ V[b,h,:,:]
//
// replaces: Attention.forward lines models_crossvit.py:87-91 (q@k^T * scale, softmax, attn@v,
// transpose/reshape) for both the ViT encoder blocks (dh=64, H=12) and the FIM self-attention
// (dh=32, H=16).  The [B,H,L,L] score tensor the reference materialises never leaves the SM.
//
// Persistent, warp-specialised kernel: one CTA per SM walks work items = (batch, head, PAIR of 128-query tiles).
// K/V are streamed in 128-key chunks by TMA straight out of the packed qkv GEMM output [B][L][3][H][dh] (no
// permute / copy) through a 3-stage ring that both tiles share and that runs on into the next work item:
//   S_t = Q_t K_c^T      tcgen05.mma  (A,B K-major in smem)      -> TMEM columns [128 + 128 t, +128)
//   P_t = exp2(S*c - m)  thread == row, S row held in registers, online max/sum in fp32 (lazy rescale)
//                                                                 -> smem (16-bit, SWIZZLE_128B)
//   O_t += P_t V_c       tcgen05.mma  (B = V chunk, MN-major)    -> TMEM columns [64 t, +dh)
// Measured (profiles/r1_attention_fwd.md): the kernel is bound by the softmax warps — MUFU.EX2 at 16/clk/SM and the
// TMEM read path at 64 B/clk/SM are each ~1024 clk per 128x128 tile — not by the tensor pipe (14 % busy).
#include "../../include/countr_b200.h"
#include "common.cuh"
#include "tma.h"

#include <stdlib.h>
#include <type_traits>

namespace countr {

// optional clock64 timeline of CTA 0 (softmax warp 0 of tile 0 and the MMA warp) for scripts/trace_attn.py; -DCOUNTR_TRACE only
#ifdef COUNTR_TRACE
__device__ long long* g_attn_trace = nullptr;
#define ATR_INIT long long* const atr_ = (g_attn_trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) ? g_attn_trace : nullptr
#define ATR(slot) do { if (atr_ != nullptr && (slot) < 4096) atr_[(slot)] = clock64(); } while (0)
// per-CTA begin / end stamps (globaltimer, ns) at [2048 + 2 * cta + which]
#define ATR_CTA(which) do { if (g_attn_trace != nullptr && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_attn_trace[2048 + 2 * blockIdx.x + (which)] = static_cast<long long>(t_); } } while (0)
#else
#define ATR_CTA(which) do {} while (0)
#define ATR_INIT do {} while (0)
#define ATR(slot) do {} while (0)
#endif

namespace {

constexpr int BQ = 128;   // query rows per tile (one tcgen05.mma M)
constexpr int KC = 128;   // keys per chunk
constexpr int kStages = 3;                 // K / V chunk ring
constexpr int kSoftmaxWarps = 8;           // 4 per query tile: thread == query row
constexpr int kMmaWarp = 8, kTmaWarp = 9;
constexpr int kThreads = 32 * (kSoftmaxWarps + 2);
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS = 128;           // S of tile t at columns [128 + 128 t, +128); O of tile t at [64 t, +dh)

template <int DH>
struct AttnSmem {
  static constexpr uint32_t kRowBytes = DH * 2;               // 128 (SW128) or 64 (SW64)
  static constexpr uint32_t kQBytes = BQ * kRowBytes;
  static constexpr uint32_t kKBytes = KC * kRowBytes;
  static constexpr uint32_t kPBytes = BQ * KC * 2;            // 2 column blocks of [128 x 128 B]
  static constexpr uint32_t kOffQ = 0;                        // [2 buffers][2 query tiles]
  static constexpr uint32_t kOffK = 4 * kQBytes;              // [kStages]
  static constexpr uint32_t kOffV = kOffK + kStages * kKBytes;
  static constexpr uint32_t kOffP = kOffV + kStages * kKBytes;   // [2]
  static constexpr uint32_t kOffBar = kOffP + 2 * kPBytes;
  static constexpr uint32_t kTotal = kOffBar + 256 + 1024;
};

// shared-memory matrix descriptor with explicit swizzle mode (2 = 128B, 4 = 64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo & 0x3FFFFu) >> 4) << 16;
  d |= static_cast<uint64_t>((sbo & 0x3FFFFu) >> 4) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  uint32_t r;
  if (bf16)
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

struct AttnArgs {
  uint16_t* out;   // [B][L][H*dh]
  float* lse;      // [B][H][L] natural-log sum-exp of the scaled scores (optional)
  int B, L, H;
  float scale_log2;  // scale * log2(e)
  int bf16;
  int q_pairs;     // pairs of 128-query tiles per (batch, head)
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_v2f(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 ld_shared_v2f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
// named barriers 1..8 pair the two softmax warps that share the rows of a split item (fwd4 kernel)
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (An FMA-pipe cubic exp2 for a share of the exponentials — the FlashAttention-4 trick — was measured: 24.6 -> 23.2 us with
// one pair in four, slower with more; the softmax warps are issue- and hand-off-bound rather than MUFU-bound, see
// profiles/README.md.  Not kept: it costs exactness against the backward's recomputed probabilities.)

// One CTA per SM = one (batch, head, PAIR of 128-query tiles), warp-specialised:
//   warps 0-3 / 4-7  softmax of tile 0 / tile 1, thread == query row (TMEM lane), no cross-thread reductions
//   warp 8           tcgen05.mma issue (one lane)
//   warp 9           TMA producer (one lane): Q tiles once, K / V chunks through a 3-stage ring shared by both tiles
// Per tile and chunk:  S = Q K_c^T (TMEM) -> softmax warps pull their row of S into registers and hand the S
// columns straight back (the MMA warp issues S of the NEXT chunk while the exponentials of this one are computed)
// -> P (16-bit, swizzled smem) -> O += P V_c (TMEM).  The two tiles ping-pong on the tensor core and share every
// K / V byte.  The running maximum is only moved when it would grow by more than 2^8 (lazy rescale), so O is
// almost never read back from TMEM.
template <int DH, bool kBf16>
__global__ void __launch_bounds__(kThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                     const AttnArgs p) {
  using SM = AttnSmem<DH>;
  constexpr uint32_t kLayout = DH == 64 ? 2u : 4u;            // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t kSboK = DH == 64 ? 1024u : 512u;         // 8 rows of the K-major tiles
  constexpr uint32_t kVStep = 16 * SM::kRowBytes;             // 16 key rows per k-step of P.V

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* q_full = bars;                 // [2]  Q tiles of work item i are in buffer i & 1
  uint64_t* q_empty = q_full + 2;          // [2]
  uint64_t* k_full = q_empty + 2;          // [kStages]
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;    // [2]  S of tile t is in TMEM
  uint64_t* s_free = s_full + 2;           // [2]  the softmax warps hold S in registers
  uint64_t* p_full = s_free + 2;           // [2]  P of tile t is in smem (and O rescaled)
  uint64_t* p_free = p_full + 2;           // [2]  P.V of tile t has completed
  uint64_t* o_full = p_free + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp-uniform for the compiler
  const int nchunks = (p.L + KC - 1) / KC;
  const int BH = p.B * p.H;
  const int nitems = BH * p.q_pairs;
  // Work item -> (batch*head, pair of query tiles).  Items are ordered full pairs first, so the ragged last pair of every
  // head (one tile of 64 rows at L = 576) lands at the end of each CTA's list.
  auto item_bh = [&](int item) { return item % BH; };
  auto item_qp = [&](int item) { return item / BH; };
  auto item_ntiles = [&](int item) { return ((2 * item_qp(item) + 1) * BQ < p.L) ? 2 : 1; };

  pdl_trigger();
  ATR_INIT;
  if (tid == 0) ATR(0);
  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full + s, 1);
      mbar_init(q_empty + s, 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full + t, 1);
      mbar_init(s_free + t, 4);      // one arrival per softmax warp of the tile
      mbar_init(p_full + t, 4);
      mbar_init(p_free + t, 1);
      mbar_init(o_full + t, 1);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // qkv is the previous kernel's output

  // Both single-thread roles (TMA producer, MMA issuer) run their loops with the WHOLE warp and predicate only the
  // issue itself with elect.sync: inside an `if (lane == 0)` region ptxas cannot prove descriptors / coordinates
  // warp-uniform and wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR loop (~100 clk per instruction).
  if (warp == kTmaWarp) {
    {
      auto load_q = [&](int item, int i) {      // i = index of the item in this CTA's list
        const int buf = i & 1;
        if (i >= 2) mbar_wait(q_empty + buf, ((i >> 1) - 1) & 1);
        const int nt = item_ntiles(item), bh = item_bh(item), qp = item_qp(item);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full + buf, nt * SM::kQBytes);
          for (int t = 0; t < nt; ++t)
            tma_load_4d(smem + SM::kOffQ + (buf * 2 + t) * SM::kQBytes, &tma_q, q_full + buf, 0, (2 * qp + t) * BQ, bh % p.H,
                        bh / p.H);
        }
        __syncwarp();
      };
      int j = 0, i = 0;                         // j: running chunk index over all items (K / V ring position)
      if (static_cast<int>(blockIdx.x) < nitems) load_q(blockIdx.x, 0);
      const int q_prefetch_chunk = min(2, nchunks - 1);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++i) {
        const int bh = item_bh(item), h = bh % p.H, b = bh / p.H;
        for (int c = 0; c < nchunks; ++c, ++j) {
          const int s = j % kStages;
          const uint32_t ph = static_cast<uint32_t>(j / kStages) & 1u;
          if (j >= kStages) mbar_wait(k_empty + s, ph ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(k_full + s, SM::kKBytes);
            tma_load_4d(smem + SM::kOffK + s * SM::kKBytes, &tma_kv, k_full + s, 0, c * KC, p.H + h, b);
          }
          __syncwarp();
          if (j >= kStages) mbar_wait(v_empty + s, ph ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(v_full + s, SM::kKBytes);
            tma_load_4d(smem + SM::kOffV + s * SM::kKBytes, &tma_kv, v_full + s, 0, c * KC, 2 * p.H + h, b);
          }
          __syncwarp();
          // the next item's Q tiles, once the ring has moved past the previous item (whose Q buffer this reuses)
          if (c == q_prefetch_chunk && item + static_cast<int>(gridDim.x) < nitems) load_q(item + gridDim.x, i + 1);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    {
      const uint32_t idesc_s = make_idesc_f16(BQ, KC, false, false, kBf16);
      const uint32_t idesc_o = make_idesc_f16(BQ, DH, false, true, kBf16);
      const uint32_t sQ = smem_u32(smem + SM::kOffQ), sK = smem_u32(smem + SM::kOffK);
      const uint32_t sV = smem_u32(smem + SM::kOffV), sP = smem_u32(smem + SM::kOffP);
      int n_s[2] = {0, 0};                      // S tiles issued so far per query tile (phase bookkeeping)
      int n_p[2] = {0, 0};                      // P.V products issued so far per query tile
      // S_t = Q_t K^T for every tile of item (index i in this CTA's list), chunk c; js = ring position of the chunk
      auto issue_s = [&](int item, int i, int c, int js) {
        const int buf = i & 1, nt = item_ntiles(item), s = js % kStages;
        if (c == 0) mbar_wait(q_full + buf, (i >> 1) & 1);
        mbar_wait(k_full + s, static_cast<uint32_t>(js / kStages) & 1u);
        for (int t = 0; t < nt; ++t) {
          if (n_s[t] > 0) mbar_wait(s_free + t, (n_s[t] - 1) & 1);   // the softmax warps hold the previous S_t in registers
          tc_fence_after();
          const uint64_t a_desc = make_desc(sQ + (buf * 2 + t) * SM::kQBytes, 16, kSboK, kLayout);
          const uint64_t b_desc = make_desc(sK + s * SM::kKBytes, 16, kSboK, kLayout);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_f16_ss(tmem_base + kTmemS + t * KC, a_desc + static_cast<uint64_t>(2 * k),
                          b_desc + static_cast<uint64_t>(2 * k), idesc_s, k != 0);
            umma_commit(s_full + t);
            if (t == nt - 1) {
              umma_commit(k_empty + s);
              if (c == nchunks - 1) umma_commit(q_empty + buf);
            }
          }
          __syncwarp();
          ++n_s[t];
        }
      };
      int j = 0, i = 0;
      if (static_cast<int>(blockIdx.x) < nitems) issue_s(blockIdx.x, 0, 0, 0);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++i) {
        const int nt = item_ntiles(item);
        for (int c = 0; c < nchunks; ++c, ++j) {
          // run one chunk ahead with S — into the next work item at the end of this one
          ATR(1024 + 8 * j);
          if (c + 1 < nchunks) issue_s(item, i, c + 1, j + 1);
          else if (item + static_cast<int>(gridDim.x) < nitems) issue_s(item + gridDim.x, i + 1, 0, j + 1);
          ATR(1025 + 8 * j);
          const int s = j % kStages;
          const int valid = min(KC, p.L - c * KC);
          const int ksteps = ((valid + 31) / 32) * 2;   // 16 keys per k-step, whole 32-key groups
          mbar_wait(v_full + s, static_cast<uint32_t>(j / kStages) & 1u);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(p_full + t, n_p[t] & 1);
            ATR(1026 + 8 * j + 2 * t);
            tc_fence_after();
            if (elect_one()) {
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t a_desc = make_desc(sP + t * SM::kPBytes + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024, 2u);
                const uint64_t b_desc = make_desc(sV + s * SM::kKBytes + k * kVStep, 16, kSboK, kLayout);
                umma_f16_ss(tmem_base + t * 64, a_desc, b_desc, idesc_o, (c | k) != 0);
              }
              umma_commit(p_free + t);
              if (c == nchunks - 1) umma_commit(o_full + t);
              if (t == nt - 1) umma_commit(v_empty + s);
            }
            __syncwarp();
            ATR(1027 + 8 * j + 2 * t);
            ++n_p[t];
          }
        }
      }
    }
  } else {
    const int t = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + kTmemS + t * KC, t_o = t_lane + t * 64;
    // 16-byte chunk j of this row's 128-byte line sits at chunk (j ^ (row & 7)) (SWIZZLE_128B): one base + an XOR immediate
    const uint32_t prow = smem_u32(smem + SM::kOffP + t * SM::kPBytes + (row >> 3) * 1024 + (row & 7) * 128) | ((row & 7) << 4);
    int n_c = 0;    // chunks of tile t seen so far (phase of s_full / s_free / p_full / p_free)
    int n_o = 0;    // items of tile t finished so far (phase of o_full)
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      if (t >= item_ntiles(item)) continue;
      const int bh = item_bh(item), h = bh % p.H, b = bh / p.H;
      const int q0 = (2 * item_qp(item) + t) * BQ;
      if (q0 + quarter * 32 >= p.L) {
        // all 32 rows of this warp are past the end of the sequence: keep the barrier protocol in step, do no work
        for (int c = 0; c < nchunks; ++c, ++n_c) {
          mbar_wait(s_full + t, n_c & 1);
          if (lane == 0) mbar_arrive(s_free + t);
          if (lane == 0) mbar_arrive(p_full + t);
          mbar_wait(p_free + t, n_c & 1);
        }
        ++n_o;
        continue;
      }
      float m_run = -INFINITY;  // running max (log2 domain, already scaled)
      float l_run = 0.f;
      // One 128-key chunk of this thread's row per iteration (whole 32-key groups are read; keys past the end of the
      // sequence are masked to -inf).
      for (int c = 0; c < nchunks; ++c) {
        constexpr bool kFull = false;
        const int valid = min(KC, p.L - c * KC);          // keys in this chunk
        const int groups = (valid + 31) / 32;             // 32-column groups that hold any valid key
        if (warp == 0) ATR(16 + 8 * n_c);
        mbar_wait(s_full + t, n_c & 1);
        if (warp == 0) ATR(17 + 8 * n_c);
        tc_fence_after();
        // ---- the whole row of S into registers; the S columns go straight back to the MMA warp ----
        uint32_t r[4][32];
        tmem_ld_32x32b_x32(t_s, r[0]);
        if (kFull || groups > 1) tmem_ld_32x32b_x32(t_s + 32, r[1]);
        if (kFull || groups > 2) tmem_ld_32x32b_x32(t_s + 64, r[2]);
        if (kFull || groups > 3) tmem_ld_32x32b_x32(t_s + 96, r[3]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free + t);
        if (warp == 0) ATR(18 + 8 * n_c);
        if (!kFull) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int lim = valid - g * 32;
            if (lim > 0 && lim < 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j >= lim) r[g][j] = 0xff800000u;  // -inf
            }
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (kFull || g < groups) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              mx4[0] = fmaxf(mx4[0], fmaxf(__uint_as_float(r[g][j + 0]), __uint_as_float(r[g][j + 1])));
              mx4[1] = fmaxf(mx4[1], fmaxf(__uint_as_float(r[g][j + 2]), __uint_as_float(r[g][j + 3])));
              mx4[2] = fmaxf(mx4[2], fmaxf(__uint_as_float(r[g][j + 4]), __uint_as_float(r[g][j + 5])));
              mx4[3] = fmaxf(mx4[3], fmaxf(__uint_as_float(r[g][j + 6]), __uint_as_float(r[g][j + 7])));
            }
          }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // Lazy rescale: the running maximum only moves when it would grow by more than 2^8 — until then
        // P = exp2(s - m_run) stays below 256 (fine in the 16-bit P tile and in the fp32 sums) and O needs no correction.
        const float m_cand = fmaxf(m_run, mx * p.scale_log2);
        float corr = 1.f;
        if (m_cand - m_run > 8.f) {    // also taken on the first chunk (m_run = -inf): corr = 0
          corr = ex2_approx(m_run - m_cand);
          m_run = m_cand;
        }
        bool rescaled = false;
        if (warp == 0) ATR(19 + 8 * n_c);
        if (n_c > 0) {
          mbar_wait(p_free + t, (n_c - 1) & 1);   // the previous P.V of this tile is done: the P tile and O are ours again
          tc_fence_after();
          if (c > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
            rescaled = true;
#pragma unroll
            for (int d0 = 0; d0 < DH; d0 += 16) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + d0, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
              tmem_st_32x32b_x16(t_o + d0, o);
            }
          }
        }
        if (warp == 0) ATR(20 + 8 * n_c);
        // ---- P = exp2(S*c - m) (packed FFMA2 / FADD2), row sum, 16-bit P row into swizzled smem ----
        const float2 sc2 = splat2(p.scale_log2), nm2 = splat2(-m_run);
        float2 ls2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (kFull || g < groups) {
            // group g covers key columns [32g, 32g+32) = 16-byte chunks (g&1)*4 .. +3 of column block g>>1
            const uint32_t pblk = prow + (g >> 1) * (BQ * 128);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t w[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 x = fma2(make_float2(__uint_as_float(r[g][8 * q + 2 * j]), __uint_as_float(r[g][8 * q + 2 * j + 1])), sc2, nm2);
                const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                ls2[j & 1] = add2(ls2[j & 1], e);
                w[j] = pack2(e.x, e.y, kBf16);
              }
              st_shared_v4(pblk ^ (((g & 1) * 4 + q) << 4), make_uint4(w[0], w[1], w[2], w[3]));
            }
          }
        l_run = l_run * corr + ((ls2[0].x + ls2[0].y) + (ls2[1].x + ls2[1].y));
        if (rescaled) tmem_st_wait();
        fence_proxy_async_smem();  // P row (generic-proxy stores) -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + t);
        if (warp == 0) ATR(21 + 8 * n_c);
        ++n_c;
      }

      // ---- epilogue: O / l -> 16 bit, one dh-wide row segment per thread ----
      mbar_wait(o_full + t, n_o & 1);
      ++n_o;
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const int q = q0 + row;
      uint16_t* orow = p.out + (static_cast<long long>(b) * p.L + q) * (p.H * DH) + h * DH;
#pragma unroll
      for (int d0 = 0; d0 < DH; d0 += 16) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(t_o + d0, o);
        tmem_ld_wait();
        if (q < p.L) {
          uint4 o0, o1;
          o0.x = pack2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l, kBf16);
          o0.y = pack2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l, kBf16);
          o0.z = pack2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l, kBf16);
          o0.w = pack2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l, kBf16);
          o1.x = pack2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l, kBf16);
          o1.y = pack2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l, kBf16);
          o1.z = pack2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l, kBf16);
          o1.w = pack2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l, kBf16);
          *reinterpret_cast<uint4*>(orow + d0) = o0;
          *reinterpret_cast<uint4*>(orow + d0 + 8) = o1;
        }
      }
      if (p.lse != nullptr && q < p.L)
        p.lse[(static_cast<long long>(b) * p.H + h) * p.L + q] = (m_run + log2f(l_run)) * 0.69314718055994531f;
      // O_t must not be overwritten by the next item's first P.V before every warp of the tile has read it: that P.V
      // waits for p_full, which this warp only arrives on after these loads (program order + the fence below)
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ================================================================================================
// Four-tile variant of the forward kernel.  The two-tile kernel above is bound by the MUFU (exp2) and keeps it busy only two
// thirds of the time: the two query tiles of a CTA move through S load -> row max -> exp -> P hand-off in lock-step, and one
// warp per SM sub-partition cannot saturate the unit on its own (profiles/r2_attention_fwd.md).  Here a CTA works on FOUR
// 128-query tiles of one (batch, head) at once with 64-key chunks: S_t is 64 TMEM columns (4 x 64 + 4 x 64 for O = all 512),
// sixteen softmax warps (four per sub-partition, 64 S values per thread and chunk) drift apart naturally, so some warp is
// always in its exp phase.  One work item = one head's four tiles (576 tokens: 4 tiles + one 64-row remainder item).
// ================================================================================================
constexpr int NT4 = 4;                     // query tiles in flight per CTA
constexpr int KC4 = 64;                    // keys per chunk
constexpr int KS4 = 4;                     // K / V ring stages
// TWO MMA-issuing warps: a tcgen05.mma costs ~60-85 clk to issue (csrc/gemm.cu), and with 16-key k-steps and 64-wide tiles this
// kernel issues 8 of them per (tile, chunk) — one warp issuing S = Q K^T and P.V in turn was the bottleneck (3300 of the
// 5100 clk per chunk step), and its in-order barrier waits kept the tiles in lock-step.
constexpr int kSoftmaxWarps4 = 16, kMmaWarp4 = 16, kPvWarp4 = 17, kTmaWarp4 = 18;
constexpr int kThreads4 = 32 * 19;

template <int DH>
struct Attn4Smem {
  static constexpr uint32_t kRowBytes = DH * 2;
  static constexpr uint32_t kQBytes = BQ * kRowBytes;
  static constexpr uint32_t kKBytes = KC4 * kRowBytes;
  static constexpr uint32_t kPBytes = BQ * KC4 * 2;            // [128 x 128 B], one SWIZZLE_128B column block
  static constexpr uint32_t kOffQ = 0;                         // [NT4]
  static constexpr uint32_t kOffK = NT4 * kQBytes;             // [KS4]
  static constexpr uint32_t kOffV = kOffK + KS4 * kKBytes;
  static constexpr uint32_t kOffP = kOffV + KS4 * kKBytes;     // [NT4]
  static constexpr uint32_t kOffBar = kOffP + NT4 * kPBytes;
  static constexpr uint32_t kTotal = kOffBar + 512 + 1024;
};

struct Attn4Args {
  uint16_t* out;
  float* lse;
  int B, L, H;
  float scale_log2;
  int n_full, n_items, full_groups;     // items [0, n_full): four-tile groups (bh-major inside a group index), then the remainders
  int split;                            // split one-tile items over two slots by key chunk
};

template <int DH, bool kBf16>
__global__ void __launch_bounds__(kThreads4, 1)
attention_fwd4_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, const Attn4Args p) {
  using SM = Attn4Smem<DH>;
  constexpr uint32_t kLayout = DH == 64 ? 2u : 4u;            // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t kSboK = DH == 64 ? 1024u : 512u;
  constexpr uint32_t kVStep = 16 * SM::kRowBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* q_empty = bars + 1;            // [1]
  uint64_t* k_full = bars + 2;             // [KS4]
  uint64_t* k_empty = k_full + KS4;
  uint64_t* v_full = k_empty + KS4;
  uint64_t* v_empty = v_full + KS4;
  uint64_t* s_full = v_empty + KS4;        // [NT4]
  uint64_t* s_free = s_full + NT4;
  uint64_t* p_full = s_free + NT4;
  uint64_t* p_free = p_full + NT4;
  uint64_t* o_full = p_free + NT4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + NT4);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nchunks = (p.L + KC4 - 1) / KC4;
  const int BH = p.B * p.H;
  const int tiles_total = (p.L + BQ - 1) / BQ;
  // it-th work item of this CTA (-1: done).  With fewer four-tile items than CTAs every such item gets its own CTA and the
  // remainder items (a fraction of a tile each) are spread over the CTAs that are left.
  auto cta_item = [&](int it) -> int {
    const int grid = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
    if (p.n_full < grid && p.n_items > p.n_full) {
      if (cta < p.n_full) return it == 0 ? cta : -1;
      const int r = (cta - p.n_full) + it * (grid - p.n_full);
      return p.n_full + r < p.n_items ? p.n_full + r : -1;
    }
    const int item = cta + it * grid;
    return item < p.n_items ? item : -1;
  };
  auto item_bh = [&](int item) { return item < p.n_full ? item % BH : item - p.n_full; };
  auto item_group = [&](int item) { return item < p.n_full ? item / BH : p.full_groups; };
  auto item_ntiles = [&](int item) { return min(NT4, tiles_total - NT4 * item_group(item)); };
  // a one-tile item would leave three quarters of the CTA idle and run its chunks strictly one after the other (S -> softmax ->
  // P.V latency per chunk): it is split over two tile slots by key chunk (even / odd) and merged at the end
  auto item_split = [&](int nt) { return p.split != 0 && nt == 1 && nchunks >= 2; };

  pdl_trigger();
  ATR_INIT;
  ATR_CTA(0);
  if (tid == 0) ATR(0);
  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KS4; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
    }
    for (int t = 0; t < NT4; ++t) {
      mbar_init(s_full + t, 1);
      mbar_init(s_free + t, 4);
      mbar_init(p_full + t, 4);
      mbar_init(p_free + t, 1);
      mbar_init(o_full + t, 1);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp4) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == kTmaWarp4) {
    int j = 0;
    for (int it = 0;; ++it) {
      const int item = cta_item(it);
      if (item < 0) break;
      const int bh = item_bh(item), h = bh % p.H, b = bh / p.H, g = item_group(item), nt = item_ntiles(item);
      if (it > 0) mbar_wait(q_empty, (it - 1) & 1);          // the previous item's last S = Q K^T has read the Q tiles
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, nt * SM::kQBytes);
        for (int t = 0; t < nt; ++t)
          tma_load_4d(smem + SM::kOffQ + t * SM::kQBytes, &tma_q, q_full, 0, (NT4 * g + t) * BQ, h, b);
      }
      __syncwarp();
      for (int c = 0; c < nchunks; ++c, ++j) {
        const int s = j % KS4;
        const uint32_t ph = static_cast<uint32_t>(j / KS4) & 1u;
        if (j >= KS4) mbar_wait(k_empty + s, ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full + s, SM::kKBytes);
          tma_load_4d(smem + SM::kOffK + s * SM::kKBytes, &tma_kv, k_full + s, 0, c * KC4, p.H + h, b);
        }
        __syncwarp();
        if (j >= KS4) mbar_wait(v_empty + s, ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(v_full + s, SM::kKBytes);
          tma_load_4d(smem + SM::kOffV + s * SM::kKBytes, &tma_kv, v_full + s, 0, c * KC4, 2 * p.H + h, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp4) {
    // ---- S = Q K^T issuer: runs as far ahead as the s_free hand-backs and the K ring allow
    const uint32_t idesc_s = make_idesc_f16(BQ, KC4, false, false, kBf16);
    const uint32_t sQ = smem_u32(smem + SM::kOffQ), sK = smem_u32(smem + SM::kOffK);
    int n_s0 = 0, n_s1 = 0, n_s2 = 0, n_s3 = 0;      // S tiles issued so far per query tile
    int j = 0;
    for (int it = 0;; ++it) {
      const int item = cta_item(it);
      if (item < 0) break;
      const int nt = item_ntiles(item);
      const bool split = item_split(nt);
      mbar_wait(q_full, it & 1);
      for (int c = 0; c < nchunks; ++c, ++j) {
        const int s = j % KS4;
        mbar_wait(k_full + s, static_cast<uint32_t>(j / KS4) & 1u);
#pragma unroll
        for (int t = 0; t < NT4; ++t) {
          if (split ? t == (c & 1) : t < nt) {
            int& n_s = t == 0 ? n_s0 : t == 1 ? n_s1 : t == 2 ? n_s2 : n_s3;
            if (n_s > 0) mbar_wait(s_free + t, (n_s - 1) & 1);
            tc_fence_after();
            const uint64_t a_desc = make_desc(sQ + (split ? 0 : t) * SM::kQBytes, 16, kSboK, kLayout);
            const uint64_t b_desc = make_desc(sK + s * SM::kKBytes, 16, kSboK, kLayout);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < DH / 16; ++k)
                umma_f16_ss(tmem_base + 256 + t * KC4, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc_s,
                            k != 0);
              umma_commit(s_full + t);
              if (split || t == nt - 1) {
                umma_commit(k_empty + s);
                if (c == nchunks - 1) umma_commit(q_empty);
              }
            }
            __syncwarp();
            ++n_s;
          }
        }
      }
    }
  } else if (warp == kPvWarp4) {
    // ---- O += P V issuer
    const uint32_t idesc_o = make_idesc_f16(BQ, DH, false, true, kBf16);
    const uint32_t sV = smem_u32(smem + SM::kOffV), sP = smem_u32(smem + SM::kOffP);
    int n_p0 = 0, n_p1 = 0, n_p2 = 0, n_p3 = 0;      // P.V products issued so far per query tile
    int j = 0;
    for (int it = 0;; ++it) {
      const int item = cta_item(it);
      if (item < 0) break;
      const int nt = item_ntiles(item);
      const bool split = item_split(nt);
      for (int c = 0; c < nchunks; ++c, ++j) {
        const int s = j % KS4;
        const int valid = min(KC4, p.L - c * KC4);
        const int ksteps = ((valid + 31) / 32) * 2;
        const int c_own = split ? c >> 1 : c;            // chunks this slot has accumulated before this one
        mbar_wait(v_full + s, static_cast<uint32_t>(j / KS4) & 1u);
#pragma unroll
        for (int t = 0; t < NT4; ++t) {
          if (split ? t == (c & 1) : t < nt) {
            int& n_p = t == 0 ? n_p0 : t == 1 ? n_p1 : t == 2 ? n_p2 : n_p3;
            mbar_wait(p_full + t, n_p & 1);
            tc_fence_after();
            if (elect_one()) {
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t a_desc = make_desc(sP + t * SM::kPBytes + k * 32, 16, 1024, 2u);
                const uint64_t b_desc = make_desc(sV + s * SM::kKBytes + k * kVStep, 16, kSboK, kLayout);
                umma_f16_ss(tmem_base + t * 64, a_desc, b_desc, idesc_o, (c_own | k) != 0);
              }
              umma_commit(p_free + t);
              if (split ? c >= nchunks - 2 : c == nchunks - 1) umma_commit(o_full + t);
              if (split || t == nt - 1) umma_commit(v_empty + s);
            }
            __syncwarp();
            ++n_p;
          }
        }
      }
    }
  } else {
    const int t = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + 256 + t * KC4, t_o = t_lane + t * 64;
    const uint32_t prow = smem_u32(smem + SM::kOffP + t * SM::kPBytes + (row >> 3) * 1024 + (row & 7) * 128) | ((row & 7) << 4);
    // hand-over buffers of a split item: O rows (fp32, 16-byte chunks XOR-swizzled by the row) in Q slots 2-3, (m, l) in P slot 3
    const uint32_t xrow = smem_u32(smem + SM::kOffQ + 2 * SM::kQBytes) + row * (DH * 4);
    const uint32_t xml = smem_u32(smem + SM::kOffP + 3 * SM::kPBytes) + row * 8;
    bool xbuf_used = false;
    int n_c = 0, n_o = 0;
    for (int it = 0;; ++it) {
      const int item = cta_item(it);
      if (item < 0) break;
      const int nt = item_ntiles(item);
      const bool split = item_split(nt);
      if (t >= (split ? 2 : nt)) continue;
      const int bh = item_bh(item), h = bh % p.H, b = bh / p.H;
      const int q0 = (NT4 * item_group(item) + (split ? 0 : t)) * BQ;
      const int c0 = split ? t : 0, cstep = split ? 2 : 1;      // a split item: this slot takes every other key chunk
      if (q0 + quarter * 32 >= p.L) {
        for (int c = c0; c < nchunks; c += cstep, ++n_c) {
          mbar_wait(s_full + t, n_c & 1);
          if (lane == 0) mbar_arrive(s_free + t);
          if (lane == 0) mbar_arrive(p_full + t);
          mbar_wait(p_free + t, n_c & 1);
        }
        ++n_o;
        continue;
      }
      float m_run = -INFINITY, l_run = 0.f;
      for (int c = c0; c < nchunks; c += cstep) {
        const int valid = min(KC4, p.L - c * KC4);
        const int groups = (valid + 31) / 32;
        if (warp == 0) ATR(16 + 8 * n_c);
        mbar_wait(s_full + t, n_c & 1);
        if (warp == 0) ATR(17 + 8 * n_c);
        tc_fence_after();
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(t_s, r[0]);
        if (groups > 1) tmem_ld_32x32b_x32(t_s + 32, r[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free + t);
        if (warp == 0) ATR(18 + 8 * n_c);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int lim = valid - g * 32;
          if (lim > 0 && lim < 32) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (jj >= lim) r[g][jj] = 0xff800000u;
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int g = 0; g < 2; ++g)
          if (g < groups) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 8) {
              mx4[0] = fmaxf(mx4[0], fmaxf(__uint_as_float(r[g][jj + 0]), __uint_as_float(r[g][jj + 1])));
              mx4[1] = fmaxf(mx4[1], fmaxf(__uint_as_float(r[g][jj + 2]), __uint_as_float(r[g][jj + 3])));
              mx4[2] = fmaxf(mx4[2], fmaxf(__uint_as_float(r[g][jj + 4]), __uint_as_float(r[g][jj + 5])));
              mx4[3] = fmaxf(mx4[3], fmaxf(__uint_as_float(r[g][jj + 6]), __uint_as_float(r[g][jj + 7])));
            }
          }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float m_cand = fmaxf(m_run, mx * p.scale_log2);
        float corr = 1.f;
        if (m_cand - m_run > 8.f) {
          corr = ex2_approx(m_run - m_cand);
          m_run = m_cand;
        }
        bool rescaled = false;
        if (warp == 0) ATR(19 + 8 * n_c);
        if (n_c > 0) {
          mbar_wait(p_free + t, (n_c - 1) & 1);
          tc_fence_after();
          if (c != c0 && __any_sync(0xffffffffu, corr != 1.f)) {
            rescaled = true;
#pragma unroll
            for (int d0 = 0; d0 < DH; d0 += 16) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + d0, o);
              tmem_ld_wait();
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * corr);
              tmem_st_32x32b_x16(t_o + d0, o);
            }
          }
        }
        if (warp == 0) ATR(20 + 8 * n_c);
        const float2 sc2 = splat2(p.scale_log2), nm2 = splat2(-m_run);
        float2 ls2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int g = 0; g < 2; ++g)
          if (g < groups) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t w[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float2 x = fma2(make_float2(__uint_as_float(r[g][8 * q + 2 * jj]), __uint_as_float(r[g][8 * q + 2 * jj + 1])), sc2, nm2);
                const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                ls2[jj & 1] = add2(ls2[jj & 1], e);
                w[jj] = pack2(e.x, e.y, kBf16);
              }
              st_shared_v4(prow ^ ((g * 4 + q) << 4), make_uint4(w[0], w[1], w[2], w[3]));
            }
          }
        l_run = l_run * corr + ((ls2[0].x + ls2[0].y) + (ls2[1].x + ls2[1].y));
        if (rescaled) tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + t);
        if (warp == 0) ATR(21 + 8 * n_c);
        ++n_c;
      }
      mbar_wait(o_full + t, n_o & 1);
      ++n_o;
      tc_fence_after();
      float f_own = 1.f, f_other = 0.f;
      if (split) {
        // the two slots hold partial softmax sums over the even / odd key chunks of the same rows: slot 1 hands its
        // (m, l, O) over through shared memory (Q slots 2-3 and P slot 3 are idle in a one-tile item), slot 0 merges
        if (t == 1) {
          if (xbuf_used) named_bar_sync(5 + quarter, 64);          // slot 0 has read the previous item's hand-over
          xbuf_used = true;
#pragma unroll
          for (int d0 = 0; d0 < DH; d0 += 16) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(t_o + d0, o);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < 4; ++v)
              st_shared_v4(xrow + ((((d0 >> 2) + v) ^ (row & (DH / 4 - 1))) << 4), make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]));
          }
          st_shared_v2f(xml, m_run, l_run);
          tc_fence_before();
          __threadfence_block();
          named_bar_arrive(1 + quarter, 64);
          continue;
        }
        named_bar_sync(1 + quarter, 64);
        const float2 ml1 = ld_shared_v2f(xml);
        const float m_all = fmaxf(m_run, ml1.x);
        f_own = ex2_approx(m_run - m_all);
        f_other = ex2_approx(ml1.x - m_all);
        l_run = l_run * f_own + ml1.y * f_other;
        m_run = m_all;
      }
      const float inv_l = 1.f / l_run;
      const int q = q0 + row;
      uint16_t* orow = p.out + (static_cast<long long>(b) * p.L + q) * (p.H * DH) + h * DH;
#pragma unroll
      for (int d0 = 0; d0 < DH; d0 += 16) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(t_o + d0, o);
        tmem_ld_wait();
        if (split) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 x = ld_shared_v4(xrow + ((((d0 >> 2) + v) ^ (row & (DH / 4 - 1))) << 4));
            o[4 * v + 0] = __float_as_uint(__uint_as_float(o[4 * v + 0]) * f_own + __uint_as_float(x.x) * f_other);
            o[4 * v + 1] = __float_as_uint(__uint_as_float(o[4 * v + 1]) * f_own + __uint_as_float(x.y) * f_other);
            o[4 * v + 2] = __float_as_uint(__uint_as_float(o[4 * v + 2]) * f_own + __uint_as_float(x.z) * f_other);
            o[4 * v + 3] = __float_as_uint(__uint_as_float(o[4 * v + 3]) * f_own + __uint_as_float(x.w) * f_other);
          }
        }
        if (q < p.L) {
          uint4 o0, o1;
          o0.x = pack2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l, kBf16);
          o0.y = pack2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l, kBf16);
          o0.z = pack2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l, kBf16);
          o0.w = pack2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l, kBf16);
          o1.x = pack2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l, kBf16);
          o1.y = pack2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l, kBf16);
          o1.z = pack2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l, kBf16);
          o1.w = pack2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l, kBf16);
          *reinterpret_cast<uint4*>(orow + d0) = o0;
          *reinterpret_cast<uint4*>(orow + d0 + 8) = o1;
        }
      }
      if (p.lse != nullptr && q < p.L)
        p.lse[(static_cast<long long>(b) * p.H + h) * p.L + q] = (m_run + log2f(l_run)) * 0.69314718055994531f;
      tc_fence_before();
      if (split && cta_item(it + 1) >= 0) {       // the next item of this CTA is a one-tile item too (they come last)
        __threadfence_block();
        named_bar_arrive(5 + quarter, 64);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  ATR_CTA(1);
  if (warp == kMmaWarp4) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int DH, bool kBf16>
int launch_attention4(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t stream) {
  using SM = Attn4Smem<DH>;
  CUtensorMap tq, tkv;
  const uint64_t dims[4] = {(uint64_t)DH, (uint64_t)L, (uint64_t)(3 * H), (uint64_t)B};
  const uint64_t str[4] = {1, (uint64_t)(3 * H * DH), (uint64_t)DH, (uint64_t)L * 3 * H * DH};
  const uint32_t boxq[4] = {DH, BQ, 1, 1}, boxk[4] = {DH, KC4, 1, 1};
  const TmapSwizzle sw = DH == 64 ? TMAP_SW_128 : TMAP_SW_64;
  int rc = make_tmap_4d_16b(&tq, qkv, dims, str, boxq, sw);
  if (rc) return rc;
  rc = make_tmap_4d_16b(&tkv, qkv, dims, str, boxk, sw);
  if (rc) return rc;
  static PerDeviceOnce attr_once;
  if (attr_once.need())
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd4_kernel<DH, kBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  Attn4Args p;
  p.out = reinterpret_cast<uint16_t*>(out);
  p.lse = lse;
  p.B = B; p.L = L; p.H = H;
  p.scale_log2 = scale * 1.44269504088896340736f;
  const int tiles = (L + BQ - 1) / BQ;
  p.full_groups = tiles / NT4;
  p.n_full = B * H * p.full_groups;
  p.n_items = p.n_full + (tiles % NT4 ? B * H : 0);
  static const int split_env = [] { const char* e = getenv("COUNTR_ATTN4_SPLIT"); return e ? atoi(e) : 1; }();
  p.split = split_env;
  const int grid = std::min(p.n_items, num_sms());
  COUNTR_CHECK_CUDA(launch_pdl(attention_fwd4_kernel<DH, kBf16>, dim3(grid), dim3(kThreads4), SM::kTotal, stream, tq, tkv, p));
  return COUNTR_OK;
}

template <int DH, bool kBf16>
int launch_attention(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, int bf16,
                     cudaStream_t stream) {
  using SM = AttnSmem<DH>;
  CUtensorMap tq, tkv;
  const uint64_t dims[4] = {(uint64_t)DH, (uint64_t)L, (uint64_t)(3 * H), (uint64_t)B};
  const uint64_t str[4] = {1, (uint64_t)(3 * H * DH), (uint64_t)DH, (uint64_t)L * 3 * H * DH};
  const uint32_t boxq[4] = {DH, BQ, 1, 1}, boxk[4] = {DH, KC, 1, 1};
  const TmapSwizzle sw = DH == 64 ? TMAP_SW_128 : TMAP_SW_64;
  int rc = make_tmap_4d_16b(&tq, qkv, dims, str, boxq, sw);
  if (rc) return rc;
  rc = make_tmap_4d_16b(&tkv, qkv, dims, str, boxk, sw);
  if (rc) return rc;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<DH, kBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  }
  AttnArgs p;
  p.out = reinterpret_cast<uint16_t*>(out);
  p.lse = lse;
  p.B = B; p.L = L; p.H = H;
  p.scale_log2 = scale * 1.44269504088896340736f;
  p.bf16 = bf16;
  p.q_pairs = ((L + BQ - 1) / BQ + 1) / 2;
  const int grid = std::min(B * H * p.q_pairs, num_sms());   // persistent: one CTA per SM walks the work items
  COUNTR_CHECK_CUDA(launch_pdl(attention_fwd_kernel<DH, kBf16>, dim3(grid), dim3(kThreads), SM::kTotal, stream, tq, tkv, p));
  return COUNTR_OK;
}


// ================================================================================================
// Fused self-attention BACKWARD (flash-style) for head_dim 32 — the FIM self-attention of the fine-tune step.
//
//   P = exp(scale*Q K^T - lse),  dP = dO V^T,  dS = scale * P * (dP - delta),  delta = rowsum(dO * O)
//   dV = P^T dO,   dK = dS^T Q,   dQ = dS K
//
// replaces: autograd of Attention.forward (models_crossvit.py:87-91): bmm / _softmax_backward_data / bmm's.
// One CTA = one (batch, head): Q, K, V, dO of the whole head (L <= 640) stay in shared memory (160 KB, TMA,
// SWIZZLE_64B), every accumulator stays in tensor memory: dV_j, dK_j (32 columns each), dQ_i for all query
// blocks (5 x 32), S and dP (128 each) = 480 of 512 columns.  No [B,H,L,L] tensor ever reaches HBM (the
// unfused path writes and re-reads four of them, 85 MB each at B=8).
// Out-of-range rows are zero-filled by TMA and get lse = delta = 0, which makes every masked contribution
// vanish without explicit masking.
// ================================================================================================
constexpr int kBwdThreads = 256;
constexpr int kBwdMaxBlk = 5;

struct AttnBwdArgs {
  const uint16_t* dO;   // [B*L][H*dh]
  const uint16_t* O;    // [B*L][H*dh]
  const float* lse;     // [B][H][L]
  uint16_t* dqkv;       // [B][L][3][H][dh]
  int B, L, H;
  float scale;
  int bf16;
};

template <int DH>
__global__ void __launch_bounds__(kBwdThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                     const AttnBwdArgs p) {
  static_assert(DH == 32, "fused attention backward is built for head_dim 32");
  constexpr uint32_t kBlkBytes = 128 * DH * 2;       // 8 KB: one 128-row block of Q / K / V / dO
  constexpr uint32_t kTileBytes = 128 * 128 * 2;     // 32 KB: P or dS tile
  constexpr uint32_t kOffQ = 0, kOffK = kBwdMaxBlk * kBlkBytes, kOffV = 2 * kBwdMaxBlk * kBlkBytes, kOffdO = 3 * kBwdMaxBlk * kBlkBytes;
  constexpr uint32_t kOffP = 4 * kBwdMaxBlk * kBlkBytes, kOffdS = kOffP + kTileBytes, kOffBar = kOffdS + kTileBytes;
  constexpr uint32_t tdV = 0, tdK = DH, tdQ = 2 * DH, tS = 256, tP = 384;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_s = bar_load + 1;
  uint64_t* bar_e = bar_load + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 3);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp: uniform for the compiler
  const int quarter = warp & 3, half = warp >> 2;
  const int r = quarter * 32 + lane;                 // row inside a 128-row block
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int nblk = (p.L + 127) / 128;
  const int D = p.H * DH;

  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_e, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  pdl_wait();   // qkv / out / dout come from earlier kernels

  // warp 0 issues every TMA / MMA: whole-warp control flow, elect.sync around the issue (see attention_fwd_kernel)
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_load, 4u * nblk * kBlkBytes);
      for (int i = 0; i < nblk; ++i) {
        tma_load_4d(smem + kOffQ + i * kBlkBytes, &tma_qkv, bar_load, 0, i * 128, h, b);
        tma_load_4d(smem + kOffK + i * kBlkBytes, &tma_qkv, bar_load, 0, i * 128, p.H + h, b);
        tma_load_4d(smem + kOffV + i * kBlkBytes, &tma_qkv, bar_load, 0, i * 128, 2 * p.H + h, b);
        tma_load_4d(smem + kOffdO + i * kBlkBytes, &tma_do, bar_load, 0, i * 128, h, b);
      }
    }
    __syncwarp();
  }

  // per-row constants of this thread's query row in every query block (overlaps the TMA loads)
  const float sl2 = p.scale * 1.44269504088896340736f;
  float lse2[kBwdMaxBlk], delta[kBwdMaxBlk];
#pragma unroll
  for (int i = 0; i < kBwdMaxBlk; ++i) {
    lse2[i] = 0.f;
    delta[i] = 0.f;
    const int q = i * 128 + r;
    if (i < nblk && q < p.L) {
      lse2[i] = p.lse[(static_cast<long long>(b) * p.H + h) * p.L + q] * 1.44269504088896340736f;
      const uint4* po = reinterpret_cast<const uint4*>(p.O + (static_cast<long long>(b) * p.L + q) * D + h * DH);
      const uint4* pd = reinterpret_cast<const uint4*>(p.dO + (static_cast<long long>(b) * p.L + q) * D + h * DH);
      float acc = 0.f;
#pragma unroll
      for (int v = 0; v < DH / 8; ++v) {
        const uint4 a = po[v], c = pd[v];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          float2 fa, fc;
          if (p.bf16) {
            fa = make_float2(__uint_as_float(aw[w] << 16), __uint_as_float(aw[w] & 0xffff0000u));
            fc = make_float2(__uint_as_float(cw[w] << 16), __uint_as_float(cw[w] & 0xffff0000u));
          } else {
            fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[w]));
            fc = __half22float2(*reinterpret_cast<const __half2*>(&cw[w]));
          }
          acc += fa.x * fc.x + fa.y * fc.y;
        }
      }
      delta[i] = acc;
    }
  }

  const uint32_t idesc_s = make_idesc_f16(128, 128, false, false, p.bf16 != 0);
  const uint32_t idesc_mm = make_idesc_f16(128, DH, true, true, p.bf16 != 0);   // A = P^T / dS^T (MN-major), B MN-major
  const uint32_t idesc_km = make_idesc_f16(128, DH, false, true, p.bf16 != 0);  // A = dS (K-major), B MN-major
  const uint32_t sP = smem_u32(smem + kOffP), sdS = smem_u32(smem + kOffdS);

  int pair = 0;
  for (int j = 0; j < nblk; ++j) {
    for (int i = 0; i < nblk; ++i, ++pair) {
      if (warp == 0) {
        if (pair == 0) mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint64_t dq = make_desc(smem_u32(smem + kOffQ + i * kBlkBytes), 16, 512, 4u);
        const uint64_t dk = make_desc(smem_u32(smem + kOffK + j * kBlkBytes), 16, 512, 4u);
        const uint64_t dd = make_desc(smem_u32(smem + kOffdO + i * kBlkBytes), 16, 512, 4u);
        const uint64_t dv = make_desc(smem_u32(smem + kOffV + j * kBlkBytes), 16, 512, 4u);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + tP, dd + 2 * k, dv + 2 * k, idesc_s, k != 0);
          umma_commit(bar_s);
        }
        __syncwarp();
      }
      // S, dP of this pair are ready (=> the dV/dK/dQ MMAs of the previous pair have drained: P/dS tiles are free)
      mbar_wait(bar_s, pair & 1);
      tc_fence_after();
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
        const int g = half * 2 + gi;
        uint32_t rs[32], rp[32];
        tmem_ld_32x32b_x32(t_lane + tS + g * 32, rs);
        tmem_ld_32x32b_x32(t_lane + tP + g * 32, rp);
        tmem_ld_wait();
        float pv[32], dsv[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          pv[c] = ex2_approx(fmaf(__uint_as_float(rs[c]), sl2, -lse2[i]));
          dsv[c] = pv[c] * (__uint_as_float(rp[c]) - delta[i]) * p.scale;
        }
        const uint32_t row_off = (g >> 1) * (128 * 128) + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o, o2;
          o.x = pack2(pv[8 * q + 0], pv[8 * q + 1], p.bf16);
          o.y = pack2(pv[8 * q + 2], pv[8 * q + 3], p.bf16);
          o.z = pack2(pv[8 * q + 4], pv[8 * q + 5], p.bf16);
          o.w = pack2(pv[8 * q + 6], pv[8 * q + 7], p.bf16);
          o2.x = pack2(dsv[8 * q + 0], dsv[8 * q + 1], p.bf16);
          o2.y = pack2(dsv[8 * q + 2], dsv[8 * q + 3], p.bf16);
          o2.z = pack2(dsv[8 * q + 4], dsv[8 * q + 5], p.bf16);
          o2.w = pack2(dsv[8 * q + 6], dsv[8 * q + 7], p.bf16);
          const uint32_t off = row_off + ((((g & 1) * 4 + q) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(smem + kOffP + off) = o;
          *reinterpret_cast<uint4*>(smem + kOffdS + off) = o2;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        const uint32_t aQ = smem_u32(smem + kOffQ + i * kBlkBytes), adO = smem_u32(smem + kOffdO + i * kBlkBytes);
        const uint32_t aK = smem_u32(smem + kOffK + j * kBlkBytes);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {   // 16 query rows (dV, dK) / 16 keys (dQ) per step
            const uint64_t p_mn = make_desc(sP + ks * 2048, 16384, 1024, 2u);
            const uint64_t ds_mn = make_desc(sdS + ks * 2048, 16384, 1024, 2u);
            const uint64_t ds_k = make_desc(sdS + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024, 2u);
            umma_f16_ss(tmem_base + tdV, p_mn, make_desc(adO + ks * 1024, 16, 512, 4u), idesc_mm, (i | ks) != 0);
            umma_f16_ss(tmem_base + tdK, ds_mn, make_desc(aQ + ks * 1024, 16, 512, 4u), idesc_mm, (i | ks) != 0);
            umma_f16_ss(tmem_base + tdQ + i * DH, ds_k, make_desc(aK + ks * 1024, 16, 512, 4u), idesc_km, (j | ks) != 0);
          }
          if (i == nblk - 1) umma_commit(bar_e);
        }
        __syncwarp();
      }
    }
    // dV_j, dK_j complete: thread (key row r, column half) writes its 16 columns of each
    mbar_wait(bar_e, j & 1);
    tc_fence_after();
    {
      const int key = j * 128 + r;
      uint32_t rv[16], rk[16];
      tmem_ld_32x32b_x16(t_lane + tdV + half * 16, rv);
      tmem_ld_32x32b_x16(t_lane + tdK + half * 16, rk);
      tmem_ld_wait();
      if (key < p.L) {
        uint16_t* base = p.dqkv + (static_cast<long long>(b) * p.L + key) * (3 * D) + h * DH + half * 16;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 ov, ok;
          ov.x = pack2(__uint_as_float(rv[8 * q + 0]), __uint_as_float(rv[8 * q + 1]), p.bf16);
          ov.y = pack2(__uint_as_float(rv[8 * q + 2]), __uint_as_float(rv[8 * q + 3]), p.bf16);
          ov.z = pack2(__uint_as_float(rv[8 * q + 4]), __uint_as_float(rv[8 * q + 5]), p.bf16);
          ov.w = pack2(__uint_as_float(rv[8 * q + 6]), __uint_as_float(rv[8 * q + 7]), p.bf16);
          ok.x = pack2(__uint_as_float(rk[8 * q + 0]), __uint_as_float(rk[8 * q + 1]), p.bf16);
          ok.y = pack2(__uint_as_float(rk[8 * q + 2]), __uint_as_float(rk[8 * q + 3]), p.bf16);
          ok.z = pack2(__uint_as_float(rk[8 * q + 4]), __uint_as_float(rk[8 * q + 5]), p.bf16);
          ok.w = pack2(__uint_as_float(rk[8 * q + 6]), __uint_as_float(rk[8 * q + 7]), p.bf16);
          *reinterpret_cast<uint4*>(base + 2 * D + 8 * q) = ov;   // V slot
          *reinterpret_cast<uint4*>(base + D + 8 * q) = ok;       // K slot
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // dV/dK TMEM columns are overwritten by the next key block
  }
  // dQ of every query block (the bar_e commit of the last key block covered all MMAs)
  for (int i = 0; i < nblk; ++i) {
    const int q = i * 128 + r;
    uint32_t rq[16];
    tmem_ld_32x32b_x16(t_lane + tdQ + i * DH + half * 16, rq);
    tmem_ld_wait();
    if (q < p.L) {
      uint16_t* base = p.dqkv + (static_cast<long long>(b) * p.L + q) * (3 * D) + h * DH + half * 16;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint4 o;
        o.x = pack2(__uint_as_float(rq[8 * c + 0]), __uint_as_float(rq[8 * c + 1]), p.bf16);
        o.y = pack2(__uint_as_float(rq[8 * c + 2]), __uint_as_float(rq[8 * c + 3]), p.bf16);
        o.z = pack2(__uint_as_float(rq[8 * c + 4]), __uint_as_float(rq[8 * c + 5]), p.bf16);
        o.w = pack2(__uint_as_float(rq[8 * c + 6]), __uint_as_float(rq[8 * c + 7]), p.bf16);
        *reinterpret_cast<uint4*>(base + 8 * c) = o;              // Q slot
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ================================================================================================
// Fused self-attention BACKWARD for head_dim 64 (the encoder of the MAE pre-training step, models_mae_noct.py:137-152
// under autograd).  Whole-head residency does not fit at dh = 64 (Q, K, V, dO of 576 tokens = 288 KB; dQ of every query
// block + dV + dK + S + dP = 704 TMEM columns), so the work is split FlashAttention-2 style:
//   one CTA = one (batch, head, 128-KEY block j): K_j, V_j stay in shared memory, the query blocks stream through
//   (Q_i, dO_i), dV_j / dK_j accumulate in tensor memory over i, and each pair's dQ_i contribution (128 x 64 fp32)
//   is added to an fp32 [B][L][H][64] workspace with vector reductions; countr_attention_bwd converts the workspace
//   into the Q slots of dqkv afterwards.  TMEM: dV 64 + dK 64 + dQ 64 + S 128 + dP 128 = 448 columns.
// Any L (one CTA per key block); out-of-range rows are zero-filled by TMA and get lse = delta = 0.
// ================================================================================================
__global__ void __launch_bounds__(kBwdThreads, 1)
attention_bwd64_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                       const AttnBwdArgs p, float* __restrict__ dq32) {
  constexpr int DH = 64;
  constexpr uint32_t kBlkBytes = 128 * DH * 2;       // 16 KB
  constexpr uint32_t kTileBytes = 128 * 128 * 2;     // 32 KB: P or dS tile
  constexpr uint32_t kOffK = 0, kOffV = kBlkBytes, kOffQ = 2 * kBlkBytes /*[2 buffers]*/, kOffdO = 4 * kBlkBytes /*[2]*/;
  constexpr uint32_t kOffP = 6 * kBlkBytes, kOffdS = kOffP + kTileBytes, kOffBar = kOffdS + kTileBytes;
  constexpr uint32_t tdV = 0, tdK = 64, tdQ = 128, tS = 256, tP = 384;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_q = bar_kv + 1;      // [2]: Q_i / dO_i landed in buffer i & 1
  uint64_t* bar_s = bar_kv + 3;      // S, dP of the pair are in TMEM
  uint64_t* bar_e = bar_kv + 4;      // dV / dK / dQ MMAs of the pair have completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_kv + 5);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int r = quarter * 32 + lane;                 // row inside a 128-row block
  const int nblk = (p.L + 127) / 128;
  const int j = blockIdx.x % nblk;
  const int bh = blockIdx.x / nblk;
  const int h = bh % p.H, b = bh / p.H;
  const int D = p.H * DH;

  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_q + 1, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_e, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  pdl_wait();   // qkv / out / dout come from earlier kernels, dq32 was zeroed by one

  auto load_q = [&](int i) {       // warp 0, whole warp
    const int buf = i & 1;
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q + buf, 2u * kBlkBytes);
      tma_load_4d(smem + kOffQ + buf * kBlkBytes, &tma_qkv, bar_q + buf, 0, i * 128, h, b);
      tma_load_4d(smem + kOffdO + buf * kBlkBytes, &tma_do, bar_q + buf, 0, i * 128, h, b);
    }
    __syncwarp();
  };
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_kv, 2u * kBlkBytes);
      tma_load_4d(smem + kOffK, &tma_qkv, bar_kv, 0, j * 128, p.H + h, b);
      tma_load_4d(smem + kOffV, &tma_qkv, bar_kv, 0, j * 128, 2 * p.H + h, b);
    }
    __syncwarp();
    load_q(0);
    if (nblk > 1) load_q(1);
  }

  const float sl2 = p.scale * 1.44269504088896340736f;
  const uint32_t idesc_s = make_idesc_f16(128, 128, false, false, p.bf16 != 0);
  const uint32_t idesc_mm = make_idesc_f16(128, DH, true, true, p.bf16 != 0);   // A = P^T / dS^T (MN-major), B MN-major
  const uint32_t idesc_km = make_idesc_f16(128, DH, false, true, p.bf16 != 0);  // A = dS (K-major), B MN-major
  const uint32_t sP = smem_u32(smem + kOffP), sdS = smem_u32(smem + kOffdS);
  const uint32_t aK = smem_u32(smem + kOffK), aV = smem_u32(smem + kOffV);

  for (int i = 0; i < nblk; ++i) {
    const int buf = i & 1;
    // per-row constants of this thread's query row (overlaps the TMA loads / the previous pair's MMAs)
    float lse2 = 0.f, delta = 0.f;
    {
      const int q = i * 128 + r;
      if (q < p.L) {
        lse2 = p.lse[(static_cast<long long>(b) * p.H + h) * p.L + q] * 1.44269504088896340736f;
        const uint4* po = reinterpret_cast<const uint4*>(p.O + (static_cast<long long>(b) * p.L + q) * D + h * DH);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dO + (static_cast<long long>(b) * p.L + q) * D + h * DH);
        float acc = 0.f;
#pragma unroll
        for (int v = 0; v < DH / 8; ++v) {
          const uint4 a = po[v], c = pd[v];
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            float2 fa, fc;
            if (p.bf16) {
              fa = make_float2(__uint_as_float(aw[w] << 16), __uint_as_float(aw[w] & 0xffff0000u));
              fc = make_float2(__uint_as_float(cw[w] << 16), __uint_as_float(cw[w] & 0xffff0000u));
            } else {
              fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[w]));
              fc = __half22float2(*reinterpret_cast<const __half2*>(&cw[w]));
            }
            acc += fa.x * fc.x + fa.y * fc.y;
          }
        }
        delta = acc;
      }
    }
    const uint32_t aQ = smem_u32(smem + kOffQ + buf * kBlkBytes), adO = smem_u32(smem + kOffdO + buf * kBlkBytes);
    if (warp == 0) {
      if (i == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q + buf, (i >> 1) & 1);
      tc_fence_after();
      const uint64_t dq = make_desc(aQ, 16, 1024, 2u), dk = make_desc(aK, 16, 1024, 2u);
      const uint64_t dd = make_desc(adO, 16, 1024, 2u), dv = make_desc(aV, 16, 1024, 2u);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + tP, dd + 2 * k, dv + 2 * k, idesc_s, k != 0);
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    mbar_wait(bar_s, i & 1);
    tc_fence_after();
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int g = half * 2 + gi;
      uint32_t rs[32], rp[32];
      tmem_ld_32x32b_x32(t_lane + tS + g * 32, rs);
      tmem_ld_32x32b_x32(t_lane + tP + g * 32, rp);
      tmem_ld_wait();
      float pv[32], dsv[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        pv[c] = ex2_approx(fmaf(__uint_as_float(rs[c]), sl2, -lse2));
        dsv[c] = pv[c] * (__uint_as_float(rp[c]) - delta) * p.scale;
      }
      const uint32_t row_off = (g >> 1) * (128 * 128) + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o, o2;
        o.x = pack2(pv[8 * q + 0], pv[8 * q + 1], p.bf16);
        o.y = pack2(pv[8 * q + 2], pv[8 * q + 3], p.bf16);
        o.z = pack2(pv[8 * q + 4], pv[8 * q + 5], p.bf16);
        o.w = pack2(pv[8 * q + 6], pv[8 * q + 7], p.bf16);
        o2.x = pack2(dsv[8 * q + 0], dsv[8 * q + 1], p.bf16);
        o2.y = pack2(dsv[8 * q + 2], dsv[8 * q + 3], p.bf16);
        o2.z = pack2(dsv[8 * q + 4], dsv[8 * q + 5], p.bf16);
        o2.w = pack2(dsv[8 * q + 6], dsv[8 * q + 7], p.bf16);
        const uint32_t off = row_off + ((((g & 1) * 4 + q) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + kOffP + off) = o;
        *reinterpret_cast<uint4*>(smem + kOffdS + off) = o2;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 query rows (dV, dK) / 16 keys (dQ) per step
          const uint64_t p_mn = make_desc(sP + ks * 2048, 16384, 1024, 2u);
          const uint64_t ds_mn = make_desc(sdS + ks * 2048, 16384, 1024, 2u);
          const uint64_t ds_k = make_desc(sdS + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024, 2u);
          umma_f16_ss(tmem_base + tdV, p_mn, make_desc(adO + ks * 2048, 16, 1024, 2u), idesc_mm, (i | ks) != 0);
          umma_f16_ss(tmem_base + tdK, ds_mn, make_desc(aQ + ks * 2048, 16, 1024, 2u), idesc_mm, (i | ks) != 0);
          umma_f16_ss(tmem_base + tdQ, ds_k, make_desc(aK + ks * 2048, 16, 1024, 2u), idesc_km, ks != 0);
        }
        umma_commit(bar_e);
      }
      __syncwarp();
    }
    // dQ contribution of this pair: thread (query row r, column half) adds its 32 columns to the fp32 workspace
    mbar_wait(bar_e, i & 1);
    tc_fence_after();
    if (warp == 0 && i + 2 < nblk) load_q(i + 2);     // this pair's Q / dO buffer is free again
    {
      const int q = i * 128 + r;
      uint32_t rq[32];
      tmem_ld_32x32b_x32(t_lane + tdQ + half * 32, rq);
      tmem_ld_wait();
      if (q < p.L) {
        float* dst = dq32 + ((static_cast<long long>(b) * p.L + q) * p.H + h) * DH + half * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c), "f"(__uint_as_float(rq[4 * c])),
                       "f"(__uint_as_float(rq[4 * c + 1])), "f"(__uint_as_float(rq[4 * c + 2])), "f"(__uint_as_float(rq[4 * c + 3]))
                       : "memory");
      }
    }
    tc_fence_before();
    __syncthreads();   // S / dP / dQ columns and the P / dS tiles are rewritten by the next pair
  }
  // dV_j, dK_j complete (the last bar_e covered every MMA): thread (key row r, column half) writes its 32 columns of each
  tc_fence_after();
  {
    const int key = j * 128 + r;
    uint32_t rv[32], rk[32];
    tmem_ld_32x32b_x32(t_lane + tdV + half * 32, rv);
    tmem_ld_32x32b_x32(t_lane + tdK + half * 32, rk);
    tmem_ld_wait();
    if (key < p.L) {
      uint16_t* base = p.dqkv + (static_cast<long long>(b) * p.L + key) * (3 * D) + h * DH + half * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 ov, ok;
        ov.x = pack2(__uint_as_float(rv[8 * q + 0]), __uint_as_float(rv[8 * q + 1]), p.bf16);
        ov.y = pack2(__uint_as_float(rv[8 * q + 2]), __uint_as_float(rv[8 * q + 3]), p.bf16);
        ov.z = pack2(__uint_as_float(rv[8 * q + 4]), __uint_as_float(rv[8 * q + 5]), p.bf16);
        ov.w = pack2(__uint_as_float(rv[8 * q + 6]), __uint_as_float(rv[8 * q + 7]), p.bf16);
        ok.x = pack2(__uint_as_float(rk[8 * q + 0]), __uint_as_float(rk[8 * q + 1]), p.bf16);
        ok.y = pack2(__uint_as_float(rk[8 * q + 2]), __uint_as_float(rk[8 * q + 3]), p.bf16);
        ok.z = pack2(__uint_as_float(rk[8 * q + 4]), __uint_as_float(rk[8 * q + 5]), p.bf16);
        ok.w = pack2(__uint_as_float(rk[8 * q + 6]), __uint_as_float(rk[8 * q + 7]), p.bf16);
        *reinterpret_cast<uint4*>(base + 2 * D + 8 * q) = ov;   // V slot
        *reinterpret_cast<uint4*>(base + D + 8 * q) = ok;       // K slot
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// Generic-head-size forward (CUDA cores, fp32): head dims the tcgen05 kernels are not instantiated for — mae_vit_huge_patch14
// (models_mae_cross.py:226-231: 1280 / 16 = 80 channels per head, 27 x 27 = 729 tokens).  One warp per query row, 8 rows per
// block; K / V stream through shared memory in 64-key chunks shared by the 8 warps; lanes own keys for Q.K^T and output
// channels for P.V; online softmax in fp32.  Same arithmetic and outputs (out, lse) as attention_fwd_kernel.  Not a
// benchmarked configuration: it exists so that every factory of the reference runs.
// ================================================================================================
constexpr int kGenMaxDh = 128, kGenKeys = 64;
__global__ void __launch_bounds__(256) attention_fwd_generic_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out,
                                                                     float* __restrict__ lse, int L, int H, int dh, float scale_log2,
                                                                     int bf16) {
  extern __shared__ float gsm[];
  const int pitch = dh + 1;                       // odd pitch in words: conflict-free row-per-lane reads
  float* sk = gsm;                                // [64][pitch]
  float* sv = sk + kGenKeys * pitch;              // [64][pitch]
  float* sq = sv + kGenKeys * pitch;              // [8][dh]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * 8 + warp;
  const long long row = 3ll * H * dh;
  pdl_trigger();
  pdl_wait();
  auto ld16 = [&](const uint16_t* p) {
    return bf16 ? __uint_as_float(static_cast<uint32_t>(*p) << 16) : __half2float(*reinterpret_cast<const __half*>(p));
  };
  if (q < L)
    for (int d = lane; d < dh; d += 32) sq[warp * dh + d] = ld16(qkv + (static_cast<long long>(b) * L + q) * row + h * dh + d);
  float m_run = -INFINITY, l_run = 0.f;
  float o[kGenMaxDh / 32];
#pragma unroll
  for (int t = 0; t < kGenMaxDh / 32; ++t) o[t] = 0.f;
  for (int k0 = 0; k0 < L; k0 += kGenKeys) {
    __syncthreads();
    for (int i = threadIdx.x; i < kGenKeys * dh; i += blockDim.x) {
      const int kk = i / dh, d = i - kk * dh;
      const int key = k0 + kk;
      float kv = 0.f, vv = 0.f;
      if (key < L) {
        const uint16_t* base = qkv + (static_cast<long long>(b) * L + key) * row + h * dh + d;
        kv = ld16(base + static_cast<long long>(H) * dh);
        vv = ld16(base + 2ll * H * dh);
      }
      sk[kk * pitch + d] = kv;
      sv[kk * pitch + d] = vv;
    }
    __syncthreads();
    if (q >= L) continue;
    float s0 = 0.f, s1 = 0.f;
    for (int d = 0; d < dh; ++d) {
      const float qd = sq[warp * dh + d];
      s0 = fmaf(qd, sk[lane * pitch + d], s0);
      s1 = fmaf(qd, sk[(lane + 32) * pitch + d], s1);
    }
    s0 = (k0 + lane < L) ? s0 * scale_log2 : -INFINITY;
    s1 = (k0 + lane + 32 < L) ? s1 * scale_log2 : -INFINITY;
    const float mx = warp_max(fmaxf(s0, s1));
    const float m_new = fmaxf(m_run, mx);
    const float corr = exp2f(m_run - m_new);         // first chunk: exp2(-inf) = 0
    const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
    l_run = l_run * corr + warp_sum(p0 + p1);
    m_run = m_new;
#pragma unroll
    for (int t = 0; t < kGenMaxDh / 32; ++t) o[t] *= corr;
    for (int j = 0; j < kGenKeys; ++j) {
      const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
#pragma unroll
      for (int t = 0; t < kGenMaxDh / 32; ++t) {
        const int d = lane + 32 * t;
        if (d < dh) o[t] = fmaf(pj, sv[j * pitch + d], o[t]);
      }
    }
  }
  if (q >= L) return;
  const float inv = 1.f / l_run;
  uint16_t* orow = out + (static_cast<long long>(b) * L + q) * (H * dh) + h * dh;
#pragma unroll
  for (int t = 0; t < kGenMaxDh / 32; ++t) {
    const int d = lane + 32 * t;
    if (d < dh) orow[d] = static_cast<uint16_t>(pack2(o[t] * inv, 0.f, bf16) & 0xffffu);
  }
  if (lse != nullptr && lane == 0) lse[(static_cast<long long>(b) * H + h) * L + q] = (m_run + log2f(l_run)) * 0.69314718055994531f;
}

// dq32 [B*L][H*64] fp32 -> the Q slots of dqkv [B*L][3][H*64] (16-bit); thread = 8 columns
__global__ void __launch_bounds__(256) dq_cast_kernel(const float* __restrict__ dq32, uint16_t* __restrict__ dqkv, long long rows, int D,
                                                      int bf16) {
  pdl_trigger();
  pdl_wait();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per_row = D / 8;
  if (idx >= rows * per_row) return;
  const long long row = idx / per_row;
  const int c = static_cast<int>(idx - row * per_row) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dq32 + row * D + c), b4 = *reinterpret_cast<const float4*>(dq32 + row * D + c + 4);
  uint4 o;
  o.x = pack2(a.x, a.y, bf16); o.y = pack2(a.z, a.w, bf16); o.z = pack2(b4.x, b4.y, bf16); o.w = pack2(b4.z, b4.w, bf16);
  *reinterpret_cast<uint4*>(dqkv + row * 3 * D + c) = o;
}

}  // namespace
}  // namespace countr

#ifdef COUNTR_TRACE
extern "C" int countr_debug_set_attn_trace(void* buf) {
  long long* pbuf = reinterpret_cast<long long*>(buf);
  COUNTR_CHECK_CUDA(cudaMemcpyToSymbol(countr::g_attn_trace, &pbuf, sizeof(pbuf)));
  return 0;
}
#endif

extern "C" int countr_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int dh, float scale,
                                    int bf16, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(qkv && out, "null pointer");
  COUNTR_REQUIRE(B > 0 && L > 0 && H > 0, "bad shape B=%d L=%d H=%d", B, L, H);
  // COUNTR_ATTN4: 0 = two-tile kernel everywhere, 1 (default) = four-tile kernel for head_dim 64 and for head_dim 32 at large
  // batch, 2 = always.  (head_dim 32, B = 8, L = 576 measured slower, 60 vs 28 us: the 64-row remainder items of 128 heads pile up
  // on the 20 CTAs that have no four-tile item.)
  static int four = -1;
  if (four < 0) {
    const char* e = getenv("COUNTR_ATTN4");
    four = e != nullptr ? atoi(e) : 1;
  }
  if (four && dh == 64) return bf16 ? launch_attention4<64, true>(qkv, out, lse, B, L, H, scale, stream)
                                    : launch_attention4<64, false>(qkv, out, lse, B, L, H, scale, stream);
  // head_dim 32: only when every CTA gets four-tile items anyway (B*H >= SMs: pre-training decoder, B = 128 inference)
  if ((four >= 2 || (four == 1 && B * H * ((L + BQ - 1) / BQ / NT4) >= num_sms())) && dh == 32) return bf16 ? launch_attention4<32, true>(qkv, out, lse, B, L, H, scale, stream)
                                    : launch_attention4<32, false>(qkv, out, lse, B, L, H, scale, stream);
  if (dh == 64) return bf16 ? launch_attention<64, true>(qkv, out, lse, B, L, H, scale, bf16, stream)
                            : launch_attention<64, false>(qkv, out, lse, B, L, H, scale, bf16, stream);
  if (dh == 32) return bf16 ? launch_attention<32, true>(qkv, out, lse, B, L, H, scale, bf16, stream)
                            : launch_attention<32, false>(qkv, out, lse, B, L, H, scale, bf16, stream);
  COUNTR_REQUIRE(dh >= 8 && dh <= kGenMaxDh && dh % 8 == 0, "attention head_dim %d not supported (32 / 64 on tensor cores, any multiple of 8 up to %d otherwise)",
                 dh, kGenMaxDh);
  {
    // head sizes without a tcgen05 instantiation (mae_vit_huge_patch14: 80): generic CUDA-core kernel
    const size_t smem = (2ull * kGenKeys * (dh + 1) + 8ull * dh) * sizeof(float);
    static PerDeviceOnce attr_gen;
    if (attr_gen.need())
      COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    COUNTR_CHECK_CUDA(launch_pdl(attention_fwd_generic_kernel, dim3((L + 7) / 8, H, B), dim3(256), smem, stream,
                                 reinterpret_cast<const uint16_t*>(qkv), reinterpret_cast<uint16_t*>(out), lse, L, H, dh,
                                 scale * 1.44269504088896340736f, bf16));
    return COUNTR_OK;
  }
}

extern "C" int64_t countr_attention_bwd_workspace_bytes(int B, int L, int H, int dh) {
  return dh == 64 ? static_cast<int64_t>(B) * L * H * dh * 4 : 0;
}

extern "C" int countr_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* workspace,
                                    int B, int L, int H, int dh, float scale, int bf16, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(qkv && out && dout && lse && dqkv, "null pointer");
  COUNTR_REQUIRE(dh == 32 || dh == 64, "fused attention backward supports head_dim 32 and 64 (got %d)", dh);
  COUNTR_REQUIRE(L >= 1 && (dh == 64 || L <= 128 * kBwdMaxBlk), "fused attention backward (head_dim 32) supports L <= %d (got %d)",
                 128 * kBwdMaxBlk, L);
  COUNTR_REQUIRE(dh == 32 || (workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0),
                 "head_dim 64 needs a 16-byte aligned workspace of countr_attention_bwd_workspace_bytes()");
  const int D = H * dh;
  const TmapSwizzle sw = dh == 64 ? TMAP_SW_128 : TMAP_SW_64;
  CUtensorMap tq, td;
  {
    const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)L, (uint64_t)(3 * H), (uint64_t)B};
    const uint64_t str[4] = {1, (uint64_t)(3 * D), (uint64_t)dh, (uint64_t)L * 3 * D};
    const uint32_t box[4] = {(uint32_t)dh, 128, 1, 1};
    int rc = make_tmap_4d_16b(&tq, qkv, dims, str, box, sw);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {1, (uint64_t)D, (uint64_t)dh, (uint64_t)L * D};
    const uint32_t box[4] = {(uint32_t)dh, 128, 1, 1};
    int rc = make_tmap_4d_16b(&td, dout, dims, str, box, sw);
    if (rc) return rc;
  }
  AttnBwdArgs a;
  a.dO = reinterpret_cast<const uint16_t*>(dout);
  a.O = reinterpret_cast<const uint16_t*>(out);
  a.lse = lse;
  a.dqkv = reinterpret_cast<uint16_t*>(dqkv);
  a.B = B; a.L = L; a.H = H;
  a.scale = scale;
  a.bf16 = bf16;
  if (dh == 64) {
    constexpr uint32_t smem64 = 6 * (128 * 64 * 2) + 2 * (128 * 128 * 2) + 64 + 1024;
    static PerDeviceOnce attr64;
    if (attr64.need())
      COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64));
    float* dq32 = reinterpret_cast<float*>(workspace);
    const long long rows = static_cast<long long>(B) * L;
    COUNTR_CHECK_CUDA(cudaMemsetAsync(dq32, 0, static_cast<size_t>(rows) * D * 4, stream));
    const int nblk = (L + 127) / 128;
    COUNTR_CHECK_CUDA(launch_pdl(attention_bwd64_kernel, dim3(B * H * nblk), dim3(kBwdThreads), smem64, stream, tq, td, a, dq32));
    const long long n = rows * (D / 8);
    COUNTR_CHECK_CUDA(launch_pdl(dq_cast_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, stream,
                                 static_cast<const float*>(dq32), a.dqkv, rows, D, bf16));
    return COUNTR_OK;
  }
  constexpr uint32_t smem_bytes = 4 * kBwdMaxBlk * (128 * 32 * 2) + 2 * (128 * 128 * 2) + 64 + 1024;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  COUNTR_CHECK_CUDA(launch_pdl(attention_bwd_kernel<32>, dim3(B * H), dim3(kBwdThreads), smem_bytes, stream, tq, td, a));
  return COUNTR_OK;
}
