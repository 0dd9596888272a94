// countr_b200 — fused multi-head self-attention forward (flash-style, tcgen05 + TMEM).
//
//   out[b, q, h*dh:(h+1)*dh] = softmax_k( scale * Q[b,h,q,:] . K[b,h,k,:] ) @ V[b,h,:,:]
//
// replaces: Attention.forward lines models_crossvit.py:87-91 (q@k^T * scale, softmax, attn@v,
// transpose/reshape) for both the ViT encoder blocks (dh=64, H=12) and the FIM self-attention
// (dh=32, H=16).  The [B,H,L,L] score tensor the reference materialises never leaves the SM.
//
// One CTA (8 warps) = one (batch, head, 128-query tile).  K/V are streamed in 128-key chunks by
// TMA straight out of the packed qkv GEMM output [B][L][3][H][dh] (no permute / copy):
//   S  = Q K_c^T        tcgen05.mma  (A,B K-major in smem)      -> TMEM columns [64,192)
//   P  = exp2(S*c - m)  thread == row, online max/sum in fp32   -> smem (fp16, SWIZZLE_128B)
//   O += P V_c          tcgen05.mma  (B = V chunk, MN-major)    -> TMEM columns [0,dh)
// Two CTAs are co-resident per SM (80 KB smem, 256 TMEM columns each) so one CTA's softmax
// (MUFU-bound: 128x576 exp2 per tile) overlaps the other's MMA/TMA.
#include "../../include/countr_b200.h"
#include "common.cuh"
#include "tma.h"

namespace countr {
namespace {

constexpr int BQ = 128;   // query rows per CTA
constexpr int KC = 128;   // keys per chunk
constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemS = 64;  // S starts at this column; O occupies [0, dh)

template <int DH>
struct AttnSmem {
  static constexpr uint32_t kRowBytes = DH * 2;               // 128 (SW128) or 64 (SW64)
  static constexpr uint32_t kQBytes = BQ * kRowBytes;
  static constexpr uint32_t kKBytes = KC * kRowBytes;
  static constexpr uint32_t kPBytes = BQ * KC * 2;            // 2 column blocks of [128 x 128 B]
  static constexpr uint32_t kOffQ = 0;
  static constexpr uint32_t kOffK = kQBytes;
  static constexpr uint32_t kOffV = kOffK + kKBytes;
  static constexpr uint32_t kOffP = kOffV + kKBytes;
  static constexpr uint32_t kOffBar = kOffP + kPBytes;
  static constexpr uint32_t kOffXchg = kOffBar + 64;                 // [2][BQ] floats: row max / row sum exchange
  static constexpr uint32_t kTotal = kOffXchg + 2 * BQ * 4 + 1024;
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
  int q_tiles;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 8 warps: warp w owns TMEM lane quarter (w & 3) — i.e. query rows 32*(w&3) .. +31 — and the column half
// (w >> 2) of every 128-key chunk.  The two threads that share a row exchange their partial row maximum /
// row sum through shared memory.  Twice the warps per SM of a thread-per-row design: the softmax is
// latency-bound (MUFU + dependent FMAs), so it needs the extra warps to keep the issue slots busy.
template <int DH>
__global__ void __launch_bounds__(kThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                     const AttnArgs p) {
  using SM = AttnSmem<DH>;
  constexpr uint32_t kLayout = DH == 64 ? 2u : 4u;            // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t kSboK = DH == 64 ? 1024u : 512u;         // 8 rows of the K-major tiles
  constexpr uint32_t kVStep = 16 * SM::kRowBytes;             // 16 key rows per k-step of P.V
  constexpr int kOHalf = DH / 2;                              // O columns rescaled / written per thread

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + SM::kOffQ;
  uint8_t* sK = smem + SM::kOffK;
  uint8_t* sV = smem + SM::kOffV;
  uint8_t* sP = smem + SM::kOffP;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* bar_k = bar_q + 1;
  uint64_t* bar_v = bar_q + 2;
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 5);
  float* xchg = reinterpret_cast<float*>(smem + SM::kOffXchg);   // [2][BQ]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int quarter = warp & 3, half = warp >> 2;
  const int qt = blockIdx.x % p.q_tiles;
  const int h = (blockIdx.x / p.q_tiles) % p.H;
  const int b = blockIdx.x / (p.q_tiles * p.H);
  const int q0 = qt * BQ;
  const int nchunks = (p.L + KC - 1) / KC;

  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  pdl_wait();   // qkv is the previous kernel's output

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, SM::kQBytes);
    tma_load_4d(sQ, &tma_q, bar_q, 0, q0, h, b);
    mbar_arrive_expect_tx(bar_k, SM::kKBytes);
    tma_load_4d(sK, &tma_kv, bar_k, 0, 0, p.H + h, b);
  }

  const uint32_t idesc_s = make_idesc_f16(BQ, KC, false, false, p.bf16 != 0);
  const uint32_t idesc_o = make_idesc_f16(BQ, DH, false, true, p.bf16 != 0);

  float m_run = -INFINITY;  // running max (log2 domain, already scaled) — identical in both threads of a row
  float l_run = 0.f;        // running sum over this thread's columns
  const int row = quarter * 32 + (tid & 31);

  for (int c = 0; c < nchunks; ++c) {
    const int valid = min(KC, p.L - c * KC);          // keys in this chunk
    const int groups = (valid + 31) / 32;             // 32-column groups that hold any valid key
    if (tid == 0) {
      if (c == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, c & 1);
      tc_fence_after();
      const uint64_t a_desc = make_desc(smem_u32(sQ), 16, kSboK, kLayout);
      const uint64_t b_desc = make_desc(smem_u32(sK), 16, kSboK, kLayout);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_f16_ss(tmem_base + kTmemS, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k),
                    idesc_s, k != 0);
      umma_commit(bar_s);
    }
    // S_c ready  (=> every earlier MMA, in particular P.V of chunk c-1, has completed:
    // the K and V buffers and the P tile are free again)
    mbar_wait(bar_s, c & 1);
    tc_fence_after();
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_v, SM::kKBytes);
      tma_load_4d(sV, &tma_kv, bar_v, 0, c * KC, 2 * p.H + h, b);
      if (c + 1 < nchunks) {
        mbar_arrive_expect_tx(bar_k, SM::kKBytes);
        tma_load_4d(sK, &tma_kv, bar_k, 0, (c + 1) * KC, p.H + h, b);
      }
    }

    // ---- pass 1: maximum over this thread's 64 columns, then over the row ----
    float mx = -INFINITY;
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int g = half * 2 + gi;
      if (g < groups) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + kTmemS + g * 32, r);
        tmem_ld_wait();
        const int lim = valid - g * 32;
        if (lim >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < lim) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      }
    }
    xchg[half * BQ + row] = mx;
    __syncthreads();
    mx = fmaxf(mx, xchg[(half ^ 1) * BQ + row]);
    const float m_new = fmaxf(m_run, mx * p.scale_log2);
    const float corr = ex2_approx(m_run - m_new);  // 0 on the first chunk (m_run = -inf)
    m_run = m_new;

    // ---- pass 2: P = exp2(S*c - m), partial row sum, fp16 P tile into swizzled smem ----
    float lsum = 0.f;
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int g = half * 2 + gi;
      if (g < groups) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + kTmemS + g * 32, r);
        tmem_ld_wait();
        const int lim = valid - g * 32;
        float pv[32];
        if (lim >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -m_new));
            lsum += pv[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -m_new));
            pv[j] = (j < lim) ? e : 0.f;
            lsum += pv[j];
          }
        }
        // group g covers key columns [32g, 32g+32) = 16-byte chunks (g&1)*4 .. +3 of column block g>>1
        uint8_t* prow = sP + (g >> 1) * (BQ * 128) + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack2(pv[8 * q + 0], pv[8 * q + 1], p.bf16);
          o.y = pack2(pv[8 * q + 2], pv[8 * q + 3], p.bf16);
          o.z = pack2(pv[8 * q + 4], pv[8 * q + 5], p.bf16);
          o.w = pack2(pv[8 * q + 6], pv[8 * q + 7], p.bf16);
          const int chunk16 = (g & 1) * 4 + q;
          *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (row & 7)) << 4)) = o;
        }
      }
    }
    l_run = l_run * corr + lsum;

    // ---- rescale this thread's half of the running output (TMEM) when the maximum moved ----
    if (c > 0) {
#pragma unroll
      for (int d0 = 0; d0 < kOHalf; d0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_lane + half * kOHalf + d0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * corr);
        tmem_st_32x32b_x16(t_lane + half * kOHalf + d0, r);
      }
      tmem_st_wait();
    }

    fence_proxy_async_smem();  // P tile (generic-proxy stores) -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, c & 1);
      tc_fence_after();
      const int ksteps = groups * 2;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t a_desc = make_desc(smem_u32(sP) + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024, 2u);
        const uint64_t b_desc = make_desc(smem_u32(sV) + k * kVStep, 16, kSboK, kLayout);
        umma_f16_ss(tmem_base, a_desc, b_desc, idesc_o, (c | k) != 0);
      }
      if (c == nchunks - 1) umma_commit(bar_o);
    }
  }

  // ---- epilogue: O / l -> 16-bit; each thread writes its half of the dh-wide row segment ----
  xchg[half * BQ + row] = l_run;
  __syncthreads();
  const float l_tot = l_run + xchg[(half ^ 1) * BQ + row];
  mbar_wait(bar_o, 0);
  tc_fence_after();
  const float inv_l = 1.f / l_tot;
  const int q = q0 + row;
  uint16_t* orow = p.out + (static_cast<long long>(b) * p.L + q) * (p.H * DH) + h * DH + half * kOHalf;
#pragma unroll
  for (int d0 = 0; d0 < kOHalf; d0 += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_lane + half * kOHalf + d0, r);
    tmem_ld_wait();
    if (q < p.L) {
      uint4 o0, o1;
      o0.x = pack2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l, p.bf16);
      o0.y = pack2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l, p.bf16);
      o0.z = pack2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l, p.bf16);
      o0.w = pack2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l, p.bf16);
      o1.x = pack2(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l, p.bf16);
      o1.y = pack2(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l, p.bf16);
      o1.z = pack2(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l, p.bf16);
      o1.w = pack2(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l, p.bf16);
      *reinterpret_cast<uint4*>(orow + d0) = o0;
      *reinterpret_cast<uint4*>(orow + d0 + 8) = o1;
    }
  }
  if (p.lse != nullptr && q < p.L && half == 0)
    p.lse[(static_cast<long long>(b) * p.H + h) * p.L + q] = (m_run + log2f(l_tot)) * 0.69314718055994531f;

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int DH>
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
  static bool attr_set = false;
  if (!attr_set) {
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_set = true;
  }
  AttnArgs p;
  p.out = reinterpret_cast<uint16_t*>(out);
  p.lse = lse;
  p.B = B; p.L = L; p.H = H;
  p.scale_log2 = scale * 1.44269504088896340736f;
  p.bf16 = bf16;
  p.q_tiles = (L + BQ - 1) / BQ;
  const int grid = B * H * p.q_tiles;
  COUNTR_CHECK_CUDA(launch_pdl(attention_fwd_kernel<DH>, dim3(grid), dim3(kThreads), SM::kTotal, stream, tq, tkv, p));
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_load, 4u * nblk * kBlkBytes);
    for (int i = 0; i < nblk; ++i) {
      tma_load_4d(smem + kOffQ + i * kBlkBytes, &tma_qkv, bar_load, 0, i * 128, h, b);
      tma_load_4d(smem + kOffK + i * kBlkBytes, &tma_qkv, bar_load, 0, i * 128, p.H + h, b);
      tma_load_4d(smem + kOffV + i * kBlkBytes, &tma_qkv, bar_load, 0, i * 128, 2 * p.H + h, b);
      tma_load_4d(smem + kOffdO + i * kBlkBytes, &tma_do, bar_load, 0, i * 128, h, b);
    }
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
      if (tid == 0) {
        if (pair == 0) mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint64_t dq = make_desc(smem_u32(smem + kOffQ + i * kBlkBytes), 16, 512, 4u);
        const uint64_t dk = make_desc(smem_u32(smem + kOffK + j * kBlkBytes), 16, 512, 4u);
        const uint64_t dd = make_desc(smem_u32(smem + kOffdO + i * kBlkBytes), 16, 512, 4u);
        const uint64_t dv = make_desc(smem_u32(smem + kOffV + j * kBlkBytes), 16, 512, 4u);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + tP, dd + 2 * k, dv + 2 * k, idesc_s, k != 0);
        umma_commit(bar_s);
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
      if (tid == 0) {
        tc_fence_after();
        const uint32_t aQ = smem_u32(smem + kOffQ + i * kBlkBytes), adO = smem_u32(smem + kOffdO + i * kBlkBytes);
        const uint32_t aK = smem_u32(smem + kOffK + j * kBlkBytes);
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

}  // namespace
}  // namespace countr

extern "C" int countr_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int dh, float scale,
                                    int bf16, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(qkv && out, "null pointer");
  COUNTR_REQUIRE(B > 0 && L > 0 && H > 0, "bad shape B=%d L=%d H=%d", B, L, H);
  if (dh == 64) return launch_attention<64>(qkv, out, lse, B, L, H, scale, bf16, stream);
  if (dh == 32) return launch_attention<32>(qkv, out, lse, B, L, H, scale, bf16, stream);
  return set_error(COUNTR_ERR_UNSUPPORTED, "attention head_dim %d not supported (32 or 64)", dh);
}

extern "C" int countr_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int L, int H,
                                    int dh, float scale, int bf16, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(qkv && out && dout && lse && dqkv, "null pointer");
  COUNTR_REQUIRE(dh == 32, "fused attention backward supports head_dim 32 (got %d)", dh);
  COUNTR_REQUIRE(L >= 1 && L <= 128 * kBwdMaxBlk, "fused attention backward supports L <= %d (got %d)", 128 * kBwdMaxBlk, L);
  const int D = H * dh;
  CUtensorMap tq, td;
  {
    const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)L, (uint64_t)(3 * H), (uint64_t)B};
    const uint64_t str[4] = {1, (uint64_t)(3 * D), (uint64_t)dh, (uint64_t)L * 3 * D};
    const uint32_t box[4] = {(uint32_t)dh, 128, 1, 1};
    int rc = make_tmap_4d_16b(&tq, qkv, dims, str, box, TMAP_SW_64);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {1, (uint64_t)D, (uint64_t)dh, (uint64_t)L * D};
    const uint32_t box[4] = {(uint32_t)dh, 128, 1, 1};
    int rc = make_tmap_4d_16b(&td, dout, dims, str, box, TMAP_SW_64);
    if (rc) return rc;
  }
  constexpr uint32_t smem_bytes = 4 * kBwdMaxBlk * (128 * 32 * 2) + 2 * (128 * 128 * 2) + 64 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  AttnBwdArgs a;
  a.dO = reinterpret_cast<const uint16_t*>(dout);
  a.O = reinterpret_cast<const uint16_t*>(out);
  a.lse = lse;
  a.dqkv = reinterpret_cast<uint16_t*>(dqkv);
  a.B = B; a.L = L; a.H = H;
  a.scale = scale;
  a.bf16 = bf16;
  COUNTR_CHECK_CUDA(launch_pdl(attention_bwd_kernel<32>, dim3(B * H), dim3(kBwdThreads), smem_bytes, stream, tq, td, a));
  return COUNTR_OK;
}
