// countr_b200 — backward-only streaming kernels of the fine-tune step (decoder side).
//
// The reference gets these from autograd (SURVEY.md §2.3 "Backward-only call sites"):
// upsample_bilinear2d_backward, native_group_norm_backward + threshold_backward,
// native_batch_norm_backward (InstanceNorm) + max_pool2d_with_indices_backward,
// _softmax_backward_data, bias-gradient sums, and the tiny-K/V cross-attention backward.
// All are HBM-bound: NHWC, 16-byte vector loads, per-block partial sums, one atomic per block.
#include "../../include/countr_b200.h"
#include "common.cuh"
#include "tma.h"

namespace countr {
namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  uint32_t r;
  if (bf16)
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack2(uint32_t v, int bf16) {
  if (bf16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], int bf16) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack2(w[i], bf16);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int bf16) {
  uint4 o;
  o.x = pack2(f[0], f[1], bf16);
  o.y = pack2(f[2], f[3], bf16);
  o.z = pack2(f[4], f[5], bf16);
  o.w = pack2(f[6], f[7], bf16);
  return o;
}
__device__ __forceinline__ float load_any(const void* p, long long i, int dtype) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  const uint16_t u = reinterpret_cast<const uint16_t*>(p)[i];
  if (dtype == 1) return __half2float(__ushort_as_half(u));
  return __uint_as_float(static_cast<uint32_t>(u) << 16);
}

// adjoint taps of bilinear x2 (align_corners=False): input index i receives from outputs
// 2i-1 (.25, i>=1), 2i (.75, or 1 at i==0), 2i+1 (.75, or 1 at i==n-1), 2i+2 (.25, i<=n-2)
__device__ __forceinline__ void up2_adjoint_taps(int i, int n, int (&o)[4], float (&w)[4]) {
  o[0] = 2 * i - 1; w[0] = i >= 1 ? 0.25f : 0.f;
  o[1] = 2 * i;     w[1] = i == 0 ? 1.f : 0.75f;
  o[2] = 2 * i + 1; w[2] = i == n - 1 ? 1.f : 0.75f;
  o[3] = 2 * i + 2; w[3] = i <= n - 2 ? 0.25f : 0.f;
  if (i < 1) o[0] = 0;
  if (i > n - 2) o[3] = 0;
}

// ------------------------------------------------------------------------------------------
// d(dmap)[B][H][W] fp32 = adjoint of the last F.interpolate(x2) applied to dOut [B][2H][2W]
// ------------------------------------------------------------------------------------------
__global__ void up2_bwd_kernel(const void* __restrict__ dy, int dtype, float* __restrict__ dx, int B, int H, int W) {
  const long long total = static_cast<long long>(B) * H * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ix = idx % W, iy = (idx / W) % H;
  const long long b = idx / (static_cast<long long>(W) * H);
  int ox[4], oy[4];
  float wx[4], wy[4];
  up2_adjoint_taps(ix, W, ox, wx);
  up2_adjoint_taps(iy, H, oy, wy);
  const long long base = b * 4ll * H * W;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (wy[a] == 0.f) continue;
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (wx[c] != 0.f) r += wx[c] * load_any(dy, base + static_cast<long long>(oy[a]) * (2 * W) + ox[c], dtype);
    acc += wy[a] * r;
  }
  dx[idx] = acc;
}

// ------------------------------------------------------------------------------------------
// GroupNorm + ReLU backward, pass A (reduce).  C == 256, 32-channel groups.
//   mode 0: dz = up2-adjoint gather of d_next [B][2H][2W][C] (gradient w.r.t. the next conv's input)
//   mode 1: dz = dmap[b][p] * w1[c]          (Conv2d 1x1 C->1 backward), also dw1/db1
// writes dyh = dz * 1[y > 0] (16-bit), accumulates dgamma/dbeta (fp32) and, per (image, group),
// S1 = sum dyh*gamma, S2 = sum dyh*gamma*xhat (double).
// ------------------------------------------------------------------------------------------
constexpr int kC = 256;
// Per channel only two running sums are needed:
//   MODE 0: R1 = sum dy, R2 = sum dy*xhat                     (dbeta = R1, dgamma = R2)
//   MODE 1: with r = dmap * 1[y>0]:  R1 = sum r, R2 = sum r*xhat
//           dbeta = w R1, dgamma = w R2, dw1 = gamma R2 + beta R1   (relu(y) = y where the mask is 1)
// and the group sums follow at the end as S1 = sum_c gamma_c dbeta_c, S2 = sum_c gamma_c dgamma_c.
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 1 ? 3 : 2) gn_relu_bwd_reduce_kernel(
    const uint16_t* __restrict__ raw, const double* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, const uint16_t* __restrict__ d_next, const float* __restrict__ dmap,
    const float* __restrict__ w1, uint16_t* __restrict__ dyh, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ dw1, float* __restrict__ db1, double* __restrict__ gsum, int H, int W, int G, float eps, int bf16,
    int R) {
  __shared__ float s_mean[8], s_rstd[8];
  __shared__ float red[8][kC + 8];
  const int b = blockIdx.y;
  const int HW = H * W;
  const int cpg = kC / G;  // 32
  if (threadIdx.x < G) {
    const double cnt = static_cast<double>(HW) * cpg;
    const double s = stats[(static_cast<long long>(b) * G + threadIdx.x) * 2];
    const double ss = stats[(static_cast<long long>(b) * G + threadIdx.x) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = rsqrtf(static_cast<float>(var) + eps);
  }
  __syncthreads();
  const int cv = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c0 = cv * 8, g = c0 / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float gam[8], bet[8], wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = gamma[c0 + j];
    bet[j] = beta[c0 + j];
    wv[j] = MODE == 1 ? w1[c0 + j] : 0.f;
  }
  float R1[8], R2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) R1[j] = R2[j] = 0.f;
  float a_b1 = 0.f;
  const uint16_t* rb = raw + static_cast<long long>(b) * HW * kC + c0;
  uint16_t* ob = dyh + static_cast<long long>(b) * HW * kC + c0;
  const int stride = gridDim.x * 8;
  if (MODE == 1) {
    // four pixels per iteration: four independent 16-byte loads in flight per thread (the loop is latency-bound)
    for (int pix = blockIdx.x * 8 + pl; pix < HW; pix += 4 * stride) {
      uint4 u[4];
      float dm[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pk = pix + k * stride;
        const bool has = pk < HW;
        u[k] = has ? *reinterpret_cast<const uint4*>(rb + static_cast<long long>(pk) * kC) : make_uint4(0, 0, 0, 0);
        dm[k] = has ? dmap[static_cast<long long>(b) * HW + pk] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pk = pix + k * stride;
        if (pk >= HW) break;
        float x[8], d[8];
        unpack8(u[k], x, bf16);
        if (cv == 0) a_b1 += dm[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (x[j] - mean) * rstd;
          const float r = fmaf(xh, gam[j], bet[j]) > 0.f ? dm[k] : 0.f;
          R1[j] += r;
          R2[j] = fmaf(r, xh, R2[j]);
          d[j] = r * wv[j];
        }
        *reinterpret_cast<uint4*>(ob + static_cast<long long>(pk) * kC) = pack8(d, bf16);
      }
    }
  } else {
    // Row-walking gather of the up-sample adjoint: the block owns 8 pixel columns x R rows (blockIdx.x = strip * x-tiles +
    // x-tile) and walks down the rows.  The horizontally blended hi-res rows 2iy-1 .. 2iy+2 live in registers and the last
    // two are re-used by the next row, so every step fetches 2 x 4 hi-res vectors instead of 4 x 4 and the rows a strip
    // shares stay in L1 (the pixel-per-iteration loop re-read each d_next element 4x through L2: 1.4 TB/s at 96^2).
    const int nxt = (W + 7) >> 3;
    const int xt = blockIdx.x % nxt, strip = blockIdx.x / nxt;
    const int ix = xt * 8 + pl;
    const int y0 = strip * R, y1 = min(H, y0 + R);
    if (ix < W) {
      int ox[4];
      float wx[4];
      up2_adjoint_taps(ix, W, ox, wx);
      const uint16_t* nb = d_next + static_cast<long long>(b) * 4 * HW * kC + c0;
      auto hrow = [&](int oy, float (&h)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = 0.f;
        const uint16_t* rp = nb + static_cast<long long>(oy) * (2 * W) * kC;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (wx[c] == 0.f) continue;
          float t[8];
          unpack8(*reinterpret_cast<const uint4*>(rp + static_cast<long long>(ox[c]) * kC), t, bf16);
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = fmaf(wx[c], t[j], h[j]);
        }
      };
      float h0[8], h1[8], h2[8], h3[8];
      hrow(max(2 * y0 - 1, 0), h0);
      hrow(2 * y0, h1);
      for (int iy = y0; iy < y1; ++iy) {
        hrow(2 * iy + 1, h2);
        hrow(min(2 * iy + 2, 2 * H - 1), h3);
        const float w0 = iy >= 1 ? 0.25f : 0.f, w1y = iy == 0 ? 1.f : 0.75f;
        const float w2 = iy == H - 1 ? 1.f : 0.75f, w3 = iy <= H - 2 ? 0.25f : 0.f;
        const int pix = iy * W + ix;
        float x[8], dy[8];
        unpack8(*reinterpret_cast<const uint4*>(rb + static_cast<long long>(pix) * kC), x, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = w0 * h0[j] + w1y * h1[j] + w2 * h2[j] + w3 * h3[j];
          const float xh = (x[j] - mean) * rstd;
          dy[j] = fmaf(xh, gam[j], bet[j]) > 0.f ? dz : 0.f;
          R1[j] += dy[j];
          R2[j] += dy[j] * xh;
          h0[j] = h2[j];
          h1[j] = h3[j];
        }
        *reinterpret_cast<uint4*>(ob + static_cast<long long>(pix) * kC) = pack8(dy, bf16);
      }
    }
  }
  float a_dg[8], a_db[8], a_dw[8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a_db[j] = MODE == 1 ? wv[j] * R1[j] : R1[j];
    a_dg[j] = MODE == 1 ? wv[j] * R2[j] : R2[j];
    a_dw[j] = gam[j] * R2[j] + bet[j] * R1[j];
    s1 += gam[j] * a_db[j];
    s2 += gam[j] * a_dg[j];
  }
  // group sums: 4 consecutive lanes share a group
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  // cross-pixel-lane reductions through smem, one quantity at a time
  for (int q = 0; q < 4; ++q) {
    if (q == 2 && MODE != 1) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl][c0 + j] = q == 0 ? a_dg[j] : q == 1 ? a_db[j] : q == 2 ? a_dw[j] : 0.f;
    if (q == 3) {
      red[pl][c0] = s1; red[pl][c0 + 1] = s2; red[pl][c0 + 2] = a_b1;
    }
    __syncthreads();
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    if (q == 0) atomicAdd(dgamma + threadIdx.x, v);
    else if (q == 1) atomicAdd(dbeta + threadIdx.x, v);
    else if (q == 2) atomicAdd(dw1 + threadIdx.x, v);
    else {
      // threadIdx.x = cv*8 + {0: s1, 1: s2, 2: db1}; one representative lane per group (cv % 4 == 0)
      const int cvv = threadIdx.x >> 3, which = threadIdx.x & 7;
      if ((cvv & 3) == 0 && which < 2) atomicAdd(gsum + (static_cast<long long>(b) * G + (cvv >> 2)) * 2 + which, static_cast<double>(v));
      if (MODE == 1 && cvv == 0 && which == 2) atomicAdd(db1, v);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm + ReLU backward, pass A for the up-sampled stages (mode 0 above), staged through shared memory.
// The register-walking kernel above is latency-bound (one dependent round trip to L2 / HBM per row step, 16 warps per SM:
// 1.9 TB/s at 96^2 -> 192^2).  Here a block owns 4 x 8 output pixels at a time; the 10 x 18 hi-res pixels they gather from
// arrive as ten row-contiguous bulk copies (cp.async.bulk, 9 KB each, mbarrier completion) into one of two stages, so the next
// tile's 92 KB are in flight while this one is reduced.  Blocks are persistent per image (the per-image group sums stay in
// registers) and run the channel / group reductions once at the end.  16 warps: warp = (row pair, column) of the tile — with
// 8 warps walking all four rows the kernel was bound by its own dependent ld.shared -> fma chains (2 warps per scheduler).
// ------------------------------------------------------------------------------------------
constexpr int GT_X = 8;
constexpr int GT_COLS = 2 * GT_X + 2;
constexpr uint32_t GT_PIX = kC * 2;                        // bytes of one pixel's channel vector
constexpr uint32_t GT_PITCH = GT_COLS * GT_PIX;
template <int GT_Y, int NS>
struct GatherCfg {
  static constexpr int kRows = 2 * GT_Y + 2;
  static constexpr uint32_t kStage = kRows * GT_PITCH;
  static constexpr uint32_t kSmem = NS * kStage + 64;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// d = a * w + c with a, w 16-bit and c, d fp32 in ONE instruction (FHFMA; PTX mixed-precision fma, sm_100): the conversion
// of every gathered value (HADD2.F32) was a fifth of the instructions this kernel issued, and it is issue-bound
template <bool kBf16>
__device__ __forceinline__ float fma_mixed(uint32_t a16, uint16_t w16, float c) {
  float d;
  const uint16_t a = static_cast<uint16_t>(a16);
  if (kBf16) asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(w16), "f"(c));
  else asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(w16), "f"(c));
  return d;
}

template <int GT_Y, int NS, bool kBf16>
__global__ void __launch_bounds__(512, 1) gn_relu_bwd_gather_kernel(
    const uint16_t* __restrict__ raw, const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
    const uint16_t* __restrict__ d_next, uint16_t* __restrict__ dyh, float* __restrict__ dgamma, float* __restrict__ dbeta,
    double* __restrict__ gsum, int H, int W, int G, float eps) {
  constexpr int bf16 = kBf16 ? 1 : 0;
  constexpr uint32_t GT_STAGE = GatherCfg<GT_Y, NS>::kStage;
  constexpr int RPW = GT_Y / 2;                      // rows per warp: warps 0-7 take the upper half of the tile, 8-15 the lower
  extern __shared__ __align__(128) uint8_t gt_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(gt_smem + NS * GT_STAGE);
  __shared__ float s_mean[8], s_rstd[8];
  const int b = blockIdx.y;
  const int HW = H * W;
  const int cpg = kC / G;
  if (threadIdx.x < G) {
    const double cnt = static_cast<double>(HW) * cpg;
    const double sm = stats[(static_cast<long long>(b) * G + threadIdx.x) * 2];
    const double ss = stats[(static_cast<long long>(b) * G + threadIdx.x) * 2 + 1];
    const double mean = sm / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = rsqrtf(static_cast<float>(var) + eps);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int cv = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c0 = cv * 8, g = c0 / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float gam[8], bet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = gamma[c0 + j];
    bet[j] = beta[c0 + j];
  }
  float R1[8], R2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) R1[j] = R2[j] = 0.f;
  const uint16_t* rb = raw + static_cast<long long>(b) * HW * kC + c0;
  uint16_t* ob = dyh + static_cast<long long>(b) * HW * kC + c0;
  const uint16_t* nb = d_next + static_cast<long long>(b) * 4 * HW * kC;
  const int ntx = (W + GT_X - 1) / GT_X, ntiles = ntx * ((H + GT_Y - 1) / GT_Y);

  // hi-res window of tile t: rows hy0..hy1, columns hx0..hx1 (clamped at the borders like the taps)
  auto window = [&](int t, int& y0, int& x0, int& hy0, int& hy1, int& hx0, int& hx1) {
    y0 = (t / ntx) * GT_Y;
    x0 = (t % ntx) * GT_X;
    const int yl = min(H, y0 + GT_Y) - 1, xl = min(W, x0 + GT_X) - 1;
    hy0 = max(2 * y0 - 1, 0), hy1 = min(2 * yl + 2, 2 * H - 1);
    hx0 = max(2 * x0 - 1, 0), hx1 = min(2 * xl + 2, 2 * W - 1);
  };
  auto issue = [&](int t, int stage) {      // one thread
    int y0, x0, hy0, hy1, hx0, hx1;
    window(t, y0, x0, hy0, hy1, hx0, hx1);
    const uint32_t row_bytes = static_cast<uint32_t>(hx1 - hx0 + 1) * GT_PIX;
    fence_proxy_async_smem();              // the stage was read with ld.shared two tiles ago
    mbar_arrive_expect_tx(bars + stage, row_bytes * static_cast<uint32_t>(hy1 - hy0 + 1));
    const uint32_t dst = smem_u32(gt_smem) + stage * GT_STAGE;
    for (int r = hy0; r <= hy1; ++r)
      bulk_g2s(dst + (r - hy0) * GT_PITCH, nb + (static_cast<long long>(r) * (2 * W) + hx0) * kC, row_bytes, bars + stage);
  };

  int it = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < NS - 1; ++k)
      if (static_cast<int>(blockIdx.x + k * gridDim.x) < ntiles) issue(blockIdx.x + k * gridDim.x, k);
  }
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int stage = it % NS;
    // NS - 1 tiles ahead, into the stage the previous iteration released
    if (threadIdx.x == 0 && t + (NS - 1) * static_cast<int>(gridDim.x) < ntiles) issue(t + (NS - 1) * gridDim.x, (it + NS - 1) % NS);
    int y0, x0, hy0, hy1, hx0, hx1;
    window(t, y0, x0, hy0, hy1, hx0, hx1);
    const int ix = x0 + (pl & 7);
    const int ys = y0 + RPW * (pl >> 3);               // this warp's rows
    const int y1 = min(H, y0 + GT_Y);
    const bool active = ix < W && ys < y1;
    // this thread's raw vectors do not depend on the staged tile: request them before waiting for it
    uint4 xr[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k)
      xr[k] = (active && ys + k < y1) ? *reinterpret_cast<const uint4*>(rb + static_cast<long long>((ys + k) * W + ix) * kC) : make_uint4(0, 0, 0, 0);
    mbar_wait(bars + stage, static_cast<uint32_t>(it / NS) & 1u);
    if (active) {
      int ox[4];
      float wx[4];
      up2_adjoint_taps(ix, W, ox, wx);
      // taps without weight (image border) read a valid neighbour and multiply it by zero: no branches in the row loop
      uint32_t off[4];
      uint16_t wh[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        off[c] = static_cast<uint32_t>((wx[c] != 0.f ? ox[c] : ox[1]) - hx0) * GT_PIX;
        wh[c] = static_cast<uint16_t>(pack2(wx[c], 0.f, bf16) & 0xffffu);         // 0, 0.25, 0.75, 1: exact in both formats
      }
      const uint8_t* st = gt_smem + stage * GT_STAGE + c0 * 2;
      auto hrow = [&](int oy, float (&h)[8]) {
        const uint8_t* rp = st + (oy - hy0) * GT_PITCH;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(rp + off[c]);
          const uint32_t w32[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[2 * i] = fma_mixed<kBf16>(w32[i] & 0xffffu, wh[c], c == 0 ? 0.f : h[2 * i]);
            h[2 * i + 1] = fma_mixed<kBf16>(w32[i] >> 16, wh[c], c == 0 ? 0.f : h[2 * i + 1]);
          }
        }
      };
      float h0[8], h1[8], h2[8], h3[8];
      hrow(max(2 * ys - 1, 0), h0);
      hrow(2 * ys, h1);
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int iy = ys + k;
        if (iy < y1) {
          hrow(2 * iy + 1, h2);
          hrow(min(2 * iy + 2, 2 * H - 1), h3);
          const float w0 = iy >= 1 ? 0.25f : 0.f, w1y = iy == 0 ? 1.f : 0.75f;
          const float w2 = iy == H - 1 ? 1.f : 0.75f, w3 = iy <= H - 2 ? 0.25f : 0.f;
          float x[8], dy[8];
          unpack8(xr[k], x, bf16);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float dz = w0 * h0[j] + w1y * h1[j] + w2 * h2[j] + w3 * h3[j];
            const float xh = (x[j] - mean) * rstd;
            dy[j] = fmaf(xh, gam[j], bet[j]) > 0.f ? dz : 0.f;
            R1[j] += dy[j];
            R2[j] = fmaf(dy[j], xh, R2[j]);
            h0[j] = h2[j];
            h1[j] = h3[j];
          }
          *reinterpret_cast<uint4*>(ob + static_cast<long long>(iy * W + ix) * kC) = pack8(dy, bf16);
        }
      }
    }
    __syncthreads();          // everyone is done with this stage before the next iteration refills it
  }

  // channel sums over the 16 pixel lanes and the per-(image, group) sums, once per block; the stages are free now
  float(*red)[kC + 8] = reinterpret_cast<float(*)[kC + 8]>(gt_smem);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1 += gam[j] * R1[j];
    s2 += gam[j] * R2[j];
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl][c0 + j] = q == 0 ? R2[j] : q == 1 ? R1[j] : 0.f;
    if (q == 2) {
      red[pl][c0] = s1; red[pl][c0 + 1] = s2;
    }
    __syncthreads();
    if (threadIdx.x < kC) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) v += red[k][threadIdx.x];
      if (q == 0) atomicAdd(dgamma + threadIdx.x, v);
      else if (q == 1) atomicAdd(dbeta + threadIdx.x, v);
      else {
        const int cvv = threadIdx.x >> 3, which = threadIdx.x & 7;
        if ((cvv & 3) == 0 && which < 2) atomicAdd(gsum + (static_cast<long long>(b) * G + (cvv >> 2)) * 2 + which, static_cast<double>(v));
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm + ReLU backward, pass A under the 1x1-conv head (mode 1 of gn_relu_bwd_reduce_kernel), staged like the gather
// kernel: the pixels of an image are one contiguous stream, a tile = 64 pixels = one 32 KB bulk copy, four stages per
// persistent block (three tiles in flight), 16 warps x 4 pixels per tile.  The direct-load kernel needs ~100 registers for its
// four loads in flight and runs at 24 % occupancy: 3.8 TB/s; this one keeps 96 KB per SM in flight whatever the register count.
// ------------------------------------------------------------------------------------------
constexpr int HT_PX = 64, HT_NS = 4;
constexpr uint32_t HT_STAGE = HT_PX * kC * 2;
constexpr uint32_t HT_SMEM = HT_NS * HT_STAGE + 64;

template <bool kBf16>
__global__ void __launch_bounds__(512, 1) gn_head_reduce_kernel(
    const uint16_t* __restrict__ raw, const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ dmap, const float* __restrict__ w1, uint16_t* __restrict__ dyh, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dw1, float* __restrict__ db1, double* __restrict__ gsum, int HW, int G, float eps) {
  constexpr int bf16 = kBf16 ? 1 : 0;
  extern __shared__ __align__(128) uint8_t ht_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(ht_smem + HT_NS * HT_STAGE);
  __shared__ float s_mean[8], s_rstd[8];
  const int b = blockIdx.y;
  const int cpg = kC / G;
  if (threadIdx.x < G) {
    const double cnt = static_cast<double>(HW) * cpg;
    const double sm = stats[(static_cast<long long>(b) * G + threadIdx.x) * 2];
    const double ss = stats[(static_cast<long long>(b) * G + threadIdx.x) * 2 + 1];
    const double mean = sm / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = rsqrtf(static_cast<float>(var) + eps);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < HT_NS; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int cv = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c0 = cv * 8, g = c0 / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float gam[8], bet[8], wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = gamma[c0 + j];
    bet[j] = beta[c0 + j];
    wv[j] = w1[c0 + j];
  }
  float R1[8], R2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) R1[j] = R2[j] = 0.f;
  float a_b1 = 0.f;
  const uint16_t* rb = raw + static_cast<long long>(b) * HW * kC;
  uint16_t* ob = dyh + static_cast<long long>(b) * HW * kC + c0;
  const float* dmb = dmap + static_cast<long long>(b) * HW;
  const int ntiles = (HW + HT_PX - 1) / HT_PX;
  const uint16_t one16 = static_cast<uint16_t>(pack2(1.f, 0.f, bf16) & 0xffffu);

  auto issue = [&](int t, int stage) {      // one thread
    const int npx = min(HT_PX, HW - t * HT_PX);
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bars + stage, static_cast<uint32_t>(npx) * kC * 2);
    const uint32_t dst = smem_u32(ht_smem) + stage * HT_STAGE;
    const uint16_t* src = rb + static_cast<long long>(t) * HT_PX * kC;
    const uint32_t half = static_cast<uint32_t>(npx / 2) * kC * 2, rest = static_cast<uint32_t>(npx) * kC * 2 - half;
    if (half) bulk_g2s(dst, src, half, bars + stage);
    bulk_g2s(dst + half, reinterpret_cast<const uint8_t*>(src) + half, rest, bars + stage);
  };

  int it = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < HT_NS - 1; ++k)
      if (static_cast<int>(blockIdx.x + k * gridDim.x) < ntiles) issue(blockIdx.x + k * gridDim.x, k);
  }
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int stage = it % HT_NS;
    if (threadIdx.x == 0 && t + (HT_NS - 1) * static_cast<int>(gridDim.x) < ntiles) issue(t + (HT_NS - 1) * gridDim.x, (it + HT_NS - 1) % HT_NS);
    const int p0 = t * HT_PX;
    float dm[HT_PX / 16];
#pragma unroll
    for (int k = 0; k < HT_PX / 16; ++k) {
      const int pk = p0 + pl + 16 * k;
      dm[k] = pk < HW ? dmb[pk] : 0.f;
    }
    mbar_wait(bars + stage, static_cast<uint32_t>(it / HT_NS) & 1u);
    const uint8_t* st = ht_smem + stage * HT_STAGE + c0 * 2;
#pragma unroll
    for (int k = 0; k < HT_PX / 16; ++k) {
      const int pk = p0 + pl + 16 * k;
      if (pk < HW) {
        const uint4 u = *reinterpret_cast<const uint4*>(st + (pl + 16 * k) * (kC * 2));
        const uint32_t w32[4] = {u.x, u.y, u.z, u.w};
        float d[8];
        if (cv == 0) a_b1 += dm[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t h16 = (j & 1) ? (w32[j >> 1] >> 16) : (w32[j >> 1] & 0xffffu);
          const float xh = fma_mixed<kBf16>(h16, one16, -mean) * rstd;           // (x - mean) * rstd, the subtraction exact
          const float r = fmaf(xh, gam[j], bet[j]) > 0.f ? dm[k] : 0.f;
          R1[j] += r;
          R2[j] = fmaf(r, xh, R2[j]);
          d[j] = r * wv[j];
        }
        *reinterpret_cast<uint4*>(ob + static_cast<long long>(pk) * kC) = pack8(d, bf16);
      }
    }
    __syncthreads();
  }

  float(*red)[kC + 8] = reinterpret_cast<float(*)[kC + 8]>(ht_smem);
  float a_dg[8], a_db[8], a_dw[8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a_db[j] = wv[j] * R1[j];
    a_dg[j] = wv[j] * R2[j];
    a_dw[j] = gam[j] * R2[j] + bet[j] * R1[j];
    s1 += gam[j] * a_db[j];
    s2 += gam[j] * a_dg[j];
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl][c0 + j] = q == 0 ? a_dg[j] : q == 1 ? a_db[j] : q == 2 ? a_dw[j] : 0.f;
    if (q == 3) {
      red[pl][c0] = s1; red[pl][c0 + 1] = s2; red[pl][c0 + 2] = a_b1;
    }
    __syncthreads();
    if (threadIdx.x < kC) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) v += red[k][threadIdx.x];
      if (q == 0) atomicAdd(dgamma + threadIdx.x, v);
      else if (q == 1) atomicAdd(dbeta + threadIdx.x, v);
      else if (q == 2) atomicAdd(dw1 + threadIdx.x, v);
      else {
        const int cvv = threadIdx.x >> 3, which = threadIdx.x & 7;
        if ((cvv & 3) == 0 && which < 2) atomicAdd(gsum + (static_cast<long long>(b) * G + (cvv >> 2)) * 2 + which, static_cast<double>(v));
        if (cvv == 0 && which == 2) atomicAdd(db1, v);
      }
    }
    __syncthreads();
  }
}

// pass B: d_raw = rstd * (dyh*gamma - S1/n - xhat*S2/n);  dbias[c] += sum d_raw
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const uint16_t* __restrict__ raw, const uint16_t* __restrict__ dyh,
                                                            const double* __restrict__ stats, const double* __restrict__ gsum,
                                                            const float* __restrict__ gamma, uint16_t* __restrict__ d_raw,
                                                            float* __restrict__ dbias, int HW, int G, float eps, int bf16) {
  __shared__ float s_mean[8], s_rstd[8], s_m1[8], s_m2[8];
  __shared__ float red[8][kC + 8];
  const int b = blockIdx.y;
  const int cpg = kC / G;
  if (threadIdx.x < G) {
    const double cnt = static_cast<double>(HW) * cpg;
    const long long o = (static_cast<long long>(b) * G + threadIdx.x) * 2;
    const double mean = stats[o] / cnt;
    double var = stats[o + 1] / cnt - mean * mean;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = rsqrtf(static_cast<float>(var) + eps);
    s_m1[threadIdx.x] = static_cast<float>(gsum[o] / cnt);
    s_m2[threadIdx.x] = static_cast<float>(gsum[o + 1] / cnt);
  }
  __syncthreads();
  const int cv = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c0 = cv * 8, g = c0 / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g], m1 = s_m1[g], m2 = s_m2[g];
  float gam[8], a_b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = gamma[c0 + j];
    a_b[j] = 0.f;
  }
  const long long base = static_cast<long long>(b) * HW * kC + c0;
  for (int pix = blockIdx.x * 8 + pl; pix < HW; pix += gridDim.x * 8) {
    float x[8], dy[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(raw + base + static_cast<long long>(pix) * kC), x, bf16);
    unpack8(*reinterpret_cast<const uint4*>(dyh + base + static_cast<long long>(pix) * kC), dy, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - mean) * rstd;
      o[j] = rstd * (dy[j] * gam[j] - m1 - xh * m2);
      a_b[j] += o[j];
    }
    *reinterpret_cast<uint4*>(d_raw + base + static_cast<long long>(pix) * kC) = pack8(o, bf16);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[pl][c0 + j] = a_b[j];
  __syncthreads();
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
  atomicAdd(dbias + threadIdx.x, v);
}

// ------------------------------------------------------------------------------------------
// out[n] += sum_r x[r][n] * scale   (bias gradients).  x is 16-bit or fp32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x, int dtype, float* __restrict__ out, long long R,
                                                      int N, long long ld, int rows_per_block) {
  // thread = 2 columns; blockDim.x = 128 threads along N, blockDim.y = 2 row lanes
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  pdl_trigger();
  pdl_wait();
  if (c >= N) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    if (dtype == 0) {
      const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(x) + r * ld + c);
      a0 += v.x; a1 += v.y;
    } else {
      const float2 v = unpack2(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(x) + r * ld + c), dtype == 2);
      a0 += v.x; a1 += v.y;
    }
  }
  __shared__ float red[2][256];
  red[threadIdx.y][threadIdx.x * 2] = a0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.y == 0) {
    atomicAdd(out + c, red[0][threadIdx.x * 2] + red[1][threadIdx.x * 2]);
    if (c + 1 < N) atomicAdd(out + c + 1, red[0][threadIdx.x * 2 + 1] + red[1][threadIdx.x * 2 + 1]);
  }
}

// ------------------------------------------------------------------------------------------
// softmax backward over materialised score rows (self-attention backward, unfused):
//   P = exp(S - lse[row]);  dS = scale * P * (dP - sum_j P_j dP_j)   in place: S <- P, dP <- dS
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(uint16_t* __restrict__ s_io, uint16_t* __restrict__ dp_io,
                                                                const float* __restrict__ lse, long long rows, int L,
                                                                float scale, int bf16) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  uint32_t* sp = reinterpret_cast<uint32_t*>(s_io + row * L);
  uint32_t* dp = reinterpret_cast<uint32_t*>(dp_io + row * L);
  const float l = lse[row];
  const int n2 = L / 2;
  float dot = 0.f;
  for (int i = lane; i < n2; i += 32) {
    const float2 s = unpack2(sp[i], bf16), d = unpack2(dp[i], bf16);
    const float p0 = __expf(s.x - l), p1 = __expf(s.y - l);
    dot += p0 * d.x + p1 * d.y;
  }
  dot = warp_sum(dot);
  for (int i = lane; i < n2; i += 32) {
    const float2 s = unpack2(sp[i], bf16), d = unpack2(dp[i], bf16);
    const float p0 = __expf(s.x - l), p1 = __expf(s.y - l);
    sp[i] = pack2(p0, p1, bf16);
    dp[i] = pack2(scale * p0 * (d.x - dot), scale * p1 * (d.y - dot), bf16);
  }
}

// ------------------------------------------------------------------------------------------
// cross-attention core backward (S <= 8 exemplar tokens, dh == 32, D % 512 == 0).
//   phase 1 (warp per token): dq16; ds (scaled) and p kept in smem
//   phase 2 (thread per 2 channels): dk/dv of this block's tokens -> atomics into dk32/dv32 [B][S][D]
// ------------------------------------------------------------------------------------------
constexpr int kMaxShots = 8;
constexpr int kTokPerBlock = 32;
__global__ void __launch_bounds__(256) cross_attn_core_bwd_kernel(
    const uint16_t* __restrict__ q16, const float* __restrict__ k32, const float* __restrict__ v32,
    const float* __restrict__ probs, const uint16_t* __restrict__ do16, uint16_t* __restrict__ dq16,
    float* __restrict__ dk32, float* __restrict__ dv32, int L, int S, int D, float scale, long long kv_bstride, int bf16) {
  extern __shared__ float sm[];
  const int Hh = D / 32;
  float* sk = sm;                       // [S][D]
  float* sv = sk + S * D;               // [S][D]
  float* sds = sv + S * D;              // [tokens][Hh][S]  (scaled ds)
  float* spr = sds + kTokPerBlock * Hh * S;  // [tokens][Hh][S]
  const int tok0 = blockIdx.x * kTokPerBlock;
  const int b = tok0 / L;
  // channel-permuted K/V (ch = lane*16 + j -> j*32 + lane inside each 512-channel chunk): conflict-free warp reads,
  // see cross_attn_core_kernel in elementwise.cu
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int s = i / D, ch = i - s * D;
    const int pi = s * D + (ch & ~511) + ((ch & 15) << 5) + ((ch & 511) >> 4);
    sk[pi] = k32[b * kv_bstride + i];
    sv[pi] = v32[b * kv_bstride + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < kTokPerBlock; t += 8) {
    const long long tok = tok0 + t;
    for (int c0 = 0; c0 < D; c0 += 512) {
      const int ch = c0 + lane * 16;
      const int hh = ch >> 5;
      float dO[16];
      {
        float a[8], c[8];
        unpack8(*reinterpret_cast<const uint4*>(do16 + tok * D + ch), a, bf16);
        unpack8(*reinterpret_cast<const uint4*>(do16 + tok * D + ch + 8), c, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) { dO[j] = a[j]; dO[8 + j] = c[j]; }
      }
      float pr[kMaxShots], dpv[kMaxShots];
      float dot = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxShots; ++s) {
        if (s < S) {
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) d += dO[j] * sv[s * D + c0 + j * 32 + lane];
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          pr[s] = probs[(tok * Hh + hh) * S + s];
          dpv[s] = d;
          dot += pr[s] * d;
        }
      }
      float dq[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) dq[j] = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxShots; ++s) {
        if (s < S) {
          const float ds = scale * pr[s] * (dpv[s] - dot);
          if ((lane & 1) == 0) {
            sds[(t * Hh + hh) * S + s] = ds;
            spr[(t * Hh + hh) * S + s] = pr[s];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) dq[j] += ds * sk[s * D + c0 + j * 32 + lane];
        }
      }
      float a[8], c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = dq[j]; c[j] = dq[8 + j]; }
      *reinterpret_cast<uint4*>(dq16 + tok * D + ch) = pack8(a, bf16);
      *reinterpret_cast<uint4*>(dq16 + tok * D + ch + 8) = pack8(c, bf16);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x * 2; c < D; c += blockDim.x * 2) {
    const int hh = c >> 5;
    float adk[kMaxShots][2], adv[kMaxShots][2];
#pragma unroll
    for (int s = 0; s < kMaxShots; ++s) adk[s][0] = adk[s][1] = adv[s][0] = adv[s][1] = 0.f;
#pragma unroll 8
    for (int t = 0; t < kTokPerBlock; ++t) {
      const long long tok = tok0 + t;
      const float2 q = unpack2(*reinterpret_cast<const uint32_t*>(q16 + tok * D + c), bf16);
      const float2 d = unpack2(*reinterpret_cast<const uint32_t*>(do16 + tok * D + c), bf16);
#pragma unroll
      for (int s = 0; s < kMaxShots; ++s) {
        if (s < S) {
          const float ds = sds[(t * Hh + hh) * S + s], p = spr[(t * Hh + hh) * S + s];
          adk[s][0] += ds * q.x; adk[s][1] += ds * q.y;
          adv[s][0] += p * d.x;  adv[s][1] += p * d.y;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < kMaxShots; ++s) {
      if (s < S) {
        const long long o = b * kv_bstride + static_cast<long long>(s) * D + c;
        atomicAdd(dk32 + o, adk[s][0]); atomicAdd(dk32 + o + 1, adk[s][1]);
        atomicAdd(dv32 + o, adv[s][0]); atomicAdd(dv32 + o + 1, adv[s][1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm + ReLU + {MaxPool2d(2) | global average} backward; one CTA per (sample, 64 channels)
//   mode 0: dpool16 [N][H/2][W/2][C]; mode 1: dpool32 [N][C]     -> d_raw16 [N][H][W][C]
// also accumulates the conv bias gradient dbias[c] += sum d_raw.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_relu_pool_bwd_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ mean_in,
                                                                   const float* __restrict__ rstd_in,
                                                                   const uint16_t* __restrict__ dpool16,
                                                                   const float* __restrict__ dpool32, uint16_t* __restrict__ d_raw,
                                                                   float* __restrict__ dbias, int H, int W, int C, int mode,
                                                                   int bf16) {
  __shared__ float red[2][8][64];
  __shared__ float s_a[64], s_b[64];
  const int n = blockIdx.y, c0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = H * W;
  const int ca = c0 + lane * 2;
  const float ma = mean_in[static_cast<long long>(n) * C + ca], mb = mean_in[static_cast<long long>(n) * C + ca + 1];
  const float ra = rstd_in[static_cast<long long>(n) * C + ca], rb = rstd_in[static_cast<long long>(n) * C + ca + 1];
  const uint16_t* xb = raw + static_cast<long long>(n) * HW * C + ca;
  uint16_t* ob = d_raw + static_cast<long long>(n) * HW * C + ca;
  const int OW = W / 2, OH = H / 2;
  float ga = 0.f, gb = 0.f;
  if (mode == 1) {
    ga = dpool32[static_cast<long long>(n) * C + ca] / HW;
    gb = dpool32[static_cast<long long>(n) * C + ca + 1] / HW;
  }
  // pass 1: A = sum dxhat, Bq = sum dxhat * xhat
  float A0 = 0.f, A1 = 0.f, B0 = 0.f, B1 = 0.f;
  if (mode == 0) {
    for (int p = warp; p < OH * OW; p += 8) {
      const int oy = p / OW, ox = p % OW;
      const uint16_t* q = xb + (static_cast<long long>(2 * oy) * W + 2 * ox) * C;
      const float2 f0 = unpack2(*reinterpret_cast<const uint32_t*>(q), bf16);
      const float2 f1 = unpack2(*reinterpret_cast<const uint32_t*>(q + C), bf16);
      const float2 f2 = unpack2(*reinterpret_cast<const uint32_t*>(q + static_cast<long long>(W) * C), bf16);
      const float2 f3 = unpack2(*reinterpret_cast<const uint32_t*>(q + static_cast<long long>(W) * C + C), bf16);
      const float va = fmaxf(fmaxf(f0.x, f1.x), fmaxf(f2.x, f3.x)), vb = fmaxf(fmaxf(f0.y, f1.y), fmaxf(f2.y, f3.y));
      const float2 d = unpack2(*reinterpret_cast<const uint32_t*>(dpool16 + (static_cast<long long>(n) * OH * OW + p) * C + ca), bf16);
      const float xa = (va - ma) * ra, xbb = (vb - mb) * rb;
      if (xa > 0.f) { A0 += d.x; B0 += d.x * xa; }
      if (xbb > 0.f) { A1 += d.y; B1 += d.y * xbb; }
    }
  } else {
    for (int p = warp; p < HW; p += 8) {
      const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(xb + static_cast<long long>(p) * C), bf16);
      const float xa = (f.x - ma) * ra, xbb = (f.y - mb) * rb;
      if (xa > 0.f) { A0 += ga; B0 += ga * xa; }
      if (xbb > 0.f) { A1 += gb; B1 += gb * xbb; }
    }
  }
  red[0][warp][lane * 2] = A0; red[0][warp][lane * 2 + 1] = A1;
  red[1][warp][lane * 2] = B0; red[1][warp][lane * 2 + 1] = B1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[0][k][threadIdx.x]; q += red[1][k][threadIdx.x]; }
    s_a[threadIdx.x] = a / HW;
    s_b[threadIdx.x] = q / HW;
  }
  __syncthreads();
  const float mA0 = s_a[lane * 2], mA1 = s_a[lane * 2 + 1], mB0 = s_b[lane * 2], mB1 = s_b[lane * 2 + 1];
  // pass 2: d_raw = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat))
  float sb0 = 0.f, sb1 = 0.f;
  if (mode == 0) {
    for (int p = warp; p < OH * OW; p += 8) {
      const int oy = p / OW, ox = p % OW;
      const long long o00 = (static_cast<long long>(2 * oy) * W + 2 * ox) * C;
      const long long offs[4] = {o00, o00 + C, o00 + static_cast<long long>(W) * C, o00 + static_cast<long long>(W) * C + C};
      float2 f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = unpack2(*reinterpret_cast<const uint32_t*>(xb + offs[k]), bf16);
      // arg-max in window scan order; first maximum wins (max_pool2d semantics)
      int ia = 0, ib = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        if (f[k].x > f[ia].x) ia = k;
        if (f[k].y > f[ib].y) ib = k;
      }
      const float2 d = unpack2(*reinterpret_cast<const uint32_t*>(dpool16 + (static_cast<long long>(n) * OH * OW + p) * C + ca), bf16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xa = (f[k].x - ma) * ra, xbb = (f[k].y - mb) * rb;
        const float da = (k == ia && xa > 0.f) ? d.x : 0.f, db = (k == ib && xbb > 0.f) ? d.y : 0.f;
        const float oa = ra * (da - mA0 - xa * mB0), obv = rb * (db - mA1 - xbb * mB1);
        sb0 += oa; sb1 += obv;
        *reinterpret_cast<uint32_t*>(ob + offs[k]) = pack2(oa, obv, bf16);
      }
    }
  } else {
    for (int p = warp; p < HW; p += 8) {
      const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(xb + static_cast<long long>(p) * C), bf16);
      const float xa = (f.x - ma) * ra, xbb = (f.y - mb) * rb;
      const float oa = ra * ((xa > 0.f ? ga : 0.f) - mA0 - xa * mB0), obv = rb * ((xbb > 0.f ? gb : 0.f) - mA1 - xbb * mB1);
      sb0 += oa; sb1 += obv;
      *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(p) * C) = pack2(oa, obv, bf16);
    }
  }
  if (dbias != nullptr) {
    __syncthreads();
    red[0][warp][lane * 2] = sb0; red[0][warp][lane * 2 + 1] = sb1;
    __syncthreads();
    if (threadIdx.x < 64) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[0][k][threadIdx.x];
      atomicAdd(dbias + c0 + threadIdx.x, a);
    }
  }
}

// ------------------------------------------------------------------------------------------
// pixel-parallel variant of the InstanceNorm+ReLU+MaxPool backward (mode 0) for large maps:
// pass 1 accumulates A = sum dxhat, Bq = sum dxhat*xhat per (sample, channel) with atomics into
// `ab` [N][C][2]; pass 2 writes d_raw.  grid = (C/64, N, pixel splits).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_pool_bwd_split_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ mean_in,
                                                                    const float* __restrict__ rstd_in,
                                                                    const uint16_t* __restrict__ dpool16, float* __restrict__ ab,
                                                                    uint16_t* __restrict__ d_raw, float* __restrict__ dbias, int H,
                                                                    int W, int C, int ppb, int pass, int bf16) {
  __shared__ float red[2][8][64];
  const int n = blockIdx.y, c0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = H * W, OW = W / 2, OH = H / 2;
  const int ca = c0 + lane * 2;
  const long long nc = static_cast<long long>(n) * C + ca;
  const float ma = mean_in[nc], mb = mean_in[nc + 1], ra = rstd_in[nc], rb = rstd_in[nc + 1];
  const uint16_t* xb = raw + static_cast<long long>(n) * HW * C + ca;
  uint16_t* ob = d_raw + static_cast<long long>(n) * HW * C + ca;
  const int p_end = min(OH * OW, (static_cast<int>(blockIdx.z) + 1) * ppb);
  float mA0 = 0.f, mA1 = 0.f, mB0 = 0.f, mB1 = 0.f;
  if (pass == 1) {
    mA0 = ab[nc * 2] / HW; mB0 = ab[nc * 2 + 1] / HW;
    mA1 = ab[nc * 2 + 2] / HW; mB1 = ab[nc * 2 + 3] / HW;
  }
  float A0 = 0.f, A1 = 0.f, B0 = 0.f, B1 = 0.f;   // pass 0: sums; pass 1: A0/A1 = bias-gradient partials
  for (int p = blockIdx.z * ppb + warp; p < p_end; p += 8) {
    const int oy = p / OW, ox = p % OW;
    const long long o00 = (static_cast<long long>(2 * oy) * W + 2 * ox) * C;
    const long long offs[4] = {o00, o00 + C, o00 + static_cast<long long>(W) * C, o00 + static_cast<long long>(W) * C + C};
    float2 f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = unpack2(*reinterpret_cast<const uint32_t*>(xb + offs[k]), bf16);
    int ia = 0, ib = 0;   // arg-max in window scan order; first maximum wins (max_pool2d semantics)
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      if (f[k].x > f[ia].x) ia = k;
      if (f[k].y > f[ib].y) ib = k;
    }
    const float2 d = unpack2(*reinterpret_cast<const uint32_t*>(dpool16 + (static_cast<long long>(n) * OH * OW + p) * C + ca), bf16);
    if (pass == 0) {
      float va = f[0].x, vb = f[0].y;
#pragma unroll
      for (int k = 1; k < 4; ++k) { va = k == ia ? f[k].x : va; vb = k == ib ? f[k].y : vb; }
      const float xa = (va - ma) * ra, xbb = (vb - mb) * rb;
      if (xa > 0.f) { A0 += d.x; B0 += d.x * xa; }
      if (xbb > 0.f) { A1 += d.y; B1 += d.y * xbb; }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xa = (f[k].x - ma) * ra, xbb = (f[k].y - mb) * rb;
        const float da = (k == ia && xa > 0.f) ? d.x : 0.f, db = (k == ib && xbb > 0.f) ? d.y : 0.f;
        const float oa = ra * (da - mA0 - xa * mB0), obv = rb * (db - mA1 - xbb * mB1);
        A0 += oa; A1 += obv;
        *reinterpret_cast<uint32_t*>(ob + offs[k]) = pack2(oa, obv, bf16);
      }
    }
  }
  if (pass == 1 && dbias == nullptr) return;
  red[0][warp][lane * 2] = A0; red[0][warp][lane * 2 + 1] = A1;
  red[1][warp][lane * 2] = B0; red[1][warp][lane * 2 + 1] = B1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[0][k][threadIdx.x]; q += red[1][k][threadIdx.x]; }
    if (pass == 0) {
      atomicAdd(ab + (static_cast<long long>(n) * C + c0 + threadIdx.x) * 2, a);
      atomicAdd(ab + (static_cast<long long>(n) * C + c0 + threadIdx.x) * 2 + 1, q);
    } else {
      atomicAdd(dbias + c0 + threadIdx.x, a);
    }
  }
}

// ------------------------------------------------------------------------------------------
// decoder_proj1[0] weight gradient: dW[co][ci][ky][kx] += sum_{n,y,x} d_raw[n,y,x,co] * box[n,ci,y+ky-1,x+kx-1]
// (Cin = 3 -> K = 27: a direct smem-tiled reduction, not a tensor-core shape)
// block = 512 pixels of one sample in 4 tiles of 128 (fewer tiles per block means more atomics: slower); smem holds the tile transposed
// ([tap][pixel] fp32, [co][pixel] fp16) so each thread (co = tid % 64, taps tid/64 + 4j) reads 4 pixels per LDS.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exemplar_conv1_dw_kernel(const void* __restrict__ boxes, int dtype, long long sB, long long sK,
                                                                 long long sC, long long sH, long long sW,
                                                                 const uint16_t* __restrict__ d_raw, float* __restrict__ dw, int S,
                                                                 int HW, int bf16, int SUB) {
  constexpr int PX = 128;
  __shared__ __align__(16) float s_in[27][PX];
  __shared__ __align__(16) uint16_t s_d[64][PX + 4];
  __shared__ float s_box[3][4][66];              // the PX / W (+2 halo) input rows of the tile, zero-padded columns -1 and W
  const int n = blockIdx.y;
  const int b = n / S, s = n % S;
  const int H = HW, W = HW;
  const int rows = PX / W;                       // image rows per tile (the host checks W == 64: two rows)
  const int co = threadIdx.x & 63, kq = threadIdx.x >> 6;
  float acc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j] = 0.f;
  for (int sub = 0; sub < SUB; ++sub) {
    const int p0 = (blockIdx.x * SUB + sub) * PX;
    const int y0 = p0 / W;
    __syncthreads();
    // both fills are independent global loads (one round trip): the halo rows of the box and the 16-byte d_raw vectors
    for (int i = threadIdx.x; i < 3 * (rows + 2) * 66; i += blockDim.x) {
      const int xx = i % 66 - 1, r = (i / 66) % (rows + 2), ci = i / (66 * (rows + 2));
      const int yy = y0 + r - 1;
      float v = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = load_any(boxes, b * sB + s * sK + ci * sC + yy * sH + xx * sW, dtype);
      s_box[ci][r][xx + 1] = v;
    }
#pragma unroll
    for (int i = threadIdx.x; i < PX * 8; i += 256) {
      const int px = i >> 3, c8 = (i & 7) * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(d_raw + (static_cast<long long>(n) * H * W + p0 + px) * 64 + c8);
      const uint32_t w32[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s_d[c8 + 2 * j][px] = static_cast<uint16_t>(w32[j] & 0xffffu);
        s_d[c8 + 2 * j + 1][px] = static_cast<uint16_t>(w32[j] >> 16);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PX * 27; i += blockDim.x) {
      const int px = i % PX, k = i / PX;
      const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
      s_in[k][px] = s_box[ci][px / W + ky][px % W + kx];
    }
    __syncthreads();
    for (int px = 0; px < PX; px += 4) {
      const uint2 dv = *reinterpret_cast<const uint2*>(&s_d[co][px]);
      const float2 d01 = unpack2(dv.x, bf16), d23 = unpack2(dv.y, bf16);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int k = kq + 4 * j;
        if (k < 27) {
          const float4 iv = *reinterpret_cast<const float4*>(&s_in[k][px]);
          acc[j] += d01.x * iv.x + d01.y * iv.y + d23.x * iv.z + d23.y * iv.w;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int k = kq + 4 * j;
    if (k < 27) atomicAdd(dw + co * 27 + k, acc[j]);
  }
}

// dw_packed [Cout][9][Cin] fp32 -> grad [Cout][Cin][3][3] fp32 (overwrite)
__global__ void conv_dw_unpack_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin) {
  const long long n = static_cast<long long>(Cout) * Cin * 9;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long t = i;
  const int tap = t % 9; t /= 9;
  const int ci = t % Cin; t /= Cin;
  const int co = static_cast<int>(t);
  dst[i] = src[(static_cast<long long>(co) * 9 + tap) * Cin + ci];
}

}  // namespace
}  // namespace countr

using namespace countr;

extern "C" int countr_upsample2x_bwd(const void* dy, int dtype, float* dx, int B, int H, int W, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(dy && dx && dtype >= 0 && dtype <= 2, "bad arguments");
  const long long total = static_cast<long long>(B) * H * W;
  up2_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(dy, dtype, dx, B, H, W);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

static int gn_grid(int HW, int B, dim3* grid) {
  long long blocks = (HW + 7) / 8;
  const long long cap = 148ll * 8 * 2;
  if (blocks * B > cap) blocks = (cap + B - 1) / B;
  if (blocks < 1) blocks = 1;
  *grid = dim3(static_cast<unsigned>(blocks), B);
  return 0;
}

extern "C" int countr_gn_relu_bwd_reduce(const void* raw, const double* stats, const float* gamma, const float* beta,
                                         const void* d_next, const float* dmap, const float* w1, void* dyh, float* dgamma,
                                         float* dbeta, float* dw1, float* db1, double* gsum, int B, int H, int W, int C, int G,
                                         float eps, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(raw && stats && gamma && beta && dyh && dgamma && dbeta && gsum, "null pointer");
  COUNTR_REQUIRE(C == kC && G == 8, "GroupNorm backward is built for C=256, G=8 (got %d, %d)", C, G);
  const int mode = d_next ? 0 : 1;
  COUNTR_REQUIRE(mode == 0 || (dmap && w1 && dw1 && db1), "1x1-conv mode needs dmap, w1, dw1, db1");
  dim3 grid;
  gn_grid(H * W, B, &grid);
  static const int gather_env = [] { const char* e = getenv("COUNTR_GN_GATHER"); return e ? atoi(e) : 1; }();
  if (mode == 0 && gather_env) {
    auto launch = [&](auto kern, int ty, uint32_t smem, PerDeviceOnce& once) -> int {
      if (once.need()) COUNTR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int ntiles = ((W + GT_X - 1) / GT_X) * ((H + ty - 1) / ty);
      const int bpi = std::max(1, std::min(ntiles, num_sms() / B));       // persistent blocks per image
      kern<<<dim3(bpi, B), 512, smem, stream>>>(reinterpret_cast<const uint16_t*>(raw), stats, gamma, beta, reinterpret_cast<const uint16_t*>(d_next),
                                                reinterpret_cast<uint16_t*>(dyh), dgamma, dbeta, gsum, H, W, G, eps);
      return COUNTR_OK;
    };
    static PerDeviceOnce once_f16, once_bf16;
    const int rc = bf16 ? launch(gn_relu_bwd_gather_kernel<4, 2, true>, 4, GatherCfg<4, 2>::kSmem, once_bf16)
                        : launch(gn_relu_bwd_gather_kernel<4, 2, false>, 4, GatherCfg<4, 2>::kSmem, once_f16);
    if (rc) return rc;
  } else if (mode == 0) {
    // strips of 8 columns x R rows; keep >= ~4 blocks per SM in flight
    const int nxt = (W + 7) / 8;
    int R = 8;
    while (R > 2 && static_cast<long long>(nxt) * ((H + R - 1) / R) * B < 4 * 148) R >>= 1;
    dim3 grid0(nxt * ((H + R - 1) / R), B);
    gn_relu_bwd_reduce_kernel<0><<<grid0, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(raw), stats, gamma, beta,
                                                            reinterpret_cast<const uint16_t*>(d_next), dmap, w1,
                                                            reinterpret_cast<uint16_t*>(dyh), dgamma, dbeta, dw1, db1, gsum, H, W, G, eps, bf16, R);
  } else if (gather_env && H * W >= 4 * HT_PX) {
    static PerDeviceOnce once_h16, once_hbf;
    const int HW = H * W;
    const int ntiles = (HW + HT_PX - 1) / HT_PX;
    const int bpi = std::max(1, std::min(ntiles, num_sms() / B));
    if (bf16) {
      if (once_hbf.need()) COUNTR_CHECK_CUDA(cudaFuncSetAttribute(gn_head_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
      gn_head_reduce_kernel<true><<<dim3(bpi, B), 512, HT_SMEM, stream>>>(reinterpret_cast<const uint16_t*>(raw), stats, gamma, beta, dmap, w1,
                                                                         reinterpret_cast<uint16_t*>(dyh), dgamma, dbeta, dw1, db1, gsum, HW, G, eps);
    } else {
      if (once_h16.need()) COUNTR_CHECK_CUDA(cudaFuncSetAttribute(gn_head_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
      gn_head_reduce_kernel<false><<<dim3(bpi, B), 512, HT_SMEM, stream>>>(reinterpret_cast<const uint16_t*>(raw), stats, gamma, beta, dmap, w1,
                                                                          reinterpret_cast<uint16_t*>(dyh), dgamma, dbeta, dw1, db1, gsum, HW, G, eps);
    }
  } else {
    gn_relu_bwd_reduce_kernel<1><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(raw), stats, gamma, beta,
                                                           reinterpret_cast<const uint16_t*>(d_next), dmap, w1,
                                                           reinterpret_cast<uint16_t*>(dyh), dgamma, dbeta, dw1, db1, gsum, H, W, G, eps, bf16, 0);
  }
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_gn_bwd_apply(const void* raw, const void* dyh, const double* stats, const double* gsum, const float* gamma,
                                   void* d_raw, float* dbias, int B, int HW, int C, int G, float eps, int bf16,
                                   countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(raw && dyh && stats && gsum && gamma && d_raw && dbias, "null pointer");
  COUNTR_REQUIRE(C == kC && G == 8, "GroupNorm backward is built for C=256, G=8 (got %d, %d)", C, G);
  dim3 grid;
  gn_grid(HW, B, &grid);
  gn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(raw), reinterpret_cast<const uint16_t*>(dyh), stats,
                                                gsum, gamma, reinterpret_cast<uint16_t*>(d_raw), dbias, HW, G, eps, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_colsum(const void* x, int dtype, float* out, int64_t R, int N, int64_t ld, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(x && out && R > 0 && N > 0 && N % 2 == 0 && ld % 2 == 0 && dtype >= 0 && dtype <= 2, "bad arguments");
  const int col_blocks = (N / 2 + 127) / 128;
  long long row_blocks = (148 * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (R + 31) / 32) row_blocks = (R + 31) / 32;
  if (row_blocks < 1) row_blocks = 1;
  const int rpb = static_cast<int>((R + row_blocks - 1) / row_blocks);
  dim3 grid(col_blocks, static_cast<unsigned>((R + rpb - 1) / rpb)), block(128, 2);
  COUNTR_CHECK_CUDA(launch_pdl(colsum_kernel, grid, block, 0, stream, x, dtype, out, static_cast<long long>(R), N, static_cast<long long>(ld), rpb));
  return COUNTR_OK;
}

extern "C" int countr_softmax_bwd_rows(void* s_io, void* dp_io, const float* lse, int64_t rows, int L, float scale, int bf16,
                                       countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(s_io && dp_io && lse && rows > 0 && L % 2 == 0, "bad arguments");
  const long long blocks = (rows + 7) / 8;
  softmax_bwd_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<uint16_t*>(s_io),
                                                                            reinterpret_cast<uint16_t*>(dp_io), lse, rows, L, scale, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_cross_attn_core_bwd(const void* q16, const float* k32, const float* v32, const float* probs, const void* do16,
                                          void* dq16, float* dk32, float* dv32, int B, int L, int S, int D, int dh, float scale,
                                          int bf16, int kv_broadcast, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(q16 && k32 && v32 && probs && do16 && dq16 && dk32 && dv32, "null pointer");
  COUNTR_REQUIRE(dh == 32 && D % 512 == 0 && S >= 1 && S <= kMaxShots && L % kTokPerBlock == 0,
                 "cross-attention backward supports dh=32, D%%512==0, S<=8, L%%32==0 (dh=%d D=%d S=%d L=%d)", dh, D, S, L);
  const size_t smem = (2ull * S * D + 2ull * kTokPerBlock * (D / 32) * S) * sizeof(float);
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  COUNTR_REQUIRE(smem <= 96 * 1024, "shared memory %zu too large", smem);
  cross_attn_core_bwd_kernel<<<B * L / kTokPerBlock, 256, smem, stream>>>(
      reinterpret_cast<const uint16_t*>(q16), k32, v32, probs, reinterpret_cast<const uint16_t*>(do16), reinterpret_cast<uint16_t*>(dq16),
      dk32, dv32, L, S, D, scale, kv_broadcast ? 0ll : static_cast<long long>(S) * D, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_inorm_relu_pool_bwd(const void* raw, const float* mean, const float* rstd, const void* dpool16,
                                          const float* dpool32, void* d_raw, float* dbias, float* scratch, int N, int H, int W,
                                          int C, int mode, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(raw && mean && rstd && d_raw && (mode == 0 ? dpool16 != nullptr : dpool32 != nullptr), "null pointer");
  COUNTR_REQUIRE(C % 64 == 0 && (mode == 1 || (H % 2 == 0 && W % 2 == 0)), "bad shape");
  if (mode == 0 && scratch != nullptr && H * W >= 256) {
    const int OHW = H * W / 4, cblocks = C / 64;
    int split = (2 * 148 + cblocks * N - 1) / (cblocks * N);
    if (split > OHW / 16) split = OHW / 16;
    if (split < 1) split = 1;
    const int ppb = (OHW + split - 1) / split;
    dim3 grid(cblocks, N, (OHW + ppb - 1) / ppb);
    COUNTR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * N * C, stream));
    for (int pass = 0; pass < 2; ++pass)
      inorm_pool_bwd_split_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(raw), mean, rstd,
                                                            reinterpret_cast<const uint16_t*>(dpool16), scratch,
                                                            reinterpret_cast<uint16_t*>(d_raw), dbias, H, W, C, ppb, pass, bf16);
    COUNTR_CHECK_CUDA(cudaGetLastError());
    return COUNTR_OK;
  }
  dim3 grid(C / 64, N);
  inorm_relu_pool_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(raw), mean, rstd,
                                                       reinterpret_cast<const uint16_t*>(dpool16), dpool32,
                                                       reinterpret_cast<uint16_t*>(d_raw), dbias, H, W, C, mode, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_exemplar_conv1_dw(const void* boxes, int dtype, int64_t sB, int64_t sK, int64_t sC, int64_t sH, int64_t sW,
                                        const void* d_raw, float* dw, int B, int S, int HW, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static const int sub_env = [] { const char* e = getenv("COUNTR_CONV1_DW_SUB"); return e ? atoi(e) : 4; }();
  const int sub = (sub_env == 2 || sub_env == 4) && (HW * HW) % (128 * sub_env) == 0 ? sub_env : 1;      // 128-pixel tiles per block
  COUNTR_REQUIRE(boxes && d_raw && dw && HW == 64, "the stage-1 exemplar weight gradient is built for 64 x 64 boxes (got %d)", HW);
  dim3 grid(HW * HW / (128 * sub), B * S);
  exemplar_conv1_dw_kernel<<<grid, 256, 0, stream>>>(boxes, dtype, sB, sK, sC, sH, sW, reinterpret_cast<const uint16_t*>(d_raw), dw, S,
                                                     HW, bf16, sub);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_conv_dw_unpack(const float* src, float* dst, int Cout, int Cin, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(src && dst, "null pointer");
  const long long n = static_cast<long long>(Cout) * Cin * 9;
  conv_dw_unpack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, dst, Cout, Cin);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}
