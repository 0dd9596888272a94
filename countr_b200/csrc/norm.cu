// countr_b200 — LayerNorm forward / backward on the fp32 residual stream.
//
// replaces: nn.LayerNorm call sites — timm Block.norm1/norm2 and SupervisedMAE.norm
// (models_mae_cross.py:32-35,146; eps 1e-6 at :214), CrossAttentionBlock.norm0/1/2
// (models_crossvit.py:137,142,147,153-155), decoder_norm (models_mae_cross.py:78,182).
//
// HBM-bound: one warp per row, float4 loads, statistics in fp32 (two-pass over registers, the
// same mean / biased-variance definition as ATen's native_layer_norm), 16-bit output written once.
#include "../../include/countr_b200.h"
#include "common.cuh"

#include <stdlib.h>

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

// D = 128 * V4 * ... : each lane owns NV float4's strided by 32 lanes.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, uint16_t* __restrict__ y16,
                                                             float* __restrict__ y32, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, int rows, int D, float eps,
                                                             int bf16) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * D);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4 v[NV], gm[NV], bt[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  // gamma / beta are fetched together with the row (not after the two reductions): one exposed memory latency, not two
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gm[i] = __ldg(g4 + lane + 32 * i);
    bt[i] = __ldg(b4 + lane + 32 * i);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[warp] = mean;
    if (rstd_out) rstd_out[warp] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = gm[i], b = bt[i];
    const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
    if (y16) {
      uint2 o;
      o.x = pack2(o0, o1, bf16);
      o.y = pack2(o2, o3, bf16);
      reinterpret_cast<uint2*>(y16 + static_cast<size_t>(warp) * D)[lane + 32 * i] = o;
    }
    if (y32) reinterpret_cast<float4*>(y32 + static_cast<size_t>(warp) * D)[lane + 32 * i] = make_float4(o0, o1, o2, o3);
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma/dbeta partials per block -> atomics.
// dy is fp32.  dx is ADDED to dx_accum when accumulate != 0 (the residual branch gradient).
template <int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ mean_in,
                                                             const float* __restrict__ rstd_in, float* __restrict__ dx,
                                                             uint16_t* __restrict__ dx16, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ dx_colsum, int rows,
                                                             int D, int accumulate, int rows_per_warp, int bf16,
                                                             float* __restrict__ partials) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  float4 dg[NV], db[NV], dxs[NV];   // dxs: column sums of the OUTPUT dx = bias gradient of the next Linear in the chain
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = dxs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  // Two blocks per SM (16 warps) hide the row-to-row latency; an explicit one-row prefetch cost 2 x NV float4 registers and kept
  // the kernel at one block per SM (156 registers).
  const int row0 = warp * rows_per_warp;
  float4 gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gm[i] = __ldg(g4 + lane + 32 * i);
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const int row = row0 + rr;
    if (row >= rows) break;
    float4 xc[NV], dc[NV];
    {
      const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
      const float4* dr = reinterpret_cast<const float4*>(dy + static_cast<size_t>(row) * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xc[i] = xr[lane + 32 * i];
        dc[i] = dr[lane + 32 * i];
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[NV], gy[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = xc[i], dv = dc[i], g = gm[i];
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      gy[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
      s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
      s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
      dg[i].x += dv.x * xh[i].x; dg[i].y += dv.y * xh[i].y; dg[i].z += dv.z * xh[i].z; dg[i].w += dv.w * xh[i].w;
      db[i].x += dv.x; db[i].y += dv.y; db[i].z += dv.z; db[i].w += dv.w;
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    float4* dxr = reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o = make_float4(rstd * (gy[i].x - s1 - xh[i].x * s2), rstd * (gy[i].y - s1 - xh[i].y * s2),
                             rstd * (gy[i].z - s1 - xh[i].z * s2), rstd * (gy[i].w - s1 - xh[i].w * s2));
      if (accumulate) {
        const float4 a = dxr[lane + 32 * i];
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      dxr[lane + 32 * i] = o;
      dxs[i].x += o.x; dxs[i].y += o.y; dxs[i].z += o.z; dxs[i].w += o.w;
      if (dx16) {
        uint2 h;
        h.x = pack2(o.x, o.y, bf16);
        h.y = pack2(o.z, o.w, bf16);
        reinterpret_cast<uint2*>(dx16 + static_cast<size_t>(row) * D)[lane + 32 * i] = h;
      }
    }
  }
  // block-level reduction of dgamma/dbeta over the 8 warps, then one atomic per column per block — or, with `partials`
  // ([3][gridDim.x][D]: dgamma, dbeta, column sums of dx), one plain store per column per block: the caller sums the
  // block rows later (grouped column sums at the end of the backward), so the kernel has no same-address atomics and can
  // run four blocks per SM
  __shared__ float4 red[8][32];
  const int w = threadIdx.x >> 5;
  if (partials != nullptr) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        red[w][lane] = pass == 0 ? dg[i] : pass == 1 ? db[i] : dxs[i];
        __syncthreads();
        if (w == 0) {
          float4 a = red[0][lane];
#pragma unroll
          for (int k = 1; k < 8; ++k) {
            a.x += red[k][lane].x; a.y += red[k][lane].y; a.z += red[k][lane].z; a.w += red[k][lane].w;
          }
          reinterpret_cast<float4*>(partials + (static_cast<size_t>(pass) * gridDim.x + blockIdx.x) * D)[lane + 32 * i] = a;
        }
        __syncthreads();
      }
    }
    return;
  }
  if (dgamma == nullptr && dx_colsum == nullptr) return;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    if (pass < 2 && dgamma == nullptr) continue;
    if (pass == 2 && dx_colsum == nullptr) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      red[w][lane] = pass == 0 ? dg[i] : pass == 1 ? db[i] : dxs[i];
      __syncthreads();
      if (w == 0) {
        float4 a = red[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
          a.x += red[k][lane].x; a.y += red[k][lane].y; a.z += red[k][lane].z; a.w += red[k][lane].w;
        }
        float* dst = (pass == 0 ? dgamma : pass == 1 ? dbeta : dx_colsum) + (lane + 32 * i) * 4;
        atomicAdd(dst, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, a.z); atomicAdd(dst + 3, a.w);
      }
      __syncthreads();
    }
  }
}

}  // namespace
}  // namespace countr

extern "C" int countr_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y16, float* y32,
                                    float* mean, float* rstd, int rows, int D, float eps, int bf16,
                                    countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(x && gamma && beta && (y16 || y32), "null pointer");
  COUNTR_REQUIRE(rows > 0 && D % 128 == 0 && D <= 1536, "LayerNorm width %d unsupported (multiple of 128, <= 1536)", D);
  const int blocks = (rows + 7) / 8;
  uint16_t* y = reinterpret_cast<uint16_t*>(y16);
#define LN_CASE(NV)                                                                                             \
  case NV:                                                                                                      \
    COUNTR_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel<NV>, dim3(blocks), dim3(256), 0, stream, x, gamma, beta, y, y32, mean, rstd, rows, D, eps, bf16)); \
    break;
  switch (D / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(8) LN_CASE(10) LN_CASE(12)
    default:
      return set_error(COUNTR_ERR_UNSUPPORTED, "LayerNorm width %d not instantiated", D);
  }
#undef LN_CASE
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

namespace {
// grid of the partial-sums mode: two 8-warp blocks per SM (what 128 registers allow; no atomics to contend on, memory latency is
// what is left)
void ln_bwd_partial_grid(int rows, int* rpw, int* blocks) {
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = getenv("COUNTR_LN_BWD_BLOCKS_PER_SM");
    per_sm = e != nullptr ? atoi(e) : 2;     // measured: 2 blocks per SM 1875 vs 1902 us (fine-tune backward), 18.2 vs 19.1 ms (pre-train step)
    if (per_sm < 1) per_sm = 1;
  }
  const int target_warps = 148 * 8 * per_sm;
  int r = (rows + target_warps - 1) / target_warps;
  if (r < 1) r = 1;
  const int warps = (rows + r - 1) / r;
  *rpw = r;
  *blocks = (warps + 7) / 8;
}
}  // namespace

extern "C" int countr_layernorm_bwd_blocks(int rows) {
  int rpw, blocks;
  ln_bwd_partial_grid(rows, &rpw, &blocks);
  return blocks;
}

extern "C" int countr_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                    const float* rstd, float* dx, void* dx16, float* dgamma, float* dbeta, float* dx_colsum,
                                    float* partials, int rows, int D, int accumulate, int bf16, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(dy && x && gamma && mean && rstd && dx, "null pointer");
  COUNTR_REQUIRE(partials == nullptr || (dgamma == nullptr && dbeta == nullptr && dx_colsum == nullptr &&
                                         (reinterpret_cast<uintptr_t>(partials) & 15u) == 0),
                 "partials mode: 16-byte aligned [3][countr_layernorm_bwd_blocks(rows)][D] buffer, dgamma / dbeta / dx_colsum NULL");
  COUNTR_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "dgamma/dbeta must both be given or both NULL");
  COUNTR_REQUIRE(rows > 0 && D % 128 == 0 && D <= 1536, "LayerNorm width %d unsupported", D);
  // one wave of 8-warp blocks: fewer blocks = fewer same-address dgamma/dbeta atomics
  const int target_warps = 148 * 8;    // (two blocks per SM measured slower: 149 vs 132 us over the 7 launches of a step — the atomics dominate)
  int rpw = (rows + target_warps - 1) / target_warps;
  if (rpw < 1) rpw = 1;
  const int warps = (rows + rpw - 1) / rpw;
  int blocks = (warps + 7) / 8;
  if (partials != nullptr) ln_bwd_partial_grid(rows, &rpw, &blocks);
#define LN_CASE(NV)                                                                                       \
  case NV:                                                                                                \
    COUNTR_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<NV>, dim3(blocks), dim3(256), 0, stream, dy, x, gamma, mean, rstd, dx,  \
                                 reinterpret_cast<uint16_t*>(dx16), dgamma, dbeta, dx_colsum, rows, D, accumulate, rpw, bf16, partials)); \
    break;
  switch (D / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(8) LN_CASE(10) LN_CASE(12)
    default:
      return set_error(COUNTR_ERR_UNSUPPORTED, "LayerNorm width %d not instantiated", D);
  }
#undef LN_CASE
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}
