// countr_b200 — HBM-bound layout / normalisation / resampling kernels around the tensor-core ops.
//
// None of these is GEMM shaped; they are coalesced, 16-byte-vectorised streaming kernels whose
// job is to touch every activation exactly once between two tensor-core kernels.
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
__device__ __forceinline__ float load_any(const void* p, long long i, int dtype) {
  // dtype: 0 fp32, 1 fp16, 2 bf16
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  const uint16_t u = reinterpret_cast<const uint16_t*>(p)[i];
  if (dtype == 1) return __half2float(__ushort_as_half(u));
  return __uint_as_float(static_cast<uint32_t>(u) << 16);
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

// ------------------------------------------------------------------------------------------
// fp32 -> 16-bit casts (weights each step, gradients into GEMM operands)
// ------------------------------------------------------------------------------------------
__global__ void cast_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n, float scale,
                            int bf16) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
    uint4 o;
    o.x = pack2(a.x * scale, a.y * scale, bf16);
    o.y = pack2(a.z * scale, a.w * scale, bf16);
    o.z = pack2(b.x * scale, b.y * scale, bf16);
    o.w = pack2(b.z * scale, b.w * scale, bf16);
    *reinterpret_cast<uint4*>(dst + i) = o;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = static_cast<uint16_t>(pack2(src[j] * scale, 0.f, bf16) & 0xffffu);
  }
}

// dst[c][r] = src[r][c]   (fp32 [R][C] -> 16-bit [C][R]); 32x32 tiles through padded smem
__global__ void cast_transpose_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int R, int C,
                                      int bf16) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[static_cast<size_t>(r) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[static_cast<size_t>(c) * R + r] = static_cast<uint16_t>(pack2(tile[threadIdx.x][i], 0.f, bf16) & 0xffffu);
  }
}

// ------------------------------------------------------------------------------------------
// PatchEmbed gather: NCHW image (any float dtype, any strides) -> A[B*gh*gw][C*P*P] 16-bit,
// column order (c, ky, kx) == Conv2d weight.view(out, -1)   (models_mae_cross.py:27,138)
// ------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const void* __restrict__ img, int dtype, long long sb, long long sc, long long sh,
                                long long sw, uint16_t* __restrict__ out, int B, int C, int H, int W, int P, int bf16) {
  // one thread = 8 consecutive kx of one (patch, c, ky)
  const int gw = W / P, gh = H / P;
  const int vec_per_row = P / 8;
  const long long total = static_cast<long long>(B) * gh * gw * C * P * vec_per_row;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long t = idx;
  const int kxv = t % vec_per_row; t /= vec_per_row;
  const int ky = t % P; t /= P;
  const int c = t % C; t /= C;
  const int px = t % gw; t /= gw;
  const int py = t % gh; t /= gh;
  const int b = static_cast<int>(t);
  const long long src = b * sb + c * sc + static_cast<long long>(py * P + ky) * sh + static_cast<long long>(px * P + kxv * 8) * sw;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = load_any(img, src + j * sw, dtype);
  const long long row = (static_cast<long long>(b) * gh + py) * gw + px;
  const long long col = (static_cast<long long>(c) * P + ky) * P + kxv * 8;
  *reinterpret_cast<uint4*>(out + row * (static_cast<long long>(C) * P * P) + col) = pack8(f, bf16);
}

// Any patch size / image size (mae_vit_huge_patch14: 14-px patches on a 384-px image -> 27 x 27 patches, the last 6 pixels of
// every row / column are dropped like Conv2d(stride = kernel) drops them): one thread per output element, rows padded with
// zeros to `ld` elements (a multiple of 8: TMA needs 16-byte row pitches).
__global__ void patchify_generic_kernel(const void* __restrict__ img, int dtype, long long sb, long long sc, long long sh, long long sw,
                                        uint16_t* __restrict__ out, int B, int C, int H, int W, int P, int ld, int bf16) {
  const int gw = W / P, gh = H / P;
  const long long total = static_cast<long long>(B) * gh * gw * ld;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int col = static_cast<int>(idx % ld);
  const long long row = idx / ld;
  float v = 0.f;
  if (col < C * P * P) {
    const int kx = col % P, ky = (col / P) % P, c = col / (P * P);
    const int px = static_cast<int>(row % gw), py = static_cast<int>((row / gw) % gh), b = static_cast<int>(row / (static_cast<long long>(gw) * gh));
    v = load_any(img, b * sb + c * sc + static_cast<long long>(py * P + ky) * sh + static_cast<long long>(px * P + kx) * sw, dtype);
  }
  float f[8] = {v, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  out[idx] = static_cast<uint16_t>(pack8(f, bf16).x & 0xffffu);
}

// ------------------------------------------------------------------------------------------
// Conv2d 3x3 weight [Cout][Cin][3][3] fp32 -> GEMM B operand, 16-bit:
//   mode 0 (forward):  out[co][(ky*3+kx)*Cin + ci] = w[co][ci][ky][kx]
//   mode 1 (dX):       out[ci][((2-ky)*3+(2-kx))*Cout + co] = w[co][ci][ky][kx]
//                      (correlation with the flipped, channel-transposed filter)
// ------------------------------------------------------------------------------------------
__global__ void conv_weight_pack_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cout, int Cin,
                                        int mode, int bf16) {
  const long long n = static_cast<long long>(Cout) * Cin * 9;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // iterate in OUTPUT order so the stores coalesce
  long long t = i;
  if (mode == 0) {
    const int ci = t % Cin; t /= Cin;
    const int tap = t % 9; t /= 9;
    const int co = static_cast<int>(t);
    out[i] = static_cast<uint16_t>(pack2(w[(static_cast<long long>(co) * Cin + ci) * 9 + tap], 0.f, bf16) & 0xffffu);
  } else {
    const int co = t % Cout; t /= Cout;
    const int tap = t % 9; t /= 9;
    const int ci = static_cast<int>(t);
    out[i] = static_cast<uint16_t>(pack2(w[(static_cast<long long>(co) * Cin + ci) * 9 + (8 - tap)], 0.f, bf16) & 0xffffu);
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm(apply) + ReLU + bilinear x2 (align_corners=False), NHWC 16-bit -> NHWC 16-bit.
//   y = up2( relu( (x - mean_g) * rstd_g * gamma_c + beta_c ) )
// replaces: nn.GroupNorm(8,256) + ReLU + F.interpolate(x2) (models_mae_cross.py:80-95,189-194)
// stats: [B][G][2] double (sum, sumsq) accumulated by the conv epilogue.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_relu_up2_kernel(const uint16_t* __restrict__ x, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           uint16_t* __restrict__ y, int H, int W, int C, int G, float eps,
                                                           int bf16) {
  // thread = one INPUT pixel x 8 channels -> the 2x2 output block it is the centre of:
  //   out[2i]   = .25 a[i-1] + .75 a[i],   out[2i+1] = .75 a[i] + .25 a[i+1]      (indices clamped)
  // so the 3x3 input neighbourhood is normalised once (9 instead of 16 activations per 4 outputs) and
  // blended separably.  The kernel is instruction-bound, not HBM-bound, hence the restructuring.
  extern __shared__ float sm_ab[];  // [2][C]
  float* sa = sm_ab;
  float* sb = sm_ab + C;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const double cnt = static_cast<double>(H) * W * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double s = stats[(static_cast<long long>(b) * G + g) * 2], ss = stats[(static_cast<long long>(b) * G + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - static_cast<float>(mean) * a;
  }
  __syncthreads();
  const int vecs = C / 8;
  const int OW = 2 * W;
  const long long total = static_cast<long long>(H) * W * vecs;
  const uint16_t* xb0 = x + static_cast<long long>(b) * H * W * C;
  uint16_t* yb0 = y + static_cast<long long>(b) * 4 * H * W * C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = idx % vecs;
    const int pix = static_cast<int>(idx / vecs);
    const int ix = pix % W, iy = pix / W;
    const int xs[3] = {max(ix - 1, 0), ix, min(ix + 1, W - 1)};
    const int ys[3] = {max(iy - 1, 0), iy, min(iy + 1, H - 1)};
    float av[8], bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      av[j] = sa[cv * 8 + j];
      bv[j] = sb[cv * 8 + j];
    }
    uint4 raw[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) raw[r][c] = *reinterpret_cast<const uint4*>(xb0 + (static_cast<long long>(ys[r]) * W + xs[c]) * C + cv * 8);
    float hl[3][8], hr[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float a0[8], a1[8], a2[8];
      unpack8(raw[r][0], a0, bf16);
      unpack8(raw[r][1], a1, bf16);
      unpack8(raw[r][2], a2, bf16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v0 = fmaxf(fmaf(a0[j], av[j], bv[j]), 0.f), v1 = fmaxf(fmaf(a1[j], av[j], bv[j]), 0.f);
        const float v2 = fmaxf(fmaf(a2[j], av[j], bv[j]), 0.f);
        hl[r][j] = 0.25f * v0 + 0.75f * v1;
        hr[r][j] = 0.75f * v1 + 0.25f * v2;
      }
    }
    float o[8];
    uint16_t* yo = yb0 + (static_cast<long long>(2 * iy) * OW + 2 * ix) * C + cv * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * hl[0][j] + 0.75f * hl[1][j];
    *reinterpret_cast<uint4*>(yo) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * hr[0][j] + 0.75f * hr[1][j];
    *reinterpret_cast<uint4*>(yo + C) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.75f * hl[1][j] + 0.25f * hl[2][j];
    *reinterpret_cast<uint4*>(yo + static_cast<long long>(OW) * C) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.75f * hr[1][j] + 0.25f * hr[2][j];
    *reinterpret_cast<uint4*>(yo + static_cast<long long>(OW) * C + C) = pack8(o, bf16);
  }
}

// Row-walking variant (used when 256 % (C/8) == 0): a block owns a strip of 256/(C/8) pixel columns x R rows and walks
// down the rows, keeping the horizontally blended, normalised rows iy-1, iy, iy+1 in registers.  Every input row of the
// strip is fetched and normalised ONCE (3 loads + 24 normalisations per 4 output pixels instead of 9 + 72), and the rows
// a step re-uses never leave the SM — the pixel-per-thread kernel above re-read its 3x3 neighbourhood from L1/L2 and was
// issue-bound (2.3 TB/s at 96^2 -> 192^2).
__global__ void __launch_bounds__(256) gn_relu_up2_rows_kernel(const uint16_t* __restrict__ x, const double* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                uint16_t* __restrict__ y, int H, int W, int C, int G, float eps,
                                                                int R, int bf16) {
  extern __shared__ float sm_ab[];  // [2][C]
  float* sa = sm_ab;
  float* sb = sm_ab + C;
  const int b = blockIdx.z;
  const int cpg = C / G;
  const double cnt = static_cast<double>(H) * W * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double s = stats[(static_cast<long long>(b) * G + g) * 2], ss = stats[(static_cast<long long>(b) * G + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - static_cast<float>(mean) * a;
  }
  __syncthreads();
  const int vecs = C / 8;
  const int cv = threadIdx.x % vecs;
  const int ix = blockIdx.x * (256 / vecs) + threadIdx.x / vecs;
  if (ix >= W) return;
  const int y0 = blockIdx.y * R, y1 = min(H, y0 + R);
  const int OW = 2 * W;
  const int xl = max(ix - 1, 0), xr = min(ix + 1, W - 1);
  const uint16_t* xb = x + static_cast<long long>(b) * H * W * C + cv * 8;
  uint16_t* yb = y + static_cast<long long>(b) * 4 * H * W * C + cv * 8;
  float av[8], bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    av[j] = sa[cv * 8 + j];
    bv[j] = sb[cv * 8 + j];
  }
  // hl / hr: the two horizontally blended output columns (2 ix, 2 ix + 1) of one normalised input row
  auto load_row = [&](int row, uint4 (&u)[3]) {
    const uint16_t* rp = xb + static_cast<long long>(row) * W * C;
    u[0] = *reinterpret_cast<const uint4*>(rp + static_cast<long long>(xl) * C);
    u[1] = *reinterpret_cast<const uint4*>(rp + static_cast<long long>(ix) * C);
    u[2] = *reinterpret_cast<const uint4*>(rp + static_cast<long long>(xr) * C);
  };
  auto blend = [&](const uint4 (&u)[3], float (&hl)[8], float (&hr)[8]) {
    float a0[8], a1[8], a2[8];
    unpack8(u[0], a0, bf16);
    unpack8(u[1], a1, bf16);
    unpack8(u[2], a2, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v0 = fmaxf(fmaf(a0[j], av[j], bv[j]), 0.f), v1 = fmaxf(fmaf(a1[j], av[j], bv[j]), 0.f);
      const float v2 = fmaxf(fmaf(a2[j], av[j], bv[j]), 0.f);
      hl[j] = 0.25f * v0 + 0.75f * v1;
      hr[j] = 0.75f * v1 + 0.25f * v2;
    }
  };
  float hl0[8], hr0[8], hl1[8], hr1[8], hl2[8], hr2[8];
  uint4 ua[3], ub[3], un[3];
  load_row(max(y0 - 1, 0), ua);
  load_row(y0, ub);
  load_row(min(y0 + 1, H - 1), un);             // the row the first iteration needs: three rows in flight at once
  blend(ua, hl0, hr0);
  blend(ub, hl1, hr1);
  for (int iy = y0; iy < y1; ++iy) {
    // one row of look-ahead: the loads of row iy + 2 are in flight while row iy's four output rows are blended and stored (the
    // walk was one dependent L2 / HBM round trip per row step at 16 warps per SM)
    uint4 uc[3] = {un[0], un[1], un[2]};
    if (iy + 1 < y1) load_row(min(iy + 2, H - 1), un);
    blend(uc, hl2, hr2);
    float o[8];
    uint16_t* yo = yb + (static_cast<long long>(2 * iy) * OW + 2 * ix) * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * hl0[j] + 0.75f * hl1[j];
    *reinterpret_cast<uint4*>(yo) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * hr0[j] + 0.75f * hr1[j];
    *reinterpret_cast<uint4*>(yo + C) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.75f * hl1[j] + 0.25f * hl2[j];
    *reinterpret_cast<uint4*>(yo + static_cast<long long>(OW) * C) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.75f * hr1[j] + 0.25f * hr2[j];
    *reinterpret_cast<uint4*>(yo + static_cast<long long>(OW) * C + C) = pack8(o, bf16);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hl0[j] = hl1[j]; hr0[j] = hr1[j];
      hl1[j] = hl2[j]; hr1[j] = hr2[j];
    }
  }
}

// Shared-memory staged variant for C == 256 (the density head): persistent blocks per image; a tile = 8 x 8 input pixels ->
// 16 x 16 output pixels, its 10 x 10 input window arrives as ten row-contiguous bulk copies (cp.async.bulk, mbarrier completion)
// into one of three 51 KB stages, so two tiles are in flight while one is blended and stored.  The row-walking kernel above waits
// for a global round trip per row step; here the walk reads shared memory only and the kernel runs at its store rate.
constexpr int UT_X = 8, UT_Y = 8, UT_NS = 3, UT_C = 256;
constexpr uint32_t UT_PIX = UT_C * 2, UT_PITCH = (UT_X + 2) * UT_PIX, UT_STAGE = (UT_Y + 2) * UT_PITCH;
constexpr uint32_t UT_SMEM = UT_NS * UT_STAGE + 64 + 2 * UT_C * sizeof(float);

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(512, 1) gn_relu_up2_staged_kernel(const uint16_t* __restrict__ x, const double* __restrict__ stats,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    uint16_t* __restrict__ y, int H, int W, int G, float eps, int bf16) {
  constexpr int C = UT_C;
  extern __shared__ __align__(128) uint8_t ut_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(ut_smem + UT_NS * UT_STAGE);
  float* sa = reinterpret_cast<float*>(ut_smem + UT_NS * UT_STAGE + 64);
  float* sb = sa + C;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const double cnt = static_cast<double>(H) * W * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double s = stats[(static_cast<long long>(b) * G + g) * 2], ss = stats[(static_cast<long long>(b) * G + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - static_cast<float>(mean) * a;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < UT_NS; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int cv = threadIdx.x & 31, pl = threadIdx.x >> 5;
  float av[8], bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    av[j] = sa[cv * 8 + j];
    bv[j] = sb[cv * 8 + j];
  }
  const uint16_t* xb = x + static_cast<long long>(b) * H * W * C;
  uint16_t* yb = y + static_cast<long long>(b) * 4 * H * W * C + cv * 8;
  const int OW = 2 * W;
  const int ntx = (W + UT_X - 1) / UT_X, ntiles = ntx * ((H + UT_Y - 1) / UT_Y);

  auto window = [&](int t, int& y0, int& x0, int& wy0, int& wy1, int& wx0, int& wx1) {
    y0 = (t / ntx) * UT_Y;
    x0 = (t % ntx) * UT_X;
    wy0 = max(y0 - 1, 0), wy1 = min(y0 + UT_Y, H - 1);
    wx0 = max(x0 - 1, 0), wx1 = min(x0 + UT_X, W - 1);
  };
  auto issue = [&](int t, int stage) {      // one thread
    int y0, x0, wy0, wy1, wx0, wx1;
    window(t, y0, x0, wy0, wy1, wx0, wx1);
    const uint32_t row_bytes = static_cast<uint32_t>(wx1 - wx0 + 1) * UT_PIX;
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bars + stage, row_bytes * static_cast<uint32_t>(wy1 - wy0 + 1));
    const uint32_t dst = smem_u32(ut_smem) + stage * UT_STAGE;
    for (int r = wy0; r <= wy1; ++r)
      bulk_g2s(dst + (r - wy0) * UT_PITCH, xb + (static_cast<long long>(r) * W + wx0) * C, row_bytes, bars + stage);
  };

  int it = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < UT_NS - 1; ++k)
      if (static_cast<int>(blockIdx.x + k * gridDim.x) < ntiles) issue(blockIdx.x + k * gridDim.x, k);
  }
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int stage = it % UT_NS;
    if (threadIdx.x == 0 && t + (UT_NS - 1) * static_cast<int>(gridDim.x) < ntiles) issue(t + (UT_NS - 1) * gridDim.x, (it + UT_NS - 1) % UT_NS);
    int y0, x0, wy0, wy1, wx0, wx1;
    window(t, y0, x0, wy0, wy1, wx0, wx1);
    const int ix = x0 + (pl & 7);
    const int ys = y0 + (UT_Y / 2) * (pl >> 3), ye = min(min(H, y0 + UT_Y), ys + UT_Y / 2);      // this warp's rows
    mbar_wait(bars + stage, static_cast<uint32_t>(it / UT_NS) & 1u);
    if (ix < W && ys < ye) {
      const uint8_t* st = ut_smem + stage * UT_STAGE + cv * 16;
      const uint32_t ol = static_cast<uint32_t>(max(ix - 1, 0) - wx0) * UT_PIX, oc = static_cast<uint32_t>(ix - wx0) * UT_PIX;
      const uint32_t orr = static_cast<uint32_t>(min(ix + 1, W - 1) - wx0) * UT_PIX;
      // hl / hr: the two horizontally blended output columns (2 ix, 2 ix + 1) of one normalised input row
      auto blend = [&](int row, float (&hl)[8], float (&hr)[8]) {
        const uint8_t* rp = st + (row - wy0) * UT_PITCH;
        float a0[8], a1[8], a2[8];
        unpack8(*reinterpret_cast<const uint4*>(rp + ol), a0, bf16);
        unpack8(*reinterpret_cast<const uint4*>(rp + oc), a1, bf16);
        unpack8(*reinterpret_cast<const uint4*>(rp + orr), a2, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v0 = fmaxf(fmaf(a0[j], av[j], bv[j]), 0.f), v1 = fmaxf(fmaf(a1[j], av[j], bv[j]), 0.f);
          const float v2 = fmaxf(fmaf(a2[j], av[j], bv[j]), 0.f);
          hl[j] = 0.25f * v0 + 0.75f * v1;
          hr[j] = 0.75f * v1 + 0.25f * v2;
        }
      };
      float hl0[8], hr0[8], hl1[8], hr1[8], hl2[8], hr2[8];
      blend(max(ys - 1, 0), hl0, hr0);
      blend(ys, hl1, hr1);
      for (int iy = ys; iy < ye; ++iy) {
        blend(min(iy + 1, H - 1), hl2, hr2);
        float o[8];
        uint16_t* yo = yb + (static_cast<long long>(2 * iy) * OW + 2 * ix) * C;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.25f * hl0[j] + 0.75f * hl1[j];
        *reinterpret_cast<uint4*>(yo) = pack8(o, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.25f * hr0[j] + 0.75f * hr1[j];
        *reinterpret_cast<uint4*>(yo + C) = pack8(o, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.75f * hl1[j] + 0.25f * hl2[j];
        *reinterpret_cast<uint4*>(yo + static_cast<long long>(OW) * C) = pack8(o, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.75f * hr1[j] + 0.25f * hr2[j];
        *reinterpret_cast<uint4*>(yo + static_cast<long long>(OW) * C + C) = pack8(o, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hl0[j] = hl1[j]; hr0[j] = hr1[j];
          hl1[j] = hl2[j]; hr1[j] = hr2[j];
        }
      }
    }
    __syncthreads();          // everyone is done with this stage before the next iteration refills it
  }
}

// GroupNorm + ReLU + Conv2d 1x1 (C -> 1): one warp per pixel.   (models_mae_cross.py:96-100)
__global__ void __launch_bounds__(256) gn_relu_dot_kernel(const uint16_t* __restrict__ x, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ out, int HW, int C, int G, float eps,
                                                           int bf16) {
  extern __shared__ float sm_abw[];  // [3][C]
  float* sa = sm_abw;
  float* sb = sm_abw + C;
  float* sw = sm_abw + 2 * C;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const double cnt = static_cast<double>(HW) * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double s = stats[(static_cast<long long>(b) * G + g) * 2], ss = stats[(static_cast<long long>(b) * G + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - static_cast<float>(mean) * a;
    sw[c] = w[c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float b0 = bias[0];
  // one warp handles 4 pixels per iteration: 4 independent 16-byte loads in flight per lane (the kernel is
  // latency-bound with one), and the four warp reductions interleave.
  if (C == 256) {
    // the lane's 8 channels never change: scale / shift / 1x1 weight live in registers (the generic loop below re-reads
    // them from shared memory with an 8-way bank conflict for every pixel group)
    const int c0 = lane * 8;
    float av[8], bv[8], wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      av[j] = sa[c0 + j];
      bv[j] = sb[c0 + j];
      wv[j] = sw[c0 + j];
    }
    const uint16_t* xb = x + static_cast<long long>(b) * HW * C + c0;
    for (int pix0 = (blockIdx.x * 8 + warp) * 4; pix0 < HW; pix0 += gridDim.x * 8 * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(min(pix0 + u, HW - 1)) * C);
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f, bf16);
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a = fmaf(fmaxf(fmaf(f[j], av[j], bv[j]), 0.f), wv[j], a);
        acc[u] = a;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
      }
      if (lane < 4 && pix0 + lane < HW) {
        const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        out[static_cast<long long>(b) * HW + pix0 + lane] = r + b0;
      }
    }
    return;
  }
  for (int pix0 = (blockIdx.x * 8 + warp) * 4; pix0 < HW; pix0 += gridDim.x * 8 * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pix = min(pix0 + u, HW - 1);
        v[u] = *reinterpret_cast<const uint4*>(x + (static_cast<long long>(b) * HW + pix) * C + c0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u] += fmaxf(fmaf(f[j], sa[c0 + j], sb[c0 + j]), 0.f) * sw[c0 + j];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane < 4 && pix0 + lane < HW) {
      const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      out[static_cast<long long>(b) * HW + pix0 + lane] = r + b0;
    }
  }
}

// Staged variant for C == 256: the image's pixels are one contiguous stream; a tile = 64 pixels = one 32 KB bulk copy, four stages
// per persistent block (three tiles in flight), 16 warps x 4 pixels per tile.
constexpr int DT_PX = 64, DT_NS = 4;
constexpr uint32_t DT_STAGE = DT_PX * UT_C * 2;
constexpr uint32_t DT_SMEM = DT_NS * DT_STAGE + 64 + 3 * UT_C * sizeof(float);

__global__ void __launch_bounds__(512, 1) gn_relu_dot_staged_kernel(const uint16_t* __restrict__ x, const double* __restrict__ stats,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    const float* __restrict__ w, const float* __restrict__ bias,
                                                                    float* __restrict__ out, int HW, int G, float eps, int bf16) {
  constexpr int C = UT_C;
  extern __shared__ __align__(128) uint8_t dt_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(dt_smem + DT_NS * DT_STAGE);
  float* sa = reinterpret_cast<float*>(dt_smem + DT_NS * DT_STAGE + 64);
  float* sb = sa + C;
  float* sw = sb + C;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const double cnt = static_cast<double>(HW) * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double s = stats[(static_cast<long long>(b) * G + g) * 2], ss = stats[(static_cast<long long>(b) * G + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - static_cast<float>(mean) * a;
    sw[c] = w[c];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < DT_NS; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float b0 = bias[0];
  const int c0 = lane * 8;
  float av[8], bv[8], wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    av[j] = sa[c0 + j];
    bv[j] = sb[c0 + j];
    wv[j] = sw[c0 + j];
  }
  const uint16_t* xb = x + static_cast<long long>(b) * HW * C;
  float* ob = out + static_cast<long long>(b) * HW;
  const int ntiles = (HW + DT_PX - 1) / DT_PX;
  auto issue = [&](int t, int stage) {      // one thread
    const int npx = min(DT_PX, HW - t * DT_PX);
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bars + stage, static_cast<uint32_t>(npx) * C * 2);
    const uint32_t dst = smem_u32(dt_smem) + stage * DT_STAGE;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(xb + static_cast<long long>(t) * DT_PX * C);
    const uint32_t half = static_cast<uint32_t>(npx / 2) * C * 2, rest = static_cast<uint32_t>(npx) * C * 2 - half;
    if (half) bulk_g2s(dst, src, half, bars + stage);
    bulk_g2s(dst + half, src + half, rest, bars + stage);
  };
  int it = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < DT_NS - 1; ++k)
      if (static_cast<int>(blockIdx.x + k * gridDim.x) < ntiles) issue(blockIdx.x + k * gridDim.x, k);
  }
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int stage = it % DT_NS;
    if (threadIdx.x == 0 && t + (DT_NS - 1) * static_cast<int>(gridDim.x) < ntiles) issue(t + (DT_NS - 1) * gridDim.x, (it + DT_NS - 1) % DT_NS);
    const int p0 = t * DT_PX;
    mbar_wait(bars + stage, static_cast<uint32_t>(it / DT_NS) & 1u);
    const uint8_t* st = dt_smem + stage * DT_STAGE + c0 * 2;
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // pixels past the end of a ragged last tile read stale shared memory; their results are not stored
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(st + (warp + 16 * u) * (C * 2)), f, bf16);
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(fmaxf(fmaf(f[j], av[j], bv[j]), 0.f), wv[j], a);
      acc[u] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane < 4) {
      const int pk = p0 + warp + 16 * lane;
      const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      if (pk < HW) ob[pk] = r + b0;
    }
    __syncthreads();
  }
}

// bilinear x2 of a single-channel fp32 map [B][H][W] -> out (fp32 / fp16 / bf16) [B][2H][2W]
__global__ void up2_f32_kernel(const float* __restrict__ x, void* __restrict__ y, int B, int H, int W, int out_dtype) {
  const int OW = 2 * W, OH = 2 * H;
  const long long total = static_cast<long long>(B) * OH * OW;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = idx % OW;
  const int oy = (idx / OW) % OH;
  const int b = idx / (static_cast<long long>(OW) * OH);
  const int ix = ox >> 1, iy = oy >> 1;
  const int x0 = (ox & 1) ? ix : max(ix - 1, 0), x1 = (ox & 1) ? min(ix + 1, W - 1) : ix;
  const int y0 = (oy & 1) ? iy : max(iy - 1, 0), y1 = (oy & 1) ? min(iy + 1, H - 1) : iy;
  const float wx1 = (ox & 1) ? 0.25f : 0.75f, wy1 = (oy & 1) ? 0.25f : 0.75f;
  const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  const float* xb = x + static_cast<long long>(b) * H * W;
  const float v = wy0 * (wx0 * xb[y0 * W + x0] + wx1 * xb[y0 * W + x1]) + wy1 * (wx0 * xb[y1 * W + x0] + wx1 * xb[y1 * W + x1]);
  if (out_dtype == 0)
    reinterpret_cast<float*>(y)[idx] = v;
  else
    reinterpret_cast<uint16_t*>(y)[idx] = static_cast<uint16_t>(pack2(v, 0.f, out_dtype == 2) & 0xffffu);
}

// ------------------------------------------------------------------------------------------
// Exemplar CNN stage 1: Conv2d(3,64,3,p=1) on [B][K][3][64][64] boxes (first S of K used),
// direct fp32 FMA (27 MACs per output; not worth a tensor-core tile) -> NHWC 16-bit raw output.
// replaces: decoder_proj1[0] (models_mae_cross.py:47-48,166); sample index n = b*S + s.
// ------------------------------------------------------------------------------------------
// block = 2 image rows (128 pixels) x 64 output channels; thread = 4 consecutive pixels x 8 channels, so every
// weight vector read from smem (2 x LDS.128) feeds 32 FMAs.
__global__ void __launch_bounds__(256) exemplar_conv1_kernel(const void* __restrict__ boxes, int dtype, long long sB,
                                                              long long sK, long long sC, long long sH, long long sW,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              uint16_t* __restrict__ out, int S, int HW, int Cout, int bf16) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float sbias[64];
  __shared__ float s_in[3][4][64 + 4];   // rows y0-1 .. y0+2, columns -1 .. 64 (zero halo), +pad
  const int n = blockIdx.y;
  const int b = n / S, s = n % S;
  const int y0 = blockIdx.x * 2;
  const int H = HW, W = HW;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;  // k = ci*9 + tap
    sw[k * Cout + co] = w[co * 27 + k];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sbias[i] = bias[i];
  for (int i = threadIdx.x; i < 3 * 4 * 66; i += blockDim.x) {
    const int xx = i % 66 - 1, r = (i / 66) % 4, ci = i / (66 * 4);
    const int yy = y0 + r - 1;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = load_any(boxes, b * sB + s * sK + ci * sC + yy * sH + xx * sW, dtype);
    s_in[ci][r][xx + 1] = v;
  }
  __syncthreads();
  const int cg = threadIdx.x & 7;           // 8 output channels
  const int pg = threadIdx.x >> 3;          // 4 consecutive pixels
  const int ry = pg >> 4, x0 = (pg & 15) * 4;
  float acc[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[q][j] = sbias[cg * 8 + j];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      float in[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) in[t] = s_in[ci][ry + ky][x0 + t];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int k = ci * 9 + ky * 3 + kx;
        const float4 w0 = *reinterpret_cast<const float4*>(&sw[k * 64 + cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sw[k * 64 + cg * 8 + 4]);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[q][j] += in[q + kx] * wv[j];
      }
    }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pix = (y0 + ry) * W + x0 + q;
    *reinterpret_cast<uint4*>(out + (static_cast<long long>(n) * H * W + pix) * 64 + cg * 8) = pack8(acc[q], bf16);
  }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm statistics, pixel-parallel: raw sums (sum x, sum x^2) per (sample, channel) -> atomics,
// then a tiny finalize turns them into (mean, rstd) in place.  Used when one CTA per (sample, 64 ch)
// would leave the GPU idle (stage 1: 24 samples x 1 channel block over 4096 pixels).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_stats_kernel(const uint16_t* __restrict__ x, float* __restrict__ partial, int HW, int C,
                                                           int N, int ppb, int bf16) {
  // partial[z][n][c][2]: per pixel-split partial sums, reduced in fixed order by the finalize kernel
  // (deterministic: no atomics — a 1e-7 wobble in these statistics is amplified ~1000x by the fp16
  // rounding / ReLU / max-pool decisions of the following stages)
  __shared__ float red[2][8][64];
  const int n = blockIdx.y, c0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint16_t* xb = x + static_cast<long long>(n) * HW * C + c0 + lane * 2;
  const int p_end = min(HW, (static_cast<int>(blockIdx.z) + 1) * ppb);
  float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int p = blockIdx.z * ppb + warp; p < p_end; p += 8) {
    const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(xb + static_cast<long long>(p) * C), bf16);
    a0 += f.x; a1 += f.y; q0 += f.x * f.x; q1 += f.y * f.y;
  }
  red[0][warp][lane * 2] = a0; red[0][warp][lane * 2 + 1] = a1;
  red[1][warp][lane * 2] = q0; red[1][warp][lane * 2 + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[0][k][threadIdx.x]; q += red[1][k][threadIdx.x]; }
    float* dst = partial + ((static_cast<long long>(blockIdx.z) * N + n) * C + c0 + threadIdx.x) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}
__global__ void inorm_finalize_kernel(const float* __restrict__ partial, float* __restrict__ mean, float* __restrict__ rstd, int total,
                                      int split, int HW, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float a = 0.f, q = 0.f;
  for (int z = 0; z < split; ++z) {
    a += partial[(static_cast<long long>(z) * total + i) * 2];
    q += partial[(static_cast<long long>(z) * total + i) * 2 + 1];
  }
  const float m = a / HW;
  const float var = fmaxf(q / HW - m * m, 0.f);
  mean[i] = m;
  rstd[i] = rsqrtf(var + eps);
}
// normalise + ReLU + MaxPool2d(2), pixel-parallel (reads the finalized mean / rstd)
__global__ void __launch_bounds__(256) inorm_apply_pool_kernel(const uint16_t* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, uint16_t* __restrict__ y16,
                                                                int H, int W, int C, int ppb, int bf16) {
  const int n = blockIdx.y, c0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ca = c0 + lane * 2;
  const float ma = mean[static_cast<long long>(n) * C + ca], mb = mean[static_cast<long long>(n) * C + ca + 1];
  const float ra = rstd[static_cast<long long>(n) * C + ca], rb = rstd[static_cast<long long>(n) * C + ca + 1];
  const int OH = H / 2, OW = W / 2;
  const uint16_t* xb = x + static_cast<long long>(n) * H * W * C + ca;
  uint16_t* yb = y16 + static_cast<long long>(n) * OH * OW * C + ca;
  const int p_end = min(OH * OW, (static_cast<int>(blockIdx.z) + 1) * ppb);
  for (int p = blockIdx.z * ppb + warp; p < p_end; p += 8) {
    const int oy = p / OW, ox = p % OW;
    const uint16_t* q = xb + (static_cast<long long>(2 * oy) * W + 2 * ox) * C;
    const float2 f0 = unpack2(*reinterpret_cast<const uint32_t*>(q), bf16);
    const float2 f1 = unpack2(*reinterpret_cast<const uint32_t*>(q + C), bf16);
    const float2 f2 = unpack2(*reinterpret_cast<const uint32_t*>(q + static_cast<long long>(W) * C), bf16);
    const float2 f3 = unpack2(*reinterpret_cast<const uint32_t*>(q + static_cast<long long>(W) * C + C), bf16);
    const float va = fmaxf(fmaxf(f0.x, f1.x), fmaxf(f2.x, f3.x)), vb = fmaxf(fmaxf(f0.y, f1.y), fmaxf(f2.y, f3.y));
    *reinterpret_cast<uint32_t*>(yb + static_cast<long long>(p) * C) = pack2(fmaxf((va - ma) * ra, 0.f), fmaxf((vb - mb) * rb, 0.f), bf16);
  }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm2d (no affine, eps, biased var) + ReLU + {MaxPool2d(2) | global average pool},
// NHWC 16-bit in.  One CTA per (sample, 64-channel block).
// replaces: decoder_proj{1..4}[1:4] (models_mae_cross.py:49-51,55-57,61-63,67-69).
// mode 0: y16 [N][H/2][W/2][C];  mode 1: y32 [N][C] (+ optional 16-bit copy y16 [N][C])
// Also stores mean/rstd [N][C] for the backward pass when asked.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_relu_pool_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y16,
                                                               float* __restrict__ y32, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, int H, int W, int C,
                                                               float eps, int mode, int bf16) {
  __shared__ float red[2][8][64];
  __shared__ float s_mean[64], s_rstd[64];
  const int n = blockIdx.y, c0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = H * W;
  const uint16_t* xb = x + static_cast<long long>(n) * HW * C + c0 + lane * 2;
  float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
  for (int p = warp; p < HW; p += 8) {
    const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(xb + static_cast<long long>(p) * C), bf16);
    s1a += f.x; s1b += f.y; s2a += f.x * f.x; s2b += f.y * f.y;
  }
  red[0][warp][lane * 2] = s1a; red[0][warp][lane * 2 + 1] = s1b;
  red[1][warp][lane * 2] = s2a; red[1][warp][lane * 2 + 1] = s2b;
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[0][k][threadIdx.x]; q += red[1][k][threadIdx.x]; }
    const float mean = a / HW;
    const float var = fmaxf(q / HW - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rstd;
    if (mean_out) mean_out[static_cast<long long>(n) * C + c0 + threadIdx.x] = mean;
    if (rstd_out) rstd_out[static_cast<long long>(n) * C + c0 + threadIdx.x] = rstd;
  }
  __syncthreads();
  const float ma = s_mean[lane * 2], mb = s_mean[lane * 2 + 1], ra = s_rstd[lane * 2], rb = s_rstd[lane * 2 + 1];
  if (mode == 0) {
    const int OH = H / 2, OW = W / 2;
    uint16_t* yb = y16 + static_cast<long long>(n) * OH * OW * C + c0 + lane * 2;
    for (int p = warp; p < OH * OW; p += 8) {
      const int oy = p / OW, ox = p % OW;
      const uint16_t* q = xb + (static_cast<long long>(2 * oy) * W + 2 * ox) * C;
      const float2 f0 = unpack2(*reinterpret_cast<const uint32_t*>(q), bf16);
      const float2 f1 = unpack2(*reinterpret_cast<const uint32_t*>(q + C), bf16);
      const float2 f2 = unpack2(*reinterpret_cast<const uint32_t*>(q + static_cast<long long>(W) * C), bf16);
      const float2 f3 = unpack2(*reinterpret_cast<const uint32_t*>(q + static_cast<long long>(W) * C + C), bf16);
      const float va = fmaxf(fmaxf(f0.x, f1.x), fmaxf(f2.x, f3.x)), vb = fmaxf(fmaxf(f0.y, f1.y), fmaxf(f2.y, f3.y));
      *reinterpret_cast<uint32_t*>(yb + static_cast<long long>(p) * C) =
          pack2(fmaxf((va - ma) * ra, 0.f), fmaxf((vb - mb) * rb, 0.f), bf16);
    }
  } else {
    float aa = 0.f, ab = 0.f;
    for (int p = warp; p < HW; p += 8) {
      const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(xb + static_cast<long long>(p) * C), bf16);
      aa += fmaxf((f.x - ma) * ra, 0.f);
      ab += fmaxf((f.y - mb) * rb, 0.f);
    }
    __syncthreads();
    red[0][warp][lane * 2] = aa; red[0][warp][lane * 2 + 1] = ab;
    __syncthreads();
    if (threadIdx.x < 64) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[0][k][threadIdx.x];
      a /= HW;
      const long long o = static_cast<long long>(n) * C + c0 + threadIdx.x;
      if (y32) y32[o] = a;
      if (y16) y16[o] = static_cast<uint16_t>(pack2(a, 0.f, bf16) & 0xffffu);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Cross-attention core (tiny K/V): per token, per head: softmax_s(q.k_s * scale) . v_s
// replaces: CrossAttention.forward lines models_crossvit.py:122-126 (attn, softmax, attn @ v).
// q16 [M][D] 16-bit (wq output), k32/v32 [B][S][D] fp32 (wk/wv outputs), out16 [M][D] 16-bit,
// probs [M][Hh][S] fp32 (optional, saved for backward).  dh == 32, D % 512 == 0, S <= 8.
// One warp per token: lane l owns channels [16l,16l+16) of each 512-chunk (= half a head).
// ------------------------------------------------------------------------------------------
constexpr int kMaxShots = 16;     // forward: evaluation passes every annotated exemplar box (FSC_test_cross(few-shot).py:261); training uses <= 3
__global__ void __launch_bounds__(256) cross_attn_core_kernel(const uint16_t* __restrict__ q16, const float* __restrict__ k32,
                                                               const float* __restrict__ v32, uint16_t* __restrict__ out16,
                                                               float* __restrict__ probs, int L, int S, int D, float scale,
                                                               int tokens_per_block, int bf16, long long kv_bstride) {
  extern __shared__ float skv[];  // [2][S][D]
  const int tok0 = blockIdx.x * tokens_per_block;
  const int b = tok0 / L;
  float* sk = skv;
  float* sv = skv + S * D;
  // K/V are stored channel-permuted inside every 512-channel chunk: channel ch = lane*16 + j lives at j*32 + lane, so the
  // 32 lanes of a warp (16 contiguous channels each) read 32 consecutive floats for a fixed j — the natural [s][ch]
  // layout makes every such read a 16-way bank conflict.
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int s = i / D, ch = i - s * D;
    const int pi = s * D + (ch & ~511) + ((ch & 15) << 5) + ((ch & 511) >> 4);
    sk[pi] = k32[b * kv_bstride + i];
    sv[pi] = v32[b * kv_bstride + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Hh = D / 32;
  for (int t = warp; t < tokens_per_block; t += 8) {
    const long long tok = tok0 + t;
    for (int c0 = 0; c0 < D; c0 += 512) {
      const int ch = c0 + lane * 16;
      float q[16];
      {
        const uint4 u0 = *reinterpret_cast<const uint4*>(q16 + tok * D + ch);
        const uint4 u1 = *reinterpret_cast<const uint4*>(q16 + tok * D + ch + 8);
        float a[8], c[8];
        unpack8(u0, a, bf16);
        unpack8(u1, c, bf16);
#pragma unroll
        for (int j = 0; j < 8; ++j) { q[j] = a[j]; q[8 + j] = c[j]; }
      }
      float sc[kMaxShots];
      float mx = -INFINITY;
#pragma unroll
      for (int s = 0; s < kMaxShots; ++s) {
        if (s < S) {
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) d += q[j] * sk[s * D + c0 + j * 32 + lane];
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          sc[s] = d * scale;
          mx = fmaxf(mx, sc[s]);
        }
      }
      float den = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxShots; ++s)
        if (s < S) { sc[s] = __expf(sc[s] - mx); den += sc[s]; }
      const float inv = 1.f / den;
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxShots; ++s) {
        if (s < S) {
          const float pr = sc[s] * inv;
          if (probs && (lane & 1) == 0) probs[(tok * Hh + (ch >> 5)) * S + s] = pr;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] += pr * sv[s * D + c0 + j * 32 + lane];
        }
      }
      float a[8], c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = o[j]; c[j] = o[8 + j]; }
      *reinterpret_cast<uint4*>(out16 + tok * D + ch) = pack8(a, bf16);
      *reinterpret_cast<uint4*>(out16 + tok * D + ch + 8) = pack8(c, bf16);
    }
  }
}


// ------------------------------------------------------------------------------------------
// MAE pre-training glue (models_mae_noct.py:110-198)
// ------------------------------------------------------------------------------------------
// dst[b][j][:] = src[b][idx[b][j]][:]   rows of row_bytes (multiple of 16) — random_masking's torch.gather
// (:124) and, with the shuffle indices, its backward.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx, uint4* __restrict__ dst,
                                   int n_src, int n_dst, int vec_per_row, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = i % vec_per_row;
  const long long r = i / vec_per_row;
  const int j = r % n_dst;
  const long long b = r / n_dst;
  const long long s = idx[b * n_dst + j];
  dst[i] = src[(b * n_src + s) * vec_per_row + v];
}

// decoder input of the MAE (:158-165): x[b][l] = (ids_restore[b][l] < Lk ? xk[b][ids_restore[b][l]] : mask_token) + pos[l]
__global__ void mae_unshuffle_kernel(const float4* __restrict__ xk, const long long* __restrict__ ids_restore,
                                     const float4* __restrict__ mask_token, const float4* __restrict__ pos, float4* __restrict__ out,
                                     int L, int Lk, int vec_per_row, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = i % vec_per_row;
  const long long r = i / vec_per_row;
  const int l = r % L;
  const long long b = r / L;
  const long long s = ids_restore[b * L + l];
  const float4 a = s < Lk ? xk[(b * Lk + s) * vec_per_row + v] : mask_token[v];
  const float4 p = pos[static_cast<long long>(l) * vec_per_row + v];
  out[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

// Pixel-reconstruction loss (:177-198): target = patchify(imgs) (per patch: (py, px, c) order, optionally normalised
// per patch with the unbiased variance), loss = mean over all patches of mean((pred - target)^2).
// One warp per patch: accumulates the loss (atomic, fp32) and writes dpred = 2 (pred - target) / (N L P) (fp32).
__global__ void __launch_bounds__(256) mae_loss_kernel(const float* __restrict__ pred, const void* __restrict__ img, int dtype,
                                                        long long sb, long long sc, long long sh, long long sw,
                                                        float* __restrict__ loss, float* __restrict__ dpred, int NL, int gw, int gh,
                                                        int P, int C, int norm_pix) {
  const int patch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (patch >= NL) return;
  const int px = patch % gw, py = (patch / gw) % gh, b = patch / (gw * gh);
  const int D = P * P * C;
  float mean = 0.f, inv = 1.f;
  if (norm_pix) {
    float s = 0.f, q = 0.f;
    for (int e = lane; e < D; e += 32) {
      const int c = e % C, qx = (e / C) % P, qy = e / (C * P);
      const float t = load_any(img, b * sb + c * sc + static_cast<long long>(py * P + qy) * sh + static_cast<long long>(px * P + qx) * sw, dtype);
      s += t; q += t * t;
    }
    s = warp_sum(s); q = warp_sum(q);
    mean = s / D;
    const float var = fmaxf((q - s * mean) / (D - 1), 0.f);   // torch.var default: unbiased
    inv = rsqrtf(var + 1e-6f);
  }
  const float gscale = 2.f / (static_cast<float>(NL) * D);
  float acc = 0.f;
  for (int e = lane; e < D; e += 32) {
    const int c = e % C, qx = (e / C) % P, qy = e / (C * P);
    float t = load_any(img, b * sb + c * sc + static_cast<long long>(py * P + qy) * sh + static_cast<long long>(px * P + qx) * sw, dtype);
    t = (t - mean) * inv;
    const float d = pred[static_cast<long long>(patch) * D + e] - t;
    acc += d * d;
    if (dpred) dpred[static_cast<long long>(patch) * D + e] = d * gscale;
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(loss, acc / (static_cast<float>(NL) * D));
}

// dst16[i] = 16-bit( src[i] * (*scale_ptr) ) : gradient recast with the (device-resident) upstream loss-scale
__global__ void cast_scaled_kernel(const float* __restrict__ src, const float* __restrict__ scale_ptr, uint16_t* __restrict__ dst,
                                   long long n, int bf16) {
  const float scale = *scale_ptr;
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
    uint4 o;
    o.x = pack2(a.x * scale, a.y * scale, bf16);
    o.y = pack2(a.z * scale, a.w * scale, bf16);
    o.z = pack2(b.x * scale, b.y * scale, bf16);
    o.w = pack2(b.z * scale, b.w * scale, bf16);
    *reinterpret_cast<uint4*>(dst + i) = o;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = static_cast<uint16_t>(pack2(src[j] * scale, 0.f, bf16) & 0xffffu);
  }
}


// ------------------------------------------------------------------------------------------
// Sliding-window blend of the evaluation path (demo.py:124-160, FSC_test_cross(few-shot).py:322-349):
// windows of width Wwin at x = starts[i] are visited left to right; a column already covered by an earlier
// window takes the average of the running map and the new window, a new column takes the window's value.
// One thread per output pixel replays that recurrence over the (few) windows that cover its column.
// ------------------------------------------------------------------------------------------
__global__ void window_blend_kernel(const void* __restrict__ outs, int dtype, const int* __restrict__ starts, int nw, int H, int Wwin,
                                    int W, float* __restrict__ density) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(H) * W) return;
  const int x = idx % W, y = idx / W;
  float v = 0.f;
  bool covered = false;
  for (int i = 0; i < nw; ++i) {
    const int s = starts[i];
    if (x >= s && x < s + Wwin) {
      const float o = load_any(outs, (static_cast<long long>(i) * H + y) * Wwin + (x - s), dtype);
      v = covered ? 0.5f * v + 0.5f * o : o;
      covered = true;
    }
  }
  density[idx] = v;
}

}  // namespace
}  // namespace countr

using namespace countr;

extern "C" int countr_cast_f32_to_16(const float* src, void* dst, int64_t n, float scale, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(src && dst && n > 0, "bad arguments");
  COUNTR_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0, "pointers must be 16-byte aligned");
  const long long threads = (n + 7) / 8;
  cast_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(src, reinterpret_cast<uint16_t*>(dst), n, scale, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_cast_transpose_f32_to_16(const float* src, void* dst, int R, int C, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(src && dst && R > 0 && C > 0, "bad arguments");
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  cast_transpose_kernel<<<grid, block, 0, stream>>>(src, reinterpret_cast<uint16_t*>(dst), R, C, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_patchify(const void* img, int dtype, int64_t sb, int64_t sc, int64_t sh, int64_t sw, void* out,
                               int B, int C, int H, int W, int P, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(img && out, "null pointer");
  COUNTR_REQUIRE(dtype >= 0 && dtype <= 2, "image dtype code %d", dtype);
  COUNTR_REQUIRE(P > 0 && H >= P && W >= P, "patch size %d vs image %dx%d", P, H, W);
  if (P % 8 != 0 || H % P != 0 || W % P != 0) {
    // generic path: row pitch = C*P*P rounded up to a multiple of 8, zero padded (see the header)
    const int ld = (C * P * P + 7) & ~7;
    const long long n = static_cast<long long>(B) * (H / P) * (W / P) * ld;
    patchify_generic_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(img, dtype, sb, sc, sh, sw,
                                                                                     reinterpret_cast<uint16_t*>(out), B, C, H, W, P, ld, bf16);
    COUNTR_CHECK_CUDA(cudaGetLastError());
    return COUNTR_OK;
  }
  const long long total = static_cast<long long>(B) * (H / P) * (W / P) * C * P * (P / 8);
  patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(img, dtype, sb, sc, sh, sw,
                                                                                 reinterpret_cast<uint16_t*>(out), B, C, H, W, P, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_conv_weight_pack(const float* w, void* out, int Cout, int Cin, int mode, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(w && out && Cout > 0 && Cin > 0 && (mode == 0 || mode == 1), "bad arguments");
  const long long n = static_cast<long long>(Cout) * Cin * 9;
  conv_weight_pack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(w, reinterpret_cast<uint16_t*>(out), Cout, Cin, mode, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_gn_relu_upsample2x(const void* x, const double* stats, const float* gamma, const float* beta, void* y,
                                         int B, int H, int W, int C, int G, float eps, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(x && stats && gamma && beta && y, "null pointer");
  COUNTR_REQUIRE(C % 8 == 0 && C % G == 0 && C <= 4096, "bad channel count %d / groups %d", C, G);
  const long long total = 1ll * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 148ll * 8 * 4;
  if (blocks * B > cap) blocks = (cap + B - 1) / B;
  if (blocks < 1) blocks = 1;
  static const int staged_env = [] { const char* e = getenv("COUNTR_GN_UP2_STAGED"); return e ? atoi(e) : 1; }();
  if (staged_env && C == UT_C && B <= 65535 && H * W >= 1024) {   // smaller maps: too few tiles to pipeline (24^2: 12.3 vs 11.0 us)
    static PerDeviceOnce attr_once;
    if (attr_once.need())
      COUNTR_CHECK_CUDA(cudaFuncSetAttribute(gn_relu_up2_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UT_SMEM));
    const int ntiles = ((W + UT_X - 1) / UT_X) * ((H + UT_Y - 1) / UT_Y);
    const int bpi = std::max(1, std::min(ntiles, num_sms() / B));       // persistent blocks per image
    gn_relu_up2_staged_kernel<<<dim3(bpi, B), 512, UT_SMEM, stream>>>(reinterpret_cast<const uint16_t*>(x), stats, gamma, beta,
                                                                       reinterpret_cast<uint16_t*>(y), H, W, G, eps, bf16);
    COUNTR_CHECK_CUDA(cudaGetLastError());
    return COUNTR_OK;
  }
  const int vecs_ = C / 8;
  if (vecs_ <= 256 && 256 % vecs_ == 0 && B <= 65535) {
    const int ppb = 256 / vecs_;                               // pixel columns per block
    int R = 8;                                                 // rows per strip: keep >= ~4 blocks per SM in flight
    while (R > 2 && static_cast<long long>((W + ppb - 1) / ppb) * ((H + R - 1) / R) * B < 4 * 148) R >>= 1;
    dim3 grid2((W + ppb - 1) / ppb, (H + R - 1) / R, B);
    gn_relu_up2_rows_kernel<<<grid2, 256, 2 * C * sizeof(float), stream>>>(reinterpret_cast<const uint16_t*>(x), stats, gamma, beta,
                                                                          reinterpret_cast<uint16_t*>(y), H, W, C, G, eps, R, bf16);
    COUNTR_CHECK_CUDA(cudaGetLastError());
    return COUNTR_OK;
  }
  dim3 grid(static_cast<unsigned>(blocks), B);
  gn_relu_up2_kernel<<<grid, 256, 2 * C * sizeof(float), stream>>>(reinterpret_cast<const uint16_t*>(x), stats, gamma, beta,
                                                                   reinterpret_cast<uint16_t*>(y), H, W, C, G, eps, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_gn_relu_conv1x1(const void* x, const double* stats, const float* gamma, const float* beta, const float* w,
                                      const float* bias, float* out, int B, int HW, int C, int G, float eps, int bf16,
                                      countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(x && stats && gamma && beta && w && bias && out, "null pointer");
  COUNTR_REQUIRE(C % 8 == 0 && C % G == 0 && C <= 4096, "bad channel count %d / groups %d", C, G);
  static const int staged_env = [] { const char* e = getenv("COUNTR_GN_DOT_STAGED"); return e ? atoi(e) : 1; }();
  if (staged_env && C == UT_C && B <= 65535 && HW >= 4096) {
    static PerDeviceOnce attr_once;
    if (attr_once.need())
      COUNTR_CHECK_CUDA(cudaFuncSetAttribute(gn_relu_dot_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
    const int ntiles = (HW + DT_PX - 1) / DT_PX;
    const int bpi = std::max(1, std::min(ntiles, num_sms() / B));
    gn_relu_dot_staged_kernel<<<dim3(bpi, B), 512, DT_SMEM, stream>>>(reinterpret_cast<const uint16_t*>(x), stats, gamma, beta, w, bias, out,
                                                                       HW, G, eps, bf16);
    COUNTR_CHECK_CUDA(cudaGetLastError());
    return COUNTR_OK;
  }
  long long blocks = (HW + 31) / 32;
  const long long cap = 148ll * 8 * 2;
  if (blocks * B > cap) blocks = (cap + B - 1) / B;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), B);
  gn_relu_dot_kernel<<<grid, 256, 3 * C * sizeof(float), stream>>>(reinterpret_cast<const uint16_t*>(x), stats, gamma, beta, w, bias,
                                                                   out, HW, C, G, eps, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_upsample2x_f32(const float* x, void* y, int B, int H, int W, int out_dtype, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(x && y && out_dtype >= 0 && out_dtype <= 2, "bad arguments");
  const long long total = 4ll * B * H * W;
  up2_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, y, B, H, W, out_dtype);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_exemplar_conv1(const void* boxes, int dtype, int64_t sB, int64_t sK, int64_t sC, int64_t sH, int64_t sW,
                                     const float* w, const float* bias, void* out, int B, int S, int HW, int Cout, int bf16,
                                     countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(boxes && w && bias && out, "null pointer");
  COUNTR_REQUIRE(Cout == 64 && HW == 64 && B > 0 && S > 0, "exemplar conv1 expects Cout=64 and 64x64 crops (got %d, %d)", Cout, HW);
  dim3 grid(HW / 2, B * S);
  exemplar_conv1_kernel<<<grid, 256, 0, stream>>>(boxes, dtype, sB, sK, sC, sH, sW, w, bias, reinterpret_cast<uint16_t*>(out), S, HW,
                                                  Cout, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_inorm_relu_pool(const void* x, void* y16, float* y32, float* mean, float* rstd, float* scratch, int N, int H,
                                      int W, int C, float eps, int mode, int bf16, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(x && (y16 || y32), "null pointer");
  COUNTR_REQUIRE(C % 64 == 0 && (mode == 1 || (H % 2 == 0 && W % 2 == 0 && y16)), "bad shape C=%d H=%d W=%d mode=%d", C, H, W, mode);
  if (mode == 0 && mean != nullptr && rstd != nullptr && scratch != nullptr && H * W >= 256) {
    // pixel-parallel path: per-split partial sums -> fixed-order finalize -> apply (3 launches, each filling the GPU)
    const int HW = H * W, cblocks = C / 64;
    int split = (2 * 148 + cblocks * N - 1) / (cblocks * N);
    if (split > HW / 64) split = HW / 64;
    if (split > 32) split = 32;          // scratch holds at most 32 partials per (sample, channel)
    if (split < 1) split = 1;
    const int ppb = (HW + split - 1) / split;
    split = (HW + ppb - 1) / ppb;
    inorm_stats_kernel<<<dim3(cblocks, N, split), 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(x), scratch, HW, C, N, ppb, bf16);
    inorm_finalize_kernel<<<(N * C + 255) / 256, 256, 0, stream>>>(scratch, mean, rstd, N * C, split, HW, eps);
    const int OHW = HW / 4;
    const int ppb2 = (OHW + split - 1) / split;
    inorm_apply_pool_kernel<<<dim3(cblocks, N, (OHW + ppb2 - 1) / ppb2), 256, 0, stream>>>(
        reinterpret_cast<const uint16_t*>(x), mean, rstd, reinterpret_cast<uint16_t*>(y16), H, W, C, ppb2, bf16);
    COUNTR_CHECK_CUDA(cudaGetLastError());
    return COUNTR_OK;
  }
  dim3 grid(C / 64, N);
  inorm_relu_pool_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(x), reinterpret_cast<uint16_t*>(y16), y32, mean,
                                                   rstd, H, W, C, eps, mode, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_cross_attn_core(const void* q16, const float* k32, const float* v32, void* out16, float* probs, int B, int L,
                                      int S, int D, int dh, float scale, int bf16, int kv_broadcast, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(q16 && k32 && v32 && out16, "null pointer");
  COUNTR_REQUIRE(dh == 32 && D % 512 == 0, "cross-attention core supports head_dim 32 and D %% 512 == 0 (got dh=%d D=%d)", dh, D);
  COUNTR_REQUIRE(S >= 1 && S <= kMaxShots, "exemplar count %d outside [1,%d]", S, kMaxShots);
  int tpb = 32;
  while (L % tpb) tpb >>= 1;
  const int blocks = B * L / tpb;
  const size_t smem = 2ull * S * D * sizeof(float);
  COUNTR_REQUIRE(smem <= 160 * 1024, "exemplar tokens do not fit in shared memory (S=%d D=%d)", S, D);
  static PerDeviceOnce attr_once;
  if (attr_once.need())
    COUNTR_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  cross_attn_core_kernel<<<blocks, 256, smem, stream>>>(reinterpret_cast<const uint16_t*>(q16), k32, v32,
                                                       reinterpret_cast<uint16_t*>(out16), probs, L, S, D, scale, tpb, bf16,
                                                       kv_broadcast ? 0ll : static_cast<long long>(S) * D);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_gather_rows(const void* src, const int64_t* idx, void* dst, int B, int n_src, int n_dst, int row_bytes,
                                  countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(src && idx && dst && row_bytes % 16 == 0 && B > 0 && n_dst > 0, "bad arguments");
  const int vpr = row_bytes / 16;
  const long long total = static_cast<long long>(B) * n_dst * vpr;
  gather_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<const long long*>(idx), reinterpret_cast<uint4*>(dst), n_src, n_dst, vpr, total);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_mae_unshuffle(const float* xk, const int64_t* ids_restore, const float* mask_token, const float* pos, float* out,
                                    int B, int L, int Lk, int D, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(xk && ids_restore && mask_token && pos && out && D % 4 == 0, "bad arguments");
  const int vpr = D / 4;
  const long long total = static_cast<long long>(B) * L * vpr;
  mae_unshuffle_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(xk), reinterpret_cast<const long long*>(ids_restore), reinterpret_cast<const float4*>(mask_token),
      reinterpret_cast<const float4*>(pos), reinterpret_cast<float4*>(out), L, Lk, vpr, total);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_mae_loss(const float* pred, const void* img, int dtype, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                               float* loss, float* dpred, int B, int C, int H, int W, int P, int norm_pix, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(pred && img && loss && H % P == 0 && W % P == 0 && dtype >= 0 && dtype <= 2, "bad arguments");
  const int NL = B * (H / P) * (W / P);
  COUNTR_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  mae_loss_kernel<<<(NL + 7) / 8, 256, 0, stream>>>(pred, img, dtype, sb, sc, sh, sw, loss, dpred, NL, W / P, H / P, P, C, norm_pix);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_cast_scaled_f32_to_16(const float* src, const float* scale_ptr, void* dst, int64_t n, int bf16,
                                            countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(src && dst && scale_ptr && n > 0, "bad arguments");
  const long long threads = (n + 7) / 8;
  cast_scaled_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(src, scale_ptr, reinterpret_cast<uint16_t*>(dst), n, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_window_blend(const void* outs, int dtype, const int32_t* starts, int nw, int H, int Wwin, int W, float* density,
                                   countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(outs && starts && density && nw > 0 && dtype >= 0 && dtype <= 2, "bad arguments");
  const long long total = static_cast<long long>(H) * W;
  window_blend_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(outs, dtype, starts, nw, H, Wwin, W, density);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

// ------------------------------------------------------------------------------------------
// Batched refresh of the 16-bit GEMM / conv operand copies of the fp32 master weights: ONE launch per step instead of one
// cast / transpose / pack launch per tensor (~50 launches of 2-3 us each after every optimizer step).
//   entry = {src fp32, dst 16-bit, kind, R, C}:   kind 0  dst[i] = src[i]                  (R*C elements)
//                                                 kind 1  dst[c][r] = src[r][c]            (fp32 [R][C] -> [C][R])
//                                                 kind 2  conv pack mode 0 (Cout = R, Cin = C): dst[co][tap][ci] = w[co][ci][tap]
//                                                 kind 3  conv pack mode 1:                     dst[ci][tap][co] = w[co][ci][8 - tap]
//   blk_prefix[e] = first block of entry e (n_entries + 1 values); 256 threads per block:
//   kind 0: 2048 elements per block; kind 1: one 64 x 64 tile per block; kind 2: one (co, 128 ci) slab per block;
//   kind 3: one 64 x 64 tile of w viewed as [Cout][Cin*9] per block  (countr_weight_refresh_blocks gives the count).
// ------------------------------------------------------------------------------------------
namespace countr {
namespace {
struct WREntry {
  const float* src;
  uint16_t* dst;
  long long kind, R, C, pad;
};

// blocks of one entry (shared by the kernel's decode and by countr_weight_refresh_blocks, which the host plan calls)
__host__ __device__ inline long long wr_blocks(long long kind, long long R, long long C) {
  if (kind == 0) return (R * C + 2047) / 2048;
  if (kind == 1) return ((R + 63) / 64) * ((C + 63) / 64);
  if (kind == 2) return R * ((C + 127) / 128);
  return ((R + 63) / 64) * ((9 * C + 63) / 64);
}

// 64 x 64 transpose tile of src viewed as [R][CC]: coalesced 128-byte row reads, 128-byte packed 16-bit row writes
// (two consecutive r per thread).  dst row of source column c: dst_row(c) * R.
template <bool kConvFlip>
__device__ __forceinline__ void wr_transpose_tile(const float* __restrict__ src, uint16_t* __restrict__ dst, int R, int CC, int lb, int bf16,
                                                  float (*tile)[65]) {
  const int tiles_c = (CC + 63) / 64;
  const int c0 = (lb % tiles_c) * 64, r0 = (lb / tiles_c) * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i;
    const float* sp = src + static_cast<size_t>(r) * CC + c0;
    tile[i][tx] = (r < R && c0 + tx < CC) ? sp[tx] : 0.f;
    tile[i][tx + 32] = (r < R && c0 + tx + 32 < CC) ? sp[tx + 32] : 0.f;
  }
  __syncthreads();
  const bool even = (R & 1) == 0;
#pragma unroll
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;
    if (c >= CC || r >= R) continue;
    size_t row = static_cast<size_t>(c);
    if (kConvFlip) {
      const int ci = c / 9, tap = c - ci * 9;
      row = static_cast<size_t>(ci) * 9 + (8 - tap);
    }
    uint16_t* dp = dst + row * R + r;
    if (even) {
      *reinterpret_cast<uint32_t*>(dp) = pack2(tile[2 * tx][i], tile[2 * tx + 1][i], bf16);
    } else {
      dp[0] = static_cast<uint16_t>(pack2(tile[2 * tx][i], 0.f, bf16) & 0xffffu);
      if (r + 1 < R) dp[1] = static_cast<uint16_t>(pack2(tile[2 * tx + 1][i], 0.f, bf16) & 0xffffu);
    }
  }
}

__global__ void __launch_bounds__(256) weight_refresh_kernel(const WREntry* __restrict__ entries, const int* __restrict__ blk_prefix,
                                                             int n_entries, int bf16) {
  __shared__ float tile[64][65];
  // entry of this block: last e with blk_prefix[e] <= blockIdx.x (every thread runs the same 7-step search: broadcast loads)
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (blk_prefix[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const WREntry en = entries[lo];
  const int lb = blockIdx.x - blk_prefix[lo];      // block index inside the entry
  const int R = static_cast<int>(en.R), C = static_cast<int>(en.C);
  if (en.kind == 0) {
    const long long n = static_cast<long long>(R) * C;
    const long long i = (static_cast<long long>(lb) * 256 + threadIdx.x) * 8;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(en.src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(en.dst) & 15u) == 0) {
      const float4 a = *reinterpret_cast<const float4*>(en.src + i), b = *reinterpret_cast<const float4*>(en.src + i + 4);
      uint4 o;
      o.x = pack2(a.x, a.y, bf16);
      o.y = pack2(a.z, a.w, bf16);
      o.z = pack2(b.x, b.y, bf16);
      o.w = pack2(b.z, b.w, bf16);
      *reinterpret_cast<uint4*>(en.dst + i) = o;
    } else {
      for (long long j = i; j < n && j < i + 8; ++j) en.dst[j] = static_cast<uint16_t>(pack2(en.src[j], 0.f, bf16) & 0xffffu);
    }
  } else if (en.kind == 1) {
    wr_transpose_tile<false>(en.src, en.dst, R, C, lb, bf16, tile);
  } else if (en.kind == 2) {
    // mode 0: dst[co][tap][ci] = w[co][ci][tap].  Block = (co, 128 input channels): 1152 contiguous floats in, 9 rows of
    // 128 contiguous 16-bit values out (Cout = R, Cin = C).
    float* cw = &tile[0][0];
    const int cblocks = (C + 127) / 128;
    const int co = lb / cblocks, ci0 = (lb % cblocks) * 128;
    const int nci = min(128, C - ci0);
    const float* sp = en.src + (static_cast<long long>(co) * C + ci0) * 9;
    for (int i = threadIdx.x; i < nci * 9; i += 256) cw[i] = sp[i];
    __syncthreads();
    uint16_t* dp = en.dst + static_cast<long long>(co) * 9 * C + ci0;
    for (int i = threadIdx.x; i < nci * 9; i += 256) {
      const int tap = i / nci, ci = i - tap * nci;
      dp[static_cast<long long>(tap) * C + ci] = static_cast<uint16_t>(pack2(cw[ci * 9 + tap], 0.f, bf16) & 0xffffu);
    }
  } else {
    // mode 1: dst[ci][tap][co] = w[co][ci][8 - tap]: the transpose of w viewed as [Cout][Cin*9], with the 9 taps of every
    // input channel written in reverse order
    wr_transpose_tile<true>(en.src, en.dst, R, C * 9, lb, bf16, tile);
  }
}
}  // namespace
}  // namespace countr

extern "C" int64_t countr_weight_refresh_blocks(int kind, int64_t R, int64_t C) {
  return (kind < 0 || kind > 3 || R < 0 || C < 0) ? -1 : countr::wr_blocks(kind, R, C);
}

extern "C" int countr_weight_refresh(const void* entries, const int32_t* blk_prefix, int n_entries, int total_blocks, int bf16,
                                     countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(entries && blk_prefix && n_entries > 0 && total_blocks > 0, "bad arguments");
  weight_refresh_kernel<<<static_cast<unsigned>(total_blocks), 256, 0, stream>>>(reinterpret_cast<const WREntry*>(entries), blk_prefix,
                                                                                n_entries, bf16);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}
