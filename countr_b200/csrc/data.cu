// countr_b200 — GPU input pipeline, first piece: ground-truth density-map synthesis.
//
// replaces (host side, numpy + scipy, one image at a time inside the DataLoader workers):
//   util/FSC147.py:262-273  (ResizeTrainImage, no-augmentation path)
//       resized_density[min(new_H-1, int(dot_y*scale_h))][min(new_W-1, int(dot_x*scale_w))] = 1
//       crop [0:384, start:start+384];  ndimage.gaussian_filter(sigma=(1,1), order=0);  * 60
//   util/FSC147.py:326-331  (ResizeValImage)   same scatter on the 384 x 384 canvas, gaussian_filter(sigma=4, radius=7), * 60
//
// Arithmetic follows scipy.ndimage.gaussian_filter on a float32 array exactly: separable, axis 0 then axis 1, each pass
// accumulated in double with the symmetric-kernel summation order of ni_filters.c (centre tap first, then the tap pairs
// from the outermost inwards), 'reflect' boundary (d c b a | a b c d | d c b a), the intermediate and the result rounded
// to float32.  The (normalised, float64) weights are computed by the caller the way scipy's _gaussian_kernel1d does.
#include "../../include/countr_b200.h"
#include "common.cuh"

#include <algorithm>

namespace countr {
namespace {

// scatter the dots of image b onto the (canvas_h x canvas_w) resized canvas and keep the H x W window at (y0, x0)
__global__ void density_scatter_kernel(const double* __restrict__ dots, const int* __restrict__ counts, float* __restrict__ map,
                                       int n_max, double scale_h, double scale_w, int canvas_h, int canvas_w, int y0, int x0,
                                       int H, int W) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[b] || i >= n_max) return;
  const double dx = dots[(static_cast<long long>(b) * n_max + i) * 2], dy = dots[(static_cast<long long>(b) * n_max + i) * 2 + 1];
  int row = static_cast<int>(dy * scale_h), col = static_cast<int>(dx * scale_w);   // Python int(): truncation toward zero
  row = min(canvas_h - 1, row) - y0;
  col = min(canvas_w - 1, col) - x0;
  if (row < 0 || row >= H || col < 0 || col >= W) return;
  map[(static_cast<long long>(b) * H + row) * W + col] = 1.0f;   // assignment, not accumulation: coincident dots count once
}

__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy 'reflect' (half-sample symmetric); radius < n is required by the caller
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - 1 - i;
  return i;
}

// one 1-D pass along `axis` (0: rows / vertical, 1: columns / horizontal); out = float32(sum in double) * gain
__global__ void gauss_pass_kernel(const float* __restrict__ in, float* __restrict__ out, const double* __restrict__ w, int radius,
                                  int B, int H, int W, int axis, float gain) {
  extern __shared__ double sw[];   // w[0] = centre, w[k] = tap at distance k
  for (int k = threadIdx.x; k <= radius; k += blockDim.x) sw[k] = w[k];
  __syncthreads();
  const long long total = static_cast<long long>(B) * H * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W), y = static_cast<int>((idx / W) % H);
  const float* img = in + (idx / (static_cast<long long>(W) * H)) * H * W;
  double acc;
  if (axis == 0) {
    acc = static_cast<double>(img[static_cast<long long>(y) * W + x]) * sw[0];
    for (int k = radius; k >= 1; --k) {
      const double a = img[static_cast<long long>(reflect_index(y - k, H)) * W + x], c = img[static_cast<long long>(reflect_index(y + k, H)) * W + x];
      acc += (a + c) * sw[k];
    }
  } else {
    acc = static_cast<double>(img[static_cast<long long>(y) * W + x]) * sw[0];
    for (int k = radius; k >= 1; --k) {
      const double a = img[static_cast<long long>(y) * W + reflect_index(x - k, W)], c = img[static_cast<long long>(y) * W + reflect_index(x + k, W)];
      acc += (a + c) * sw[k];
    }
  }
  const float r = static_cast<float>(acc);
  out[idx] = gain == 1.0f ? r : r * gain;
}

}  // namespace
}  // namespace countr

extern "C" int countr_density_from_dots(const double* dots, const int32_t* counts, int B, int n_max, double scale_h, double scale_w,
                                        int canvas_h, int canvas_w, int y0, int x0, int H, int W, const double* weights, int radius,
                                        float gain, float* tmp, float* out, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(dots && counts && weights && tmp && out && tmp != out, "null or aliased pointer");
  COUNTR_REQUIRE(B > 0 && n_max >= 0 && H > 0 && W > 0 && radius >= 0 && radius < H && radius < W && radius <= 512,
                 "bad shape B=%d n_max=%d H=%d W=%d radius=%d", B, n_max, H, W, radius);
  const long long total = static_cast<long long>(B) * H * W;
  COUNTR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * total, stream));
  if (n_max > 0) {
    dim3 grid((n_max + 127) / 128, B);
    density_scatter_kernel<<<grid, 128, 0, stream>>>(dots, counts, out, n_max, scale_h, scale_w, canvas_h, canvas_w, y0, x0, H, W);
  }
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  const size_t smem = sizeof(double) * (radius + 1);
  gauss_pass_kernel<<<blocks, 256, smem, stream>>>(out, tmp, weights, radius, B, H, W, 0, 1.0f);
  gauss_pass_kernel<<<blocks, 256, smem, stream>>>(tmp, out, weights, radius, B, H, W, 1, gain);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

// ------------------------------------------------------------------------------------------
// Exemplar crops: boxes[b][s] = Resize((64, 64))(image[b][:, y1:y2+1, x1:x2+1])     (util/FSC147.py:285-298, 343-351)
// torchvision 0.14.1 (the reference's pin) resizes a float tensor with torch.nn.functional.interpolate(mode="bilinear",
// align_corners=False) and no antialiasing; this is that arithmetic (ATen upsample_bilinear2d: source index
// scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out in float, value = h0 (w0 a + w1 b) + h1 (w0 c + w1 d)).
// ------------------------------------------------------------------------------------------
namespace countr {
namespace {
__global__ void crop_resize_kernel(const float* __restrict__ img, long long sb, long long sc, long long sh, long long sw,
                                   const int* __restrict__ rects, float* __restrict__ out, int B, int S, int C, int H, int W, int OH, int OW) {
  const long long total = static_cast<long long>(B) * S * C * OH * OW;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % OW), oy = static_cast<int>((idx / OW) % OH);
  const long long plane = static_cast<long long>(OH) * OW;
  const int c = static_cast<int>((idx / plane) % C);
  const int s = static_cast<int>((idx / (plane * C)) % S);
  const int b = static_cast<int>(idx / (plane * C * S));
  const int* r = rects + (static_cast<long long>(b) * S + s) * 4;
  const int y1 = max(0, min(H - 1, r[0])), x1 = max(0, min(W - 1, r[1]));
  const int y2 = max(y1, min(H - 1, r[2])), x2 = max(x1, min(W - 1, r[3]));   // python slicing clips at the image border
  const int h = y2 - y1 + 1, w = x2 - x1 + 1;
  const float scale_h = static_cast<float>(h) / OH, scale_w = static_cast<float>(w) / OW;
  float ry = scale_h * (oy + 0.5f) - 0.5f, rx = scale_w * (ox + 0.5f) - 0.5f;
  ry = ry < 0.f ? 0.f : ry;
  rx = rx < 0.f ? 0.f : rx;
  const int iy0 = min(static_cast<int>(ry), h - 1), ix0 = min(static_cast<int>(rx), w - 1);
  const int iy1 = iy0 + (iy0 < h - 1 ? 1 : 0), ix1 = ix0 + (ix0 < w - 1 ? 1 : 0);
  const float ly1 = fminf(fmaxf(ry - iy0, 0.f), 1.f), lx1 = fminf(fmaxf(rx - ix0, 0.f), 1.f);
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* p = img + b * sb + c * sc;
  const float a = p[(y1 + iy0) * sh + (x1 + ix0) * sw], bq = p[(y1 + iy0) * sh + (x1 + ix1) * sw];
  const float cq = p[(y1 + iy1) * sh + (x1 + ix0) * sw], d = p[(y1 + iy1) * sh + (x1 + ix1) * sw];
  out[idx] = ly0 * (lx0 * a + lx1 * bq) + ly1 * (lx0 * cq + lx1 * d);
}

// sum over the inclusive pixel rectangles (y1, x1, y2, x2) of map / divisor, all rectangles added into out[0]
// (test-time normalisation: the density mass under the exemplar boxes, FSC_test_cross(few-shot).py:353-359)
__global__ void __launch_bounds__(256) rect_mass_kernel(const float* __restrict__ map, int H, int W, const int* __restrict__ rects,
                                                        float divisor, float* __restrict__ out) {
  const int* r = rects + blockIdx.x * 4;
  // python slicing: negative / oversized bounds clip at the border
  const int y1 = max(0, r[0]), x1 = max(0, r[1]), y2 = min(H - 1, r[2]), x2 = min(W - 1, r[3]);
  float acc = 0.f;
  if (y2 >= y1 && x2 >= x1) {
    const int w = x2 - x1 + 1, n = (y2 - y1 + 1) * w;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += map[static_cast<long long>(y1 + i / w) * W + x1 + i % w] / divisor;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}
}  // namespace
}  // namespace countr

extern "C" int countr_crop_resize(const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int32_t* rects, float* out,
                                  int B, int S, int C, int H, int W, int out_h, int out_w, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(img && rects && out, "null pointer");
  COUNTR_REQUIRE(B > 0 && S > 0 && C > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "bad shape");
  const long long total = static_cast<long long>(B) * S * C * out_h * out_w;
  crop_resize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(img, sb, sc, sh, sw, rects, out, B, S, C, H, W, out_h,
                                                                                     out_w);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_crop_resize_boxes(const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int32_t* rects,
                                        float* out, int B, int S, int C, int H, int W, int out_hw, countr_stream_t stream_) {
  return countr_crop_resize(img, sb, sc, sh, sw, rects, out, B, S, C, H, W, out_hw, out_hw, stream_);
}

extern "C" int countr_rect_mass(const float* map, int H, int W, const int32_t* rects, int n_rects, float divisor, float* out,
                                countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(map && rects && out && H > 0 && W > 0 && n_rects > 0 && divisor != 0.f, "bad arguments");
  rect_mass_kernel<<<n_rects, 256, 0, stream>>>(map, H, W, rects, divisor, out);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}


// ==========================================================================================
// Training-time augmentations on the device (util/FSC147.py:133-143, 371-374): Gaussian noise + clamp, torchvision's
// ColorJitter (brightness / contrast / saturation / hue in a sampled order, tensor arithmetic of
// torchvision.transforms.functional) and GaussianBlur(kernel_size=(7, 9), sigma ~ U[0.1, 2]), horizontal flip.  The random
// PARAMETERS (factors, order, sigma, flip decision) are the caller's (host RNG streams cannot be reproduced on the device);
// given the same parameters the pixel arithmetic is torchvision's.  Images are fp32 [B][3][H][W] contiguous.
// ==========================================================================================
namespace countr {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// out = clamp(img + N(0, std), 0, 1); counter-based generator (one 64-bit hash per pixel pair, Box-Muller)
__global__ void __launch_bounds__(256) noise_clamp_kernel(const float* __restrict__ img, float* __restrict__ out, long long n, float stddev,
                                                          unsigned long long seed) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i >= n) return;
  const uint64_t h = mix64(seed ^ mix64(static_cast<uint64_t>(i)));
  const float u1 = (static_cast<float>(static_cast<uint32_t>(h)) + 0.5f) * 2.3283064365386963e-10f;          // (0, 1)
  const float u2 = (static_cast<float>(static_cast<uint32_t>(h >> 32)) + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.f * logf(u1)) * stddev;
  float sn, cs;
  sincospif(2.f * u2, &sn, &cs);
  out[i] = fminf(fmaxf(img[i] + r * cs, 0.f), 1.f);
  if (i + 1 < n) out[i + 1] = fminf(fmaxf(img[i + 1] + r * sn, 0.f), 1.f);
}

__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }

// mean[b] += sum of the grayscale image (fp64 accumulation); the caller divides by H*W
__global__ void __launch_bounds__(256) gray_sum_kernel(const float* __restrict__ img, double* __restrict__ sum, int HW) {
  const int b = blockIdx.y;
  const float* p = img + static_cast<long long>(b) * 3 * HW;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
    acc += static_cast<double>(gray_of(p[i], p[HW + i], p[2 * HW + i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(sum + b, t);
  }
}

// one ColorJitter operation on every pixel of image b: op[b] in {0 brightness, 1 contrast, 2 saturation, 3 hue, -1 skip}
__global__ void __launch_bounds__(256) color_op_kernel(float* __restrict__ img, const int* __restrict__ op, const float* __restrict__ factor,
                                                       const double* __restrict__ gray_sum, int HW) {
  const int b = blockIdx.y;
  const int which = op[b];
  if (which < 0) return;
  const float f = factor[b];
  float* p = img + static_cast<long long>(b) * 3 * HW;
  const float mean = which == 1 ? static_cast<float>(gray_sum[b] / HW) : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float r = p[i], g = p[HW + i], bl = p[2 * HW + i];
    if (which == 0) {            // _blend(img, 0, f)
      r = fminf(fmaxf(f * r, 0.f), 1.f); g = fminf(fmaxf(f * g, 0.f), 1.f); bl = fminf(fmaxf(f * bl, 0.f), 1.f);
    } else if (which == 1 || which == 2) {
      const float m = which == 1 ? mean : gray_of(r, g, bl);
      r = fminf(fmaxf(f * r + (1.f - f) * m, 0.f), 1.f);
      g = fminf(fmaxf(f * g + (1.f - f) * m, 0.f), 1.f);
      bl = fminf(fmaxf(f * bl + (1.f - f) * m, 0.f), 1.f);
    } else {                     // hue: _rgb2hsv, h = (h + f) % 1, _hsv2rgb
      const float maxc = fmaxf(r, fmaxf(g, bl)), minc = fminf(r, fminf(g, bl));
      const bool eqc = maxc == minc;
      const float cr = maxc - minc;
      const float sat = cr / (eqc ? 1.f : maxc);
      const float div = eqc ? 1.f : cr;
      const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - bl) / div;
      float h = 0.f;
      if (maxc == r) h = bc - gc;
      else if (maxc == g) h = 2.f + rc - bc;
      else h = 4.f + gc - rc;
      h = fmodf(h / 6.f + 1.f, 1.f);
      h = h + f;
      h = h - floorf(h);         // python % 1.0
      const float v = maxc;
      const float i6 = floorf(h * 6.f);
      const float fr = h * 6.f - i6;
      int ii = static_cast<int>(i6) % 6;
      if (ii < 0) ii += 6;
      const float pp = fminf(fmaxf(v * (1.f - sat), 0.f), 1.f);
      const float qq = fminf(fmaxf(v * (1.f - sat * fr), 0.f), 1.f);
      const float tt = fminf(fmaxf(v * (1.f - sat * (1.f - fr)), 0.f), 1.f);
      switch (ii) {
        case 0: r = v; g = tt; bl = pp; break;
        case 1: r = qq; g = v; bl = pp; break;
        case 2: r = pp; g = v; bl = tt; break;
        case 3: r = pp; g = qq; bl = v; break;
        case 4: r = tt; g = pp; bl = v; break;
        default: r = v; g = pp; bl = qq; break;
      }
    }
    p[i] = r; p[HW + i] = g; p[2 * HW + i] = bl;
  }
}

// separable Gaussian blur with reflect padding; weights as torchvision's _get_gaussian_kernel1d (fp32), sigma per image.
// horizontal != 0: along x with `k` taps, else along y.
__global__ void __launch_bounds__(256) gauss_blur_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ sigma,
                                                         int H, int W, int k, int horizontal) {
  const int b = blockIdx.z;
  const long long plane = static_cast<long long>(H) * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= 3 * plane) return;
  const int x = static_cast<int>(idx % W), y = static_cast<int>((idx / W) % H), c = static_cast<int>(idx / plane);
  const float sg = sigma[b];
  const float half = (k - 1) * 0.5f;
  float wsum = 0.f, acc = 0.f;
  const float* p = in + (static_cast<long long>(b) * 3 + c) * plane;
  for (int t = 0; t < k; ++t) {
    const float xx = (-half + t) / sg;
    const float w = expf(-0.5f * xx * xx);
    wsum += w;
    int q = (horizontal ? x : y) + t - k / 2;
    const int n = horizontal ? W : H;
    if (q < 0) q = -q;                       // reflect (no edge repeat)
    if (q >= n) q = 2 * n - 2 - q;
    acc += w * (horizontal ? p[static_cast<long long>(y) * W + q] : p[static_cast<long long>(q) * W + x]);
  }
  out[(static_cast<long long>(b) * 3 + c) * plane + static_cast<long long>(y) * W + x] = acc / wsum;
}

// horizontal flip of the images whose flag is set (planes = channels per image); otherwise copy
__global__ void __launch_bounds__(256) hflip_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ flag, int planes,
                                                    int H, int W) {
  const int b = blockIdx.y;
  const long long n = static_cast<long long>(planes) * H * W;
  const bool flip = flag[b] != 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    out[b * n + i] = in[b * n + (flip ? i - x + (W - 1 - x) : i)];
  }
}

}  // namespace
}  // namespace countr

extern "C" int countr_aug_noise_clamp(const float* img, float* out, int64_t n, float stddev, uint64_t seed, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(img && out && n > 0 && stddev >= 0.f, "bad arguments");
  const long long threads = (n + 1) / 2;
  noise_clamp_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(img, out, n, stddev, seed);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_aug_color_jitter(float* img, const int32_t* ops, const float* factors, double* scratch, int B, int H, int W,
                                       countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(img && ops && factors && scratch && B > 0 && H > 0 && W > 0, "bad arguments");
  const int HW = H * W;
  dim3 grid(static_cast<unsigned>(std::min(592, (HW + 255) / 256)), B);
  // four passes (ColorJitter applies its four functions one after the other in the sampled order); pass j of image b runs
  // ops[j][b] with factors[j][b]; the contrast pass needs the mean of the CURRENT grayscale image, recomputed per pass
  for (int j = 0; j < 4; ++j) {
    COUNTR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * B, stream));
    gray_sum_kernel<<<grid, 256, 0, stream>>>(img, scratch, HW);
    color_op_kernel<<<grid, 256, 0, stream>>>(img, ops + static_cast<long long>(j) * B, factors + static_cast<long long>(j) * B, scratch, HW);
  }
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_aug_gaussian_blur(const float* img, float* tmp, float* out, const float* sigma, int B, int H, int W, int kx, int ky,
                                        countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(img && tmp && out && sigma && B > 0 && H > ky / 2 && W > kx / 2 && kx % 2 == 1 && ky % 2 == 1 && kx <= 31 && ky <= 31,
                 "bad arguments (odd kernel sizes smaller than twice the image)");
  const long long n = 3ll * H * W;
  dim3 grid(static_cast<unsigned>((n + 255) / 256), 1, B);
  gauss_blur_kernel<<<grid, 256, 0, stream>>>(img, tmp, sigma, H, W, kx, 1);
  gauss_blur_kernel<<<grid, 256, 0, stream>>>(tmp, out, sigma, H, W, ky, 0);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_aug_hflip(const float* in, float* out, const int32_t* flags, int B, int planes, int H, int W, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(in && out && flags && in != out && B > 0 && planes > 0 && H > 0 && W > 0, "bad arguments (out of place)");
  const long long n = static_cast<long long>(planes) * H * W;
  dim3 grid(static_cast<unsigned>(std::min<long long>(1184, (n + 255) / 256)), B);
  hflip_kernel<<<grid, 256, 0, stream>>>(in, out, flags, planes, H, W);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

// ------------------------------------------------------------------------------------------
// Mosaic collage (util/FSC147.py:183-262) and the affine warp with key points (:146-171).
// ------------------------------------------------------------------------------------------
namespace countr {
namespace {

struct MosaicArgs {
  countr_mosaic_src src[4];
  int rl, bl, half, C;
};

// tile t at (row r, column c) of its rl x rl resize: Resize((rl, rl))(TF.crop(image, top, left, length, length)), the same
// bilinear arithmetic as crop_resize_kernel
__device__ __forceinline__ float mosaic_tile(const countr_mosaic_src& s, int ch, int rl, int r, int c) {
  const int n = s.length;
  const float scale = static_cast<float>(n) / rl;
  float ry = scale * (r + 0.5f) - 0.5f, rx = scale * (c + 0.5f) - 0.5f;
  ry = ry < 0.f ? 0.f : ry;
  rx = rx < 0.f ? 0.f : rx;
  const int iy0 = min(static_cast<int>(ry), n - 1), ix0 = min(static_cast<int>(rx), n - 1);
  const int iy1 = iy0 + (iy0 < n - 1 ? 1 : 0), ix1 = ix0 + (ix0 < n - 1 ? 1 : 0);
  const float ly1 = fminf(fmaxf(ry - iy0, 0.f), 1.f), lx1 = fminf(fmaxf(rx - ix0, 0.f), 1.f);
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* p = s.img + ch * s.sc + s.top * s.sh + s.left * s.sw;
  const float a = p[iy0 * s.sh + ix0 * s.sw], b = p[iy0 * s.sh + ix1 * s.sw];
  const float cq = p[iy1 * s.sh + ix0 * s.sw], d = p[iy1 * s.sh + ix1 * s.sw];
  return ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * cq + lx1 * d);
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// the reference's seam recurrence `far * (bl - i) / (2 bl) + near * (i + bl) / (2 bl)` in its fp32 operation order
__device__ __forceinline__ float seam(float far_v, float near_v, int i, int bl) {
  const float d = static_cast<float>(2 * bl);
  return __fadd_rn(__fdiv_rn(__fmul_rn(far_v, static_cast<float>(bl - i)), d), __fdiv_rn(__fmul_rn(near_v, static_cast<float>(i + bl)), d));
}

// column `c` (tile coordinates, 0..rl) of the vertical strip built from tiles (2k, 2k+1) at collage row y   (:239-245 / :247-253)
__device__ __forceinline__ float mosaic_strip(const MosaicArgs& a, int k, int ch, int y, int c) {
  const countr_mosaic_src& top = a.src[2 * k];
  const countr_mosaic_src& bot = a.src[2 * k + 1];
  const int rl = a.rl, bl = a.bl, half = a.half;
  float v;
  if (y < half) {
    v = mosaic_tile(top, ch, rl, bl + y, c);
    const int i = half - 1 - y;
    if (i < bl) v = seam(mosaic_tile(bot, ch, rl, bl - i, c), v, i, bl);
  } else {
    v = mosaic_tile(bot, ch, rl, bl + y - half, c);
    const int i = y - half;
    if (i < bl) v = seam(mosaic_tile(top, ch, rl, rl - 1 - bl + i, c), v, i, bl);
  }
  return clamp01(v);
}

__global__ void __launch_bounds__(256) mosaic_kernel(const MosaicArgs a, float* __restrict__ out) {
  const int side = 2 * a.half;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.C * side * side) return;
  const int x = idx % side, y = (idx / side) % side, ch = idx / (side * side);
  const int rl = a.rl, bl = a.bl, half = a.half;
  float v;
  if (x < half) {                                                                       // :255-261
    v = mosaic_strip(a, 0, ch, y, bl + x);
    const int i = half - 1 - x;
    if (i < bl) v = seam(mosaic_strip(a, 1, ch, y, bl - i), v, i, bl);
  } else {
    v = mosaic_strip(a, 1, ch, y, bl + x - half);
    const int i = x - half;
    if (i < bl) v = seam(mosaic_strip(a, 0, ch, y, rl - 1 - bl + i), v, i, bl);
  }
  out[idx] = clamp01(v);
}

// dot maps of the four tiles written straight into the collage canvas (:190-196, :228-232, :240, :248, :256)
__global__ void __launch_bounds__(128) mosaic_dots_kernel(const MosaicArgs a, const double* __restrict__ dots, float* __restrict__ canvas) {
  const int t = blockIdx.y;
  const countr_mosaic_src& s = a.src[t];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.dot_count) return;
  const double dx = dots[(static_cast<long long>(s.dot_begin) + i) * 2], dy = dots[(static_cast<long long>(s.dot_begin) + i) * 2 + 1];
  const int py = min(s.H - 1, static_cast<int>(dy * s.scale_h)), px = min(s.W - 1, static_cast<int>(dx * s.scale_w));
  if (py < s.top || py >= s.top + s.length || px < s.left || px >= s.left + s.length) return;
  // int((p - start) * rl / length): a quotient of small integers, the float64 division cannot reach the next integer
  const int r = min(a.rl - 1, (py - s.top) * a.rl / s.length), c = min(a.rl - 1, (px - s.left) * a.rl / s.length);
  if (r < a.bl || r >= a.rl - a.bl || c < a.bl || c >= a.rl - a.bl) return;
  const int side = 2 * a.half;
  canvas[((t & 1) * a.half + r - a.bl) * side + (t >> 1) * a.half + c - a.bl] = 1.0f;
}

struct Affine6 { double m[6]; };

// out(y, x) = bilinear sample of img at M^-1 (x, y), zero outside (constant border 0, order 1)
__global__ void __launch_bounds__(256) affine_warp_kernel(const float* __restrict__ img, float* __restrict__ out, int C, int H, int W, const Affine6 inv) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int x = idx % W, y = idx / W;
  const double sx = inv.m[0] * x + inv.m[1] * y + inv.m[2], sy = inv.m[3] * x + inv.m[4] * y + inv.m[5];
  const double fx = floor(sx), fy = floor(sy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float wx = static_cast<float>(sx - fx), wy = static_cast<float>(sy - fy);
  const bool in_x0 = x0 >= 0 && x0 < W, in_x1 = x0 + 1 >= 0 && x0 + 1 < W, in_y0 = y0 >= 0 && y0 < H, in_y1 = y0 + 1 >= 0 && y0 + 1 < H;
  const bool far_out = sx <= -1.0 || sx >= W || sy <= -1.0 || sy >= H;
  for (int ch = 0; ch < C; ++ch) {
    const float* p = img + static_cast<long long>(ch) * H * W;
    float v = 0.f;
    if (!far_out) {
      const float a = in_y0 && in_x0 ? p[y0 * W + x0] : 0.f, b = in_y0 && in_x1 ? p[y0 * W + x0 + 1] : 0.f;
      const float c = in_y1 && in_x0 ? p[(y0 + 1) * W + x0] : 0.f, d = in_y1 && in_x1 ? p[(y0 + 1) * W + x0 + 1] : 0.f;
      v = (1.f - wy) * ((1.f - wx) * a + wx * b) + wy * ((1.f - wx) * c + wx * d);
    }
    out[static_cast<long long>(ch) * H * W + idx] = v;
  }
}

// key points through the forward matrix, then the reference's dot map (:146-166): truncated coordinates, points that leave
// the image are dropped
__global__ void __launch_bounds__(128) affine_dots_kernel(const double* __restrict__ dots, int n, double scale_h, double scale_w, int H, int W,
                                                          const Affine6 fwd, float* __restrict__ canvas) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double kx = min(W - 1, static_cast<int>(dots[2 * i] * scale_w)), ky = min(H - 1, static_cast<int>(dots[2 * i + 1] * scale_h));
  const double ax = fwd.m[0] * kx + fwd.m[1] * ky + fwd.m[2], ay = fwd.m[3] * kx + fwd.m[4] * ky + fwd.m[5];
  if (!(ax >= 0.0 && ax < W && ay >= 0.0 && ay < H)) return;                  // Keypoint.is_out_of_image
  const int r = static_cast<int>(ay), c = static_cast<int>(ax);
  if (r > H - 1 || c > W - 1) return;
  canvas[r * W + c] = 1.0f;
}

bool mosaic_args(const countr_mosaic_src* src, int rl, int bl, int C, MosaicArgs* a, const char** why) {
  a->rl = rl, a->bl = bl, a->half = rl - 2 * bl, a->C = C;
  if (!src || C <= 0 || bl <= 0 || a->half <= 0 || bl >= a->half) { *why = "bad rl / bl"; return false; }
  for (int t = 0; t < 4; ++t) {
    const countr_mosaic_src& s = src[t];
    if (!s.img || s.H <= 0 || s.W <= 0 || s.length <= 0 || s.top < 0 || s.left < 0 || s.top + s.length > s.H || s.left + s.length > s.W ||
        s.dot_count < 0 || s.dot_begin < 0) {
      *why = "a source crop leaves its image";
      return false;
    }
    a->src[t] = s;
  }
  return true;
}

}  // namespace
}  // namespace countr

extern "C" int countr_aug_mosaic(const countr_mosaic_src* src, int rl, int bl, int C, float* out, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MosaicArgs a;
  const char* why = "";
  COUNTR_REQUIRE(out && mosaic_args(src, rl, bl, C, &a, &why), "mosaic: %s (rl=%d bl=%d)", why, rl, bl);
  const int total = C * 4 * a.half * a.half;
  mosaic_kernel<<<(total + 255) / 256, 256, 0, stream>>>(a, out);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_aug_mosaic_dots(const countr_mosaic_src* src, int rl, int bl, const double* dots, float* canvas, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MosaicArgs a;
  const char* why = "";
  COUNTR_REQUIRE(canvas && mosaic_args(src, rl, bl, 1, &a, &why), "mosaic dots: %s (rl=%d bl=%d)", why, rl, bl);
  const int side = 2 * a.half;
  COUNTR_CHECK_CUDA(cudaMemsetAsync(canvas, 0, sizeof(float) * side * side, stream));
  int n_max = 0;
  for (int t = 0; t < 4; ++t) n_max = std::max(n_max, a.src[t].dot_count);
  if (n_max > 0) {
    COUNTR_REQUIRE(dots, "null dots");
    mosaic_dots_kernel<<<dim3((n_max + 127) / 128, 4), 128, 0, stream>>>(a, dots, canvas);
  }
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_density_filter(const float* canvas, float* tmp, float* out, int B, int H, int W, const double* weights, int radius,
                                     float gain, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(canvas && tmp && out && weights && tmp != out && tmp != canvas, "null or aliased pointer");
  COUNTR_REQUIRE(B > 0 && H > 0 && W > 0 && radius >= 0 && radius < H && radius < W && radius <= 512, "bad shape B=%d H=%d W=%d radius=%d", B, H, W,
                 radius);
  const long long total = static_cast<long long>(B) * H * W;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  const size_t smem = sizeof(double) * (radius + 1);
  gauss_pass_kernel<<<blocks, 256, smem, stream>>>(canvas, tmp, weights, radius, B, H, W, 0, 1.0f);
  gauss_pass_kernel<<<blocks, 256, smem, stream>>>(tmp, out, weights, radius, B, H, W, 1, gain);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_aug_affine(const float* img, float* out, int C, int H, int W, const double* inverse, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(img && out && img != out && inverse && C > 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30), "bad arguments");
  Affine6 m;
  for (int i = 0; i < 6; ++i) m.m[i] = inverse[i];
  affine_warp_kernel<<<(H * W + 255) / 256, 256, 0, stream>>>(img, out, C, H, W, m);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_aug_affine_dots(const double* dots, int n, double scale_h, double scale_w, int H, int W, const double* forward,
                                      float* canvas, countr_stream_t stream_) {
  using namespace countr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(canvas && forward && n >= 0 && H > 0 && W > 0 && (n == 0 || dots), "bad arguments");
  COUNTR_CHECK_CUDA(cudaMemsetAsync(canvas, 0, sizeof(float) * H * W, stream));
  if (n > 0) {
    Affine6 m;
    for (int i = 0; i < 6; ++i) m.m[i] = forward[i];
    affine_dots_kernel<<<(n + 127) / 128, 128, 0, stream>>>(dots, n, scale_h, scale_w, H, W, m, canvas);
  }
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}
