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
