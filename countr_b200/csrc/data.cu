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
