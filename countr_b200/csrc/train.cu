// countr_b200 — script-side pieces of the fine-tune step as single kernels:
// the masked-MSE density loss with its gradient, and unscale + AdamW over all decoder parameters at once.
//
// replaces: FSC_finetune_cross.py:290-295 (loss), util/misc.py:266-280 + torch.optim.AdamW (:235) — about 25
// tiny ATen launches per step in the reference script.
#include "../../include/countr_b200.h"
#include "common.cuh"

namespace countr {
namespace {

__device__ __forceinline__ float load_any(const void* p, long long i, int dtype) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  const uint16_t u = reinterpret_cast<const uint16_t*>(p)[i];
  if (dtype == 1) return __half2float(__ushort_as_half(u));
  return __uint_as_float(static_cast<uint32_t>(u) << 16);
}

// loss = sum_b,p (out - gt)^2 * mask[p] / (HW) / B ;  dout = 2 (out - gt) mask / (HW B) * grad_scale   (fp32)
__global__ void __launch_bounds__(256) masked_mse_kernel(const void* __restrict__ out, int out_dtype, const float* __restrict__ gt,
                                                          const float* __restrict__ mask, float* __restrict__ loss,
                                                          float* __restrict__ dout, long long total, int HW, float inv_norm,
                                                          float grad_scale) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float m = mask[i % HW];
    const float d = load_any(out, i, out_dtype) - gt[i];
    acc += d * d * m;
    if (dout) dout[i] = 2.f * d * m * inv_norm * grad_scale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(loss, s * inv_norm);
  }
}

struct AdamTensor {
  float* param;          // fp32 master parameter
  long long grad_off;    // offset of its gradient inside the flat gradient arena of this step
  long long moment_off;  // offset of its Adam moments inside the (all-parameter) moment arenas
  long long numel;
  float weight_decay;
  int step_idx;          // index of this parameter's own step counter (torch.optim keeps one per parameter)
};

__global__ void adam_step_inc_kernel(const AdamTensor* __restrict__ tensors, int n, float* __restrict__ steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) steps[tensors[i].step_idx] += 1.f;
}

// one block = 1024 consecutive elements of one tensor (chunk table built by the host once)
__global__ void __launch_bounds__(256) adamw_kernel(const AdamTensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                     const float* __restrict__ grad, float* __restrict__ exp_avg,
                                                     float* __restrict__ exp_avg_sq, const float* __restrict__ step_ptr, float lr,
                                                     float beta1, float beta2, float eps, float inv_scale) {
  const int2 ch = chunks[blockIdx.x];
  const AdamTensor t = tensors[ch.x];
  const float step = step_ptr[t.step_idx];
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const long long base = static_cast<long long>(ch.y) * 1024;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long e = base + u * 256 + threadIdx.x;
    if (e < t.numel) {
      const float g = grad[t.grad_off + e] * inv_scale;
      float p = t.param[e];
      float m = exp_avg[t.moment_off + e], v = exp_avg_sq[t.moment_off + e];
      p *= 1.f - lr * t.weight_decay;                        // decoupled weight decay (torch.optim.AdamW)
      m = beta1 * m + (1.f - beta1) * g;
      v = beta2 * v + (1.f - beta2) * g * g;
      const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
      p -= step_size * m / denom;
      t.param[e] = p;
      exp_avg[t.moment_off + e] = m;
      exp_avg_sq[t.moment_off + e] = v;
    }
  }
}

}  // namespace
}  // namespace countr

using namespace countr;

extern "C" int countr_masked_mse(const void* out, int out_dtype, const float* gt, const float* mask, float* loss, float* dout, int B,
                                 int H, int W, float grad_scale, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(out && gt && mask && loss && out_dtype >= 0 && out_dtype <= 2, "bad arguments");
  const long long total = static_cast<long long>(B) * H * W;
  COUNTR_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  const int blocks = static_cast<int>(total / 256 / 4 < 148 * 4 ? (total + 1023) / 1024 : 148 * 4);
  masked_mse_kernel<<<blocks, 256, 0, stream>>>(out, out_dtype, gt, mask, loss, dout, total, H * W, 1.f / (static_cast<float>(H) * W * B),
                                               grad_scale);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int countr_adamw_step(const void* tensors, int num_tensors, const void* chunks, int num_chunks, const float* grad,
                                 float* exp_avg, float* exp_avg_sq, float* step, float lr, float beta1, float beta2, float eps,
                                 float inv_scale, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(tensors && chunks && grad && exp_avg && exp_avg_sq && step && num_chunks > 0, "bad arguments");
  adam_step_inc_kernel<<<(num_tensors + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const AdamTensor*>(tensors), num_tensors, step);
  adamw_kernel<<<num_chunks, 256, 0, stream>>>(reinterpret_cast<const AdamTensor*>(tensors), reinterpret_cast<const int2*>(chunks), grad,
                                               exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, inv_scale);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}
