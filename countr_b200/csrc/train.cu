// countr_b200 — script-side pieces of the fine-tune step as single kernels:
//   * the masked-MSE density loss with its gradient, the per-image counts (sum/60) and the batch MAE / MSE of the
//     counts, with the Bernoulli(0.8) pixel mask either supplied or drawn on the device
//   * unscale + inf/NaN check + global gradient norm over the flat gradient arena
//   * AdamW over all decoder parameters at once, skipped on overflow, followed by the GradScaler update
// Loss scale, learning rate, found_inf and the step counter live in a small DEVICE state block, so a captured CUDA graph
// follows the lr schedule and the dynamic loss scale without being re-captured.
//
// replaces: FSC_finetune_cross.py:290-303 (mask, loss, counts), util/misc.py:260-301 (GradScaler scale / unscale_ /
// inf-skip / update, get_grad_norm_) + torch.optim.AdamW (:235) — about 25 tiny ATen launches and one H2D per step.
#include "../../include/countr_b200.h"
#include "common.cuh"

namespace countr {
namespace {

enum : int { ST_SCALE = 0, ST_GROWTH = 1, ST_FOUND_INF = 2, ST_GRAD_NORM = 3, ST_LR = 4, ST_STEP = 5 };

__device__ __forceinline__ float load_any(const void* p, long long i, int dtype) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  const uint16_t u = reinterpret_cast<const uint16_t*>(p)[i];
  if (dtype == 1) return __half2float(__ushort_as_half(u));
  return __uint_as_float(static_cast<uint32_t>(u) << 16);
}

// Counter-based Bernoulli draw: splitmix64 finaliser over (seed, step, pixel); the tests restate it in numpy.
__device__ __forceinline__ bool bernoulli_keep(unsigned long long seed, unsigned long long step, unsigned long long pixel,
                                               unsigned int threshold24) {
  unsigned long long z = seed + step * 0xD1B54A32D192ED03ull + (pixel + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<unsigned int>(z >> 40) < threshold24;
}

constexpr int kLossBlocksPerImg = 36;

// grid (kLossBlocksPerImg, B).  partial: [B][kLossBlocksPerImg][3] doubles; ticket: one int (left at 0).
// result[0] = loss, [1] = batch MAE of the counts, [2] = batch MSE of the counts; counts[b] = {pred, gt} / 60.
__global__ void __launch_bounds__(256) finetune_loss_kernel(const void* __restrict__ out, int out_dtype, const void* __restrict__ gt,
                                                             int gt_dtype, const float* __restrict__ mask, long long mask_bstride,
                                                             unsigned long long seed, unsigned int threshold24,
                                                             const float* __restrict__ state, float grad_scale, float* __restrict__ dout,
                                                             unsigned char* __restrict__ mask_out, double* __restrict__ partial,
                                                             int* __restrict__ ticket, float* __restrict__ result,
                                                             float* __restrict__ counts, int B, int HW) {
  __shared__ double red[3][8];
  __shared__ int last;
  const int b = blockIdx.y;
  const float scale = state != nullptr ? state[ST_SCALE] : grad_scale;
  const unsigned long long step = state != nullptr ? static_cast<unsigned long long>(state[ST_STEP]) : 0ull;
  const float inv_norm = 1.f / (static_cast<float>(HW) * B);
  const float gmul = 2.f * inv_norm * scale;
  float a_loss = 0.f, a_pred = 0.f, a_gt = 0.f;
  const long long base = static_cast<long long>(b) * HW;
#pragma unroll 4
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float m;
    if (mask != nullptr) m = mask[b * mask_bstride + i];
    else m = bernoulli_keep(seed, step, static_cast<unsigned long long>(i), threshold24) ? 1.f : 0.f;
    if (mask_out != nullptr && b == 0) mask_out[i] = m != 0.f;
    const float o = load_any(out, base + i, out_dtype), g = load_any(gt, base + i, gt_dtype);
    const float d = o - g;
    a_loss = fmaf(d * d, m, a_loss);
    a_pred += o;
    a_gt += g;
    if (dout != nullptr) dout[base + i] = d * m * gmul;
  }
  a_loss = warp_sum(a_loss); a_pred = warp_sum(a_pred); a_gt = warp_sum(a_gt);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a_loss; red[1][threadIdx.x >> 5] = a_pred; red[2][threadIdx.x >> 5] = a_gt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0., s1 = 0., s2 = 0.;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s0 += red[0][k]; s1 += red[1][k]; s2 += red[2][k]; }
    double* pp = partial + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 3;
    pp[0] = s0; pp[1] = s1; pp[2] = s2;
    __threadfence();
    last = atomicAdd(ticket, 1) == static_cast<int>(gridDim.x * gridDim.y) - 1;
  }
  __syncthreads();
  if (last) {
    // fixed summation order (thread b sums image b's partials in index order, thread 0 combines the images in order): the
    // loss and the metrics are bit-reproducible run to run
    __shared__ double fin[3][256];
    __threadfence();
    const int nblk = static_cast<int>(gridDim.x);
    for (int bb = threadIdx.x; bb < B; bb += blockDim.x) {
      double l = 0., pr = 0., g = 0.;
      for (int k = 0; k < nblk; ++k) {
        const double* pp = partial + (static_cast<long long>(bb) * nblk + k) * 3;
        l += __ldcg(pp); pr += __ldcg(pp + 1); g += __ldcg(pp + 2);
      }
      const float pc = static_cast<float>(pr / 60.0), gc = static_cast<float>(g / 60.0);
      if (counts != nullptr) { counts[2 * bb] = pc; counts[2 * bb + 1] = gc; }
      const float err = fabsf(pc - gc);
      fin[0][bb] = l; fin[1][bb] = static_cast<double>(err); fin[2][bb] = static_cast<double>(err * err);     // B <= 256 (host check)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double loss = 0., mae = 0., mse = 0.;
      for (int bb = 0; bb < B; ++bb) { loss += fin[0][bb]; mae += fin[1][bb]; mse += fin[2][bb]; }
      result[0] = static_cast<float>(loss * static_cast<double>(inv_norm));
      result[1] = static_cast<float>(mae / B);
      result[2] = static_cast<float>(mse / B);
      *ticket = 0;
    }
  }
}

// sum of squares + non-finite check of the (still scaled) gradient arena; last block finalises into the state block
__global__ void __launch_bounds__(256) grad_stats_kernel(const float* __restrict__ grad, long long n, double* __restrict__ partial,
                                                          int* __restrict__ ticket, float* __restrict__ state) {
  __shared__ float red[8];
  __shared__ int bad[8];
  __shared__ int last;
  float acc = 0.f;
  int nonfinite = 0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(grad);
#pragma unroll 4
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc = fmaf(grad[i], grad[i], acc);
  // any inf / NaN element makes its square (hence the partial sum) non-finite
  nonfinite = !isfinite(acc);
  acc = warp_sum(acc);
  nonfinite = __any_sync(0xffffffffu, nonfinite);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; bad[threadIdx.x >> 5] = nonfinite; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.;
    int nf = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += red[k]; nf |= bad[k]; }
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = nf ? 1.0 : 0.0;
    __threadfence();
    last = atomicAdd(ticket, 1) == static_cast<int>(gridDim.x) - 1;
  }
  __syncthreads();
  if (last) {
    // fixed order: thread i sums partials i, i + 256, ...; thread 0 combines the 256 sums in index order
    __shared__ double fin[256];
    __shared__ int finbad[256];
    __threadfence();
    double s = 0.;
    int nf = 0;
    for (int k = threadIdx.x; k < static_cast<int>(gridDim.x); k += blockDim.x) {
      s += __ldcg(partial + 2 * k);
      nf |= __ldcg(partial + 2 * k + 1) != 0.0;
    }
    fin[threadIdx.x] = s;
    finbad[threadIdx.x] = nf;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.;
      int bad_any = 0;
      for (int k = 0; k < 256; ++k) { tot += fin[k]; bad_any |= finbad[k]; }
      const float scale = state[ST_SCALE];
      const float norm = static_cast<float>(sqrt(tot)) / scale;
      state[ST_FOUND_INF] = (bad_any || !isfinite(norm)) ? 1.f : 0.f;
      state[ST_GRAD_NORM] = norm;
      *ticket = 0;
    }
  }
}

struct AdamTensor {
  float* param;          // fp32 master parameter
  long long grad_off;    // offset of its gradient inside the flat gradient arena
  long long moment_off;  // offset of its Adam moments inside the moment arenas
  long long numel;
  float weight_decay;
  int step_idx;          // index of this parameter's own step counter (torch.optim keeps one per parameter)
  int flag_idx;          // 0: always updated; k > 0: only when flags[k - 1] != 0 (parameter group used this step on some rank)
  int pad;
};

__device__ __forceinline__ bool tensor_active(const AdamTensor& t, const float* flags) {
  return t.flag_idx == 0 || flags == nullptr || flags[t.flag_idx - 1] != 0.f;
}

// one block = 1024 consecutive elements of one tensor (chunk table built by the host once)
__global__ void __launch_bounds__(256) adamw_kernel(const AdamTensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                     const float* __restrict__ grad, const float* __restrict__ flags,
                                                     float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                     const float* __restrict__ step_ptr, const float* __restrict__ state, float beta1,
                                                     float beta2, float eps) {
  if (state[ST_FOUND_INF] != 0.f) return;      // GradScaler.step: skip the whole update on overflow
  const int2 ch = chunks[blockIdx.x];
  const AdamTensor t = tensors[ch.x];
  if (!tensor_active(t, flags)) return;
  const float lr = state[ST_LR], inv_scale = 1.f / state[ST_SCALE];
  const float step = step_ptr[t.step_idx] + 1.f;            // the counters are advanced by adam_finish_kernel
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

// per-parameter step counters + torch.cuda.amp.GradScaler.update() (growth 2.0 / backoff 0.5 / interval 2000 by default)
__global__ void adam_finish_kernel(const AdamTensor* __restrict__ tensors, int n, const float* __restrict__ flags,
                                   float* __restrict__ steps, float* __restrict__ state, float growth, float backoff, int interval) {
  const bool found_inf = state[ST_FOUND_INF] != 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (!found_inf && tensor_active(tensors[i], flags)) steps[tensors[i].step_idx] += 1.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (interval > 0) {
      if (found_inf) {
        state[ST_SCALE] *= backoff;
        state[ST_GROWTH] = 0.f;
      } else {
        const float g = state[ST_GROWTH] + 1.f;
        if (g >= static_cast<float>(interval)) {
          state[ST_SCALE] *= growth;
          state[ST_GROWTH] = 0.f;
        } else {
          state[ST_GROWTH] = g;
        }
      }
    }
    state[ST_STEP] += 1.f;
  }
}

}  // namespace
}  // namespace countr

using namespace countr;

extern "C" int countr_finetune_loss(const void* out, int out_dtype, const void* gt, int gt_dtype, const float* mask,
                                    int64_t mask_bstride, uint64_t seed, float keep_prob, const float* state, float grad_scale,
                                    float* dout, uint8_t* mask_out, void* scratch, float* result, float* counts, int B, int H, int W,
                                    countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(out && gt && scratch && result && out_dtype >= 0 && out_dtype <= 2 && gt_dtype >= 0 && gt_dtype <= 2, "bad arguments");
  COUNTR_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 256, "bad shape B=%d H=%d W=%d (B <= 256)", B, H, W);
  COUNTR_REQUIRE(mask != nullptr || (keep_prob >= 0.f && keep_prob <= 1.f), "keep_prob must be in [0, 1]");
  COUNTR_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, "scratch must be 8-byte aligned");
  // scratch: one int ticket (8 bytes, zero before the FIRST call; the kernel leaves it at zero) + B * 36 * 3 doubles
  int* ticket = reinterpret_cast<int*>(scratch);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 8);
  const unsigned int thr = static_cast<unsigned int>(static_cast<double>(keep_prob) * 16777216.0);
  finetune_loss_kernel<<<dim3(kLossBlocksPerImg, B), 256, 0, stream>>>(out, out_dtype, gt, gt_dtype, mask, mask_bstride, seed, thr, state,
                                                                        grad_scale, dout, mask_out, partial, ticket, result, counts, B,
                                                                        H * W);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int64_t countr_finetune_loss_scratch_bytes(int B) { return 8 + static_cast<int64_t>(B) * kLossBlocksPerImg * 3 * 8; }

extern "C" int countr_grad_stats(const float* grad, int64_t n, void* scratch, float* state, countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(grad && scratch && state && n > 0, "bad arguments");
  COUNTR_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15u) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, "misaligned buffers");
  int* ticket = reinterpret_cast<int*>(scratch);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 8);
  const long long want = (n / 4 + 255) / 256;
  const int blocks = static_cast<int>(want < 1 ? 1 : (want > 592 ? 592 : want));     // 4 x 148; scratch holds 592 partial pairs
  grad_stats_kernel<<<blocks, 256, 0, stream>>>(grad, n, partial, ticket, state);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}

extern "C" int64_t countr_grad_stats_scratch_bytes(void) { return 8 + 592 * 2 * 8; }

extern "C" int countr_adamw_update(const void* tensors, int num_tensors, const void* chunks, int num_chunks, const float* grad,
                                   const float* flags, float* exp_avg, float* exp_avg_sq, float* step, float* state, float beta1,
                                   float beta2, float eps, float growth_factor, float backoff_factor, int growth_interval,
                                   countr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  COUNTR_REQUIRE(tensors && chunks && grad && exp_avg && exp_avg_sq && step && state && num_chunks > 0 && num_tensors > 0, "bad arguments");
  adamw_kernel<<<num_chunks, 256, 0, stream>>>(reinterpret_cast<const AdamTensor*>(tensors), reinterpret_cast<const int2*>(chunks), grad,
                                               flags, exp_avg, exp_avg_sq, step, state, beta1, beta2, eps);
  adam_finish_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const AdamTensor*>(tensors), num_tensors, flags, step, state, growth_factor,
                                            backoff_factor, growth_interval);
  COUNTR_CHECK_CUDA(cudaGetLastError());
  return COUNTR_OK;
}
