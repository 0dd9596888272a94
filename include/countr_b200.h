/* countr_b200 — C ABI of the sm_100a kernel library (libcountr_sm100.so).
 *
 * The reference (Verg-Avesta/CounTR) is pure PyTorch and has no FFI of its own: every device
 * operation on its hot path is an implicit ATen call made from models_mae_cross.py /
 * models_crossvit.py.  Each entry point below replaces one family of those
 * call sites; the citation after "replaces:" is the reference line whose arithmetic it performs.
 * The Python modules in countr_b200/ (same class names and state_dict keys as the reference)
 * are the only callers; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; no torch types; every pointer is DEVICE memory owned by the caller
 *   - nothing here allocates, frees or synchronises; work is enqueued on `stream`
 *   - return 0 on success, <0 on error (countr_last_error() gives the message, thread-local)
 *   - 16-bit activations are IEEE fp16 unless the descriptor says bf16; accumulation is fp32
 */
#ifndef COUNTR_B200_H_
#define COUNTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* countr_stream_t; /* cudaStream_t */

const char* countr_last_error(void);
const char* countr_version(void);
int countr_check_device(void);
int countr_num_sms(void);
/* Persistent kernels launch one CTA (or CTA pair) per SM.  A data-parallel trainer that overlaps its NCCL gradient all-reduce with
 * compute sets a budget below the device's SM count (e.g. 144 of 148) so that the collective's CTAs find free SMs instead of
 * queueing behind — and splitting into a second wave — the persistent grids.  n <= 0 restores "all SMs"; returns the value in
 * effect (also settable with the COUNTR_SM_BUDGET environment variable). */
int countr_set_sm_budget(int n);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (statistics / split-K accumulators) */
int countr_memset_zero(void* ptr, size_t bytes, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM 3x3 convolution (tcgen05.mma, TMA-fed, TMEM accumulators)
 *
 *   C[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * B[b][n][k] )
 *
 * replaces: every nn.Linear on the path — Attention.qkv/proj (models_crossvit.py:77,79,84,92),
 *   Mlp.fc1/fc2 (:55,58,62,65), CrossAttention.wq/wk/wv/proj (:104-108,115-127),
 *   decoder_embed (models_mae_cross.py:39,152), PatchEmbed.proj as a patch GEMM (:27,138);
 *   the bmm's of the attention backward; and, in conv mode, the Conv2d 3x3 of decode_head0..3
 *   (:80-100) and decoder_proj2..4 (:53-71) plus their dX.
 *
 * Operand layouts (16-bit):
 *   a_mn == 0 : A[m][k] at a + m*lda + k      (K contiguous,  "K-major")
 *   a_mn == 1 : A[m][k] at a + k*lda + m      (M contiguous, "MN-major"; used for dW = dY^T X)
 *   b_mn likewise for B[n][k].   lda/ldb/batch strides must be multiples of 8 elements.
 * Batch index = b1*nb2 + b2 with independent strides (so a [B,L,3,H,dh] qkv buffer can be
 * addressed per (batch, head) without a copy).
 *
 * conv mode (conv_h > 0): A is an NHWC activation [nb1=B][H][W][Cin] read through shifted
 * TMA boxes (zero-filled halo = padding 1); B is the weight as [N=Cout][9*Cin] with k ordered
 * (ky, kx, cin); M tiles are conv_bx x conv_by pixel rectangles (bx*by == 128); C is NHWC
 * [B][H][W][ldc].
 *
 * conv weight-gradient mode (conv_h > 0 && conv_dw != 0): dW[co][tap][ci] += sum over pixels of
 * dY[b,y,x,co] * X[b,y+ky-1,x+kx-1,ci].  A = dY NHWC [conv_batch][H][W][M=Cout] (lda = Cout),
 * B = X NHWC [conv_batch][H][W][N=Cin] (ldb = Cin), both read as MN-major 64-pixel TMA boxes
 * (conv_bx*conv_by == 64, B shifted by the tap); K = conv_batch * pixel_tiles * 64, nb2 = 9 taps,
 * sc2 = Cin, ldc = 9*Cin, fp32 atomic output (split_k over the pixel tiles).
 * ------------------------------------------------------------------------------------------ */
typedef struct countr_gemm_desc {
  const void* a;
  const void* b;
  int64_t lda, sa1, sa2; /* elements */
  int64_t ldb, sb1, sb2;
  int32_t a_mn, b_mn;
  int32_t M, N, K;
  int32_t nb1, nb2;
  int32_t bf16; /* 0: fp16 operands/outputs, 1: bf16 */
  /* tiling */
  int32_t bn;      /* N tile: 0 = choose; else multiple of 32 (64 if b_mn), <= 256 */
  int32_t split_k; /* >= 1; > 1 requires atomic == 1 */
  int32_t cluster; /* CTAs per cluster sharing the B tile by TMA multicast: 0 = choose, else 1, 2 or 4 */
  int32_t cta_pair; /* tcgen05 cta_group::2 — two SMs compute one 256 x bn tile (K-major B only):
                       0 = choose (on for large convolutions), 1 = on, -1 = off */
  /* conv mode */
  int32_t conv_h, conv_w, conv_cin, conv_bx, conv_by;
  int32_t conv_dw, conv_batch; /* conv_dw != 0: weight-gradient mode, see below */
  /* epilogue */
  void* c;
  int64_t ldc, sc1, sc2;
  int32_t out_f32;       /* 1: C is fp32, 0: C is 16-bit */
  int32_t atomic;        /* 1: C (fp32, pre-zeroed by caller) += result via red.global.add */
  float alpha;
  const float* bias;     /* [N] fp32 or NULL */
  int32_t act;           /* 0 none, 1 GELU(erf), 2 multiply by GELU'(aux) */
  void* aux;             /* 16-bit [M][ldaux]: act==1 -> pre-activation is stored here (may be NULL);
                            act==2 -> pre-activation is read from here */
  int64_t ldaux;
  const float* residual; /* fp32, added last; row index = m % res_mod when res_mod > 0 */
  int64_t ldr;
  int32_t res_mod;
  double* gn_stats;      /* conv mode: [B][N/32][2] (sum, sum of squares) += over 32-channel groups */
  /* LayerNorm folded into the GEMMs on either side of it (the frozen encoder: Block.norm1 / norm2 between proj|fc2 and
   * qkv|fc1, timm Block.forward).  PRODUCER (fp32 out + residual): ln_x16 = 16-bit copy of the output rows [M][ld_x16] and
   * ln_stats = float2 [M][8] partial (sum, sum of squares) of every output row, slot 2 * n_tile + column half (no atomics:
   * bit-reproducible).  CONSUMER (16-bit out; A = that 16-bit copy, B = W * diag(gamma)): ln_stats != NULL with ln_colsum:
   *   C[r][c] = rstd_r * acc[r][c] - rstd_r * mean_r * ln_colsum[c] + bias[c],
   * ln_colsum[c] = sum_k B[c][k], bias = b + W beta, mean / rstd from the eight partials over ln_dim columns (eps ln_eps).
   * Algebraically LayerNorm(x) W^T + b with the statistics taken from the fp32 values. */
  void* ln_x16;
  int64_t ld_x16;
  float* ln_stats;
  const float* ln_colsum;
  int32_t ln_dim;
  float ln_eps;
} countr_gemm_desc;

int countr_gemm(const countr_gemm_desc* d, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm on the fp32 residual stream.  y16 (16-bit GEMM operand) and/or y32 are written;
 * mean/rstd ([rows], optional) are saved for the backward pass.
 * replaces: nn.LayerNorm call sites — timm Block.norm1/norm2, SupervisedMAE.norm / decoder_norm
 *   (models_mae_cross.py:32-35,78,146,182; eps=1e-6 via :214), CrossAttentionBlock.norm0/1/2
 *   (models_crossvit.py:137,142,147,153-155).  Backward: native_layer_norm_backward.
 * ------------------------------------------------------------------------------------------ */
int countr_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y16, float* y32,
                         float* mean, float* rstd, int rows, int D, float eps, int bf16,
                         countr_stream_t stream);
/* dx (+)= LN'(dy); dx16 (optional) = 16-bit copy of the updated dx (next GEMM's operand);
 * dgamma/dbeta (optional, pre-zeroed or carrying earlier contributions) +=;
 * dx_colsum [D] (optional) += column sums of the updated dx = bias gradient of the next Linear up the chain.
 * partials != NULL (then dgamma / dbeta / dx_colsum must be NULL): no atomics — every block writes its sums of
 * dgamma, dbeta and of the updated dx to partials[3][countr_layernorm_bwd_blocks(rows)][D]; the caller adds the block rows up
 * later (countr_grouped_colsum at the end of the backward). */
int countr_layernorm_bwd_blocks(int rows);
int countr_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                         const float* rstd, float* dx, void* dx16, float* dgamma, float* dbeta, float* dx_colsum, float* partials,
                         int rows, int D, int accumulate, int bf16, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-head self-attention forward: softmax(scale * Q K^T) V, flash-style on tcgen05.
 * qkv is the packed output of the qkv Linear, [B][L][3][H][dh] 16-bit; out is [B][L][H*dh].
 * lse ([B][H][L] fp32, optional) receives log-sum-exp of the scaled scores.
 * replaces: Attention.forward, models_crossvit.py:85-91 (reshape/permute, q@k^T*scale, softmax,
 *   attn@v, transpose) — used by the 12 encoder blocks and the FIM self-attention.
 * ------------------------------------------------------------------------------------------ */
int countr_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int dh, float scale,
                         int bf16, countr_stream_t stream);

/* Fused flash-style backward of the same op: given qkv, the forward output `out`, its gradient `dout` ([B*L][H*dh]) and
 * lse, writes dqkv [B][L][3][H][dh].  No [B,H,L,L] tensor is materialised.
 *   head_dim 32 (the FIM self-attention, L <= 640): one CTA per (batch, head), everything resident, workspace unused (NULL).
 *   head_dim 64 (the MAE pre-training encoder, any L): one CTA per (batch, head, 128-key block); dQ is accumulated in the
 *   fp32 workspace (countr_attention_bwd_workspace_bytes, 16-byte aligned; zeroed here) and converted into dqkv at the end.
 * replaces: autograd of models_crossvit.py:87-91 / timm Attention (bmm / _softmax_backward_data / bmm). */
int64_t countr_attention_bwd_workspace_bytes(int B, int L, int H, int dh);
int countr_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* workspace, int B,
                         int L, int H, int dh, float scale, int bf16, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Cross-attention core for a handful of exemplar tokens (S <= 8): per token and head
 * softmax_s(scale * q . k_s) applied to v_s.   q16 [B*L][D] 16-bit, k32/v32 [B][S][D] fp32,
 * out16 [B*L][D] 16-bit, probs [B*L][D/dh][S] fp32 (optional, kept for backward).
 * kv_broadcast != 0: k32/v32 are [S][D], shared by every image (zero-shot shot_token, :176).
 * replaces: CrossAttention.forward, models_crossvit.py:122-126.
 * ------------------------------------------------------------------------------------------ */
int countr_cross_attn_core(const void* q16, const float* k32, const float* v32, void* out16, float* probs,
                           int B, int L, int S, int D, int dh, float scale, int bf16, int kv_broadcast,
                           countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Layout / cast helpers (HBM-bound streaming kernels).
 * ------------------------------------------------------------------------------------------ */
/* dst16[i] = (16-bit) (src[i] * scale) : fp32 master weights / gradients -> GEMM operands */
int countr_cast_f32_to_16(const float* src, void* dst, int64_t n, float scale, int bf16, countr_stream_t stream);
/* dst16[c][r] = src[r][c] : W^T operand for dX = dY W */
int countr_cast_transpose_f32_to_16(const float* src, void* dst, int R, int C, int bf16, countr_stream_t stream);
/* PatchEmbed gather (timm PatchEmbed.proj as a GEMM; models_mae_cross.py:27,138): NCHW image of
 * dtype code {0 fp32, 1 fp16, 2 bf16} with element strides (sb,sc,sh,sw) -> out16 [B*gh*gw][C*P*P],
 * columns ordered (c, ky, kx) like Conv2d.weight.view(out, -1).  Patch sizes that are not a multiple of 8 or do not divide
 * the image (mae_vit_huge_patch14: P = 14 on 384 px -> 27 x 27 patches, trailing pixels dropped like the strided Conv2d
 * drops them): rows are [B*(H/P)*(W/P)][ld], ld = C*P*P rounded up to a multiple of 8, zero padded. */
int countr_patchify(const void* img, int dtype, int64_t sb, int64_t sc, int64_t sh, int64_t sw, void* out,
                    int B, int C, int H, int W, int P, int bf16, countr_stream_t stream);
/* Conv2d 3x3 weight [Cout][Cin][3][3] fp32 -> B operand of the implicit GEMM:
 * mode 0: [Cout][(ky,kx,ci)] (forward);  mode 1: [Cin][(2-ky,2-kx,co)] (dX = conv with flipped filter) */
int countr_conv_weight_pack(const float* w, void* out, int Cout, int Cin, int mode, int bf16, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Density-head glue (models_mae_cross.py:80-100,189-197); activations NHWC 16-bit,
 * stats = [B][G][2] doubles (sum, sum of squares) produced by the conv epilogue (gn_stats).
 * ------------------------------------------------------------------------------------------ */
/* y = bilinear_x2( relu( GroupNorm(x) ) ),  [B][H][W][C] -> [B][2H][2W][C]
 * replaces: nn.GroupNorm(8,256)+ReLU (:82-83,87-88,92-93) + F.interpolate(x2, bilinear) (:189-194) */
int countr_gn_relu_upsample2x(const void* x, const double* stats, const float* gamma, const float* beta, void* y,
                              int B, int H, int W, int C, int G, float eps, int bf16, countr_stream_t stream);
/* out[b][p] = bias + sum_c w[c] * relu(GroupNorm(x))[b][p][c]   (GN + ReLU + Conv2d 1x1 C->1, :97-99) */
int countr_gn_relu_conv1x1(const void* x, const double* stats, const float* gamma, const float* beta, const float* w,
                           const float* bias, float* out, int B, int HW, int C, int G, float eps, int bf16,
                           countr_stream_t stream);
/* single-channel bilinear x2 [B][H][W] fp32 -> [B][2H][2W] of dtype code {0 fp32,1 fp16,2 bf16}
 * (last F.interpolate + squeeze, :195-197) */
int countr_upsample2x_f32(const float* x, void* y, int B, int H, int W, int out_dtype, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Exemplar encoder glue (models_mae_cross.py:47-71,157-177); sample n = b*S + s.
 * ------------------------------------------------------------------------------------------ */
/* decoder_proj1[0]: Conv2d(3,64,3,p=1) on boxes [B][K][3][HW][HW] (first S of K) -> NHWC 16-bit raw */
int countr_exemplar_conv1(const void* boxes, int dtype, int64_t sB, int64_t sK, int64_t sC, int64_t sH, int64_t sW,
                          const float* w, const float* bias, void* out, int B, int S, int HW, int Cout, int bf16,
                          countr_stream_t stream);
/* InstanceNorm2d(eps, no affine) + ReLU + MaxPool2d(2) (mode 0 -> y16 [N][H/2][W/2][C]) or
 * AdaptiveAvgPool2d(1) (mode 1 -> y32 [N][C] and/or y16 [N][C]); mean/rstd [N][C] optional
 * (when mean, rstd and scratch [32][N][C][2] fp32 are given, large maps take a pixel-parallel, deterministic
 * partial-sums -> finalize -> apply path) */
int countr_inorm_relu_pool(const void* x, void* y16, float* y32, float* mean, float* rstd, float* scratch, int N, int H, int W,
                           int C, float eps, int mode, int bf16, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backward-only kernels of the fine-tune step (autograd call sites listed in SURVEY.md §2.3).
 * ------------------------------------------------------------------------------------------ */
/* upsample_bilinear2d_backward of the last F.interpolate (models_mae_cross.py:195-196):
 * dy [B][2H][2W] (dtype code) -> dx fp32 [B][H][W] */
int countr_upsample2x_bwd(const void* dy, int dtype, float* dx, int B, int H, int W, countr_stream_t stream);
/* GroupNorm(8,256)+ReLU backward, pass A.  The incoming gradient is either the bilinear-x2 adjoint
 * of d_next [B][2H][2W][C] (stages 0-2) or dmap[b][p]*w1[c] (Conv2d 1x1 of decode_head3, then
 * dw1/db1 are accumulated too).  Writes dyh = dz*1[y>0]; accumulates dgamma/dbeta [C] and the
 * per-(image,group) sums gsum [B][G][2]. */
int countr_gn_relu_bwd_reduce(const void* raw, const double* stats, const float* gamma, const float* beta,
                              const void* d_next, const float* dmap, const float* w1, void* dyh, float* dgamma,
                              float* dbeta, float* dw1, float* db1, double* gsum, int B, int H, int W, int C, int G,
                              float eps, int bf16, countr_stream_t stream);
/* pass B: d_raw (gradient w.r.t. the conv output) and the conv bias gradient dbias [C] += */
int countr_gn_bwd_apply(const void* raw, const void* dyh, const double* stats, const double* gsum, const float* gamma,
                        void* d_raw, float* dbias, int B, int HW, int C, int G, float eps, int bf16,
                        countr_stream_t stream);
/* out[n] += sum_r x[r][n]  (Linear bias gradients); dtype code of x: 0 fp32, 1 fp16, 2 bf16 */
int countr_colsum(const void* x, int dtype, float* out, int64_t R, int N, int64_t ld, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Deferred weight / bias gradients of a backward pass in ONE launch each.
 * countr_grouped_dw: for every problem i  dw_i[n_out][k_in] (fp32, pre-zeroed or carrying earlier contributions) +=
 *   dy_i[tokens][n_out]^T x_i[tokens][k_in]  (16-bit operands, fp32 accumulate; one persistent tcgen05 kernel walks the
 *   concatenated, token-split tile list of all problems and accumulates through TMA reduce-add).
 * countr_grouped_colsum: out_i[cols] += column sums of x_i[rows][cols] (dtype 0 fp32, 1 fp16, 2 bf16).
 * replaces: the mm(dY^T, X) and sum(dY, 0) halves of autograd's addmm backward for every nn.Linear of the decoder
 *   (models_crossvit.py:55-57,77,80,104-108; models_mae_cross.py:39) and of models_mae_noct under FSC_pretrain.py.
 * ------------------------------------------------------------------------------------------ */
#define COUNTR_MAX_GROUP 32
typedef struct countr_dw_problem {
  const void* dy;   /* [tokens][ld_dy] 16-bit */
  const void* x;    /* [tokens][ld_x] 16-bit */
  float* dw;        /* [n_out][ld_dw] fp32 */
  int64_t ld_dy, ld_x, ld_dw;
  int32_t tokens, n_out, k_in, pad_;
} countr_dw_problem;
int countr_grouped_dw(const countr_dw_problem* probs, int n, int bf16, countr_stream_t stream);
typedef struct countr_colsum_problem {
  const void* x;
  float* out;
  int64_t rows, ld;
  int32_t cols, dtype;
} countr_colsum_problem;
int countr_grouped_colsum(const countr_colsum_problem* probs, int n, countr_stream_t stream);
/* _softmax_backward_data on materialised rows (self-attention backward of the FIM):
 * in place S <- P = exp(S - lse), dP <- dS = scale * P * (dP - sum_j P_j dP_j) */
int countr_softmax_bwd_rows(void* s_io, void* dp_io, const float* lse, int64_t rows, int L, float scale, int bf16,
                            countr_stream_t stream);
/* backward of countr_cross_attn_core: dq16 [B*L][D]; dk32/dv32 [B][S][D] (or [S][D]) += */
int countr_cross_attn_core_bwd(const void* q16, const float* k32, const float* v32, const float* probs, const void* do16,
                               void* dq16, float* dk32, float* dv32, int B, int L, int S, int D, int dh, float scale,
                               int bf16, int kv_broadcast, countr_stream_t stream);
/* backward of countr_inorm_relu_pool (InstanceNorm + ReLU + MaxPool2d(2) | avg-pool);
 * dbias [C] (optional) += sum d_raw = gradient of the preceding conv bias;
 * scratch (optional, [N][C][2] fp32) enables the pixel-parallel two-pass path for large maps */
int countr_inorm_relu_pool_bwd(const void* raw, const float* mean, const float* rstd, const void* dpool16,
                               const float* dpool32, void* d_raw, float* dbias, float* scratch, int N, int H, int W, int C,
                               int mode, int bf16, countr_stream_t stream);
/* decoder_proj1[0] weight gradient dw [64][3][3][3] += (direct reduction, K = 27) */
int countr_exemplar_conv1_dw(const void* boxes, int dtype, int64_t sB, int64_t sK, int64_t sC, int64_t sH, int64_t sW,
                             const void* d_raw, float* dw, int B, int S, int HW, int bf16, countr_stream_t stream);
/* [Cout][9][Cin] (implicit-GEMM dW layout) -> Conv2d.weight.grad layout [Cout][Cin][3][3] */
int countr_conv_dw_unpack(const float* src, float* dst, int Cout, int Cin, countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MAE pre-training glue (models_mae_noct.py).
 * ------------------------------------------------------------------------------------------ */
/* dst[b][j][:] = src[b][idx[b][j]][:]; rows of row_bytes (multiple of 16).  torch.gather of random_masking
 * (models_mae_noct.py:124) and, with the shuffle indices, its backward. */
int countr_gather_rows(const void* src, const int64_t* idx, void* dst, int B, int n_src, int n_dst, int row_bytes,
                       countr_stream_t stream);
/* MAE decoder input (:158-165): out[b][l] = (ids_restore[b][l] < Lk ? xk[b][ids_restore[b][l]] : mask_token) + pos[l] */
int countr_mae_unshuffle(const float* xk, const int64_t* ids_restore, const float* mask_token, const float* pos, float* out,
                         int B, int L, int Lk, int D, countr_stream_t stream);
/* forward_loss (:177-198): loss = mean_patches mean((pred - patchify(img))^2) (optionally per-patch normalised
 * targets); also writes dpred = d loss / d pred (fp32, optional) */
int countr_mae_loss(const float* pred, const void* img, int dtype, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                    float* loss, float* dpred, int B, int C, int H, int W, int P, int norm_pix, countr_stream_t stream);
/* dst16 = 16-bit(src * *scale_ptr): the upstream gradient scalar stays on the device (no host sync) */
int countr_cast_scaled_f32_to_16(const float* src, const float* scale_ptr, void* dst, int64_t n, int bf16,
                                 countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Script-side pieces of the fine-tune step as single kernels (used by countr_b200.train.FineTuner).
 * ------------------------------------------------------------------------------------------ */
/* Device state block of the fused step (fp32[8], owned by the caller): [0] loss scale, [1] GradScaler growth tracker,
 * [2] found_inf of the last countr_grad_stats (0/1), [3] global L2 norm of the UNSCALED gradients (util/misc.py:289-301
 * get_grad_norm_), [4] learning rate, [5] step counter (drives the device-side mask draw).  Keeping these on the device
 * lets a captured CUDA graph follow the lr schedule (lr_sched.adjust_learning_rate, FSC_finetune_cross.py:270) and the
 * dynamic loss scale (torch.cuda.amp.GradScaler, util/misc.py:260-280) without re-capture. */
#define COUNTR_ST_SCALE 0
#define COUNTR_ST_GROWTH 1
#define COUNTR_ST_FOUND_INF 2
#define COUNTR_ST_GRAD_NORM 3
#define COUNTR_ST_LR 4
#define COUNTR_ST_STEP 5

/* FSC_finetune_cross.py:290-303 in one launch.  out / gt: [B][H][W] of dtype code 0 fp32 / 1 fp16 / 2 bf16.
 *   mask != NULL : fp32 keep-mask, element (b, i) at mask[b * mask_bstride + i] (mask_bstride 0 = one [H][W] mask tiled over
 *                  the batch, as np.tile does at :291-293)
 *   mask == NULL : the mask is drawn on the device, Bernoulli(keep_prob) per pixel, the same draw for every image of the
 *                  batch, from a counter-based generator keyed by (seed, state[5], pixel); mask_out (optional, uint8
 *                  [H][W]) receives it
 *   result[0] = sum((out-gt)^2 * mask / (H*W)) / B;  result[1] = batch MAE and result[2] = batch MSE of the counts
 *   counts (optional) [B][2] = {out.sum()/60, gt.sum()/60} per image (:298-303)
 *   dout (optional, fp32 [B][H][W]) = d(loss * scale)/d out with scale = state[0] if state != NULL else grad_scale
 * scratch: countr_finetune_loss_scratch_bytes(B) bytes, 8-byte aligned, its first 8 bytes zero before the first call.
 * All reductions run in a fixed order (bit-reproducible). */
int countr_finetune_loss(const void* out, int out_dtype, const void* gt, int gt_dtype, const float* mask, int64_t mask_bstride,
                         uint64_t seed, float keep_prob, const float* state, float grad_scale, float* dout, uint8_t* mask_out,
                         void* scratch, float* result, float* counts, int B, int H, int W, countr_stream_t stream);
int64_t countr_finetune_loss_scratch_bytes(int B);
/* GradScaler.unscale_ inf check + get_grad_norm_ over the flat (still scaled) gradient arena: state[2] = any non-finite,
 * state[3] = ||grad||_2 / state[0].  scratch: countr_grad_stats_scratch_bytes() bytes, first 8 bytes zero before the first call. */
int countr_grad_stats(const float* grad, int64_t n, void* scratch, float* state, countr_stream_t stream);
int64_t countr_grad_stats_scratch_bytes(void);
/* unscale + torch.optim.AdamW (decoupled weight decay, bias correction) over every tensor of a flat arena in one launch,
 * then the per-parameter step counters and GradScaler.update().  Skipped entirely (parameters, moments and counters
 * untouched, scale *= backoff_factor) when state[2] != 0; lr = state[4], 1/scale = 1/state[0].
 * tensors: device array of {float* param; int64 grad_off; int64 moment_off; int64 numel; float weight_decay; int step_idx;
 * int flag_idx; int pad} (48 bytes); a tensor with flag_idx k > 0 is updated only when flags[k-1] != 0 (flags: device
 * floats, e.g. the tail of the all-reduced gradient arena saying which optional parameter groups were used on any rank —
 * DDP(find_unused_parameters=True) semantics, FSC_finetune_cross.py:230).
 * chunks: device array of {int tensor, int chunk_of_1024}; step: device fp32 per-parameter step counters.
 * growth_interval <= 0 keeps the loss scale static. */
int countr_adamw_update(const void* tensors, int num_tensors, const void* chunks, int num_chunks, const float* grad,
                        const float* flags, float* exp_avg, float* exp_avg_sq, float* step, float* state, float beta1, float beta2,
                        float eps, float growth_factor, float backoff_factor, int growth_interval, countr_stream_t stream);

/* Sliding-window evaluation blend (demo.py:124-160; FSC_test_cross(few-shot).py:322-349): outs [nw][H][Wwin] are the
 * density maps of the windows at columns starts[i] (visited left to right); density [H][W] receives the running
 * average the reference's loop produces (overlap -> (old + new) / 2, new columns -> new). */
int countr_window_blend(const void* outs, int dtype, const int32_t* starts, int nw, int H, int Wwin, int W, float* density,
                        countr_stream_t stream);

/* Ground-truth density-map synthesis on the GPU (util/FSC147.py:262-273 ResizeTrainImage no-augmentation
 * path, :326-331 ResizeValImage): dots [B][n_max][2] (x, y; float64 as the annotation files hold them), counts [B];
 * a 1 is written at (min(canvas_h-1, int(y*scale_h)), min(canvas_w-1, int(x*scale_w))) of the resized canvas, the H x W
 * window at (y0, x0) is kept, filtered with scipy.ndimage.gaussian_filter's arithmetic (separable, axis 0 then axis 1,
 * double accumulation, 'reflect' boundary; weights[0..radius] = the normalised float64 half kernel, centre first) and
 * multiplied by gain (60).  tmp, out: [B][H][W] fp32. */
int countr_density_from_dots(const double* dots, const int32_t* counts, int B, int n_max, double scale_h, double scale_w,
                             int canvas_h, int canvas_w, int y0, int x0, int H, int W, const double* weights, int radius,
                             float gain, float* tmp, float* out, countr_stream_t stream);

/* Batched refresh of the 16-bit operand copies of fp32 master weights (one launch per optimizer step instead of one per
 * tensor).  entries: device array of {const float* src; uint16* dst; int64 kind, R, C, pad} with kind 0 = cast of R*C
 * elements, 1 = cast + transpose [R][C] -> [C][R], 2 / 3 = countr_conv_weight_pack mode 0 / 1 with Cout = R, Cin = C;
 * blk_prefix[e] (n_entries + 1 values) = first 256-thread block of entry e, entry e taking
 * countr_weight_refresh_blocks(kind, R, C) blocks (2048 elements per block for kind 0, one 64x64 tile per block for kind 1,
 * one (co, 128 input channels) slab per block for kind 2, one 64x64 tile of the filter viewed as [Cout][Cin*9] for kind 3). */
int countr_weight_refresh(const void* entries, const int32_t* blk_prefix, int n_entries, int total_blocks, int bf16,
                          countr_stream_t stream);
int64_t countr_weight_refresh_blocks(int kind, int64_t R, int64_t C);   /* host-only helper, -1 on bad arguments */

/* Exemplar crops (util/FSC147.py:285-298, 343-351): out[b][s] = Resize((out_hw, out_hw))(img[b][:, y1:y2+1, x1:x2+1]) with
 * torchvision 0.14.1's tensor semantics (bilinear, align_corners=False, no antialias).  img: fp32, element strides
 * (sb, sc, sh, sw); rects: int32 [B][S][4] = (y1, x1, y2, x2), inclusive, clipped to the image; out: fp32 [B][S][C][out_hw][out_hw]. */
int countr_crop_resize_boxes(const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int32_t* rects, float* out,
                             int B, int S, int C, int H, int W, int out_hw, countr_stream_t stream);
/* Same resize arithmetic with a rectangular output: out fp32 [B][S][C][out_h][out_w].  Used by the evaluation path that tiles an
 * image with tiny exemplars into 3 x 3 crops and blows each crop up to the full frame (FSC_test_cross(few-shot).py:273-285:
 * TF.crop + transforms.Resize((h, w))). */
int countr_crop_resize(const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int32_t* rects, float* out, int B,
                       int S, int C, int H, int W, int out_h, int out_w, countr_stream_t stream);
/* out[0] += sum over n_rects inclusive pixel rectangles (y1, x1, y2, x2; clipped like python slices) of map[y][x] / divisor —
 * the exemplar-box density mass of the test-time normalisation (FSC_test_cross(few-shot).py:353-359, demo.py:162-169). */
int countr_rect_mass(const float* map, int H, int W, const int32_t* rects, int n_rects, float divisor, float* out,
                     countr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training-time augmentations on the device (util/FSC147.py:133-143, 371-374, 176-180).  The random PARAMETERS are the caller's;
 * given the same parameters the pixel arithmetic is numpy's / torchvision's.  Images: fp32 [B][3][H][W], contiguous.
 *   countr_aug_noise_clamp   out = clamp(img + N(0, stddev), 0, 1)   (counter-based generator keyed by `seed`; :133-137)
 *   countr_aug_color_jitter  torchvision ColorJitter in place: pass j (0..3) applies ops[j][b] in {0 brightness, 1 contrast,
 *                            2 saturation, 3 hue, -1 none} with factors[j][b] to image b (transforms.functional.adjust_*, :372);
 *                            scratch: B doubles
 *   countr_aug_gaussian_blur torchvision GaussianBlur(kernel_size=(kx, ky)) with sigma[b] (reflect padding, :373); tmp: image-sized
 *   countr_aug_hflip         TF.hflip of the images whose flag is set, `planes` channels per image (image and density map, :176-180)
 *   countr_aug_mosaic        the 2 x 2 self-/cross-image collage (:183-262): quadrant t (0 top-left, 1 bottom-left, 2 top-right,
 *                            3 bottom-right) = Resize((rl, rl))(TF.crop(src[t].img, top, left, length, length)) with rl = 192 + 2 bl,
 *                            rows/columns bl..rl-bl kept and the 2 bl-wide seams blended with the reference's recurrence and
 *                            index offsets (:241-245, :257-261); out: fp32 [C][2(rl-2bl)][2(rl-2bl)].  `src` is a HOST array of 4.
 *   countr_aug_mosaic_dots   the collage's dot map (:190-196, :228-232): dots [.][2] float64 (x, y) on the device, quadrant t uses
 *                            dots[dot_begin .. dot_begin+dot_count) (count 0 for an image of another class); canvas [2(rl-2bl)]^2
 *   countr_density_filter    scipy.ndimage.gaussian_filter + gain on an existing dot map (the tail of countr_density_from_dots, :265-269)
 *   countr_aug_affine        bilinear (order 1), zero-border warp of fp32 [C][H][W] by the 2 x 3 matrix `inverse` (output pixel ->
 *                            source position; HOST array of 6 doubles) — the image half of iaa.Affine (:151-159)
 *   countr_aug_affine_dots   the key points through the 2 x 3 `forward` matrix and the reference's dot map of the survivors
 *                            (:146-149, :162-166); canvas [H][W]
 * imgaug (0.4.0, requirements.txt) is not installable here: the matrix composition in countr_b200/data.py:affine_matrix restates
 * its published order and is NOT pinned against imgaug itself; cv2's 1/32-pixel coordinate quantisation is not reproduced.
 * ------------------------------------------------------------------------------------------ */
int countr_aug_noise_clamp(const float* img, float* out, int64_t n, float stddev, uint64_t seed, countr_stream_t stream);
int countr_aug_color_jitter(float* img, const int32_t* ops, const float* factors, double* scratch, int B, int H, int W,
                            countr_stream_t stream);
int countr_aug_gaussian_blur(const float* img, float* tmp, float* out, const float* sigma, int B, int H, int W, int kx, int ky,
                             countr_stream_t stream);
int countr_aug_hflip(const float* in, float* out, const int32_t* flags, int B, int planes, int H, int W, countr_stream_t stream);
typedef struct countr_mosaic_src {
  const float* img;             /* fp32 [C][H][W] view of the resized source image of this quadrant (device) */
  int64_t sc, sh, sw;           /* its element strides */
  int32_t H, W;                 /* new_TH, new_TW */
  int32_t top, left, length;    /* start_H, start_W, length of the square crop */
  int32_t dot_begin, dot_count; /* this quadrant's annotated points */
  int32_t pad_;
  double scale_h, scale_w;      /* Tscale_factor_h / Tscale_factor_w */
} countr_mosaic_src;
int countr_aug_mosaic(const countr_mosaic_src* src, int rl, int bl, int C, float* out, countr_stream_t stream);
int countr_aug_mosaic_dots(const countr_mosaic_src* src, int rl, int bl, const double* dots, float* canvas, countr_stream_t stream);
int countr_density_filter(const float* canvas, float* tmp, float* out, int B, int H, int W, const double* weights, int radius, float gain,
                          countr_stream_t stream);
int countr_aug_affine(const float* img, float* out, int C, int H, int W, const double* inverse, countr_stream_t stream);
int countr_aug_affine_dots(const double* dots, int n, double scale_h, double scale_w, int H, int W, const double* forward, float* canvas,
                           countr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COUNTR_B200_H_ */
