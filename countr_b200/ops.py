"""Thin tensor-level wrappers over the C ABI (include/countr_b200.h).

Every function enqueues hand-written sm_100a kernels on torch's current CUDA stream and
returns immediately.  PyTorch is used for memory ownership only: no ATen math runs here.
"""
import ctypes

import torch

from ._lib import GemmDesc, check, lib

F16 = torch.float16
BF16 = torch.bfloat16
_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _is_bf16(t):
    return 1 if t.dtype == BF16 else 0


LAUNCHES = [0]  # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n=1):
    LAUNCHES[0] += n


def gemm(a, b, c, M, N, K, *, lda, ldb, ldc, a_mn=False, b_mn=False, nb1=1, nb2=1,
         sa=(0, 0), sb=(0, 0), sc=(0, 0), bias=None, act=0, aux=None, ldaux=0, residual=None, ldr=0,
         res_mod=0, alpha=1.0, atomic=False, split_k=1, bn=0, conv=None, gn_stats=None, conv_dw=None, cluster=0, pair=0,
         ln_x16=None, ln_stats=None, ln_colsum=None, ln_dim=0, ln_eps=0.0):
    """C = epilogue(alpha * A @ B^T); see countr_gemm_desc for the layout rules."""
    assert a.dtype in (F16, BF16) and b.dtype == a.dtype
    d = GemmDesc()
    d.a, d.b = a.data_ptr(), b.data_ptr()
    d.lda, d.sa1, d.sa2 = lda, sa[0], sa[1]
    d.ldb, d.sb1, d.sb2 = ldb, sb[0], sb[1]
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.M, d.N, d.K = M, N, K
    d.nb1, d.nb2 = nb1, nb2
    d.bf16 = _is_bf16(a)
    d.bn, d.split_k, d.cluster, d.cta_pair = bn, split_k, cluster, pair
    if conv is not None:
        d.conv_h, d.conv_w, d.conv_cin, d.conv_bx, d.conv_by = conv
    if conv_dw is not None:
        d.conv_h, d.conv_w, d.conv_bx, d.conv_by, d.conv_batch = conv_dw
        d.conv_dw = 1
    d.c = c.data_ptr()
    d.ldc, d.sc1, d.sc2 = ldc, sc[0], sc[1]
    d.out_f32 = 1 if c.dtype == torch.float32 else 0
    d.atomic = int(atomic)
    d.alpha = alpha
    d.bias = bias.data_ptr() if bias is not None else None
    d.act = act
    d.aux = aux.data_ptr() if aux is not None else None
    d.ldaux = ldaux
    d.residual = residual.data_ptr() if residual is not None else None
    d.ldr = ldr
    d.res_mod = res_mod
    d.gn_stats = gn_stats.data_ptr() if gn_stats is not None else None
    if ln_stats is not None:        # LayerNorm folded into this GEMM's epilogue (producer: ln_x16, consumer: ln_colsum)
        d.ln_stats = ln_stats.data_ptr()
        d.ln_x16 = ln_x16.data_ptr() if ln_x16 is not None else None
        d.ld_x16 = ln_x16.stride(0) if ln_x16 is not None else 0
        d.ln_colsum = ln_colsum.data_ptr() if ln_colsum is not None else None
        d.ln_dim, d.ln_eps = ln_dim, ln_eps
    check(lib().countr_gemm(ctypes.byref(d), _stream()))
    _count()
    return c


def linear(x16, w16, out, bias=None, act=0, aux=None, residual=None, res_mod=0, alpha=1.0, bn=0, ln_x16=None, ln_stats=None,
           ln_colsum=None, ln_eps=0.0):
    """out[M,N] = epi(x16[M,K] @ w16[N,K]^T); residual is fp32 [M or res_mod, N] (may alias out).
    LayerNorm folding (countr_gemm_desc): producer — ln_x16 [M,N] 16-bit + ln_stats [M,8,2] fp32 receive a 16-bit copy of `out` and
    its row-statistics partials; consumer — x16 is such a copy, w16 = W * diag(gamma), ln_colsum[n] = sum_k w16[n,k], bias = b + W beta."""
    M, K = x16.shape
    N = w16.shape[0]
    assert w16.shape[1] == K and out.shape == (M, N)
    return gemm(x16, w16, out, M, N, K, lda=x16.stride(0), ldb=w16.stride(0), ldc=out.stride(0), bias=bias,
                act=act, aux=aux, ldaux=aux.stride(0) if aux is not None else 0, residual=residual,
                ldr=residual.stride(0) if residual is not None else 0, res_mod=res_mod, alpha=alpha, bn=bn, ln_x16=ln_x16,
                ln_stats=ln_stats, ln_colsum=ln_colsum, ln_dim=K, ln_eps=ln_eps)


def conv3x3(x16, w16, out, bias=None, gn_stats=None, pair=0):
    """NHWC implicit-GEMM conv: x16 [B,H,W,Cin], w16 [Cout, 9*Cin] (ky,kx,ci), out [B,H,W,Cout]."""
    B, H, W, Cin = x16.shape
    Cout = w16.shape[0]
    bx = next((t for t in (128, 64, 32, 16, 8) if W % t == 0), 8)
    by = 128 // bx
    return gemm(x16, w16, out, B * H * W, Cout, 9 * Cin, lda=Cin, ldb=w16.stride(0), ldc=Cout, nb1=B,
                sa=(H * W * Cin, 0), sc=(H * W * Cout, 0), bias=bias, conv=(H, W, Cin, bx, by), gn_stats=gn_stats, pair=pair)


def conv3x3_dw(dy16, x16, dw32, split_k=0, pair=0):
    """dw32 [Cout, 9*Cin] fp32 (tap-major, PRE-ZEROED or accumulating) += dY^T (*) X over all pixels."""
    B, H, W, Cout = dy16.shape
    Cin = x16.shape[-1]
    bx = next((t for t in (64, 32, 16, 8) if W % t == 0), 8)
    by = 64 // bx
    tiles = ((W + bx - 1) // bx) * ((H + by - 1) // by)
    kblocks = B * tiles
    if split_k <= 0:
        out_tiles = 9 * ((Cout + 127) // 128) * ((Cin + 255) // 256)
        split_k = max(1, min(kblocks, 148 // out_tiles))
    return gemm(dy16, x16, dw32, Cout, Cin, kblocks * 64, lda=Cout, ldb=Cin, ldc=9 * Cin, a_mn=True, b_mn=True, nb2=9,
                sc=(0, Cin), atomic=True, split_k=split_k, conv_dw=(H, W, bx, by, B), pair=pair)


def layernorm_fwd(x, gamma, beta, eps, y16=None, y32=None, mean=None, rstd=None):
    rows, D = x.shape
    check(lib().countr_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y16), _ptr(y32), _ptr(mean), _ptr(rstd),
                                     rows, D, eps, _is_bf16(y16) if y16 is not None else 0, _stream()))
    _count()


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma=None, dbeta=None, accumulate=False, dx16=None, dx_colsum=None, jobs=None):
    """jobs (backward.GradJobs): dgamma / dbeta / dx_colsum are not accumulated with atomics by this kernel; it writes per-block
    partial sums and three column-sum jobs are recorded instead (summed by the grouped launch at the end of the backward), which
    also lets the kernel run four blocks per SM."""
    rows, D = x.shape
    partials = None
    if jobs is not None and dgamma is not None:
        nblk = int(lib().countr_layernorm_bwd_blocks(rows))
        partials = torch.empty(3, nblk, D, dtype=torch.float32, device=x.device)
        jobs.dB(partials[0], dgamma)
        jobs.dB(partials[1], dbeta)
        if dx_colsum is not None:
            jobs.dB(partials[2], dx_colsum)
        dgamma = dbeta = dx_colsum = None
    check(lib().countr_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dx16), _ptr(dgamma),
                                     _ptr(dbeta), _ptr(dx_colsum), _ptr(partials), rows, D, int(accumulate),
                                     _is_bf16(dx16) if dx16 is not None else 0, _stream()))
    _count()


def attention_fwd(qkv16, out16, B, L, H, dh, scale, lse=None):
    check(lib().countr_attention_fwd(_ptr(qkv16), _ptr(out16), _ptr(lse), B, L, H, dh, scale, _is_bf16(qkv16), _stream()))
    _count()


def attention_bwd(qkv16, out16, dout16, lse, dqkv16, B, L, H, dh, scale):
    ws = None
    nbytes = int(lib().countr_attention_bwd_workspace_bytes(B, L, H, dh))
    if nbytes:
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=qkv16.device)     # fp32 dQ accumulator (head_dim 64)
    check(lib().countr_attention_bwd(_ptr(qkv16), _ptr(out16), _ptr(dout16), _ptr(lse), _ptr(dqkv16), _ptr(ws), B, L, H, dh, scale,
                                     _is_bf16(qkv16), _stream()))
    _count(3 if nbytes else 1)


def cross_attn_core(q16, k32, v32, out16, B, L, S, D, dh, scale, probs=None, kv_broadcast=False):
    check(lib().countr_cross_attn_core(_ptr(q16), _ptr(k32), _ptr(v32), _ptr(out16), _ptr(probs), B, L, S, D, dh, scale,
                                       _is_bf16(q16), int(kv_broadcast), _stream()))
    _count()


def cast16(src32, dst16, scale=1.0):
    check(lib().countr_cast_f32_to_16(_ptr(src32), _ptr(dst16), src32.numel(), scale, _is_bf16(dst16), _stream()))
    _count()


def weight_refresh(entries_dev, prefix_dev, n_entries, total_blocks, bf16=False):
    """One launch for a whole table of cast / transpose / conv-pack jobs (countr_weight_refresh)."""
    check(lib().countr_weight_refresh(_ptr(entries_dev), _ptr(prefix_dev), n_entries, total_blocks, int(bf16), _stream()))
    _count()


def cast16_transpose(src32, dst16):
    R, C = src32.shape
    check(lib().countr_cast_transpose_f32_to_16(_ptr(src32), _ptr(dst16), R, C, _is_bf16(dst16), _stream()))
    _count()


def patchify(img, out16, P):
    B, C, H, W = img.shape
    sb, sc, sh, sw = img.stride()
    check(lib().countr_patchify(_ptr(img), _DTYPE_CODE[img.dtype], sb, sc, sh, sw, _ptr(out16), B, C, H, W, P,
                                _is_bf16(out16), _stream()))
    _count()


def conv_weight_pack(w32, out16, mode=0):
    Cout, Cin = w32.shape[0], w32.shape[1]
    check(lib().countr_conv_weight_pack(_ptr(w32), _ptr(out16), Cout, Cin, mode, _is_bf16(out16), _stream()))
    _count()


def gn_relu_upsample2x(x16, stats, gamma, beta, y16, G, eps):
    B, H, W, C = x16.shape
    check(lib().countr_gn_relu_upsample2x(_ptr(x16), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(y16), B, H, W, C, G, eps,
                                          _is_bf16(x16), _stream()))
    _count()


def gn_relu_conv1x1(x16, stats, gamma, beta, w, bias, out32, G, eps):
    B, H, W, C = x16.shape
    check(lib().countr_gn_relu_conv1x1(_ptr(x16), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(w), _ptr(bias), _ptr(out32), B,
                                       H * W, C, G, eps, _is_bf16(x16), _stream()))
    _count()


def upsample2x_f32(x32, y):
    B, H, W = x32.shape
    check(lib().countr_upsample2x_f32(_ptr(x32), _ptr(y), B, H, W, _DTYPE_CODE[y.dtype], _stream()))
    _count()


def exemplar_conv1(boxes, S, w, bias, out16):
    B, K, C, H, W = boxes.shape
    assert C == 3 and H == W and S <= K
    sB, sK, sC, sH, sW = boxes.stride()
    check(lib().countr_exemplar_conv1(_ptr(boxes), _DTYPE_CODE[boxes.dtype], sB, sK, sC, sH, sW, _ptr(w), _ptr(bias),
                                      _ptr(out16), B, S, H, w.shape[0], _is_bf16(out16), _stream()))
    _count()


def inorm_relu_pool(x16, mode, eps, y16=None, y32=None, mean=None, rstd=None, scratch=None):
    N, H, W, C = x16.shape
    split_path = mode == 0 and mean is not None and rstd is not None and scratch is not None and H * W >= 256
    assert scratch is None or scratch.numel() >= 64 * N * C
    check(lib().countr_inorm_relu_pool(_ptr(x16), _ptr(y16), _ptr(y32), _ptr(mean), _ptr(rstd), _ptr(scratch), N, H, W, C, eps, mode,
                                       _is_bf16(x16), _stream()))
    _count(3 if split_path else 1)

def zero_(t):
    check(lib().countr_memset_zero(_ptr(t), t.numel() * t.element_size(), _stream()))
    return t


def upsample2x_bwd(dy, dx32):
    B, H, W = dx32.shape
    check(lib().countr_upsample2x_bwd(_ptr(dy), _DTYPE_CODE[dy.dtype], _ptr(dx32), B, H, W, _stream()))
    _count()


def gn_relu_bwd_reduce(raw16, stats, gamma, beta, dyh16, dgamma, dbeta, gsum, G, eps, d_next=None, dmap=None, w1=None,
                       dw1=None, db1=None):
    B, H, W, C = raw16.shape
    check(lib().countr_gn_relu_bwd_reduce(_ptr(raw16), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(d_next), _ptr(dmap), _ptr(w1),
                                          _ptr(dyh16), _ptr(dgamma), _ptr(dbeta), _ptr(dw1), _ptr(db1), _ptr(gsum), B, H, W, C, G,
                                          eps, _is_bf16(raw16), _stream()))
    _count()


def gn_bwd_apply(raw16, dyh16, stats, gsum, gamma, d_raw16, dbias, G, eps):
    B, H, W, C = raw16.shape
    check(lib().countr_gn_bwd_apply(_ptr(raw16), _ptr(dyh16), _ptr(stats), _ptr(gsum), _ptr(gamma), _ptr(d_raw16), _ptr(dbias), B,
                                    H * W, C, G, eps, _is_bf16(raw16), _stream()))
    _count()


def colsum(x, out32):
    """out32[n] += sum over all leading dims of x[..., n]"""
    N = x.shape[-1]
    R = x.numel() // N
    check(lib().countr_colsum(_ptr(x), _DTYPE_CODE[x.dtype], _ptr(out32), R, N, N, _stream()))
    _count()


class _DwProblem(ctypes.Structure):
    _fields_ = [("dy", ctypes.c_void_p), ("x", ctypes.c_void_p), ("dw", ctypes.c_void_p), ("ld_dy", ctypes.c_int64),
                ("ld_x", ctypes.c_int64), ("ld_dw", ctypes.c_int64), ("tokens", ctypes.c_int32), ("n_out", ctypes.c_int32),
                ("k_in", ctypes.c_int32), ("pad_", ctypes.c_int32)]


class _ColsumProblem(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("out", ctypes.c_void_p), ("rows", ctypes.c_int64), ("ld", ctypes.c_int64),
                ("cols", ctypes.c_int32), ("dtype", ctypes.c_int32)]


MAX_GROUP = 32


def grouped_dw(problems):
    """problems: list of (dy16 [tokens, n_out], x16 [tokens, k_in], dw32 [n_out, k_in]);  dw32 += dy16^T @ x16 for all of them in
    one persistent launch per 32 problems (countr_grouped_dw)."""
    for i in range(0, len(problems), MAX_GROUP):
        chunk = problems[i:i + MAX_GROUP]
        arr = (_DwProblem * len(chunk))()
        for j, (dy, x, dw) in enumerate(chunk):
            assert dy.dtype in (F16, BF16) and x.dtype == dy.dtype and dw.dtype == torch.float32
            assert dy.shape[0] == x.shape[0] and dw.shape == (dy.shape[1], x.shape[1]) and dy.stride(1) == 1 and x.stride(1) == 1
            assert dw.is_contiguous()
            arr[j] = _DwProblem(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), dy.stride(0), x.stride(0), dw.stride(0), dy.shape[0],
                                dy.shape[1], x.shape[1], 0)
        check(lib().countr_grouped_dw(arr, len(chunk), _is_bf16(chunk[0][0]), _stream()))
        _count()


def grouped_colsum(problems):
    """problems: list of (x [rows, cols], out32 [cols]);  out32 += column sums, one launch per 32 problems."""
    for i in range(0, len(problems), MAX_GROUP):
        chunk = problems[i:i + MAX_GROUP]
        arr = (_ColsumProblem * len(chunk))()
        for j, (x, out) in enumerate(chunk):
            cols = x.shape[-1]
            assert x.is_contiguous() and out.dtype == torch.float32 and out.numel() == cols
            arr[j] = _ColsumProblem(x.data_ptr(), out.data_ptr(), x.numel() // cols, cols, cols, _DTYPE_CODE[x.dtype])
        check(lib().countr_grouped_colsum(arr, len(chunk), _stream()))
        _count()


def softmax_bwd_rows(s16, dp16, lse, scale):
    L = s16.shape[-1]
    check(lib().countr_softmax_bwd_rows(_ptr(s16), _ptr(dp16), _ptr(lse), s16.numel() // L, L, scale, _is_bf16(s16), _stream()))
    _count()


def cross_attn_core_bwd(q16, k32, v32, probs, do16, dq16, dk32, dv32, B, L, S, D, dh, scale, kv_broadcast=False):
    check(lib().countr_cross_attn_core_bwd(_ptr(q16), _ptr(k32), _ptr(v32), _ptr(probs), _ptr(do16), _ptr(dq16), _ptr(dk32),
                                           _ptr(dv32), B, L, S, D, dh, scale, _is_bf16(q16), int(kv_broadcast), _stream()))
    _count()


def inorm_relu_pool_bwd(raw16, mean, rstd, d_raw16, mode, dpool16=None, dpool32=None, dbias=None, scratch=None):
    N, H, W, C = raw16.shape
    check(lib().countr_inorm_relu_pool_bwd(_ptr(raw16), _ptr(mean), _ptr(rstd), _ptr(dpool16), _ptr(dpool32), _ptr(d_raw16),
                                           _ptr(dbias), _ptr(scratch), N, H, W, C, mode, _is_bf16(raw16), _stream()))
    _count(3 if (scratch is not None and mode == 0 and H * W >= 256) else 1)


def exemplar_conv1_dw(boxes, S, d_raw16, dw32):
    B, K, C, H, W = boxes.shape
    sB, sK, sC, sH, sW = boxes.stride()
    check(lib().countr_exemplar_conv1_dw(_ptr(boxes), _DTYPE_CODE[boxes.dtype], sB, sK, sC, sH, sW, _ptr(d_raw16), _ptr(dw32), B, S,
                                         H, _is_bf16(d_raw16), _stream()))
    _count()


def conv_dw_unpack(src32, dst32, Cout, Cin):
    check(lib().countr_conv_dw_unpack(_ptr(src32), _ptr(dst32), Cout, Cin, _stream()))
    _count()


def gather_rows(src, idx, dst):
    """dst[b, j] = src[b, idx[b, j]] for [B, n, D] row tensors of any dtype (rows must be a multiple of 16 bytes)."""
    B, n_src = src.shape[0], src.shape[1]
    n_dst = idx.shape[1]
    row_bytes = src.shape[2] * src.element_size()
    check(lib().countr_gather_rows(_ptr(src), _ptr(idx), _ptr(dst), B, n_src, n_dst, row_bytes, _stream()))
    _count()


def mae_unshuffle(xk, ids_restore, mask_token, pos, out):
    B, L, D = out.shape
    check(lib().countr_mae_unshuffle(_ptr(xk), _ptr(ids_restore), _ptr(mask_token), _ptr(pos), _ptr(out), B, L, xk.shape[1], D, _stream()))
    _count()


def mae_loss(pred, imgs, loss, dpred, P, norm_pix):
    B, C, H, W = imgs.shape
    sb, sc, sh, sw = imgs.stride()
    check(lib().countr_mae_loss(_ptr(pred), _ptr(imgs), _DTYPE_CODE[imgs.dtype], sb, sc, sh, sw, _ptr(loss), _ptr(dpred), B, C, H, W, P,
                                int(bool(norm_pix)), _stream()))
    _count()


def cast16_scaled(src32, scale_tensor, dst16):
    check(lib().countr_cast_scaled_f32_to_16(_ptr(src32), _ptr(scale_tensor), _ptr(dst16), src32.numel(), _is_bf16(dst16), _stream()))
    _count()
