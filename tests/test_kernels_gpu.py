"""GPU unit tests of the individual sm_100a kernels against plain fp32 torch references.

These are floating-point kernels, so the reference for each op is the fp32 torch expression of
the same arithmetic on the SAME 16-bit-rounded operands; tolerances are written per test.
The model-level parity tests (CUDA path vs. oracle vs. committed goldens) live in test_parity_gpu.py.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _describe(got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    bad = err > (1e-2 * ref.abs().max().clamp_min(1e-6))
    msg = [f"max|err|={err.max().item():.4g} max|ref|={ref.abs().max().item():.4g} bad={bad.float().mean().item():.4f}"]
    if bad.any() and got.dim() == 2:
        rows = bad.any(1).nonzero().flatten()[:12].tolist()
        cols = bad.any(0).nonzero().flatten()[:12].tolist()
        msg.append(f"first bad rows {rows} cols {cols}")
        r0 = rows[0]
        c0 = bad[r0].nonzero().flatten()[0].item()
        msg.append(f"got[{r0},{c0}:{c0+8}]={got[r0, c0:c0+8].tolist()} ref={ref[r0, c0:c0+8].tolist()}")
    return " | ".join(msg)


def assert_close(got, ref, rtol, what=""):
    """max|got-ref| <= rtol * max|ref|  (matrix-level relative tolerance)"""
    assert torch.isfinite(got.float()).all(), f"{what}: non-finite output"
    scale = ref.float().abs().max().clamp_min(1e-6)
    err = (got.float() - ref.float()).abs().max()
    assert err <= rtol * scale, f"{what}: {_describe(got, ref)}"


def rel_close(a, b, tol=2e-3):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item() < tol


def _rand16(shape, dev, dtype=torch.float16, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(dev)


# ------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (256, 256, 128, 128), (128, 64, 64, 64), (128, 256, 256, 256), (384, 192, 192, 192),
    (200, 136, 72, 0), (4608, 768, 768, 0), (4608, 2304, 768, 0), (1152, 3072, 768, 0), (24, 512, 512, 0),
])
def test_gemm_kmajor(cuda, M, N, K, bn):
    from countr_b200 import ops
    a = _rand16((M, K), cuda, seed=1)
    b = _rand16((N, K), cuda, seed=2)
    c = torch.empty(M, N, device=cuda, dtype=torch.float32)
    ops.gemm(a, b, c, M, N, K, lda=K, ldb=K, ldc=N, bn=bn)
    torch.cuda.synchronize()
    assert_close(c, a.float() @ b.float().t(), 2e-5, f"gemm f32 {M}x{N}x{K} bn={bn}")
    c16 = torch.empty(M, N, device=cuda, dtype=torch.float16)
    ops.gemm(a, b, c16, M, N, K, lda=K, ldb=K, ldc=N, bn=bn)
    torch.cuda.synchronize()
    assert_close(c16, a.float() @ b.float().t(), 1e-3, f"gemm f16 {M}x{N}x{K} bn={bn}")


def test_gemm_bf16(cuda):
    from countr_b200 import ops
    M, N, K = 256, 384, 320
    a = _rand16((M, K), cuda, torch.bfloat16, 1)
    b = _rand16((N, K), cuda, torch.bfloat16, 2)
    c = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, b, c, M, N, K, lda=K, ldb=K, ldc=N)
    torch.cuda.synchronize()
    assert_close(c, a.float() @ b.float().t(), 8e-3, "gemm bf16")


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 192, 128), (512, 512, 4608), (200, 136, 72)])
def test_gemm_mn_major(cuda, a_mn, b_mn, M, N, K):
    from countr_b200 import ops
    a = _rand16((M, K), cuda, seed=3)
    b = _rand16((N, K), cuda, seed=4)
    a_store = a.t().contiguous() if a_mn else a          # [K, M] when MN-major
    b_store = b.t().contiguous() if b_mn else b
    c = torch.empty(M, N, device=cuda, dtype=torch.float32)
    ops.gemm(a_store, b_store, c, M, N, K, lda=a_store.stride(0), ldb=b_store.stride(0), ldc=N, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    assert_close(c, a.float() @ b.float().t(), 2e-5, f"gemm a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")


def test_gemm_split_k_atomic(cuda):
    from countr_b200 import ops
    M, N, K = 512, 768, 4608
    a = _rand16((K, M), cuda, seed=5)   # MN-major operands: dW = dY^T X
    b = _rand16((K, N), cuda, seed=6)
    c = torch.zeros(M, N, device=cuda, dtype=torch.float32)
    ops.gemm(a, b, c, M, N, K, lda=M, ldb=N, ldc=N, a_mn=True, b_mn=True, atomic=True, split_k=6)
    torch.cuda.synchronize()
    assert_close(c, a.float().t() @ b.float(), 3e-5, "split-k atomic")


def test_gemm_batched_strided(cuda):
    """S = Q K^T per (batch, head) straight out of a packed [B, L, 3, H, dh] buffer."""
    from countr_b200 import ops
    B, L, H, dh = 2, 576, 4, 32
    qkv = _rand16((B, L, 3, H, dh), cuda, seed=7)
    s = torch.empty(B, H, L, L, device=cuda, dtype=torch.float16)
    q = qkv[:, :, 0]
    k = qkv[:, :, 1]
    row = 3 * H * dh
    ops.gemm(q, k, s, L, L, dh, lda=row, ldb=row, ldc=L, nb1=B, nb2=H, sa=(L * row, dh), sb=(L * row, dh),
             sc=(H * L * L, L * L), alpha=0.25)
    torch.cuda.synchronize()
    ref = torch.einsum("blhd,bmhd->bhlm", q.float(), k.float()) * 0.25
    assert_close(s.reshape(-1, L), ref.reshape(-1, L), 1e-3, "batched strided")


def test_gemm_epilogues(cuda):
    from countr_b200 import ops
    M, N, K = 1152, 512, 768
    a = _rand16((M, K), cuda, seed=8, scale=0.5)
    w = _rand16((N, K), cuda, seed=9, scale=0.05)
    bias = torch.randn(N, device=cuda)
    res = torch.randn(M, N, device=cuda)
    lin = a.float() @ w.float().t() + bias
    # bias + residual, fp32 out, in place over the residual
    x = res.clone()
    ops.linear(a, w, x, bias=bias, residual=x)
    torch.cuda.synchronize()
    assert_close(x, lin + res, 2e-5, "bias+residual in place")
    # bias + GELU with pre-activation side output
    pre = torch.empty(M, N, device=cuda, dtype=torch.float16)
    u = torch.empty(M, N, device=cuda, dtype=torch.float16)
    ops.linear(a, w, u, bias=bias, act=1, aux=pre)
    torch.cuda.synchronize()
    assert_close(pre, lin, 1e-3, "pre-activation")
    assert_close(u, F.gelu(lin), 1e-3, "gelu")
    # GELU backward multiply
    g = torch.empty(M, N, device=cuda, dtype=torch.float16)
    ops.linear(a, w, g, act=2, aux=pre)
    torch.cuda.synchronize()
    x_ = pre.float().requires_grad_(True)
    F.gelu(x_).sum().backward()
    assert_close(g, (a.float() @ w.float().t()) * x_.grad, 1.5e-3, "gelu bwd")
    # positional-embedding style broadcast residual (row % res_mod)
    pos = torch.randn(576, N, device=cuda)
    y = torch.empty(M, N, device=cuda, dtype=torch.float32)
    ops.linear(a, w, y, bias=bias, residual=pos, res_mod=576)
    torch.cuda.synchronize()
    assert_close(y, lin + pos.repeat(M // 576, 1), 2e-5, "pos-embed residual")


@pytest.mark.parametrize("M,D,N2,act", [(4608, 768, 2304, 0), (1000, 768, 3072, 1), (4608, 512, 2048, 1), (333, 1024, 1024, 0)])
def test_layernorm_folded_into_gemms(cuda, M, D, N2, act):
    """LayerNorm between two Linear layers without a LayerNorm kernel: the producer GEMM (fp32 out + residual) also emits the
    16-bit copy of its rows and row-statistics partials, the consumer GEMM multiplies by W diag(gamma) and normalises in its
    epilogue.  Against fp32 torch: x = a W1^T + b1 + res; y = act(LayerNorm(x) W2^T + b2)."""
    from countr_b200 import ops
    K1 = 512
    a = _rand16((M, K1), cuda, seed=70, scale=0.5)
    w1 = _rand16((D, K1), cuda, seed=71, scale=0.05)
    b1 = torch.randn(D, device=cuda)
    res = torch.randn(M, D, device=cuda) * 2 + 0.5            # non-zero row means
    gamma, beta = torch.rand(D, device=cuda) + 0.5, torch.randn(D, device=cuda) * 0.1
    w2 = torch.randn(N2, D, device=cuda) * 0.05
    b2 = torch.randn(N2, device=cuda)
    x = res.clone()
    x16 = torch.zeros(M, D, device=cuda, dtype=torch.float16)
    st = torch.zeros(M, 8, 2, device=cuda)
    ops.linear(a, w1, x, bias=b1, residual=x, bn=D // 4, ln_x16=x16, ln_stats=st)
    x_ref = a.float() @ w1.float().t() + b1 + res
    torch.cuda.synchronize()
    assert_close(x, x_ref, 2e-5, "producer output")
    assert_close(x16, x_ref, 1e-3, "16-bit copy")
    assert_close(st.view(M, 8, 2)[:, :, 0].sum(1), x_ref.sum(1), 1e-4, "row sums")
    w2f = (w2 * gamma[None, :]).half().contiguous()
    colsum = w2f.double().sum(1).float().contiguous()
    bias2 = (b2.double() + w2.double() @ beta.double()).float().contiguous()
    y = torch.zeros(M, N2, device=cuda, dtype=torch.float16)
    ops.linear(x16, w2f, y, bias=bias2, act=act, ln_stats=st, ln_colsum=colsum, ln_eps=1e-6)
    torch.cuda.synchronize()
    ref = F.layer_norm(x_ref, (D,), gamma, beta, 1e-6) @ w2.t() + b2
    if act:
        ref = F.gelu(ref)
    assert_close(y, ref, 2e-3, "folded LayerNorm + Linear")


def test_grouped_dw_and_colsum(cuda):
    """Every dW = dY^T X (and bias gradient) of a backward pass in one launch each: mixed shapes, a token count that is not a
    multiple of 64 (zero-filled k tail), a 24-token problem (the k/v projections of the exemplars), output widths that are not
    multiples of the 128 x 256 tile, accumulation on top of earlier contributions."""
    from countr_b200 import ops
    shapes = [(4608, 512, 2048), (4608, 2048, 512), (4608, 1536, 512), (24, 512, 512), (1000, 512, 768), (4608, 512, 512), (333, 192, 320)]
    dws, refs, probs, cprobs, crefs = [], [], [], [], []
    for i, (tok, n_out, k_in) in enumerate(shapes):
        dy = _rand16((tok, n_out), cuda, seed=40 + i, scale=0.5)
        x = _rand16((tok, k_in), cuda, seed=60 + i, scale=0.5)
        dw = torch.randn(n_out, k_in, device=cuda)
        refs.append(dw.double() + dy.double().t() @ x.double())
        probs.append((dy, x, dw))
        db = torch.randn(n_out, device=cuda)
        crefs.append(db.double() + dy.double().sum(0))
        cprobs.append((dy, db))
    x32 = torch.randn(24, 512, device=cuda)
    db32 = torch.zeros(512, device=cuda)
    cprobs.append((x32, db32))
    crefs.append(x32.double().sum(0))
    for shape, dt in (((100, 6), torch.float16), ((33, 10), torch.float32), ((700, 264), torch.float16)):   # widths off the 8-column fast path
        xo = torch.randn(shape, device=cuda).to(dt)
        cprobs.append((xo, torch.zeros(shape[1], device=cuda)))
        crefs.append(xo.double().sum(0))
    ops.grouped_dw(probs)
    ops.grouped_colsum(cprobs)
    torch.cuda.synchronize()
    for (tok, n_out, k_in), (_, _, dw), ref in zip(shapes, probs, refs):
        assert_close(dw, ref.float(), 2e-5, f"grouped dW tokens={tok} {n_out}x{k_in}")
    for (_, db), ref in zip(cprobs, crefs):
        assert_close(db, ref.float(), 2e-5, "grouped colsum")


@pytest.mark.parametrize("M,N,K", [(4608, 2048, 512), (1000, 2048, 512), (4608, 512, 2048), (333, 1536, 512)])
def test_gemm_aux_epilogues_tma(cuda, M, N, K):
    """fc1 forward with the saved pre-activation (second 16-bit output) and fc2's dX with the GELU' multiply (16-bit input),
    both through the TMA epilogue, at the FIM / encoder shapes and with ragged row counts (the TMA unit clips the tails)."""
    from countr_b200 import ops
    a = _rand16((M, K), cuda, seed=18, scale=0.5)
    w = _rand16((N, K), cuda, seed=19, scale=0.05)
    bias = torch.randn(N, device=cuda)
    lin = a.float() @ w.float().t() + bias
    pre = torch.zeros(M, N, device=cuda, dtype=torch.float16)
    u = torch.zeros(M, N, device=cuda, dtype=torch.float16)
    ops.linear(a, w, u, bias=bias, act=1, aux=pre)
    torch.cuda.synchronize()
    assert_close(pre, lin, 1e-3, "pre-activation")
    assert_close(u, F.gelu(lin), 1e-3, "gelu")
    g = torch.zeros(M, N, device=cuda, dtype=torch.float16)
    ops.linear(a, w, g, act=2, aux=pre)
    torch.cuda.synchronize()
    x_ = pre.float().requires_grad_(True)
    F.gelu(x_).sum().backward()
    assert_close(g, (a.float() @ w.float().t()) * x_.grad, 1.5e-3, "gelu bwd")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 24, 24, 512, 256), (2, 48, 48, 256, 256), (1, 96, 96, 256, 256),
                                            (3, 32, 32, 64, 128), (3, 16, 16, 128, 256), (3, 8, 8, 256, 512)])
def test_conv3x3(cuda, B, H, W, Cin, Cout):
    from countr_b200 import ops
    x = _rand16((B, H, W, Cin), cuda, seed=10)
    w = torch.randn(Cout, Cin, 3, 3, device=cuda) * 0.05
    bias = torch.randn(Cout, device=cuda)
    w16 = torch.empty(Cout, 9 * Cin, device=cuda, dtype=torch.float16)
    ops.conv_weight_pack(w, w16, 0)
    y = torch.empty(B, H, W, Cout, device=cuda, dtype=torch.float16)
    stats = torch.zeros(B, Cout // 32, 2, device=cuda, dtype=torch.float64)
    ops.conv3x3(x, w16, y, bias=bias, gn_stats=stats)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), bias, padding=1).permute(0, 2, 3, 1)
    assert_close(y.reshape(-1, Cout), ref.reshape(-1, Cout), 1.5e-3, f"conv {B}x{H}x{W} {Cin}->{Cout}")
    rg = ref.reshape(B, H * W, Cout // 32, 32).double()
    ref_stats = torch.stack([rg.sum((1, 3)), (rg * rg).sum((1, 3))], -1)
    assert_close(stats.reshape(-1, 2), ref_stats.reshape(-1, 2), 1e-4, "gn stats")


def test_conv3x3_dx_weights(cuda):
    """dX of a 3x3/p1 conv == conv of dY with the flipped, channel-transposed filter (pack mode 1)."""
    from countr_b200 import ops
    B, H, W, Cin, Cout = 2, 24, 24, 128, 256
    dy = _rand16((B, H, W, Cout), cuda, seed=11)
    w = torch.randn(Cout, Cin, 3, 3, device=cuda) * 0.05
    wt16 = torch.empty(Cin, 9 * Cout, device=cuda, dtype=torch.float16)
    ops.conv_weight_pack(w, wt16, 1)
    dx = torch.empty(B, H, W, Cin, device=cuda, dtype=torch.float16)
    ops.conv3x3(dy, wt16, dx)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.half().float(), padding=1).permute(0, 2, 3, 1)
    assert_close(dx.reshape(-1, Cin), ref.reshape(-1, Cin), 1.5e-3, "conv dX")


# ------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,D", [(4608, 768), (1152, 512), (37, 1024)])
def test_layernorm(cuda, rows, D):
    from countr_b200 import ops
    x = torch.randn(rows, D, device=cuda) * 3 + 0.5
    g = torch.randn(D, device=cuda)
    b = torch.randn(D, device=cuda)
    y16 = torch.empty(rows, D, device=cuda, dtype=torch.float16)
    y32 = torch.empty(rows, D, device=cuda)
    mean = torch.empty(rows, device=cuda)
    rstd = torch.empty(rows, device=cuda)
    ops.layernorm_fwd(x, g, b, 1e-6, y16=y16, y32=y32, mean=mean, rstd=rstd)
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (D,), g, b, 1e-6)
    assert_close(y32, ref, 1e-5, "ln f32")
    assert_close(y16, ref, 1e-3, "ln f16")
    assert_close(mean[:, None], x.mean(1, keepdim=True), 1e-5, "ln mean")
    # backward
    dy = torch.randn(rows, D, device=cuda)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gr, br, 1e-6).backward(dy)
    dx = torch.ones(rows, D, device=cuda)
    dg = torch.zeros(D, device=cuda)
    db = torch.zeros(D, device=cuda)
    cs = torch.zeros(D, device=cuda)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, accumulate=True, dx_colsum=cs)
    torch.cuda.synchronize()
    assert_close(dx, xr.grad + 1.0, 2e-5, "ln dx (accumulated)")
    assert_close(cs[None], dx.sum(0)[None], 1e-4, "ln dx column sums")
    assert_close(dg[None], gr.grad[None], 1e-4, "ln dgamma")
    assert_close(db[None], br.grad[None], 1e-4, "ln dbeta")


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,L,H,dh", [(1, 128, 1, 64), (2, 576, 12, 64), (2, 576, 16, 32), (1, 288, 12, 64), (1, 200, 2, 32),
                                      (1, 729, 16, 80), (2, 100, 3, 48)])     # the last two: generic head sizes (ViT-H: 80)
def test_attention_fwd(cuda, B, L, H, dh):
    from countr_b200 import ops
    qkv = _rand16((B, L, 3, H, dh), cuda, seed=12, scale=1.5)
    out = torch.empty(B, L, H * dh, device=cuda, dtype=torch.float16)
    lse = torch.empty(B, H, L, device=cuda)
    scale = dh ** -0.5
    ops.attention_fwd(qkv, out, B, L, H, dh, scale, lse=lse)
    torch.cuda.synchronize()
    q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * scale
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, L, H * dh)
    assert_close(out.reshape(-1, H * dh), ref.reshape(-1, H * dh), 2e-3, f"attention {B},{L},{H},{dh}")
    assert_close(lse.reshape(-1, L), torch.logsumexp(s, -1).reshape(-1, L), 1e-4, "lse")


def test_attention_fwd_growing_scores(cuda):
    """Scores that keep growing along the key axis: the running maximum moves in every 128-key chunk, by more and by less
    than the lazy-rescale threshold (2^8), for some rows and not for others."""
    from countr_b200 import ops
    B, L, H, dh = 1, 576, 2, 64
    qkv = _rand16((B, L, 3, H, dh), cuda, seed=21, scale=1.0).float()
    ramp = torch.linspace(0.2, 6.0, L, device=cuda)[None, :, None, None]
    qkv[:, :, 1] *= ramp                                   # |k| grows 30x from the first to the last key
    qkv[:, ::3, 0] *= 0.05                                 # every third query barely moves its maximum
    qkv = qkv.half()
    out = torch.empty(B, L, H * dh, device=cuda, dtype=torch.float16)
    lse = torch.empty(B, H, L, device=cuda)
    scale = dh ** -0.5
    ops.attention_fwd(qkv, out, B, L, H, dh, scale, lse=lse)
    torch.cuda.synchronize()
    q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * scale
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, L, H * dh)
    assert_close(out.reshape(-1, H * dh), ref.reshape(-1, H * dh), 2e-3, "attention, growing scores")
    assert_close(lse.reshape(-1, L), torch.logsumexp(s, -1).reshape(-1, L), 1e-4, "lse, growing scores")


def test_cross_attn_core(cuda):
    from countr_b200 import ops
    B, L, D, dh = 2, 576, 512, 32
    for S in (1, 2, 3, 5, 9, 16):      # evaluation passes every annotated box (up to 16); training uses <= 3
        q = _rand16((B * L, D), cuda, seed=13)
        k = torch.randn(B, S, D, device=cuda)
        v = torch.randn(B, S, D, device=cuda)
        out = torch.empty(B * L, D, device=cuda, dtype=torch.float16)
        probs = torch.empty(B * L, D // dh, S, device=cuda)
        ops.cross_attn_core(q, k, v, out, B, L, S, D, dh, dh ** -0.5, probs=probs)
        torch.cuda.synchronize()
        qh = q.float().reshape(B, L, D // dh, dh).permute(0, 2, 1, 3)
        kh = k.reshape(B, S, D // dh, dh).permute(0, 2, 1, 3)
        vh = v.reshape(B, S, D // dh, dh).permute(0, 2, 1, 3)
        p = ((qh @ kh.transpose(-1, -2)) * dh ** -0.5).softmax(-1)
        ref = (p @ vh).permute(0, 2, 1, 3).reshape(B * L, D)
        assert_close(out, ref, 1e-3, f"cross attn S={S}")
        assert_close(probs.reshape(-1, S), p.permute(0, 2, 1, 3).reshape(-1, S), 1e-5, f"cross attn probs S={S}")


# ------------------------------------------------------------------------------------------
# layout / norm / resample glue
# ------------------------------------------------------------------------------------------
def test_casts_and_patchify(cuda):
    from countr_b200 import ops
    w = torch.randn(300, 520, device=cuda)
    o = torch.empty(300, 520, device=cuda, dtype=torch.float16)
    ops.cast16(w, o)
    ot = torch.empty(520, 300, device=cuda, dtype=torch.float16)
    ops.cast16_transpose(w, ot)
    torch.cuda.synchronize()
    assert torch.equal(o, w.half()) and torch.equal(ot, w.t().half())
    img = torch.rand(2, 3, 384, 512, device=cuda)[:, :, :, 64:448]          # non-contiguous view
    for t in (img, img.half()):
        a = torch.empty(2 * 576, 768, device=cuda, dtype=torch.float16)
        ops.patchify(t, a, 16)
        torch.cuda.synchronize()
        ref = F.unfold(t.float(), 16, stride=16).transpose(1, 2).reshape(2 * 576, 768).half()
        assert torch.equal(a, ref)


@pytest.mark.parametrize("B,H,W", [(2, 24, 24), (1, 96, 96), (1, 40, 37), (3, 5, 9), (8, 48, 48), (2, 100, 100)])
def test_gn_relu_upsample(cuda, B, H, W):
    """Whole 8 x 8 tiles, ragged tiles on both axes, a map below the staged kernel's minimum size (row-walking kernel), several tiles per
    persistent block with the three-stage ring wrapping."""
    from countr_b200 import ops
    C, G = 256, 8
    x = _rand16((B, H, W, C), cuda, seed=14, scale=2.0)
    gamma = torch.randn(C, device=cuda)
    beta = torch.randn(C, device=cuda)
    xg = x.double().reshape(B, H * W, G, C // G)
    stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()
    y = torch.empty(B, 2 * H, 2 * W, C, device=cuda, dtype=torch.float16)
    ops.gn_relu_upsample2x(x, stats, gamma, beta, y, G, 1e-5)
    xn = x.float().permute(0, 3, 1, 2)
    act = F.relu(F.group_norm(xn, G, gamma, beta, 1e-5))
    ref = F.interpolate(act, scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert_close(y.reshape(-1, C), ref.reshape(-1, C), 1e-3, "gn+relu+up2")
    w = torch.randn(C, device=cuda) * 0.1
    bias = torch.randn(1, device=cuda)
    d = torch.empty(B, H, W, device=cuda)
    ops.gn_relu_conv1x1(x, stats, gamma, beta, w, bias, d, G, 1e-5)
    up = torch.empty(B, 2 * H, 2 * W, device=cuda)
    ops.upsample2x_f32(d, up)
    torch.cuda.synchronize()
    refd = F.conv2d(act, w.reshape(1, C, 1, 1), bias)
    assert_close(d.reshape(B, -1), refd.reshape(B, -1), 1e-5, "gn+relu+1x1")
    refu = F.interpolate(refd, scale_factor=2, mode="bilinear", align_corners=False).squeeze(1)
    assert_close(up.reshape(B, -1), refu.reshape(B, -1), 1e-5, "up2 f32")


def test_exemplar_stage1_and_inorm(cuda):
    from countr_b200 import ops
    B, K, S = 2, 3, 2
    torch.manual_seed(341)     # fp16 outputs against a 1e-3 max-scaled bound: keep the draw fixed
    boxes = torch.rand(B, K, 3, 64, 64, device=cuda)
    w = torch.randn(64, 3, 3, 3, device=cuda) * 0.2
    bias = torch.randn(64, device=cuda) * 0.1
    raw = torch.empty(B * S, 64, 64, 64, device=cuda, dtype=torch.float16)
    ops.exemplar_conv1(boxes, S, w, bias, raw)
    torch.cuda.synchronize()
    xin = boxes[:, :S].reshape(B * S, 3, 64, 64)
    ref = F.conv2d(xin, w, bias, padding=1)
    assert_close(raw.permute(0, 3, 1, 2).reshape(B * S * 64, -1), ref.reshape(B * S * 64, -1), 1e-3, "exemplar conv1")
    pooled = torch.empty(B * S, 32, 32, 64, device=cuda, dtype=torch.float16)
    mean = torch.empty(B * S, 64, device=cuda)
    rstd = torch.empty(B * S, 64, device=cuda)
    scratch = torch.empty(64 * B * S * 64, device=cuda)
    ops.inorm_relu_pool(raw, 0, 1e-5, y16=pooled, mean=mean, rstd=rstd, scratch=scratch)      # pixel-parallel path
    pooled1 = torch.empty_like(pooled)
    ops.inorm_relu_pool(raw, 0, 1e-5, y16=pooled1)                                            # one CTA per (sample, 64 ch)
    torch.cuda.synchronize()
    assert rel_close(pooled1, pooled)
    r = raw.float().permute(0, 3, 1, 2)
    refp = F.max_pool2d(F.relu(F.instance_norm(r, eps=1e-5)), 2)
    assert_close(pooled.permute(0, 3, 1, 2).reshape(B * S * 64, -1), refp.reshape(B * S * 64, -1), 1e-3, "IN+relu+maxpool")
    y32 = torch.empty(B * S, 64, device=cuda)
    y16 = torch.empty(B * S, 64, device=cuda, dtype=torch.float16)
    ops.inorm_relu_pool(raw, 1, 1e-5, y16=y16, y32=y32)
    torch.cuda.synchronize()
    refa = F.relu(F.instance_norm(r, eps=1e-5)).mean((2, 3))
    assert_close(y32, refa, 1e-5, "IN+relu+avgpool")


# ------------------------------------------------------------------------------------------
# cta_group::2 (two SMs per 256 x N tile)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, 256), (256, 128, 128, 128), (512, 512, 512, 256), (4608, 768, 768, 0),
                                      (4608, 2304, 768, 256), (1000, 520, 200, 0), (8192, 1024, 1024, 256)])
def test_gemm_cta_pair(cuda, M, N, K, bn):
    from countr_b200 import ops
    a = _rand16((M, K), cuda, seed=71)
    b = _rand16((N, K), cuda, seed=72)
    bias = torch.randn(N, device=cuda)
    c = torch.empty(M, N, device=cuda, dtype=torch.float32)
    ops.gemm(a, b, c, M, N, K, lda=K, ldb=K, ldc=N, bn=bn, bias=bias, pair=1)
    torch.cuda.synchronize()
    assert_close(c, a.float() @ b.float().t() + bias, 2e-5, f"pair gemm {M}x{N}x{K} bn={bn}")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 24, 24, 512, 256), (2, 48, 48, 256, 256), (1, 96, 96, 256, 256), (3, 16, 16, 128, 256)])
def test_conv3x3_cta_pair(cuda, B, H, W, Cin, Cout):
    from countr_b200 import ops
    x = _rand16((B, H, W, Cin), cuda, seed=73)
    w = torch.randn(Cout, Cin, 3, 3, device=cuda) * 0.05
    bias = torch.randn(Cout, device=cuda)
    w16 = torch.empty(Cout, 9 * Cin, device=cuda, dtype=torch.float16)
    ops.conv_weight_pack(w, w16, 0)
    y = torch.empty(B, H, W, Cout, device=cuda, dtype=torch.float16)
    stats = torch.zeros(B, Cout // 32, 2, device=cuda, dtype=torch.float64)
    ops.conv3x3(x, w16, y, bias=bias, gn_stats=stats, pair=1)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), bias, padding=1).permute(0, 2, 3, 1)
    assert_close(y.reshape(-1, Cout), ref.reshape(-1, Cout), 1.5e-3, f"pair conv {B}x{H}x{W} {Cin}->{Cout}")
    rg = ref.reshape(B, H * W, Cout // 32, 32).double()
    ref_stats = torch.stack([rg.sum((1, 3)), (rg * rg).sum((1, 3))], -1)
    assert_close(stats.reshape(-1, 2), ref_stats.reshape(-1, 2), 1e-4, "gn stats (pair)")


@pytest.mark.parametrize("pair", [1, -1])
def test_conv_dw_and_linear_dw_cta_pair(cuda, pair):
    from countr_b200 import ops
    for (B, H, W, Cin, Cout) in [(2, 24, 24, 512, 256), (2, 48, 48, 256, 256), (3, 16, 16, 128, 256)]:
        x = _rand16((B, H, W, Cin), cuda, seed=81)
        dy = _rand16((B, H, W, Cout), cuda, seed=82)
        w = torch.zeros(Cout, Cin, 3, 3, device=cuda, requires_grad=True)
        F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        dwp = torch.zeros(Cout, 9 * Cin, device=cuda)
        ops.conv3x3_dw(dy, x, dwp, pair=pair)
        dw = torch.empty(Cout, Cin, 3, 3, device=cuda)
        ops.conv_dw_unpack(dwp, dw, Cout, Cin)
        torch.cuda.synchronize()
        assert_close(dw.reshape(Cout, -1), w.grad.reshape(Cout, -1), 1e-4, f"conv dW pair={pair} {B}x{H}x{W} {Cin}->{Cout}")
    M, N, K = 512, 2048, 4608          # Linear dW = dY^T X, both operands MN-major
    a = _rand16((K, M), cuda, seed=83)
    b = _rand16((K, N), cuda, seed=84)
    c = torch.zeros(M, N, device=cuda, dtype=torch.float32)
    ops.gemm(a, b, c, M, N, K, lda=M, ldb=N, ldc=N, a_mn=True, b_mn=True, atomic=True, split_k=4, bn=256, pair=pair)
    torch.cuda.synchronize()
    assert_close(c, a.float().t() @ b.float(), 3e-5, f"linear dW pair={pair}")


def test_weight_refresh_all_kinds_ragged_shapes(cuda):
    """countr_weight_refresh: cast, transpose and both conv filter packings in one launch, on shapes that do not fill the 64 x 64
    tiles / 128-channel slabs (odd row counts take the unpacked store path)."""
    from countr_b200 import ops
    from countr_b200.engine import WeightCache
    g = torch.Generator().manual_seed(5)
    mk = lambda *shape: torch.nn.Parameter(torch.randn(*shape, generator=g).to(cuda))  # noqa: E731
    lin = [mk(512, 2048), mk(100, 70), mk(33, 65), mk(1, 7)]
    conv = [mk(256, 512, 3, 3), mk(64, 3, 3, 3), mk(37, 21, 3, 3)]
    wc = WeightCache()
    plan = [(p, "w") for p in lin] + [(p, "wt") for p in lin] + [(p, "c0") for p in conv] + [(p, "c1") for p in conv]
    before = ops.LAUNCHES[0]
    assert wc.refresh_batch(plan) == len(plan)
    assert ops.LAUNCHES[0] - before <= 1
    for p in lin:
        assert torch.equal(wc.w16(p), p.detach().half())
        assert torch.equal(wc.w16_t(p), p.detach().t().contiguous().half())
    for p in conv:
        co, ci = p.shape[:2]
        w = p.detach()
        assert torch.equal(wc.conv16(p, 0).reshape(co, 9, ci), w.reshape(co, ci, 9).permute(0, 2, 1).contiguous().half())
        flipped = w.reshape(co, ci, 9).flip(-1)                               # tap -> 8 - tap
        assert torch.equal(wc.conv16(p, 1).reshape(ci, 9, co), flipped.permute(1, 2, 0).contiguous().half())
