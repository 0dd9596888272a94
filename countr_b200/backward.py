"""Decoder backward schedule of the fine-tune step (host side).

Reverse of `Engine.decoder_forward`: density head -> decoder_norm -> FIM blocks -> decoder_embed,
plus the exemplar CNN.  Every gradient is produced by the sm_100a kernels of libcountr_sm100.so:
  dX of a Linear      tcgen05 GEMM against the transposed 16-bit weight copy
  dW of a Linear      tcgen05 GEMM with both operands MN-major (dY^T X), split-K, fp32 atomics
  dX / dW of Conv3x3  implicit GEMM with the flipped filter / pixels-as-K implicit GEMM
  attention backward  batched GEMMs over materialised [B,H,L,L] scores + a row softmax-backward
  everything else     the streaming kernels of csrc/backward.cu and csrc/norm.cu

The reference obtains all of this from autograd (scaler.scale(loss).backward(),
util/misc.py:266-270); which parameters receive a gradient follows models_mae_cross.py:150-207
(encoder frozen under no_grad; shot_token only when shot_num == 0; decoder_proj* only when > 0).
Parameter gradients are views into one flat fp32 arena (one memset, ready for a single
all-reduce); they are returned to autograd, which accumulates them into `.grad` as usual.
"""
import os

import torch

from . import ops
from .dist import ARENA_TAIL, build_grad_arena, param_flag_index, usage_flags
from .engine import F16, F32, _contig32


def _dw_linear(dy16, x16, dw32):
    """dw32[N_out, K_in] += dy16[M, N_out]^T @ x16[M, K_in]   (both operands MN-major, split-K)."""
    Mtok, n_out = dy16.shape
    k_in = x16.shape[1]
    out_tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    kblocks = (Mtok + 63) // 64
    split = max(1, min(kblocks, 148 // max(1, out_tiles)))
    ops.gemm(dy16, x16, dw32, n_out, k_in, Mtok, lda=dy16.stride(0), ldb=x16.stride(0), ldc=k_in, a_mn=True, b_mn=True,
             atomic=True, split_k=split)


class GradJobs:
    """Deferred weight / bias gradients of a backward pass.  They are leaves of the backward graph, so the schedules only
    RECORD them (operands must stay alive and unmodified) and `flush()` computes all of them with one grouped tcgen05 launch
    per 32 weight gradients and one grouped column-sum launch per 32 bias gradients (csrc/grouped.cu) — instead of a
    launch + prologue + pipeline fill + exposed epilogue per layer in the middle of the latency-bound dX chain."""

    def __init__(self):
        self.dw, self.cs = [], []

    def dW(self, dy16, x16, dw32):
        self.dw.append((dy16, x16, dw32.view(dw32.shape[0], -1)))

    def dB(self, x, out32):
        self.cs.append((x, out32.view(-1)))

    def flush(self):
        if self.cs:
            ops.grouped_colsum(self.cs)
        if self.dw:
            ops.grouped_dw(self.dw)
        self.dw, self.cs = [], []


def attention_backward(qkv, lse, datt, B, L, H, dh, scale, att=None, fused=True):
    """qkv fp16 [B*L, 3*H*dh] (packed), lse fp32 [B,H,L], datt fp16 [B*L, H*dh] -> dqkv fp16 [B*L, 3*H*dh].
    head_dim 32 (the FIM, the pre-training decoder) and 64 (the pre-training encoder) run the fused flash-style kernels (they
    need the forward output `att`); other head sizes use batched tcgen05 GEMMs over materialised [B,H,L,L] scores plus a row
    softmax-backward."""
    dev = qkv.device
    D = H * dh
    row = 3 * D
    if fused and att is not None and ((dh == 32 and L <= 640) or dh == 64):
        dqkv = torch.empty(B * L, row, dtype=F16, device=dev)
        ops.attention_bwd(qkv, att, datt, lse, dqkv, B, L, H, dh, scale)
        return dqkv
    flat = qkv.view(-1)
    q, k, v = flat, flat[D:], flat[2 * D:]
    s16 = torch.empty(B, H, L, L, dtype=F16, device=dev)
    dp16 = torch.empty(B, H, L, L, dtype=F16, device=dev)
    bs = dict(nb1=B, nb2=H)
    # S = scale * Q K^T ; dP = dO V^T
    ops.gemm(q, k, s16, L, L, dh, lda=row, ldb=row, ldc=L, sa=(L * row, dh), sb=(L * row, dh), sc=(H * L * L, L * L),
             alpha=scale, **bs)
    ops.gemm(datt, v, dp16, L, L, dh, lda=D, ldb=row, ldc=L, sa=(L * D, dh), sb=(L * row, dh), sc=(H * L * L, L * L), **bs)
    ops.softmax_bwd_rows(s16, dp16, lse, scale)          # s16 <- P, dp16 <- dS (already * scale)
    dqkv = torch.empty(B * L, row, dtype=F16, device=dev)
    dflat = dqkv.view(-1)
    # dV = P^T dO
    ops.gemm(s16, datt, dflat[2 * D:], L, dh, L, lda=L, ldb=D, ldc=row, a_mn=True, b_mn=True, sa=(H * L * L, L * L),
             sb=(L * D, dh), sc=(L * row, dh), **bs)
    # dQ = dS K
    ops.gemm(dp16, k, dflat, L, dh, L, lda=L, ldb=row, ldc=row, b_mn=True, sa=(H * L * L, L * L), sb=(L * row, dh),
             sc=(L * row, dh), **bs)
    # dK = dS^T Q
    ops.gemm(dp16, q, dflat[D:], L, dh, L, lda=L, ldb=row, ldc=row, a_mn=True, b_mn=True, sa=(H * L * L, L * L),
             sb=(L * row, dh), sc=(L * row, dh), **bs)
    return dqkv


def _conv_grads(d_raw, inp, conv, grads, wc, dx_dtype):
    """dW (+ unpack into Conv2d layout) and dX of a 3x3/p1 conv; d_raw, inp are NHWC fp16."""
    cout, cin = conv.weight.shape[0], conv.weight.shape[1]
    name_w = grads["__names__"][id(conv.weight)]
    dwp = torch.empty(cout, 9 * cin, dtype=F32, device=d_raw.device)
    ops.zero_(dwp)
    ops.conv3x3_dw(d_raw, inp, dwp)
    ops.conv_dw_unpack(dwp, grads[name_w], cout, cin)
    if dx_dtype is None:
        return None
    d_in = torch.empty(inp.shape, dtype=dx_dtype, device=d_raw.device)
    ops.conv3x3(d_raw, wc.conv16(conv.weight, mode=1), d_in)
    return d_in


def _exemplar_backward(m, sv, boxes, S, dy32, grads, wc, G):
    """decoder_proj1..4 backward (models_mae_cross.py:157-177 reversed) given dL/dy [B*S, C]."""
    dev = dy32.device
    ex = sv["exemplar"]
    convs = [m.decoder_proj1[0], m.decoder_proj2[0], m.decoder_proj3[0], m.decoder_proj4[0]]
    raw4 = ex["raw"][3]
    d_raw = torch.empty_like(raw4)
    ops.inorm_relu_pool_bwd(raw4, ex["mean"][3], ex["rstd"][3], d_raw, 1, dpool32=dy32, dbias=G(convs[3].bias))
    for i in (3, 2, 1):
        d_pool = _conv_grads(d_raw, ex["pooled"][i - 1], convs[i], grads, wc, F16)
        raw = ex["raw"][i - 1]
        d_raw = torch.empty_like(raw)
        scratch = torch.empty(raw.shape[0], raw.shape[3], 2, dtype=F32, device=dev)
        ops.inorm_relu_pool_bwd(raw, ex["mean"][i - 1], ex["rstd"][i - 1], d_raw, 0, dpool16=d_pool, dbias=G(convs[i - 1].bias),
                                scratch=scratch)
    ops.exemplar_conv1_dw(boxes, S, d_raw, G(convs[0].weight))


def decoder_backward(eng, m, sv, boxes, grad_out):
    dev = grad_out.device
    wc = eng.wc
    B, L, shot_num, S = sv["B"], sv["L"], sv["shot_num"], sv["S"]
    M = B * L
    Dd = m.decoder_embed.weight.shape[0]

    # ---- flat gradient arena: every decoder parameter (also the ones this shot_num does not reach, left at zero) + usage
    # flags, so that all data-parallel ranks reduce the same layout whatever shot_num each of them drew (dist.py)
    names, params = m._decoder_params(None)
    arena, views = build_grad_arena(names, params, dev, tail=ARENA_TAIL)
    ops.zero_(arena)
    arena[arena.numel() - ARENA_TAIL:].copy_(eng.usage_flags(shot_num, dev), non_blocking=True)
    grads = {"__names__": {id(p): n for n, p in zip(names, params)}}
    grads.update(views)

    def G(p):
        return grads[grads["__names__"][id(p)]]

    # ---- density head (models_mae_cross.py:189-197 reversed)
    heads = [m.decode_head0, m.decode_head1, m.decode_head2, m.decode_head3]
    hs = sv["heads"]
    go = grad_out if grad_out.is_contiguous() else grad_out.contiguous()
    raw3 = hs[3]["raw"]
    dmap = torch.empty(B, raw3.shape[1], raw3.shape[2], dtype=F32, device=dev)
    ops.upsample2x_bwd(go, dmap)
    d_next = None
    for i in (3, 2, 1, 0):
        conv, gn = heads[i][0], heads[i][1]
        raw, stats, inp = hs[i]["raw"], hs[i]["stats"], hs[i]["inp"]
        dyh = torch.empty_like(raw)
        gsum = torch.empty(B, gn.num_groups, 2, dtype=torch.float64, device=dev)
        ops.zero_(gsum)
        gamma, beta = _contig32(gn.weight), _contig32(gn.bias)
        if i == 3:
            c1 = heads[3][3]
            ops.gn_relu_bwd_reduce(raw, stats, gamma, beta, dyh, G(gn.weight), G(gn.bias), gsum, gn.num_groups, gn.eps,
                                   dmap=dmap, w1=_contig32(c1.weight).reshape(-1), dw1=G(c1.weight).view(-1), db1=G(c1.bias))
        else:
            ops.gn_relu_bwd_reduce(raw, stats, gamma, beta, dyh, G(gn.weight), G(gn.bias), gsum, gn.num_groups, gn.eps,
                                   d_next=d_next)
        ops.gn_bwd_apply(raw, dyh, stats, gsum, gamma, dyh, G(conv.bias), gn.num_groups, gn.eps)   # in place: dyh -> d_raw
        d_next = _conv_grads(dyh, inp, conv, grads, wc, F32 if i == 0 else F16)

    # weight / bias / LayerNorm-parameter gradients are deferred to grouped launches at the end (see GradJobs)
    defer = getattr(eng, "defer_dw", True)
    jobs = GradJobs() if defer else None
    dw_jobs, cs_jobs = (jobs.dw, jobs.cs) if defer else ([], [])

    # ---- decoder_norm
    g = torch.empty(M, Dd, dtype=F32, device=dev)       # gradient of the fp32 residual stream
    g16 = torch.empty(M, Dd, dtype=F16, device=dev)
    dn = m.decoder_norm
    blocks_rev = list(reversed(list(m.decoder_blocks)))
    # every LayerNorm backward also emits the column sums of the residual gradient it produces = the bias gradient of the
    # Linear that consumes it next (fc2 of the block below, the two attention projections, decoder_embed)
    ops.layernorm_bwd(d_next.view(M, Dd), sv["x_final"], _contig32(dn.weight), sv["meanf"], sv["rstdf"], g, G(dn.weight),
                      G(dn.bias), accumulate=False, dx16=g16, dx_colsum=G(blocks_rev[0].mlp.fc2.bias), jobs=jobs)

    # ---- FIM blocks (models_crossvit.py:152-156 reversed)
    y16 = sv["y16"]
    ny = y16.shape[0]
    kvb = sv["kv_broadcast"]
    if shot_num == 0:
        dy32 = G(m.shot_token).view(1, Dd)              # accumulate straight into shot_token.grad
    else:
        dy32 = torch.empty(ny, Dd, dtype=F32, device=dev)
        ops.zero_(dy32)
    dh = torch.empty(M, Dd, dtype=F32, device=dev)
    side = None
    kv_done = None
    n_blocks = len(sv["blocks"])
    # Weight / bias gradients are leaves of the backward graph: the chain LayerNorm' -> dX GEMM -> attention' -> ... never
    # reads them.  They run on the side stream (one fork per producer, one join at the end), so the latency-bound
    # dX chain on the main stream is not serialised behind ~45 dW GEMMs and column sums.
    main = torch.cuda.current_stream()
    wstream = eng.side_stream(dev) if eng.overlap_dw else None
    keep = []           # tensors the side stream reads: kept alive until the join (the allocator recycles by stream order)

    def off(fn, *tensors):
        if wstream is None:
            fn()
            return
        keep.extend(tensors)
        wstream.wait_stream(main)
        with torch.cuda.stream(wstream):
            fn()

    def new_g16():
        # the 16-bit residual gradient is read by the deferred dW GEMMs: a fresh buffer per LayerNorm backward
        return torch.empty(M, Dd, dtype=F16, device=dev)

    # Weight and bias gradients of the Linear layers are leaves of the backward graph: they are only RECORDED here and
    # computed at the end by one grouped tcgen05 launch (+ one grouped column-sum launch) over all of them — 17 GEMMs and
    # 8 column sums that would otherwise sit between the dX GEMMs of the latency-bound chain, each with its own launch,
    # prologue, pipeline fill and exposed epilogue (csrc/grouped.cu).  eng.defer_dw = False restores the per-layer launches.

    def dW(dy16, x16, p):
        if defer:
            dw_jobs.append((dy16, x16, G(p).view(p.shape[0], -1)))
        else:
            off(lambda: _dw_linear(dy16, x16, G(p)), dy16)

    def dB(dy, p):
        if defer:
            cs_jobs.append((dy, G(p)))
        else:
            off(lambda: ops.colsum(dy, G(p)), dy)
    for bi, (blk, s) in enumerate(zip(blocks_rev, reversed(sv["blocks"]))):
        H = blk.selfattn.num_heads
        dhd = Dd // H
        hid = blk.mlp.fc1.weight.shape[0]
        # --- MLP: x3 = x2 + fc2(gelu(fc1(LN2 x2)))
        dW(g16, s["u"], blk.mlp.fc2.weight)
        dpre = torch.empty(M, hid, dtype=F16, device=dev)
        ops.linear(g16, wc.w16_t(blk.mlp.fc2.weight), dpre, act=2, aux=s["pre"])
        dB(dpre, blk.mlp.fc1.bias)
        dW(dpre, s["h2"], blk.mlp.fc1.weight)
        ops.linear(dpre, wc.w16_t(blk.mlp.fc1.weight), dh)
        g16 = new_g16()
        ops.layernorm_bwd(dh, s["x2"], _contig32(blk.norm2.weight), s["mean2"], s["rstd2"], g, G(blk.norm2.weight),
                          G(blk.norm2.bias), accumulate=True, dx16=g16, dx_colsum=G(blk.attn.proj.bias), jobs=jobs)
        # --- cross attention: x2 = x1 + proj(core(wq(LN1 x1), wk(y), wv(y)))
        ca = blk.attn
        dW(g16, s["c16"], ca.proj.weight)
        dc = torch.empty(M, Dd, dtype=F16, device=dev)
        ops.linear(g16, wc.w16_t(ca.proj.weight), dc)
        dq = torch.empty(M, Dd, dtype=F16, device=dev)
        dk32 = torch.empty(ny, Dd, dtype=F32, device=dev)
        dv32 = torch.empty(ny, Dd, dtype=F32, device=dev)
        ops.zero_(dk32)
        ops.zero_(dv32)
        ops.cross_attn_core_bwd(s["q16"], s["k32"], s["v32"], s["probs"], dc, dq, dk32, dv32, B, L, S, Dd, dhd, ca.scale,
                                kv_broadcast=kvb)
        dB(dq, ca.wq.bias)
        dW(dq, s["h1"], ca.wq.weight)
        ops.linear(dq, wc.w16_t(ca.wq.weight), dh)
        def _kv_grads(dk32=dk32, dv32=dv32):
            # everything downstream of dK / dV only feeds the exemplar branch (dL/dy), never the token stream
            dk16 = torch.empty(ny, Dd, dtype=F16, device=dev)
            dv16 = torch.empty(ny, Dd, dtype=F16, device=dev)
            keep.extend((dk16, dv16))
            ops.cast16(dk32, dk16)
            ops.cast16(dv32, dv16)
            if defer:
                cs_jobs.extend(((dk32, G(ca.wk.bias)), (dv32, G(ca.wv.bias))))
                dw_jobs.extend(((dk16, y16, G(ca.wk.weight)), (dv16, y16, G(ca.wv.weight))))
            else:
                ops.colsum(dk32, G(ca.wk.bias))
                ops.colsum(dv32, G(ca.wv.bias))
                _dw_linear(dk16, y16, G(ca.wk.weight))
                _dw_linear(dv16, y16, G(ca.wv.weight))
            ops.linear(dk16, wc.w16_t(ca.wk.weight), dy32, residual=dy32)
            ops.linear(dv16, wc.w16_t(ca.wv.weight), dy32, residual=dy32)
        off(_kv_grads, dk32, dv32)
        if bi == n_blocks - 1 and shot_num > 0:
            # dL/dy is complete: the exemplar-CNN backward (~45 tiny launches) runs on the side stream while this
            # stream finishes block 0's self-attention backward and decoder_embed
            if eng.overlap_exemplar:
                side = eng.side_stream(dev)
                side.wait_stream(torch.cuda.current_stream())
                if wstream is not None and defer:
                    # the deferred weight-gradient jobs only need the k/v gradients that are already queued on the side
                    # stream, not the exemplar-CNN backward behind them: the grouped launch at the end waits for this event
                    kv_done = torch.cuda.Event()
                    kv_done.record(side)
                with torch.cuda.stream(side):
                    _exemplar_backward(m, sv, boxes, S, dy32, grads, wc, G)
            else:
                if wstream is not None:
                    main.wait_stream(wstream)      # dL/dy is accumulated on the side stream
                _exemplar_backward(m, sv, boxes, S, dy32, grads, wc, G)
        g16 = new_g16()
        ops.layernorm_bwd(dh, s["x1"], _contig32(blk.norm1.weight), s["mean1"], s["rstd1"], g, G(blk.norm1.weight),
                          G(blk.norm1.bias), accumulate=True, dx16=g16, dx_colsum=G(blk.selfattn.proj.bias), jobs=jobs)
        # --- self attention: x1 = x0 + proj(attn(qkv(LN0 x0)))
        sa = blk.selfattn
        dW(g16, s["att"], sa.proj.weight)
        datt = torch.empty(M, Dd, dtype=F16, device=dev)
        ops.linear(g16, wc.w16_t(sa.proj.weight), datt)
        dqkv = attention_backward(s["qkv"], s["lse"], datt, B, L, H, dhd, sa.scale, att=s["att"])

        dB(dqkv, sa.qkv.bias)
        dW(dqkv, s["h0"], sa.qkv.weight)
        ops.linear(dqkv, wc.w16_t(sa.qkv.weight), dh)
        nxt_bias = blocks_rev[bi + 1].mlp.fc2.bias if bi + 1 < n_blocks else m.decoder_embed.bias
        g16 = new_g16()
        ops.layernorm_bwd(dh, s["x0"], _contig32(blk.norm0.weight), s["mean0"], s["rstd0"], g, G(blk.norm0.weight),
                          G(blk.norm0.bias), accumulate=True, dx16=g16, dx_colsum=G(nxt_bias), jobs=jobs)

    # ---- decoder_embed: weight / bias only (its input is the frozen encoder's output)
    de = m.decoder_embed
    if defer:
        dw_jobs.append((g16, sv["lat16"], G(de.weight)))
    else:
        _dw_linear(g16, sv["lat16"], G(de.weight))

    if kv_done is not None and getattr(eng, "early_flush", os.environ.get("COUNTR_EARLY_FLUSH", "1") != "0"):
        torch.cuda.current_stream().wait_event(kv_done)
        jobs.flush()
        torch.cuda.current_stream().wait_stream(eng.side_stream(dev))     # join the exemplar-CNN backward
    else:
        if side is not None or wstream is not None:
            torch.cuda.current_stream().wait_stream(eng.side_stream(dev))     # join the exemplar-CNN backward and the k/v work
        if defer:
            jobs.flush()
    keep.clear()
    # Every parameter that just received a gradient is about to be changed by an optimizer, and torch's version counter
    # cannot be relied on to say so (torch.optim.AdamW(fused=True) updates parameters without bumping `_version`, and so
    # does anything that writes through `.data`): mark their 16-bit copies stale now, the next forward re-casts them.
    wc.bump(params)
    eng.last_arena = arena             # trainers that all-reduce outside the autograd node pick the arena up here
    if eng.grad_allreduce is not None:
        eng.grad_allreduce(arena)      # data-parallel mean of every decoder gradient in one collective
        # DDP(find_unused_parameters=True) semantics for eager trainers: a parameter group this rank did not use but another
        # rank did receives the averaged gradient too (autograd only returns gradients for this rank's own inputs).  Reading
        # the two flags is a host sync; the graph-captured FineTuner path resolves them on the device instead.
        flags = arena[arena.numel() - ARENA_TAIL:].tolist()
        local = set(m._decoder_params(shot_num)[0])
        for n, p in zip(names, params):
            k = param_flag_index(n)
            if k and n not in local and flags[k - 1] > 0:
                p.grad = views[n].clone() if p.grad is None else p.grad + views[n]
    return grads
