"""GPU tests of the backward kernels (each against torch autograd of the same fp32 expression)
and of the whole decoder backward (against autograd through the CPU oracle and the goldens the
reference produced).  Gradient tolerances are looser than forward ones: activations' gradients
travel as fp16 GEMM operands, exactly as in the reference's fp16-autocast training."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import countr_oracle as O
from oracle import synth
from test_kernels_gpu import _rand16, assert_close
from test_parity_gpu import build, rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_upsample2x_bwd(cuda):
    from countr_b200 import ops
    x = torch.randn(2, 24, 24, device=cuda, requires_grad=True)
    dy = torch.randn(2, 48, 48, device=cuda)
    F.interpolate(x[:, None], scale_factor=2, mode="bilinear", align_corners=False)[:, 0].backward(dy)
    dx = torch.empty(2, 24, 24, device=cuda)
    ops.upsample2x_bwd(dy, dx)
    assert_close(dx.reshape(2, -1), x.grad.reshape(2, -1), 1e-5, "up2 bwd f32")
    ops.upsample2x_bwd(dy.half(), dx)
    assert_close(dx.reshape(2, -1), x.grad.reshape(2, -1), 1e-3, "up2 bwd f16")


@pytest.mark.parametrize("mode,B,H,W", [(0, 2, 24, 24), (1, 2, 24, 24), (0, 1, 30, 21), (0, 3, 5, 9), (0, 8, 48, 48), (0, 1, 96, 100),
                                        (1, 1, 30, 21), (1, 8, 96, 96), (1, 3, 5, 9), (1, 2, 100, 100)])
def test_gn_relu_backward(cuda, mode, B, H, W):
    """mode 0 (up-sample adjoint gather, shared-memory staged tiles): whole tiles, ragged tiles on both axes, maps smaller than a
    tile, several tiles per persistent block; mode 1: the 1x1-conv head (64-pixel bulk-copy tiles: a ragged last tile, the stage ring
    wrapping, a map below the staged kernel's minimum size)."""
    from countr_b200 import ops
    C, G = 256, 8
    raw = _rand16((B, H, W, C), cuda, seed=20, scale=2.0)
    gamma = (torch.randn(C, device=cuda) * 0.5 + 1).requires_grad_(True)
    beta = (torch.randn(C, device=cuda) * 0.3).requires_grad_(True)
    x = raw.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    z = F.relu(F.group_norm(x, G, gamma, beta, 1e-5))
    xg = raw.double().reshape(B, H * W, G, C // G)
    stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()
    dyh = torch.empty_like(raw)
    dgamma = torch.zeros(C, device=cuda)
    dbeta = torch.zeros(C, device=cuda)
    gsum = torch.zeros(B, G, 2, device=cuda, dtype=torch.float64)
    dbias = torch.zeros(C, device=cuda)
    if mode == 0:
        d_next = _rand16((B, 2 * H, 2 * W, C), cuda, seed=21)
        F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False).backward(d_next.float().permute(0, 3, 1, 2))
        ops.gn_relu_bwd_reduce(raw, stats, gamma.detach(), beta.detach(), dyh, dgamma, dbeta, gsum, G, 1e-5, d_next=d_next)
    else:
        w1 = (torch.randn(C, device=cuda) * 0.1).requires_grad_(True)
        b1 = torch.zeros(1, device=cuda, requires_grad=True)
        dmap = torch.randn(B, H, W, device=cuda)
        F.conv2d(z, w1.reshape(1, C, 1, 1), b1).squeeze(1).backward(dmap)
        dw1 = torch.zeros(C, device=cuda)
        db1 = torch.zeros(1, device=cuda)
        ops.gn_relu_bwd_reduce(raw, stats, gamma.detach(), beta.detach(), dyh, dgamma, dbeta, gsum, G, 1e-5, dmap=dmap,
                               w1=w1.detach(), dw1=dw1, db1=db1)
    ops.gn_bwd_apply(raw, dyh, stats, gsum, gamma.detach(), dyh, dbias, G, 1e-5)
    torch.cuda.synchronize()
    ref_dx = x.grad.permute(0, 2, 3, 1)
    assert_close(dyh.reshape(-1, C), ref_dx.reshape(-1, C), 2e-3, f"gn bwd dx mode {mode}")
    assert_close(dgamma[None], gamma.grad[None], 2e-3, "gn dgamma")
    assert_close(dbeta[None], beta.grad[None], 2e-3, "gn dbeta")
    assert_close(dbias[None], ref_dx.reshape(-1, C).sum(0)[None], 3e-2, "conv dbias")   # sums to ~0 in exact arithmetic
    if mode == 1:
        assert_close(dw1[None], w1.grad[None], 1e-4, "dw1")
        assert_close(db1[None], b1.grad[None], 1e-4, "db1")


def test_colsum_and_softmax_bwd(cuda):
    from countr_b200 import ops
    x = torch.randn(4608, 1536, device=cuda)
    out = torch.zeros(1536, device=cuda)
    ops.colsum(x, out)
    assert_close(out[None], x.sum(0)[None], 1e-5, "colsum f32")
    out.zero_()
    ops.colsum(x.half(), out)
    assert_close(out[None], x.half().float().sum(0)[None], 1e-5, "colsum f16")
    rows, L = 1000, 576
    s = torch.randn(rows, L, device=cuda).half()
    dp = torch.randn(rows, L, device=cuda).half()
    lse = torch.logsumexp(s.float(), -1)
    sr = s.float().requires_grad_(True)
    sr.softmax(-1).backward(dp.float())
    s_io, dp_io = s.clone(), dp.clone()
    ops.softmax_bwd_rows(s_io, dp_io, lse, 0.5)
    assert_close(s_io, s.float().softmax(-1), 1e-3, "P recompute")
    assert_close(dp_io, sr.grad * 0.5, 2e-3, "softmax bwd")


@pytest.mark.parametrize("B,L,H,dh,fused", [(2, 576, 16, 32, True), (2, 576, 16, 32, False), (1, 200, 3, 32, True), (1, 640, 2, 32, True),
                                             (1, 288, 4, 64, False), (2, 288, 12, 64, True), (1, 576, 3, 64, True), (1, 200, 2, 64, True),
                                             (3, 144, 2, 64, True)])
def test_attention_backward(cuda, B, L, H, dh, fused):
    from countr_b200 import ops
    from countr_b200.backward import attention_backward
    qkv = _rand16((B, L, 3, H, dh), cuda, seed=30)
    datt = _rand16((B * L, H * dh), cuda, seed=31)
    out = torch.empty(B, L, H * dh, device=cuda, dtype=torch.float16)
    lse = torch.empty(B, H, L, device=cuda)
    scale = dh ** -0.5
    ops.attention_fwd(qkv, out, B, L, H, dh, scale, lse=lse)
    dqkv = attention_backward(qkv.view(B * L, -1), lse, datt, B, L, H, dh, scale, att=out.view(B * L, -1), fused=fused)
    x = qkv.float().requires_grad_(True)
    q, k, v = [x[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    o = ((q @ k.transpose(-1, -2) * scale).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * L, H * dh)
    o.backward(datt.float())
    torch.cuda.synchronize()
    assert_close(dqkv, x.grad.reshape(B * L, -1), 4e-3, f"attention bwd fused={fused} {B},{L},{H},{dh}")


@pytest.mark.parametrize("S,bcast", [(3, False), (1, False), (5, False), (1, True)])
def test_cross_attn_core_bwd(cuda, S, bcast):
    from countr_b200 import ops
    B, L, D, dh = 2, 576, 512, 32
    Hh = D // dh
    q = _rand16((B * L, D), cuda, seed=40)
    nb = 1 if bcast else B
    k = torch.randn(nb, S, D, device=cuda)
    v = torch.randn(nb, S, D, device=cuda)
    do = _rand16((B * L, D), cuda, seed=41)
    out = torch.empty(B * L, D, device=cuda, dtype=torch.float16)
    probs = torch.empty(B * L, Hh, S, device=cuda)
    ops.cross_attn_core(q, k, v, out, B, L, S, D, dh, dh ** -0.5, probs=probs, kv_broadcast=bcast)
    dq = torch.empty_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    ops.cross_attn_core_bwd(q, k, v, probs, do, dq, dk, dv, B, L, S, D, dh, dh ** -0.5, kv_broadcast=bcast)
    qr = q.float().requires_grad_(True)
    kr = k.clone().requires_grad_(True)
    vr = v.clone().requires_grad_(True)
    qh = qr.reshape(B, L, Hh, dh).permute(0, 2, 1, 3)
    kh = kr.expand(B, S, D).reshape(B, S, Hh, dh).permute(0, 2, 1, 3)
    vh = vr.expand(B, S, D).reshape(B, S, Hh, dh).permute(0, 2, 1, 3)
    o = (((qh @ kh.transpose(-1, -2)) * dh ** -0.5).softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B * L, D)
    o.backward(do.float())
    torch.cuda.synchronize()
    assert_close(dq, qr.grad, 2e-3, "cross dq")
    assert_close(dk.reshape(-1, D), kr.grad.reshape(-1, D), 1e-4, "cross dk")
    assert_close(dv.reshape(-1, D), vr.grad.reshape(-1, D), 1e-4, "cross dv")


@pytest.mark.parametrize("mode", [0, 1])
def test_inorm_relu_pool_bwd(cuda, mode):
    from countr_b200 import ops
    N, H, W, C = 3, 16, 16, 128
    raw = _rand16((N, H, W, C), cuda, seed=50, scale=2.0)
    mean = torch.empty(N, C, device=cuda)
    rstd = torch.empty(N, C, device=cuda)
    x = raw.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    z = F.relu(F.instance_norm(x, eps=1e-5))
    d_raw = torch.empty_like(raw)
    dbias = torch.zeros(C, device=cuda)
    if mode == 0:
        pooled = torch.empty(N, H // 2, W // 2, C, device=cuda, dtype=torch.float16)
        ops.inorm_relu_pool(raw, 0, 1e-5, y16=pooled, mean=mean, rstd=rstd)
        dp = _rand16((N, H // 2, W // 2, C), cuda, seed=51)
        F.max_pool2d(z, 2).backward(dp.float().permute(0, 3, 1, 2))
        ops.inorm_relu_pool_bwd(raw, mean, rstd, d_raw, 0, dpool16=dp, dbias=dbias)
    else:
        y32 = torch.empty(N, C, device=cuda)
        ops.inorm_relu_pool(raw, 1, 1e-5, y32=y32, mean=mean, rstd=rstd)
        dp = torch.randn(N, C, device=cuda)
        z.mean((2, 3)).backward(dp)
        ops.inorm_relu_pool_bwd(raw, mean, rstd, d_raw, 1, dpool32=dp, dbias=dbias)
    torch.cuda.synchronize()
    assert_close(d_raw.reshape(-1, C), x.grad.permute(0, 2, 3, 1).reshape(-1, C), 2e-3, f"IN bwd mode {mode}")


def test_conv_weight_grads(cuda):
    from countr_b200 import ops
    for (B, H, W, Cin, Cout) in [(2, 24, 24, 512, 256), (2, 48, 48, 256, 256), (3, 8, 8, 256, 512), (3, 32, 32, 64, 128)]:
        x = _rand16((B, H, W, Cin), cuda, seed=60)
        dy = _rand16((B, H, W, Cout), cuda, seed=61)
        w = torch.zeros(Cout, Cin, 3, 3, device=cuda, requires_grad=True)
        F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        dwp = torch.zeros(Cout, 9 * Cin, device=cuda)
        ops.conv3x3_dw(dy, x, dwp)
        dw = torch.empty(Cout, Cin, 3, 3, device=cuda)
        ops.conv_dw_unpack(dwp, dw, Cout, Cin)
        torch.cuda.synchronize()
        assert_close(dw.reshape(Cout, -1), w.grad.reshape(Cout, -1), 1e-4, f"conv dW {B}x{H}x{W} {Cin}->{Cout}")
    # stage-1 exemplar conv (Cin = 3): direct kernel
    boxes = torch.rand(2, 3, 3, 64, 64, device=cuda)
    d_raw = _rand16((2 * 2, 64, 64, 64), cuda, seed=62)
    w = torch.zeros(64, 3, 3, 3, device=cuda, requires_grad=True)
    F.conv2d(boxes[:, :2].reshape(4, 3, 64, 64), w, padding=1).backward(d_raw.float().permute(0, 3, 1, 2))
    dw = torch.zeros(64, 3, 3, 3, device=cuda)
    ops.exemplar_conv1_dw(boxes, 2, d_raw, dw)
    torch.cuda.synchronize()
    assert_close(dw.reshape(64, -1), w.grad.reshape(64, -1), 1e-4, "exemplar conv1 dW")


LOSS_SCALE = 4096.0   # stand-in for GradScaler (util/misc.py:260-270): keeps fp16 gradients out of the subnormals


@pytest.mark.parametrize("shot", [3, 0])
def test_decoder_gradients_match_oracle_and_golden(cuda, shot):
    g = np.load(os.path.join(GOLD, "small_grads.npz"))
    m, sd, cfg = build("small", 1, cuda)
    m.train()
    imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
    gt, mask = synth.make_targets(2, seed=78)
    bx = boxes.to(cuda) if shot else torch.empty(2, 0, device=cuda)
    out = m(imgs.to(cuda), bx, shot)
    loss = O.finetune_loss(out, gt.to(cuda), mask.to(cuda))
    (loss * LOSS_SCALE).backward()
    # oracle autograd on CPU
    names = O.decoder_param_names(sd, shot)
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    ref_out = O.forward(sd, cfg, imgs, boxes if shot else torch.empty(2, 0), shot)
    O.finetune_loss(ref_out, gt, mask).backward()
    got_names = sorted(n for n, p in m.named_parameters() if p.grad is not None)
    assert got_names == sorted(names)          # same set of parameters receives a gradient as in the reference
    # conv biases in front of InstanceNorm have an exactly-zero true gradient (the norm removes the
    # mean): both sides only hold rounding noise there, so norms are compared with a floor tied to
    # the overall gradient scale.
    all_ref = torch.cat([sd[n].grad.flatten().double() for n in names])
    floor = 1e-5 * all_ref.norm().item()
    num, den, table = 0.0, 0.0, []
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        got = (p.grad / LOSS_SCALE).double().cpu()
        ref = sd[n].grad.double()
        num += (got - ref).pow(2).sum().item()
        den += ref.pow(2).sum().item()
        table.append((n, ref.norm().item(), (got - ref).norm().item(), float(g[f"s{shot}/{n}/norm"])))
    total = (num / den) ** 0.5
    lref = float(np.load(os.path.join(GOLD, "small_fwd.npz"))[f"loss_s{shot}"])
    # per-parameter: error relative to the parameter's own gradient norm, with a floor tied to the overall
    # gradient scale (wk.bias shifts every exemplar logit equally -> softmax-invariant -> true gradient 0)
    worst = max(table, key=lambda t: t[2] / (t[1] + floor))
    out_dir = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"grad_parity_s{shot}.txt"), "w") as f:
            for t in sorted(table, key=lambda t: -t[2] / (t[1] + floor)):
                f.write(f"{t[0]:48s} ref_norm={t[1]:.3e} err_norm={t[2]:.3e} rel={t[2] / (t[1] + floor):.3e} golden_norm={t[3]:.3e}\n")
    print(f"\n[grad parity small shot={shot}] loss rel={abs(loss.item() - lref) / lref:.1e} all-grads relL2={total:.3e} "
          f"worst param {worst[0]} rel={worst[2] / (worst[1] + floor):.3e}")
    assert abs(loss.item() - lref) / lref < 2e-3
    assert total < 1e-2
    for n, ref_norm, err_norm, gold_norm in table:
        assert abs(ref_norm - gold_norm) <= 1e-3 * gold_norm + floor, n          # oracle autograd == reference autograd
        # The exemplar CNN routes gradients through MaxPool arg-max / ReLU masks.  With 16-bit activation
        # storage (ours, and the reference's own fp16-autocast training) ~1 % of those decisions flip
        # against an fp32 run, which moves the conv weight gradients by 7-10 % in L2 although every
        # kernel is exact (measured on the oracle itself, see test below): loose bound here, tight
        # bound against the oracle with the same storage rounding in the next test.
        tol = 0.15 if n.startswith("decoder_proj") else 5e-2
        assert err_norm <= tol * ref_norm + 20 * floor, (n, ref_norm, err_norm)


def test_exemplar_cnn_gradients_with_matched_storage_rounding(cuda):
    """Exemplar-CNN parameter gradients against oracle autograd that rounds the conv outputs / pooled
    maps to fp16 at the same points (straight-through), i.e. takes the same arg-max / ReLU routing."""
    m, sd, cfg = build("small", 1, cuda)
    m.train()
    imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
    gt, mask = synth.make_targets(2, seed=78)
    out = m(imgs.to(cuda), boxes.to(cuda), 3)
    (O.finetune_loss(out, gt.to(cuda), mask.to(cuda)) * LOSS_SCALE).backward()
    names = [n for n in O.decoder_param_names(sd, 3) if n.startswith("decoder_proj")]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    ref_out = O.forward(sd, cfg, imgs, boxes, 3, exemplar_round16=True)
    O.finetune_loss(ref_out, gt, mask).backward()
    params = dict(m.named_parameters())
    for n in names:
        if n.endswith(".bias"):
            continue          # exactly-zero true gradient (InstanceNorm removes the mean)
        got = (params[n].grad / LOSS_SCALE).double().cpu()
        ref = sd[n].grad.double()
        e = ((got - ref).norm() / ref.norm()).item()
        print(f"[exemplar grads, matched rounding] {n}: relL2={e:.3e}")
        # even with matched storage points the two pipelines round slightly different fp32 values, so a
        # (different) ~1 % of the arg-max / ReLU decisions still flips; the stage-level kernels are checked
        # tightly (2e-3) on identical inputs in test_inorm_relu_pool_bwd / test_conv_weight_grads.
        assert e < 0.15, (n, e)


def test_groupnorm_kernels_bf16(cuda):
    """The bf16 switch of the staged GroupNorm kernels (forward up-sample and 1x1 head, backward gather and head reduce): the
    mixed-precision FMA and the packers take a different instruction path than fp16."""
    from countr_b200 import ops
    bf = torch.bfloat16
    B, C, G = 2, 256, 8

    def stats_of(raw):
        xg = raw.double().reshape(raw.shape[0], -1, G, C // G)
        return torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()

    gamma0 = torch.randn(C, device=cuda) * 0.5 + 1
    beta0 = torch.randn(C, device=cuda) * 0.3
    # forward: up-sample (48^2 -> staged kernel) and 1x1 head (64^2 = 4096 pixels -> staged kernel)
    x = _rand16((B, 48, 48, C), cuda, bf, seed=70, scale=2.0)
    y = torch.empty(B, 96, 96, C, device=cuda, dtype=bf)
    ops.gn_relu_upsample2x(x, stats_of(x), gamma0, beta0, y, G, 1e-5)
    act = F.relu(F.group_norm(x.float().permute(0, 3, 1, 2), G, gamma0, beta0, 1e-5))
    ref = F.interpolate(act, scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert_close(y.reshape(-1, C), ref.reshape(-1, C), 8e-3, "bf16 gn+relu+up2")
    x2 = _rand16((B, 64, 64, C), cuda, bf, seed=71, scale=2.0)
    w = torch.randn(C, device=cuda) * 0.1
    bias = torch.randn(1, device=cuda)
    d = torch.empty(B, 64, 64, device=cuda)
    ops.gn_relu_conv1x1(x2, stats_of(x2), gamma0, beta0, w, bias, d, G, 1e-5)
    act2 = F.relu(F.group_norm(x2.float().permute(0, 3, 1, 2), G, gamma0, beta0, 1e-5))
    assert_close(d.reshape(B, -1), F.conv2d(act2, w.reshape(1, C, 1, 1), bias).reshape(B, -1), 1e-4, "bf16 gn+relu+1x1")
    # backward, both modes
    for mode, (H, W) in ((0, (24, 24)), (1, (24, 24))):
        raw = _rand16((B, H, W, C), cuda, bf, seed=72 + mode, scale=2.0)
        gamma = gamma0.clone().requires_grad_(True)
        beta = beta0.clone().requires_grad_(True)
        xin = raw.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        z = F.relu(F.group_norm(xin, G, gamma, beta, 1e-5))
        stats = stats_of(raw)
        dyh = torch.empty_like(raw)
        dgamma, dbeta = torch.zeros(C, device=cuda), torch.zeros(C, device=cuda)
        gsum = torch.zeros(B, G, 2, device=cuda, dtype=torch.float64)
        dbias = torch.zeros(C, device=cuda)
        if mode == 0:
            d_next = _rand16((B, 2 * H, 2 * W, C), cuda, bf, seed=80)
            F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False).backward(d_next.float().permute(0, 3, 1, 2))
            ops.gn_relu_bwd_reduce(raw, stats, gamma.detach(), beta.detach(), dyh, dgamma, dbeta, gsum, G, 1e-5, d_next=d_next)
        else:
            w1 = (torch.randn(C, device=cuda) * 0.1).requires_grad_(True)
            b1 = torch.zeros(1, device=cuda, requires_grad=True)
            dmap = torch.randn(B, H, W, device=cuda)
            F.conv2d(z, w1.reshape(1, C, 1, 1), b1).squeeze(1).backward(dmap)
            dw1, db1 = torch.zeros(C, device=cuda), torch.zeros(1, device=cuda)
            ops.gn_relu_bwd_reduce(raw, stats, gamma.detach(), beta.detach(), dyh, dgamma, dbeta, gsum, G, 1e-5, dmap=dmap,
                                   w1=w1.detach(), dw1=dw1, db1=db1)
            assert_close(dw1[None], w1.grad[None], 2e-3, "bf16 dw1")
        ops.gn_bwd_apply(raw, dyh, stats, gsum, gamma.detach(), dyh, dbias, G, 1e-5)
        torch.cuda.synchronize()
        assert_close(dyh.reshape(-1, C), xin.grad.permute(0, 2, 3, 1).reshape(-1, C), 1.5e-2, f"bf16 gn bwd dx mode {mode}")
        assert_close(dgamma[None], gamma.grad[None], 1.5e-2, f"bf16 dgamma mode {mode}")
        assert_close(dbeta[None], beta.grad[None], 1.5e-2, f"bf16 dbeta mode {mode}")
