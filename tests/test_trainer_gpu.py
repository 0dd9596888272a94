"""FineTuner (fused loss + flat-arena AdamW) against the reference loop's pieces: the same model stepped with
autograd + torch.optim.AdamW (the script's optimizer, FSC_finetune_cross.py:235) must move every parameter the same way."""
import copy
import ctypes

import pytest
import torch

from oracle import synth
from test_parity_gpu import build

pytestmark = pytest.mark.gpu


def _loss_call(out, gt, mask, state=None, scale=1.0, seed=0, keep=0.8, want_mask=False, bstride=0):
    from countr_b200 import ops
    from countr_b200._lib import check, lib
    B, H, W = out.shape
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())     # noqa: E731
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
    dev = out.device
    scratch = torch.zeros(int(lib().countr_finetune_loss_scratch_bytes(B)) // 8, dtype=torch.float64, device=dev)
    result = torch.zeros(3, device=dev)
    counts = torch.zeros(B, 2, device=dev)
    dout = torch.empty(B, H, W, device=dev)
    mask_out = torch.zeros(H, W, dtype=torch.uint8, device=dev) if want_mask else None
    for _ in range(2):        # twice: the ticket must be left at zero for the next launch (CUDA-graph replay)
        check(lib().countr_finetune_loss(P(out), code[out.dtype], P(gt), code[gt.dtype], P(mask), bstride, seed, keep, P(state), scale,
                                         P(dout), P(mask_out), P(scratch), P(result), P(counts), B, H, W, ops._stream()))
    torch.cuda.synchronize()
    return result, counts, dout, mask_out


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("per_image_mask", [False, True])
def test_finetune_loss_kernel(cuda, dtype, per_image_mask):
    """FSC_finetune_cross.py:290-303 in one launch: loss, its gradient, per-image counts, batch MAE / MSE."""
    from oracle import train_oracle as TO
    B, H, W = 3, 96, 96
    out = torch.randn(B, H, W, device=cuda).to(dtype)
    gt = torch.rand(B, H, W, device=cuda).to(dtype)
    mask = (torch.rand((B, H, W) if per_image_mask else (H, W), device=cuda) < 0.8).float()
    out32 = out.float().requires_grad_(True)         # the kernel reads the 16-bit values and computes in fp32
    ref, pred, gtc, mae, mse = TO.loss_and_counts(out32, gt.float(), mask)
    (ref * 7.0).backward()
    result, counts, dout, _ = _loss_call(out, gt, mask, scale=7.0, bstride=H * W if per_image_mask else 0)
    assert abs(result[0].item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-7
    assert torch.allclose(dout, out32.grad, rtol=1e-5, atol=1e-9)
    assert torch.allclose(counts[:, 0], pred.detach(), rtol=1e-5, atol=1e-5) and torch.allclose(counts[:, 1], gtc, rtol=1e-5, atol=1e-5)
    assert abs(result[1].item() - mae.item()) < 1e-4 * abs(mae.item()) + 1e-6
    assert abs(result[2].item() - mse.item()) < 1e-4 * abs(mse.item()) + 1e-6


def test_device_side_bernoulli_mask(cuda):
    """mask == NULL: the Bernoulli(0.8) pixel mask (np.random.binomial at FSC_finetune_cross.py:290) is drawn on the device,
    one draw per step shared by the batch; bit-identical to the numpy restatement, a fresh draw every step."""
    from oracle import train_oracle as TO
    B, H, W = 2, 384, 384
    out = torch.randn(B, H, W, device=cuda)
    gt = torch.rand(B, H, W, device=cuda)
    state = torch.zeros(8, device=cuda)
    state[0] = 512.0
    masks = []
    for step in (0, 1, 5):
        state[5] = float(step)
        result, counts, dout, mk = _loss_call(out, gt, None, state=state, seed=1234, want_mask=True)
        want = torch.from_numpy(TO.bernoulli_mask(1234, step, H * W).reshape(H, W))
        assert torch.equal(mk.cpu(), want)
        ref = TO.loss_and_counts(out, gt, want.to(cuda).float())[0]
        assert abs(result[0].item() - ref.item()) < 1e-5 * abs(ref.item())
        assert torch.allclose(dout, 2 * (out - gt) * want.to(cuda).float() / (H * W * B) * 512.0, rtol=1e-5, atol=1e-9)
        masks.append(want)
    assert all(abs(m.float().mean().item() - 0.8) < 5e-3 for m in masks)
    assert (masks[0] != masks[1]).float().mean() > 0.25          # independent draws differ on 2 * .8 * .2 = 32 % of the pixels


def test_grad_stats_and_overflow_skip(cuda):
    """util/misc.py:260-301: unscale_ + inf check + get_grad_norm_, optimizer step skipped and scale halved on overflow,
    scale doubled after growth_interval clean steps; lr and the scale are read from the device state block."""
    from countr_b200 import ops
    from countr_b200._lib import check, lib
    from oracle import train_oracle as TO
    import numpy as np
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())     # noqa: E731
    n = 4 * 1000 + 8
    ps = [torch.randn(1000, device=cuda), torch.randn(30, 100, device=cuda), torch.randn(8, device=cuda)]
    flag_idx = [0, 0, 1]
    arena = torch.zeros(n + 4, device=cuda)
    offs = [0, 1000, 4000]
    rec = np.zeros(3, dtype=np.dtype([("param", "<u8"), ("goff", "<i8"), ("moff", "<i8"), ("numel", "<i8"), ("wd", "<f4"),
                                      ("step_idx", "<i4"), ("flag_idx", "<i4"), ("pad", "<i4")]))
    chunks = []
    for i, p in enumerate(ps):
        rec[i] = (p.data_ptr(), offs[i], offs[i], p.numel(), 0.05 if p.ndim > 1 else 0.0, i, flag_idx[i], 0)
        chunks += [(i, c) for c in range((p.numel() + 1023) // 1024)]
    tensors = torch.from_numpy(rec.view(np.uint8).copy()).to(cuda)
    chunks_t = torch.tensor(chunks, dtype=torch.int32, device=cuda)
    m1, m2, steps = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda), torch.zeros(3, device=cuda)
    state = torch.zeros(8, device=cuda)
    state[0], state[4] = 1024.0, 1e-3
    scratch = torch.zeros(int(lib().countr_grad_stats_scratch_bytes()) // 8, dtype=torch.float64, device=cuda)
    ref_p = [p.clone() for p in ps]
    opt = torch.optim.AdamW([{"params": [ref_p[0], ref_p[2]], "weight_decay": 0.0}, {"params": [ref_p[1]], "weight_decay": 0.05}],
                            lr=1e-3, betas=(0.9, 0.95))
    scaler = TO.ScalerState(1024.0, interval=3)

    def run(grads, flags, lr):
        state[4] = lr
        arena.zero_()
        for g, o in zip(grads, offs):
            arena[o:o + g.numel()] = g.flatten() * state[0]
        arena[n:n + 4] = torch.tensor(flags, device=cuda)
        check(lib().countr_grad_stats(P(arena), n, P(scratch), P(state), ops._stream()))
        check(lib().countr_adamw_update(P(tensors), 3, P(chunks_t), len(chunks), P(arena), ctypes.c_void_p(arena.data_ptr() + 4 * n),
                                        P(m1), P(m2), P(steps), P(state), 0.9, 0.95, 1e-8, 2.0, 0.5, 3, ops._stream()))
        torch.cuda.synchronize()

    g = torch.Generator(device="cpu").manual_seed(0)
    plan = [("ok", 1e-3), ("inf", 1e-3), ("ok", 5e-4), ("ok", 5e-4), ("nan", 2e-4), ("ok", 2e-4), ("ok", 2e-4), ("ok", 2e-4), ("unused", 2e-4)]
    for kind, lr in plan:
        grads = [torch.randn(p.shape, generator=g).to(cuda) for p in ps]
        flags = [1.0, 0.0, 0.0, 0.0]
        if kind == "inf":
            grads[1][3, 7] = float("inf")
        if kind == "nan":
            grads[0][11] = float("nan")
        if kind == "unused":
            flags[0] = 0.0
            grads[2].zero_()                  # absent gradients are zero-filled in the arena
        before = [p.clone() for p in ps]
        scale_before = state[0].item()
        run(grads, flags, lr)
        bad = kind in ("inf", "nan")
        assert bool(state[2].item()) == bad
        if bad:
            assert all(torch.equal(a, b) for a, b in zip(ps, before)), "an overflowed step must not touch the parameters"
        else:
            want = TO.grad_norm(grads if kind != "unused" else grads[:2]).item()
            assert abs(state[3].item() - want) < 1e-4 * want
            for grp in opt.param_groups:
                grp["lr"] = lr
            for rp, gg in zip(ref_p, grads):
                rp.grad = gg.clone()
            if kind == "unused":
                ref_p[2].grad = None          # DDP find_unused_parameters: no gradient anywhere -> the optimizer skips it
            opt.step()
            for a, b in zip(ps, ref_p):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
        scaler.update(bad)
        assert state[0].item() == scaler.scale, (kind, state[0].item(), scaler.scale, scale_before)
    assert steps.tolist() == [7.0, 7.0, 6.0] and state[5].item() == len(plan)


@pytest.mark.parametrize("shots", [[3, 3, 3], [3, 0, 2]])
def test_finetuner_matches_autograd_plus_torch_adamw(cuda, shots):
    from countr_b200.train import FineTuner
    m1, sd, cfg = build("small", 1, cuda)
    m2 = copy.deepcopy(m1)
    m1.train(); m2.train()
    scale = 1024.0
    decay, no_decay = [], []
    for n, p in m1.named_parameters():
        if p.requires_grad:
            (no_decay if (p.ndim == 1 or n.endswith(".bias")) else decay).append(p)
    opt = torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}], lr=1e-3, betas=(0.9, 0.95))
    tuner = FineTuner(m2, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.95), loss_scale=scale)
    start = {n: p.detach().clone() for n, p in m1.named_parameters()}
    for it, shot in enumerate(shots):
        imgs, boxes = synth.make_inputs(2, seed=100 + it)
        gt, mask = synth.make_targets(2, seed=200 + it)
        imgs, boxes, gt, mask = imgs.to(cuda), boxes.to(cuda), gt.to(cuda), mask.to(cuda)
        bx = boxes if shot else torch.empty(2, 0, device=cuda)
        out = m1(imgs, bx, shot)
        loss1 = ((out - gt) ** 2 * mask / (384 * 384)).sum() / 2
        opt.zero_grad(set_to_none=True)
        (loss1 * scale).backward()
        torch._foreach_mul_([p.grad for p in m1.parameters() if p.grad is not None], 1.0 / scale)
        opt.step()
        loss2 = tuner.step(imgs, bx, gt, mask, shot)
        torch.cuda.synchronize()
        assert abs(loss1.item() - loss2.item()) < 1e-4 * abs(loss1.item())
    p2 = dict(m2.named_parameters())
    rows = []
    for n, p in m1.named_parameters():
        if n.endswith(".attn.wk.bias") or (n.startswith("decoder_proj") and n.endswith(".bias")):
            continue      # exactly-zero true gradient (softmax shift / InstanceNorm mean removal): Adam amplifies pure noise
        d1 = (p.detach() - start[n]).double()
        d2 = (p2[n].detach() - start[n]).double()
        if d1.norm() == 0:
            assert d2.norm() == 0, n          # frozen encoder / unused parameters do not move
            continue
        rows.append(((d1 - d2).norm().item() / d1.norm().item(), n, d1.flatten()[:3].tolist(), d2.flatten()[:3].tolist()))
    rows.sort(reverse=True)
    print("\n[finetuner] largest relative differences of a parameter update vs autograd + torch.optim.AdamW:")
    for r in rows[:6]:
        print("   ", r)
    # Adam turns a gradient into a step of size ~lr * g / (|g| + eps): parameters whose gradient is at the eps = 1e-8
    # level (shot_token on its first zero-shot step) are ill-conditioned, so the bound is on the bulk of the update
    num = sum((e * 1.0) ** 2 for e, *_ in rows)
    assert sorted(e for e, *_ in rows)[int(0.9 * len(rows))] < 2e-2
    assert rows[0][0] < 0.5


@pytest.mark.parametrize("api", ["script_loop", "finetuner"])
def test_loss_curve_tracks_reference(cuda, api):
    """north_star: "training loss curves track the reference step-for-step".  tests/golden/small_curve.npz is the
    UNMODIFIED reference stepped 16 times the way FSC_finetune_cross.py:234-315 steps it (AdamW, timm weight-decay groups,
    mixed shot counts so shot_token / the exemplar CNN come and go from the optimizer's view); the CPU restatement
    reproduces it bit for bit (tests/test_oracle.py).  Here the same 16 steps run on the sm_100a path — through
    autograd + torch.optim.AdamW, and through FineTuner — with fp16 operands and a static loss scale."""
    import os
    import numpy as np
    from countr_b200.train import FineTuner
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_curve.npz"))
    C = synth.CURVE
    m, sd, cfg = build("small", 1, cuda)
    m.train()
    scale = 4096.0
    if api == "finetuner":
        tuner = FineTuner(m, lr=C["lr"], weight_decay=C["weight_decay"], betas=C["betas"], loss_scale=scale)
    else:
        opt = torch.optim.AdamW(synth.weight_decay_groups(m.named_parameters(), C["weight_decay"]), lr=C["lr"], betas=C["betas"])
    batches = [tuple(t.to(cuda) for t in b) for b in synth.curve_batches()]
    devs, cdevs = [], []
    for it in range(C["steps"]):
        imgs, boxes, gt, mask = batches[it % 2]
        shot = C["shots"][it]
        bx = boxes[:, :shot].contiguous() if shot else torch.empty(C["batch"], 0, device=cuda)
        if api == "finetuner":
            loss = tuner.step(imgs, bx, gt, mask, shot)
            out = tuner.last_output if hasattr(tuner, "last_output") else None
        else:
            out = m(imgs, bx, shot)
            loss = ((out - gt) ** 2 * mask / (384 * 384)).sum() / out.shape[0]
            opt.zero_grad(set_to_none=True)
            (loss * scale).backward()
            torch._foreach_mul_([p.grad for p in m.parameters() if p.grad is not None], 1.0 / scale)
            opt.step()
        devs.append(abs(float(loss.detach()) - g["loss"][it]) / g["loss"][it])
        if out is not None:
            cnt = (out.detach().sum((1, 2)) / 60).cpu().numpy()
            cdevs.append(float(np.abs(cnt - g["count"][it]).max()))
    print(f"\n[curve/{api}] relative loss deviation per step:", " ".join(f"{d:.1e}" for d in devs))
    if cdevs:
        print(f"[curve/{api}] max |count - reference count| per step:", " ".join(f"{d:.2f}" for d in cdevs))
    # the loss is quadratic in the map, so the 1e-3 map tolerance is 2e-3 on the loss before any optimizer feedback
    assert max(devs[:2]) < 2e-3
    assert max(devs) < 1e-2              # 16 Adam steps of fp16-operand arithmetic later (measured: <= 5e-3)


@pytest.mark.parametrize("kind", ["fused", "data_write"])
def test_parameter_updates_that_skip_the_version_counter_are_seen(cuda, kind):
    """torch.optim.AdamW(fused=True) — what a fast training script uses — updates parameters WITHOUT bumping `Tensor._version`,
    and so does a write through `.data`; the 16-bit operand copies must still follow the fp32 masters.  After one such update the
    model must produce exactly what a freshly built model with the updated state_dict produces (the eval forward is
    bit-reproducible)."""
    import copy
    m, sd, cfg = build("small", 1, cuda)
    m.train()
    imgs, boxes = synth.make_inputs(2, seed=300)
    imgs, boxes = imgs.to(cuda), boxes.to(cuda)
    out0 = m(imgs, boxes, 3)
    (out0.float().sum() * 64.0).backward()
    if kind == "fused":
        opt = torch.optim.AdamW([p for p in m.parameters() if p.grad is not None], lr=3e-3, betas=(0.9, 0.95), fused=True)
        v0 = m.decode_head3[0].weight._version
        opt.step()
        assert m.decode_head3[0].weight._version == v0, "torch now bumps the version in the fused path: the test premise changed"
    else:
        with torch.no_grad():
            for p in m.parameters():
                if p.grad is not None:
                    p.data.add_(p.grad.sign(), alpha=-3e-3)
    m.eval()
    with torch.no_grad():
        out1 = m(imgs, boxes, 3)
    fresh, _, _ = build("small", 1, cuda)
    fresh.load_state_dict(copy.deepcopy(m.state_dict()), strict=True)
    fresh.eval()
    with torch.no_grad():
        ref = fresh(imgs, boxes, 3)
    torch.cuda.synchronize()
    assert not torch.equal(out1, out0.detach()), "the update did not change the output at all"
    assert torch.equal(out1, ref), f"stale 16-bit weight copies: max |diff| = {(out1 - ref).abs().max().item():.3e}"


@pytest.mark.parametrize("which", ["finetune", "pretrain"])
def test_steady_training_step_refreshes_weights_in_one_launch(cuda, which):
    """After an optimizer step every trainable tensor needs new 16-bit copies; the forward's refresh plan must cover every
    (parameter, layout) the kernels ask for, so no per-tensor cast / transpose / pack launch happens (WeightCache.lazy_fills)."""
    from countr_b200.engine import engine
    eng = engine()
    if which == "finetune":
        m, sd, cfg = build("small", 1, cuda)
        m.train()
        imgs, boxes = synth.make_inputs(2, seed=310)
        imgs, boxes = imgs.to(cuda), boxes.to(cuda)
        run = lambda shot: m(imgs, boxes if shot else torch.empty(2, 0, device=cuda), shot).float().sum()
        shots = [3, 0, 2]
    else:
        import models_mae_noct as N
        m = N.MaskedAutoencoderViTNoCT(embed_dim=256, depth=2, num_heads=4, decoder_depth=1).to(cuda).train()
        imgs = torch.rand(2, 3, 384, 384, device=cuda)
        run = lambda shot: m(imgs, mask_ratio=0.5)[0]
        shots = [0, 0, 0]
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, fused=True)
    for i, shot in enumerate(shots + shots):
        if i == len(shots):
            before = eng.wc.lazy_fills          # every shot count has been seen once: caches and job tables exist
        opt.zero_grad(set_to_none=True)
        (run(shot) * 16.0).backward()
        opt.step()
    torch.cuda.synchronize()
    assert eng.wc.lazy_fills == before, f"{eng.wc.lazy_fills - before} per-tensor weight refreshes in steady-state steps"


def test_arena_adamw_matches_torch(cuda):
    """ArenaAdamW (the pre-training step's optimizer: unscale + inf check + grad norm + AdamW + GradScaler.update over a flat arena)
    against torch.optim.AdamW with the timm add_weight_decay grouping, three steps, plus the overflow skip."""
    from countr_b200.dist import build_grad_arena
    from countr_b200.train import ArenaAdamW
    torch.manual_seed(3)
    shapes = {"a.weight": (64, 48), "a.bias": (64,), "norm.weight": (48,), "tok": (1, 1, 48), "b.weight": (33, 7, 3, 3)}
    params = [torch.nn.Parameter(torch.randn(s, device=cuda)) for s in shapes.values()]
    names = list(shapes)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in params]
    decay = [r for n, r in zip(names, ref) if not (r.ndim == 1 or n.endswith(".bias"))]
    no_decay = [r for n, r in zip(names, ref) if r.ndim == 1 or n.endswith(".bias")]
    topt = torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}], lr=1e-3, betas=(0.9, 0.95))
    opt = ArenaAdamW(names, params, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.95), loss_scale=512.0, dynamic_scale=True)
    for step in range(3):
        arena, views = build_grad_arena(names, params, cuda)
        arena.zero_()
        total = 0.0
        for n, r in zip(names, ref):
            gr = torch.randn_like(r)
            r.grad = gr.clone()
            views[n].copy_(gr * 512.0)
            total += gr.double().pow(2).sum().item()
        opt.step(arena)
        topt.step()
        m = opt.metrics()
        assert not m["found_inf"] and abs(m["grad_norm"] - total ** 0.5) < 1e-4 * total ** 0.5
    for n, p, r in zip(names, params, ref):
        assert torch.allclose(p, r, rtol=2e-5, atol=2e-6), n
    before = [p.detach().clone() for p in params]
    arena, views = build_grad_arena(names, params, cuda)
    arena.zero_()
    views["a.weight"][0, 0] = float("inf")
    opt.step(arena)
    m = opt.metrics()
    assert m["found_inf"] and m["loss_scale"] == 256.0                       # skipped, scale backed off
    assert all(torch.equal(p, b) for p, b in zip(params, before))
