"""CPU tests: the oracle restatement (oracle/countr_oracle.py) against golden vectors produced by
the reference itself (scripts/gen_golden.py, run in the build container).  fp32 on CPU is
reproducible to ~1e-6 relative across thread counts, so the tolerance is 2e-5 (matrix-level)."""
import os

import numpy as np
import pytest
import torch

from oracle import countr_oracle as O
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_state_dict_spec_counts():
    """99,690,625 parameters in the reference base model."""
    n = sum(int(np.prod(s)) for _, s, _ in synth.state_dict_spec(synth.CONFIGS["base"]))
    assert n == 99_690_625


def test_base_c1_forward_matches_reference():
    g = np.load(os.path.join(GOLD, "base_c1.npz"))
    cfg = synth.CONFIGS["base"]
    sd = synth.make_state_dict(cfg, seed=0)
    imgs, boxes = synth.make_inputs(1, seed=1234)
    taps = {}
    with torch.no_grad():
        out = O.forward(sd, cfg, imgs, boxes, 3, taps)
        out0 = O.forward(sd, cfg, imgs, torch.empty(1, 0), 0)
    assert out.shape == (1, 384, 384)
    assert rel(out, g["out"]) < 2e-5
    assert rel(taps["latent"][0, :8, :32], g["latent_head"]) < 2e-5
    assert rel(taps["latent"][0].sum(-1), g["latent_rowsum"]) < 2e-4
    assert abs(out.sum().item() / 60 - g["out"].sum() / 60) < 1e-2          # the count
    assert rel(torch.nn.functional.avg_pool2d(out0[:, None], 8)[:, 0], g["out_zero_pool8"]) < 2e-5


@pytest.mark.parametrize("shot", [0, 1, 2, 3, 5])
def test_small_forward_and_loss_match_reference(shot):
    g = np.load(os.path.join(GOLD, "small_fwd.npz"))
    cfg = synth.CONFIGS["small"]
    sd = synth.make_state_dict(cfg, seed=1)
    imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
    gt, mask = synth.make_targets(2, seed=78)
    with torch.no_grad():
        out = O.forward(sd, cfg, imgs, boxes if shot else torch.empty(2, 0), shot)
    assert rel(torch.nn.functional.avg_pool2d(out[:, None], 8)[:, 0], g[f"out_pool8_s{shot}"]) < 2e-5
    assert rel(out[:, [0, 100, 383]], g[f"out_rows_s{shot}"]) < 2e-5
    assert rel(out.sum((1, 2)), g[f"sum_s{shot}"]) < 2e-5
    assert rel(O.finetune_loss(out, gt, mask), g[f"loss_s{shot}"]) < 2e-5


@pytest.mark.parametrize("shot", [0, 3])
def test_small_decoder_grads_match_reference(shot):
    g = np.load(os.path.join(GOLD, "small_grads.npz"))
    cfg = synth.CONFIGS["small"]
    sd = synth.make_state_dict(cfg, seed=1)
    names = O.decoder_param_names(sd, shot)
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
    gt, mask = synth.make_targets(2, seed=78)
    out = O.forward(sd, cfg, imgs, boxes if shot else torch.empty(2, 0), shot)
    O.finetune_loss(out, gt, mask).backward()
    golden_names = sorted({k.split("/")[1] for k in g.files if k.startswith(f"s{shot}/")})
    assert sorted(names) == golden_names          # exactly the parameters the reference gives a grad
    for n in names:
        gr = sd[n].grad.flatten()
        ref_norm = float(g[f"s{shot}/{n}/norm"])
        assert abs(gr.norm().item() - ref_norm) <= 5e-4 * ref_norm + 1e-12, n
        assert rel(gr[:16], g[f"s{shot}/{n}/head"]) < 1e-3 or ref_norm < 1e-10, n


def test_finetune_loss_curve_tracks_reference():
    """north_star: "training loss curves track the reference step-for-step".  The oracle stepped with torch.optim.AdamW the
    way FSC_finetune_cross.py:234-315 steps the reference reproduces the reference's own 16-step curve."""
    g = np.load(os.path.join(GOLD, "small_curve.npz"))
    cfg = synth.CONFIGS["small"]
    sd = synth.make_state_dict(cfg, seed=1)
    start = {k: v.clone() for k, v in sd.items()}
    # few-shot steps leave shot_token without a gradient and zero-shot steps leave the exemplar CNN without one: AdamW
    # skips a parameter whose .grad is None (no decay, no moment update), so the optimizer holds the union
    names = sorted(set(O.decoder_param_names(sd, 3)) | set(O.decoder_param_names(sd, 0)))
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    C = synth.CURVE
    opt = torch.optim.AdamW(synth.weight_decay_groups([(n, sd[n]) for n in names], C["weight_decay"]), lr=C["lr"], betas=C["betas"])
    batches = synth.curve_batches()
    devs = []
    for it in range(C["steps"]):
        imgs, boxes, gt, mask = batches[it % 2]
        shot = C["shots"][it]
        out = O.forward(sd, cfg, imgs, boxes[:, :shot] if shot else torch.empty(C["batch"], 0), shot)
        loss = O.finetune_loss(out, gt, mask)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        devs.append(abs(loss.item() - g["loss"][it]) / g["loss"][it])
        assert rel(out.detach().sum((1, 2)) / 60, g["count"][it]) < 2e-3
    print("\n[curve] oracle vs reference, relative loss deviation per step:", " ".join(f"{d:.1e}" for d in devs))
    assert max(devs) < 2e-4
    for n in names:
        d = (sd[n].detach() - start[n]).norm().item()
        ref = float(g[f"final/{n}/delta_norm"])
        assert abs(d - ref) <= 2e-3 * ref + 1e-9, (n, d, ref)


# ------------------------------------------------------------------------------------------------
# MAE pre-training model (models_mae_noct.py) — oracle vs reference-generated goldens
# ------------------------------------------------------------------------------------------------
NOCT_SMALL = dict(img_size=384, patch_size=16, embed_dim=256, depth=2, num_heads=4, decoder_embed_dim=512, decoder_depth=2,
                  decoder_num_heads=16, mlp_ratio=4, eps=1e-6)


def noct_noise(n, l, seed=123):
    torch.manual_seed(seed)         # the reference's forward draws torch.rand(N, L) first (models_mae_noct.py:118)
    return torch.rand(n, l)


@pytest.mark.parametrize("norm_pix", [False, True])
def test_noct_oracle_matches_reference(norm_pix):
    from oracle import noct_oracle as NO
    g = np.load(os.path.join(GOLD, "noct_small.npz"))
    tag = "np1" if norm_pix else "np0"
    sd = NO.make_state_dict(NOCT_SMALL, seed=2)
    train = [k for k in sd if k not in ("pos_embed", "decoder_pos_embed")]
    for k in train:
        sd[k] = sd[k].clone().requires_grad_(True)
    imgs, _ = synth.make_inputs(2, seed=91)
    loss, pred, mask = NO.forward(sd, NOCT_SMALL, imgs, 0.5, noct_noise(2, 576), norm_pix)
    loss.backward()
    assert rel(loss, g[f"{tag}/loss"]) < 1e-5
    assert torch.equal(mask, torch.from_numpy(g[f"{tag}/mask"]))
    assert rel(pred[:, :4, :64], g[f"{tag}/pred_head"]) < 2e-5
    assert rel(pred.sum(-1), g[f"{tag}/pred_rowsum"]) < 2e-4
    for k in train:
        ref_norm = float(g[f"{tag}/g/{k}/norm"])
        assert abs(sd[k].grad.norm().item() - ref_norm) <= 1e-3 * ref_norm + 1e-9, k


def test_oracle_matches_reference_at_the_benchmarked_config():
    """tests/golden/base_b8.npz is the UNMODIFIED reference on the bench workload (base model, batch 8, 3 shots).  The images of
    a batch are independent, so the oracle is pinned on the first two of them here (the GPU test pins the gradients)."""
    g = np.load(os.path.join(GOLD, "base_b8.npz"))
    cfg = synth.CONFIGS["base"]
    sd = synth.make_state_dict(cfg, seed=0)
    imgs, boxes = synth.make_inputs(8, seed=1234)
    with torch.no_grad():
        lat = O.forward_encoder(sd, cfg, imgs[:2])
        out = O.forward(sd, cfg, imgs[:2], boxes[:2], 3)
    pool = torch.nn.functional.avg_pool2d(out[:, None], 8)[:, 0]
    assert rel(lat[:, ::8, ::8], g["latent_sub"][:2]) < 2e-5
    assert rel(lat.norm(dim=-1), g["latent_rownorm"][:2]) < 2e-5
    assert rel(pool, g["out_pool8"][:2]) < 2e-5
    assert rel(out.sum((1, 2)), g["out_sum"][:2]) < 2e-5


def test_scaler_state_follows_torch_gradscaler():
    """oracle/train_oracle.ScalerState (what the device-side loss-scale update of countr_b200.train is checked against) against
    torch's own GradScaler — the class util/misc.py:260-287 wraps — over a sequence with overflow steps; the skipped updates too."""
    from oracle import train_oracle as TO
    p = torch.nn.Parameter(torch.ones(4))
    opt = torch.optim.SGD([p], lr=0.1)
    sc = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=3)
    st = TO.ScalerState(1024.0, interval=3)
    for bad in [0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0]:
        before = p.detach().clone()
        opt.zero_grad()
        sc.scale((p * p).sum()).backward()
        if bad:
            p.grad[0] = float("inf")
        sc.unscale_(opt)
        sc.step(opt)
        sc.update()
        st.update(bool(bad))
        assert sc.get_scale() == st.scale
        assert torch.equal(p.detach(), before) == bool(bad)          # an overflow step leaves the parameters alone
