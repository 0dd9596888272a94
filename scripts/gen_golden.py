"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Runs only in the build container (the GPU box has no /root/reference).  The reference needs
timm==0.4.9, which is not installed and cannot be (no network): a minimal shim provides
`timm.models.vision_transformer.{PatchEmbed,Block}` built from the reference's OWN
models_crossvit.Attention / Mlp (verbatim copies of timm's, models_crossvit.py:46-94), and
`np.float` is aliased for util/pos_embed.py:56 under numpy >= 1.24.  Nothing is copied from the
reference; its files are imported by path.

    python scripts/gen_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle import countr_oracle as O  # noqa: E402


def import_reference():
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    import models_crossvit as ref_xvit  # the reference's own file

    class PatchEmbed(nn.Module):  # timm 0.4.9 vision_transformer.PatchEmbed
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
            super().__init__()
            self.img_size = (img_size, img_size)
            self.patch_size = (patch_size, patch_size)
            self.num_patches = (img_size // patch_size) ** 2
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

        def forward(self, x):
            B, C, H, W = x.shape
            assert H == self.img_size[0] and W == self.img_size[1]
            return self.proj(x).flatten(2).transpose(1, 2)

    class Block(nn.Module):  # timm 0.4.9 vision_transformer.Block (drop_path=0 -> identity)
        def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                     drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
            super().__init__()
            self.norm1 = norm_layer(dim)
            self.attn = ref_xvit.Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                           attn_drop=attn_drop, proj_drop=drop)
            self.drop_path = nn.Identity()
            self.norm2 = norm_layer(dim)
            self.mlp = ref_xvit.Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

        def forward(self, x):
            x = x + self.drop_path(self.attn(self.norm1(x)))
            x = x + self.drop_path(self.mlp(self.norm2(x)))
            return x

    timm = types.ModuleType("timm")
    timm.__version__ = "0.4.9"
    timm.models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.PatchEmbed, vt.Block = PatchEmbed, Block
    timm.models.vision_transformer = vt
    sys.modules.update({"timm": timm, "timm.models": timm.models, "timm.models.vision_transformer": vt})
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            tv = types.ModuleType("torchvision")
            tv.utils = types.ModuleType("torchvision.utils")
            sys.modules.update({"torchvision": tv, "torchvision.utils": tv.utils})
    spec = importlib.util.spec_from_file_location("ref_models_mae_cross", os.path.join(REF, "models_mae_cross.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    spec2 = importlib.util.spec_from_file_location("ref_models_mae_noct", os.path.join(REF, "models_mae_noct.py"))
    mod.noct = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(mod.noct)
    return mod


def build_ref_model(ref, cfg, sd):
    from functools import partial
    m = ref.SupervisedMAE(img_size=cfg["img_size"], patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"],
                          depth=cfg["depth"], num_heads=cfg["num_heads"], decoder_embed_dim=cfg["decoder_embed_dim"],
                          decoder_depth=cfg["decoder_depth"], decoder_num_heads=cfg["decoder_num_heads"],
                          mlp_ratio=cfg["mlp_ratio"], norm_layer=partial(nn.LayerNorm, eps=cfg["eps"]))
    # the synthetic spec must reproduce the reference's key set, order and shapes exactly
    ref_sd = m.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "state_dict key order differs from the reference"
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    # the sin-cos tables of the oracle must equal the reference's bit for bit
    assert torch.equal(ref_sd["pos_embed"], sd["pos_embed"]) and torch.equal(ref_sd["decoder_pos_embed"], sd["decoder_pos_embed"])
    m.load_state_dict(sd, strict=True)
    return m


NOCT_SMALL = dict(img_size=384, patch_size=16, embed_dim=256, depth=2, num_heads=4, decoder_embed_dim=512, decoder_depth=2,
                  decoder_num_heads=16, mlp_ratio=4, eps=1e-6)


def gen_noct(ref, out_dir):
    """MAE pre-training model (BASELINE config 5 path): loss / pred / mask and every parameter-gradient norm."""
    from functools import partial
    from oracle import noct_oracle as NO
    cfg = NOCT_SMALL
    sd = NO.make_state_dict(cfg, seed=2)
    out = {}
    for norm_pix in (False, True):
        m = ref.noct.MaskedAutoencoderViTNoCT(img_size=384, patch_size=16, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                                              num_heads=cfg["num_heads"], decoder_embed_dim=512, decoder_depth=cfg["decoder_depth"],
                                              decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                                              norm_pix_loss=norm_pix)
        assert list(m.state_dict().keys()) == list(sd.keys())
        m.load_state_dict(sd, strict=True)
        m.train()
        imgs, _ = synth.make_inputs(2, seed=91)
        torch.manual_seed(123)                     # the reference draws its masking noise from the global RNG
        loss, pred, mask = m(imgs, mask_ratio=0.5)
        loss.backward()
        tag = "np1" if norm_pix else "np0"
        out[f"{tag}/loss"] = loss.detach().numpy()
        out[f"{tag}/mask"] = mask.numpy()
        out[f"{tag}/pred_head"] = pred.detach()[:, :4, :64].numpy()
        out[f"{tag}/pred_rowsum"] = pred.detach().sum(-1).numpy()
        for name, p in m.named_parameters():
            if p.grad is not None:
                out[f"{tag}/g/{name}/norm"] = p.grad.norm().numpy()
                out[f"{tag}/g/{name}/head"] = p.grad.flatten()[:16].numpy()
        print(f"noct norm_pix={norm_pix}: loss={loss.item():.6f} kept={int((mask == 0).sum())}")
    np.savez_compressed(os.path.join(out_dir, "noct_small.npz"), **out)


def gen_curve(ref, out_dir):
    """The reference model stepped the way FSC_finetune_cross.py:234-315 steps it (fp32 here: no autocast on CPU):
    per-step loss, and where a few parameters end up."""
    cfg = synth.CONFIGS["small"]
    sd = synth.make_state_dict(cfg, seed=1)
    m = build_ref_model(ref, cfg, sd).train()
    opt = torch.optim.AdamW(synth.weight_decay_groups(m.named_parameters(), synth.CURVE["weight_decay"]), lr=synth.CURVE["lr"], betas=synth.CURVE["betas"])
    batches = synth.curve_batches()
    losses, counts = [], []
    opt.zero_grad()
    for it in range(synth.CURVE["steps"]):
        imgs, boxes, gt, mask = batches[it % 2]
        shot = synth.CURVE["shots"][it]
        out = m(imgs, boxes[:, :shot] if shot else torch.empty(synth.CURVE["batch"], 0), shot)
        loss = O.finetune_loss(out, gt, mask)
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
        counts.append((out.detach().sum((1, 2)) / 60).numpy())
        print(f"curve step {it} shot={shot}: loss={loss.item():.6f}")
    res = {"loss": np.array(losses, np.float64), "count": np.stack(counts)}
    for name, p in m.named_parameters():
        if p.requires_grad:
            res[f"final/{name}/delta_norm"] = (p.detach() - sd[name]).norm().numpy()
    np.savez_compressed(os.path.join(out_dir, "small_curve.npz"), **res)


def gen_base_b8(ref, out_dir):
    """The benchmarked configuration itself (BASELINE configs[1]): base model, batch 8, 3 exemplars — density map,
    encoder latent and every decoder gradient of the fine-tune loss from the UNMODIFIED reference (fp32, CPU)."""
    cfg = synth.CONFIGS["base"]
    sd = synth.make_state_dict(cfg, seed=0)
    m = build_ref_model(ref, cfg, sd).train()
    imgs, boxes = synth.make_inputs(8, seed=1234)
    gt, mask = synth.make_targets(8, seed=4321)
    with torch.no_grad():
        latent = m.forward_encoder(imgs)
    out = m(imgs, boxes, 3)
    loss = O.finetune_loss(out, gt, mask)
    loss.backward()
    res = {"loss": loss.detach().numpy(), "out_pool8": pool8(out.detach()).numpy(), "out_sum": out.detach().sum((1, 2)).numpy(),
           "out_rows": out.detach()[:, [0, 100, 383]].numpy(),
           "latent_sub": latent[:, ::8, ::8].numpy(), "latent_rownorm": latent.norm(dim=-1).numpy(),
           "latent_norm": latent.norm().numpy()}
    tot = 0.0
    for name, p in m.named_parameters():
        if p.grad is None:
            continue
        gflat = p.grad.flatten()
        res[f"g/{name}/norm"] = gflat.norm().numpy()
        res[f"g/{name}/sum"] = gflat.sum().numpy()
        res[f"g/{name}/head"] = gflat[:16].numpy()
        tot += float(gflat.double().norm() ** 2)
    res["g_total_norm"] = np.array(tot ** 0.5)
    np.savez_compressed(os.path.join(out_dir, "base_b8.npz"), **res)
    print(f"base_b8: loss={loss.item():.6f} counts={(out.detach().sum((1, 2)) / 60).tolist()} |g|={tot ** 0.5:.6e}")


NOCT_BASE = dict(img_size=384, patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=8,
                 decoder_num_heads=16, mlp_ratio=4, eps=1e-6)


def gen_noct_base(ref, out_dir):
    """MAE pre-training at the benchmarked geometry (models_mae_noct.mae_vit_base_patch16, BASELINE configs[4]), batch 4,
    mask_ratio 0.5: loss, prediction summary and every parameter-gradient norm from the UNMODIFIED reference."""
    from functools import partial
    from oracle import noct_oracle as NO
    cfg = NOCT_BASE
    sd = NO.make_state_dict(cfg, seed=5)
    m = ref.noct.MaskedAutoencoderViTNoCT(img_size=384, patch_size=16, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                                          num_heads=cfg["num_heads"], decoder_embed_dim=512, decoder_depth=cfg["decoder_depth"],
                                          decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                                          norm_pix_loss=False)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    m.train()
    imgs, _ = synth.make_inputs(4, seed=92)
    torch.manual_seed(321)
    loss, pred, mask = m(imgs, mask_ratio=0.5)
    loss.backward()
    out = {"loss": loss.detach().numpy(), "mask": mask.numpy(), "pred_rowsum": pred.detach().sum(-1).numpy(),
           "pred_head": pred.detach()[:, :4, :64].numpy()}
    tot = 0.0
    for name, p in m.named_parameters():
        if p.grad is not None:
            out[f"g/{name}/norm"] = p.grad.norm().numpy()
            out[f"g/{name}/head"] = p.grad.flatten()[:16].numpy()
            tot += float(p.grad.double().norm() ** 2)
    out["g_total_norm"] = np.array(tot ** 0.5)
    np.savez_compressed(os.path.join(out_dir, "noct_base.npz"), **out)
    print(f"noct_base: loss={loss.item():.6f} |g|={tot ** 0.5:.6e}")


def pool8(x):
    return torch.nn.functional.avg_pool2d(x[:, None], 8)[:, 0]


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    ref = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--curve-only" in sys.argv:
        return gen_curve(ref, out_dir)
    if "--base-b8-only" in sys.argv:
        return gen_base_b8(ref, out_dir)
    if "--noct-base-only" in sys.argv:
        return gen_noct_base(ref, out_dir)

    # ---- C1: base model, one image, 3 exemplars, eval, fp32 (demo.py path) ----
    cfg = synth.CONFIGS["base"]
    sd = synth.make_state_dict(cfg, seed=0)
    m = build_ref_model(ref, cfg, sd).eval()
    imgs, boxes = synth.make_inputs(1, seed=1234)
    with torch.no_grad():
        latent = m.forward_encoder(imgs)
        out = m(imgs, boxes, 3)
        out0 = m(imgs, torch.empty(1, 0), 0)
    np.savez_compressed(os.path.join(out_dir, "base_c1.npz"), out=out.numpy(), out_zero_pool8=pool8(out0).numpy(),
                        out_zero_sum=out0.sum().numpy(), latent_head=latent[0, :8, :32].numpy(),
                        latent_mean=latent.mean().numpy(), latent_std=latent.std().numpy(),
                        latent_rowsum=latent[0].sum(-1).numpy())
    print("base_c1: count=%.4f zero-shot count=%.4f" % (out.sum().item() / 60, out0.sum().item() / 60))

    # ---- small model: every shot count, B=2, plus decoder gradients of the fine-tune loss ----
    cfg = synth.CONFIGS["small"]
    sd = synth.make_state_dict(cfg, seed=1)
    m = build_ref_model(ref, cfg, sd).train()   # train(): no dropout / BN in the model, same arithmetic as eval
    imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
    gt, mask = synth.make_targets(2, seed=78)
    fwd, grads = {}, {}
    for shot in (0, 1, 2, 3, 5):
        m.zero_grad(set_to_none=True)
        bx = boxes if shot > 0 else torch.empty(2, 0)
        out = m(imgs, bx, shot)
        loss = O.finetune_loss(out, gt, mask)
        loss.backward()
        fwd[f"out_pool8_s{shot}"] = pool8(out.detach()).numpy()
        fwd[f"out_rows_s{shot}"] = out.detach()[:, [0, 100, 383]].numpy()
        fwd[f"sum_s{shot}"] = out.detach().sum((1, 2)).numpy()
        fwd[f"loss_s{shot}"] = loss.detach().numpy()
        if shot in (0, 3):
            for name, p in m.named_parameters():
                if p.grad is None:
                    continue
                gflat = p.grad.flatten()
                grads[f"s{shot}/{name}/norm"] = gflat.norm().numpy()
                grads[f"s{shot}/{name}/sum"] = gflat.sum().numpy()
                grads[f"s{shot}/{name}/head"] = gflat[:16].numpy()
        print(f"small shot={shot}: loss={loss.item():.6f} sums={out.detach().sum((1,2)).tolist()}")
    np.savez_compressed(os.path.join(out_dir, "small_fwd.npz"), **fwd)
    np.savez_compressed(os.path.join(out_dir, "small_grads.npz"), **grads)
    gen_noct(ref, out_dir)
    gen_curve(ref, out_dir)
    gen_base_b8(ref, out_dir)
    gen_noct_base(ref, out_dir)
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
