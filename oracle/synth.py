"""Deterministic synthetic weights and inputs shared by the golden generator, the oracle tests
and the GPU parity tests (test infrastructure; see oracle/countr_oracle.py header).

The real FSC147 checkpoint / dataset are not available offline, so parity is checked on seeded
random weights.  Unlike the reference's init (zero biases, unit LN gains — models_mae_cross.py:
126-134) every bias / gain here is non-trivial so that a dropped bias or a swapped gamma/beta
cannot hide.
"""
import math

import numpy as np

import torch

from . import countr_oracle as O

CONFIGS = {
    # the reference's mae_vit_base_patch16 (models_mae_cross.py:210-215)
    "base": dict(img_size=384, patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512,
                 decoder_depth=2, decoder_num_heads=16, mlp_ratio=4, eps=1e-6),
    # a shallow, narrow encoder with the real decoder: quick enough for many CPU parity cases
    # ViT-L/16 geometry (embed 1024, 16 heads of 64; models_mae_cross.py:218-223) with a 4-block FIM (:232-237), shallow
    "large_fim4": dict(img_size=384, patch_size=16, embed_dim=1024, depth=2, num_heads=16, decoder_embed_dim=512,
                       decoder_depth=4, decoder_num_heads=16, mlp_ratio=4, eps=1e-6),
    # ViT-H/14 geometry (models_mae_cross.py:226-231): 14-px patches -> 27 x 27 = 729 tokens, 1280 / 16 = 80 channels per head,
    # 432 x 432 density map; shallow encoder
    "huge_d2": dict(img_size=384, patch_size=14, embed_dim=1280, depth=2, num_heads=16, decoder_embed_dim=512,
                    decoder_depth=2, decoder_num_heads=16, mlp_ratio=4, eps=1e-6),
    "small": dict(img_size=384, patch_size=16, embed_dim=256, depth=2, num_heads=4, decoder_embed_dim=512,
                  decoder_depth=2, decoder_num_heads=16, mlp_ratio=4, eps=1e-6),
}


def state_dict_spec(cfg):
    """(key, shape, kind) in the reference's state_dict order."""
    D, Dd, P = cfg["embed_dim"], cfg["decoder_embed_dim"], cfg["patch_size"]
    L = (cfg["img_size"] // P) ** 2
    Hd = int(D * cfg["mlp_ratio"])
    Hdd = int(Dd * cfg["mlp_ratio"])
    spec = [("pos_embed", (1, L, D), "pos"), ("decoder_pos_embed", (1, L, Dd), "pos"), ("shot_token", (Dd,), "token"),
            ("patch_embed.proj.weight", (D, 3, P, P), "w"), ("patch_embed.proj.bias", (D,), "b")]

    def lin(prefix, out_f, in_f):
        spec.append((prefix + ".weight", (out_f, in_f), "w"))
        spec.append((prefix + ".bias", (out_f,), "b"))

    def norm(prefix, dim):
        spec.append((prefix + ".weight", (dim,), "g"))
        spec.append((prefix + ".bias", (dim,), "b"))

    for i in range(cfg["depth"]):
        p = f"blocks.{i}"
        norm(p + ".norm1", D); lin(p + ".attn.qkv", 3 * D, D); lin(p + ".attn.proj", D, D)
        norm(p + ".norm2", D); lin(p + ".mlp.fc1", Hd, D); lin(p + ".mlp.fc2", D, Hd)
    norm("norm", D)
    lin("decoder_embed", Dd, D)
    for i, (ci, co) in enumerate([(3, 64), (64, 128), (128, 256), (256, Dd)], 1):
        spec.append((f"decoder_proj{i}.0.weight", (co, ci, 3, 3), "w"))
        spec.append((f"decoder_proj{i}.0.bias", (co,), "b"))
    for j in range(cfg["decoder_depth"]):
        p = f"decoder_blocks.{j}"
        norm(p + ".norm0", Dd); lin(p + ".selfattn.qkv", 3 * Dd, Dd); lin(p + ".selfattn.proj", Dd, Dd)
        norm(p + ".norm1", Dd)
        for n in ("wq", "wk", "wv", "proj"):
            lin(p + ".attn." + n, Dd, Dd)
        norm(p + ".norm2", Dd); lin(p + ".mlp.fc1", Hdd, Dd); lin(p + ".mlp.fc2", Dd, Hdd)
    norm("decoder_norm", Dd)
    for i, ci in enumerate([Dd, 256, 256, 256]):
        spec.append((f"decode_head{i}.0.weight", (256, ci, 3, 3), "w"))
        spec.append((f"decode_head{i}.0.bias", (256,), "b"))
        norm(f"decode_head{i}.1", 256)
    spec.append(("decode_head3.3.weight", (1, 256, 1, 1), "w"))
    spec.append(("decode_head3.3.bias", (1,), "b"))
    return spec


def make_state_dict(cfg, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    grid = cfg["img_size"] // cfg["patch_size"]
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "pos":
            sd[key] = O.sincos_2d(shape[-1], grid).unsqueeze(0)
        elif kind == "token":
            sd[key] = torch.randn(shape, generator=g) * 0.02
        elif kind == "w":
            fan_out = shape[0]
            fan_in = int(math.prod(shape[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "g":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[key] = 0.02 * torch.randn(shape, generator=g)
    return sd


def make_inputs(B, seed=1234, shots=3, img_size=384):
    """BASELINE.md §3 inputs: uniform [0,1] images (no mean/std normalisation, util/FSC147.py:367-369)
    and 64x64 exemplar crops."""
    g = torch.Generator().manual_seed(seed)
    imgs = torch.rand(B, 3, img_size, img_size, generator=g)
    boxes = torch.rand(B, shots, 3, 64, 64, generator=g)
    return imgs, boxes


def make_targets(B, seed=4321, img_size=384):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, img_size, img_size, generator=g) * 0.5
    mask = (torch.rand(img_size, img_size, generator=g) < 0.8).float()   # np.random.binomial(1, .8) stand-in
    return gt, mask


# ---- the fine-tune loss-curve case (tests/golden/small_curve.npz) ----
CURVE = dict(steps=16, lr=2e-5, weight_decay=0.05, betas=(0.9, 0.95), batch=2, shots=[3, 3, 0, 2, 3, 1, 3, 3, 0, 3, 3, 2, 3, 3, 3, 3])


def curve_batches():
    """Two fixed batches visited alternately, so the loss moves a long way in few steps."""
    out = []
    for i in range(2):
        imgs, boxes = make_inputs(CURVE["batch"], seed=500 + i)
        gt, mask = make_targets(CURVE["batch"], seed=600 + i)
        out.append((imgs, boxes, gt, mask))
    return out


def weight_decay_groups(named_params, weight_decay):
    """timm.optim.optim_factory.add_weight_decay as FSC_finetune_cross.py:234 calls it: trainable 1-D tensors and
    biases get no decay."""
    decay, no_decay = [], []
    for n, p in named_params:
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim == 1 or n.endswith(".bias")) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def smooth_image(h, w, seed):
    """Smooth uint8 RGB test pattern (sums of sinusoids): deterministic, so large test images need not be stored in fixtures."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        for _ in range(4):
            fy, fx, ph = rng.uniform(0.01, 0.12), rng.uniform(0.01, 0.12), rng.uniform(0, 6.28)
            img[..., c] += rng.uniform(0.3, 1.0) * np.sin(fy * yy + fx * xx + ph)
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)
