"""CPU ORACLE — test infrastructure only.  NOT part of the product path.

A functional, fp32, CPU restatement of the reference hot path
(`SupervisedMAE.forward` = ViT encoder -> exemplar CNN -> FIM -> density head) written against
plain tensors in a state_dict, each function citing the reference lines it follows.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference leg may
import this file; nothing under `countr_b200/` does.

Third-party arithmetic: `timm==0.4.9` (requirements.txt:5) is not vendored in /root/reference
and is not installed here.  `timm.models.vision_transformer.{PatchEmbed,Block}` are restated
below from their published 0.4.9 definition; the reference's own `models_crossvit.Attention` /
`Mlp` (models_crossvit.py:46-94) are verbatim copies of timm's and anchor the restatement.

Pinning: the reference has no tests or golden vectors, so this oracle is pinned
against the reference ITSELF: `scripts/gen_golden.py` imports /root/reference/models_mae_cross.py
(with a timm shim) in the build container, runs it on the seeded inputs of `oracle/synth.py`
and commits the outputs under tests/golden/; tests/test_oracle.py checks this file against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# util/pos_embed.py:20-67 — fixed 2-D sin-cos table, float64 numpy -> float32; "w goes first"
# ---------------------------------------------------------------------------------------------
def sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float64)        # pos_embed.py:56 (np.float == float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)         # :60-61
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)  # :63-66


def sincos_2d(embed_dim, grid_size):
    grid_h = np.arange(grid_size, dtype=np.float32)
    grid_w = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape(2, 1, grid_size, grid_size)   # :26-31
    emb_h = sincos_1d(embed_dim // 2, grid[0])                 # :43-44
    emb_w = sincos_1d(embed_dim // 2, grid[1])
    return torch.from_numpy(np.concatenate([emb_h, emb_w], axis=1)).float()                    # :46, cast as models_mae_cross.py:112


# ---------------------------------------------------------------------------------------------
# transformer pieces
# ---------------------------------------------------------------------------------------------
def layer_norm(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def attention(x, sd, prefix, num_heads):
    """models_crossvit.py:82-94 (== timm 0.4.9 Attention.forward); dropout p=0."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = linear(x, sd, prefix + ".qkv").reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)   # :84
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * hd ** -0.5                                                 # :87 (scale :75)
    attn = attn.softmax(dim=-1)                                                                   # :88
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)                                               # :91
    return linear(x, sd, prefix + ".proj")                                                        # :92


def mlp(x, sd, prefix):
    """models_crossvit.py:61-67: fc1 -> exact (erf) GELU -> fc2."""
    return linear(F.gelu(linear(x, sd, prefix + ".fc1")), sd, prefix + ".fc2")


def vit_block(x, sd, prefix, num_heads, eps):
    """timm 0.4.9 Block.forward: x += attn(norm1(x)); x += mlp(norm2(x)) (drop_path = identity)."""
    x = x + attention(layer_norm(x, sd, prefix + ".norm1", eps), sd, prefix + ".attn", num_heads)
    x = x + mlp(layer_norm(x, sd, prefix + ".norm2", eps), sd, prefix + ".mlp")
    return x


def cross_attention(x, y, sd, prefix, num_heads):
    """models_crossvit.py:111-128."""
    B, Nx, C = x.shape
    Ny = y.shape[1]
    hd = C // num_heads
    q = linear(x, sd, prefix + ".wq").reshape(B, Nx, num_heads, hd).permute(0, 2, 1, 3)   # :115
    k = linear(y, sd, prefix + ".wk").reshape(B, Ny, num_heads, hd).permute(0, 2, 1, 3)   # :117
    v = linear(y, sd, prefix + ".wv").reshape(B, Ny, num_heads, hd).permute(0, 2, 1, 3)   # :119
    attn = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)                       # :121-122
    x = (attn @ v).transpose(1, 2).reshape(B, Nx, C)                                      # :125
    return linear(x, sd, prefix + ".proj")                                                # :126


def fim_block(x, y, sd, prefix, num_heads, eps):
    """models_crossvit.py:152-156 — y is NOT normalised."""
    x = x + attention(layer_norm(x, sd, prefix + ".norm0", eps), sd, prefix + ".selfattn", num_heads)
    x = x + cross_attention(layer_norm(x, sd, prefix + ".norm1", eps), y, sd, prefix + ".attn", num_heads)
    x = x + mlp(layer_norm(x, sd, prefix + ".norm2", eps), sd, prefix + ".mlp")
    return x


# ---------------------------------------------------------------------------------------------
# SupervisedMAE
# ---------------------------------------------------------------------------------------------
def forward_encoder(sd, cfg, imgs):
    """models_mae_cross.py:136-148 (+ timm PatchEmbed: Conv2d(k=s=patch) -> flatten(2).transpose(1,2))."""
    p = cfg["patch_size"]
    x = F.conv2d(imgs, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p)
    x = x.flatten(2).transpose(1, 2)
    x = x + sd["pos_embed"]                                                               # :141
    for i in range(cfg["depth"]):
        x = vit_block(x, sd, f"blocks.{i}", cfg["num_heads"], cfg["eps"])                 # :144-145
    return layer_norm(x, sd, "norm", cfg["eps"])                                          # :146


class _Round16STE(torch.autograd.Function):
    """fp16 storage rounding with a straight-through gradient.  Test aid only: lets the oracle mimic
    the 16-bit activation storage of the reference's fp16-autocast training (FSC_finetune_cross.py:286)
    at chosen points, so gradients that are routed by max-pool arg-max / ReLU masks can be compared
    against a reference that takes the same routing decisions."""

    @staticmethod
    def forward(ctx, t):
        return t.half().float()

    @staticmethod
    def backward(ctx, g):
        return g


def exemplar_encoder(sd, boxes_one_shot, round16=False):
    """decoder_proj1..4, models_mae_cross.py:47-71: conv3x3 -> InstanceNorm2d (no affine, eps 1e-5)
    -> ReLU -> MaxPool2d(2) x3, last stage AdaptiveAvgPool2d(1)."""
    y = boxes_one_shot
    for i in (1, 2, 3, 4):
        y = F.conv2d(y, sd[f"decoder_proj{i}.0.weight"], sd[f"decoder_proj{i}.0.bias"], padding=1)
        if round16:
            y = _Round16STE.apply(y)
        y = F.relu(F.instance_norm(y, eps=1e-5))
        y = F.max_pool2d(y, 2) if i < 4 else y.mean((2, 3), keepdim=True)
        if round16 and i < 4:
            y = _Round16STE.apply(y)
    return y.squeeze(-1).squeeze(-1)


def forward_decoder(sd, cfg, latent, boxes, shot_num, taps=None, exemplar_round16=False):
    """models_mae_cross.py:150-199."""
    x = linear(latent, sd, "decoder_embed") + sd["decoder_pos_embed"]                     # :152-154
    N = latent.shape[0]
    if shot_num > 0:
        ys = [exemplar_encoder(sd, boxes[:, s], exemplar_round16) for s in range(shot_num)]                 # :157-171 (first shot_num boxes)
        y = torch.stack(ys, 0).transpose(0, 1)                                            # :174,177 -> [N, S, C]
    else:
        y = sd["shot_token"].repeat(N, 1).unsqueeze(0).transpose(0, 1)                    # :176-177 -> [N, 1, C]
    if taps is not None:
        taps["y"] = y
    for j in range(cfg["decoder_depth"]):
        x = fim_block(x, y, sd, f"decoder_blocks.{j}", cfg["decoder_num_heads"], cfg["eps"])   # :180-181
    x = layer_norm(x, sd, "decoder_norm", cfg["eps"])                                     # :182
    if taps is not None:
        taps["fim"] = x
    n, hw, c = x.shape
    h = w = int(math.sqrt(hw))
    x = x.transpose(1, 2).reshape(n, c, h, w)                                             # :185-187
    for i in range(4):                                                                    # :189-196
        x = F.conv2d(x, sd[f"decode_head{i}.0.weight"], sd[f"decode_head{i}.0.bias"], padding=1)
        x = F.relu(F.group_norm(x, 8, sd[f"decode_head{i}.1.weight"], sd[f"decode_head{i}.1.bias"], 1e-5))
        if i == 3:
            x = F.conv2d(x, sd["decode_head3.3.weight"], sd["decode_head3.3.bias"])       # :99
        x = F.interpolate(x, size=x.shape[-1] * 2, mode="bilinear", align_corners=False)
    return x.squeeze(-3)                                                                  # :197


def forward(sd, cfg, imgs, boxes, shot_num, taps=None, exemplar_round16=False):
    """models_mae_cross.py:201-207: encoder under no_grad, then decoder."""
    with torch.no_grad():
        latent = forward_encoder(sd, cfg, imgs)
    if taps is not None:
        taps["latent"] = latent
    return forward_decoder(sd, cfg, latent, boxes, shot_num, taps, exemplar_round16)


def decoder_param_names(sd, shot_num):
    """Parameters that receive a gradient in the fine-tune step (encoder is frozen by no_grad,
    models_mae_cross.py:204; pos tables have requires_grad=False, :30,42)."""
    names = []
    for k in sd:
        if k.startswith(("patch_embed", "blocks.", "norm.")) or k in ("pos_embed", "decoder_pos_embed"):
            continue
        if shot_num > 0 and k == "shot_token":
            continue
        if shot_num == 0 and k.startswith("decoder_proj"):
            continue
        names.append(k)
    return names


def finetune_loss(output, gt_density, mask):
    """FSC_finetune_cross.py:290-295: masked squared error / (384*384), summed, / batch."""
    loss = (output - gt_density) ** 2
    loss = (loss * mask / (output.shape[-1] * output.shape[-2])).sum() / output.shape[0]
    return loss
