"""CPU ORACLE (test infrastructure only) for the MAE pre-training model, models_mae_noct.py.
Functional fp32 restatement over a state_dict; see oracle/countr_oracle.py for the rules that apply."""
import torch
import torch.nn.functional as F

from . import countr_oracle as O


def state_dict_spec(cfg):
    """(key, shape, kind) in the reference's state_dict order (models_mae_noct.py:19-48)."""
    D, Dd, P = cfg["embed_dim"], cfg["decoder_embed_dim"], cfg["patch_size"]
    L = (cfg["img_size"] // P) ** 2
    spec = [("pos_embed", (1, L, D), "pos"), ("mask_token", (1, 1, Dd), "token"), ("decoder_pos_embed", (1, L, Dd), "pos"),
            ("patch_embed.proj.weight", (D, 3, P, P), "w"), ("patch_embed.proj.bias", (D,), "b")]

    def lin(prefix, out_f, in_f):
        spec.append((prefix + ".weight", (out_f, in_f), "w"))
        spec.append((prefix + ".bias", (out_f,), "b"))

    def norm(prefix, dim):
        spec.append((prefix + ".weight", (dim,), "g"))
        spec.append((prefix + ".bias", (dim,), "b"))

    def blocks(prefix, n, dim):
        hid = int(dim * cfg["mlp_ratio"])
        for i in range(n):
            p = f"{prefix}.{i}"
            norm(p + ".norm1", dim); lin(p + ".attn.qkv", 3 * dim, dim); lin(p + ".attn.proj", dim, dim)
            norm(p + ".norm2", dim); lin(p + ".mlp.fc1", hid, dim); lin(p + ".mlp.fc2", dim, hid)

    blocks("blocks", cfg["depth"], D)
    norm("norm", D)
    lin("decoder_embed", Dd, D)
    blocks("decoder_blocks", cfg["decoder_depth"], Dd)
    norm("decoder_norm", Dd)
    lin("decoder_pred", P * P * 3, Dd)
    return spec


def make_state_dict(cfg, seed=0):
    import math
    g = torch.Generator().manual_seed(seed)
    sd = {}
    grid = cfg["img_size"] // cfg["patch_size"]
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "pos":
            sd[key] = O.sincos_2d(shape[-1], grid).unsqueeze(0)
        elif kind == "token":
            sd[key] = torch.randn(shape, generator=g) * 0.02
        elif kind == "w":
            a = math.sqrt(6.0 / (int(math.prod(shape[1:])) + shape[0]))
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "g":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[key] = 0.02 * torch.randn(shape, generator=g)
    return sd


def patchify(imgs, p):
    """models_mae_noct.py:82-94."""
    h = w = imgs.shape[2] // p
    x = imgs.reshape(imgs.shape[0], 3, h, p, w, p)
    return torch.einsum("nchpwq->nhwpqc", x).reshape(imgs.shape[0], h * w, p * p * 3)


def forward(sd, cfg, imgs, mask_ratio, noise, norm_pix_loss=False):
    """MaskedAutoencoderViTNoCT.forward (models_mae_noct.py:200-204) with the masking noise passed in
    (the reference draws it with torch.rand(N, L), :118)."""
    p = cfg["patch_size"]
    x = F.conv2d(imgs, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p).flatten(2).transpose(1, 2)
    x = x + sd["pos_embed"]                                                                   # :139-142
    N, L, D = x.shape
    len_keep = int(L * (1 - mask_ratio))                                                      # :116
    ids_shuffle = torch.argsort(noise, dim=1)                                                 # :121-122
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    ids_keep = ids_shuffle[:, :len_keep]
    x = torch.gather(x, 1, ids_keep.unsqueeze(-1).repeat(1, 1, D))                            # :125-126
    mask = torch.ones(N, L)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, 1, ids_restore)                                                 # :129-133
    for i in range(cfg["depth"]):
        x = O.vit_block(x, sd, f"blocks.{i}", cfg["num_heads"], cfg["eps"])                   # :148-149
    x = O.layer_norm(x, sd, "norm", cfg["eps"])
    x = O.linear(x, sd, "decoder_embed")                                                      # :156
    mask_tokens = sd["mask_token"].repeat(N, L - x.shape[1], 1)                               # :159
    x_ = torch.cat([x, mask_tokens], dim=1)
    x = torch.gather(x_, 1, ids_restore.unsqueeze(-1).repeat(1, 1, x.shape[2]))               # :161
    x = x + sd["decoder_pos_embed"]                                                           # :165
    for i in range(cfg["decoder_depth"]):
        x = O.vit_block(x, sd, f"decoder_blocks.{i}", cfg["decoder_num_heads"], cfg["eps"])   # :168-169
    x = O.layer_norm(x, sd, "decoder_norm", cfg["eps"])
    pred = O.linear(x, sd, "decoder_pred")                                                    # :173
    target = patchify(imgs, p)                                                                # :183
    if norm_pix_loss:
        mean = target.mean(dim=-1, keepdim=True)
        var = target.var(dim=-1, keepdim=True)
        target = (target - mean) / (var + 1.e-6) ** .5                                        # :184-187
    loss = ((pred - target) ** 2).mean(dim=-1)                                                # :189-190
    loss = loss.sum() / (N * L)                                                               # :193-195 (mask_s = ones)
    return loss, pred, mask
