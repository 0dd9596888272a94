"""timm==0.4.9 `vision_transformer.PatchEmbed` / `Block` equivalents (the reference imports them at
models_mae_cross.py:13; timm is not vendored there).  Same attribute names, so the state_dict keys
`patch_embed.proj.*`, `blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*` line up with
the published FSC147.pth checkpoint."""
import torch
import torch.nn as nn

from . import ops
from .engine import F16, F32, _contig32, engine
from .models_crossvit import Attention, Mlp, to_2tuple


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = tuple(to_2tuple(img_size))
        patch_size = tuple(to_2tuple(patch_size))
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    @torch.no_grad()
    def forward(self, x):
        """[B, C, H, W] -> [B, num_patches, embed_dim]: proj(x).flatten(2).transpose(1, 2) (timm 0.4.9 PatchEmbed.forward; called at
        models_mae_cross.py:138, models_mae_noct.py:139) as a patch gather + one tcgen05 GEMM."""
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        P = self.patch_size[0]
        M, D = B * self.num_patches, self.proj.weight.shape[0]
        patches = torch.empty(M, C * P * P, dtype=F16, device=x.device)
        ops.patchify(x, patches, P)
        y = torch.empty(M, D, dtype=F32, device=x.device)
        ops.linear(patches, engine().wc.w16(self.proj.weight), y, bias=_contig32(self.proj.bias))
        y = y.view(B, self.num_patches, D)
        return y if x.dtype == F32 else y.to(x.dtype)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if drop_path:
            raise NotImplementedError("countr_b200: drop_path > 0 is not on the reference's hot path")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    @torch.no_grad()
    def forward(self, x):
        """x = x + attn(norm1(x)); x = x + mlp(norm2(x)) (timm 0.4.9 Block.forward; called at models_mae_cross.py:145,
        models_mae_noct.py:149,169) on an fp32 residual stream: the schedule of Engine.encoder_forward for one block."""
        B, L, D = x.shape
        M = B * L
        dev = x.device
        wc = engine().wc
        H = self.attn.num_heads
        r = x.detach().to(F32).contiguous().view(M, D).clone()
        h = torch.empty(M, D, dtype=F16, device=dev)
        qkv = torch.empty(M, 3 * D, dtype=F16, device=dev)
        att = torch.empty(M, D, dtype=F16, device=dev)
        u = torch.empty(M, self.mlp.fc1.weight.shape[0], dtype=F16, device=dev)
        qb = None if self.attn.qkv.bias is None else _contig32(self.attn.qkv.bias)
        ops.layernorm_fwd(r, _contig32(self.norm1.weight), _contig32(self.norm1.bias), self.norm1.eps, y16=h)
        ops.linear(h, wc.w16(self.attn.qkv.weight), qkv, bias=qb)
        ops.attention_fwd(qkv, att, B, L, H, D // H, self.attn.scale)
        ops.linear(att, wc.w16(self.attn.proj.weight), r, bias=_contig32(self.attn.proj.bias), residual=r)
        ops.layernorm_fwd(r, _contig32(self.norm2.weight), _contig32(self.norm2.bias), self.norm2.eps, y16=h)
        ops.linear(h, wc.w16(self.mlp.fc1.weight), u, bias=_contig32(self.mlp.fc1.bias), act=1)
        ops.linear(u, wc.w16(self.mlp.fc2.weight), r, bias=_contig32(self.mlp.fc2.bias), residual=r)
        r = r.view(B, L, D)
        return r if x.dtype == F32 else r.to(x.dtype)
