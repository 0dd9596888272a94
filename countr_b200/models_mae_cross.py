"""Drop-in for the reference's models_mae_cross.py (SupervisedMAE + factories, :18-253).

Same constructor signature, factory names, `forward(imgs, boxes, shot_num)` contract and
state_dict keys, so FSC_finetune_cross.py / FSC_test_cross(*).py / demo.py import it unchanged.
All device work is done by the sm_100a kernels behind include/countr_b200.h:
  * the frozen encoder runs outside autograd (the reference wraps it in torch.no_grad(), :204-205)
  * the decoder (exemplar CNN + FIM + density head) is ONE autograd node whose backward launches
    the hand-written backward kernels and hands the parameter gradients to autograd, so
    GradScaler / DDP(find_unused_parameters=True) / AdamW in the reference scripts keep working.
"""
import math
from functools import partial

import torch
import torch.nn as nn

from . import _lib
from .engine import F16, F32, engine
from .models_crossvit import CrossAttentionBlock
from .pos_embed import get_2d_sincos_pos_embed
from .vit import Block, PatchEmbed


class _DecoderFn(torch.autograd.Function):
    """forward_decoder as a single autograd node over the decoder parameters."""

    @staticmethod
    def forward(ctx, model, lat16, boxes, shot_num, B, out_dtype, names, pre, *params):
        save = {}
        out = engine().decoder_forward(model, lat16, boxes, shot_num, B, out_dtype, save=save, pre=pre)
        ctx.model, ctx.saved, ctx.names, ctx.boxes = model, save, names, boxes
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from .backward import decoder_backward
        grads = decoder_backward(engine(), ctx.model, ctx.saved, ctx.boxes, grad_out)
        ctx.saved = None
        return (None, None, None, None, None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


class SupervisedMAE(nn.Module):
    def __init__(self, img_size=384, patch_size=16, in_chans=3,
                 embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=2, decoder_num_heads=16,
                 mlp_ratio=4., norm_layer=nn.LayerNorm, norm_pix_loss=False):
        super().__init__()
        # encoder (models_mae_cross.py:25-35)
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True, qk_scale=None, norm_layer=norm_layer)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        # decoder (:39-100)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches, decoder_embed_dim), requires_grad=False)
        self.shot_token = nn.Parameter(torch.zeros(512))

        def proj(cin, cout, last=False):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.InstanceNorm2d(cout),
                                 nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d((1, 1)) if last else nn.MaxPool2d(2))

        self.decoder_proj1 = proj(3, 64)
        self.decoder_proj2 = proj(64, 128)
        self.decoder_proj3 = proj(128, 256)
        self.decoder_proj4 = proj(256, decoder_embed_dim, last=True)
        self.decoder_blocks = nn.ModuleList([
            CrossAttentionBlock(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=True, qk_scale=None,
                                norm_layer=norm_layer)
            for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)

        def head(cin, final=False):
            layers = [nn.Conv2d(cin, 256, kernel_size=3, stride=1, padding=1), nn.GroupNorm(8, 256), nn.ReLU(inplace=True)]
            if final:
                layers.append(nn.Conv2d(256, 1, kernel_size=1, stride=1))
            return nn.Sequential(*layers)

        self.decode_head0 = head(decoder_embed_dim)
        self.decode_head1 = head(256)
        self.decode_head2 = head(256)
        self.decode_head3 = head(256, final=True)
        self.norm_pix_loss = norm_pix_loss
        self.initialize_weights()

    # -- init: same distributions as models_mae_cross.py:108-134 (host side, init time only)
    def initialize_weights(self):
        grid = int(self.patch_embed.num_patches ** .5)
        pos_embed = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid, cls_token=False)
        self.pos_embed.data.copy_(torch.from_numpy(pos_embed).float().unsqueeze(0))
        decoder_pos_embed = get_2d_sincos_pos_embed(self.decoder_pos_embed.shape[-1], grid, cls_token=False)
        self.decoder_pos_embed.data.copy_(torch.from_numpy(decoder_pos_embed).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        torch.nn.init.normal_(self.shot_token, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # -- hot path
    def _check(self, t):
        if not t.is_cuda:
            raise _lib.CountrError("countr_b200 runs on a B200 (sm_100a) only: got a CPU tensor and there is no CPU fallback")
        _lib.require_device()

    def _encode(self, imgs):
        self._check(imgs)
        train = self._needs_grad()     # the backward keeps the 16-bit latent: it must not live in the shared workspace
        with torch.no_grad():
            return engine().encoder_forward(self, imgs, keep=train)

    def forward_encoder(self, x):
        """[N,3,H,W] -> [N, L, D] (patch embed + pos + blocks + norm, models_mae_cross.py:136-148)."""
        lat32, _ = self._encode(x)
        return lat32 if x.dtype == F32 else lat32.to(x.dtype)

    def _decoder_params(self, shot_num=None):
        """Trainable decoder parameters in `named_parameters()` order; shot_num=None -> all of them (the gradient arena
        layout), otherwise only the ones a step with `shot_num` exemplars reaches."""
        names, params = [], []
        for n, p in self.named_parameters():
            if not p.requires_grad or n.startswith(("patch_embed.", "blocks.", "norm.")):
                continue
            if shot_num is not None and ((shot_num > 0 and n == "shot_token") or (shot_num == 0 and n.startswith("decoder_proj"))):
                continue   # unused on this path: no gradient, like the reference (DDP find_unused_parameters)
            names.append(n)
            params.append(p)
        return names, params

    def _needs_grad(self):
        if not torch.is_grad_enabled():
            return False
        return any(p.requires_grad for n, p in self.named_parameters() if not n.startswith(("patch_embed.", "blocks.", "norm.")))

    def _decode(self, lat16, y_, shot_num, B, out_dtype, pre=None):
        if self._needs_grad():
            names, params = self._decoder_params(shot_num)
            return _DecoderFn.apply(self, lat16, y_, shot_num, B, out_dtype, tuple(names), pre, *params)
        with torch.no_grad():
            return engine().decoder_forward(self, lat16, y_, shot_num, B, out_dtype, pre=pre)

    def forward_decoder(self, x, y_, shot_num=3):
        """x: encoder output [N, L, D]; y_: boxes [N, K, 3, 64, 64] or an empty tensor (:150-199)."""
        self._check(x)
        from . import ops
        B, L, D = x.shape
        x32 = x.detach().to(F32).contiguous().view(B * L, D)
        lat16 = torch.empty(B * L, D, dtype=F16, device=x.device)
        ops.cast16(x32, lat16)
        return self._decode(lat16, y_, shot_num, B, x.dtype)

    def forward(self, imgs, boxes, shot_num):
        """-> density map [N, H, W]; count = map.sum() / 60 (models_mae_cross.py:201-207)."""
        self._check(imgs)
        eng = engine()
        pre = None
        # every stale 16-bit decoder weight copy is refreshed in one launch — on the side stream, ahead of the exemplar CNN
        # (the frozen encoder does not read them, so the ~50 us stay off the critical path)
        refreshed = eng.refresh_decoder_weights(self, shot_num, self._needs_grad(), imgs.device)
        if shot_num > 0 and eng.overlap_exemplar:
            assert boxes.dim() == 5 and boxes.shape[1] >= shot_num, "boxes must be [N, K>=shot_num, 3, 64, 64]"
            train = self._needs_grad()
            with torch.no_grad():
                pre = eng.exemplar_async(self, boxes, shot_num, train=train)   # overlaps the encoder
        _, lat16 = self._encode(imgs)
        if refreshed is not None:
            torch.cuda.current_stream().wait_event(refreshed)
        out_dtype = imgs.dtype if imgs.dtype in (F32, F16, torch.bfloat16) else F32
        return self._decode(lat16, boxes, shot_num, imgs.shape[0], out_dtype, pre)


def mae_vit_base_patch16_dec512d8b(**kwargs):
    return SupervisedMAE(patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=2,
                         decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_large_patch16_dec512d8b(**kwargs):
    return SupervisedMAE(patch_size=16, embed_dim=1024, depth=24, num_heads=16, decoder_embed_dim=512, decoder_depth=2,
                         decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_huge_patch14_dec512d8b(**kwargs):
    return SupervisedMAE(patch_size=14, embed_dim=1280, depth=32, num_heads=16, decoder_embed_dim=512, decoder_depth=2,
                         decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_base_patch16_fim4(**kwargs):
    return SupervisedMAE(patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=4,
                         decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_base_patch16_fim6(**kwargs):
    return SupervisedMAE(patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=6,
                         decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


# recommended archs (models_mae_cross.py:248-253)
mae_vit_base_patch16 = mae_vit_base_patch16_dec512d8b
mae_vit_base4_patch16 = mae_vit_base_patch16_fim4
mae_vit_base6_patch16 = mae_vit_base_patch16_fim6
mae_vit_large_patch16 = mae_vit_large_patch16_dec512d8b
mae_vit_huge_patch14 = mae_vit_huge_patch14_dec512d8b
