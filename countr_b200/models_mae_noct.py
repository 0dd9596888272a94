"""Drop-in for the reference's models_mae_noct.py (MaskedAutoencoderViTNoCT, the MAE pre-training model
without a class token; FSC_pretrain.py:202,263).  Same constructor, factories, method names and
state_dict keys; `forward(imgs, mask_ratio)` returns `(loss, pred, mask)`.

Every dense op (patch embedding, 12 encoder blocks on the kept tokens, 8 decoder blocks on all 576,
predictor, reconstruction loss) and its backward runs in the sm_100a kernels behind
include/countr_b200.h.  The whole model is ONE autograd node: `loss.backward()` (through GradScaler in
the reference script) launches the backward schedule and hands the parameter gradients — views of one
flat fp32 arena — to autograd.

random_masking (models_mae_noct.py:110-135) draws its noise with torch.rand and ranks it with
torch.argsort exactly like the reference (index bookkeeping on [N, 576] values); the token gather /
un-shuffle that moves the activations is done by our kernels.
"""
from functools import partial

import torch
import torch.nn as nn

from . import _lib, ops
from .blocks import vit_block_backward, vit_block_forward
from .dist import build_grad_arena
from .engine import F16, F32, _contig32, engine
from .backward import GradJobs
from .pos_embed import get_2d_sincos_pos_embed
from .vit import Block, PatchEmbed


class _MAEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, imgs, mask_ratio, names, *params):
        tape = {}
        loss, pred, mask = model._forward_impl(imgs, mask_ratio, tape)
        ctx.model, ctx.tape, ctx.names = model, tape, names
        ctx.mark_non_differentiable(pred, mask)
        return loss, pred, mask

    @staticmethod
    def backward(ctx, g_loss, _g_pred, _g_mask):
        grads = ctx.model._backward_impl(ctx.tape, g_loss)
        ctx.tape = None
        return (None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


class MaskedAutoencoderViTNoCT(nn.Module):
    """Masked Autoencoder with VisionTransformer backbone, no class token (models_mae_noct.py:10-204)."""

    def __init__(self, img_size=384, patch_size=16, in_chans=3,
                 embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16,
                 mlp_ratio=4., norm_layer=nn.LayerNorm, norm_pix_loss=False):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True, qk_scale=None, norm_layer=norm_layer)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList([
            Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=True, qk_scale=None, norm_layer=norm_layer)
            for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.decoder_pred = nn.Linear(decoder_embed_dim, patch_size ** 2 * in_chans, bias=True)
        self.norm_pix_loss = norm_pix_loss
        self._noise_override = None        # tests inject the reference's noise here to compare maskings
        self.initialize_weights()

    def initialize_weights(self):
        grid = int(self.patch_embed.num_patches ** .5)
        pos_embed = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid, cls_token=False)
        self.pos_embed.data.copy_(torch.from_numpy(pos_embed).float().unsqueeze(0))
        decoder_pos_embed = get_2d_sincos_pos_embed(self.decoder_pos_embed.shape[-1], grid, cls_token=False)
        self.decoder_pos_embed.data.copy_(torch.from_numpy(decoder_pos_embed).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        torch.nn.init.normal_(self.mask_token, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # -- host-side layout helpers of the reference API (used by FSC_pretrain.py:268 for visualisation only)
    def patchify(self, imgs):
        p = self.patch_embed.patch_size[0]
        assert imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
        h = w = imgs.shape[2] // p
        x = imgs.reshape(shape=(imgs.shape[0], 3, h, p, w, p))
        x = torch.einsum('nchpwq->nhwpqc', x)
        return x.reshape(shape=(imgs.shape[0], h * w, p ** 2 * 3))

    def unpatchify(self, x):
        p = self.patch_embed.patch_size[0]
        h = w = int(x.shape[1] ** .5)
        assert h * w == x.shape[1]
        x = x.reshape(shape=(x.shape[0], h, w, p, p, 3))
        x = torch.einsum('nhwpqc->nchpwq', x)
        return x.reshape(shape=(x.shape[0], 3, h * p, h * p))

    def _masking_indices(self, N, L, mask_ratio, device):
        """Index part of random_masking (models_mae_noct.py:110-135)."""
        len_keep = int(L * (1 - mask_ratio))
        noise = self._noise_override.to(device) if self._noise_override is not None else torch.rand(N, L, device=device)
        ids_shuffle = torch.argsort(noise, dim=1)
        ids_restore = torch.argsort(ids_shuffle, dim=1)
        mask = torch.ones([N, L], device=device)
        mask[:, :len_keep] = 0
        mask = torch.gather(mask, dim=1, index=ids_restore)
        return len_keep, ids_shuffle.contiguous(), ids_restore.contiguous(), mask

    def random_masking(self, x, mask_ratio):
        N, L, D = x.shape
        len_keep, ids_shuffle, ids_restore, mask = self._masking_indices(N, L, mask_ratio, x.device)
        x32 = x.detach().to(F32).contiguous()
        x_masked = torch.empty(N, len_keep, D, dtype=F32, device=x.device)
        ops.gather_rows(x32, ids_shuffle[:, :len_keep].contiguous(), x_masked)
        return x_masked.to(x.dtype), mask, ids_restore

    # -- kernel schedules
    def _trainable(self):
        names, params = [], []
        for n, p in self.named_parameters():
            if p.requires_grad:
                names.append(n)
                params.append(p)
        return names, params

    def _forward_impl(self, imgs, mask_ratio, tape):
        if not imgs.is_cuda:
            raise _lib.CountrError("countr_b200 runs on a B200 (sm_100a) only: got a CPU tensor and there is no CPU fallback")
        _lib.require_device()
        wc = engine().wc
        dev = imgs.device
        # all stale 16-bit weight copies (every trainable tensor after an optimizer step) in one launch
        train_plan = tape is not None
        lin = [self.patch_embed.proj, self.decoder_embed, self.decoder_pred]
        for blk in list(self.blocks) + list(self.decoder_blocks):
            lin += [blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2]
        plan = [(l.weight, "w") for l in lin]
        if train_plan:
            plan += [(l.weight, "wt") for l in lin[1:]]       # the patch embedding needs no dX
        wc.refresh_batch(plan)
        B, C, Himg, Wimg = imgs.shape
        P = self.patch_embed.patch_size[0]
        assert Himg == self.patch_embed.img_size[0] and Wimg == self.patch_embed.img_size[1], \
            f"Input image size ({Himg}*{Wimg}) doesn't match model ({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]})."
        L = (Himg // P) * (Wimg // P)
        D = self.pos_embed.shape[-1]
        Dd = self.decoder_embed.weight.shape[0]
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        # encoder (models_mae_noct.py:137-152)
        patches = e((B * L, C * P * P), F16)
        ops.patchify(imgs, patches, P)
        x_full = e((B * L, D), F32)
        pe = self.patch_embed.proj
        ops.linear(patches, wc.w16(pe.weight), x_full, bias=_contig32(pe.bias), residual=_contig32(self.pos_embed).reshape(L, D), res_mod=L)
        Lk, ids_shuffle, ids_restore, mask = self._masking_indices(B, L, mask_ratio, dev)
        ids_keep = ids_shuffle[:, :Lk].contiguous()
        x = e((B * Lk, D), F32)
        ops.gather_rows(x_full.view(B, L, D), ids_keep, x.view(B, Lk, D))
        enc_tape = []
        for blk in self.blocks:
            x = vit_block_forward(wc, blk, x, B, Lk, enc_tape)
        lat16, mean_e, rstd_e = e((B * Lk, D), F16), e((B * Lk,), F32), e((B * Lk,), F32)
        ops.layernorm_fwd(x, _contig32(self.norm.weight), _contig32(self.norm.bias), self.norm.eps, y16=lat16, mean=mean_e, rstd=rstd_e)
        # decoder (:154-175)
        xk = e((B * Lk, Dd), F32)
        ops.linear(lat16, wc.w16(self.decoder_embed.weight), xk, bias=_contig32(self.decoder_embed.bias))
        xd = e((B * L, Dd), F32)
        ops.mae_unshuffle(xk.view(B, Lk, Dd), ids_restore, _contig32(self.mask_token).reshape(Dd),
                          _contig32(self.decoder_pos_embed).reshape(L, Dd), xd.view(B, L, Dd))
        dec_tape = []
        for blk in self.decoder_blocks:
            xd = vit_block_forward(wc, blk, xd, B, L, dec_tape)
        f16, mean_d, rstd_d = e((B * L, Dd), F16), e((B * L,), F32), e((B * L,), F32)
        ops.layernorm_fwd(xd, _contig32(self.decoder_norm.weight), _contig32(self.decoder_norm.bias), self.decoder_norm.eps, y16=f16,
                          mean=mean_d, rstd=rstd_d)
        pred = e((B, L, C * P * P), F32)
        ops.linear(f16, wc.w16(self.decoder_pred.weight), pred.view(B * L, -1), bias=_contig32(self.decoder_pred.bias))
        # loss (:177-198) — mean over ALL patches (mask_s = ones)
        loss = e((), F32)
        dpred = e((B * L, C * P * P), F32) if tape is not None else None
        ops.mae_loss(pred, imgs, loss, dpred, P, self.norm_pix_loss)
        if tape is not None:
            tape.update(B=B, L=L, Lk=Lk, patches=patches, ids_keep=ids_keep, ids_masked=ids_shuffle[:, Lk:].contiguous(), enc=enc_tape, x_enc=x, lat16=lat16,
                        mean_e=mean_e, rstd_e=rstd_e, dec=dec_tape, x_dec=xd, f16=f16, mean_d=mean_d, rstd_d=rstd_d, dpred=dpred)
        return loss, pred, mask

    def _backward_impl(self, t, g_loss):
        eng = engine()
        wc = eng.wc
        dev = g_loss.device
        B, L, Lk = t["B"], t["L"], t["Lk"]
        D = self.pos_embed.shape[-1]
        Dd = self.decoder_embed.weight.shape[0]
        names, params = self._trainable()
        arena, views = build_grad_arena(names, params, dev)
        ops.zero_(arena)
        ids = {id(p): n for n, p in zip(names, params)}
        G = lambda p: views[ids[id(p)]]  # noqa: E731
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        jobs = GradJobs()     # every weight / bias gradient is deferred to two grouped launches at the end (backward.GradJobs)
        # predictor
        dp16 = e(t["dpred"].shape, F16)
        ops.cast16_scaled(t["dpred"], g_loss.detach().to(F32).contiguous(), dp16)
        jobs.dB(dp16, G(self.decoder_pred.bias))
        jobs.dW(dp16, t["f16"], G(self.decoder_pred.weight))
        dh = e((B * L, Dd), F32)
        ops.linear(dp16, wc.w16_t(self.decoder_pred.weight), dh)
        g, g16 = e((B * L, Dd), F32), e((B * L, Dd), F16)
        dn = self.decoder_norm
        dblocks, eblocks = list(self.decoder_blocks), list(self.blocks)
        ops.layernorm_bwd(dh, t["x_dec"], _contig32(dn.weight), t["mean_d"], t["rstd_d"], g, G(dn.weight), G(dn.bias), accumulate=False, dx16=g16,
                          dx_colsum=G(dblocks[-1].mlp.fc2.bias), jobs=jobs)
        for k in range(len(dblocks) - 1, -1, -1):
            g16 = vit_block_backward(wc, dblocks[k], t["dec"][k], g, g16, G, jobs, next_bias=dblocks[k - 1].mlp.fc2.bias if k > 0 else None)
        # un-shuffle backward: d x_[b, j] = g[b, ids_shuffle[b, j]]; rows j < Lk belong to the kept tokens, the rest
        # all received the shared mask_token
        if L > Lk:
            gm = e((B, L - Lk, Dd), F32)
            ops.gather_rows(g.view(B, L, Dd), t["ids_masked"], gm)
            jobs.dB(gm.view(B * (L - Lk), Dd), G(self.mask_token))
        gk = e((B, Lk, Dd), F32)
        ops.gather_rows(g.view(B, L, Dd), t["ids_keep"], gk)
        gk16 = e((B * Lk, Dd), F16)
        ops.cast16(gk.view(B * Lk, Dd), gk16)
        de = self.decoder_embed
        jobs.dB(gk.view(B * Lk, Dd), G(de.bias))
        jobs.dW(gk16, t["lat16"], G(de.weight))
        # Data parallel: gradients become final from the END of the arena backwards (named_parameters order = forward order).
        # Whenever a suffix is final — after the decoder backward, then after every third encoder block — its deferred jobs are
        # flushed and the slice goes to the trainer's hook, so its all-reduce travels while the rest of the backward computes;
        # only the last slice (patch embedding + the first three blocks, a fifth of the 446.6 MB) is reduced after the backward.
        slice_hook = getattr(eng, "grad_slice_hook", None)
        off = lambda p: (G(p).data_ptr() - arena.data_ptr()) // 4  # noqa: E731
        hi = [arena.numel()]
        if slice_hook is not None:
            dec_lo = off(de.weight)
            if not all((off(p) >= dec_lo) == n.startswith(("decoder_embed", "decoder_blocks", "decoder_norm", "decoder_pred"))
                       for n, p in zip(names, params)):
                slice_hook = None          # unexpected parameter order: one all-reduce at the end

        def release(lo):
            if slice_hook is not None and lo < hi[0]:
                jobs.flush()
                slice_hook(arena[lo:hi[0]])
                hi[0] = lo

        if slice_hook is not None:
            release(dec_lo)
        dhe = e((B * Lk, D), F32)
        ops.linear(gk16, wc.w16_t(de.weight), dhe)
        ge, ge16 = e((B * Lk, D), F32), e((B * Lk, D), F16)
        ops.layernorm_bwd(dhe, t["x_enc"], _contig32(self.norm.weight), t["mean_e"], t["rstd_e"], ge, G(self.norm.weight), G(self.norm.bias),
                          accumulate=False, dx16=ge16, dx_colsum=G(eblocks[-1].mlp.fc2.bias), jobs=jobs)
        pe = self.patch_embed.proj
        for k in range(len(eblocks) - 1, -1, -1):
            ge16 = vit_block_backward(wc, eblocks[k], t["enc"][k], ge, ge16, G, jobs, next_bias=eblocks[k - 1].mlp.fc2.bias if k > 0 else pe.bias)
            if k > 0 and k % 3 == 0:
                # blocks k.. are final, except that block k's backward just added to fc2.bias of block k - 1 (outside the slice)
                release(off(eblocks[k].norm1.weight))
        # patch embedding: only the kept tokens carry gradient
        pk = e((B, Lk, t["patches"].shape[1]), F16)
        ops.gather_rows(t["patches"].view(B, L, -1), t["ids_keep"], pk)
        jobs.dW(ge16, pk.view(B * Lk, -1), G(pe.weight))
        jobs.flush()
        wc.bump(params)      # gradient received => about to be updated; fused optimizers do not bump `_version` (see backward.py)
        eng.last_arena = arena
        if slice_hook is not None:
            release(0)
        elif eng.grad_allreduce is not None:
            eng.grad_allreduce(arena)
        return views

    # -- reference API: the staged calls (models_mae_noct.py:137-198).  forward() runs the same schedule fused and is the only
    # differentiable entry point (FSC_pretrain.py:263 calls nothing else); the staged calls are inference-only inspection hooks.
    def _refresh(self):
        lin = [self.patch_embed.proj, self.decoder_embed, self.decoder_pred]
        for blk in list(self.blocks) + list(self.decoder_blocks):
            lin += [blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2]
        engine().wc.refresh_batch([(l.weight, "w") for l in lin])

    @torch.no_grad()
    def forward_encoder(self, x, mask_ratio):
        """imgs [N, 3, H, W] -> (latent fp32 [N, L_keep, D], mask [N, L], ids_restore [N, L])   (models_mae_noct.py:137-152)"""
        if not x.is_cuda:
            raise _lib.CountrError("countr_b200 runs on a B200 (sm_100a) only: got a CPU tensor and there is no CPU fallback")
        _lib.require_device()
        self._refresh()
        wc, dev = engine().wc, x.device
        B, C, Himg, Wimg = x.shape
        P = self.patch_embed.patch_size[0]
        L, D = (Himg // P) * (Wimg // P), self.pos_embed.shape[-1]
        patches = torch.empty(B * L, C * P * P, dtype=F16, device=dev)
        ops.patchify(x, patches, P)
        x_full = torch.empty(B * L, D, dtype=F32, device=dev)
        pe = self.patch_embed.proj
        ops.linear(patches, wc.w16(pe.weight), x_full, bias=_contig32(pe.bias), residual=_contig32(self.pos_embed).reshape(L, D), res_mod=L)
        Lk, ids_shuffle, ids_restore, mask = self._masking_indices(B, L, mask_ratio, dev)
        h = torch.empty(B * Lk, D, dtype=F32, device=dev)
        ops.gather_rows(x_full.view(B, L, D), ids_shuffle[:, :Lk].contiguous(), h.view(B, Lk, D))
        for blk in self.blocks:
            h = vit_block_forward(wc, blk, h, B, Lk, None)
        lat = torch.empty(B, Lk, D, dtype=F32, device=dev)
        ops.layernorm_fwd(h, _contig32(self.norm.weight), _contig32(self.norm.bias), self.norm.eps, y32=lat.view(B * Lk, D))
        return lat, mask, ids_restore

    @torch.no_grad()
    def forward_decoder(self, x, ids_restore):
        """latent [N, L_keep, D], ids_restore [N, L] -> pred fp32 [N, L, p*p*3]   (models_mae_noct.py:154-175)"""
        self._refresh()
        wc, dev = engine().wc, x.device
        B, Lk, D = x.shape
        L = ids_restore.shape[1]
        Dd = self.decoder_embed.weight.shape[0]
        lat16 = torch.empty(B * Lk, D, dtype=F16, device=dev)
        ops.cast16(x.detach().to(F32).contiguous().view(B * Lk, D), lat16)
        xk = torch.empty(B * Lk, Dd, dtype=F32, device=dev)
        ops.linear(lat16, wc.w16(self.decoder_embed.weight), xk, bias=_contig32(self.decoder_embed.bias))
        xd = torch.empty(B * L, Dd, dtype=F32, device=dev)
        ops.mae_unshuffle(xk.view(B, Lk, Dd), ids_restore.contiguous(), _contig32(self.mask_token).reshape(Dd),
                          _contig32(self.decoder_pos_embed).reshape(L, Dd), xd.view(B, L, Dd))
        for blk in self.decoder_blocks:
            xd = vit_block_forward(wc, blk, xd, B, L, None)
        f16 = torch.empty(B * L, Dd, dtype=F16, device=dev)
        ops.layernorm_fwd(xd, _contig32(self.decoder_norm.weight), _contig32(self.decoder_norm.bias), self.decoder_norm.eps, y16=f16)
        pred = torch.empty(B, L, self.decoder_pred.weight.shape[0], dtype=F32, device=dev)
        ops.linear(f16, wc.w16(self.decoder_pred.weight), pred.view(B * L, -1), bias=_contig32(self.decoder_pred.bias))
        return pred

    @torch.no_grad()
    def forward_loss(self, imgs, pred, mask):
        """Mean over ALL patches of the per-patch MSE — the reference replaces `mask` by ones (models_mae_noct.py:193-195)."""
        loss = torch.empty((), dtype=F32, device=imgs.device)
        ops.mae_loss(pred.detach().to(F32).contiguous(), imgs, loss, None, self.patch_embed.patch_size[0], self.norm_pix_loss)
        return loss

    def forward(self, imgs, mask_ratio=0.75):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            names, params = self._trainable()
            return _MAEFn.apply(self, imgs, mask_ratio, tuple(names), *params)
        with torch.no_grad():
            return self._forward_impl(imgs, mask_ratio, None)


def mae_vit_base_patch16_dec512d8b(**kwargs):
    return MaskedAutoencoderViTNoCT(patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=8,
                                    decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_large_patch16_dec512d8b(**kwargs):
    return MaskedAutoencoderViTNoCT(patch_size=16, embed_dim=1024, depth=24, num_heads=16, decoder_embed_dim=512, decoder_depth=8,
                                    decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_huge_patch14_dec512d8b(**kwargs):
    return MaskedAutoencoderViTNoCT(patch_size=14, embed_dim=1280, depth=32, num_heads=16, decoder_embed_dim=512, decoder_depth=8,
                                    decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


mae_vit_base_patch16 = mae_vit_base_patch16_dec512d8b
mae_vit_large_patch16 = mae_vit_large_patch16_dec512d8b
mae_vit_huge_patch14 = mae_vit_huge_patch14_dec512d8b
