"""Kernel schedules of the CounTR hot path (host side).

`encoder_forward` / `decoder_forward` / `decoder_backward` enqueue the sm_100a kernels of
libcountr_sm100.so in order on torch's current stream.  torch is used for buffer ownership only.

Data layout in HBM (per step, B images, L = 576 tokens, M = B*L):
  residual stream          fp32  [M, D]              (updated in place in eval; chained in training)
  GEMM / conv operands     fp16  [M, K] row-major / NHWC (TMA SWIZZLE_128B tiles)
  qkv                      fp16  [B, L, 3, H, dh]    (read in place by the attention kernel)
  weights                  fp16 copies of the fp32 master parameters, refreshed when a parameter's
                           version counter changes (every optimizer step for the decoder, never for
                           the frozen encoder)

Reference correspondence: encoder_forward = SupervisedMAE.forward_encoder (models_mae_cross.py:
136-148); decoder_forward = forward_decoder (:150-199).
"""
import math
import os
import weakref

import torch

from . import ops
from ._lib import lib

F16 = torch.float16
F32 = torch.float32


class Workspace:
    """Named device buffers reused across steps (no allocator traffic on the hot path)."""

    def __init__(self):
        self.bufs = {}

    def get(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self.bufs[key] = t
        return t


class WeightCache:
    """16-bit GEMM-operand copies of fp32 master parameters."""

    def __init__(self):
        self.cache = {}
        self.ext = {}          # id(p) -> counter bumped by trainers that update parameters with their own kernels
        self._tables = {}      # (device, stale set) -> device job table of refresh_batch
        self.lazy_fills = 0    # per-tensor refreshes done outside refresh_batch (diagnostic; tests/test_trainer_gpu.py)

    def bump(self, params):
        """Mark parameters as modified outside torch (torch's version counter did not move)."""
        for p in params:
            self.ext[id(p)] = self.ext.get(id(p), 0) + 1

    def _lookup(self, p, kind, shape, fill):
        # keyed by id(p) but validated with a weak reference: ids (and device pointers) are recycled when a
        # model is freed and another one is built, so identity must be checked, not assumed.
        key = (id(p), kind)
        ent = self.cache.get(key)
        ver = (p.data_ptr(), p._version, self.ext.get(id(p), 0))
        if ent is None or ent[0]() is not p or ent[1] != ver or ent[2].device != p.device:
            reuse = ent is not None and ent[2].device == p.device and ent[2].shape == torch.Size(shape)
            buf = ent[2] if reuse else torch.empty(shape, dtype=F16, device=p.device)
            fill(p.detach(), buf)
            self.lazy_fills += 1       # one launch per tensor: the batched refresh_batch() plans exist to keep this at zero
            ent = (weakref.ref(p), ver, buf)
            self.cache[key] = ent
            if len(self.cache) > 4096:      # drop entries of dead parameters
                self.cache = {k: e for k, e in self.cache.items() if e[0]() is not None}
        return ent[2]

    # kind -> (countr_weight_refresh job code, shape of the 16-bit copy)
    @staticmethod
    def _job(p, kind):
        n = p.shape[0]
        k = p.numel() // n if n else 0
        if kind == "w":
            return 0, (n, k), (n, k)
        if kind == "v":
            return 0, tuple(p.shape), (1, p.numel())
        if kind == "wt":
            return 1, (k, n), (n, k)
        cout, cin = p.shape[0], p.shape[1]
        if kind == "c0":
            return 2, (cout, 9 * cin), (cout, cin)
        return 3, (cin, 9 * cout), (cout, cin)          # "c1"

    def refresh_batch(self, plan):
        """Bring the 16-bit copies of every (parameter, kind) in `plan` up to date with ONE kernel launch (instead of one
        cast / transpose / pack launch per stale tensor — ~50 launches after every optimizer step of the fine-tune loop).
        The device job table is cached per set of stale tensors, so a steady training loop (and a CUDA-graph capture after
        warm-up) does no host-to-device traffic here."""
        work, pending = [], []
        for p, kind in plan:
            key = (id(p), kind)
            ent = self.cache.get(key)
            ver = (p.data_ptr(), p._version, self.ext.get(id(p), 0))
            if ent is not None and ent[0]() is p and ent[1] == ver and ent[2].device == p.device:
                continue
            code, shape, rc = self._job(p, kind)
            reuse = ent is not None and ent[2].device == p.device and ent[2].shape == torch.Size(shape)
            buf = ent[2] if reuse else torch.empty(shape, dtype=F16, device=p.device)
            pending.append((key, (weakref.ref(p), ver, buf)))      # committed only once the refresh kernel is enqueued
            work.append((p.data_ptr(), buf.data_ptr(), code, rc[0], rc[1]))
        if not work:
            return 0
        dev = plan[0][0].device
        tkey = (str(dev), tuple(work))
        tab = self._tables.get(tkey)
        if tab is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("countr_b200: the weight-refresh table must be built outside CUDA-graph capture "
                                   "(run one eager warm-up step before capturing)")
            rows, prefix, nblk = [], [0], 0
            for src, dst, code, r, c in work:
                rows += [src, dst, code, r, c, 0]
                nb = lib().countr_weight_refresh_blocks(code, r, c)
                assert nb >= 0
                nblk += nb
                prefix.append(nblk)
            tab = (torch.tensor(rows, dtype=torch.int64, device=dev), torch.tensor(prefix, dtype=torch.int32, device=dev), len(work), nblk)
            if len(self._tables) > 64:
                self._tables.clear()
            self._tables[tkey] = tab
        ops.weight_refresh(tab[0], tab[1], tab[2], tab[3])
        # a failed table build / launch raised above: the cache then still describes the old (stale, but initialised) copies
        self.cache.update(pending)
        return len(work)

    def invalidate(self, params=None):
        """Out-of-band weight edits (EMA / weight averaging, `p.data.copy_()` on encoder weights, anything that neither moves
        torch's version counter nor goes through a countr backward): mark the 16-bit copies of `params` (default: every
        cached parameter) stale so the next forward re-casts them."""
        if params is None:
            for key in list(self.cache):
                self.ext[key[0]] = self.ext.get(key[0], 0) + 1
        else:
            self.bump(params)

    def w16(self, p):
        """[N, K] row-major copy (B operand of y = x W^T)."""
        shape = (p.shape[0], p.numel() // p.shape[0])
        return self._lookup(p, "w", shape, lambda src, dst: ops.cast16(src, dst))

    def w16_t(self, p):
        """[K, N] transposed copy (B operand of dX = dY W)."""
        n, k = p.shape[0], p.numel() // p.shape[0]
        return self._lookup(p, "wt", (k, n), lambda src, dst: ops.cast16_transpose(src.reshape(n, k), dst))

    def conv16(self, p, mode=0):
        cout, cin = p.shape[0], p.shape[1]
        shape = (cout, 9 * cin) if mode == 0 else (cin, 9 * cout)
        return self._lookup(p, "c%d" % mode, shape, lambda src, dst: ops.conv_weight_pack(src, dst, mode))

    def v16(self, p):
        return self._lookup(p, "v", tuple(p.shape), lambda src, dst: ops.cast16(src, dst))


def _contig32(p):
    t = p.detach()
    assert t.dtype == F32 and t.is_contiguous(), "countr_b200 expects contiguous fp32 parameters"
    return t


class Engine:
    def __init__(self):
        self.ws = Workspace()
        self.wc = WeightCache()
        # optional callable(arena): averages the flat fp32 gradient arena over the data-parallel ranks
        # (one NCCL all-reduce per step, set by the trainer / bench when WORLD_SIZE > 1)
        self.grad_allreduce = None
        self.last_arena = None          # flat fp32 gradient arena of the most recent backward
        self._side = {}                 # per-device side stream: the exemplar CNN runs concurrently with the encoder
        self._flags = {}
        self.overlap_exemplar = True
        self.overlap_dw = True          # FIM weight / bias gradients on the side stream (backward.py)
        self.fold_layernorm = os.environ.get("COUNTR_FOLD_LN", "0") == "1"     # encoder: LayerNorm folded into the neighbouring GEMM epilogues
        self._folded = {}

    # ------------------------------------------------------------------ encoder
    def folded_linear(self, norm, lin):
        """Operands of a Linear with the LayerNorm in front of it folded in (frozen encoder; cached until any of the four tensors
        changes): W' = 16-bit(W diag(gamma)), colsum[n] = sum_k W'[n, k] (of the ROUNDED values the tensor core multiplies),
        bias' = b + W beta.  Init-time torch arithmetic, not on the hot path."""
        ps = (norm.weight, norm.bias, lin.weight, lin.bias)
        key = tuple(id(p) for p in ps)
        stamp = tuple((p.data_ptr(), p._version, self.wc.ext.get(id(p), 0)) for p in ps)
        ent = self._folded.get(key)
        if ent is None or ent[0] != stamp:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("countr_b200: folded LayerNorm operands must be built outside CUDA-graph capture "
                                   "(run one eager warm-up step before capturing)")
            with torch.no_grad():
                w = lin.weight.detach().to(F32)
                w16 = (w * norm.weight.detach().to(F32)[None, :]).to(F16).contiguous()
                colsum = w16.double().sum(1).to(F32).contiguous()
                bias = (lin.bias.detach().double() + w.double() @ norm.bias.detach().double()).to(F32).contiguous()
            ent = (stamp, w16, colsum, bias)
            self._folded[key] = ent
        return ent[1], ent[2], ent[3]

    def encoder_forward(self, m, imgs, keep=False):
        """m: SupervisedMAE-like module (patch_embed, pos_embed, blocks, norm).
        Returns (latent fp32 [B, L, D], latent fp16 [B*L, D]).  keep=True (training): the fp16 latent is saved for the decoder
        backward (decoder_embed's weight gradient), so it gets its own buffer instead of the shared workspace slot that the next
        forward — a validation pass, a second micro-batch, a sliding-window evaluation — would overwrite."""
        ops_ = ops
        dev = imgs.device
        B, C, Himg, Wimg = imgs.shape
        P = m.patch_embed.patch_size[0]
        assert Himg == m.patch_embed.img_size[0] and Wimg == m.patch_embed.img_size[1], \
            f"Input image size ({Himg}*{Wimg}) doesn't match model ({m.patch_embed.img_size[0]}*{m.patch_embed.img_size[1]})."
        L = (Himg // P) * (Wimg // P)
        M = B * L
        D = m.pos_embed.shape[-1]
        ws, wc = self.ws, self.wc
        Kp = C * P * P
        ld = (Kp + 7) & ~7             # TMA row pitch: multiple of 16 bytes (14-px patches: 588 -> 592, zero padded)
        patches = ws.get("patches", (M, ld), F16, dev)
        ops_.patchify(imgs, patches, P)
        x = ws.get("enc_x", (M, D), F32, dev)
        pe = m.patch_embed.proj
        w_pe = wc.w16(pe.weight)
        if ld != Kp:
            # not a benchmarked configuration (mae_vit_huge_patch14 only): zero-padded copy of the filter matrix per forward
            wpad = ws.get("pe_wpad", (D, ld), F16, dev)
            wpad.zero_()
            wpad[:, :Kp].copy_(w_pe)
            w_pe = wpad
        ops_.linear(patches, w_pe, x, bias=_contig32(pe.bias), residual=_contig32(m.pos_embed).reshape(L, D), res_mod=L)
        h = ws.get("enc_h", (M, D), F16, dev)
        # LayerNorm folded into the GEMMs around it (csrc/gemm.cu, countr_gemm_desc): proj / fc2 also emit a 16-bit copy of the
        # residual stream and per-row statistics partials, fc1 / the next block's qkv multiply that copy by W diag(gamma) and
        # normalise in their epilogue.  23 of the encoder's 25 LayerNorm launches disappear; block 0's norm1 (fed by the patch
        # embedding) and the final norm (fp32 + 16-bit outputs) stay kernels.  Needs D / 4 to be a GEMM tile width.
        # OFF by default (COUNTR_FOLD_LN=1 / Engine.fold_layernorm turn it on): measured on B200 at B = 8 inside the CUDA graph the
        # 23 saved launches (~5.6 us each) are paid back by ~3 us of extra epilogue per GEMM on 46 GEMMs — encoder 1347 us folded
        # vs 1331 us with LayerNorm kernels (profiles/README.md).  It wins where launches are synchronous (the scripts'
        # CUDA_LAUNCH_BLOCKING=1).
        fold = self.fold_layernorm and D % 4 == 0 and (D // 4) in (128, 192, 256)
        x16 = ws.get("enc_x16", (M, D), F16, dev) if fold else None
        st = ws.get("enc_lnstats", (M, 8, 2), F32, dev) if fold else None
        nblk = len(m.blocks)
        for bi, blk in enumerate(m.blocks):
            H = blk.attn.num_heads
            dh = D // H
            hid = blk.mlp.fc1.weight.shape[0]
            qkv = ws.get("enc_qkv", (M, 3 * D), F16, dev)
            att = ws.get("enc_att", (M, D), F16, dev)
            u = ws.get("enc_u", (M, hid), F16, dev)
            if fold and bi > 0:
                wq, sq, bq = self.folded_linear(blk.norm1, blk.attn.qkv)
                ops_.linear(x16, wq, qkv, bias=bq, ln_stats=st, ln_colsum=sq, ln_eps=blk.norm1.eps)
            else:
                ops_.layernorm_fwd(x, _contig32(blk.norm1.weight), _contig32(blk.norm1.bias), blk.norm1.eps, y16=h)
                ops_.linear(h, wc.w16(blk.attn.qkv.weight), qkv, bias=_contig32(blk.attn.qkv.bias))
            ops_.attention_fwd(qkv, att, B, L, H, dh, blk.attn.scale)
            if fold:
                ops_.linear(att, wc.w16(blk.attn.proj.weight), x, bias=_contig32(blk.attn.proj.bias), residual=x, bn=D // 4, ln_x16=x16,
                            ln_stats=st)
                w1, s1, b1 = self.folded_linear(blk.norm2, blk.mlp.fc1)
                ops_.linear(x16, w1, u, bias=b1, act=1, ln_stats=st, ln_colsum=s1, ln_eps=blk.norm2.eps)
            else:
                ops_.linear(att, wc.w16(blk.attn.proj.weight), x, bias=_contig32(blk.attn.proj.bias), residual=x)
                ops_.layernorm_fwd(x, _contig32(blk.norm2.weight), _contig32(blk.norm2.bias), blk.norm2.eps, y16=h)
                ops_.linear(h, wc.w16(blk.mlp.fc1.weight), u, bias=_contig32(blk.mlp.fc1.bias), act=1)
            if fold and bi + 1 < nblk:
                ops_.linear(u, wc.w16(blk.mlp.fc2.weight), x, bias=_contig32(blk.mlp.fc2.bias), residual=x, bn=D // 4, ln_x16=x16,
                            ln_stats=st)
            else:
                ops_.linear(u, wc.w16(blk.mlp.fc2.weight), x, bias=_contig32(blk.mlp.fc2.bias), residual=x)
        lat32 = torch.empty(B, L, D, dtype=F32, device=dev)
        lat16 = torch.empty(M, D, dtype=F16, device=dev) if keep else ws.get("lat16", (M, D), F16, dev)
        ops_.layernorm_fwd(x, _contig32(m.norm.weight), _contig32(m.norm.bias), m.norm.eps, y16=lat16, y32=lat32.view(M, D))
        return lat32, lat16

    def usage_flags(self, shot_num, dev):
        """Device constant written to the tail of the gradient arena (dist.usage_flags); built on first use, which must
        happen outside CUDA-graph capture (the eager warm-up step)."""
        from .dist import usage_flags
        key = (str(dev), shot_num == 0)
        t = self._flags.get(key)
        if t is None:
            t = torch.tensor(usage_flags(shot_num), dtype=F32, device=dev)
            self._flags[key] = t
        return t

    # ------------------------------------------------------------------ exemplar encoder
    def side_stream(self, dev):
        key = str(dev)
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=dev)
        return self._side[key]

    def refresh_decoder_weights(self, m, shot_num, train, dev):
        """WeightCache.refresh_batch over the decoder's plan.  With the side stream enabled the launch goes there and an event
        recorded behind it is returned: the caller makes its stream wait for it before the first decoder kernel
        (decoder_embed reads a refreshed copy before decoder_forward joins the exemplar branch).  Otherwise the refresh runs on
        the current stream and None is returned."""
        plan = self.decoder_weight_plan(m, shot_num, train)
        if not self.overlap_exemplar:
            self.wc.refresh_batch(plan)
            return None
        side = self.side_stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            if self.wc.refresh_batch(plan) == 0:
                return None
            ev = torch.cuda.Event()
            ev.record(side)
        return ev

    def exemplar_async(self, m, boxes, S, train):
        """Start decoder_proj1..4 on the side stream (it only depends on the boxes): ~45 tiny latency-bound
        launches that would otherwise sit on the critical path overlap the encoder instead.  The consumer calls
        `current_stream().wait_stream(side_stream)` before touching the result (decoder_forward does)."""
        side = self.side_stream(boxes.device)
        side.wait_stream(torch.cuda.current_stream())
        saved = {} if train else None
        with torch.cuda.stream(side):
            y32, y16 = self.exemplar_forward(m, boxes, S, saved)
            kv = self.kv_project(m, y16, train)     # wk(y), wv(y) of every FIM block: they only depend on y
        return dict(y32=y32, y16=y16, kv=kv, saved=None if saved is None else saved["exemplar"], side=side)

    def kv_project(self, m, y16, train):
        """k = wk(y), v = wv(y) of every CrossAttentionBlock (models_crossvit.py:115-116).  y is the same for all blocks
        (models_mae_cross.py:180-181) and does not depend on the image tokens, so these tiny (M = B*S rows) latency-bound
        GEMMs are taken off the critical path: they run right behind the exemplar CNN on the side stream."""
        dev = y16.device
        ny = y16.shape[0]
        out = []
        for i, blk in enumerate(m.decoder_blocks):
            Dd = blk.attn.wk.weight.shape[0]
            if train:
                k32 = torch.empty(ny, Dd, dtype=F32, device=dev)
                v32 = torch.empty(ny, Dd, dtype=F32, device=dev)
            else:
                k32 = self.ws.get(f"dec_k{i}", (ny, Dd), F32, dev)
                v32 = self.ws.get(f"dec_v{i}", (ny, Dd), F32, dev)
            ops.linear(y16, self.wc.w16(blk.attn.wk.weight), k32, bias=_contig32(blk.attn.wk.bias))
            ops.linear(y16, self.wc.w16(blk.attn.wv.weight), v32, bias=_contig32(blk.attn.wv.bias))
            out.append((k32, v32))
        return out

    def exemplar_forward(self, m, boxes, S, save):
        """decoder_proj1..4 on the first S boxes of every image -> y32 [B*S, C], y16 [B*S, C]."""
        dev = boxes.device
        B = boxes.shape[0]
        N = B * S
        ws, wc = self.ws, self.wc
        get = (lambda name, shape, dt: torch.empty(shape, dtype=dt, device=dev)) if save is not None else \
            (lambda name, shape, dt: ws.get(name, shape, dt, dev))
        c1 = m.decoder_proj1[0]
        side = boxes.shape[-1]
        raw = get("ex_raw1", (N, side, side, c1.weight.shape[0]), F16)
        ops.exemplar_conv1(boxes, S, _contig32(c1.weight), _contig32(c1.bias), raw)
        stages = [m.decoder_proj2[0], m.decoder_proj3[0], m.decoder_proj4[0]]
        saved = {"raw": [raw], "pooled": [], "mean": [], "rstd": []}
        cur = raw
        for i, conv in enumerate(stages):
            n, hh, ww, cc = cur.shape
            pooled = get(f"ex_pool{i}", (n, hh // 2, ww // 2, cc), F16)
            mean = get(f"ex_mean{i}", (n, cc), F32)       # always: enables the pixel-parallel statistics path
            rstd = get(f"ex_rstd{i}", (n, cc), F32)
            scratch = ws.get("ex_scratch", (64 * n * cc,), F32, dev)   # per-split partial sums (deterministic reduction)
            ops.inorm_relu_pool(cur, 0, 1e-5, y16=pooled, mean=mean, rstd=rstd, scratch=scratch)
            cout = conv.weight.shape[0]
            nxt = get(f"ex_raw{i + 2}", (n, hh // 2, ww // 2, cout), F16)
            ops.conv3x3(pooled, wc.conv16(conv.weight), nxt, bias=_contig32(conv.bias))
            saved["pooled"].append(pooled); saved["mean"].append(mean); saved["rstd"].append(rstd); saved["raw"].append(nxt)
            cur = nxt
        n, hh, ww, cc = cur.shape
        y32 = get("ex_y32", (N, cc), F32)
        y16 = get("ex_y16", (N, cc), F16)
        mean = get("ex_mean3", (n, cc), F32) if save is not None else None
        rstd = get("ex_rstd3", (n, cc), F32) if save is not None else None
        ops.inorm_relu_pool(cur, 1, 1e-5, y16=y16, y32=y32, mean=mean, rstd=rstd)
        saved["mean"].append(mean); saved["rstd"].append(rstd)
        if save is not None:
            save["exemplar"] = saved
        return y32, y16

    # ------------------------------------------------------------------ decoder
    @staticmethod
    def decoder_weight_plan(m, shot_num, train):
        """Every (parameter, operand-copy kind) the decoder forward — and, when training, its backward — will ask the
        WeightCache for: 'w' row-major, 'wt' transposed (dX GEMMs), 'c0' / 'c1' packed conv filters (forward / dX)."""
        plan = [(m.decoder_embed.weight, "w")]
        ex = [m.decoder_proj2[0], m.decoder_proj3[0], m.decoder_proj4[0]]
        if shot_num > 0:
            plan += [(c.weight, "c0") for c in ex]
        else:
            plan.append((m.shot_token, "v"))
        for blk in m.decoder_blocks:
            lin = [blk.selfattn.qkv, blk.selfattn.proj, blk.attn.wq, blk.attn.wk, blk.attn.wv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2]
            plan += [(l.weight, "w") for l in lin]
            if train:
                plan += [(l.weight, "wt") for l in lin]
        heads = [m.decode_head0[0], m.decode_head1[0], m.decode_head2[0], m.decode_head3[0]]
        plan += [(c.weight, "c0") for c in heads]
        if train:
            plan += [(c.weight, "c1") for c in heads]
            if shot_num > 0:
                plan += [(c.weight, "c1") for c in ex]
        return plan

    def decoder_forward(self, m, lat16, boxes, shot_num, B, out_dtype, save=None, pre=None):
        """lat16: fp16 [B*L, D] encoder output.  Returns the density map [B, 2^4*h, 2^4*w]."""
        dev = lat16.device
        ws, wc = self.ws, self.wc
        M, D = lat16.shape
        L = M // B
        Dd = m.decoder_embed.weight.shape[0]
        train = save is not None
        # in training every activation the backward needs gets its own buffer
        get = (lambda name, shape, dt: torch.empty(shape, dtype=dt, device=dev)) if train else \
            (lambda name, shape, dt: ws.get(name, shape, dt, dev))

        x = get("dec_x", (M, Dd), F32)
        ops.linear(lat16, wc.w16(m.decoder_embed.weight), x, bias=_contig32(m.decoder_embed.bias),
                   residual=_contig32(m.decoder_pos_embed).reshape(L, Dd), res_mod=L)

        if shot_num > 0:
            assert boxes.dim() == 5 and boxes.shape[1] >= shot_num, "boxes must be [N, K>=shot_num, 3, 64, 64]"
            S = shot_num
            if pre is not None:       # computed concurrently on the side stream (exemplar_async)
                y32, y16, kvs = pre["y32"], pre["y16"], pre["kv"]
                if train:
                    save["exemplar"] = pre["saved"]
                torch.cuda.current_stream().wait_stream(pre["side"])
            else:
                y32, y16 = self.exemplar_forward(m, boxes, S, save)
                kvs = self.kv_project(m, y16, train)
            kv_broadcast = False
        else:
            S = 1
            y16 = wc.v16(m.shot_token).reshape(1, Dd)     # same token for every image (models_mae_cross.py:176)
            kvs = self.kv_project(m, y16, train)
            kv_broadcast = True

        blocks_saved = []
        for bi, blk in enumerate(m.decoder_blocks):
            H = blk.selfattn.num_heads
            dh = Dd // H
            hid = blk.mlp.fc1.weight.shape[0]
            sv = {}
            # --- self attention
            h0 = get("dec_h0", (M, Dd), F16)
            mean0 = get("m0", (M,), F32) if train else None
            rstd0 = get("r0", (M,), F32) if train else None
            ops.layernorm_fwd(x, _contig32(blk.norm0.weight), _contig32(blk.norm0.bias), blk.norm0.eps, y16=h0, mean=mean0, rstd=rstd0)
            qkv = get("dec_qkv", (M, 3 * Dd), F16)
            ops.linear(h0, wc.w16(blk.selfattn.qkv.weight), qkv, bias=_contig32(blk.selfattn.qkv.bias))
            att = get("dec_att", (M, Dd), F16)
            lse = get("dec_lse", (B, H, L), F32) if train else None
            ops.attention_fwd(qkv, att, B, L, H, dh, blk.selfattn.scale, lse=lse)
            x1 = get("dec_x1", (M, Dd), F32) if train else x
            ops.linear(att, wc.w16(blk.selfattn.proj.weight), x1, bias=_contig32(blk.selfattn.proj.bias), residual=x)
            # --- cross attention with the exemplar tokens
            h1 = get("dec_h1", (M, Dd), F16)
            mean1 = get("m1", (M,), F32) if train else None
            rstd1 = get("r1", (M,), F32) if train else None
            ops.layernorm_fwd(x1, _contig32(blk.norm1.weight), _contig32(blk.norm1.bias), blk.norm1.eps, y16=h1, mean=mean1, rstd=rstd1)
            q16 = get("dec_q", (M, Dd), F16)
            ops.linear(h1, wc.w16(blk.attn.wq.weight), q16, bias=_contig32(blk.attn.wq.bias))
            k32, v32 = kvs[bi]
            c16 = get("dec_c", (M, Dd), F16)
            probs = get("dec_probs", (M, H, S), F32) if train else None
            ops.cross_attn_core(q16, k32, v32, c16, B, L, S, Dd, dh, blk.attn.scale, probs=probs, kv_broadcast=kv_broadcast)
            x2 = get("dec_x2", (M, Dd), F32) if train else x1
            ops.linear(c16, wc.w16(blk.attn.proj.weight), x2, bias=_contig32(blk.attn.proj.bias), residual=x1)
            # --- MLP
            h2 = get("dec_h2", (M, Dd), F16)
            mean2 = get("m2", (M,), F32) if train else None
            rstd2 = get("r2", (M,), F32) if train else None
            ops.layernorm_fwd(x2, _contig32(blk.norm2.weight), _contig32(blk.norm2.bias), blk.norm2.eps, y16=h2, mean=mean2, rstd=rstd2)
            u = get("dec_u", (M, hid), F16)
            pre = get("dec_pre", (M, hid), F16) if train else None
            ops.linear(h2, wc.w16(blk.mlp.fc1.weight), u, bias=_contig32(blk.mlp.fc1.bias), act=1, aux=pre)
            x3 = get("dec_x3", (M, Dd), F32) if train else x2
            ops.linear(u, wc.w16(blk.mlp.fc2.weight), x3, bias=_contig32(blk.mlp.fc2.bias), residual=x2)
            if train:
                sv.update(x0=x, h0=h0, mean0=mean0, rstd0=rstd0, qkv=qkv, att=att, lse=lse, x1=x1, h1=h1, mean1=mean1,
                          rstd1=rstd1, q16=q16, k32=k32, v32=v32, c16=c16, probs=probs, x2=x2, h2=h2, mean2=mean2,
                          rstd2=rstd2, u=u, pre=pre)
                blocks_saved.append(sv)
            x = x3

        f16 = get("dec_f", (M, Dd), F16)
        meanf = get("mf", (M,), F32) if train else None
        rstdf = get("rf", (M,), F32) if train else None
        ops.layernorm_fwd(x, _contig32(m.decoder_norm.weight), _contig32(m.decoder_norm.bias), m.decoder_norm.eps, y16=f16,
                          mean=meanf, rstd=rstdf)

        # --- density head: tokens [B, L, C] are already NHWC [B, h, w, C]
        hgt = wdt = int(math.sqrt(L))
        cur = f16.view(B, hgt, wdt, Dd)
        heads = [m.decode_head0, m.decode_head1, m.decode_head2, m.decode_head3]
        head_saved = []
        out = None
        for i, head in enumerate(heads):
            conv, gn = head[0], head[1]
            cout = conv.weight.shape[0]
            G = gn.num_groups
            assert cout // G == 32, "GroupNorm statistics are fused for 32-channel groups"
            raw = get(f"head_raw{i}", (B, hgt, wdt, cout), F16)
            stats = get(f"head_stats{i}", (B, G, 2), torch.float64)
            ops.zero_(stats)
            ops.conv3x3(cur, wc.conv16(conv.weight), raw, bias=_contig32(conv.bias), gn_stats=stats)
            head_saved.append(dict(inp=cur, raw=raw, stats=stats))
            if i < 3:
                nxt = get(f"head_in{i + 1}", (B, 2 * hgt, 2 * wdt, cout), F16)
                ops.gn_relu_upsample2x(raw, stats, _contig32(gn.weight), _contig32(gn.bias), nxt, G, gn.eps)
                cur = nxt
                hgt, wdt = 2 * hgt, 2 * wdt
            else:
                c1 = head[3]
                dmap = get("head_dmap", (B, hgt, wdt), F32)
                ops.gn_relu_conv1x1(raw, stats, _contig32(gn.weight), _contig32(gn.bias), _contig32(c1.weight).reshape(-1),
                                    _contig32(c1.bias), dmap, G, gn.eps)
                out = torch.empty(B, 2 * hgt, 2 * wdt, dtype=out_dtype, device=dev)
                ops.upsample2x_f32(dmap, out)
        if train:
            save.update(blocks=blocks_saved, x_final=x, meanf=meanf, rstdf=rstdf, f16=f16, heads=head_saved, lat16=lat16,
                        y16=y16, S=S, kv_broadcast=kv_broadcast, B=B, L=L, shot_num=shot_num)
        return out


_ENGINE = None


def engine():
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine()
    return _ENGINE
