"""FineTuner — the fine-tune step of FSC_finetune_cross.py (:265-319) as one kernel schedule.

    tuner = FineTuner(model, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95))
    tuner.set_lr(lr_t)                                                 # per-iteration schedule (lr_sched, :270)
    loss = tuner.step(imgs, boxes, gt_density, mask, shot_num)         # device scalar tensor
    tuner.metrics()                                                    # loss, batch MAE / MSE, grad norm, scale, found_inf

Same arithmetic as the reference loop — frozen-encoder forward, decoder forward, masked-MSE loss, per-image counts
(sum/60), decoder backward, (gradient all-reduce), GradScaler unscale / inf check / skip / update, get_grad_norm_,
AdamW with the timm `add_weight_decay` grouping — but without autograd, GradScaler host bookkeeping or per-parameter
optimizer launches: gradients and both Adam moments live in flat fp32 arenas, the loss (+ gradient + counts + optional
device-side Bernoulli mask) is one kernel, the inf check + norm one kernel and the whole optimizer update one kernel.
Loss scale, learning rate, found_inf and the step counter live in a device state block, so nothing here synchronises
with the host and a step captured in a CUDA graph follows the lr schedule and the dynamic loss scale.
"""
import ctypes
import math

import numpy as np
import torch

from . import ops
from ._lib import check, lib
from .backward import decoder_backward
from .dist import ARENA_TAIL, arena_size, param_flag_index
from .engine import F32, engine

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
ST_SCALE, ST_GROWTH, ST_FOUND_INF, ST_GRAD_NORM, ST_LR, ST_STEP = range(6)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def scheduled_lr(epoch, lr, min_lr=0.0, warmup_epochs=10, epochs=200):
    """The learning rate the scripts set before every iteration (util/lr_sched.py:9-21, called with the fractional epoch
    `data_iter_step / len(loader) + epoch`, FSC_finetune_cross.py:270-271): linear warm-up, then half a cosine down to min_lr.
    Host arithmetic; feed the result to `FineTuner.set_lr` / `ArenaAdamW.set_lr`."""
    if epoch < warmup_epochs:
        return lr * epoch / warmup_epochs
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)))


class FineTuner:
    def __init__(self, model, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95), eps=1e-8, loss_scale=65536.0, allreduce=None,
                 dynamic_scale=True, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, mask_keep_prob=0.8, seed=0):
        """loss_scale / growth_* / backoff_* default to torch.cuda.amp.GradScaler's (util/misc.py:263);
        dynamic_scale=False keeps loss_scale fixed (overflowed steps are still skipped)."""
        self.model, self.wd, self.betas, self.eps = model, weight_decay, betas, eps
        self.allreduce = allreduce          # callable(arena) or None
        self.growth, self.backoff = float(growth_factor), float(backoff_factor)
        self.interval = int(growth_interval) if dynamic_scale else 0
        self.keep_prob, self.seed = float(mask_keep_prob), int(seed)
        self.dev = next(model.parameters()).device
        st = [0.0] * 8
        st[ST_SCALE], st[ST_LR] = float(loss_scale), float(lr)
        self.state = torch.tensor(st, dtype=F32, device=self.dev)
        self._lr_host = torch.zeros(1, dtype=F32).pin_memory() if torch.cuda.is_available() else torch.zeros(1)
        self.result = torch.zeros(3, dtype=F32, device=self.dev)         # loss, batch MAE, batch MSE of the counts
        self.loss = self.result[0]
        self.counts = None
        self._loss_scratch = {}
        self._stats_scratch = torch.zeros(int(lib().countr_grad_stats_scratch_bytes()) // 8, dtype=torch.float64, device=self.dev)
        self._build_tables()

    # ------------------------------------------------------------------ optimizer tables (built once)
    def _build_tables(self):
        names, params = self.model._decoder_params(None)
        n_grad = arena_size(params)
        self.exp_avg = torch.zeros(n_grad, dtype=F32, device=self.dev)
        self.exp_avg_sq = torch.zeros(n_grad, dtype=F32, device=self.dev)
        self.step_count = torch.zeros(len(names), dtype=F32, device=self.dev)   # one counter per parameter, like torch.optim
        rec = np.zeros(len(names), dtype=np.dtype([("param", "<u8"), ("goff", "<i8"), ("moff", "<i8"), ("numel", "<i8"), ("wd", "<f4"),
                                                   ("step_idx", "<i4"), ("flag_idx", "<i4"), ("pad", "<i4")]))
        assert rec.dtype.itemsize == 48
        chunks, off = [], 0
        for i, (n, p) in enumerate(zip(names, params)):
            assert p.dtype == F32 and p.is_contiguous()
            wd = 0.0 if (p.ndim == 1 or n.endswith(".bias")) else self.wd          # timm add_weight_decay (FSC_finetune_cross.py:234)
            rec[i] = (p.data_ptr(), off, off, p.numel(), wd, i, param_flag_index(n), 0)
            chunks += [(i, c) for c in range((p.numel() + 1023) // 1024)]
            off += (p.numel() + 3) // 4 * 4
        self.names, self.params = names, params
        self.n_grad = n_grad
        self._tensors = torch.from_numpy(rec.view(np.uint8).copy()).to(self.dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int32, device=self.dev)
        self._n_chunks = len(chunks)

    # ------------------------------------------------------------------ schedule / state
    def set_lr(self, lr):
        """Per-iteration learning rate (the script's lr_sched.adjust_learning_rate): a 4-byte async H2D into the state
        block, which a replayed CUDA graph reads."""
        self._lr_host[0] = float(lr)
        self.state[ST_LR:ST_LR + 1].copy_(self._lr_host, non_blocking=True)

    def metrics(self):
        """One D2H of everything the script logs per step (:298-319)."""
        r, s = self.result.tolist(), self.state.tolist()
        return dict(loss=r[0], batch_mae=r[1], batch_mse=r[2], grad_norm=s[ST_GRAD_NORM], loss_scale=s[ST_SCALE],
                    found_inf=bool(s[ST_FOUND_INF]), lr=s[ST_LR], step=int(s[ST_STEP]))

    # ------------------------------------------------------------------ the step
    def _loss(self, out, gt_density, mask, dout):
        B, H, W = out.shape
        assert gt_density.shape == out.shape and gt_density.is_contiguous() and gt_density.device == out.device, \
            "gt_density must be a contiguous [B, H, W] tensor on the model's device"
        assert gt_density.dtype in _DTYPE_CODE, f"gt_density dtype {gt_density.dtype} not supported (fp32 / fp16 / bf16)"
        bstride = 0
        if mask is not None:
            # the script builds an int64 mask tiled to [B, H, W] (:290-293); one [H, W] float mask is the compact form of it
            if mask.dtype != F32:
                mask = mask.to(F32)
            assert mask.device == out.device and mask.shape in ((H, W), (B, H, W)), "mask must be [H, W] or [B, H, W]"
            mask = mask.contiguous()
            bstride = H * W if mask.dim() == 3 else 0
        scr = self._loss_scratch.get(B)
        if scr is None:
            scr = torch.zeros(int(lib().countr_finetune_loss_scratch_bytes(B)) // 8, dtype=torch.float64, device=self.dev)
            self._loss_scratch[B] = scr
        if self.counts is None or self.counts.shape[0] != B:
            self.counts = torch.zeros(B, 2, dtype=F32, device=self.dev)
        check(lib().countr_finetune_loss(_p(out), _DTYPE_CODE[out.dtype], _p(gt_density), _DTYPE_CODE[gt_density.dtype], _p(mask), bstride,
                                         self.seed, self.keep_prob, _p(self.state), 0.0, _p(dout), None, _p(scr), _p(self.result),
                                         _p(self.counts), B, H, W, ops._stream()))
        ops._count()
        self._keep = mask       # alive until the kernel has run

    @torch.no_grad()
    def encode(self, imgs):
        """Frozen-encoder forward (models_mae_cross.py:204-205) -> fp16 latent [B*L, D] in the engine's workspace.  It does not
        depend on any trainable parameter, so a data-parallel loop may run it for batch i+1 while the gradient all-reduce
        of batch i is in flight (bench.py); pass a COPY of the result to forward_backward(lat16=...)."""
        return engine().encoder_forward(self.model, imgs, keep=False)[1]

    @torch.no_grad()
    def forward_backward(self, imgs, boxes, gt_density, mask, shot_num, lat16=None):
        """Forward, loss and backward; leaves the (scaled) gradients in `self.arena`.  Returns the loss (device scalar).
        mask=None draws the Bernoulli(0.8) pixel mask on the device (FSC_finetune_cross.py:290); lat16: the encoder output of
        `imgs` if it was computed ahead (see encode)."""
        m, eng = self.model, engine()
        B = imgs.shape[0]
        on_side = eng.refresh_decoder_weights(m, shot_num, True, imgs.device)    # all stale 16-bit weight copies, one launch
        pre = eng.exemplar_async(m, boxes, shot_num, train=True) if (shot_num > 0 and eng.overlap_exemplar) else None
        if lat16 is None:
            _, lat16 = eng.encoder_forward(m, imgs, keep=True)
        if on_side is not None:
            torch.cuda.current_stream().wait_event(on_side)
        save = {}
        out = eng.decoder_forward(m, lat16, boxes, shot_num, B, F32, save=save, pre=pre)
        dout = torch.empty_like(out)
        self._loss(out, gt_density, mask, dout)
        self.last_output = out
        hook, eng.grad_allreduce = eng.grad_allreduce, None      # the all-reduce is issued by step() / the caller, not inside
        try:
            decoder_backward(eng, m, save, boxes, dout)
        finally:
            eng.grad_allreduce = hook
        self.arena = eng.last_arena
        return self.loss

    @torch.no_grad()
    def update(self):
        """inf check + gradient norm + unscale + AdamW + GradScaler.update on the arena left by forward_backward
        (all-reduce it first when data-parallel)."""
        eng = engine()
        arena = self.arena
        assert arena.numel() == self.n_grad + ARENA_TAIL
        check(lib().countr_grad_stats(_p(arena), self.n_grad, _p(self._stats_scratch), _p(self.state), ops._stream()))
        flags = ctypes.c_void_p(arena.data_ptr() + 4 * self.n_grad)
        check(lib().countr_adamw_update(_p(self._tensors), len(self.names), _p(self._chunks), self._n_chunks, _p(arena), flags,
                                        _p(self.exp_avg), _p(self.exp_avg_sq), _p(self.step_count), _p(self.state), self.betas[0],
                                        self.betas[1], self.eps, self.growth, self.backoff, self.interval, ops._stream()))
        ops._count(3)
        eng.wc.bump(self.params)      # parameters changed behind torch's back: refresh their fp16 copies next step

    def step(self, imgs, boxes, gt_density, mask, shot_num):
        loss = self.forward_backward(imgs, boxes, gt_density, mask, shot_num)
        if self.allreduce is not None:
            self.allreduce(self.arena)
        self.update()
        return loss


class ArenaAdamW:
    """GradScaler.unscale_ / inf check / get_grad_norm_ / AdamW / GradScaler.update (util/misc.py:260-301 + torch.optim.AdamW with
    the timm add_weight_decay grouping, FSC_pretrain.py:226-228) over a flat fp32 gradient arena, as three kernels: the same
    device-side state block and update kernels FineTuner uses, for any model whose backward leaves its gradients in one arena
    (models_mae_noct: `engine().last_arena`).

        opt = ArenaAdamW(*model._trainable(), lr=..., weight_decay=0.05, betas=(0.9, 0.95), loss_scale=1024.0)
        (loss * opt.scale()).backward();  opt.step(engine().last_arena)
    """

    def __init__(self, names, params, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95), eps=1e-8, loss_scale=65536.0, dynamic_scale=True,
                 growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.names, self.params = list(names), list(params)
        self.betas, self.eps = betas, eps
        self.growth, self.backoff = float(growth_factor), float(backoff_factor)
        self.interval = int(growth_interval) if dynamic_scale else 0
        self.dev = self.params[0].device
        st = [0.0] * 8
        st[ST_SCALE], st[ST_LR] = float(loss_scale), float(lr)
        self.state = torch.tensor(st, dtype=F32, device=self.dev)
        self._lr_host = torch.zeros(1, dtype=F32).pin_memory() if torch.cuda.is_available() else torch.zeros(1)
        self._stats_scratch = torch.zeros(int(lib().countr_grad_stats_scratch_bytes()) // 8, dtype=torch.float64, device=self.dev)
        self.n_grad = arena_size(self.params)
        self.exp_avg = torch.zeros(self.n_grad, dtype=F32, device=self.dev)
        self.exp_avg_sq = torch.zeros(self.n_grad, dtype=F32, device=self.dev)
        self.step_count = torch.zeros(len(self.names), dtype=F32, device=self.dev)
        self._flags = torch.zeros(ARENA_TAIL, dtype=F32, device=self.dev)          # no optional parameter groups
        rec = np.zeros(len(self.names), dtype=np.dtype([("param", "<u8"), ("goff", "<i8"), ("moff", "<i8"), ("numel", "<i8"), ("wd", "<f4"),
                                                        ("step_idx", "<i4"), ("flag_idx", "<i4"), ("pad", "<i4")]))
        chunks, off = [], 0
        for i, (n, p) in enumerate(zip(self.names, self.params)):
            assert p.dtype == F32 and p.is_contiguous()
            wd = 0.0 if (p.ndim == 1 or n.endswith(".bias")) else weight_decay
            rec[i] = (p.data_ptr(), off, off, p.numel(), wd, i, 0, 0)
            chunks += [(i, c) for c in range((p.numel() + 1023) // 1024)]
            off += (p.numel() + 3) // 4 * 4
        self._tensors = torch.from_numpy(rec.view(np.uint8).copy()).to(self.dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int32, device=self.dev)
        self._n_chunks = len(chunks)

    def scale(self):
        """The current loss scale as a device scalar (multiply the loss by it before backward)."""
        return self.state[ST_SCALE]

    def set_lr(self, lr):
        self._lr_host[0] = float(lr)
        self.state[ST_LR:ST_LR + 1].copy_(self._lr_host, non_blocking=True)

    def metrics(self):
        s = self.state.tolist()
        return dict(grad_norm=s[ST_GRAD_NORM], loss_scale=s[ST_SCALE], found_inf=bool(s[ST_FOUND_INF]), lr=s[ST_LR], step=int(s[ST_STEP]))

    @torch.no_grad()
    def step(self, arena):
        assert arena.numel() >= self.n_grad and arena.dtype == F32
        check(lib().countr_grad_stats(_p(arena), self.n_grad, _p(self._stats_scratch), _p(self.state), ops._stream()))
        check(lib().countr_adamw_update(_p(self._tensors), len(self.names), _p(self._chunks), self._n_chunks, _p(arena), _p(self._flags),
                                        _p(self.exp_avg), _p(self.exp_avg_sq), _p(self.step_count), _p(self.state), self.betas[0],
                                        self.betas[1], self.eps, self.growth, self.backoff, self.interval, ops._stream()))
        ops._count(3)
        engine().wc.bump(self.params)
