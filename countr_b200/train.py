"""FineTuner — the fine-tune step of FSC_finetune_cross.py (:265-319) as one kernel schedule.

    tuner = FineTuner(model, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95), loss_scale=4096.)
    loss = tuner.step(imgs, boxes, gt_density, mask, shot_num)      # device scalar tensor

Same arithmetic as the reference loop — frozen-encoder forward, decoder forward, masked-MSE loss, decoder
backward, (gradient all-reduce), unscale, AdamW with the timm `add_weight_decay` grouping — but without
autograd, GradScaler bookkeeping or per-parameter optimizer launches: gradients and both Adam moments live in
flat fp32 arenas, the loss (+ its gradient) is one kernel and the whole optimizer update is one kernel.
Nothing here synchronises with the host, so a step can be captured in a CUDA graph.
"""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import check, lib
from .backward import decoder_backward
from .dist import build_grad_arena
from .engine import F32, engine

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class FineTuner:
    def __init__(self, model, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95), eps=1e-8, loss_scale=4096.0, allreduce=None):
        self.model, self.lr, self.wd, self.betas, self.eps = model, lr, weight_decay, betas, eps
        self.loss_scale = float(loss_scale)
        self.allreduce = allreduce          # callable(arena) or None
        self.dev = next(model.parameters()).device
        self._opt = {}                      # per shot-mode (shot_num > 0 / == 0): the parameter set differs
        self.loss = torch.zeros((), dtype=F32, device=self.dev)

    def _optimizer_tables(self, shot_num):
        key = shot_num > 0
        if key in self._opt:
            return self._opt[key]
        names, params = self.model._decoder_params(shot_num)
        # Adam moments are kept per PARAMETER NAME in one arena covering every decoder parameter, so switching
        # between few-shot and zero-shot steps keeps each parameter's history (like the reference's single AdamW).
        if not hasattr(self, "_moment_index"):
            all_names = [n for n, p in self.model.named_parameters()
                         if p.requires_grad and not n.startswith(("patch_embed.", "blocks.", "norm."))]
            all_params = dict(self.model.named_parameters())
            arena, views = build_grad_arena(all_names, [all_params[n] for n in all_names], self.dev)
            self.exp_avg = torch.zeros_like(arena)
            self.exp_avg_sq = torch.zeros_like(arena)
            base = arena.data_ptr()
            self._moment_index = {n: (views[n].data_ptr() - base) // 4 for n in all_names}
            self._step_index = {n: i for i, n in enumerate(all_names)}
            self.step_count = torch.zeros(len(all_names), dtype=F32, device=self.dev)   # one counter per parameter, like torch.optim
        g_arena, g_views = build_grad_arena(names, params, self.dev)
        base = g_arena.data_ptr()
        rec = np.zeros(len(names), dtype=np.dtype([("param", "<u8"), ("goff", "<i8"), ("moff", "<i8"), ("numel", "<i8"), ("wd", "<f4"), ("step_idx", "<i4")]))
        chunks = []
        for i, (n, p) in enumerate(zip(names, params)):
            rec[i] = (p.data_ptr(), (g_views[n].data_ptr() - base) // 4, self._moment_index[n], p.numel(),
                      0.0 if (p.ndim == 1 or n.endswith(".bias")) else self.wd, self._step_index[n])     # timm add_weight_decay
            chunks += [(i, c) for c in range((p.numel() + 1023) // 1024)]
        t = dict(names=names, params=params, tensors=torch.from_numpy(rec.view(np.uint8).copy()).to(self.dev),
                 chunks=torch.tensor(chunks, dtype=torch.int32, device=self.dev), n_chunks=len(chunks))
        self._opt[key] = t
        return t

    @torch.no_grad()
    def forward_backward(self, imgs, boxes, gt_density, mask, shot_num):
        """Forward, loss and backward; leaves the (scaled) gradients in `self.arena`.  Returns the loss (device scalar)."""
        m, eng = self.model, engine()
        B = imgs.shape[0]
        on_side = eng.refresh_decoder_weights(m, shot_num, True, imgs.device)    # all stale 16-bit weight copies, one launch
        pre = eng.exemplar_async(m, boxes, shot_num, train=True) if (shot_num > 0 and eng.overlap_exemplar) else None
        _, lat16 = eng.encoder_forward(m, imgs)
        if on_side is not None:
            torch.cuda.current_stream().wait_event(on_side)
        save = {}
        out = eng.decoder_forward(m, lat16, boxes, shot_num, B, F32, save=save, pre=pre)
        dout = torch.empty_like(out)
        check(lib().countr_masked_mse(ctypes.c_void_p(out.data_ptr()), _DTYPE_CODE[out.dtype], ctypes.c_void_p(gt_density.data_ptr()),
                                      ctypes.c_void_p(mask.data_ptr()), ctypes.c_void_p(self.loss.data_ptr()),
                                      ctypes.c_void_p(dout.data_ptr()), B, out.shape[1], out.shape[2], self.loss_scale, ops._stream()))
        ops._count()
        decoder_backward(eng, m, save, boxes, dout)
        self.arena = eng.last_arena
        self._shot = shot_num
        return self.loss

    @torch.no_grad()
    def update(self):
        """Unscale + AdamW on the arena left by forward_backward (all-reduce it first when data-parallel)."""
        eng = engine()
        t = self._optimizer_tables(self._shot)
        check(lib().countr_adamw_step(ctypes.c_void_p(t["tensors"].data_ptr()), len(t["names"]), ctypes.c_void_p(t["chunks"].data_ptr()),
                                      t["n_chunks"], ctypes.c_void_p(self.arena.data_ptr()), ctypes.c_void_p(self.exp_avg.data_ptr()),
                                      ctypes.c_void_p(self.exp_avg_sq.data_ptr()), ctypes.c_void_p(self.step_count.data_ptr()),
                                      self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.loss_scale, ops._stream()))
        ops._count(2)
        eng.wc.bump(t["params"])      # parameters changed behind torch's back: refresh their fp16 copies next step

    def step(self, imgs, boxes, gt_density, mask, shot_num):
        loss = self.forward_backward(imgs, boxes, gt_density, mask, shot_num)
        if self.allreduce is not None:
            self.allreduce(self.arena)
        self.update()
        return loss
