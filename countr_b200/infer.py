"""Sliding-window evaluation: the reference walks a 384-px window over wide images with
stride 128, one batch-1 forward per window, and blends the windows with a chain of ZeroPad2d allocations
(demo.py:124-160, FSC_test_cross(few-shot).py:322-349).  Here every window of an image goes through ONE batched
forward and the blend is one kernel that replays the same recurrence per pixel."""
import ctypes

import torch

from . import ops
from ._lib import check, lib

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def window_starts(w, win=384, stride=128):
    """Window positions exactly as the reference loop generates them (demo.py:124-160)."""
    starts, start = [], 0
    while start + win - 1 < w:
        starts.append(start)
        start += stride
        if start + win - 1 >= w:
            if start == w - win + stride:
                break
            start = w - win
    return starts


@torch.no_grad()
def sliding_window_density(model, samples, boxes, shot_num, win=384, stride=128):
    """samples [1, 3, 384, W] (W >= 384), boxes [1, K, 3, 64, 64] or empty -> (density [384, W] fp32, count)."""
    _, _, h, w = samples.shape
    starts = window_starts(w, win, stride)
    nw = len(starts)
    batch = torch.stack([samples[0, :, :, s:s + win] for s in starts])                    # [nw, 3, 384, 384]
    bx = boxes.expand(nw, *boxes.shape[1:]) if boxes.dim() == 5 else torch.empty(nw, 0, device=samples.device)
    outs = model(batch, bx, shot_num)                                                       # one batched forward
    density = torch.empty(h, w, dtype=torch.float32, device=samples.device)
    st = torch.tensor(starts, dtype=torch.int32, device=samples.device)
    check(lib().countr_window_blend(ctypes.c_void_p(outs.data_ptr()), _DTYPE_CODE[outs.dtype], ctypes.c_void_p(st.data_ptr()), nw, h, win, w,
                                    ctypes.c_void_p(density.data_ptr()), ops._stream()))
    ops._count()
    return density, density.sum() / 60
