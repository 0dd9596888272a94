"""Sliding-window evaluation: the reference walks a 384-px window over wide images with
stride 128, one batch-1 forward per window, and blends the windows with a chain of ZeroPad2d allocations
(demo.py:124-160, FSC_test_cross(few-shot).py:322-349).  Here every window of an image goes through ONE batched
forward and the blend is one kernel that replays the same recurrence per pixel.

`evaluate_image` is the whole per-image evaluation of FSC_test_cross(few-shot).py:258-359: the 3 x 3 tiling for images
whose exemplars are tiny (every crop blown up to the full frame, all crops x windows in one batched forward), the count,
and the test-time normalisation by the density mass under the exemplar boxes — crop / resize, blend and box sums are
kernels of libcountr_sm100.so, nothing synchronises with the host."""
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


def _blend(outs, starts_t, nw, h, win, w, density):
    check(lib().countr_window_blend(ctypes.c_void_p(outs.data_ptr()), _DTYPE_CODE[outs.dtype], ctypes.c_void_p(starts_t.data_ptr()), nw, h,
                                    win, w, ctypes.c_void_p(density.data_ptr()), ops._stream()))
    ops._count()


def tile_rects(h, w, order="test"):
    """The nine crops (top, left, height, width) in the order the reference builds them: FSC_test_cross(few-shot).py:276-284
    walks column by column, demo.py:86-94 row by row."""
    ch, cw = int(h / 3), int(w / 3)
    tops, lefts = [0, int(h / 3), int(h * 2 / 3)], [0, int(w / 3), int(w * 2 / 3)]
    if order == "test":
        idx = [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (2, 1), (0, 2), (1, 2), (2, 2)]       # (row, col)
    else:
        idx = [(r, c) for r in range(3) for c in range(3)]
    return [(tops[r], lefts[c], ch, cw) for r, c in idx]


def small_exemplar_count(pos, limit=3):
    """FSC_test_cross(few-shot).py:263-271: exemplars (among the first `limit` boxes) smaller than 10 px on both sides;
    demo.py:80-83 looks at every box (limit=None).  pos: iterable of (y1, x1, y2, x2)."""
    s_cnt = 0
    for k, rect in enumerate(pos):
        if limit is not None and k >= limit:
            break
        if rect[2] - rect[0] < 10 and rect[3] - rect[1] < 10:
            s_cnt += 1
    return s_cnt


@torch.no_grad()
def evaluate_image(model, samples, boxes, pos, shot_num=None, max_s_cnt=1, normalization=True, win=384, stride=128, semantics="test"):
    """Per-image evaluation of FSC_test_cross(few-shot).py:258-359 (semantics="test") or demo.py:76-169 (semantics="demo").
    samples [1, 3, 384, W] fp32, boxes [1, K, 3, 64, 64] (or empty), pos: the K exemplar rectangles (y1, x1, y2, x2) in pixels.
    Returns (pred_cnt: 0-d device tensor, density map(s) [384, W] or [9, 384, W] fp32).

    Reference quirk kept: with tiling the normalisation uses the LAST crop's density map (the loop variable the reference reads
    after its loop; the bottom-right crop in both scripts).  demo.py differs from the test script in the crop order (row by row),
    in counting small exemplars over ALL boxes, and in passing the literal shot count 3."""
    dev = samples.device
    _, _, h, w = samples.shape
    K = boxes.shape[1] if boxes.dim() == 5 else 0
    if shot_num is None:
        shot_num = K if semantics == "test" else 3           # :261 num_boxes / demo.py:111 literal 3
    pos = [tuple(int(v) for v in r) for r in pos]
    s_cnt = small_exemplar_count(pos, 3 if semantics == "test" else None)
    starts = window_starts(w, win, stride)
    nw = len(starts)
    st = torch.tensor(starts, dtype=torch.int32, device=dev)
    if s_cnt >= max_s_cnt:
        rects = tile_rects(h, w, "test" if semantics == "test" else "demo")
        rt = torch.tensor([[t, l, t + ch - 1, l + cw - 1] for t, l, ch, cw in rects], dtype=torch.int32, device=dev).view(1, 9, 4)
        img = samples if samples.dtype == torch.float32 else samples.float()
        tiles = torch.empty(1, 9, 3, h, w, dtype=torch.float32, device=dev)
        sb, sc, sh, sw = img.stride()
        check(lib().countr_crop_resize(ctypes.c_void_p(img.data_ptr()), sb, sc, sh, sw, ctypes.c_void_p(rt.data_ptr()),
                                       ctypes.c_void_p(tiles.data_ptr()), 1, 9, 3, h, w, h, w, ops._stream()))
        ops._count()
        tiles = tiles[0]                                                                           # [9, 3, h, w]
        batch = torch.stack([tiles[k, :, :, s:s + win] for k in range(9) for s in starts])      # [9 * nw, 3, 384, 384]
        bx = boxes.expand(9 * nw, *boxes.shape[1:]) if boxes.dim() == 5 else torch.empty(9 * nw, 0, device=dev)
        outs = model(batch, bx, shot_num)                                                         # ONE batched forward
        density = torch.empty(9, h, w, dtype=torch.float32, device=dev)
        for k in range(9):
            _blend(outs[k * nw:(k + 1) * nw], st, nw, h, win, w, density[k])
        pred = density.sum() / 60                    # both scripts add up all nine crops (demo.py:127, few-shot:320)
        last = density[8]
    else:
        batch = torch.stack([samples[0, :, :, s:s + win] for s in starts])
        bx = boxes.expand(nw, *boxes.shape[1:]) if boxes.dim() == 5 else torch.empty(nw, 0, device=dev)
        outs = model(batch, bx, shot_num)
        density = torch.empty(h, w, dtype=torch.float32, device=dev)
        _blend(outs, st, nw, h, win, w, density)
        pred = density.sum() / 60
        last = density
    if normalization and len(pos) > 0:
        e = torch.zeros(1, dtype=torch.float32, device=dev)
        pr = torch.tensor(pos, dtype=torch.int32, device=dev)
        check(lib().countr_rect_mass(ctypes.c_void_p(last.data_ptr()), h, w, ctypes.c_void_p(pr.data_ptr()), len(pos), 60.0,
                                     ctypes.c_void_p(e.data_ptr()), ops._stream()))
        ops._count()
        e_cnt = e[0] / 3
        pred = torch.where(e_cnt > 1.8, pred / e_cnt, pred)
    return pred, density
