"""The evaluation path pinned against the REFERENCE'S OWN script: tests/golden/eval_fewshot.npz holds the per-image counts that
`FSC_test_cross(few-shot).py` itself (its TestData dataset and its main() loop: sliding window, 3 x 3 tiling for tiny exemplars,
test-time normalisation) printed for three synthetic images with the seeded synthetic base-model weights — produced by
scripts/gen_golden_eval.py.  The CPU test drives the oracle (oracle/infer_oracle.py + oracle/countr_oracle.py) over the two
un-tiled images; the GPU test drives countr_b200.infer.evaluate_image over all three."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "eval_fewshot.npz")


def _inputs(g, key):
    """TestData.__getitem__ (FSC_test_cross(few-shot).py:134-181): resized image, 64 x 64 exemplars, scaled boxes."""
    arr = g["img_" + key]
    H, W = arr.shape[:2]
    new_H = 384
    new_W = 16 * int((W / H * 384) / 16)
    sw, sh = float(new_W) / W, float(new_H) / H
    image = D.resize_pil(arr, (new_H, new_W))
    rects = [[int(y1 * sh), int(x1 * sw), int(y2 * sh), int(x2 * sw)] for y1, x1, y2, x2 in g["boxes_" + key].tolist()]
    boxes = D.crop_resize_boxes(image, rects)
    return image[None], boxes[None], rects


def _eval_state_dict():
    cfg = synth.CONFIGS["base"]
    sd = synth.make_state_dict(cfg, seed=0)
    sd["decode_head3.3.bias"] = torch.full_like(sd["decode_head3.3.bias"], 0.5)      # scripts/gen_golden_eval.py:eval_state_dict
    return cfg, sd


def test_evaluation_oracle_matches_the_reference_script():
    from oracle import countr_oracle as O
    from oracle import infer_oracle as IO
    g = np.load(GOLD)
    cfg, sd = _eval_state_dict()
    names = [str(n) for n in g["names"]]
    for key in ("r", "p"):                        # one window / three overlapping windows; both take the normalisation branch
        i = names.index(key + ".png")
        assert not bool(g["tiled"][i])
        samples, boxes, pos = _inputs(g, key)
        with torch.no_grad():
            cnt, _ = IO.evaluate_image(lambda im, bx, s: O.forward(sd, cfg, im.contiguous(), bx, s), samples, boxes, pos)
            raw, _ = IO.evaluate_image(lambda im, bx, s: O.forward(sd, cfg, im.contiguous(), bx, s), samples, boxes, pos, normalization=False)
        ref = float(g["pred_cnt"][i])
        assert abs(cnt - ref) < 2e-4 * abs(ref), (key, cnt, ref)
        assert raw > 5 * cnt                      # the mass under the exemplar boxes did rescale the count (e_cnt > 1.8, :353-359)


@pytest.mark.gpu
def test_evaluate_image_matches_the_reference_script(cuda):
    from countr_b200.infer import evaluate_image, small_exemplar_count
    from test_parity_gpu import build
    g = np.load(GOLD)
    m, _, _ = build("base", 0, cuda)
    with torch.no_grad():
        m.decode_head3[3].bias.fill_(0.5)
    m.eval()
    names = [str(n) for n in g["names"]]
    for key in ("r", "p", "q"):
        i = names.index(key + ".png")
        samples, boxes, pos = _inputs(g, key)
        assert (small_exemplar_count(pos) >= 1) == bool(g["tiled"][i])
        cnt, _ = evaluate_image(m, samples.to(cuda), boxes.to(cuda), pos)
        ref = float(g["pred_cnt"][i])
        print(f"[eval vs reference script] {key}: {cnt.item():.4f} vs {ref:.4f} (tiled: {bool(g['tiled'][i])})")
        assert abs(cnt.item() - ref) < 3e-3 * abs(ref), (key, cnt.item(), ref)


def test_zero_shot_oracle_matches_the_reference_script():
    """`FSC_test_cross(zero-shot).py`: model(window, boxes, 0) over the sliding window, no tiling, no normalisation."""
    from oracle import countr_oracle as O
    from oracle import infer_oracle as IO
    g = np.load(GOLD)
    cfg, sd = _eval_state_dict()
    names = [str(n) for n in g["zs_names"]]
    samples, _, _ = _inputs(g, "p")
    with torch.no_grad():
        den = IO.window_pass(lambda im, bx, s: O.forward(sd, cfg, im.contiguous(), bx, s), samples, torch.empty(1, 0), 0)
    ref = float(g["zs_pred_cnt"][names.index("p.png")])
    assert abs(float(den.sum() / 60) - ref) < 2e-4 * abs(ref)


@pytest.mark.gpu
def test_zero_shot_evaluation_matches_the_reference_script(cuda):
    from countr_b200.infer import sliding_window_density
    from test_parity_gpu import build
    g = np.load(GOLD)
    m, _, _ = build("base", 0, cuda)
    with torch.no_grad():
        m.decode_head3[3].bias.fill_(0.5)
    m.eval()
    names = [str(n) for n in g["zs_names"]]
    for key in ("r", "p", "q"):
        samples, _, _ = _inputs(g, key)
        _, cnt = sliding_window_density(m, samples.to(cuda), torch.empty(1, 0, device=cuda), 0)
        ref = float(g["zs_pred_cnt"][names.index(key + ".png")])
        print(f"[zero-shot vs reference script] {key}: {cnt.item():.3f} vs {ref:.3f}")
        assert abs(cnt.item() - ref) < 3e-3 * abs(ref), (key, cnt.item(), ref)


DEMO_GOLD = os.path.join(os.path.dirname(__file__), "golden", "demo_counts.npz")
DEMO_BBOXES = [[[136, 98], [173, 127]], [[209, 125], [242, 150]], [[212, 168], [258, 200]]]      # demo.py:52-56, (x, y) corners


def _demo_inputs(h, w, seed):
    """demo.py:34-73 (load_image) on the synthetic image the generator used."""
    arr = synth.smooth_image(h, w, seed)
    new_H = 384
    new_W = 16 * int((w / h * 384) / 16)
    sh, sw = float(new_H) / h, float(new_W) / w
    image = D.resize_pil(arr, (new_H, new_W))
    rects = [[int(b[0][1] * sh), int(b[0][0] * sw), int(b[1][1] * sh), int(b[1][0] * sw)] for b in DEMO_BBOXES]
    return image[None], D.crop_resize_boxes(image, rects)[None], rects


def test_demo_oracle_matches_demo_py():
    """demo.py executed by scripts/gen_golden_demo.py (plain sliding-window case) against the oracle's demo mode."""
    from oracle import countr_oracle as O
    from oracle import infer_oracle as IO
    g = np.load(DEMO_GOLD)
    cfg, sd = _eval_state_dict()
    h, w, seed = (int(v) for v in g["plain_hw_seed"])
    samples, boxes, pos = _demo_inputs(h, w, seed)
    with torch.no_grad():
        cnt, maps = IO.evaluate_image(lambda im, bx, s: O.forward(sd, cfg, im.contiguous(), bx, s), samples, boxes, pos, demo=True)
    assert len(maps) == 1
    ref = float(g["plain_count"])
    assert abs(cnt - ref) < 2e-4 * abs(ref), (cnt, ref)


@pytest.mark.gpu
def test_evaluate_image_demo_semantics_match_demo_py(cuda):
    from countr_b200.infer import evaluate_image
    from test_parity_gpu import build
    g = np.load(DEMO_GOLD)
    m, _, _ = build("base", 0, cuda)
    with torch.no_grad():
        m.decode_head3[3].bias.fill_(0.5)
    m.eval()
    for name, tiled in (("plain", False), ("tiled", True)):
        h, w, seed = (int(v) for v in g[name + "_hw_seed"])
        samples, boxes, pos = _demo_inputs(h, w, seed)
        cnt, dens = evaluate_image(m, samples.to(cuda), boxes.to(cuda), pos, semantics="demo")
        assert (dens.dim() == 3) == tiled
        ref = float(g[name + "_count"])
        print(f"[demo.py] {name}: {cnt.item():.4f} vs {ref:.4f}")
        assert abs(cnt.item() - ref) < 3e-3 * abs(ref), (name, cnt.item(), ref)


@pytest.mark.parametrize("w", [384, 400, 511, 512, 513, 576, 640, 700, 768, 896, 1000, 1152])
def test_window_starts_follow_the_reference_loop(w):
    """countr_b200.infer.window_starts against the window positions the reference's while-loop visits (:323-349), read back from the
    oracle's literal restatement through a probe image whose pixel value is its column index."""
    from countr_b200.infer import window_starts
    from oracle import infer_oracle as IO
    probe = torch.arange(w, dtype=torch.float32).expand(1, 3, 384, w).contiguous()
    seen = []

    def fake_forward(im, bx, s):
        seen.append(int(im[0, 0, 0, 0].item()))
        return torch.zeros(1, 384, 384)

    IO.window_pass(fake_forward, probe, torch.empty(1, 0), 0)
    assert window_starts(w) == seen


@pytest.mark.parametrize("demo", [False, True])
def test_tile_rects_follow_the_reference_crop_order(demo):
    """countr_b200.infer.tile_rects against the order in which the scripts build their nine crops (few-shot :276-284 column by column,
    demo.py:86-94 row by row): the probe's pixel value encodes (row, column), the blown-up crop keeps its top-left pixel."""
    from countr_b200.infer import tile_rects
    from oracle import infer_oracle as IO
    h, w = 384, 500
    probe = (torch.arange(h, dtype=torch.float32)[:, None] * 1000 + torch.arange(w, dtype=torch.float32)[None, :]).expand(1, 3, h, w).contiguous()
    seen = []

    def fake_forward(im, bx, s):
        if im.shape[-1] == 384 and len(seen) < 9 * 2:
            seen.append(float(im[0, 0, 0, 0].item()))
        return torch.zeros(1, 384, 384)

    tiny = [(10, 10, 15, 16), (50, 60, 120, 130), (200, 210, 300, 310)]           # one exemplar under 10 px: tiling
    IO.evaluate_image(fake_forward, probe, torch.zeros(1, 3, 3, 64, 64), tiny, demo=demo)
    firsts = seen[::2]                                                               # w = 500: two windows per crop, the first starts at column 0
    rects = tile_rects(h, w, "demo" if demo else "test")
    assert [(int(v // 1000), int(v % 1000)) for v in firsts] == [(t, l) for t, l, _, _ in rects]
    assert all((ch, cw) == (int(h / 3), int(w / 3)) for _, _, ch, cw in rects)
