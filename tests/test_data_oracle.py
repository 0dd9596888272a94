"""Density-map synthesis: the CPU oracle against its committed fixture, the kernel-weight helper against
scipy's own, and (GPU) the CUDA kernels against the oracle — bit for bit, it is the same double-precision arithmetic."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D

GOLD = os.path.join(os.path.dirname(__file__), "golden", "density_synth.npz")


def test_oracle_matches_fixture():
    g = np.load(GOLD)
    H, W = (int(v) for v in g["hw"])
    v = D.val_density(g["dots"], H, W)
    assert v.dtype == np.float32 and v.shape == (384, 384)
    assert np.array_equal(v[:48, :48], g["val_crop"]) and np.array_equal(v[-8:, -8:], g["val_corner"])
    assert abs(v.astype(np.float64).sum() - float(g["val_sum"])) < 1e-6
    new_H, new_W, start = (int(x) for x in g["train_meta"])
    t = D.train_density(g["dots"], H, W, new_H, new_W, start)
    assert np.array_equal(t[100:148, 200:248], g["train_crop"])
    assert abs(t.astype(np.float64).sum() - float(g["train_sum"])) < 1e-6
    # one coincident pair and the 60x gain: the count is recovered as sum / 60 (FSC_finetune_cross.py:299)
    assert abs(v.sum() / 60 - (len(g["dots"]) - 1)) < 1e-3


def test_half_kernel_is_scipys():
    from scipy.ndimage import _filters
    from countr_b200.data import gaussian_half_kernel
    for sigma, radius in ((1.0, None), (4.0, 7), (2.5, None)):
        w, r = gaussian_half_kernel(sigma, radius)
        ref = _filters._gaussian_kernel1d(sigma, 0, r)
        assert np.array_equal(w, ref[r:]) and np.array_equal(ref[:r][::-1], ref[r + 1:])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["val", "train"])
def test_density_kernels_match_oracle(cuda, mode):
    from countr_b200.data import density_from_dots
    rng = np.random.default_rng(11)
    sizes = [(480, 640), (333, 1000), (384, 384), (700, 400)]
    B = len(sizes)
    n = [61, 0, 300, 17]
    n_max = max(n)
    dots = np.zeros((B, n_max, 2))
    outs, refs = [], []
    for b, (H, W) in enumerate(sizes):
        d = rng.random((n[b], 2)) * np.array([W, H])
        if n[b] > 3:
            d[1] = d[0]
            d[2] = [W - 1e-9, H - 1e-9]
        dots[b, :n[b]] = d
    for b, (H, W) in enumerate(sizes):     # scale / canvas / window differ per image: one call per image, as the dataset does
        dd = torch.from_numpy(dots[b:b + 1]).to(cuda).contiguous()
        cc = torch.tensor([n[b]], dtype=torch.int32, device=cuda)
        if mode == "val":
            ref = D.val_density(dots[b, :n[b]], H, W)
            got = density_from_dots(dd, cc, scale=(384.0 / H, 384.0 / W), sigma=4.0, radius=7)
        else:
            new_H, new_W = 384, max(384, 16 * int(W * 384 / H / 16))
            start = (new_W - 384) // 3
            ref = D.train_density(dots[b, :n[b]], H, W, new_H, new_W, start)
            got = density_from_dots(dd, cc, scale=(float(new_H) / H, float(new_W) / W), canvas_hw=(new_H, new_W), origin=(0, start),
                                    sigma=1.0)
        torch.cuda.synchronize()
        got = got[0].cpu().numpy()
        assert got.shape == ref.shape
        diff = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()
        assert np.array_equal(got, ref) or diff <= 1e-6 * max(1.0, float(np.abs(ref).max())), (b, diff)
        outs.append(got); refs.append(ref)
    # a batched call (shared geometry) equals the per-image calls
    dd = torch.from_numpy(np.stack([dots[0], dots[0][::-1].copy()])).to(cuda).contiguous()
    cc = torch.tensor([n[0], n_max], dtype=torch.int32, device=cuda)
    both = density_from_dots(dd, cc, scale=(384.0 / 480, 384.0 / 640), sigma=4.0, radius=7)
    one = density_from_dots(dd[:1].contiguous(), cc[:1].contiguous(), scale=(384.0 / 480, 384.0 / 640), sigma=4.0, radius=7)
    assert torch.equal(both[0], one[0])


def test_crop_resize_oracle_is_torchvision_without_antialias():
    """The oracle's restatement equals torchvision's own Resize with antialias disabled (the 0.14.1 behaviour for tensors)."""
    tv = pytest.importorskip("torchvision")
    from torchvision import transforms
    g = torch.Generator().manual_seed(5)
    img = torch.rand(3, 120, 200, generator=g)
    rects = [(10, 20, 57, 90), (0, 0, 119, 199), (30, 40, 33, 44), (100, 150, 119, 199)]
    ref = D.crop_resize_boxes(img, rects)
    for i, (y1, x1, y2, x2) in enumerate(rects):
        tvb = transforms.Resize((64, 64), antialias=False)(img[:, y1:y2 + 1, x1:x2 + 1])
        assert torch.allclose(ref[i], tvb, rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_crop_resize_kernel_matches_oracle(cuda):
    from countr_b200.data import crop_resize_boxes
    g = torch.Generator().manual_seed(6)
    imgs = torch.rand(3, 3, 384, 416, generator=g)
    rects = torch.tensor([[[10, 20, 57, 90], [0, 0, 383, 415], [30, 40, 33, 44]],
                          [[100, 150, 119, 199], [5, 5, 5, 5], [300, 350, 383, 415]],
                          [[0, 0, 63, 63], [17, 3, 200, 9], [200, 100, 210, 300]]], dtype=torch.int32)
    got = crop_resize_boxes(imgs.to(cuda), rects.to(cuda))
    torch.cuda.synchronize()
    assert got.shape == (3, 3, 3, 64, 64)
    for b in range(3):
        ref = D.crop_resize_boxes(imgs[b], [tuple(int(v) for v in r) for r in rects[b]])
        assert torch.allclose(got[b].cpu(), ref, rtol=0, atol=2e-6), (b, (got[b].cpu() - ref).abs().max().item())
    # a non-contiguous view (a 384-wide window of a wider image) gives the same crops as its contiguous copy
    win = imgs.to(cuda)[:, :, :, 16:400]
    r2 = torch.tensor([[[10, 20, 57, 90]], [[0, 0, 383, 383]], [[30, 40, 33, 44]]], dtype=torch.int32, device=cuda)
    assert torch.equal(crop_resize_boxes(win, r2), crop_resize_boxes(win.contiguous(), r2))


def test_host_side_draw_helpers():
    """countr_b200.data host logic (no kernel calls): ColorJitter draws stay inside torchvision's ranges and use every function once;
    the affine matrix composes centre -> scale -> shear -> rotate -> translate -> back."""
    from countr_b200 import data
    g = torch.Generator().manual_seed(0)
    ops_t, fac = data.sample_color_jitter(64, generator=g)
    assert ops_t.shape == (4, 64) and fac.shape == (4, 64)
    assert all(sorted(ops_t[:, b].tolist()) == [0, 1, 2, 3] for b in range(64))
    lo = {0: 0.75, 1: 0.85, 2: 0.85, 3: -0.15}
    hi = {0: 1.25, 1: 1.15, 2: 1.15, 3: 0.15}
    for j in range(4):
        for b in range(64):
            fn = int(ops_t[j, b])
            assert lo[fn] <= float(fac[j, b]) <= hi[fn]
    assert len({tuple(ops_t[:, b].tolist()) for b in range(64)}) > 4             # the order is drawn per image
    H, W = 384, 512
    assert np.allclose(data.affine_matrix(H, W), np.eye(3))
    c = np.array([W / 2 - 0.5, H / 2 - 0.5, 1.0])
    M = data.affine_matrix(H, W, rotate_deg=15, scale=1.2, shear_deg=-10, translate_frac=(0.1, -0.2))
    assert np.allclose(M @ c, c + np.array([round(0.1 * W), round(-0.2 * H), 0.0]))   # the centre only moves by the translation
    R = data.affine_matrix(H, W, rotate_deg=90)
    p = R @ np.array([W / 2 - 0.5 + 10, H / 2 - 0.5, 1.0])                       # +x of the centre -> +y (clockwise on screen, y down)
    assert np.allclose(p[:2], [W / 2 - 0.5, H / 2 - 0.5 + 10])
    S = data.affine_matrix(H, W, scale=0.8)
    assert np.allclose((S @ np.array([W / 2 - 0.5 + 10, H / 2 - 0.5 + 5, 1.0]))[:2], [W / 2 - 0.5 + 8, H / 2 - 0.5 + 4])
