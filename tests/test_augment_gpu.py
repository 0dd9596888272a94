"""GPU augmentation kernels (util/FSC147.py:133-180, 371-374) against torchvision's own tensor arithmetic run on the CPU with
the SAME sampled parameters (order / factors / sigma / flip flags), and the noise kernel's distribution."""
import pytest
import torch

pytestmark = pytest.mark.gpu
tvF = pytest.importorskip("torchvision.transforms.functional")


def _images(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, H, W, generator=g)
    x[0, :, : H // 4] = 0.5                       # a grey patch: max == min exercises the hue branch without a dominant channel
    x[-1, 0] = 1.0                                # saturated red channel
    return x


def test_color_jitter_matches_torchvision(cuda):
    from countr_b200 import data
    B, H, W = 5, 96, 130
    x = _images(B, H, W, 1)
    g = torch.Generator().manual_seed(2)
    ops_t, fac = data.sample_color_jitter(B, generator=g)
    fixed = {0: 1.2, 1: 0.88, 2: 1.13, 3: -0.12}       # factor per function for the two hand-picked orders below
    for b, order in ((0, [3, 1, 0, 2]), (1, [1, 3, 2, 0])):      # hue first / contrast after hue are covered for sure
        for j, fn in enumerate(order):
            ops_t[j, b] = fn
            fac[j, b] = fixed[fn]
    got = data.color_jitter(x.to(cuda), ops_t, fac).cpu()
    fns = [tvF.adjust_brightness, tvF.adjust_contrast, tvF.adjust_saturation, tvF.adjust_hue]
    for b in range(B):
        ref = x[b]
        for j in range(4):
            ref = fns[int(ops_t[j, b])](ref, float(fac[j, b]))
        # hue near a sector boundary can pick the neighbouring sector under different rounding: compare robustly
        diff = (got[b] - ref).abs()
        assert diff.mean().item() < 2e-6, (b, diff.mean().item())
        assert (diff > 1e-4).float().mean().item() < 1e-4, (b, diff.max().item())


@pytest.mark.parametrize("H,W", [(384, 384), (100, 77)])
def test_gaussian_blur_matches_torchvision(cuda, H, W):
    from countr_b200 import data
    B = 3
    x = _images(B, H, W, 3)
    sigma = torch.tensor([0.1, 0.9, 2.0])
    got = data.gaussian_blur(x.to(cuda), sigma).cpu()
    for b in range(B):
        ref = tvF.gaussian_blur(x[b], kernel_size=[7, 9], sigma=[float(sigma[b]), float(sigma[b])])
        assert torch.allclose(got[b], ref, rtol=0, atol=2e-6), (b, (got[b] - ref).abs().max().item())


def test_hflip_and_noise(cuda):
    from countr_b200 import data
    x = _images(4, 64, 50, 4).to(cuda)
    flags = torch.tensor([1, 0, 1, 0])
    got = data.hflip(x, flags)
    for b in range(4):
        assert torch.equal(got[b], x[b].flip(-1) if flags[b] else x[b])
    d = torch.rand(4, 64, 50, device=cuda)
    gd = data.hflip(d, flags)
    assert torch.equal(gd[0], d[0].flip(-1)) and torch.equal(gd[1], d[1])
    base = torch.full((2, 3, 384, 384), 0.5, device=cuda)
    n1, n2 = data.augment_noise(base, 0.1, seed=7), data.augment_noise(base, 0.1, seed=8)
    r = (n1 - 0.5).flatten()
    assert abs(r.mean().item()) < 5e-4 and abs(r.std().item() - 0.1) < 5e-4           # N(0, 0.1): 884736 samples, nothing clamps at 5 sigma
    assert not torch.equal(n1, n2) and torch.equal(n1, data.augment_noise(base, 0.1, seed=7))
    k = (((r / 0.1) ** 4).mean()).item()
    assert abs(k - 3.0) < 0.05                                                        # Gaussian kurtosis
    edge = data.augment_noise(torch.zeros(1, 3, 64, 64, device=cuda), 0.1, seed=1)
    assert edge.min().item() == 0.0 and 0.4 < (edge == 0).float().mean().item() < 0.6     # clamp at 0 takes the negative half


def _dots(n, H, W, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)


@pytest.mark.parametrize("bl,self_mosaic", [(10, True), (20, False), (13, False)])
def test_mosaic_matches_reference_recurrence(cuda, bl, self_mosaic):
    """util/FSC147.py:183-262: collage image (seam offsets included) and its dot map / density."""
    import numpy as np
    from countr_b200 import data
    from oracle import data_oracle as D
    g = torch.Generator().manual_seed(bl)
    if self_mosaic:                                                    # >= 70 objects: four crops of the same resized image
        base = torch.rand(3, 384, 683, generator=g)
        images = [base] * 4
        crops = [(0, 0, 384), (10, 200, 150), (100, 299, 284), (33, 480, 203)]
        orig = [(600, 1067)] * 4
    else:                                                              # the sample and three other training images
        images = [torch.rand(3, 384, w, generator=g) for w in (512, 384, 700, 421)]
        crops = [(0, 0, 384), (50, 30, 250), (0, 316, 384), (84, 121, 300)]
        orig = [(768, 1024), (500, 500), (1000, 1823), (384, 421)]
    scales = [(im.shape[1] / oh, im.shape[2] / ow) for im, (oh, ow) in zip(images, orig)]
    dots = [_dots(150 if self_mosaic else 40 + 11 * t, oh, ow, 7 + t) for t, (oh, ow) in enumerate(orig)]
    if self_mosaic:
        dots = [dots[0]] * 4
    same = [True, True, False, True] if not self_mosaic else [True] * 4
    ref_img, ref_dots = D.mosaic(images, crops, bl, dots, scales, same)
    dev_images = [im.to(cuda) for im in images] if not self_mosaic else [images[0].to(cuda)] * 4
    got = data.mosaic(dev_images, crops, bl).cpu()
    assert got.shape == (3, 384, 384)
    assert torch.allclose(got, ref_img, rtol=0, atol=2e-6), (got - ref_img).abs().max().item()
    all_dots = torch.from_numpy(np.concatenate(dots)).to(cuda)
    begins = np.cumsum([0] + [len(d) for d in dots[:-1]])
    ranges = [(int(b), len(d) if s else 0) for b, d, s in zip(begins, dots, same)]
    den = data.mosaic_density(dev_images, crops, bl, all_dots, ranges, scales).cpu().numpy()
    ref_den = D.filter_density(ref_dots.numpy())
    assert ref_dots.sum() > 20
    assert np.allclose(den, ref_den, rtol=0, atol=1e-6), np.abs(den - ref_den).max()
    assert abs(den.sum() / 60 - ref_dots.sum().item()) < 1e-2


def test_mosaic_rejects_a_crop_outside_its_image(cuda):
    from countr_b200 import data
    img = torch.rand(3, 384, 400, device=cuda)
    with pytest.raises(RuntimeError):
        data.mosaic([img] * 4, [(0, 0, 384), (0, 0, 384), (0, 100, 384), (0, 0, 384)], 10)


@pytest.mark.parametrize("params", [dict(rotate_deg=15, scale=1.2, shear_deg=-10, translate_frac=(0.2, -0.2)),
                                    dict(rotate_deg=-7.5, scale=0.8, shear_deg=4, translate_frac=(-0.11, 0.07)),
                                    dict()])
def test_affine_warp_and_keypoints(cuda, params):
    """iaa.Affine (util/FSC147.py:146-171): bilinear zero-border warp and the dot map of the transformed key points."""
    import numpy as np
    from countr_b200 import data
    from oracle import data_oracle as D
    H0, W0, H, W = 600, 900, 384, 576
    img = _images(1, H, W, 11)[0]
    M = data.affine_matrix(H, W, **params)
    got = data.affine_warp(img.to(cuda), M).cpu().numpy()
    ref = D.affine_warp(img.numpy(), M)
    assert np.allclose(got, ref, rtol=0, atol=2e-6), np.abs(got - ref).max()
    if not params:
        assert np.array_equal(got, img.numpy())                        # identity parameters: the image itself
    dots = _dots(300, H0, W0, 5)
    canvas = data.affine_dot_canvas(torch.from_numpy(dots).to(cuda), (H / H0, W / W0), (H, W), M).cpu().numpy()
    ref_canvas = D.affine_dot_canvas(dots, H0, W0, H, W, M)
    assert np.array_equal(canvas, ref_canvas)
    assert 50 < ref_canvas.sum() <= 300
    # the centre of the image moves by the translation only
    c = M @ np.array([W / 2 - 0.5, H / 2 - 0.5, 1.0])
    t = params.get("translate_frac", (0, 0))
    assert abs(c[0] - (W / 2 - 0.5 + round(t[0] * W))) < 1e-9 and abs(c[1] - (H / 2 - 0.5 + round(t[1] * H))) < 1e-9


def test_train_transform_composes_like_the_reference(cuda):
    """ResizeTrainImage.__call__ (util/FSC147.py:117-306), non-mosaic branch, with the draws fixed: torchvision's CPU functions
    for jitter / blur, the oracle for the rest."""
    import numpy as np
    from countr_b200 import data
    from oracle import data_oracle as D
    H0, W0, H, W = 512, 1000, 384, 752
    img = _images(1, H, W, 21)[0]
    dots = _dots(200, H0, W0, 22)
    scale = (H / H0, W / W0)
    rects = torch.tensor([[10, 20, 60, 90], [100, 300, 140, 333], [200, 500, 290, 560]], dtype=torch.int32)
    ops_t = torch.tensor([[1], [0], [2], [3]], dtype=torch.int32)
    fac = torch.tensor([[0.9], [1.15], [1.1], [0.05]])
    draws = dict(mosaic=None, noise_seed=0, jitter=(ops_t, fac), blur_sigma=torch.tensor([1.3]),
                 affine=dict(rotate_deg=9.0, scale=0.93, shear_deg=6.0, translate_frac=(0.05, -0.1)), flip=True, crop=(0, 201))
    out = data.train_transform(img.to(cuda), torch.from_numpy(dots).to(cuda), scale, rects, draws, noise_std=0.0)
    x = torch.clamp(img, 0, 1)
    for fn, f in zip([tvF.adjust_contrast, tvF.adjust_brightness, tvF.adjust_saturation, tvF.adjust_hue], fac[:, 0].tolist()):
        x = fn(x, f)
    x = tvF.gaussian_blur(x, kernel_size=[7, 9], sigma=[1.3, 1.3])
    M = data.affine_matrix(H, W, **draws["affine"])
    x = torch.from_numpy(D.affine_warp(x.numpy(), M)).flip(-1)[:, 0:384, 201:201 + 384]
    canvas = np.ascontiguousarray(D.affine_dot_canvas(dots, H0, W0, H, W, M)[:, ::-1])[0:384, 201:201 + 384]
    diff = (out["image"].cpu() - x).abs()
    assert diff.mean().item() < 2e-6 and (diff > 1e-4).float().mean().item() < 1e-4, (diff.mean().item(), diff.max().item())
    ref_den = D.filter_density(canvas)
    assert np.allclose(out["gt_density"].cpu().numpy(), ref_den, rtol=0, atol=1e-6)
    assert canvas.sum() > 30
    ref_boxes = D.crop_resize_boxes(img, [tuple(int(v) for v in r) for r in rects])
    assert torch.allclose(out["boxes"].cpu(), ref_boxes, rtol=0, atol=2e-6)
    assert out["pos"].numel() == 0 and out["image"].shape == (3, 384, 384)
