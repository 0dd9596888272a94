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
