"""GPU input pipeline, first piece: ground-truth density maps from dot annotations.

`density_from_dots` reproduces, for a whole batch on the device, what the reference's dataset transforms compute per image
on the host with numpy + scipy (util/FSC147.py:262-273 for training without augmentation, :326-331 for validation): scatter
a 1 per annotated point on the resized canvas, crop the 384-wide window, `ndimage.gaussian_filter`, multiply by 60.
There is no CPU fallback: the kernels live in libcountr_sm100.so (csrc/data.cu).
"""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import check, lib


def gaussian_half_kernel(sigma, radius=None, truncate=4.0):
    """The float64 weights scipy.ndimage uses (_gaussian_kernel1d, order 0), centre first: w[k] is the tap at distance k."""
    sigma = float(sigma)
    if radius is None:
        radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:].copy(), radius


def density_from_dots(dots, counts, out_hw=(384, 384), scale=(1.0, 1.0), canvas_hw=None, origin=(0, 0), sigma=1.0, radius=None,
                      gain=60.0):
    """dots: float64 [B, n_max, 2] (x, y) on the device, counts: int32 [B]; scale = (scale_factor_h, scale_factor_w);
    canvas_hw = (new_H, new_W) of the resized image (defaults to out_hw); origin = (y0, x0) of the kept window.
    Training (FSC147.py:262-273): sigma=1 (radius 4); validation (:326-331): sigma=4, radius=7.  Returns fp32 [B, H, W]."""
    assert dots.is_cuda and dots.dtype == torch.float64 and dots.dim() == 3 and dots.shape[-1] == 2 and dots.is_contiguous()
    assert counts.is_cuda and counts.dtype == torch.int32 and counts.shape == (dots.shape[0],)
    B, n_max, _ = dots.shape
    H, W = out_hw
    ch, cw = canvas_hw if canvas_hw is not None else out_hw
    w, r = gaussian_half_kernel(sigma, radius)
    wd = torch.from_numpy(w).to(dots.device)
    tmp = torch.empty(B, H, W, dtype=torch.float32, device=dots.device)
    out = torch.empty(B, H, W, dtype=torch.float32, device=dots.device)
    check(lib().countr_density_from_dots(ctypes.c_void_p(dots.data_ptr()), ctypes.c_void_p(counts.data_ptr()), B, n_max,
                                         float(scale[0]), float(scale[1]), int(ch), int(cw), int(origin[0]), int(origin[1]), H, W,
                                         ctypes.c_void_p(wd.data_ptr()), r, float(gain), ctypes.c_void_p(tmp.data_ptr()),
                                         ctypes.c_void_p(out.data_ptr()), ops._stream()))
    return out


def crop_resize_boxes(images, rects, out_hw=64):
    """images: fp32 [B, C, H, W] on the device (any strides); rects: int32 [B, S, 4] = (y1, x1, y2, x2) inclusive, in pixels of
    `images` (FSC147.py:287-296: the scaled exemplar boxes).  Returns fp32 [B, S, C, out_hw, out_hw] — the `boxes` input of
    SupervisedMAE.forward — with torchvision 0.14.1's `transforms.Resize((64, 64))` arithmetic for tensors."""
    assert images.is_cuda and images.dtype == torch.float32 and images.dim() == 4
    assert rects.is_cuda and rects.dtype == torch.int32 and rects.dim() == 3 and rects.shape[0] == images.shape[0] and rects.shape[2] == 4
    rects = rects.contiguous()
    B, C, H, W = images.shape
    S = rects.shape[1]
    out = torch.empty(B, S, C, out_hw, out_hw, dtype=torch.float32, device=images.device)
    sb, sc, sh, sw = images.stride()
    check(lib().countr_crop_resize_boxes(ctypes.c_void_p(images.data_ptr()), sb, sc, sh, sw, ctypes.c_void_p(rects.data_ptr()),
                                         ctypes.c_void_p(out.data_ptr()), B, S, C, H, W, out_hw, ops._stream()))
    return out


# ---------------------------------------------------------------------------------------------- augmentations (util/FSC147.py:133-180)
def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


def augment_noise(images, std=0.1, seed=0):
    """clamp(images + N(0, std), 0, 1)  (util/FSC147.py:133-137: np.random.normal(0, 0.1) + torch.clamp)."""
    assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
    out = torch.empty_like(images)
    check(lib().countr_aug_noise_clamp(_vp(images), _vp(out), images.numel(), float(std), int(seed) & (2 ** 64 - 1), ops._stream()))
    return out


JITTER_OPS = {"brightness": 0, "contrast": 1, "saturation": 2, "hue": 3}


def sample_color_jitter(B, brightness=0.25, contrast=0.15, saturation=0.15, hue=0.15, generator=None):
    """The parameters torchvision.transforms.ColorJitter.get_params draws (a permutation of the four functions and one factor
    each), for B images: (ops int32 [4, B], factors fp32 [4, B]) on the host."""
    ops_t = torch.empty(4, B, dtype=torch.int32)
    fac = torch.empty(4, B, dtype=torch.float32)
    rng = [(1 - brightness, 1 + brightness), (1 - contrast, 1 + contrast), (1 - saturation, 1 + saturation), (-hue, hue)]
    for b in range(B):
        perm = torch.randperm(4, generator=generator)
        for j, fn in enumerate(perm.tolist()):
            lo, hi = rng[fn]
            ops_t[j, b] = fn
            fac[j, b] = float(torch.empty(1).uniform_(lo, hi, generator=generator))
    return ops_t, fac


def color_jitter(images, ops_t, factors):
    """torchvision ColorJitter (util/FSC147.py:372) on fp32 [B, 3, H, W] images, given the sampled order / factors [4, B]."""
    assert images.is_cuda and images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3
    B, _, H, W = images.shape
    out = images.contiguous().clone()
    o = ops_t.to(device=images.device, dtype=torch.int32).contiguous()
    f = factors.to(device=images.device, dtype=torch.float32).contiguous()
    assert o.shape == (4, B) and f.shape == (4, B)
    scratch = torch.empty(B, dtype=torch.float64, device=images.device)
    check(lib().countr_aug_color_jitter(_vp(out), _vp(o), _vp(f), _vp(scratch), B, H, W, ops._stream()))
    return out


def gaussian_blur(images, sigma, kernel_size=(7, 9)):
    """torchvision GaussianBlur(kernel_size=(7, 9)) (util/FSC147.py:373) with one sigma per image (torchvision draws U[0.1, 2.0])."""
    assert images.is_cuda and images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3 and images.is_contiguous()
    B, _, H, W = images.shape
    sg = sigma.to(device=images.device, dtype=torch.float32).contiguous()
    assert sg.shape == (B,)
    tmp, out = torch.empty_like(images), torch.empty_like(images)
    check(lib().countr_aug_gaussian_blur(_vp(images), _vp(tmp), _vp(out), _vp(sg), B, H, W, int(kernel_size[0]), int(kernel_size[1]),
                                         ops._stream()))
    return out


def hflip(x, flags):
    """TF.hflip of x[b] where flags[b] != 0 (util/FSC147.py:176-180); x: fp32 [B, C, H, W] (images) or [B, H, W] (density maps)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() in (3, 4)
    B = x.shape[0]
    planes = x.shape[1] if x.dim() == 4 else 1
    fl = flags.to(device=x.device, dtype=torch.int32).contiguous()
    out = torch.empty_like(x)
    check(lib().countr_aug_hflip(_vp(x), _vp(out), _vp(fl), B, planes, x.shape[-2], x.shape[-1], ops._stream()))
    return out
