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


# ---------------------------------------------------------------------------------------------- mosaic (util/FSC147.py:183-262)
class MosaicSrc(ctypes.Structure):
    """Mirror of `countr_mosaic_src` (include/countr_b200.h)."""

    _fields_ = [("img", ctypes.c_void_p), ("sc", ctypes.c_int64), ("sh", ctypes.c_int64), ("sw", ctypes.c_int64),
                ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("top", ctypes.c_int32), ("left", ctypes.c_int32),
                ("length", ctypes.c_int32), ("dot_begin", ctypes.c_int32), ("dot_count", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("scale_h", ctypes.c_double), ("scale_w", ctypes.c_double)]


def _mosaic_srcs(images, crops, dot_ranges=None, scales=None):
    arr = (MosaicSrc * 4)()
    for t in range(4):
        img = images[t]
        assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 3
        top, left, length = (int(v) for v in crops[t])
        begin, count = (0, 0) if dot_ranges is None else (int(v) for v in dot_ranges[t])
        sh_, sw_ = (1.0, 1.0) if scales is None else (float(v) for v in scales[t])
        arr[t] = MosaicSrc(img.data_ptr(), img.stride(0), img.stride(1), img.stride(2), img.shape[1], img.shape[2], top, left, length,
                           begin, count, 0, sh_, sw_)
    return arr


def mosaic(images, crops, blending_l):
    """The 2 x 2 collage of util/FSC147.py:183-262.  images: four fp32 [C, H_t, W_t] device tensors (the resized image four times
    for >= 70 objects, :187-199; one copy of it and three other training images otherwise, :207-236) in the reference's
    `image_array` order; crops[t] = (start_H, start_W, length); blending_l in 10..20.  Returns fp32 [C, 384, 384]."""
    bl = int(blending_l)
    rl = 192 + 2 * bl
    C = images[0].shape[0]
    assert all(im.shape[0] == C for im in images)
    out = torch.empty(C, 384, 384, dtype=torch.float32, device=images[0].device)
    check(lib().countr_aug_mosaic(_mosaic_srcs(images, crops), rl, bl, C, _vp(out), ops._stream()))
    ops._count()
    return out


def mosaic_density(images, crops, blending_l, dots, dot_ranges, scales, sigma=1.0, gain=60.0):
    """Ground truth of the collage: the per-quadrant dot maps (:190-196, :228-232) assembled like the image (:240, :248, :256),
    then gaussian_filter(sigma=1) * 60 (:265-269).  dots: float64 [n, 2] (x, y) on the device, the four quadrants' points
    concatenated; dot_ranges[t] = (begin, count) — count 0 for a quadrant whose image is of another class (:228);
    scales[t] = (Tscale_factor_h, Tscale_factor_w).  Returns fp32 [384, 384]."""
    bl = int(blending_l)
    rl = 192 + 2 * bl
    dev = images[0].device
    assert dots.is_cuda and dots.dtype == torch.float64 and dots.is_contiguous()
    canvas = torch.empty(384, 384, dtype=torch.float32, device=dev)
    check(lib().countr_aug_mosaic_dots(_mosaic_srcs(images, crops, dot_ranges, scales), rl, bl, _vp(dots), _vp(canvas), ops._stream()))
    ops._count()
    return density_filter(canvas[None], sigma=sigma, gain=gain)[0]


def density_filter(canvas, sigma=1.0, radius=None, gain=60.0):
    """ndimage.gaussian_filter(canvas, sigma) * gain on fp32 [B, H, W] dot maps (util/FSC147.py:265-269)."""
    assert canvas.is_cuda and canvas.dtype == torch.float32 and canvas.dim() == 3 and canvas.is_contiguous()
    B, H, W = canvas.shape
    w, r = gaussian_half_kernel(sigma, radius)
    wd = torch.from_numpy(w).to(canvas.device)
    tmp, out = torch.empty_like(canvas), torch.empty_like(canvas)
    check(lib().countr_density_filter(_vp(canvas), _vp(tmp), _vp(out), B, H, W, _vp(wd), r, float(gain), ops._stream()))
    ops._count(2)
    return out


# ---------------------------------------------------------------------------------------------- affine (util/FSC147.py:146-171)
def affine_matrix(h, w, rotate_deg=0.0, scale=1.0, shear_deg=0.0, translate_frac=(0.0, 0.0)):
    """Forward 3 x 3 matrix (input pixel -> output pixel, (x, y, 1) columns) of iaa.Affine(rotate, scale, shear,
    translate_percent) as imgaug 0.4.0 composes it: centre on (w/2 - 0.5, h/2 - 0.5), scale, shear along x (sign flipped),
    rotate, translate by the rounded pixel offset, move back.  RESTATED from imgaug's published behaviour; imgaug is not
    installable here, so this composition is not pinned against the library."""
    sx, sy = w / 2.0 - 0.5, h / 2.0 - 0.5
    rot, shr = np.deg2rad(rotate_deg), np.deg2rad(shear_deg)
    tx, ty = float(np.round(translate_frac[0] * w)), float(np.round(translate_frac[1] * h))

    def T(x, y):
        return np.array([[1, 0, x], [0, 1, y], [0, 0, 1]], dtype=np.float64)

    S = np.diag([scale, scale, 1.0])
    Sh = np.array([[1, np.tan(-shr), 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    R = np.array([[np.cos(rot), -np.sin(rot), 0], [np.sin(rot), np.cos(rot), 0], [0, 0, 1]], dtype=np.float64)
    return T(sx, sy) @ T(tx, ty) @ R @ Sh @ S @ T(-sx, -sy)


def affine_warp(image, matrix):
    """image fp32 [C, H, W] on the device warped by the forward `matrix` (3 x 3 or 2 x 3): bilinear, zero outside."""
    assert image.is_cuda and image.dtype == torch.float32 and image.dim() == 3 and image.is_contiguous()
    m = np.eye(3)
    m[:2] = np.asarray(matrix, dtype=np.float64)[:2]
    inv = np.ascontiguousarray(np.linalg.inv(m)[:2].reshape(6))
    out = torch.empty_like(image)
    C, H, W = image.shape
    check(lib().countr_aug_affine(_vp(image), _vp(out), C, H, W, inv.ctypes.data_as(ctypes.c_void_p), ops._stream()))
    ops._count()
    return out


def affine_dot_canvas(dots, scale, canvas_hw, matrix):
    """Dot map after the affine transform (:146-149, :162-166).  dots: float64 [n, 2] (x, y) on the device in original-image
    pixels, scale = (scale_factor_h, scale_factor_w), canvas_hw = (new_H, new_W).  Returns fp32 [new_H, new_W] of 0 / 1."""
    assert dots.is_cuda and dots.dtype == torch.float64 and dots.is_contiguous() and dots.dim() == 2 and dots.shape[1] == 2
    H, W = canvas_hw
    fwd = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64)[:2].reshape(6))
    canvas = torch.empty(H, W, dtype=torch.float32, device=dots.device)
    check(lib().countr_aug_affine_dots(_vp(dots), dots.shape[0], float(scale[0]), float(scale[1]), H, W,
                                       fwd.ctypes.data_as(ctypes.c_void_p), _vp(canvas), ops._stream()))
    ops._count()
    return canvas


# ---------------------------------------------------------------------------------------------- the whole training transform
def train_transform(resized_image, dots, scale, box_rects, draws, noise_std=0.1):
    """ResizeTrainImage.__call__ with do_aug (util/FSC147.py:117-306) for one sample, every pixel operation on the device; the
    random draws the reference makes along the way are passed in `draws`:
      mosaic: None, or dict(images, crops, blending_l, dots, dot_ranges, scales) — see `mosaic` / `mosaic_density` (:183-262);
      otherwise noise_seed, jitter=(ops, factors), blur_sigma (tensor [1]), affine=dict(rotate_deg, scale, shear_deg,
      translate_frac), flip (bool), crop=(start_H, start_W) (:133-180, :263-267).
    resized_image: fp32 [3, new_H, new_W]; dots: float64 [n, 2] (x, y) in original pixels; scale = (scale_factor_h,
    scale_factor_w); box_rects: int32 [S, 4] = the scaled exemplar boxes (y1, x1, y2, x2) (:279-285).
    Returns the reference's sample dict: image [3, 384, 384], boxes [S, 3, 64, 64] (cropped from the UN-augmented image, as the
    reference does, :286-288), gt_density [384, 384], pos (empty, :292)."""
    assert resized_image.is_cuda and resized_image.dtype == torch.float32 and resized_image.dim() == 3
    dev = resized_image.device
    _, H, W = resized_image.shape
    m = draws.get("mosaic")
    if m is not None:
        image = mosaic(m["images"], m["crops"], m["blending_l"])
        density = mosaic_density(m["images"], m["crops"], m["blending_l"], m["dots"], m["dot_ranges"], m["scales"])
    else:
        x = augment_noise(resized_image[None].contiguous(), std=noise_std, seed=draws["noise_seed"])
        x = color_jitter(x, *draws["jitter"])
        x = gaussian_blur(x, draws["blur_sigma"])
        M = affine_matrix(H, W, **draws["affine"])
        x = affine_warp(x[0], M)
        canvas = affine_dot_canvas(dots, scale, (H, W), M)
        flip = torch.tensor([1 if draws["flip"] else 0], dtype=torch.int32)
        x = hflip(x[None], flip)[0]
        canvas = hflip(canvas[None], flip)[0]
        top, left = draws["crop"]
        image = x[:, top:top + 384, left:left + 384].contiguous()
        density = density_filter(canvas[top:top + 384, left:left + 384].contiguous()[None])[0]
    boxes = crop_resize_boxes(resized_image[None], box_rects.to(device=dev, dtype=torch.int32)[None])[0]
    return {"image": image, "boxes": boxes, "pos": torch.empty(0, device=dev), "gt_density": density}
