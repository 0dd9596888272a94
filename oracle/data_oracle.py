"""TEST INFRASTRUCTURE — CPU restatement of the reference's density-map synthesis (never imported by the product path).

Follows util/FSC147.py:262-273 (ResizeTrainImage without augmentation: scatter on the resized canvas, crop the 384-wide
window at `start`, `ndimage.gaussian_filter(sigma=(1, 1), order=0)`, * 60) and :326-331 (ResizeValImage: 384 x 384 canvas,
`gaussian_filter(sigma=4, radius=7, order=0)`, * 60) line by line, with the same numpy / scipy calls the reference makes.
util/FSC147.py itself cannot be imported here (imgaug / torchvision transforms pipeline), the arithmetic is all scipy's.
"""
import numpy as np
from scipy import ndimage


def train_density(dots, H, W, new_H, new_W, start, max_hw=384):
    """FSC147.py:262-273.  dots: float64 [n, 2] (x, y) in original-image pixels; (H, W) original size."""
    scale_factor_h = float(new_H) / H
    scale_factor_w = float(new_W) / W
    resized_density = np.zeros((new_H, new_W), dtype='float32')
    for i in range(dots.shape[0]):
        resized_density[min(new_H - 1, int(dots[i][1] * scale_factor_h))][min(new_W - 1, int(dots[i][0] * scale_factor_w))] = 1
    reresized_density = resized_density[0:max_hw, start:start + max_hw]
    reresized_density = ndimage.gaussian_filter(reresized_density, sigma=(1, 1), order=0)
    return reresized_density * 60


def val_density(dots, H, W, max_hw=384):
    """FSC147.py:326-331."""
    new_H = new_W = max_hw
    scale_factor_h = float(new_H) / H
    scale_factor_w = float(new_W) / W
    resized_density = np.zeros((new_H, new_W), dtype='float32')
    for i in range(dots.shape[0]):
        resized_density[min(new_H - 1, int(dots[i][1] * scale_factor_h))][min(new_W - 1, int(dots[i][0] * scale_factor_w))] = 1
    resized_density = ndimage.gaussian_filter(resized_density, sigma=4, radius=7, order=0)
    return resized_density * 60


def crop_resize_boxes(resized_image, rects, out_hw=64):
    """FSC147.py:285-298 / 343-351: `bbox = resized_image[:, y1:y2 + 1, x1:x2 + 1]; bbox = transforms.Resize((64, 64))(bbox)`.
    With the reference's pin (torchvision==0.14.1, requirements.txt:4) `Resize` on a float tensor is
    torch.nn.functional.interpolate(mode="bilinear", align_corners=False) without antialiasing
    (torchvision/transforms/functional_tensor.py:resize, antialias=None -> False); torchvision >= 0.17 would antialias.
    resized_image: float32 torch tensor [C, H, W]; rects: iterable of (y1, x1, y2, x2).  Returns [S, C, 64, 64]."""
    import torch
    import torch.nn.functional as F
    boxes = []
    for y1, x1, y2, x2 in rects:
        bbox = resized_image[:, y1:y2 + 1, x1:x2 + 1]
        boxes.append(F.interpolate(bbox[None], size=(out_hw, out_hw), mode="bilinear", align_corners=False)[0])
    return torch.stack(boxes)
