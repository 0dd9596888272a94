"""TEST INFRASTRUCTURE — CPU restatement of the reference's density-map synthesis (never imported by the product path).

Follows util/FSC147.py:262-273 (ResizeTrainImage without augmentation: scatter on the resized canvas, crop the 384-wide
window at `start`, `ndimage.gaussian_filter(sigma=(1, 1), order=0)`, * 60) and :326-331 (ResizeValImage: 384 x 384 canvas,
`gaussian_filter(sigma=4, radius=7, order=0)`, * 60) line by line, with the same numpy / scipy calls the reference makes.
util/FSC147.py itself cannot be imported here (imgaug / torchvision transforms pipeline), the arithmetic is all scipy's.
"""
import numpy as np
from scipy import ndimage


def _dot_canvas(dots, H, W, new_H, new_W):
    """One 1 per annotated point on the (new_H x new_W) resized canvas: row = min(new_H-1, int(y * new_H/H)), column likewise
    (FSC147.py:263-265, 327-329).  Assignment, not accumulation: coincident points count once.  float64 products truncated
    toward zero, exactly like Python's int() on the reference's numpy scalars."""
    sh, sw = float(new_H) / H, float(new_W) / W
    canvas = np.zeros((new_H, new_W), dtype=np.float32)
    if len(dots):
        d = np.asarray(dots, dtype=np.float64)
        rows = np.minimum(new_H - 1, np.trunc(d[:, 1] * sh).astype(np.int64))
        cols = np.minimum(new_W - 1, np.trunc(d[:, 0] * sw).astype(np.int64))
        canvas[rows, cols] = 1
    return canvas


def train_density(dots, H, W, new_H, new_W, start, max_hw=384):
    """FSC147.py:262-273 (no-augmentation path).  dots: float64 [n, 2] (x, y) in original-image pixels; (H, W) original size;
    the max_hw-wide window starting at column `start` is filtered with sigma = 1 (scipy's default truncate: radius 4)."""
    window = _dot_canvas(dots, H, W, new_H, new_W)[:max_hw, start:start + max_hw]
    return ndimage.gaussian_filter(window, sigma=(1, 1), order=0) * 60


def val_density(dots, H, W, max_hw=384):
    """FSC147.py:326-331: the whole image resized to max_hw x max_hw, sigma = 4 with an explicit radius of 7."""
    return ndimage.gaussian_filter(_dot_canvas(dots, H, W, max_hw, max_hw), sigma=4, radius=7, order=0) * 60


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
