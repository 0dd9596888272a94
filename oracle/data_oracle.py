"""TEST INFRASTRUCTURE — CPU restatement of the reference's density-map synthesis (never imported by the product path).

Follows util/FSC147.py:262-273 (ResizeTrainImage without augmentation: scatter on the resized canvas, crop the 384-wide
window at `start`, `ndimage.gaussian_filter(sigma=(1, 1), order=0)`, * 60) and :326-331 (ResizeValImage: 384 x 384 canvas,
`gaussian_filter(sigma=4, radius=7, order=0)`, * 60) line by line, with the same numpy / scipy calls the reference makes.
Pinned against the reference itself: scripts/gen_golden_aug.py runs util/FSC147.py's own ResizeTrainImage / ResizeValImage
(stand-ins for the absent cv2 / imgaug) and tests/test_reference_transforms.py checks these functions bit for bit against its
outputs (tests/golden/fsc147_transforms.npz).  The affine step (imgaug) is the one piece that stays unpinned.
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


# ---------------------------------------------------------------------------------------------- mosaic (util/FSC147.py:183-262)
def _resize(t, size):
    """transforms.Resize((size, size)) on a float tensor under the reference's torchvision pin: bilinear, no antialias."""
    import torch.nn.functional as F
    return F.interpolate(t[None], size=(size, size), mode="bilinear", align_corners=False)[0]


def mosaic(images, crops, blending_l, dots=None, scales=None, same_class=None):
    """util/FSC147.py:183-262 with the random draws passed in.  images[t]: float32 torch [C, H_t, W_t] (the four `r_image`s /
    `resized_image`), crops[t] = (start_H, start_W, length), dots[t]: float64 array [n_t, 2] (x, y) in original pixels of image
    t, scales[t] = (scale_factor_h, scale_factor_w), same_class[t]: whether image t's class equals the sample's (:228).
    Returns (image [C, 384, 384], dot map [384, 384]) — the statements below are the reference's, in its order."""
    import torch
    bl = blending_l
    resize_l = 192 + 2 * bl
    image_array, map_array = [], []
    for t in range(4):
        start_H, start_W, length = crops[t]
        new_TH, new_TW = images[t].shape[1:]
        tile = _resize(images[t][:, start_H:start_H + length, start_W:start_W + length], resize_l)          # :225-226
        dmap = np.zeros((resize_l, resize_l), dtype="float32")
        if dots is not None and (same_class is None or same_class[t]):
            sfh, sfw = scales[t]
            for i in range(len(dots[t])):                                                                   # :229-231
                py = min(new_TH - 1, int(dots[t][i][1] * sfh))
                px = min(new_TW - 1, int(dots[t][i][0] * sfw))
                if start_H <= py < start_H + length and start_W <= px < start_W + length:
                    dmap[min(resize_l - 1, int((py - start_H) * resize_l / length))][min(resize_l - 1, int((px - start_W) * resize_l / length))] = 1
        image_array.append(tile)
        map_array.append(torch.from_numpy(dmap))

    def stack_vertical(a, b, ma, mb):                                                                       # :239-245 / :247-253
        img = torch.cat((a[:, bl:resize_l - bl], b[:, bl:resize_l - bl]), 1)
        den = torch.cat((ma[bl:resize_l - bl], mb[bl:resize_l - bl]), 0)
        for i in range(bl):
            img[:, 192 + i] = a[:, resize_l - 1 - bl + i] * (bl - i) / (2 * bl) + img[:, 192 + i] * (i + bl) / (2 * bl)
            img[:, 191 - i] = b[:, bl - i] * (bl - i) / (2 * bl) + img[:, 191 - i] * (i + bl) / (2 * bl)
        return torch.clamp(img, 0, 1), den

    img5, den5 = stack_vertical(image_array[0], image_array[1], map_array[0], map_array[1])
    img6, den6 = stack_vertical(image_array[2], image_array[3], map_array[2], map_array[3])
    img = torch.cat((img5[:, :, bl:resize_l - bl], img6[:, :, bl:resize_l - bl]), 2)                        # :255-261
    den = torch.cat((den5[:, bl:resize_l - bl], den6[:, bl:resize_l - bl]), 1)
    for i in range(bl):
        img[:, :, 192 + i] = img5[:, :, resize_l - 1 - bl + i] * (bl - i) / (2 * bl) + img[:, :, 192 + i] * (i + bl) / (2 * bl)
        img[:, :, 191 - i] = img6[:, :, bl - i] * (bl - i) / (2 * bl) + img[:, :, 191 - i] * (i + bl) / (2 * bl)
    return torch.clamp(img, 0, 1), den


def filter_density(dot_map, sigma=1):
    """util/FSC147.py:265-269."""
    return ndimage.gaussian_filter(np.asarray(dot_map, dtype=np.float32), sigma=(sigma, sigma), order=0) * 60


# ---------------------------------------------------------------------------------------------- affine (util/FSC147.py:146-171)
def affine_warp(image, matrix):
    """Bilinear (order 1) warp with a zero border by the forward 3 x 3 `matrix`: out(x, y) = image(M^-1 (x, y)).  imgaug / cv2 are
    absent here — PARITY UNPINNED for this function: it states the sampling rule the CUDA kernel is checked against, not cv2's
    fixed-point arithmetic.  image: float32 numpy [C, H, W]."""
    C, H, W = image.shape
    inv = np.linalg.inv(np.asarray(matrix, dtype=np.float64))[:2].reshape(6)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    sx = inv[0] * xs + inv[1] * ys + inv[2]
    sy = inv[3] * xs + inv[4] * ys + inv[5]
    fx, fy = np.floor(sx), np.floor(sy)
    wx, wy = (sx - fx).astype(np.float32), (sy - fy).astype(np.float32)
    x0, y0 = fx.astype(np.int64), fy.astype(np.int64)
    pad = np.zeros((C, H + 2, W + 2), dtype=np.float32)
    pad[:, 1:-1, 1:-1] = image

    def tap(yy, xx):
        ok = (yy >= -1) & (yy <= H) & (xx >= -1) & (xx <= W)
        return np.where(ok, pad[:, np.clip(yy + 1, 0, H + 1), np.clip(xx + 1, 0, W + 1)], np.float32(0))

    one = np.float32(1)
    top = (one - wx) * tap(y0, x0) + wx * tap(y0, x0 + 1)
    bot = (one - wx) * tap(y0 + 1, x0) + wx * tap(y0 + 1, x0 + 1)
    return ((one - wy) * top + wy * bot).astype(np.float32)


def affine_dot_canvas(dots, H, W, new_H, new_W, matrix):
    """util/FSC147.py:146-149 (key points at the truncated, clamped resized coordinates) and :162-166 (dot map of the transformed
    points that stay inside the image), the transform being x' = M (x, y, 1)."""
    sfh, sfw = float(new_H) / H, float(new_W) / W
    m = np.asarray(matrix, dtype=np.float64)
    canvas = np.zeros((new_H, new_W), dtype="float32")
    for i in range(len(dots)):
        kx, ky = min(new_W - 1, int(dots[i][0] * sfw)), min(new_H - 1, int(dots[i][1] * sfh))
        ax = m[0, 0] * kx + m[0, 1] * ky + m[0, 2]
        ay = m[1, 0] * kx + m[1, 1] * ky + m[1, 2]
        out_of_image = not (0 <= ax < new_W and 0 <= ay < new_H)
        if int(ay) <= new_H - 1 and int(ax) <= new_W - 1 and not out_of_image:
            canvas[int(ay)][int(ax)] = 1
    return canvas


# ---------------------------------------------------------------------------------------------- host-side resize (util/FSC147.py:102-126)
def flex_resize(h, w, max_hw=384):
    """ResizeTrainImage.flex_resize (util/FSC147.py:102-115): the smaller side to max_hw, or both rounded down to multiples of 16."""
    if h < max_hw <= w or h <= w < max_hw:
        new_h = max_hw
        new_w = round(w * new_h / h)
    elif w < max_hw <= h or w < h < max_hw:
        new_w = max_hw
        new_h = round(h * new_w / w)
    else:
        new_w = 16 * int(w / 16)
        new_h = 16 * int(h / 16)
    return new_h, new_w


def resize_pil(arr_u8, new_hw):
    """`TTensor(transforms.Resize((new_H, new_W))(image))` for a PIL image (util/FSC147.py:125-126, 324-325): PIL's antialiased
    bilinear resize, then uint8 -> float32 / 255.  arr_u8: uint8 [H, W, 3].  Returns float32 torch [3, new_H, new_W]."""
    import torch
    from PIL import Image
    img = Image.fromarray(arr_u8).resize((new_hw[1], new_hw[0]), Image.BILINEAR)
    return torch.from_numpy(np.asarray(img).astype(np.float32) / 255.0).permute(2, 0, 1).contiguous()
