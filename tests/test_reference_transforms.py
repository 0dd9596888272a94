"""The data-path oracle (oracle/data_oracle.py) pinned against the REFERENCE'S OWN transforms: tests/golden/fsc147_transforms.npz
holds outputs of util/FSC147.py's ResizeTrainImage (collage branch, and the no-augmentation branch) and ResizeValImage executed
by scripts/gen_golden_aug.py, together with the random draws they made.  CPU tests replay the draws through the oracle; the GPU
tests replay them through the CUDA kernels (countr_b200.data)."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fsc147_transforms.npz")


def _sample(a):
    return dict(grid=a[:, ::6, ::6], rows=a[:, 168:216, ::3], cols=a[:, ::3, 168:216], sum=float(a.astype(np.float64).sum()))


def _check_image(g, prefix, img, atol):
    s = _sample(np.asarray(img, dtype=np.float32))
    for k in ("grid", "rows", "cols"):
        ref = g[f"{prefix}_image_{k}"]
        assert np.allclose(s[k], ref, rtol=0, atol=atol), (prefix, k, np.abs(s[k] - ref).max())
    assert abs(s["sum"] - float(g[f"{prefix}_image_sum"])) < max(atol * 3 * 384 * 384, 1e-3)


def _collage_inputs(g, case):
    """Images, crops, dots, scales, same-class flags of collage case 1 (self) / 2 (four training images) as the reference drew them."""
    names = [str(n) for n in g["names"]]
    cls = dict(zip(names, (str(c) for c in g["classes"])))
    ids = ["a.png"] * 4 if case == 1 else [str(i) for i in g["m2_ids"]]
    own = "a.png" if case == 1 else "b.png"
    resized, scales, dots = {}, {}, {}
    for i in set(ids):
        arr = g["img_" + i[0]]
        h, w = arr.shape[:2]
        nh, nw = D.flex_resize(h, w)
        resized[i] = D.resize_pil(arr, (nh, nw))
        scales[i] = (float(nh) / h, float(nw) / w)
        dots[i] = g["dots_" + i[0]]
    crops = [tuple(int(v) for v in r) for r in g[f"m{case}_crops"]]
    return ([resized[i] for i in ids], crops, int(g[f"m{case}_blending_l"]), [dots[i] for i in ids], [scales[i] for i in ids],
            [cls[i] == cls[own] for i in ids])


@pytest.mark.parametrize("case", [1, 2])
def test_collage_oracle_matches_the_reference(case):
    """util/FSC147.py:183-262 + :265-269, run by the reference itself."""
    g = np.load(GOLD)
    images, crops, bl, dots, scales, same = _collage_inputs(g, case)
    img, dmap = D.mosaic(images, crops, bl, dots, scales, same)
    _check_image(g, f"m{case}", img.numpy(), 0.0)                      # same torch operations in the same order: bit-exact
    den = D.filter_density(dmap.numpy())
    assert np.array_equal(den, g[f"m{case}_density"])
    if case == 2:
        assert not all(same) and sum(same) >= 2                        # a quadrant of another class contributes no objects (:228)


def test_no_augmentation_and_validation_oracles_match_the_reference():
    """util/FSC147.py:263-306 (do_aug=False) and :316-357 (ResizeValImage), run by the reference itself."""
    g = np.load(GOLD)
    a = g["img_a"]
    h, w = a.shape[:2]
    nh, nw = D.flex_resize(h, w)
    ra = D.resize_pil(a, (nh, nw))
    start = int(g["t_start"])
    assert np.array_equal(D.train_density(g["dots_a"], h, w, nh, nw, start), g["t_density"])
    _check_image(g, "t", ra[:, :384, start:start + 384].numpy(), 0.0)
    sh, sw = float(nh) / h, float(nw) / w
    rects = [(int(int(b[0]) * sh), int(int(b[1]) * sw), int(int(b[2]) * sh), int(int(b[3]) * sw)) for b in g["boxes_a"]]
    assert torch.equal(D.crop_resize_boxes(ra, rects), torch.from_numpy(g["t_boxes"]))
    pos = [[y1, max(0, x1 - start), y2, min(384, x2 - start)] for y1, x1, y2, x2 in rects]       # :289
    assert np.array_equal(np.array(pos), g["t_pos"])
    # collage case 1 crops its exemplars from the same un-augmented image (:286-288)
    assert torch.equal(D.crop_resize_boxes(ra, rects), torch.from_numpy(g["m1_boxes"]))
    b = g["img_b"]
    h, w = b.shape[:2]
    rb = D.resize_pil(b, (384, 384))
    assert np.array_equal(D.val_density(g["dots_b"], h, w), g["v_density"])
    _check_image(g, "v", rb.numpy(), 0.0)
    sh, sw = 384.0 / h, 384.0 / w
    rects = [(int(int(q[0]) * sh), int(int(q[1]) * sw), int(int(q[2]) * sh), int(int(q[3]) * sw)) for q in g["boxes_b"]]
    assert torch.equal(D.crop_resize_boxes(rb, rects), torch.from_numpy(g["v_boxes"]))
    assert np.array_equal(np.array(rects), g["v_pos"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [1, 2])
def test_collage_kernels_match_the_reference(cuda, case):
    """countr_b200.data.mosaic / mosaic_density against the reference's own collage output."""
    from countr_b200 import data
    g = np.load(GOLD)
    images, crops, bl, dots, scales, same = _collage_inputs(g, case)
    dev = [im.to(cuda) for im in images]
    got = data.mosaic(dev, crops, bl).cpu().numpy()
    _check_image(g, f"m{case}", got, 2e-6)
    all_dots = torch.from_numpy(np.concatenate(dots)).to(cuda)
    begins = np.cumsum([0] + [len(d) for d in dots[:-1]])
    ranges = [(int(b), len(d) if s else 0) for b, d, s in zip(begins, dots, same)]
    den = data.mosaic_density(dev, crops, bl, all_dots, ranges, scales).cpu().numpy()
    assert np.allclose(den, g[f"m{case}_density"], rtol=0, atol=1e-6), np.abs(den - g[f"m{case}_density"]).max()


@pytest.mark.gpu
def test_density_and_box_kernels_match_the_reference(cuda):
    """density_from_dots / crop_resize_boxes against the reference's no-augmentation and validation transforms."""
    from countr_b200 import data
    g = np.load(GOLD)
    for key, img_key, val in (("t", "img_a", False), ("v", "img_b", True)):
        arr = g[img_key]
        h, w = arr.shape[:2]
        nh, nw = (384, 384) if val else D.flex_resize(h, w)
        r = D.resize_pil(arr, (nh, nw)).to(cuda)
        sh, sw = float(nh) / h, float(nw) / w
        dots = torch.from_numpy(g["dots_" + img_key[-1]])[None].contiguous().to(cuda)
        counts = torch.tensor([dots.shape[1]], dtype=torch.int32, device=cuda)
        start = 0 if val else int(g["t_start"])
        den = data.density_from_dots(dots, counts, scale=(sh, sw), canvas_hw=(nh, nw), origin=(0, start),
                                     sigma=4.0 if val else 1.0, radius=7 if val else None)[0].cpu().numpy()
        assert np.array_equal(den, g[key + "_density"]), np.abs(den - g[key + "_density"]).max()
        boxes = g["boxes_" + img_key[-1]]
        rects = torch.tensor([[int(int(b[0]) * sh), int(int(b[1]) * sw), int(int(b[2]) * sh), int(int(b[3]) * sw)] for b in boxes],
                             dtype=torch.int32, device=cuda)[None]
        got = data.crop_resize_boxes(r[None], rects)[0].cpu()
        assert torch.allclose(got, torch.from_numpy(g[key + "_boxes"]), rtol=0, atol=2e-6)


def _aug_draws(g):
    order = [int(v) for v in g["a_jitter_order"]]
    by_fn = [float(v) for v in g["a_jitter_factors"]]                  # brightness, contrast, saturation, hue
    top, left = (int(v) for v in g["a_crop"])
    return order, by_fn, float(g["a_sigma"]), (top, left)


def _robust_image_check(g, prefix, img):
    """fp32 arithmetic against the reference's float64 pipeline; hue can pick the neighbouring sector at a boundary."""
    s = _sample(np.asarray(img, dtype=np.float32))
    for k in ("grid", "rows", "cols"):
        diff = np.abs(s[k] - g[f"{prefix}_image_{k}"])
        assert diff.mean() < 3e-6 and (diff > 1e-4).mean() < 1e-3, (prefix, k, diff.mean(), diff.max())


def test_augmentation_branch_composition_matches_the_reference():
    """util/FSC147.py:133-180, 263-269 run by the reference (zero noise, identity affine, logged ColorJitter / GaussianBlur / flip /
    crop draws) against the same steps composed from torchvision's functional ops and the oracle's dot map, in fp32."""
    tvF = pytest.importorskip("torchvision.transforms.functional")
    g = np.load(GOLD)
    order, by_fn, sigma, (top, left) = _aug_draws(g)
    a = g["img_a"]
    h, w = a.shape[:2]
    nh, nw = D.flex_resize(h, w)
    x = torch.clamp(D.resize_pil(a, (nh, nw)), 0, 1)
    fns = [tvF.adjust_brightness, tvF.adjust_contrast, tvF.adjust_saturation, tvF.adjust_hue]
    for fn in order:
        x = fns[fn](x, by_fn[fn])
    x = tvF.gaussian_blur(x, kernel_size=[7, 9], sigma=[sigma, sigma])
    x = x.flip(-1)[:, top:top + 384, left:left + 384]
    _robust_image_check(g, "a", x.numpy())
    canvas = np.ascontiguousarray(D.affine_dot_canvas(g["dots_a"], h, w, nh, nw, np.eye(3))[:, ::-1])[top:top + 384, left:left + 384]
    assert np.array_equal(D.filter_density(canvas), g["a_density"])


@pytest.mark.gpu
def test_train_transform_kernels_match_the_reference(cuda):
    """countr_b200.data.train_transform (every pixel operation a kernel) against the reference's own augmentation branch."""
    from countr_b200 import data
    g = np.load(GOLD)
    order, by_fn, sigma, crop = _aug_draws(g)
    a = g["img_a"]
    h, w = a.shape[:2]
    nh, nw = D.flex_resize(h, w)
    sh, sw = float(nh) / h, float(nw) / w
    ra = D.resize_pil(a, (nh, nw)).to(cuda)
    ops_t = torch.tensor(order, dtype=torch.int32).view(4, 1)
    fac = torch.tensor([by_fn[fn] for fn in order], dtype=torch.float32).view(4, 1)
    rects = torch.tensor([[int(int(b[0]) * sh), int(int(b[1]) * sw), int(int(b[2]) * sh), int(int(b[3]) * sw)] for b in g["boxes_a"]],
                         dtype=torch.int32)
    draws = dict(mosaic=None, noise_seed=0, jitter=(ops_t, fac), blur_sigma=torch.tensor([sigma]),
                 affine=dict(rotate_deg=0.0, scale=1.0, shear_deg=0.0, translate_frac=(0.0, 0.0)), flip=True, crop=crop)
    out = data.train_transform(ra, torch.from_numpy(g["dots_a"]).to(cuda), (sh, sw), rects, draws, noise_std=0.0)
    _robust_image_check(g, "a", out["image"].cpu().numpy())
    den = out["gt_density"].cpu().numpy()
    assert np.allclose(den, g["a_density"], rtol=0, atol=1e-6), np.abs(den - g["a_density"]).max()
    assert torch.allclose(out["boxes"].cpu(), torch.from_numpy(g["m1_boxes"]), rtol=0, atol=2e-6)      # same un-augmented crops (:286-288)
