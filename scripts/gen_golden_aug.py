"""Writes tests/golden/fsc147_transforms.npz by running the REFERENCE'S OWN dataset transforms (util/FSC147.py:
ResizeTrainImage.__call__ with and without augmentation, ResizeValImage.__call__) in this container on small synthetic images.

What is shimmed, and why (nothing of the transform code itself is replaced):
  * `cv2` and `imgaug` are absent here: empty stand-in modules; `iaa.Sequential([...])` becomes the identity on (image, key points).
    The collage branch (:183-262), the no-augmentation branch (:263-273) and the validation transform never look at the affine
    result, so their outputs are the reference's; the affine step itself stays unpinned.
  * the reference pins torchvision==0.14.1, where `transforms.Resize` on a TENSOR does not antialias; torchvision 0.26 (this image)
    would: `transforms.Resize` is wrapped to pass antialias=False (PIL inputs ignore the flag).
  * `random.random()` is scripted (so that the collage branch is taken) and every `random.randint` / `TF.crop` call is logged:
    the logged draws are what the tests replay through the oracle and the kernels.
The fixture holds the small uint8 source images (the tests redo the PIL resize exactly as the reference does), the draws, and
the reference's outputs (density maps whole; images as a strided sample plus the seam bands plus a float64 checksum)."""
import importlib.util
import json
import os
import random as _random
import sys
import tempfile
import types

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/util/FSC147.py"

# ---- stand-ins for the absent packages
sys.modules["cv2"] = types.ModuleType("cv2")
ia = types.ModuleType("imgaug")
iaa = types.ModuleType("imgaug.augmenters")
iab = types.ModuleType("imgaug.augmentables")


class Keypoint:
    def __init__(self, x, y):
        self.x, self.y = x, y

    def is_out_of_image(self, image):
        h, w = image.shape[:2]
        return not (0 <= self.y < h and 0 <= self.x < w)


class KeypointsOnImage:
    def __init__(self, keypoints, shape):
        self.keypoints, self.shape = keypoints, shape


class _Identity:
    def __call__(self, image, keypoints):
        return image, keypoints


iaa.Affine = lambda **kw: None
iaa.Sequential = lambda lst: _Identity()
iab.Keypoint, iab.KeypointsOnImage = Keypoint, KeypointsOnImage
ia.augmenters, ia.augmentables = iaa, iab
sys.modules.update({"imgaug": ia, "imgaug.augmenters": iaa, "imgaug.augmentables": iab})

from torchvision import transforms  # noqa: E402

_Resize = transforms.Resize
transforms.Resize = lambda size, **kw: _Resize(size, antialias=False)      # torchvision 0.14.1 semantics for tensors

spec = importlib.util.spec_from_file_location("ref_fsc147", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class RandomProxy:
    """The stdlib generator with scripted random() values first, and a log of every randint."""

    def __init__(self, seed, scripted):
        self.rng = _random.Random(seed)
        self.scripted = list(scripted)
        self.ints = []

    def random(self):
        return self.scripted.pop(0) if self.scripted else self.rng.random()

    def randint(self, a, b):
        v = self.rng.randint(a, b)
        self.ints.append((a, b, v))
        return v


class TFProxy:
    def __init__(self, real):
        self.real, self.crops = real, []

    def __getattr__(self, name):
        return getattr(self.real, name)

    def crop(self, img, top, left, height, width):
        self.crops.append((int(top), int(left), int(height), int(width)))
        return self.real.crop(img, top, left, height, width)


def synth_image(h, w, seed):
    """Smooth uint8 RGB pattern (compresses well, has gradients in both directions)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        for _ in range(4):
            fy, fx, ph = rng.uniform(0.01, 0.12), rng.uniform(0.01, 0.12), rng.uniform(0, 6.28)
            img[..., c] += rng.uniform(0.3, 1.0) * np.sin(fy * yy + fx * xx + ph)
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)


def image_sample(t):
    """What the fixture keeps of a [3, 384, 384] image."""
    a = t.numpy().astype(np.float32)
    return dict(grid=a[:, ::6, ::6].copy(), rows=a[:, 168:216, ::3].copy(), cols=a[:, ::3, 168:216].copy(),
                sum=np.float64(a.astype(np.float64).sum()))


def main():
    out = {}
    tmp = tempfile.mkdtemp(prefix="fsc_")
    sizes = {"a.png": (150, 206), "b.png": (171, 160), "c.png": (140, 230), "d.png": (200, 150), "e.png": (133, 201)}
    classes = {"a.png": "apples", "b.png": "apples", "c.png": "birds", "d.png": "apples", "e.png": "birds"}
    ann, rng = {}, np.random.default_rng(11)
    for k, (name, (h, w)) in enumerate(sizes.items()):
        arr = synth_image(h, w, 100 + k)
        Image.fromarray(arr).save(os.path.join(tmp, name))
        out["img_" + name[0]] = arr
        n = 90 if name == "a.png" else 25 + 7 * k
        pts = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1)
        ann[name] = {"points": pts.tolist()}
        out["dots_" + name[0]] = pts
    with open(os.path.join(tmp, "anno.json"), "w") as f:
        json.dump(ann, f)
    with open(os.path.join(tmp, "split.json"), "w") as f:
        json.dump({"train": list(sizes)}, f)
    with open(os.path.join(tmp, "classes.txt"), "w") as f:
        for name, c in classes.items():
            f.write(f"{name} {c}\n")
    args = types.SimpleNamespace(im_dir=tmp, anno_file=os.path.join(tmp, "anno.json"), data_split_file=os.path.join(tmp, "split.json"),
                                 class_file=os.path.join(tmp, "classes.txt"), do_aug=True)
    out["names"] = np.array(list(sizes))
    out["classes"] = np.array([classes[n] for n in sizes])
    boxes = {"a.png": [[20, 30, 60, 75], [80, 100, 120, 160], [5, 5, 33, 40]], "b.png": [[10, 12, 50, 44], [60, 70, 99, 120], [100, 20, 140, 60]]}
    out["boxes_a"], out["boxes_b"] = np.array(boxes["a.png"]), np.array(boxes["b.png"])

    def run(cls, name, scripted, seed, **kw):
        ref.random = RandomProxy(seed, scripted)
        ref.TF = TFProxy(sys.modules["torchvision.transforms.functional"])
        np.random.seed(seed)
        torch.manual_seed(seed)
        t = cls(args, 384, **kw) if kw else cls(args, 384)
        img = Image.open(os.path.join(tmp, name))
        img.load()
        sample = {"image": img, "lines_boxes": boxes.get(name, boxes["a.png"]), "dots": np.array(ann[name]["points"]), "id": name, "m_flag": 0}
        res = t(sample)
        return res, ref.random.ints, ref.TF.crops

    # --- case 1: self-collage (>= 70 objects): four crops of the sample itself
    res, ints, crops = run(ref.ResizeTrainImage, "a.png", [0.1, 0.3], 5, do_aug=True)
    assert len(crops) == 4 and ints[0][:2] == (10, 20)
    out["m1_blending_l"] = np.int64(ints[0][2])
    out["m1_crops"] = np.array([(t, l, h) for t, l, h, w in crops])
    out["m1_density"] = res["gt_density"].numpy()
    for k, v in image_sample(res["image"]).items():
        out["m1_image_" + k] = v
    out["m1_boxes"] = res["boxes"].numpy()
    assert res["m_flag"] == 0 and res["pos"].numel() == 0

    # --- case 2: collage of the sample and three other training images (< 70 objects), one slot of another class
    for seed in range(8, 64):          # first seed whose draw puts an image of ANOTHER class into the collage (:228)
        res, ints, crops = run(ref.ResizeTrainImage, "b.png", [0.1, 0.3, 0.9], seed, do_aug=True)
        assert len(crops) == 4 and ints[0][:2] == (10, 20) and ints[1][:2] == (0, 3)
        gt_pos, k, ids = ints[1][2], 2, []
        for i in range(4):
            if i == gt_pos:
                ids.append("b.png")
            else:
                ids.append(list(sizes)[ints[k][2]])
                k += 1
            k += 3
        if any(classes[i] != classes["b.png"] for i in ids) and sum(classes[i] == classes["b.png"] for i in ids) >= 2:
            break
    out["m2_blending_l"] = np.int64(ints[0][2])
    out["m2_ids"] = np.array(ids)
    out["m2_crops"] = np.array([(t, l, h) for t, l, h, w in crops])
    out["m2_density"] = res["gt_density"].numpy()
    for kk, v in image_sample(res["image"]).items():
        out["m2_image_" + kk] = v
    assert res["m_flag"] == 1
    print("case 2: quadrant images", ids, "blending_l", ints[0][2], "objects", float(res["gt_density"].sum()) / 60)

    # --- case 3: training transform without augmentation (:263-273, :279-306)
    res, ints, crops = run(ref.ResizeTrainImage, "a.png", [0.9], 3, do_aug=False)
    out["t_start"] = np.int64(ints[0][2])
    out["t_density"] = res["gt_density"].numpy()
    out["t_boxes"] = res["boxes"].numpy()
    out["t_pos"] = res["pos"].numpy()
    for kk, v in image_sample(res["image"]).items():
        out["t_image_" + kk] = v

    # --- case 4: validation transform (:316-357)
    res, ints, crops = run(ref.ResizeValImage, "b.png", [], 4)
    out["v_density"] = res["gt_density"].numpy()
    out["v_boxes"] = res["boxes"].numpy()
    out["v_pos"] = res["pos"].numpy()
    for kk, v in image_sample(res["image"]).items():
        out["v_image_" + kk] = v

    # --- case 5: the augmentation branch without collage (:133-180, :263-267): the Gaussian noise is replaced by zeros (the GPU path
    # has its own generator), the affine is the identity stand-in; ColorJitter / GaussianBlur draws are logged
    class _NoNoise:
        def __getattr__(self, name):
            return getattr(np.random, name)

        def normal(self, loc, scale, size):
            return np.zeros(size)

    class _NP:
        random = _NoNoise()

        def __getattr__(self, name):
            return getattr(np, name)

    log = {}
    cj, gb = ref.Augmentation.transforms
    real_cj, real_gb = type(cj).get_params, type(gb).get_params

    def cj_params(*a, **k):
        r = real_cj(*a, **k)
        log["jitter"] = ([int(i) for i in r[0]], [float(v) for v in r[1:]])
        return r

    def gb_params(*a, **k):
        r = real_gb(*a, **k)
        log["sigma"] = float(r)
        return r

    type(cj).get_params = staticmethod(cj_params)
    type(gb).get_params = staticmethod(gb_params)
    ref.np = _NP()
    try:
        res, ints, crops = run(ref.ResizeTrainImage, "a.png", [0.9, 0.7], 21, do_aug=True)
    finally:
        ref.np = np
        type(cj).get_params, type(gb).get_params = staticmethod(real_cj), staticmethod(real_gb)
    assert len(crops) == 1 and crops[0][2:] == (384, 384)
    out["a_jitter_order"] = np.array(log["jitter"][0])
    out["a_jitter_factors"] = np.array(log["jitter"][1])          # brightness, contrast, saturation, hue
    out["a_sigma"] = np.float64(log["sigma"])
    out["a_crop"] = np.array(crops[0][:2])
    out["a_density"] = res["gt_density"].numpy()
    for kk, v in image_sample(res["image"].float()).items():
        out["a_image_" + kk] = v
    print("case 5: jitter", log["jitter"], "sigma", log["sigma"], "crop", crops[0][:2], "objects", float(res["gt_density"].sum()) / 60)

    path = os.path.join(ROOT, "tests", "golden", "fsc147_transforms.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB;", {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("m1_") or k.startswith("t_")})


if __name__ == "__main__":
    main()
