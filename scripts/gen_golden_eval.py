"""Writes tests/golden/eval_fewshot.npz by running the REFERENCE'S OWN evaluation script, `FSC_test_cross(few-shot).py` — its
`TestData` dataset and its `main()` loop (sliding window, 3 x 3 tiling for tiny exemplars, test-time normalisation) — in this
container, on CPU, over three small synthetic images, with the seeded synthetic base-model weights of oracle/synth.py.

What is shimmed (none of the script's own code is replaced):
  * absent packages: timm (the shim of scripts/gen_golden.py, built from the reference's own Attention / Mlp), matplotlib,
    wandb, cv2 (`rectangle` only draws the visualisation), `torch._six.inf`;
  * `misc.load_model_FSC` loads the synthetic state dict instead of a checkpoint file; `torch.cuda.synchronize` is a no-op;
  * torchvision's tensor `Resize` is held at the reference's pinned 0.14.1 behaviour (no antialias);
  * the module-level name `abs` logs its argument, which is how the exact `pred_cnt` of every image is read back
    (`cnt_err = abs(pred_cnt - gt_cnt)`, :361; the script only prints three decimals).
The fixture holds the uint8 images, the annotations, and per image: name, pred_cnt, gt_cnt, whether the 3 x 3 tiling ran."""
import builtins
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import gen_golden  # noqa: E402  (timm shim + reference import)
from oracle import synth  # noqa: E402

REF = "/root/reference"


def synth_image(h, w, seed):
    """Smooth uint8 RGB pattern (the same recipe as scripts/gen_golden_aug.py)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        for _ in range(4):
            fy, fx, ph = rng.uniform(0.01, 0.12), rng.uniform(0.01, 0.12), rng.uniform(0, 6.28)
            img[..., c] += rng.uniform(0.3, 1.0) * np.sin(fy * yy + fx * xx + ph)
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)


def stubs():
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("scatter", "xlabel", "ylabel", "savefig", "figure", "close"):
        setattr(plt, name, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    cv2 = types.ModuleType("cv2")
    cv2.rectangle = lambda img, *a, **k: img
    six = types.ModuleType("torch._six")
    six.inf = float("inf")
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "cv2": cv2, "wandb": types.ModuleType("wandb"), "torch._six": six})
    from torchvision import transforms
    _Resize = transforms.Resize
    transforms.Resize = lambda size, **kw: _Resize(size, antialias=False)


def eval_state_dict():
    """The weights both sides use: the seeded synthetic base model with a positive density offset, so that the mass under the
    exemplar boxes crosses the 1.8 threshold of the test-time normalisation for large boxes and stays below it for tiny ones."""
    cfg = synth.CONFIGS["base"]
    sd = synth.make_state_dict(cfg, seed=0)
    sd["decode_head3.3.bias"] = torch.full_like(sd["decode_head3.3.bias"], 0.5)
    return cfg, sd


def corners(y1, x1, y2, x2):
    return [[x1, y1], [x1, y2], [x2, y2], [x2, y1]]


def main():
    stubs()
    ref_models = gen_golden.import_reference()
    spec = importlib.util.spec_from_file_location("ref_fewshot", os.path.join(REF, "FSC_test_cross(few-shot).py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    misc = sys.modules["util.misc"]

    tmp = tempfile.mkdtemp(prefix="fsc_eval_")
    im_dir = os.path.join(tmp, "images")
    os.makedirs(im_dir)
    out = {}
    # name -> (H, W, exemplar boxes (y1, x1, y2, x2) in original pixels)
    images = {
        "p.png": (200, 300, [(20, 30, 70, 95), (100, 150, 160, 230), (40, 200, 90, 280)]),          # w = 576: three windows
        "q.png": (240, 320, [(30, 40, 34, 45), (100, 200, 105, 204), (150, 60, 155, 66)]),          # tiny exemplars: 3 x 3 tiling
        "r.png": (210, 210, [(15, 20, 80, 100), (120, 30, 190, 90), (60, 130, 110, 200)]),          # w = 384: one window
    }
    annotations, rng = {}, np.random.default_rng(3)
    for k, (name, (h, w, bxs)) in enumerate(images.items()):
        arr = synth_image(h, w, 200 + k)
        Image.fromarray(arr).save(os.path.join(im_dir, name))
        n = 17 + 9 * k
        pts = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1)
        annotations[name] = {"points": pts.tolist(), "box_examples_coordinates": [corners(*b) for b in bxs]}
        out["img_" + name[0]] = arr
        out["boxes_" + name[0]] = np.array(bxs)
        out["npoints_" + name[0]] = np.int64(n)
    ref.annotations = annotations
    ref.data_split = {"test": list(images)}
    ref.im_dir = im_dir

    cfg, sd = eval_state_dict()
    misc.load_model_FSC = lambda args, model_without_ddp: model_without_ddp.load_state_dict(sd, strict=True)
    torch.cuda.synchronize = lambda *a, **k: None
    logged = []

    def logging_abs(v):
        logged.append(float(v))
        return builtins.abs(v)

    ref.abs = logging_abs
    args = ref.get_args_parser().parse_args(["--device", "cpu", "--output_dir", os.path.join(tmp, "out"), "--resume", "", "--no_pin_mem"])
    os.makedirs(args.output_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    prints = []
    real_print = builtins.print
    ref.print = lambda *a, **k: (prints.append(" ".join(str(x) for x in a)), real_print(*a, **k))
    ref.main(args)

    # per-image lines: "i/n: pred_cnt: ..., gt_cnt: ..., ..., id: name, s_cnt: True/False"
    lines = [p for p in prints if "pred_cnt:" in p and "id:" in p]
    assert len(lines) == len(images) == len(logged), (len(lines), len(logged))
    names, tiled, gts = [], [], []
    for ln in lines:
        names.append(ln.split("id: ")[1].split(",")[0])
        tiled.append(ln.strip().endswith("True"))
        gts.append(float(ln.split("gt_cnt:")[1].split(",")[0]))
    pred = [err_arg + gt for err_arg, gt in zip(logged, gts)]         # abs() saw pred_cnt - gt_cnt
    for ln, p in zip(lines, pred):
        assert abs(float(ln.split("pred_cnt:")[1].split(",")[0]) - p) < 1e-3, (ln, p)
    out["names"] = np.array(names)
    out["tiled"] = np.array(tiled)
    out["gt_cnt"] = np.array(gts)
    out["pred_cnt"] = np.array(pred, dtype=np.float64)
    # ---- the zero-shot script on the same images: model(window, boxes, 0), no tiling (s_cnt >= 100 never holds), no normalisation
    spec0 = importlib.util.spec_from_file_location("ref_zeroshot", os.path.join(REF, "FSC_test_cross(zero-shot).py"))
    ref0 = importlib.util.module_from_spec(spec0)
    spec0.loader.exec_module(ref0)
    ref0.annotations, ref0.data_split, ref0.im_dir = annotations, {"test": list(images)}, im_dir
    prints0 = []
    ref0.print = lambda *a, **k: (prints0.append(" ".join(str(x) for x in a)), real_print(*a, **k))
    args0 = ref0.get_args_parser().parse_args(["--device", "cpu", "--output_dir", os.path.join(tmp, "out0"), "--resume", "", "--no_pin_mem",
                                               "--num_workers", "0"])
    os.makedirs(args0.output_dir, exist_ok=True)
    ref0.main(args0)
    lines0 = [p for p in prints0 if "pred_cnt:" in p and "id:" in p]
    assert len(lines0) == len(images), lines0
    out["zs_names"] = np.array([ln.split("id: ")[1].strip() for ln in lines0])
    out["zs_pred_cnt"] = np.array([float(ln.split("pred_cnt:")[1].split(",")[0]) for ln in lines0], dtype=np.float64)

    path = os.path.join(ROOT, "tests", "golden", "eval_fewshot.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB", dict(zip(names, zip(pred, tiled))))


if __name__ == "__main__":
    main()
