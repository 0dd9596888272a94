"""TEST INFRASTRUCTURE — CPU restatement of the reference's per-image evaluation, FSC_test_cross(few-shot).py:258-359
(the sliding 384-px window with overlap blending, the 3 x 3 tiling for tiny exemplars, the count, the test-time
normalisation).  Only tests/ may import this; the product path (countr_b200/infer.py) never does.

`forward(imgs, boxes, shot_num) -> [N, 384, 384]` is the model under test on the CPU (oracle/countr_oracle.forward with
the state dict bound).  torchvision is not needed: TF.crop is a slice and transforms.Resize((h, w)) on a float tensor is
F.interpolate(mode="bilinear", align_corners=False) without antialiasing in the reference's torchvision 0.14.1 pin (up-scaling, so
antialiasing would not matter either).
Pinned against the script itself: scripts/gen_golden_eval.py runs the reference's own TestData + main() and
tests/test_reference_eval.py compares this restatement (and the CUDA path) with the counts it printed."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def window_pass(forward, image, boxes, shot_num):
    """:322-349 (and :289-318 inside the tiling loop): literal control flow and ZeroPad2d blending."""
    _, _, h, w = image.shape
    density_map = torch.zeros([h, w])
    start, prev = 0, -1
    while start + 383 < w:
        output, = forward(image[:, :, :, start:start + 384], boxes, shot_num)
        output = output.squeeze(0)
        b1 = nn.ZeroPad2d(padding=(start, w - prev - 1, 0, 0))
        d1 = b1(output[:, 0:prev - start + 1])
        b2 = nn.ZeroPad2d(padding=(prev + 1, w - start - 384, 0, 0))
        d2 = b2(output[:, prev - start + 1:384])
        b3 = nn.ZeroPad2d(padding=(0, w - start, 0, 0))
        density_map_l = b3(density_map[:, 0:start])
        density_map_m = b1(density_map[:, start:prev + 1])
        b4 = nn.ZeroPad2d(padding=(prev + 1, 0, 0, 0))
        density_map_r = b4(density_map[:, prev + 1:w])
        density_map = density_map_l + density_map_r + density_map_m / 2 + d1 / 2 + d2
        prev = start + 383
        start = start + 128
        if start + 383 >= w:
            if start == w - 384 + 128:
                break
            else:
                start = w - 384
    return density_map


def evaluate_image(forward, samples, boxes, pos, max_s_cnt=1, normalization=True, demo=False):
    """:258-359 — or, with demo=True, demo.py:76-169 (run_one_image): small exemplars counted over ALL boxes, the literal shot
    count 3, the nine crops taken row by row.  Returns (pred_cnt float, list of density maps: one, or the nine crops' maps)."""
    num_boxes = 3 if demo else (boxes.shape[1] if boxes.nelement() > 0 else 0)
    _, _, h, w = samples.shape
    r_cnt = s_cnt = 0
    for rect in pos:
        r_cnt += 1
        if r_cnt > 3 and not demo:
            break
        if rect[2] - rect[0] < 10 and rect[3] - rect[1] < 10:
            s_cnt += 1
    if s_cnt >= max_s_cnt:
        crop = lambda top, left, hh, ww: samples[0][:, top:top + hh, left:left + ww]          # TF.crop, no padding needed  # noqa: E731
        if demo:                                                                                 # demo.py:86-94
            r_images = [crop(0, 0, int(h / 3), int(w / 3)), crop(0, int(w / 3), int(h / 3), int(w / 3)),
                        crop(0, int(w * 2 / 3), int(h / 3), int(w / 3)), crop(int(h / 3), 0, int(h / 3), int(w / 3)),
                        crop(int(h / 3), int(w / 3), int(h / 3), int(w / 3)), crop(int(h / 3), int(w * 2 / 3), int(h / 3), int(w / 3)),
                        crop(int(h * 2 / 3), 0, int(h / 3), int(w / 3)), crop(int(h * 2 / 3), int(w / 3), int(h / 3), int(w / 3)),
                        crop(int(h * 2 / 3), int(w * 2 / 3), int(h / 3), int(w / 3))]
        else:
            r_images = [crop(0, 0, int(h / 3), int(w / 3)), crop(int(h / 3), 0, int(h / 3), int(w / 3)),
                        crop(0, int(w / 3), int(h / 3), int(w / 3)), crop(int(h / 3), int(w / 3), int(h / 3), int(w / 3)),
                        crop(int(h * 2 / 3), 0, int(h / 3), int(w / 3)), crop(int(h * 2 / 3), int(w / 3), int(h / 3), int(w / 3)),
                        crop(0, int(w * 2 / 3), int(h / 3), int(w / 3)), crop(int(h / 3), int(w * 2 / 3), int(h / 3), int(w / 3)),
                        crop(int(h * 2 / 3), int(w * 2 / 3), int(h / 3), int(w / 3))]
        pred_cnt = 0
        maps = []
        for r_image in r_images:
            r_image = F.interpolate(r_image.unsqueeze(0), size=(h, w), mode="bilinear", align_corners=False)
            density_map = window_pass(forward, r_image, boxes, num_boxes)
            pred_cnt += torch.sum(density_map / 60).item()
            maps.append(density_map)
    else:
        density_map = window_pass(forward, samples, boxes, num_boxes)
        pred_cnt = torch.sum(density_map / 60).item()
        maps = [density_map]
    if normalization:
        e_cnt = 0
        for rect in pos:
            e_cnt += torch.sum(density_map[rect[0]:rect[2] + 1, rect[1]:rect[3] + 1] / 60).item()     # the LAST map (reference quirk)
        e_cnt = e_cnt / 3
        if e_cnt > 1.8:
            pred_cnt /= e_cnt
    return pred_cnt, maps
