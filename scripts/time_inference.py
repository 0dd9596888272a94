"""BASELINE configs[3]: zero-shot inference (0 exemplars, learnable shot token), batch 128, 1 x B200 — device time,
CUDA-graph replay, plus the 3-shot B=8 / B=1 forward for reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import models_mae_cross as M

dev = torch.device("cuda:0")
m = M.mae_vit_base_patch16().to(dev).eval()
for B, shot in ((128, 0), (8, 3), (1, 3)):
    imgs = torch.rand(B, 3, 384, 384, device=dev)
    boxes = torch.rand(B, 3, 3, 64, 64, device=dev) if shot else torch.empty(B, 0, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m(imgs, boxes, shot)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m(imgs, boxes, shot)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    gf = 179.48 if shot == 0 else 180.89
    print(f"inference B={B} shots={shot}: {ms:.3f} ms -> {B / ms * 1e3:.1f} img/s  ({B / ms * gf:.1f} TFLOP/s algorithmic)")
