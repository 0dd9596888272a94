"""Quick device-side timing of the forward path per kernel family (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import models_mae_cross as M
from countr_b200 import ops

dev = torch.device("cuda:0")
m = M.mae_vit_base_patch16().to(dev).eval()
for B, shot in ((8, 3), (128, 0), (1, 3)):
    imgs = torch.rand(B, 3, 384, 384, device=dev)
    boxes = torch.rand(B, 3, 3, 64, 64, device=dev) if shot else torch.empty(B, 0, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m(imgs, boxes, shot)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES[0]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        iters = 5
        te = td = 0.0
        t0 = time.time()
        for _ in range(iters):
            e0.record()
            _, lat16 = m._encode(imgs)
            e1.record()
            m._decode(lat16, boxes, shot, B, torch.float32)
            e2.record()
            torch.cuda.synchronize()
            te += e0.elapsed_time(e1); td += e1.elapsed_time(e2)
        wall = (time.time() - t0) / iters * 1e3
    print(f"B={B} shot={shot}: encoder {te/iters:.3f} ms, decoder {td/iters:.3f} ms, total {(te+td)/iters:.3f} ms "
          f"-> {B/((te+td)/iters)*1e3:.1f} img/s (wall {wall:.2f} ms/iter, {(ops.LAUNCHES[0]-n0)//iters} launches)")
