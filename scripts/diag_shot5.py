import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import countr_oracle as O, synth
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_parity_gpu import build, rel
from countr_b200.engine import engine
dev = torch.device("cuda:0")
m, sd, cfg = build("small", 1, dev); m.eval()
imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
for shot in (3, 5, 4, 5, 3):
    taps = {}
    with torch.no_grad():
        ref = O.forward(sd, cfg, imgs, boxes, shot, taps)
        out = m(imgs.to(dev), boxes.to(dev), shot)
        y32, y16 = engine().exemplar_forward(m, boxes.to(dev), shot, None)
        torch.cuda.synchronize()
    yr = taps["y"].reshape(-1, 512)
    print(f"shot={shot}: out rel={rel(out, ref):.3e}  y rel={rel(y32, yr):.3e}  per-sample y err:",
          [round(rel(y32[i], yr[i]), 4) for i in range(y32.shape[0])])
