import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import synth
from test_parity_gpu import build, rel
from countr_b200.engine import engine
dev = torch.device("cuda:0")
m, sd, cfg = build("small", 1, dev); m.eval()
imgs, boxes = synth.make_inputs(3, seed=5)
imgs, boxes = imgs.to(dev), boxes.to(dev)
eng = engine()
with torch.no_grad():
    outs = []
    for it in range(3):
        lat32, lat16 = eng.encoder_forward(m, imgs)
        y32, y16 = eng.exemplar_forward(m, boxes, 3, None)
        raws = {k: v.clone() for k, v in eng.ws.bufs.items() if k[0].startswith("ex_")}
        out = eng.decoder_forward(m, lat16, boxes, 3, 3, torch.float32)
        torch.cuda.synchronize()
        bufs = {k: v.clone() for k, v in eng.ws.bufs.items()}
        outs.append((lat32.clone(), y32.clone(), out.clone(), raws, bufs))
    for it in (1, 2):
        print(f"run {it} vs 0: latent {rel(outs[it][0], outs[0][0]):.3e}  y {rel(outs[it][1], outs[0][1]):.3e}  out {rel(outs[it][2], outs[0][2]):.3e}")
        for k in sorted(outs[0][3], key=str):
            a, b = outs[0][3][k].float(), outs[it][3][k].float()
            print("   ex", k[0], k[1], f"{rel(b, a):.3e}")
        for k in sorted(outs[0][4], key=str):
            if k[0].startswith(("dec_", "head_")):
                a, b = outs[0][4][k].double(), outs[it][4][k].double()
                print("   ", k[0], k[1], f"{rel(b, a):.3e}")
