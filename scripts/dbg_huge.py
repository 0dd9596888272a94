"""Debug aid: stage-by-stage comparison of the decoder at the ViT-H/14 geometry (27 x 27 tokens) against the CPU oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from oracle import synth, countr_oracle as O
from test_parity_gpu import build, rel
from countr_b200.engine import engine, F32
cuda = torch.device("cuda:0")
m, sd, cfg = build("huge_d2", int(os.environ.get("SEED", "5")), cuda); m.eval()
imgs, boxes = synth.make_inputs(1, seed=10)
with torch.no_grad():
    rlat = O.forward_encoder(sd, cfg, imgs)
    taps = {}
    ref = O.forward_decoder(sd, cfg, rlat, boxes, 3, taps=taps)
    eng = engine()
    lat16 = rlat.to(cuda).half().view(-1, rlat.shape[-1]).contiguous()
    save = {}
    out = eng.decoder_forward(m, lat16, boxes.to(cuda), 3, 1, F32, save=save)
    print("map rel", rel(out, ref))
    print("y rel", rel(save["y16"].float().view(1, 3, 512), taps["y"]))
    print("fim (after decoder_norm) rel", rel(save["f16"].float().view(1, 729, 512), taps["fim"]))
    for bi, sv in enumerate(save["blocks"]):
        print(f" block {bi}: x1 (after self-attn) |x|={sv['x1'].norm().item():.3f} x2 {sv['x2'].norm().item():.3f}")
    # oracle head stages
    x = taps["fim"].transpose(1, 2).reshape(1, 512, 27, 27)
    for i in range(4):
        raw = F.conv2d(x, sd[f"decode_head{i}.0.weight"], sd[f"decode_head{i}.0.bias"], padding=1)
        got = save["heads"][i]["raw"].float().permute(0, 3, 1, 2)
        st = save["heads"][i]["stats"]
        n = raw.shape[2] * raw.shape[3] * 32
        rs = raw.view(1, 8, -1).double()
        print(f" head{i} raw rel {rel(got, raw):.3e}; GN sum rel {rel(st[0, :, 0], rs.sum(-1)[0]):.3e} sumsq rel {rel(st[0, :, 1], (rs * rs).sum(-1)[0]):.3e}")
        x = F.relu(F.group_norm(raw, 8, sd[f"decode_head{i}.1.weight"], sd[f"decode_head{i}.1.bias"], 1e-5))
        if i < 3:
            x = F.interpolate(x, size=x.shape[-1] * 2, mode="bilinear", align_corners=False)
            got_in = save["heads"][i + 1]["inp"].float().permute(0, 3, 1, 2)
            print(f"   next input rel {rel(got_in, x):.3e}")
