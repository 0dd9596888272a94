"""Where does the in-graph fine-tune step go?  Graph-captured pieces of the B=8 step, device-timed (CUDA events around
graph replays): encoder forward | decoder forward | loss + backward | optimizer update, and the whole step for reference.
The pieces are timed alone, so work that overlaps across pieces in the whole step (side streams) shows as their sum being
larger than the whole."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import models_mae_cross
from countr_b200.engine import engine, F32
from countr_b200.train import FineTuner
from countr_b200.backward import decoder_backward

dev = torch.device("cuda:0")
B, S = int(os.environ.get("B", "8")), 3
torch.manual_seed(0)
model = models_mae_cross.mae_vit_base_patch16(norm_pix_loss=False).to(dev).train()
g = torch.Generator().manual_seed(1)
imgs = torch.rand(B, 3, 384, 384, generator=g).to(dev)
boxes = torch.rand(B, S, 3, 64, 64, generator=g).to(dev)
gt = (torch.rand(B, 384, 384, generator=g) * 0.5).to(dev)
mask = (torch.rand(384, 384, generator=g) < 0.8).float().to(dev)
tuner = FineTuner(model, lr=1e-5, loss_scale=4096.0)
eng = engine()


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


state = {}


def whole():
    tuner.forward_backward(imgs, boxes, gt, mask, S)
    tuner.update()


def enc():
    with torch.no_grad():
        state["lat16"] = eng.encoder_forward(model, imgs, keep=True)[1]


def dec_fwd():
    with torch.no_grad():
        ev = eng.refresh_decoder_weights(model, S, True, dev)
        pre = eng.exemplar_async(model, boxes, S, train=True)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        save = {}
        out = eng.decoder_forward(model, state["lat16"], boxes, S, B, F32, save=save, pre=pre)
        state["save"], state["out"] = save, out


def loss_bwd():
    with torch.no_grad():
        dout = torch.empty_like(state["out"])
        tuner._loss(state["out"], gt, mask, dout)
        decoder_backward(eng, model, state["save"], boxes, dout)
        tuner.arena = eng.last_arena


def upd():
    tuner.update()


t_whole = timed(whole)
t_enc = timed(enc)
enc()
t_dec = timed(dec_fwd)
dec_fwd()
t_bwd = timed(loss_bwd)
loss_bwd()
t_upd = timed(upd)
print(f"B={B}: whole step {t_whole:7.1f} us | encoder fwd {t_enc:7.1f} | decoder fwd (+refresh, exemplar) {t_dec:7.1f} | loss + backward {t_bwd:7.1f} | update {t_upd:6.1f} | sum {t_enc + t_dec + t_bwd + t_upd:7.1f}")
eng.overlap_dw = False
eng.overlap_exemplar = False
t_bwd1 = timed(loss_bwd)
t_dec1 = timed(dec_fwd)
print(f"single stream: decoder fwd {t_dec1:7.1f} | loss + backward {t_bwd1:7.1f}")
