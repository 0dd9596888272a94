"""Development aid: the SAME fine-tune step through stock PyTorch on the B200 (cuBLAS / cuDNN / ATen, fp16 autocast +
GradScaler-style static scale, fused AdamW), i.e. the 'existing Blackwell software stack' the reference would run on.
Uses the functional restatement in oracle/ on CUDA tensors; not part of the product or of bench.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import countr_oracle as O, synth

dev = torch.device("cuda:0")
cfg = synth.CONFIGS["base"]
sd = {k: v.to(dev) for k, v in synth.make_state_dict(cfg, 0).items()}
names = O.decoder_param_names(sd, 3)
params = []
for n in names:
    sd[n] = sd[n].clone().requires_grad_(True)
    params.append(sd[n])
opt = torch.optim.AdamW(params, lr=1e-5, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
B = 8
imgs, boxes = synth.make_inputs(B, seed=1)
gt, mask = synth.make_targets(B, seed=2)
imgs, boxes, gt, mask = imgs.to(dev), boxes.to(dev), gt.to(dev), mask.to(dev)
scale = 4096.0


def step():
    with torch.autocast("cuda", dtype=torch.float16):
        out = O.forward(sd, cfg, imgs, boxes, 3)
    loss = O.finetune_loss(out.float(), gt, mask)
    opt.zero_grad(set_to_none=True)
    (loss * scale).backward()
    torch._foreach_mul_([p.grad for p in params], 1.0 / scale)
    opt.step()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"torch eager fp16-autocast fine-tune step, B={B}: {ms:.3f} ms/step -> {B / ms * 1e3:.1f} img/s")
with torch.no_grad():
    def fwd():
        with torch.autocast("cuda", dtype=torch.float16):
            return O.forward(sd, cfg, imgs, boxes, 3)
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"torch eager fp16-autocast forward, B={B}: {ms:.3f} ms -> {B / ms * 1e3:.1f} img/s")
