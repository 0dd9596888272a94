"""BASELINE configs[4]: MAE pre-training step (models_mae_noct.mae_vit_base_patch16, mask_ratio 0.5), per-GPU batch 32 —
device time of forward + full backward (encoder included) + AdamW, CUDA-graph replay, synthetic images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import models_mae_noct as N

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "32"))
torch.manual_seed(0)
m = N.mae_vit_base_patch16(norm_pix_loss=True).to(dev).train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-5, betas=(0.9, 0.95), weight_decay=0.05, fused=True, capturable=True)
imgs = torch.rand(B, 3, 384, 384, device=dev)
scale = 1024.0


def step():
    loss, _, _ = m(imgs, mask_ratio=0.5)
    (loss * scale).backward()
    grads = [p.grad for p in m.parameters() if p.grad is not None]
    torch._foreach_mul_(grads, 1.0 / scale)
    opt.step()
    return loss


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
mode = "eager"
try:
    opt.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    run = g.replay
    mode = "cuda graph"
except Exception as e:  # noqa: BLE001
    print("graph capture failed:", type(e).__name__, e)
    torch.cuda.synchronize()

    def run():
        opt.zero_grad(set_to_none=True)
        step()
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"pre-train step B={B} mask 0.5 ({mode}): {ms:.3f} ms -> {B / ms * 1e3:.1f} img/s  ({B / ms * 262.62:.1f} TFLOP/s algorithmic, "
      f"{B / ms * 262.62 / 1375.4:.3f} of the sustained bf16 peak)")
