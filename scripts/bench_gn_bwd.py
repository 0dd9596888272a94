"""Device time of the GroupNorm+ReLU backward reduce pass (up-sample adjoint gather) on the decoder-head shapes, B=8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops

dev = torch.device("cuda:0")
B, C, G = 8, 256, 8
for H in ((96,) if os.environ.get("GN_ONE") else (24, 48, 96)):
    raw = torch.randn(B, H, H, C, device=dev).half()
    d_next = torch.randn(B, 2 * H, 2 * H, C, device=dev).half()
    xg = raw.double().reshape(B, H * H, G, C // G)
    stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    dyh = torch.empty_like(raw)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    gsum = torch.zeros(B, G, 2, device=dev, dtype=torch.float64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    f = lambda: ops.gn_relu_bwd_reduce(raw, stats, gamma, beta, dyh, dg, db, gsum, G, 1e-5, d_next=d_next)
    for _ in range(3):
        f()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    mb = (raw.numel() * 2 * 2 + d_next.numel() * 2) / 1e6
    print(f"gn_relu_bwd_reduce mode 0  {H}^2 -> {2*H}^2: {ts[len(ts)//2]:7.1f} us (cold L2)  {mb:6.1f} MB  {mb / ts[len(ts)//2]:5.2f} TB/s", flush=True)

# mode 1: the 1x1-conv head at 192^2 (reads raw, writes dyh)
H = 192
raw = torch.randn(B, H, H, C, device=dev).half()
xg = raw.double().reshape(B, H * H, G, C // G)
stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
dyh = torch.empty_like(raw)
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
gsum = torch.zeros(B, G, 2, device=dev, dtype=torch.float64)
dmap = torch.randn(B, H, H, device=dev)
w1, dw1, db1 = torch.randn(C, device=dev), torch.zeros(C, device=dev), torch.zeros(1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
f = lambda: ops.gn_relu_bwd_reduce(raw, stats, gamma, beta, dyh, dg, db, gsum, G, 1e-5, dmap=dmap, w1=w1, dw1=dw1, db1=db1)
for _ in range(3):
    f()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
mb = raw.numel() * 2 * 2 / 1e6
print(f"gn_relu_bwd_reduce mode 1  {H}^2: {ts[len(ts)//2]:7.1f} us (cold L2)  {mb:6.1f} MB  {mb / ts[len(ts)//2]:5.2f} TB/s", flush=True)

# forward: GroupNorm + ReLU + bilinear x2 (decode_head0..2)
for H in ((96,) if os.environ.get("GN_ONE") else (24, 48, 96)):
    raw = torch.randn(B, H, H, C, device=dev).half()
    xg = raw.double().reshape(B, H * H, G, C // G)
    stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()
    y = torch.empty(B, 2 * H, 2 * H, C, device=dev, dtype=torch.float16)
    f = lambda: ops.gn_relu_upsample2x(raw, stats, gamma, beta, y, G, 1e-5)
    for _ in range(3):
        f()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    mb = (raw.numel() + y.numel()) * 2 / 1e6
    print(f"gn_relu_upsample2x  {H}^2 -> {2*H}^2: {ts[len(ts)//2]:7.1f} us (cold L2)  {mb:6.1f} MB  {mb / ts[len(ts)//2]:5.2f} TB/s", flush=True)

# forward: GroupNorm + ReLU + 1x1 conv (decode_head3) at 192^2
H = 192
raw = torch.randn(B, H, H, C, device=dev).half()
xg = raw.double().reshape(B, H * H, G, C // G)
stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1).contiguous()
d = torch.empty(B, H, H, device=dev)
w1, bias1 = torch.randn(C, device=dev), torch.zeros(1, device=dev)
f = lambda: ops.gn_relu_conv1x1(raw, stats, gamma, beta, w1, bias1, d, G, 1e-5)
for _ in range(3):
    f()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
mb = raw.numel() * 2 / 1e6
print(f"gn_relu_conv1x1  {H}^2: {ts[len(ts)//2]:7.1f} us (cold L2)  {mb:6.1f} MB  {mb / ts[len(ts)//2]:5.2f} TB/s", flush=True)
