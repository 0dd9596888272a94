"""Device time of the small direct backward kernels of the exemplar branch (B=8, 3 shots)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops

dev = torch.device("cuda:0")
B, S = 8, 3
boxes = torch.rand(B, S, 3, 64, 64, device=dev)
d_raw = torch.randn(B * S, 64, 64, 64, device=dev).half()
dw = torch.zeros(64, 3, 3, 3, device=dev)
f = lambda: ops.exemplar_conv1_dw(boxes, S, d_raw, dw)
for _ in range(3):
    f()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        f()
g.replay()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
print(f"exemplar_conv1_dw B={B} S={S}: {best:6.1f} us  (COUNTR_CONV1_DW_SUB={os.environ.get('COUNTR_CONV1_DW_SUB', '1')})")
