"""Device time of the fused self-attention forward on the shapes of the step (CUDA graph of 20 calls)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


SHAPES = ((8, 576, 12, 64), (8, 576, 16, 32), (32, 288, 12, 64), (128, 576, 12, 64), (32, 576, 16, 32), (128, 576, 16, 32))
for B, L, H, dh in (SHAPES[:1] if os.environ.get("ATTN_ONE") else SHAPES):
    qkv = torch.randn(B, L, 3, H, dh, device=dev).half()
    out = torch.empty(B, L, H * dh, device=dev, dtype=torch.float16)
    lse = torch.empty(B, H, L, device=dev)
    us = timeit(lambda: ops.attention_fwd(qkv, out, B, L, H, dh, dh ** -0.5, lse=lse))
    print(f"attention fwd B={B} L={L} H={H} dh={dh}: {us:8.1f} us  {4.0*B*H*L*L*dh/us/1e6:7.1f} TFLOP/s (dense)", flush=True)
