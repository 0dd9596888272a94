"""How far are our Linear-layer GEMMs from cuBLASLt on the same shapes?  (measurement only — cuBLAS is not on the product path)
Device time per call from a CUDA graph of 20 back-to-back calls, B=8 fine-tune shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops

dev = torch.device("cuda:0")
M = int(os.environ.get("B", "8")) * 576


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
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


SHAPES = [("enc qkv", 2304, 768), ("enc proj", 768, 768), ("enc fc1", 3072, 768), ("enc fc2", 768, 3072),
          ("dec embed", 512, 768), ("fim qkv", 1536, 512), ("fim proj/wq", 512, 512), ("fim fc1", 2048, 512),
          ("fim fc2", 512, 2048), ("kv (M=24)", 512, 512)]
for name, n, k in SHAPES:
    m = 24 if name.startswith("kv") else M
    a = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev).half()
    c = torch.empty(m, n, device=dev, dtype=torch.float16)
    bias32 = torch.zeros(n, device=dev)
    ours = timeit(lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias32))
    cub = timeit(lambda: torch.addmm(bias, a, w.t(), out=c))
    fl = 2.0 * m * n * k
    print(f"{name:12s} M={m:5d} N={n:5d} K={k:5d}  ours {ours:6.1f} us ({fl/ours/1e6:6.0f} TF)   cuBLAS {cub:6.1f} us ({fl/cub/1e6:6.0f} TF)", flush=True)
