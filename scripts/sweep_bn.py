"""Sweep the N tile (and CTA-pair mode) of the Linear-layer GEMM shapes of the B=8 fine-tune step against the automatic
choice (pick_bn in csrc/gemm.cu).  Device time per call from a CUDA graph of 20 back-to-back calls."""
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


def make(m, n, k, mode):
    a = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev)
    if mode == "f16":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        return lambda bn, pr: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, bn=bn, pair=pr)
    if mode == "gelu":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        aux = torch.empty(m, n, device=dev, dtype=torch.float16)
        return lambda bn, pr: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, act=1, bn=bn, pair=pr)
    if mode == "gelubwd":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        aux = torch.randn(m, n, device=dev).half()
        return lambda bn, pr: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, act=2, aux=aux, ldaux=n, bn=bn, pair=pr)
    if mode == "res":
        c = torch.zeros(m, n, device=dev)
        return lambda bn, pr: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n, bn=bn, pair=pr)
    c = torch.zeros(m, n, device=dev)
    return lambda bn, pr: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bn=bn, pair=pr)


SHAPES = [("enc qkv", 2304, 768, "f16"), ("enc proj", 768, 768, "res"), ("enc fc1", 3072, 768, "gelu"), ("enc fc2", 768, 3072, "res"),
          ("dec embed", 512, 768, "res"), ("fim qkv", 1536, 512, "f16"), ("fim proj/wq", 512, 512, "res"), ("fim fc1", 2048, 512, "gelu"),
          ("fim fc2", 512, 2048, "res"), ("fim dX fc2", 2048, 512, "gelubwd"), ("fim dX fc1", 512, 2048, "f32"), ("fim dX qkv", 512, 1536, "f32")]
for name, n, k, mode in SHAPES:
    f = make(M, n, k, mode)
    auto = timeit(lambda: f(0, 0))
    res = []
    for pr in (-1, 1):
        for bn in (64, 96, 128, 160, 192, 224, 256):
            try:
                res.append((timeit(lambda: f(bn, pr)), bn, pr))
            except Exception as e:  # noqa: BLE001
                res.append((float("inf"), bn, pr))
    res.sort()
    fl = 2.0 * M * n * k
    top = "  ".join(f"bn{bn}{'p' if pr > 0 else ''}={us:.1f}" for us, bn, pr in res[:4])
    print(f"{name:12s} N={n:5d} K={k:5d} {mode:8s} auto {auto:6.1f} us ({fl/auto/1e6:6.0f} TF)  best: {top}   worst {res[-1][0]:.1f}", flush=True)
