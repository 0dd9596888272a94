"""Tile-policy sweep of the Linear-layer GEMMs of the B=8 step (graph-timed): N tile, CTA pairs, multicast clusters,
epilogue kinds.  Development aid for pick_bn / the auto pair policy in csrc/gemm.cu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops
from bench_gemm import timeit, dev, M  # noqa

SHAPES = [("enc qkv", 2304, 768, "f16"), ("enc proj", 768, 768, "res"), ("enc fc1", 3072, 768, "gelu"), ("enc fc2", 768, 3072, "res"),
          ("fim qkv", 1536, 512, "f16"), ("fim proj", 512, 512, "res"), ("fim fc1", 2048, 512, "gelu"), ("fim fc2", 512, 2048, "res"),
          ("fim dx fc1", 512, 2048, "f32"), ("fim dx fc2", 2048, 512, "f16")]
VARIANTS = [(0, 0, 0), (256, 0, 0), (192, 0, 0), (128, 0, 0), (64, 0, 0), (256, 0, 1), (192, 0, 1), (128, 0, 1), (64, 0, 1), (256, 2, 0), (128, 2, 0)]
only = os.environ.get("ONLY")
for name, n, k, mode in SHAPES:
    if only and only not in name:
        continue
    a = torch.randn(M, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev)
    res = []
    for bn, cl, pr in VARIANTS:
        if mode == "res":
            c = torch.zeros(M, n, device=dev)
            f = lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n, bn=bn, cluster=cl, pair=pr if pr else -1)
        elif mode == "f32":
            c = torch.zeros(M, n, device=dev)
            f = lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bn=bn, cluster=cl, pair=pr if pr else -1)
        else:
            c = torch.empty(M, n, device=dev, dtype=torch.float16)
            f = lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bias=bias, act=1 if mode == "gelu" else 0, bn=bn, cluster=cl,
                                 pair=pr if pr else -1)
        try:
            us = timeit(f)
        except Exception as e:
            us = float("nan")
        res.append(us)
    fl = 2.0 * M * n * k
    best = min(r for r in res if r == r)
    print(f"{name:11s} N={n:5d} K={k:5d} {mode:5s} | " + " ".join(f"bn{bn}{'p' if pr else ''}{'c' if cl else ''}={us:5.1f}" for (bn, cl, pr), us in zip(VARIANTS, res))
          + f" | best {best:5.1f} us = {fl / best / 1e6:5.0f} TF", flush=True)
