"""Fixed cost vs per-k-block cost of the Linear GEMM (production build): time at K = 768 ... 6144 for a one-wave
shape (N = 768: 144 tiles of 128 x 192) and a three-wave shape (N = 2304), ours and cuBLASLt.  The slope is the
steady-state k-block period, the intercept the launch + prologue + exposed epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops
from cublas_compare import timeit, dev  # noqa

M = 4608
for n, modes in ((768, ("f16", "res")), (2304, ("f16",)), (512, ("f16", "res")), (3072, ("gelu",))):
    for mode in modes:
        rows = []
        for k in (512, 768, 1536, 3072, 6144):
            a = torch.randn(M, k, device=dev).half()
            w = torch.randn(n, k, device=dev).half() * 0.05
            bias32 = torch.zeros(n, device=dev)
            bias16 = torch.zeros(n, device=dev).half()
            if mode == "res":
                c = torch.zeros(M, n, device=dev)
                ours = timeit(lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bias=bias32, residual=c, ldr=n))
                c16 = torch.empty(M, n, device=dev, dtype=torch.float16)
                cub = timeit(lambda: torch.addmm(bias16, a, w.t(), out=c16))
            else:
                c = torch.empty(M, n, device=dev, dtype=torch.float16)
                ours = timeit(lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bias=bias32, act=1 if mode == "gelu" else 0))
                cub = timeit(lambda: torch.addmm(bias16, a, w.t(), out=c))
            rows.append((k, ours, cub))
        (k0, o0, c0), (k1, o1, c1) = rows[1], rows[-1]
        so, sc = (o1 - o0) / ((k1 - k0) / 64), (c1 - c0) / ((k1 - k0) / 64)
        print(f"N={n:5d} {mode:4s} " + " ".join(f"K={k}: {o:5.1f}/{c:5.1f}" for k, o, c in rows) +
              f" | per k-block ours {so * 1e3:5.0f} ns cuBLAS {sc * 1e3:5.0f} ns; intercept ours {o0 - so * k0 / 64:5.1f} us cuBLAS {c0 - sc * k0 / 64:5.1f} us", flush=True)
