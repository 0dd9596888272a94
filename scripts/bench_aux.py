"""A/B of the aux epilogues (COUNTR_EPI_AUX=0/1): FIM / encoder-decoder MLP shapes, graph-timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops
from cublas_compare import timeit, dev
M = 4608
for name, n, k in (("fim fc1 (+pre)", 2048, 512), ("fim fc2 dX (gelu')", 2048, 512), ("noct dec fc1", 2048, 512), ("enc fc1 (+pre)", 3072, 768), ("enc fc2 dX", 3072, 768)):
    a = torch.randn(M, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev)
    pre = torch.randn(M, n, device=dev).half()
    u = torch.empty(M, n, device=dev, dtype=torch.float16)
    if "dX" in name:
        t = timeit(lambda: ops.linear(a, w, u, act=2, aux=pre))
    else:
        t = timeit(lambda: ops.linear(a, w, u, bias=bias, act=1, aux=pre))
    print(f"{name:20s} N={n} K={k}: {t:6.1f} us", flush=True)
