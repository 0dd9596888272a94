"""One Linear-layer GEMM shape of the B=8 step, launched a few times, for `ncu --set full` (see profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "proj"
m = 4608
n, k, mode = {"proj": (768, 768, "res"), "fc2": (768, 3072, "res"), "qkv": (2304, 768, "f16"), "fc1": (3072, 768, "gelu")}[which]
a = torch.randn(m, k, device=dev).half()
w = torch.randn(n, k, device=dev).half() * 0.05
bias = torch.zeros(n, device=dev)
if mode == "res":
    c = torch.zeros(m, n, device=dev)
    f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n)
elif mode == "gelu":
    c = torch.empty(m, n, device=dev, dtype=torch.float16)
    aux = torch.empty(m, n, device=dev, dtype=torch.float16)
    f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, act=1, aux=aux, ldaux=n)
else:
    c = torch.empty(m, n, device=dev, dtype=torch.float16)
    f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias)
for _ in range(8):
    f()
torch.cuda.synchronize()
