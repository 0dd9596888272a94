import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops
dev = torch.device("cuda:0")
which = sys.argv[1]
M = 4608
if which == "proj":
    m, n, k = M, 768, 768
    a = torch.randn(m, k, device=dev).half(); w = torch.randn(n, k, device=dev).half() * 0.05
    c = torch.zeros(m, n, device=dev); bias = torch.zeros(n, device=dev)
    f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n)
elif which == "qkv":
    m, n, k = M, 2304, 768
    a = torch.randn(m, k, device=dev).half(); w = torch.randn(n, k, device=dev).half() * 0.05
    c = torch.empty(m, n, device=dev, dtype=torch.float16); bias = torch.zeros(n, device=dev)
    f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias)
elif which == "conv":
    x16 = torch.randn(8, 192, 192, 256, device=dev).half()
    w16 = torch.randn(256, 9 * 256, device=dev).half() * 0.02
    y16 = torch.empty(8, 192, 192, 256, device=dev, dtype=torch.float16)
    bias = torch.zeros(256, device=dev)
    stats = torch.zeros(8, 8, 2, device=dev, dtype=torch.float64)
    f = lambda: ops.conv3x3(x16, w16, y16, bias=bias, gn_stats=stats)
elif which == "attn":
    qkv = torch.randn(8, 576, 3, 12, 64, device=dev).half()
    out = torch.empty(8, 576, 768, device=dev, dtype=torch.float16)
    f = lambda: ops.attention_fwd(qkv, out, 8, 576, 12, 64, 0.125)
for _ in range(6):
    f()
torch.cuda.synchronize()
