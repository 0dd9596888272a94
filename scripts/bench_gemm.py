"""Device-time microbenchmark of the GEMM / conv / attention kernels on the shapes of the B=8 fine-tune step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import ops

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "8"))
M = B * 576


def timeit(fn, reps=20):
    """GPU time per call: the calls are captured in a CUDA graph so host launch cost (ctypes + tensor-map
    encode, ~15 us per GEMM call) does not bound the measurement."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def lin(name, m, n, k, mode, bn=0, cl=0, pr=0):
    a = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev)
    if mode == "f16":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, bn=bn, cluster=cl, pair=pr)
    elif mode == "gelu":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, act=1, bn=bn, cluster=cl, pair=pr)
    elif mode == "gelubwd":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        aux = torch.randn(m, n, device=dev).half()
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, act=2, aux=aux, ldaux=n, bn=bn, cluster=cl, pair=pr)
    elif mode == "res":
        c = torch.zeros(m, n, device=dev)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n, bn=bn, cluster=cl, pair=pr)
    elif mode == "f32":
        c = torch.zeros(m, n, device=dev)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bn=bn, cluster=cl, pair=pr)
    us = timeit(f)
    print(f"{name:28s} M={m:6d} N={n:5d} K={k:5d} {mode:8s} bn={bn:3d} cl={cl} pair={pr}: {us:8.1f} us  {2.0*m*n*k/us/1e6:7.1f} TFLOP/s")


def dw(name, m_tok, n_out, k_in, split):
    dy = torch.randn(m_tok, n_out, device=dev).half()
    x = torch.randn(m_tok, k_in, device=dev).half()
    c = torch.zeros(n_out, k_in, device=dev)
    f = lambda: ops.gemm(dy, x, c, n_out, k_in, m_tok, lda=n_out, ldb=k_in, ldc=k_in, a_mn=True, b_mn=True, atomic=True, split_k=split)
    us = timeit(f)
    print(f"{name:28s} dW [{n_out}x{k_in}] over {m_tok} split={split:2d}: {us:8.1f} us  {2.0*m_tok*n_out*k_in/us/1e6:7.1f} TFLOP/s")


def conv(name, h, cin, cout, pr=0):
    x = torch.randn(B, h, h, cin, device=dev).half()
    w = torch.randn(cout, 9 * cin, device=dev).half() * 0.02
    y = torch.empty(B, h, h, cout, device=dev, dtype=torch.float16)
    bias = torch.zeros(cout, device=dev)
    stats = torch.zeros(B, cout // 32, 2, device=dev, dtype=torch.float64)
    us = timeit(lambda: ops.conv3x3(x, w, y, bias=bias, gn_stats=stats, pair=pr))
    fl = 2.0 * B * h * h * cout * 9 * cin
    dyv = torch.randn(B, h, h, cout, device=dev).half()
    dwp = torch.zeros(cout, 9 * cin, device=dev)
    us2 = timeit(lambda: ops.conv3x3_dw(dyv, x, dwp))
    print(f"{name:28s} pair={pr} conv {h}x{h} {cin}->{cout}: fwd {us:8.1f} us {fl/us/1e6:7.1f} TF | dW {us2:8.1f} us {fl/us2/1e6:7.1f} TF")


def attn(name, H, dh):
    qkv = torch.randn(B, 576, 3, H, dh, device=dev).half()
    out = torch.empty(B, 576, H * dh, device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.attention_fwd(qkv, out, B, 576, H, dh, dh ** -0.5))
    print(f"{name:28s} attention H={H} dh={dh}: {us:8.1f} us  {4.0*B*H*576*576*dh/us/1e6:7.1f} TFLOP/s")


print(f"B={B}")
def dwp(name, m_tok, n_out, k_in, split, pr):
    dy = torch.randn(m_tok, n_out, device=dev).half()
    x = torch.randn(m_tok, k_in, device=dev).half()
    c = torch.zeros(n_out, k_in, device=dev)
    f = lambda: ops.gemm(dy, x, c, n_out, k_in, m_tok, lda=n_out, ldb=k_in, ldc=k_in, a_mn=True, b_mn=True, atomic=True, split_k=split, bn=256, pair=pr)
    us = timeit(f)
    print(f"{name:28s} dW [{n_out}x{k_in}] over {m_tok} split={split:2d} pair={pr}: {us:8.1f} us  {2.0*m_tok*n_out*k_in/us/1e6:7.1f} TFLOP/s")
for pr in (-1, 1):
    dwp("fim fc2 dW", M, 512, 2048, 4, pr)
    dwp("fim fc1 dW", M, 2048, 512, 4, pr)
    dwp("fim qkv dW", M, 1536, 512, 6, pr)
    dwp("fim proj dW", M, 512, 512, 16, pr)
    for h, cin in ((24, 512), (48, 256), (96, 256), (192, 256)):
        x = torch.randn(B, h, h, cin, device=dev).half()
        dyv = torch.randn(B, h, h, 256, device=dev).half()
        dwq = torch.zeros(256, 9 * cin, device=dev)
        us2 = timeit(lambda: ops.conv3x3_dw(dyv, x, dwq, pair=pr))
        fl = 2.0 * B * h * h * 256 * 9 * cin
        print(f"conv dW {h}x{h} {cin}->256 pair={pr}: {us2:8.1f} us {fl/us2/1e6:7.1f} TF")
