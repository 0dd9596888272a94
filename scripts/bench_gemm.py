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
lin("enc qkv", M, 2304, 768, "f16")
lin("enc fc1 (gelu)", M, 3072, 768, "gelu")
lin("enc fc2 (+res)", M, 768, 3072, "res")
lin("fim fc1 (gelu)", M, 2048, 512, "gelu")
lin("fim fc2 dX (gelu bwd)", M, 2048, 512, "gelubwd")


def head_elementwise():
    C, G = 256, 8
    for h in (24, 48, 96):
        x = torch.randn(B, h, h, C, device=dev).half()
        stats = torch.stack([x.double().reshape(B, h * h, G, 32).sum((1, 3)), (x.double() ** 2).reshape(B, h * h, G, 32).sum((1, 3))], -1).contiguous()
        gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
        y = torch.empty(B, 2 * h, 2 * h, C, device=dev, dtype=torch.float16)
        us = timeit(lambda: ops.gn_relu_upsample2x(x, stats, gamma, beta, y, G, 1e-5))
        mb = (x.numel() + y.numel()) * 2 / 1e6
        print(f"gn_relu_up2 {h}->{2*h}: {us:7.1f} us  {mb/us*1e3:7.1f} GB/s")
    h = 192
    x = torch.randn(B, h, h, C, device=dev).half()
    stats = torch.stack([x.double().reshape(B, h * h, G, 32).sum((1, 3)), (x.double() ** 2).reshape(B, h * h, G, 32).sum((1, 3))], -1).contiguous()
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    w = torch.randn(C, device=dev); bias = torch.zeros(1, device=dev)
    d = torch.empty(B, h, h, device=dev)
    us = timeit(lambda: ops.gn_relu_conv1x1(x, stats, gamma, beta, w, bias, d, G, 1e-5))
    print(f"gn_relu_conv1x1 192: {us:7.1f} us  {x.numel()*2/1e6/us*1e3:7.1f} GB/s")
    dyh = torch.empty_like(x); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    gsum = torch.zeros(B, G, 2, device=dev, dtype=torch.float64); dw1 = torch.zeros(C, device=dev); db1 = torch.zeros(1, device=dev)
    dmap = torch.randn(B, h, h, device=dev)
    us = timeit(lambda: ops.gn_relu_bwd_reduce(x, stats, gamma, beta, dyh, dg, db, gsum, G, 1e-5, dmap=dmap, w1=w, dw1=dw1, db1=db1))
    print(f"gn_relu_bwd_reduce mode1 192: {us:7.1f} us  {2*x.numel()*2/1e6/us*1e3:7.1f} GB/s")
    dbias = torch.zeros(C, device=dev)
    us = timeit(lambda: ops.gn_bwd_apply(x, dyh, stats, gsum, gamma, dyh, dbias, G, 1e-5))
    print(f"gn_bwd_apply 192: {us:7.1f} us  {3*x.numel()*2/1e6/us*1e3:7.1f} GB/s")
    for hh in (96, 48):
        xs = torch.randn(B, hh, hh, C, device=dev).half()
        st = torch.stack([xs.double().reshape(B, hh * hh, G, 32).sum((1, 3)), (xs.double() ** 2).reshape(B, hh * hh, G, 32).sum((1, 3))], -1).contiguous()
        dn = torch.randn(B, 2 * hh, 2 * hh, C, device=dev).half()
        dy2 = torch.empty_like(xs)
        us = timeit(lambda: ops.gn_relu_bwd_reduce(xs, st, gamma, beta, dy2, dg, db, gsum, G, 1e-5, d_next=dn))
        print(f"gn_relu_bwd_reduce mode0 {hh}: {us:7.1f} us  {(2*xs.numel()+dn.numel())*2/1e6/us*1e3:7.1f} GB/s")


head_elementwise()
