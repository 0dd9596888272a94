"""In-kernel timeline of the GEMM kernel (clock64 stamps of one CTA) — needs the instrumented build:
   make -C countr_b200/csrc trace   ->  countr_b200/lib/libcountr_sm100_trace.so
Prints, per shape/variant, when the producer issued each k-block, when the MMA thread saw it land, and the epilogue times."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libcountr_sm100.so", "libcountr_sm100_trace.so")
from countr_b200 import ops

dev = torch.device("cuda:0")
lib = _lib.lib()
lib.countr_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
trace = torch.zeros(2048, dtype=torch.int64, device=dev)


def run(name, m, n, k, mode, bn=0, pair=0, cluster=0, cta=0, show=14):
    a = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev)
    if mode == "gelu":
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, act=1, bn=bn, pair=pair, cluster=cluster)
    elif mode == "res":
        c = torch.zeros(m, n, device=dev)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n, bn=bn, pair=pair, cluster=cluster)
    else:
        c = torch.empty(m, n, device=dev, dtype=torch.float16)
        f = lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, bn=bn, pair=pair, cluster=cluster)
    lib.countr_debug_set_trace(None, 0)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    trace.zero_()
    lib.countr_debug_set_trace(trace.data_ptr(), cta)
    torch.cuda.synchronize()
    f()
    torch.cuda.synchronize()
    t = trace.cpu().tolist()
    t0 = t[0]
    rel = lambda v: (v - t0) if v else -1
    P = [rel(v) for v in t[16:256] if v]
    P2 = [rel(v) for v in t[256:512] if v]
    Fm = [rel(v) for v in t[512:768] if v]
    F2 = [rel(v) for v in t[768:1024] if v]
    E = [rel(v) for v in t[1024:1100] if v]
    print(f"--- {name} M={m} N={n} K={k} {mode} bn={bn} pair={pair} cl={cluster} cta={cta}")
    print(f"    prologue sync +{rel(t[1])}  pdl_wait +{rel(t[2])}  end +{rel(t[3])} clk ({rel(t[3]) / 1965:.2f} us @1.965GHz)")
    print("    producer issue:", P[:show], "..." if len(P) > show else "")
    print("    producer done :", P2[:show])
    print("    mma full-wait :", Fm[:show], "..." if len(Fm) > show else "")
    print("    mma issued    :", F2[:show])
    if len(Fm) > 8:
        d = [Fm[i + 1] - Fm[i] for i in range(4, len(Fm) - 1)]
        print(f"    steady k-block period: mean {sum(d) / len(d):.0f} clk  (min {min(d)}, max {max(d)}); k-blocks {len(Fm)}; last full-wait +{Fm[-1]}")
    print("    epilogue (acc ready, drained) per tile:", E[:8])
    C = [rel(v) for v in t[1200:1264] if v]
    print("    first tile, warp 2, per chunk (start, tmem loaded, transposed, stores issued):", [tuple(C[i:i + 4]) for i in range(0, len(C), 4)])


lib.countr_debug_set_knobs.argtypes = [ctypes.c_int]
M = 4608
lib.countr_debug_set_knobs(0)
run("enc fc2 res", M, 768, 3072, "res", show=8)
run("fim fc2 res", M, 512, 2048, "res", show=8)
run("enc qkv f16", M, 2304, 768, "f16", show=8)
run("enc proj res", M, 768, 768, "res", show=8)
run("enc fc1 gelu", M, 3072, 768, "gelu", show=8)
run("fim fc2 res pair", M, 512, 2048, "res", bn=128, pair=1, show=8)
run("N=1024 bn256 pair", M, 1024, 4096, "f16", bn=256, pair=1, show=8)
