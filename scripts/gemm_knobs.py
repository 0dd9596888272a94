"""Which resource paces the GEMM main loop?  Instrumented build (make trace), graph-timed, with the debug knobs:
   0 = normal, 4 = producer issues no TMA loads (operands stay whatever is in smem), 8 = MMA warp issues no MMAs,
   12 = neither (pure barrier hand-off), 1 = no C stores."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libcountr_sm100.so", "libcountr_sm100_trace.so")
from countr_b200 import ops
from bench_gemm import timeit, dev

lib = _lib.lib()
M = 4608
for name, n, k, mode, bn in [("enc fc2", 768, 3072, "res", 0), ("enc fc2 K=12288", 768, 12288, "res", 0), ("fim fc2", 512, 2048, "res", 0),
                             ("N=512 K=8192", 512, 8192, "res", 0), ("enc qkv", 2304, 768, "f16", 0), ("N=1024 bn256 K=8192", 1024, 8192, "f16", 256),
                             ("N=256 bn64 K=8192", 256, 8192, "f16", 64)]:
    a = torch.randn(M, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half() * 0.05
    bias = torch.zeros(n, device=dev)
    if mode == "res":
        c = torch.zeros(M, n, device=dev)
        f = lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bias=bias, residual=c, ldr=n, bn=bn)
    else:
        c = torch.empty(M, n, device=dev, dtype=torch.float16)
        f = lambda: ops.gemm(a, w, c, M, n, k, lda=k, ldb=k, ldc=n, bias=bias, bn=bn)
    out = []
    for knob in (0, 4, 8, 12, 1):
        lib.countr_debug_set_knobs(knob)
        torch.cuda.synchronize()
        out.append((knob, timeit(f)))
    lib.countr_debug_set_knobs(0)
    kb = k // 64
    print(f"{name:22s} N={n} K={k}: " + "  ".join(f"knob{kn}={us:6.1f}us ({us * 1965 / kb:5.0f} clk/kb)" for kn, us in out), flush=True)
