"""Timeline of the attention forward kernel (CTA 0: softmax warp 0 of tile 0, MMA warp); needs `make -C countr_b200/csrc trace`."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from countr_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.__file__), "lib", "libcountr_sm100_trace.so")
from countr_b200 import ops

dev = torch.device("cuda:0")
lib = _lib.lib()
lib.countr_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
B, L, H, dh = 8, 576, 12, 64
qkv = torch.randn(B, L, 3, H, dh, device=dev).half()
out = torch.empty(B, L, H * dh, device=dev, dtype=torch.float16)
f = lambda: ops.attention_fwd(qkv, out, B, L, H, dh, dh ** -0.5)
for _ in range(3):
    f()
torch.cuda.synchronize()
tr = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.countr_debug_set_attn_trace(tr.data_ptr())
f()
torch.cuda.synchronize()
t = tr.cpu().tolist()
t0 = t[0]
rel = lambda v: v - t0 if v else -1
print("softmax warp 0 (tile 0), per chunk: [before s_full wait, S ready, S in regs (s_free), max done, p_free ok, P written (p_full)]")
for c in range(20):
    v = [rel(x) for x in t[16 + 8 * c:16 + 8 * c + 6]]
    if v[1] >= 0:
        print(f"  chunk {c:2d}: {v}   wait S {v[1]-v[0]:5d}  load {v[2]-v[1]:5d}  max {v[3]-v[2]:5d}  wait PV {v[4]-v[3]:5d}  exp+store {v[5]-v[4]:5d}")
if os.environ.get("COUNTR_ATTN4", "1") != "0":
    # per-CTA begin / end (globaltimer ns) of the four-tile kernel
    spans = [(t[2048 + 2 * i], t[2048 + 2 * i + 1]) for i in range(148) if t[2048 + 2 * i]]
    if spans:
        g0 = min(a for a, _ in spans)
        n_full = B * H * ((L + 127) // 128 // 4)
        full = [(a - g0, b - g0) for a, b in spans[:n_full]]
        rest = [(a - g0, b - g0) for a, b in spans[n_full:]]
        print(f"CTA spans (ns): {len(full)} four-tile CTAs begin {min(a for a, _ in full)}..{max(a for a, _ in full)} end {min(b for _, b in full)}..{max(b for _, b in full)}")
        if rest:
            print(f"                {len(rest)} remainder CTAs begin {min(a for a, _ in rest)}..{max(a for a, _ in rest)} end {min(b for _, b in rest)}..{max(b for _, b in rest)}")
    sys.exit(0)
print("MMA warp, per ring chunk j: [start, S(next) issued, p_full t0 ok, PV t0 issued, p_full t1 ok, PV t1 issued]")
for j in range(12):
    v = [rel(x) for x in t[1024 + 8 * j:1024 + 8 * j + 6]]
    if v[0] >= 0:
        print(f"  j {j:2d}: {v}   issue S {v[1]-v[0]:5d}  wait P0 {v[2]-v[1]:5d}  PV0 {v[3]-v[2]:5d}  wait P1 {v[4]-v[3]:5d}  PV1 {v[5]-v[4]:5d}")
