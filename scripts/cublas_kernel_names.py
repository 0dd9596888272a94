"""Which cuBLASLt kernels (tile / cluster shapes are in the names) does torch pick for the Linear shapes of the step?
Run under `ncu --metrics gpu__time_duration.sum`; measurement aid only."""
import torch
dev = torch.device("cuda:0")
M = 4608
for name, n, k in [("enc qkv", 2304, 768), ("enc proj", 768, 768), ("enc fc1", 3072, 768), ("enc fc2", 768, 3072), ("dec embed", 512, 768),
                   ("fim qkv", 1536, 512), ("fim proj", 512, 512), ("fim fc1", 2048, 512), ("fim fc2", 512, 2048)]:
    a = torch.randn(M, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half()
    b = torch.zeros(n, device=dev).half()
    c = torch.empty(M, n, device=dev, dtype=torch.float16)
    for _ in range(3):
        torch.addmm(b, a, w.t(), out=c)
    torch.cuda.synchronize()
