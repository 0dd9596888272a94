"""2-rank check (torchrun): the progressive slice all-reduce of the MAE pre-training backward (models_mae_noct: decoder
slice first, then every third encoder block, overlapped on a comm stream) leaves exactly the gradients of ONE all-reduce of
the whole arena after the backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import models_mae_noct as N
from countr_b200.engine import engine

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = N.MaskedAutoencoderViTNoCT(embed_dim=256, depth=7, num_heads=4, decoder_depth=2).to(dev).train()
g = torch.Generator().manual_seed(5 + rank)
imgs = torch.rand(2, 3, 384, 384, generator=g).to(dev)
m._noise_override = torch.rand(2, 576, generator=g)
eng = engine()
comm = torch.cuda.Stream()
pieces = []


def hook(sl):
    pieces.append(sl.numel())
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        dist.all_reduce(sl, op=dist.ReduceOp.AVG)


def grads(use_hook):
    for p in m.parameters():
        p.grad = None
    eng.grad_slice_hook = hook if use_hook else None
    loss, _, _ = m(imgs, mask_ratio=0.5)
    (loss * 64.0).backward()
    eng.grad_slice_hook = None
    torch.cuda.current_stream().wait_stream(comm)
    a = eng.last_arena
    if not use_hook:
        dist.all_reduce(a, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    return a.clone()


ref = grads(False)
got = grads(True)
same = torch.equal(ref, got)
# split-K dW accumulates with fp32 atomics: allow the run-to-run reduction-order noise of the local backward itself
rel = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
if rank == 0:
    print(f"[slice all-reduce] pieces {pieces} (sum {sum(pieces)} of {ref.numel()}); identical={same}; relL2={rel:.2e}")
assert sum(pieces) == ref.numel() and rel < 1e-5
dist.destroy_process_group()
