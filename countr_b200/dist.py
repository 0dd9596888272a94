"""Data-parallel plumbing: the only exchange step of the path is the gradient mean (the reference uses
DDP, FSC_finetune_cross.py:230).  Gradients live in ONE flat fp32 arena, so the mean is one collective."""
import torch
import torch.distributed as dist


def build_grad_arena(names, params, device):
    """One flat fp32 buffer with a 16-byte-aligned view per parameter.  Returns (arena, {name: view})."""
    total = sum((p.numel() + 3) // 4 * 4 for p in params)
    arena = torch.empty(total, dtype=torch.float32, device=device)
    views, off = {}, 0
    for n, p in zip(names, params):
        views[n] = arena[off:off + p.numel()].view(p.shape)
        off += (p.numel() + 3) // 4 * 4
    return arena, views


def make_grad_allreduce(group=None):
    """callable(arena): in-place mean over the ranks of `group` (NCCL: one AVG all-reduce over NVLink/NVSwitch;
    gloo, used by the CPU tests: SUM then scale)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)

    def allreduce(arena):
        if backend == "nccl":
            dist.all_reduce(arena, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=group)
            arena.mul_(1.0 / world)
    return allreduce
