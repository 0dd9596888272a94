"""Data-parallel plumbing: the only exchange step of the path is the gradient mean (the reference uses
DDP, FSC_finetune_cross.py:230).  Gradients live in ONE flat fp32 arena, so the mean is one collective.

The arena always covers EVERY decoder parameter in `named_parameters()` order — also the ones a given
`shot_num` does not touch (`shot_token` for shot_num > 0, `decoder_proj*` for shot_num == 0; the script draws
`shot_num` per rank from an unseeded `random.randint`, FSC_finetune_cross.py:278-284) — with absent gradients
zero-filled, followed by ARENA_TAIL usage flags.  Every rank therefore reduces a buffer of the same size and
layout whatever it drew, and after the mean the flags say which optional parameter groups were used on ANY
rank: exactly what DDP(find_unused_parameters=True) tracks with its used-parameter bitmap.
"""
import torch
import torch.distributed as dist

ARENA_TAIL = 4          # [0] shot_token used, [1] decoder_proj* (exemplar CNN) used, [2..3] spare (keeps 16-byte alignment)
FLAG_SHOT_TOKEN, FLAG_EXEMPLAR = 0, 1


def param_flag_index(name):
    """0: the parameter takes part in every step; k > 0: only when usage flag k - 1 is set."""
    if name == "shot_token":
        return FLAG_SHOT_TOKEN + 1
    if name.startswith("decoder_proj"):
        return FLAG_EXEMPLAR + 1
    return 0


def arena_size(params, tail=0):
    return sum((p.numel() + 3) // 4 * 4 for p in params) + tail


def build_grad_arena(names, params, device, tail=0):
    """One flat fp32 buffer with a 16-byte-aligned view per parameter (+ `tail` trailing floats).
    Returns (arena, {name: view})."""
    total = arena_size(params, tail)
    arena = torch.empty(total, dtype=torch.float32, device=device)
    views, off = {}, 0
    for n, p in zip(names, params):
        views[n] = arena[off:off + p.numel()].view(p.shape)
        off += (p.numel() + 3) // 4 * 4
    return arena, views


def usage_flags(shot_num):
    """Tail of the arena for a step run with `shot_num` exemplars (models_mae_cross.py:157-177)."""
    f = [0.0] * ARENA_TAIL
    f[FLAG_SHOT_TOKEN if shot_num == 0 else FLAG_EXEMPLAR] = 1.0
    return f


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
