"""World-size-2 gloo tests (CPU) of the N>1 host logic: the flat gradient arena and its single mean all-reduce reproduce
DDP's per-parameter gradient averaging — including the case the reference script creates on purpose: every rank draws its
own `shot_num` from an unseeded `random.randint` (FSC_finetune_cross.py:278-284), so the ranks of one step can disagree on
which parameters are used (`shot_token` vs the exemplar CNN `decoder_proj*`).  The arena therefore always has the layout of
ALL decoder parameters, absent gradients are zero, and usage flags travel in its tail (countr_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shots, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models_mae_cross as M
    from countr_b200.dist import ARENA_TAIL, build_grad_arena, make_grad_allreduce, param_flag_index, usage_flags
    torch.manual_seed(0)
    m = M.SupervisedMAE(embed_dim=128, depth=1, num_heads=2)
    shot_num = shots[rank]
    names, params = m._decoder_params(None)                 # the arena layout: every decoder parameter
    used = set(m._decoder_params(shot_num)[0])              # what this rank's backward writes
    arena, views = build_grad_arena(names, params, "cpu", tail=ARENA_TAIL)
    arena.zero_()
    arena[-ARENA_TAIL:] = torch.tensor(usage_flags(shot_num))
    g = torch.Generator().manual_seed(100 + rank)
    for n in names:
        if n in used:
            views[n].copy_(torch.randn(views[n].shape, generator=g))
    local = {n: views[n].clone() for n in names}
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([arena.numel()]))
    ok = len({int(s) for s in sizes}) == 1                  # same buffer size on every rank whatever it drew
    make_grad_allreduce()(arena)
    # reference: gather every rank's local gradients (zeros where a rank did not use the parameter) and average
    for n in names:
        parts = [torch.empty_like(local[n]) for _ in range(world)]
        dist.all_gather(parts, local[n])
        ok &= torch.allclose(views[n], sum(parts) / world, atol=1e-6)
    ok &= all(v.data_ptr() % 16 == 0 for v in views.values())
    # usage flags after the mean: > 0 exactly for the groups some rank used (DDP's globally-used bitmap)
    flags = arena[-ARENA_TAIL:].tolist()
    any_zero, any_few = any(s == 0 for s in shots), any(s > 0 for s in shots)
    ok &= (flags[0] > 0) == any_zero and (flags[1] > 0) == any_few
    for n in names:
        k = param_flag_index(n)
        ok &= (k == 1) == (n == "shot_token") and (k == 2) == n.startswith("decoder_proj")
    ok &= ("shot_token" in used) == (shot_num == 0) and any(n.startswith("decoder_proj") for n in used) == (shot_num > 0)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_arena_allreduce_world2():
    for shots in ((3, 3), (0, 0), (3, 0)):                  # the last: rank 0 few-shot, rank 1 zero-shot in the same step
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_worker, args=(2, port, shots, ret), nprocs=2, join=True)
            assert ret[0] and ret[1], shots


def test_scheduled_lr_matches_the_reference_schedule():
    """countr_b200.train.scheduled_lr against util/lr_sched.py:adjust_learning_rate run by scripts/gen_golden_lr.py."""
    import os
    import numpy as np
    from countr_b200.train import scheduled_lr
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lr_sched.npz"))
    for name in ("finetune", "pretrain"):
        lr, min_lr, warm, epochs = (float(v) for v in g[name + "_args"])
        got = np.array([scheduled_lr(float(e), lr, min_lr, warm, epochs) for e in g[name + "_epochs"]])
        assert np.array_equal(got, g[name + "_lr"])
