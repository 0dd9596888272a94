"""World-size-2 gloo test (CPU) of the N>1 host logic: the flat gradient arena and its single
mean all-reduce reproduce DDP's per-parameter gradient averaging, including parameters that are
unused on this step (absent from the arena, like DDP's find_unused_parameters)."""
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


def _worker(rank, world, port, shot_num, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models_mae_cross as M
    from countr_b200.dist import build_grad_arena, make_grad_allreduce
    torch.manual_seed(0)
    m = M.SupervisedMAE(embed_dim=128, depth=1, num_heads=2)
    names, params = m._decoder_params(shot_num)
    arena, views = build_grad_arena(names, params, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    for n in names:
        views[n].copy_(torch.randn(views[n].shape, generator=g))
    local = {n: views[n].clone() for n in names}
    make_grad_allreduce()(arena)
    # reference: gather every rank's local gradients and average per parameter
    ok = True
    for n in names:
        parts = [torch.empty_like(local[n]) for _ in range(world)]
        dist.all_gather(parts, local[n])
        ok &= torch.allclose(views[n], sum(parts) / world, atol=1e-6)
    ok &= all(v.data_ptr() % 16 == 0 for v in views.values())
    ok &= ("shot_token" in names) == (shot_num == 0) and any(n.startswith("decoder_proj") for n in names) == (shot_num > 0)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_arena_allreduce_world2():
    for shot in (3, 0):
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_worker, args=(2, port, shot, ret), nprocs=2, join=True)
            assert ret[0] and ret[1]
