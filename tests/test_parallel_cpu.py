"""Host-side logic of the data-parallel path, on CPU: bucket planning, and a world_size-2 gloo run showing that the
sum-all-reduce + 1/world grad scale of tpgan_b200.parallel reproduces the single-process gradient of the concatenated
batch (exact for this model family: G and D have no BatchNorm at config defaults - SURVEY 8e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from tpgan_b200.parallel import allreduce_sum, plan_buckets


def test_plan_buckets_cover_everything_in_order():
    sizes = [5, 5, 5, 5, 3, 40, 1]
    ready = [1, 2, 3, 4, 9, 12, 12]
    b = plan_buckets(sizes, ready, 8)
    assert b == [(0, 2, 0, 10), (2, 4, 10, 10), (4, 6, 20, 43), (6, 7, 63, 1)]
    assert sum(x[3] for x in b) == sum(sizes)
    assert plan_buckets([3], [0], 100) == [(0, 1, 0, 3)]
    assert plan_buckets([], [], 8) == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import model_port as mp, step as ostep
    from tpgan_b200 import D_and_G_model as M
    torch.manual_seed(0)
    torch.set_num_threads(2)
    D = M.Discriminator(False)
    pd = {k: v.clone().requires_grad_(True) for k, v in D.state_dict().items()}
    b = ostep.make_batch(2)                        # global batch 2 = 1 image per rank
    x = b["img"][rank:rank + 1]
    loss = mp.discriminator(pd, x).mean()          # local mean
    grads = torch.autograd.grad(loss, list(pd.values()))
    flat = torch.cat([g.flatten() for g in grads])
    allreduce_sum(flat)      # stream / call-ordered (async_op=False): complete on return for gloo
    flat *= 1.0 / world                            # the optimizer's grad_scale
    if rank == 0:
        full = mp.discriminator(pd, b["img"]).mean()
        gfull = torch.cat([g.flatten() for g in torch.autograd.grad(full, list(pd.values()))])
        ret["err"] = float((flat - gfull).norm() / gfull.norm())
    dist.destroy_process_group()


def test_gloo_world2_gradient_equivalence():
    mgr = tmp.Manager()
    ret = mgr.dict()
    tmp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["err"] < 1e-5, ret["err"]
