"""CPU, world_size 2, gloo: the data-parallel step logic (one flat all-reduce averaging gradients + BN running
statistics, parameter broadcast) -- the N > 1 path of bench.py / SURVEY.md 8e."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fabric_b200.distributed import DataParallelStep
    torch.manual_seed(100 + rank)                      # different initial weights per rank on purpose
    model = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(), nn.Conv2d(4, 2, 1))
    dp = DataParallelStep(model)
    dp.broadcast_parameters(0)
    w0 = [p.detach().clone() for p in model.parameters()]
    torch.manual_seed(7 + rank)                        # per-rank shard of the batch
    x = torch.randn(4, 3, 8, 8)
    model.train()
    model(x).square().mean().backward()
    local = [p.grad.clone() for p in model.parameters()]
    local_rm = model[1].running_mean.clone()
    dp.sync()
    gathered = [torch.zeros_like(torch.cat([g.flatten() for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.flatten() for g in local]))
    rms = [torch.zeros_like(local_rm) for _ in range(world)]
    dist.all_gather(rms, local_rm)
    ok = torch.allclose(torch.cat([p.grad.flatten() for p in model.parameters()]), sum(gathered) / world, atol=1e-7)
    ok = ok and torch.allclose(model[1].running_mean, sum(rms) / world, atol=1e-7)
    wall = [torch.zeros_like(torch.cat([w.flatten() for w in w0])) for _ in range(world)]
    dist.all_gather(wall, torch.cat([w.flatten() for w in w0]))
    ok = ok and all(torch.equal(wall[0], w) for w in wall)          # broadcast made the replicas identical
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_data_parallel_step_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
