"""Data-parallel plumbing on CPU (gloo, world_size 2): flat gradient bucket + averaged all-reduce + broadcast.

The compute path needs CUDA; what is tested here is the host-side N > 1 logic bench.py uses: ranks
own disjoint meshes, the only exchange is ONE all-reduce over a contiguous gradient buffer.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from surfacenetworks_b200 import dist as D
    from surfacenetworks_b200 import models as M

    r, lr, w = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                       # ranks start DIFFERENT, broadcast must fix that
    model = nn.Sequential(nn.Linear(6, 8), nn.BatchNorm1d(8), nn.Linear(8, 3))
    D.broadcast_module(model)
    flat0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.zeros_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, flat0)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "broadcast_module left ranks out of sync"

    bucket = D.FlatGradAllReduce(model)
    assert bucket.numel == sum(p.numel() for p in model.parameters())
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    # each rank: its own shard of the "batch" (meshes are independent units)
    torch.manual_seed(7 + rank)
    x, y = torch.randn(16, 6), torch.randn(16, 3)
    for step in range(2):                               # second step checks zero() keeps the views alive
        bucket.zero()
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()
        local = bucket.flat.clone()
        views_ok = all(p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off
                       for p, off in zip(bucket.params, _offsets(bucket.params)))
        assert views_ok, "param.grad stopped aliasing the flat bucket"
        bucket.allreduce()
        both = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        expect = sum(both) / world
        assert torch.allclose(bucket.flat, expect, rtol=1e-6, atol=1e-7)
        assert torch.allclose(model[0].weight.grad.reshape(-1), expect[:48], rtol=1e-6, atol=1e-7)
    # the model the bench shards (parameter count = the all-reduce size quoted in SURVEY.md 8(e))
    if rank == 0:
        n = sum(p.numel() for p in M.ArapDirModel().parameters())
        out.put(n)
    dist.barrier()
    dist.destroy_process_group()


class _Stack(nn.Module):
    def __init__(self):
        super().__init__()
        for i in range(6):
            setattr(self, "rn%d" % i, nn.Sequential(nn.Linear(8, 8), nn.Tanh()))

    def forward(self, x):
        for i in range(6):
            x = x + getattr(self, "rn%d" % i)(x)
        return x


def _step_worker(rank, world, port, out):
    """graph.CapturedTrainStep's N > 1 path on CPU tensors (eager; gloo): after a step every rank's gradients are the
    mean of the ranks' local gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import copy
    from surfacenetworks_b200 import dist as D, graph as G
    D.init_from_env(backend="gloo")
    torch.manual_seed(3 + rank)
    x, y = torch.randn(32, 8), torch.randn(32, 8)
    torch.manual_seed(0)
    model = _Stack()
    D.broadcast_module(model)
    ref = copy.deepcopy(model)
    ((ref(x) - y) ** 2).mean().backward()
    expect = []
    for p in ref.parameters():
        parts = [torch.zeros_like(p.grad) for _ in range(world)]
        dist.all_gather(parts, p.grad)
        expect.append(sum(parts[1:], parts[0]) * (1.0 / world))
    opt = torch.optim.SGD(model.parameters(), lr=0.0)            # lr 0: parameters stay put, gradients are inspected
    step = G.CapturedTrainStep(model, lambda m, t, o: ((m(t["x"]) - t["y"]) ** 2).mean(), opt, {"x": x, "y": y},
                               warmup=0, capture=False)
    for _ in range(2):
        step.eager_step()
        for p, e in zip(model.parameters(), expect):
            assert torch.allclose(p.grad, e, rtol=1e-6, atol=1e-7)
    if rank == 0:
        out.put(step.grad_bytes)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_captured_step_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0, "rank failed"
    assert out.get(timeout=5) == 6 * (8 * 8 + 8) * 4


def _offsets(params):
    off = 0
    for p in params:
        yield off
        off += p.numel()


@pytest.mark.timeout(180)
def test_flat_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0, "rank failed"
    assert out.get(timeout=5) == 1018872                # 4.08 MB of fp32 gradients per all-reduce


def test_single_process_is_a_noop():
    from surfacenetworks_b200 import dist as D
    model = nn.Linear(4, 2)
    b = D.FlatGradAllReduce(model)
    model(torch.ones(3, 4)).sum().backward()
    before = b.flat.clone()
    b.allreduce()
    assert torch.equal(before, b.flat)
    D.broadcast_module(model)
