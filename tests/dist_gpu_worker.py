"""Two-rank NCCL check run by tests/test_gpu_graph.py under torchrun (one process per GPU).

Every rank: (1) single-GPU gradients of its own shard on model copy A; all-gathered and averaged -> expected;
(2) CapturedTrainStep (graph with the NCCL all-reduce inside) on copy B -> p.grad must equal the expectation;
(3) three replays == three eager multi-rank steps on copy C (parameters bit-equal)."""
import copy
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _mark(rank, what):
    if rank == 0:
        print("[dist_gpu_worker] " + what, flush=True)


def main():
    from surfacenetworks_b200 import dist as D, graph as G, models as M, operators as OP, workloads as W
    rank, local_rank, world = D.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = 3
    meshes = W.make_mesh_ops(300, range(rank * B, rank * B + B))
    host = W.arap_batch(meshes, seed=rank)
    t = {k: host[k].to(dev) for k in ("inputs", "targets", "mask")}
    o = {k: OP.Bsr4Operator.from_torch_coo(host[k].to(dev)) for k in ("Di", "DiA")}
    torch.manual_seed(0)
    model_a = M.ArapDirModel().to(dev).train()
    D.broadcast_module(model_a)
    model_b, model_c = copy.deepcopy(model_a), copy.deepcopy(model_a)

    def loss_fn(m, t, o):
        return M.arap_loss(m(o["Di"], o["DiA"], t["mask"], t["inputs"]), t["targets"], t["mask"], B)

    def adam(m):
        return torch.optim.Adam(m.parameters(), 1e-3, weight_decay=1e-5, fused=True, capturable=True)

    _mark(rank, "setup done")
    # (1) local gradients, averaged by hand
    loss_fn(model_a, t, o).backward()
    expected = []
    for p in model_a.parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g.contiguous())
        expected.append(sum(parts[1:], parts[0]) * (1.0 / world))
    # (2) captured step, first replay happens on the SAME initial parameters? No: warm-up steps move them.  Use
    #     warmup=0 and capture=False for the gradient check, then the captured variant for (3).
    _mark(rank, "expected gradients gathered")
    sb = G.CapturedTrainStep(model_b, loss_fn, adam(model_b), t, o, warmup=0, capture=False)
    sb.eager_step()
    for p, e in zip(model_b.parameters(), expected):
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.equal(got, e), "rank %d: all-reduced gradient differs from the mean of the local gradients" % rank
    _mark(rank, "all-reduced gradients == mean of the local gradients")
    # (3) graph replay (NCCL inside the graph) == eager multi-rank steps
    model_b.load_state_dict(model_a.state_dict())        # (model_a's BatchNorm buffers moved in (1): both copies start from them)
    model_c.load_state_dict(model_a.state_dict())
    sg = G.CapturedTrainStep(model_b, loss_fn, adam(model_b), t, o, warmup=1, capture=True)
    assert sg.mode == "cuda_graph_replay", sg.mode
    _mark(rank, "step captured")
    se = G.CapturedTrainStep(model_c, loss_fn, adam(model_c), t, o, warmup=1, capture=False)
    for _ in range(2):
        se.eager_step()
    for _ in range(3):
        lg, le = float(sg.replay()), float(se.replay())
        assert lg == le, (lg, le)
    for (k, pa), (_, pb) in zip(model_b.state_dict().items(), model_c.state_dict().items()):
        assert torch.equal(pa, pb), k
    _mark(rank, "three replays == three eager steps")
    # every rank holds the same parameters after the averaged steps
    for p in model_b.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, p.detach())
    # captured graphs hold NCCL kernels: release them before the process group goes away
    del sg, se, sb
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
