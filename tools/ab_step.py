#!/usr/bin/env python
"""In-process A/B of a module-level switch on the captured training step (cfg3 size).  Box-to-box and run-to-run
clock differences (power capping: 1867 ... 1957 MHz between bench runs) are +-2 % -- larger than most single changes --
so both variants are captured in ONE process and replayed alternately, round after round.

    python tools/ab_step.py --toggle module.SWITCH [--rounds 12] [--steps 10]
"""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--toggle", default=None, help="module.ATTRIBUTE inside surfacenetworks_b200 (A: True, B: False)")
    ap.add_argument("--toggle-kw", default=None, help="boolean keyword of CapturedTrainStep (A: True, B: False); works "
                    "under torchrun (one process per GPU, max over ranks)")
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--vertices", type=int, default=2000)
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from surfacenetworks_b200 import dist as D, graph as G, models as M, operators as OP, workloads as W
    import torch.distributed as tdist
    rank, local_rank, world = D.init_from_env()
    torch.cuda.set_device(local_rank)
    mod = attr = None
    if args.toggle:
        modname, attr = args.toggle.rsplit(".", 1)
        mod = importlib.import_module("surfacenetworks_b200." + modname)
    dev = torch.device("cuda", local_rank)
    meshes = W.make_mesh_ops(args.vertices, range(rank * args.meshes, (rank + 1) * args.meshes))
    host = W.arap_batch(meshes, seed=rank)
    t = {k: host[k].to(dev) for k in ("inputs", "targets", "mask")}
    o = {"Di": OP.Bsr4Operator.from_torch_coo(host["Di"].to(dev)), "DiA": OP.Bsr4Operator.from_torch_coo(host["DiA"].to(dev))}
    B = args.meshes

    def loss_fn(m, t, o):
        return M.arap_loss(m(o["Di"], o["DiA"], t["mask"], t["inputs"]), t["targets"], t["mask"], B)

    steps = {}
    for name, val in (("A", True), ("B", False)):
        if mod is not None:
            setattr(mod, attr, val)
        kw = {args.toggle_kw: val} if args.toggle_kw else {}
        torch.manual_seed(0)
        model = M.ArapDirModel().to(dev).train()
        D.broadcast_module(model)
        opt = torch.optim.Adam(model.parameters(), 1e-3, weight_decay=1e-5, fused=True, capturable=True)
        steps[name] = G.CapturedTrainStep(model, loss_fn, opt, t, o, warmup=2, **kw)
        assert steps[name].graph is not None, steps[name].mode
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = {"A": [], "B": []}
    for r in range(args.rounds):
        for name in (("A", "B") if r % 2 == 0 else ("B", "A")):
            st = steps[name]
            st.replay()
            if world > 1:
                tdist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                st.replay()
            e1.record()
            e1.synchronize()
            v = e0.elapsed_time(e1) / args.steps
            if world > 1:
                tv = torch.tensor([v], device=dev, dtype=torch.float64)
                tdist.all_reduce(tv, op=tdist.ReduceOp.MAX)
                v = float(tv.item())
            ms[name].append(v)
    losses = {name: float(steps[name].loss) for name in ("A", "B")}
    if world > 1:                       # every rank leaves through here (a rank that returns early hangs the others' teardown)
        del steps
        torch.cuda.synchronize()
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    out = {"toggle": args.toggle or args.toggle_kw, "world": world}
    for name in ("A", "B"):
        v = sorted(ms[name])
        out[name] = {"median_ms": round(v[len(v) // 2], 4), "min_ms": round(v[0], 4), "max_ms": round(v[-1], 4),
                     "loss": losses[name]}
    out["A_minus_B_ms"] = round(out["A"]["median_ms"] - out["B"]["median_ms"], 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
