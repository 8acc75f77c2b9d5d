#!/usr/bin/env python
"""One eager ArapDirModel training step at the cfg3 size between cudaProfilerStart / Stop (for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv ...`: the launch list of exactly one step).

    python tools/step_once.py [--meshes 64] [--vertices 2000] [--warm 2]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--vertices", type=int, default=2000)
    ap.add_argument("--warm", type=int, default=2)
    args = ap.parse_args()
    from surfacenetworks_b200 import graph as G, models as M, operators as OP, workloads as W
    dev = torch.device("cuda")
    meshes = W.make_mesh_ops(args.vertices, range(args.meshes))
    host = W.arap_batch(meshes, seed=0)
    t = {k: host[k].to(dev) for k in ("inputs", "targets", "mask")}
    o = {"Di": OP.Bsr4Operator.from_torch_coo(host["Di"].to(dev)), "DiA": OP.Bsr4Operator.from_torch_coo(host["DiA"].to(dev))}
    torch.manual_seed(0)
    model = M.ArapDirModel().to(dev).train()
    opt = torch.optim.Adam(model.parameters(), 1e-3, weight_decay=1e-5, fused=True, capturable=True)
    B = args.meshes

    def loss_fn(m, t, o):
        return M.arap_loss(m(o["Di"], o["DiA"], t["mask"], t["inputs"]), t["targets"], t["mask"], B)

    step = G.CapturedTrainStep(model, loss_fn, opt, t, o, warmup=args.warm, capture=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = step.eager_step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", float(loss.detach()))


if __name__ == "__main__":
    main()
