#!/usr/bin/env python
"""What does each kernel family cost INSIDE the captured training step?  (ncu's per-launch times are cold-cache and
serialised; event pairs around eager launches add their own overhead.)  Knock-out measurement: the step is captured
once per entry in --skip with the named C entry points turned into no-ops (their outputs keep the values of the last
eager step, so everything stays finite; the results of such a step are garbage by construction -- only the replay time
matters), and the difference to the complete step is that family's cost in situ, launch gaps and cache state included.

    python tools/ablate_step.py --skip sn_avg_fold_fwd_f32 sn_bn_fold_fwd_f32,sn_bn_fold_bwd_f32 ...
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--vertices", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--skip", nargs="*", default=[], help="comma-separated entry-point groups, one capture per group")
    args = ap.parse_args()
    from surfacenetworks_b200 import _native as N, graph as G, models as M, operators as OP, workloads as W
    dev = torch.device("cuda")
    meshes = W.make_mesh_ops(args.vertices, range(args.meshes))
    host = W.arap_batch(meshes, seed=0)
    t = {k: host[k].to(dev) for k in ("inputs", "targets", "mask")}
    o = {"Di": OP.Bsr4Operator.from_torch_coo(host["Di"].to(dev)), "DiA": OP.Bsr4Operator.from_torch_coo(host["DiA"].to(dev))}
    B = args.meshes

    def loss_fn(m, t, o):
        return M.arap_loss(m(o["Di"], o["DiA"], t["mask"], t["inputs"]), t["targets"], t["mask"], B)

    real_call = N.call
    skipped = set()
    counts = {}

    def call(name, *a, **kw):
        if name in skipped:
            counts[name] = counts.get(name, 0) + 1
            return N.SN_OK
        return real_call(name, *a, **kw)

    # modules bind `_native` as N and call N.call(...): patching the attribute reaches all of them
    N.call = call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for group in ["", ""] + [g if g != "full" else "" for g in args.skip]:
        torch.manual_seed(0)
        model = M.ArapDirModel().to(dev).train()
        opt = torch.optim.Adam(model.parameters(), 1e-3, weight_decay=1e-5, fused=True, capturable=True)
        skipped.clear()
        step = G.CapturedTrainStep(model, loss_fn, opt, t, o, warmup=2, capture=False)
        skipped.update(x for x in group.split(",") if x)
        counts.clear()
        step._capture()
        n_skipped = sum(counts.values()) // 3           # _capture runs the step 2 + 1 times
        if step.graph is None:
            out[group or "full"] = step.mode
            continue
        for _ in range(3):
            step.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            step.replay()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        key = group or "full"
        while key in out and key == "full":
            key += "'"
        out[key] = {"ms_per_step": round(ms, 4), "launches_skipped": n_skipped}
        if group and "full" in out:
            d = out["full"]["ms_per_step"] - ms
            out[key]["cost_ms"] = round(d, 4)
            out[key]["cost_us_per_launch"] = round(1e3 * d / max(n_skipped, 1), 2)
        print(json.dumps({key: out[key]}), flush=True)
        del step, model, opt


if __name__ == "__main__":
    main()
