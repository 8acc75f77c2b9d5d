#!/usr/bin/env python
"""Operator-only micro-benchmark at the BASELINE sizes: Dirac D / D* and Laplacian SpMM.

    python tools/spmm_bench.py [--meshes 64] [--vertices 2000] [--features 128] [--reps 20] [--variants ...]

Prints one JSON line per (operator, variant): time per launch (CUDA events, L2 flushed between launches when the
working set is smaller than L2), canonical algorithmic GB/s and fraction of MEASURED_PEAKS.json's hbm_gbs.
Used under ncu for the profiles/ captures.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--vertices", type=int, default=2000)
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--distinct", type=int, default=8, help="distinct meshes (repeated to fill the batch)")
    ap.add_argument("--ops", default="D,Dstar,L")
    ap.add_argument("--variants", default="rg,rg1,rg2,rg3,rg5,rg9,smem,direct",
                    help="rg[N] = row-group kernel (tuning variant N), smem = cp.async streaming kernel (BSR4 only), "
                         "direct = first-generation direct-gather kernel; +elu = ELU on load")
    ap.add_argument("--order", default="none", help="none | bisect | morton | morton_xy | rcm: renumber every mesh with geometry.locality_order")
    args = ap.parse_args()
    from surfacenetworks_b200 import geometry, operators as OP, workloads as W
    dev = torch.device("cuda", 0)
    peak = 6551.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    base = W.make_mesh_ops(args.vertices, range(args.distinct), order=args.order)
    meshes = [base[i % args.distinct] for i in range(args.meshes)]
    C = args.features
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_it(fn):
        for _ in range(3):
            fn()
        tot = 0.0
        best = 1e30
        for _ in range(args.reps):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            t = e0.elapsed_time(e1)
            tot += t
            best = min(best, t)
        return tot / args.reps, best

    ops = {}
    if "D" in args.ops.split(",") or "Dstar" in args.ops.split(","):
        b = W.arap_batch(meshes, 0)
        ops["D"] = OP.as_bsr4(b["Di"].to(dev))
        ops["Dstar"] = OP.as_bsr4(b["DiA"].to(dev))
    if "L" in args.ops.split(","):
        ops["L"] = OP.as_csr(W.lap_batch(meshes)["L"].to(dev))
    # copy bandwidth of the same box, same timing method, for context
    a = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    c = torch.empty_like(a)
    ms, best = time_it(lambda: c.copy_(a))
    print(json.dumps({"op": "copy 256MB", "us": ms * 1e3, "GBps": 2 * a.numel() * 4 / ms / 1e6}))
    eye = torch.sparse_coo_tensor(torch.arange(32).repeat(2, 1), torch.ones(32), (32, 32)).coalesce()
    tiny, xt = OP.as_csr(eye.to(dev)), torch.randn(32, 16, device=dev)
    ms, best = time_it(lambda: tiny.apply(xt))
    print(json.dumps({"op": "launch floor (32-row operator)", "us": ms * 1e3, "us_best": best * 1e3}))
    for name in args.ops.split(","):
        op = ops[name]
        ncols = op.n_bcols if op.kind == "bsr4" else op.n_cols
        X = torch.randn(ncols, C, device=dev)
        Y = op.apply(X)
        variants = [v for v in args.variants.split(",") if op.kind == "bsr4" or not v.startswith("smem")]
        for v in variants:
            kw = {"elu_input": "elu" in v, "direct_gather": v.startswith("direct")}
            if op.kind == "bsr4":
                kw["smem_stream"] = v.startswith("smem")
            if v.startswith("rg") and v[2:].replace("+elu", "").isdigit():
                kw["variant"] = int(v[2:].replace("+elu", ""))
            ms, best = time_it(lambda: op.apply(X, out=Y, **kw))
            gb = op.algorithmic_bytes(C) / 1e9
            print(json.dumps({"op": name, "variant": v, "order": args.order, "rows": op.n_brows if op.kind == "bsr4" else op.n_rows,
                              "C": C, "us": ms * 1e3, "us_best": best * 1e3, "alg_MB": gb * 1e3,
                              "GBps": gb / (ms / 1e3), "frac_of_measured_peak": gb / (ms / 1e3) / peak,
                              "GFLOPs": op.flops(C) / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
