#!/usr/bin/env python
"""A/B timings of the round-2 producer-side fusions at the cfg3 size (64 meshes x 2000 V, C = 128):
the passes a Dirac block used to run (GEMM, sn_elu_f32, sn_elu_colstats_f32, SpMM, sn_colstats_f32) against the fused
launches (GEMM + activation + statistics epilogue, SpMM + statistics store path).  L2 is flushed between timed launches.
Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from surfacenetworks_b200 import _native as N, fused, operators as OP, workloads as W, ops
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return round(tot / reps * 1e3, 1)

    out = {}
    C = 128
    st = torch.cuda.current_stream().cuda_stream
    if "--ncu-act" in sys.argv:          # three launches of the activation-epilogue GEMM only (for ncu -k ... -s 1 -c 1)
        M = 255168
        A = torch.randn(M, 2 * C, device=dev)
        Wt = torch.randn(C, 2 * C, device=dev) / 16
        hi, lo = torch.empty_like(Wt), torch.empty_like(Wt)
        N.call("sn_split_tf32_f32", Wt.data_ptr(), 2 * C, C, 2 * C, hi.data_ptr(), lo.data_ptr(), st)
        bias = torch.randn(C, device=dev)
        Zn = torch.empty(M, 2 * C, device=dev)
        mean, var = torch.empty(C, device=dev), torch.empty(C, device=dev)
        for _ in range(3):
            fused.gemm_tf32_act(A, hi, lo, bias=bias, act_out=Zn[:, :C], mean=mean, var=var, want_raw=False)
        torch.cuda.synchronize()
        return
    if "--ncu-spmm-stats" in sys.argv:
        meshes = W.make_mesh_ops(2000, range(64))
        b = W.arap_batch(meshes, 0)
        D = OP.Bsr4Operator.from_torch_coo(b["Di"].to(dev))
        X = torch.randn(D.n_bcols, 2 * C, device=dev)
        Z = torch.empty(D.n_brows, 2 * C, device=dev)
        mean, var = torch.empty(C, device=dev), torch.empty(C, device=dev)
        for _ in range(3):
            D.apply_stats(X[:, :C], Z[:, C:], mean, var)
        torch.cuda.synchronize()
        return
    for M in (255168, 128000):
        A = torch.randn(M, 2 * C, device=dev)
        Wt = torch.randn(C, 2 * C, device=dev) / 16
        hi, lo = torch.empty_like(Wt), torch.empty_like(Wt)
        N.call("sn_split_tf32_f32", Wt.data_ptr(), 2 * C, C, 2 * C, hi.data_ptr(), lo.data_ptr(), st)
        bias = torch.randn(C, device=dev)
        R = torch.randn(M, C, device=dev)
        Y = torch.empty(M, C, device=dev)
        Zn = torch.empty(M, 2 * C, device=dev)
        act = torch.empty(M, C, device=dev)
        mean, var = torch.empty(C, device=dev), torch.empty(C, device=dev)
        k = "M%d" % M
        out[k + " gemm"] = t(lambda: fused.gemm_tf32(A, hi, bias=bias, B_lo=lo, out=Y))
        out[k + " gemm+R"] = t(lambda: fused.gemm_tf32(A, hi, bias=bias, R=R, B_lo=lo, out=Y))
        out[k + " gemm_act(act only, stats, strided)"] = t(lambda: fused.gemm_tf32_act(A, hi, lo, bias=bias, act_out=Zn[:, :C], mean=mean, var=var, want_raw=False))
        out[k + " gemm_act(act only, no stats)"] = t(lambda: fused.gemm_tf32_act(A, hi, lo, bias=bias, act_out=act, want_raw=False))
        out[k + " gemm_act(raw+act, stats, +R)"] = t(lambda: fused.gemm_tf32_act(A, hi, lo, bias=bias, R=R, out=Y, act_out=act, mean=mean, var=var))
        out[k + " elu"] = t(lambda: ops.elu_into(Y, act))
        out[k + " elu_colstats"] = t(lambda: fused.elu_colstats(Y, Zn[:, :C], mean, var))
        out[k + " colstats(right half)"] = t(lambda: fused.colstats(Zn[:, C:], mean, var))
        del A, R, Y, Zn, act
    meshes = W.make_mesh_ops(2000, range(64))
    b = W.arap_batch(meshes, 0)
    D, DA = OP.Bsr4Operator.from_torch_coo(b["Di"].to(dev)), OP.Bsr4Operator.from_torch_coo(b["DiA"].to(dev))
    for name, op in (("D", D), ("D*", DA)):
        X = torch.randn(op.n_bcols, 2 * C, device=dev)
        Z = torch.empty(op.n_brows, 2 * C, device=dev)
        mean, var = torch.empty(C, device=dev), torch.empty(C, device=dev)
        out[name + " spmm"] = t(lambda: op.apply(X[:, :C], out=Z[:, C:]))
        out[name + " spmm+stats"] = t(lambda: op.apply_stats(X[:, :C], Z[:, C:], mean, var))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
