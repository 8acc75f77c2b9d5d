#!/usr/bin/env python
"""GPU operator construction (sn_mesh_dirac_bsr4 / sn_mesh_laplacian_csr) vs the host path it replaces
(vectorised scipy builder + sparse_diag_cat + upload + COO -> CSR32 -> BSR4 conversion) at the BASELINE cfg3 batch size.

    python tools/mesh_ops_bench.py [--meshes 64] [--vertices 2000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--vertices", type=int, default=2000)
    ap.add_argument("--distinct", type=int, default=4)
    args = ap.parse_args()
    from surfacenetworks_b200 import geometry, operators as OP, utils_pt as U
    dev = torch.device("cuda", 0)
    base = [geometry.synth_mesh(args.vertices, s) for s in range(args.distinct)]
    meshes = [base[i % args.distinct] for i in range(args.meshes)]
    nv = max(v.shape[0] for v, _ in meshes)
    nf = max(f.shape[0] for _, f in meshes)
    Vg, Fg = OP.pack_meshes(meshes, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gpu_ms(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"meshes": args.meshes, "vertices": nv, "faces": nf, "host_cores": os.cpu_count(),
           "gpu_dirac_pair_ms": gpu_ms(lambda: OP.build_dirac_operators(Vg, Fg)),
           "gpu_laplacian_ms": gpu_ms(lambda: OP.build_laplacian_operator(Vg, Fg))}
    # the host path for the same batch: per-mesh scipy construction (distinct meshes only, scaled), block-diagonal
    # assembly, upload, conversion
    t0 = time.perf_counter()
    DD = [geometry.build_dirac(v, f) for v, f in base]
    LL = [geometry.build_laplacian(v, f) for v, f in base]
    out["host_build_ms_per_mesh"] = (time.perf_counter() - t0) * 1e3 / len(base)
    t0 = time.perf_counter()
    Dh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(DD[i % len(base)][0]) for i in range(args.meshes)], 4 * nf, 4 * nv)
    DAh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(DD[i % len(base)][1]) for i in range(args.meshes)], 4 * nv, 4 * nf)
    Lh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(LL[i % len(base)]) for i in range(args.meshes)], nv, nv)
    out["host_assemble_ms"] = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops = (OP.Bsr4Operator.from_torch_coo(Dh.to(dev)), OP.Bsr4Operator.from_torch_coo(DAh.to(dev)),
           OP.CsrOperator.from_torch_coo(Lh.to(dev)))
    torch.cuda.synchronize()
    out["upload_convert_ms"] = (time.perf_counter() - t0) * 1e3
    out["host_path_total_ms"] = out["host_build_ms_per_mesh"] * args.meshes + out["host_assemble_ms"] + out["upload_convert_ms"]
    out["blocks"] = ops[0].n_blocks
    print(json.dumps(out))


if __name__ == "__main__":
    main()
