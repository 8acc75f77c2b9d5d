#!/usr/bin/env python
"""One DirResNet2(128) block, forward + backward, at the BASELINE cfg3 size -- a short target for ncu captures of the
forward (plain) and backward (epilogue) row-group SpMM launches:

    ncu --set full --clock-control none --import-source on -k regex:rowgroup -c 8 -o out python tools/block_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from surfacenetworks_b200 import operators as OP, utils_pt as U, workloads as W
    dev = torch.device("cuda", 0)
    base = W.make_mesh_ops(2000, range(4))
    meshes = [base[i % 4] for i in range(64)]
    b = W.arap_batch(meshes, 0)
    D, DA = OP.as_bsr4(b["Di"].to(dev)), OP.as_bsr4(b["DiA"].to(dev))
    D.T, DA.T
    blk = U.DirResNet2(128).to(dev).train()
    v = torch.randn(64, b["num_vertices"], 128, device=dev, requires_grad=True)
    f = torch.randn(64, b["num_faces"], 128, device=dev, requires_grad=True)
    for _ in range(2):
        vo, fo = blk(D, DA, v, f)
        (vo.sum() + fo.sum()).backward()
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
