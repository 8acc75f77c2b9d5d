#!/usr/bin/env python
"""Which stage bounds gemm_tf32_ts_kernel?  Times the kernel with individual stages switched off through the debug bits of
`flags` (bits 8..: 1 = epilogue does nothing, 2 = epilogue skips the tensor-memory loads, 4 = no MMAs issued, 8 = split warps
skip the shared-memory read + conversion, 16 = no weight (B) traffic).  Results of these runs are garbage by construction;
only the times matter.  Prints one JSON line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from surfacenetworks_b200 import _native as N
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream

    def t(fn, reps=15):
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

    for M, Nn, K, res in ((16000, 128, 128, True), (64000, 128, 128, True), (128000, 128, 128, False), (128000, 128, 128, True),
                          (128000, 128, 256, True), (255168, 128, 256, False), (255168, 256, 128, True), (512000, 128, 128, True)):
        A = torch.randn(M, K, device=dev)
        B = torch.randn(Nn, K, device=dev) / K ** 0.5
        hi, lo = torch.empty_like(B), torch.empty_like(B)
        N.call("sn_split_tf32_f32", B.data_ptr(), K, Nn, K, hi.data_ptr(), lo.data_ptr(), st)
        bias = torch.randn(Nn, device=dev)
        R = torch.randn(M, Nn, device=dev)
        C = torch.empty(M, Nn, device=dev)

        def run(dbg):
            N.call("sn_gemm_tf32_presplit_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, bias.data_ptr(),
                   R.data_ptr() if res else 0, Nn, 0, 0, 0, C.data_ptr(), Nn, M, Nn, K, dbg << 8, st)

        out = {"M": M, "N": Nn, "K": K, "R": res, "tiles_per_cta": round(M / 128 / 148, 2)}

        def b2b(dbg, n=20):       # n launches back to back (launch overhead hidden; what a graph replay sees)
            run(dbg)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                run(dbg)
            e1.record()
            e1.synchronize()
            return round(e0.elapsed_time(e1) / n * 1e3, 1)
        out["full_b2b"] = b2b(0)
        out["no_epilogue_b2b"] = b2b(1)
        out["only_A_stream_b2b"] = b2b(29)
        for nm, dbg in (("A+mma", 25), ("A+split", 21), ("A+B", 13), ("A+mma+split", 17), ("A+mma+B", 9), ("A+split+B", 5)):
            out[nm + "_b2b"] = b2b(dbg)
        for name, dbg in (("full", 0), ("no_epilogue", 1), ("epilogue_no_tmem_ld", 2), ("no_mma", 4), ("no_split_math", 8),
                          ("no_B_traffic", 16), ("no_mma_no_split", 12), ("no_mma_no_split_no_B", 28), ("only_A_stream", 29), ("A+mma", 25), ("A+split", 21),
                          ("A+B", 13), ("A+mma+split", 17), ("A+mma+B", 9), ("A+split+B", 5)):
            out[name] = t(lambda: run(dbg))
        print(json.dumps(out), flush=True)
        del A, R, C


if __name__ == "__main__":
    main()
