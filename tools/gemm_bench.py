#!/usr/bin/env python
"""Dense-stage GEMM micro-benchmark at the BASELINE shapes: sn_gemm_tf32_f32 vs torch (cuBLAS fp32 / TF32)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from surfacenetworks_b200 import _native as N
    dev = torch.device("cuda", 0)
    if "--ncu" in sys.argv:              # short run for ncu --set full: the two BASELINE dense-stage shapes, TS kernel only
        st = torch.cuda.current_stream().cuda_stream
        for M, Nn, K in ((255168, 128, 256), (128000, 256, 128)):
            A = torch.randn(M, K, device=dev)
            B = torch.randn(Nn, K, device=dev) / K ** 0.5
            hi, lo = torch.empty_like(B), torch.empty_like(B)
            N.call("sn_split_tf32_f32", B.data_ptr(), K, Nn, K, hi.data_ptr(), lo.data_ptr(), st)
            bias, R, C = torch.randn(Nn, device=dev), torch.randn(M, Nn, device=dev), torch.empty(M, Nn, device=dev)
            for _ in range(3):
                N.call("sn_gemm_tf32_presplit_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, bias.data_ptr(),
                       R.data_ptr(), Nn, 0, 0, 0, C.data_ptr(), Nn, M, Nn, K, 0, st)
        for R, Nn in ((255168, 256), (128000, 256)):
            A = torch.randn(R, 128, device=dev)
            B = torch.randn(R, Nn, device=dev)
            G, cs = torch.empty(128, Nn, device=dev), torch.empty(128, device=dev)
            nb = N.lib.sn_gemm_tn_tf32_ws_bytes(R, Nn)
            WS = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
            for _ in range(3):
                N.call("sn_gemm_tn_colsum_tf32_f32", A.data_ptr(), 128, B.data_ptr(), Nn, G.data_ptr(), Nn, cs.data_ptr(), R, 128,
                       Nn, 0, WS.data_ptr(), nb, st)
        torch.cuda.synchronize()
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_it(fn, reps=10):
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
        return tot / reps

    for M, Nn, K in ((128000, 128, 256), (255168, 128, 256), (128000, 256, 128), (255168, 256, 128), (16000, 128, 256), (128000, 128, 128)):
        A = torch.randn(M, K, device=dev)
        B = torch.randn(Nn, K, device=dev) / K ** 0.5
        bias = torch.randn(Nn, device=dev)
        R = torch.randn(M, Nn, device=dev)
        C = torch.empty(M, Nn, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        WS = torch.empty(1 << 20, dtype=torch.uint8, device=dev)

        def ours(flags=0, res=True):
            N.call("sn_gemm_tf32_f32", A.data_ptr(), K, B.data_ptr(), K, bias.data_ptr(), R.data_ptr() if res else 0, Nn, 0, 0, 0,
                   C.data_ptr(), Nn, M, Nn, K, flags, WS.data_ptr(), WS.numel(), st)

        flops = 2.0 * M * Nn * K
        bytes_ = 4.0 * (M * K + 2 * M * Nn + Nn * K)
        out = {"M": M, "N": Nn, "K": K, "hbm_floor_us": bytes_ / 6551e3, "hbm_floor_nores_us": (bytes_ - 4.0 * M * Nn) / 6551e3}
        hi, lo = torch.empty_like(B), torch.empty_like(B)
        N.call("sn_split_tf32_f32", B.data_ptr(), K, Nn, K, hi.data_ptr(), lo.data_ptr(), st)

        def presplit(flags=0, res=True):
            N.call("sn_gemm_tf32_presplit_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, bias.data_ptr(),
                   R.data_ptr() if res else 0, Nn, 0, 0, 0, C.data_ptr(), Nn, M, Nn, K, flags, st)

        for name, fn in (("sn_3xtf32", lambda: ours(0)), ("sn_tf32", lambda: ours(N.SN_GEMM_SINGLE_PASS)),
                         ("sn_3xtf32_presplit", lambda: presplit(0)), ("sn_3xtf32_presplit_nores", lambda: presplit(0, False)),
                         ("legacy_ss_3xtf32", lambda: ours(N.SN_GEMM_LEGACY_SS)),
                         ("legacy_ss_3xtf32_nores", lambda: ours(N.SN_GEMM_LEGACY_SS, False)),
                         ("sn_3xtf32_nores", lambda: ours(0, False)),
                         ("sn_3xtf32_nopf", lambda: ours(N.SN_GEMM_NO_L2_PREFETCH)),
                         ("sn_3xtf32_nores_nopf", lambda: ours(N.SN_GEMM_NO_L2_PREFETCH, False)),
                         ("torch_fp32_addmm", lambda: torch.addmm(bias, A, B.t(), out=C).add_(R))):
            ms = time_it(fn)
            out[name] = {"us": ms * 1e3, "TFLOPs": flops / ms / 1e9, "GBps": bytes_ / ms / 1e6}
        torch.backends.cuda.matmul.allow_tf32 = True
        ms = time_it(lambda: torch.addmm(bias, A, B.t(), out=C).add_(R))
        torch.backends.cuda.matmul.allow_tf32 = False
        out["torch_tf32_addmm"] = {"us": ms * 1e3, "TFLOPs": flops / ms / 1e9}
        print(json.dumps(out), flush=True)

    # weight-gradient product G = dY^T Z (split-K over the SMs)
    for R, Nn in ((128000, 256), (255168, 256), (128000, 128)):
        A = torch.randn(R, 128, device=dev)
        B = torch.randn(R, Nn, device=dev)
        G = torch.empty(128, Nn, device=dev)
        nb = N.lib.sn_gemm_tn_tf32_ws_bytes(R, Nn)
        WS = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def tn(flags=0):
            N.call("sn_gemm_tn_tf32_f32", A.data_ptr(), 128, B.data_ptr(), Nn, G.data_ptr(), Nn, R, 128, Nn, flags,
                   WS.data_ptr(), nb, st)

        cs = torch.empty(128, device=dev)

        def tn_cs():
            N.call("sn_gemm_tn_colsum_tf32_f32", A.data_ptr(), 128, B.data_ptr(), Nn, G.data_ptr(), Nn, cs.data_ptr(), R, 128, Nn, 0,
                   WS.data_ptr(), nb, st)

        out = {"op": "gemm_tn", "R": R, "M": 128, "N": Nn, "alg_MB": 4.0 * R * (128 + Nn) / 1e6,
               "hbm_floor_us": 4.0 * R * (128 + Nn) / 6551e3}
        for name, fn in (("sn_3xtf32", lambda: tn(0)), ("sn_3xtf32_colsum", tn_cs), ("legacy_ss_3xtf32", lambda: tn(N.SN_GEMM_LEGACY_SS)),
                         ("sn_3xtf32_nopf", lambda: tn(N.SN_GEMM_NO_L2_PREFETCH)),
                         ("sn_tf32", lambda: tn(N.SN_GEMM_SINGLE_PASS)), ("torch_fp32", lambda: torch.mm(A.t(), B, out=G))):
            ms = time_it(fn)
            out[name] = {"us": ms * 1e3, "GBps": out["alg_MB"] / ms / 1e3}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
