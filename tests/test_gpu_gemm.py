"""tcgen05 TF32 GEMM (sn_gemm_tf32_f32) vs an fp64 reference.

Tolerance: the reference Linear is fp32; 3xTF32 (default) must stay within 8 * eps_fp32 * K-independent bound
|err| <= 2e-6 * (|A| |B|^T + |bias| + |rscale R|); the single-pass mode is checked at 2e-3 of the same magnitude.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_gemm(A, B, bias=None, R=None, rscale=None, single_pass=False, out=None):
    from surfacenetworks_b200 import _native as N
    M, K = A.shape
    Nn = B.shape[0]
    C = torch.empty(M, Nn, device=DEV) if out is None else out
    p = lambda t: 0 if t is None else t.data_ptr()
    wsb = N.lib.sn_gemm_tf32_ws_bytes(Nn, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    N.call("sn_gemm_tf32_f32", p(A), A.stride(0), p(B), B.stride(0), p(bias), p(R), 0 if R is None else R.stride(0),
           p(rscale), 0, 0, p(C), C.stride(0), M, Nn, K, N.SN_GEMM_SINGLE_PASS if single_pass else 0, p(ws), wsb,
           torch.cuda.current_stream().cuda_stream)
    return C


def reference(A, B, bias, R, rscale):
    A64, B64 = A.double(), B.double()
    y = A64 @ B64.t()
    mag = A64.abs() @ B64.abs().t()
    if bias is not None:
        y = y + bias.double()
        mag = mag + bias.double().abs()
    if R is not None:
        s = rscale.double() if rscale is not None else torch.ones(B.shape[0], device=DEV, dtype=torch.float64)
        y = y + s * R.double()
        mag = mag + (s * R.double()).abs()
    return y, mag


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1000, 128, 256), (4096 + 77, 256, 128), (300, 64, 64), (20000, 128, 128),
                                   (1000, 256, 256), (3000, 256, 64), (700, 64, 128), (148 * 128 * 3 + 5, 128, 512)])
@pytest.mark.parametrize("mode", ["plain", "bias", "full"])
def test_gemm_matches_fp64(M, N, K, mode):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g) * torch.exp(torch.randn(M, 1, device=DEV, generator=g))
    B = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g) if mode != "plain" else None
    R = torch.randn(M, N, device=DEV, generator=g) if mode == "full" else None
    rscale = torch.randn(N, device=DEV, generator=g) if mode == "full" else None
    y64, mag = reference(A, B, bias, R, rscale)
    C = run_gemm(A, B, bias, R, rscale)
    torch.cuda.synchronize()
    err = (C.double() - y64).abs()
    assert torch.all(err <= 2e-6 * mag + 1e-30), "3xTF32: max err/mag %g" % float((err / (mag + 1e-300)).max())
    C1 = run_gemm(A, B, bias, R, rscale, single_pass=True)
    err1 = (C1.double() - y64).abs()
    assert torch.all(err1 <= 2e-3 * mag + 1e-30), "TF32: max err/mag %g" % float((err1 / (mag + 1e-300)).max())
    # the fp32 cuBLAS result (what the reference's nn.Linear computes on a GPU) is at most 8x closer to fp64 than we are
    ref32 = torch.nn.functional.linear(A, B, bias)
    if R is not None:
        ref32 = ref32 + rscale * R
    e32 = float(((ref32.double() - y64).abs() / (mag + 1e-300)).max())
    assert float((err / (mag + 1e-300)).max()) <= max(8 * e32, 2e-6)


@pytest.mark.parametrize("M,N,K", [(5000, 128, 256), (5000, 256, 128), (2000, 256, 256), (900, 64, 96)])
def test_gemm_presplit_and_legacy_kernel_agree(M, N, K):
    """sn_gemm_tf32_presplit_f32 (weights split by the caller / by sn_bn_fold_*) is bit-identical to sn_gemm_tf32_f32, and the
    TS-mode kernel (A operand in tensor memory) matches the round-1 shared-memory kernel within the 3xTF32 bound."""
    from surfacenetworks_b200 import _native as Nt
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + K)
    A = torch.randn(M, K, device=DEV, generator=g)
    B = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    R = torch.randn(M, N, device=DEV, generator=g)
    rs = torch.randn(N, device=DEV, generator=g)
    st = torch.cuda.current_stream().cuda_stream
    C0 = run_gemm(A, B, bias, R, rs)
    hi, lo = torch.empty_like(B), torch.empty_like(B)
    Nt.call("sn_split_tf32_f32", B.data_ptr(), K, N, K, hi.data_ptr(), lo.data_ptr(), st)
    assert torch.all((hi + lo - B).abs() <= B.abs() * 2.0 ** -22)   # lo is rounded to tf32 as well (the tensor core truncates)
    assert torch.all((lo.view(torch.int32) & 0x1fff) == 0)
    assert torch.all((hi.view(torch.int32) & 0x1fff) == 0)            # tf32: low 13 mantissa bits clear
    C1 = torch.empty_like(C0)
    Nt.call("sn_gemm_tf32_presplit_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, bias.data_ptr(), R.data_ptr(), N,
            rs.data_ptr(), 0, 0, C1.data_ptr(), N, M, N, K, 0, st)
    assert torch.equal(C0, C1)
    C2 = torch.empty_like(C0)
    wsb = Nt.lib.sn_gemm_tf32_ws_bytes(N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    Nt.call("sn_gemm_tf32_f32", A.data_ptr(), K, B.data_ptr(), K, bias.data_ptr(), R.data_ptr(), N, rs.data_ptr(), 0, 0,
            C2.data_ptr(), N, M, N, K, Nt.SN_GEMM_LEGACY_SS, ws.data_ptr(), wsb, st)
    y64, mag = reference(A, B, bias, R, rs)
    assert torch.all((C2.double() - y64).abs() <= 2e-6 * mag)
    assert torch.all((C0.double() - C2.double()).abs() <= 4e-6 * mag)


def test_gemm_strided_operands_and_repeatability():
    """A = the [rows, 2C] concat buffer view, C written into a wider buffer; two launches give identical bits."""
    g = torch.Generator(device=DEV).manual_seed(5)
    M, N, K = 3000, 128, 256
    Zbuf = torch.randn(M, K + 64, device=DEV, generator=g)
    A = Zbuf[:, :K]
    B = torch.randn(N, K, device=DEV, generator=g) / 16
    out = torch.zeros(M, 2 * N, device=DEV)
    run_gemm(A, B, out=out[:, N:])
    y64, mag = reference(A, B, None, None, None)
    assert torch.all((out[:, N:].double() - y64).abs() <= 2e-6 * mag)
    assert torch.all(out[:, :N] == 0)
    again = torch.zeros(M, 2 * N, device=DEV)
    run_gemm(A, B, out=again[:, N:])
    assert torch.equal(out, again)


def test_gemm_argument_errors():
    from surfacenetworks_b200 import _native as N
    A = torch.zeros(128, 64, device=DEV)
    B = torch.zeros(100, 64, device=DEV)
    with pytest.raises(N.SurfnetError):
        run_gemm(A, B)                       # N = 100 unsupported
    with pytest.raises(N.SurfnetError):
        run_gemm(torch.zeros(128, 48, device=DEV), torch.zeros(128, 48, device=DEV))   # K % 32 != 0


# ------------------------------------------------------------------------------------------- fused BN + Linear stage
def _composite64(Z, gamma, beta, W, b, residual, rm, rv, training, eps=1e-5, momentum=0.1):
    """GraphConv1x1("pre") in float64 with torch ops (the reference arithmetic: BatchNorm1d then Linear)."""
    import torch.nn.functional as F
    y = F.linear(F.batch_norm(Z, rm, rv, gamma, beta, training, momentum, eps), W, b)
    return y if residual is None else y + residual


@pytest.mark.parametrize("C,rows", [(128, 5000), (64, 1111), (256, 700)])
@pytest.mark.parametrize("training,with_res", [(True, True), (True, False), (False, True)])
def test_bn_linear_fused_vs_float64(C, rows, training, with_res):
    """fused.bn_linear (colstats + fold + tcgen05 GEMM, hand-derived backward) vs BatchNorm1d + Linear in float64.
    Tolerance 1e-4 * (|ref| + max|ref|): fp32 statistics and 3xTF32 products, no normalised copy of Z."""
    import torch.nn as nn
    from surfacenetworks_b200 import fused
    g = torch.Generator(device=DEV).manual_seed(C + rows)
    Z = (torch.randn(rows, 2 * C, device=DEV, generator=g) * (1 + torch.rand(2 * C, device=DEV, generator=g) * 3)
         + torch.randn(2 * C, device=DEV, generator=g)).requires_grad_(True)
    res = torch.randn(rows, C, device=DEV, generator=g).requires_grad_(True) if with_res else None
    bn, fc = nn.BatchNorm1d(2 * C).to(DEV), nn.Linear(2 * C, C).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
        bn.running_mean.normal_(0, 0.5)
        bn.running_var.uniform_(0.5, 2.0)
    bn.train(training)
    assert fused.fused_supported(Z, fc.weight)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    wout = torch.randn(rows, C, device=DEV, generator=g)
    Y = fused.bn_linear(Z, bn, fc, res)
    (Y * wout).sum().backward()
    # float64 reference
    Z64 = Z.detach().double().requires_grad_(True)
    r64 = res.detach().double().requires_grad_(True) if with_res else None
    P = [t.detach().double().requires_grad_(True) for t in (bn.weight, bn.bias, fc.weight, fc.bias)]
    rm, rv = rm0.double(), rv0.double()
    Y64 = _composite64(Z64, P[0], P[1], P[2], P[3], r64, rm, rv, training)
    (Y64 * wout.double()).sum().backward()

    def close(a, b, what, tol=1e-4):
        b = b.double()
        err = (a.double() - b).abs()
        assert torch.all(err <= tol * (b.abs() + b.abs().max())), "%s: max err %g (scale %g)" % (what, float(err.max()), float(b.abs().max()))

    close(Y, Y64, "Y")
    close(Z.grad, Z64.grad, "dZ")
    for name, p, p64 in zip(("dgamma", "dbeta", "dW", "db"), (bn.weight, bn.bias, fc.weight, fc.bias), P):
        close(p.grad, p64.grad, name, 2e-4)
    if with_res:
        close(res.grad, r64.grad, "dres")
    if training:
        close(bn.running_mean, rm, "running_mean", 1e-5)
        close(bn.running_var, rv, "running_var", 1e-5)
        assert int(bn.num_batches_tracked) == 1


def test_colstats_matches_float64():
    from surfacenetworks_b200 import fused
    g = torch.Generator(device=DEV).manual_seed(3)
    for rows, C in ((1, 256), (37, 128), (100000, 256), (4099, 512), (999, 16)):
        Zb = torch.randn(rows, C + 12, device=DEV, generator=g) * 5 + 3
        Z = Zb[:, :C]
        mean, var = fused.colstats(Z)
        m64 = Z.double().mean(0)
        v64 = Z.double().var(0, unbiased=False)
        assert torch.allclose(mean.double(), m64, rtol=1e-6, atol=1e-6)
        assert torch.allclose(var.double(), v64, rtol=2e-5, atol=1e-6), (rows, C, float((var.double() - v64).abs().max()))
        if rows == 1:
            assert torch.all(var == 0)                                   # constant column -> exactly zero
        m2, v2 = fused.colstats(Z)
        assert torch.equal(mean, m2) and torch.equal(var, v2)          # deterministic


@pytest.mark.parametrize("kind", ["lap", "dir", "avg"])
def test_blocks_width128_vs_oracle(kind):
    """One LapResNet2(128) / DirResNet2(128) block -- the fused tensor-core stage inside the real layer -- against the
    oracle port run live on the CPU (fp32), forward and all gradients; tolerance 2e-4 * (|ref| + max|ref|)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from det import det_fill, det_tensor
    from oracle import layers as O
    from surfacenetworks_b200 import utils_pt as U, workloads as W
    C = 128
    meshes = W.make_mesh_ops(150, [0, 1]) + W.make_mesh_ops(140, [2])
    B = len(meshes)
    host = W.arap_batch(meshes, 0, dirac=(kind != "lap"))
    nv, nf = host["num_vertices"], host["num_faces"]
    block = det_fill({"lap": U.LapResNet2, "dir": U.DirResNet2, "avg": U.AvgResNet2}[kind](C), 5, gain=0.5)
    P = {}
    for k, v in block.state_dict().items():
        v = v.clone()
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
        P[k] = v
    x = det_tensor((B, nv, C), 1)
    f = det_tensor((B, nf, C), 2)
    wv, wf = det_tensor((B, nv, C), 3), det_tensor((B, nf, C), 4)
    xc, fc_ = x.clone().requires_grad_(True), f.clone().requires_grad_(True)
    if kind == "lap":
        ref = (O.lap_resnet2(P, host["L"], xc),)
        (ref[0] * wv).sum().backward()
    elif kind == "avg":
        ref = (O.avg_resnet2(P, host["mask"], xc),)
        (ref[0] * wv).sum().backward()
    else:
        ref = O.dir_resnet2(P, host["Di"], host["DiA"], xc, fc_)
        ((ref[0] * wv).sum() + (ref[1] * wf).sum()).backward()
    blk = block.to(DEV).train()
    xg, fg = x.to(DEV).requires_grad_(True), f.to(DEV).requires_grad_(True)
    if kind == "lap":
        out = (blk(host["L"].to(DEV), None, xg),)
        (out[0] * wv.to(DEV)).sum().backward()
    elif kind == "avg":
        out = (blk(None, host["mask"].to(DEV), xg),)
        (out[0] * wv.to(DEV)).sum().backward()
    else:
        out = blk(host["Di"].to(DEV), host["DiA"].to(DEV), xg, fg)
        ((out[0] * wv.to(DEV)).sum() + (out[1] * wf.to(DEV)).sum()).backward()

    def close(a, b, what, tol=2e-4, floor=0.0):
        a, b = a.detach().cpu().double(), b.detach().double()
        err = (a - b).abs()
        assert torch.all(err <= tol * (b.abs() + max(float(b.abs().max()), floor))), "%s: max err %g scale %g" % (what, float(err.max()), float(b.abs().max()))

    for i, (o, r) in enumerate(zip(out, ref)):
        close(o, r, "out%d" % i)
    close(xg.grad, xc.grad, "dx")
    if kind == "dir":
        close(fg.grad, fc_.grad, "df")
    gscale = max(float(P[k].grad.abs().max()) for k, _ in blk.named_parameters())
    for k, p in blk.named_parameters():
        close(p.grad, P[k].grad, "grad " + k, 1e-3, floor=1e-2 * gscale)


# ------------------------------------------------------------------------------------------- G = A^T B (split-K, MN-major)
def run_gemm_tn(A, B, single_pass=False):
    from surfacenetworks_b200 import _native as N
    R, M = A.shape
    Nn = B.shape[1]
    G = torch.empty(M, Nn, device=DEV)
    wsb = N.lib.sn_gemm_tn_tf32_ws_bytes(R, Nn)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    N.call("sn_gemm_tn_tf32_f32", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), G.data_ptr(), G.stride(0), R, M, Nn,
           N.SN_GEMM_SINGLE_PASS if single_pass else 0, ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
    return G


@pytest.mark.parametrize("R,N", [(32, 256), (8, 64), (1000, 256), (4133, 128), (128000, 256), (255168, 256)])
def test_gemm_tn_matches_fp64(R, N):
    g = torch.Generator(device=DEV).manual_seed(R + N)
    A = torch.randn(R, 128, device=DEV, generator=g)
    Bb = torch.randn(R, N + 32, device=DEV, generator=g) + 0.5
    B = Bb[:, :N]                                   # strided operand (the concat buffer case)
    G = run_gemm_tn(A, B)
    torch.cuda.synchronize()
    G64 = A.double().t() @ B.double()
    mag = A.double().abs().t() @ B.double().abs()
    err = (G.double() - G64).abs()
    assert torch.all(err <= 2e-6 * mag), "3xTF32 G: max err/mag %g" % float((err / mag).max())
    assert torch.equal(G, run_gemm_tn(A, B))          # deterministic split-K reduction
    G1 = run_gemm_tn(A, B, single_pass=True)
    assert torch.all((G1.double() - G64).abs() <= 2e-3 * mag)


@pytest.mark.parametrize("R,N", [(32, 256), (5, 32), (1000, 256), (4133, 128), (128000, 256), (255168, 128)])
def test_gemm_tn_colsum_and_legacy_kernel(R, N):
    """sn_gemm_tn_colsum_tf32_f32: same G bits as sn_gemm_tn_tf32_f32, column sums of A within the fp32 summation bound
    of the fp64 sums, deterministic; the TS kernel (dY through tensor memory) agrees with the round-1 kernel."""
    from surfacenetworks_b200 import _native as Nt
    g = torch.Generator(device=DEV).manual_seed(7 * R + N)
    A = torch.randn(R, 128, device=DEV, generator=g) + 0.25
    B = torch.randn(R, N, device=DEV, generator=g)
    G0 = run_gemm_tn(A, B)
    wsb = Nt.lib.sn_gemm_tn_tf32_ws_bytes(R, N)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream

    def with_colsum():
        G, cs = torch.empty(128, N, device=DEV), torch.empty(128, device=DEV)
        Nt.call("sn_gemm_tn_colsum_tf32_f32", A.data_ptr(), 128, B.data_ptr(), N, G.data_ptr(), N, cs.data_ptr(), R, 128, N, 0,
                ws.data_ptr(), wsb, st)
        return G, cs

    G1, cs1 = with_colsum()
    G2, cs2 = with_colsum()
    assert torch.equal(G0, G1) and torch.equal(G1, G2) and torch.equal(cs1, cs2)
    cs64 = A.double().sum(0)
    mag = A.double().abs().sum(0)
    # sequential fp32 summation of <= ceil(R / 148) rows per CTA, then <= 148 partials: error <= (n + 148) eps |A| / 2
    n_seq = (R + 147) // 148 + 32 + 148
    assert torch.all((cs1.double() - cs64).abs() <= n_seq * 6e-8 * mag + 1e-30), float(((cs1.double() - cs64).abs() / mag).max())
    Gl = torch.empty(128, N, device=DEV)
    Nt.call("sn_gemm_tn_tf32_f32", A.data_ptr(), 128, B.data_ptr(), N, Gl.data_ptr(), N, R, 128, N, Nt.SN_GEMM_LEGACY_SS,
            ws.data_ptr(), wsb, st)
    G64 = A.double().t() @ B.double()
    magG = A.double().abs().t() @ B.double().abs()
    assert torch.all((Gl.double() - G64).abs() <= 2e-6 * magG)
    assert torch.all((G0.double() - Gl.double()).abs() <= 4e-6 * magG)


def test_tf32_operand_truncation_probe():
    """Records how the tensor core reads an fp32 container as tf32 (printed, not asserted): single-pass product with
    raw operands vs operands with the low 13 mantissa bits cleared."""
    from surfacenetworks_b200 import _native as Nt
    g = torch.Generator(device=DEV).manual_seed(11)
    M, N, K = 1024, 128, 128
    A = torch.randn(M, K, device=DEV, generator=g)
    B = torch.randn(N, K, device=DEV, generator=g)
    trunc = lambda t: (t.view(torch.int32) & ~0x1fff).view(torch.float32)
    outs = []
    for a, b in ((A, B), (trunc(A), trunc(B))):
        C = torch.empty(M, N, device=DEV)
        Nt.call("sn_gemm_tf32_f32", a.data_ptr(), K, b.data_ptr(), K, 0, 0, 0, 0, 0, 0, C.data_ptr(), N, M, N, K,
                Nt.SN_GEMM_SINGLE_PASS | Nt.SN_GEMM_LEGACY_SS, 0, 0, torch.cuda.current_stream().cuda_stream)
        outs.append(C)
    print("tf32 operand read == truncation:", bool(torch.equal(outs[0], outs[1])),
          "max diff", float((outs[0] - outs[1]).abs().max()))


@pytest.mark.parametrize("M,N,K", [(1000, 256, 128), (777, 64, 32), (4096, 128, 64)])
def test_gemm_elu_bwd_left_epilogue(M, N, K):
    """SN_GEMM_ELU_BWD_LEFT: columns [0, N/2) of dZ = dY Ws + p.*Z + q leave the epilogue multiplied by elu'(Z) (Z holds
    activated values) -- identical bits to the unfused product followed by the elementwise multiply."""
    from surfacenetworks_b200 import fused
    g = torch.Generator(device=DEV).manual_seed(M + N)
    dY = torch.randn(M, K, device=DEV, generator=g)
    Ws = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    Z = torch.randn(M, N, device=DEV, generator=g)
    Z[:, :N // 2] = torch.nn.functional.elu(Z[:, :N // 2])                 # activated left half, as in the stage buffer
    p, q = torch.randn(N, device=DEV, generator=g), torch.randn(N, device=DEV, generator=g)
    plain = fused.gemm_tf32(dY, Ws, bias=q, R=Z, rscale=p)
    fusedz = fused.gemm_tf32(dY, Ws, bias=q, R=Z, rscale=p, elu_bwd_left=True)
    a = Z[:, :N // 2]
    expect = plain.clone()
    expect[:, :N // 2] = plain[:, :N // 2] * torch.where(a > 0, torch.ones_like(a), a + 1)
    assert torch.equal(fusedz, expect)
    with pytest.raises(ValueError):
        fused.gemm_tf32(dY, Ws, bias=q, elu_bwd_left=True)                 # needs the residual operand


@pytest.mark.parametrize("Na,Nb,K", [(7000, 7000, 120), (300, 1000, 120), (129, 132, 32), (5000, 2052, 128), (64, 260, 64), (129, 4, 8)])
def test_wide_correlation_gemm_matches_fp64(Na, Nb, K):
    """dense_correspondence correlation FA . FB^T (models.py:199-203) on sn_gemm_nt_wide_tf32_f32: K and N tails are handled
    by the tensor maps (no padding), every output element within 2e-6 |A||B|^T of the fp64 product; gradients flow."""
    from surfacenetworks_b200 import fused
    g = torch.Generator(device=DEV).manual_seed(Na + Nb + K)
    FA = torch.randn(2, Na, K, device=DEV, generator=g).requires_grad_(True)
    FB = torch.randn(2, Nb, K, device=DEV, generator=g).requires_grad_(True)
    guard = torch.full((2, Na, Nb + 4), 7.0, device=DEV)
    out = fused.correlation(FA, FB)
    assert out.shape == (2, Na, Nb)
    ref = torch.bmm(FA.detach().double(), FB.detach().double().transpose(1, 2))
    mag = torch.bmm(FA.detach().double().abs(), FB.detach().double().abs().transpose(1, 2))
    err = (out.detach().double() - ref).abs()
    assert torch.all(err <= 2e-6 * mag + 1e-30), "max err/mag %g" % float((err / (mag + 1e-300)).max())
    w = torch.randn(2, Na, Nb, device=DEV, generator=g)
    (out * w).sum().backward()
    assert torch.allclose(FA.grad, torch.bmm(w, FB.detach()), rtol=1e-4, atol=1e-3)
    assert torch.allclose(FB.grad, torch.bmm(w.transpose(1, 2), FA.detach()), rtol=1e-4, atol=1e-3)
    assert torch.equal(out, fused.correlation(FA.detach(), FB.detach())), "bit-reproducible"
    del guard
