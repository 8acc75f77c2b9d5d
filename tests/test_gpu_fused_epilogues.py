"""Round-2 producer-side fusions, each against the separate passes it replaces (all through the C ABI):

* sn_gemm_tf32_presplit_act_f32 -- the dense stage that also emits elu(result) (the next stage's gather operand / left
  half) and its BatchNorm statistics: raw output bit-equal to sn_gemm_tf32_presplit_f32, activated output bit-equal to
  sn_elu_f32 of that, statistics within fp32 summation noise of an fp64 reduction, bit-stable run to run.
* sn_{bsr4,csr}_spmm_stats_f32 -- the SpMM that also emits the column statistics of its output: Y bit-equal to
  sn_*_spmm_f32, statistics vs fp64, bit-stable.

Reference lines: F.elu + torch.cat + BatchNorm1d statistics of utils_pt.py:161-168, 195-204, 208-216 and :98.
Tolerance for the statistics: |mean - mean64| <= 1e-5 (|mean64| + std64), |var - var64| <= 2e-5 (var64 + mean64^2)
(un-shifted fp32 partial sums over <= ~2000 rows per accumulator, fp64 across accumulators).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def check_stats(mean, var, Y, what):
    Y64 = Y.double()
    m64 = Y64.mean(0)
    v64 = Y64.var(0, unbiased=False)
    em = (mean.double() - m64).abs()
    ev = (var.double() - v64).abs()
    assert torch.all(em <= 1e-5 * (m64.abs() + v64.sqrt()) + 1e-30), "%s mean: worst %g" % (what, float(em.max()))
    assert torch.all(ev <= 2e-5 * (v64 + m64 * m64) + 1e-30), "%s var: worst rel %g" % (what, float((ev / (v64 + m64 * m64 + 1e-300)).max()))


@pytest.mark.parametrize("M,N,K", [(1000, 128, 256), (128 * 148 * 2 + 77, 128, 256), (5000, 256, 128), (700, 64, 128),
                                   (20000, 128, 128)])
@pytest.mark.parametrize("mode", ["act_only", "raw_and_act", "residual", "strided_no_stats"])
def test_gemm_act_epilogue(M, N, K, mode):
    from surfacenetworks_b200 import _native as Nv, fused
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    R = torch.randn(M, N, device=DEV, generator=g) * 2 if mode == "residual" else None
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    Nv.call("sn_split_tf32_f32", W.data_ptr(), K, N, K, hi.data_ptr(), lo.data_ptr(), torch.cuda.current_stream().cuda_stream)
    raw_ref = fused.gemm_tf32(A, hi, bias=bias, R=R, B_lo=lo)
    act_ref = torch.empty_like(raw_ref)
    Nv.call("sn_elu_f32", raw_ref.data_ptr(), N, act_ref.data_ptr(), N, M, N, torch.cuda.current_stream().cuda_stream)
    if mode == "strided_no_stats":
        Z = torch.full((M, 2 * N), 7.0, device=DEV)
        raw, act = fused.gemm_tf32_act(A, hi, lo, bias=bias, act_out=Z[:, :N], want_raw=False)
        assert raw is None
        assert torch.equal(act, act_ref)
        assert torch.all(Z[:, N:] == 7.0)
        return
    mean, var = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    raw, act = fused.gemm_tf32_act(A, hi, lo, bias=bias, R=R, mean=mean, var=var, want_raw=(mode != "act_only"))
    if mode == "act_only":
        assert raw is None
    else:
        assert torch.equal(raw, raw_ref)
    assert torch.equal(act, act_ref)
    check_stats(mean, var, act_ref, "gemm act")
    mean2, var2 = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    fused.gemm_tf32_act(A, hi, lo, bias=bias, R=R, mean=mean2, var=var2, want_raw=False)
    assert torch.equal(mean, mean2) and torch.equal(var, var2), "statistics are not bit-stable"


def test_gemm_act_group_bias_and_errors():
    from surfacenetworks_b200 import _native as Nv, fused
    M, N, K, rps = 3 * 700, 128, 128, 700
    g = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    gb = torch.randn(3, N, device=DEV, generator=g)
    R = torch.randn(M, N, device=DEV, generator=g)
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    st = torch.cuda.current_stream().cuda_stream
    Nv.call("sn_split_tf32_f32", W.data_ptr(), K, N, K, hi.data_ptr(), lo.data_ptr(), st)
    raw_ref = fused.gemm_tf32(A, hi, R=R, group_bias=gb, rows_per_group=rps, B_lo=lo)
    mean, var = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    raw, act = fused.gemm_tf32_act(A, hi, lo, R=R, group_bias=gb, rows_per_group=rps, mean=mean, var=var)
    assert torch.equal(raw, raw_ref)
    act_ref = torch.empty_like(raw_ref)
    Nv.call("sn_elu_f32", raw_ref.data_ptr(), N, act_ref.data_ptr(), N, M, N, st)
    assert torch.equal(act, act_ref)
    check_stats(mean, var, act, "gemm act group bias")
    # argument errors: no activated destination, mean without var, flags
    with pytest.raises(Nv.SurfnetError):
        Nv.call("sn_gemm_tf32_presplit_act_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, 0, 0, 0, 0, 0, 0, raw.data_ptr(), N,
                0, N, 0, 0, M, N, K, 0, 0, 0, st)
    with pytest.raises(Nv.SurfnetError):
        Nv.call("sn_gemm_tf32_presplit_act_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, 0, 0, 0, 0, 0, 0, raw.data_ptr(), N,
                act.data_ptr(), N, mean.data_ptr(), 0, M, N, K, 0, 0, 0, st)
    with pytest.raises(Nv.SurfnetError):
        Nv.call("sn_gemm_tf32_presplit_act_f32", A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, 0, 0, 0, 0, 0, 0, raw.data_ptr(), N,
                act.data_ptr(), N, 0, 0, M, N, K, Nv.SN_GEMM_SINGLE_PASS, 0, 0, st)


def _dirac_ops(num_vertices, B):
    from surfacenetworks_b200 import operators as OP, workloads as W
    meshes = W.make_mesh_ops(num_vertices, range(B))
    b = W.arap_batch(meshes, 0)
    lb = W.lap_batch(meshes)
    return (OP.Bsr4Operator.from_torch_coo(b["Di"].to(DEV)), OP.Bsr4Operator.from_torch_coo(b["DiA"].to(DEV)),
            OP.CsrOperator.from_torch_coo(lb["L"].to(DEV)))


@pytest.mark.parametrize("C", [32, 64, 128, 256, 512])
@pytest.mark.parametrize("shape", [(300, 3), (2000, 8)])
def test_spmm_stats_store_path(C, shape):
    D, DA, L = _dirac_ops(*shape)
    g = torch.Generator(device=DEV).manual_seed(C)
    for name, op, n_in in (("D", D, D.n_bcols), ("D*", DA, DA.n_bcols), ("L", L, L.n_cols)):
        X = torch.randn(n_in, C, device=DEV, generator=g)
        Y_ref = op.apply(X)
        Z = torch.full((Y_ref.shape[0], 2 * C), -3.0, device=DEV)           # strided destination: right half of a concat buffer
        mean, var = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        out = op.apply_stats(X, Z[:, C:], mean, var)
        assert out is not None, "row-group statistics path unsupported at C=%d" % C
        assert torch.equal(Z[:, C:], Y_ref), name
        assert torch.all(Z[:, :C] == -3.0)
        check_stats(mean, var, Y_ref, "%s C=%d" % (name, C))
        mean2, var2 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        op.apply_stats(X, Z[:, C:], mean2, var2)
        assert torch.equal(mean, mean2) and torch.equal(var, var2), "statistics are not bit-stable"


def test_spmm_stats_large_mean_and_unsupported_width():
    """A column with |mean| >> std (the un-shifted sums' worst case) stays inside the stated bound; C = 16 reports
    unsupported (the caller then runs the two passes)."""
    D, DA, L = _dirac_ops(500, 4)
    C = 128
    X = torch.randn(L.n_cols, C, device=DEV) * 0.01 + 5.0
    Lp = L                                       # L @ const ~ 0: use an operator with non-zero row sums instead
    X2 = torch.randn(DA.n_bcols, C, device=DEV)
    Y = DA.apply(X2)
    mean, var = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    DA.apply_stats(X2, torch.empty_like(Y), mean, var)
    check_stats(mean, var, Y, "D*")
    Y = Lp.apply(X)
    Lp.apply_stats(X, torch.empty_like(Y), mean, var)
    check_stats(mean, var, Y, "L shifted input")
    X16 = torch.randn(D.n_bcols, 16, device=DEV)
    assert D.apply_stats(X16, torch.empty(D.n_brows, 16, device=DEV), torch.empty(16, device=DEV), torch.empty(16, device=DEV)) is None


@pytest.mark.parametrize("C", [64, 128])
def test_chained_dirac_blocks_match_the_block_by_block_path(C):
    """Three DirResNet2 blocks: ``forward`` three times (raw f handed over, reference utils_pt.py:191-220) against
    ``forward_chained`` (activated f handed over inside the next block's concat buffer).  Same arithmetic up to the
    statistics' summation order: outputs, input / parameter gradients and BatchNorm buffers within 2e-4 of scale."""
    import copy
    from det import det_fill, det_tensor
    from surfacenetworks_b200 import ops, utils_pt as U
    D, DA, _ = _dirac_ops(400, 3)
    B, nv, nf = 3, D.n_bcols // 3, D.n_brows // 3
    blocks_a = [det_fill(U.DirResNet2(C), 10 + i).to(DEV).train() for i in range(3)]
    blocks_b = copy.deepcopy(blocks_a)
    w = det_tensor((B, nv, C), 77).to(DEV)

    def leaves():
        return det_tensor((B, nv, C), 5).to(DEV).requires_grad_(True), det_tensor((B, nf, C), 6).to(DEV).requires_grad_(True)

    va, fa = leaves()
    v, f = va, fa
    for blk in blocks_a:
        v, f = blk(D, DA, v, f)
    out_a = v
    (out_a * w).sum().backward()
    vb, fb = leaves()
    assert blocks_b[0].chain_supported(vb, fb)
    state = ops.face_chain_start(fb.reshape(-1, C))
    v = vb
    for i, blk in enumerate(blocks_b):
        v, state = blk.forward_chained(D, DA, v, state, last=(i == 2))
    out_b = v
    (out_b * w).sum().backward()

    def close(a, b, what):
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        assert err <= 2e-4 * max(scale, 1e-6), "%s: max err %g at scale %g" % (what, err, scale)

    close(out_b.detach(), out_a.detach(), "v_out")
    close(vb.grad, va.grad, "grad v")
    close(fb.grad, fa.grad, "grad f")
    for ba, bb in zip(blocks_a, blocks_b):
        gs = max(float(p.grad.abs().max()) for p in ba.parameters())
        for (k, pa), (_, pb) in zip(ba.named_parameters(), bb.named_parameters()):
            err = float((pa.grad - pb.grad).abs().max())
            assert err <= 2e-4 * max(float(pa.grad.abs().max()), 1e-2 * gs), "grad %s: %g" % (k, err)
        for (k, xa), (_, xb) in zip(ba.named_buffers(), bb.named_buffers()):
            close(xb.float(), xa.float(), "buffer " + k)


@pytest.mark.parametrize("C,B,nv", [(128, 3, 150), (256, 2, 90), (128, 70, 40)])
def test_avg_block_fused_stage_vs_oracle(C, B, nv):
    """AvgResNet2 (utils_pt.py:222-243) on the fused stage kernels (sn_avg_stage_pre_f32 / sn_avg_fold_{fwd,bwd}_f32) against
    the oracle port on the CPU: forward, input and parameter gradients, BatchNorm buffers.  B = 70 exercises the 64-mesh
    chunking of the fold kernels, ragged masks exercise the masked average.  Tolerance 2e-4 * (|ref| + max|ref|)."""
    from det import det_fill, det_tensor
    from oracle import layers as O
    from surfacenetworks_b200 import utils_pt as U
    block = det_fill(U.AvgResNet2(C), 5, gain=0.5)
    P = {}
    for k, v in block.state_dict().items():
        v = v.clone()
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
        P[k] = v
    x = det_tensor((B, nv, C), 1)
    mask = torch.ones(B, nv, 1)
    for b in range(B):
        mask[b, nv - (b % 7):] = 0           # ragged meshes
    x = x * mask
    w = det_tensor((B, nv, C), 3)
    xc = x.clone().requires_grad_(True)
    ref = O.avg_resnet2(P, mask, xc)
    (ref * w).sum().backward()
    blk = block.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    out = blk(None, mask.to(DEV), xg)
    (out * w.to(DEV)).sum().backward()

    def close(a, b, what, tol=2e-4, floor=0.0):
        a, b = a.detach().cpu().double(), b.detach().double()
        err = (a - b).abs()
        assert torch.all(err <= tol * (b.abs() + max(float(b.abs().max()), floor))), "%s: max err %g scale %g" % (what, float(err.max()), float(b.abs().max()))

    close(out, ref, "out")
    close(xg.grad, xc.grad, "dx")
    gscale = max(float(P[k].grad.abs().max()) for k, _ in blk.named_parameters())
    for k, p in blk.named_parameters():
        close(p.grad, P[k].grad, "grad " + k, 1e-3, floor=1e-2 * gscale)


@pytest.mark.parametrize("with_r2", [False, True])
@pytest.mark.parametrize("rows_per_seg,n_seg,C,K", [(700, 3, 128, 128), (2000, 8, 128, 128), (333, 5, 256, 256)])
def test_gemm_elu_backward_epilogue(rows_per_seg, n_seg, C, K, with_r2):
    """sn_gemm_tf32_presplit_elubwd_f32 (AvgResNet2 stage backward, utils_pt.py:231-243 through autograd) against the two
    launches it replaces: sn_gemm_tf32_presplit_f32 (dZ = dY W + q + p .* a) then sn_elu_bwd_group_f32
    ((dZ + mask * gb[mesh]) .* elu'(a) + g3).  Same arithmetic per element up to one fused multiply-add: 4 ulps of the
    summed magnitudes."""
    from surfacenetworks_b200 import _native as Nv, fused
    M = rows_per_seg * n_seg
    g = torch.Generator(device=DEV).manual_seed(M + C + K)
    dY = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(C, K, device=DEV, generator=g) / K ** 0.5
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    st = torch.cuda.current_stream().cuda_stream
    Nv.call("sn_split_tf32_f32", W.data_ptr(), K, C, K, hi.data_ptr(), lo.data_ptr(), st)
    a = torch.nn.functional.elu(torch.randn(M, C, device=DEV, generator=g))
    p, q = torch.randn(C, device=DEV, generator=g) * 0.1, torch.randn(C, device=DEV, generator=g)
    gb = torch.randn(n_seg, C, device=DEV, generator=g)
    maskw = (torch.rand(M, device=DEV, generator=g) > 0.1).float()
    g3 = torch.randn(M, C, device=DEV, generator=g) if with_r2 else None
    dZ = fused.gemm_tf32(dY, hi, bias=q, R=a, rscale=p, B_lo=lo)
    ref = torch.empty_like(a)
    Nv.call("sn_elu_bwd_group_f32", a.data_ptr(), C, dZ.data_ptr(), C, gb.data_ptr(), maskw.data_ptr(), rows_per_seg,
            0 if g3 is None else g3.data_ptr(), C, ref.data_ptr(), C, M, C, st)
    out = torch.empty_like(a)
    Nv.call("sn_gemm_tf32_presplit_elubwd_f32", dY.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, q.data_ptr(), a.data_ptr(), C,
            p.data_ptr(), gb.data_ptr(), rows_per_seg, maskw.data_ptr(), 0 if g3 is None else g3.data_ptr(), C, out.data_ptr(), C,
            M, C, K, st)
    # magnitude of the summed terms (the two paths add them in different orders)
    mag = ((dZ - q - p * a).abs() + q.abs() + (p * a).abs() + gb.abs()[torch.arange(M, device=DEV) // rows_per_seg]
           + (0 if g3 is None else g3.abs()) + 1e-6)
    assert torch.all((out - ref).abs() <= 5e-7 * mag), float(((out - ref).abs() / mag).max())
