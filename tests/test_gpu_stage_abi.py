"""a whole DirResNet2 / LapResNet2 block driven through the four single-call stage entry points ONLY
(sn_dir_stage_fwd/_bwd_f32, sn_lap_stage_fwd/_bwd_f32) -- what a non-Python host would call -- against the package's own
module path (itself pinned to the reference goldens and the oracle) and, for the forward, against the oracle port on the CPU.
Reference: src/utils/utils_pt.py:151-180 (LapResNet2), :182-220 (DirResNet2).  Tolerance 2e-4 * (|ref| + max|ref|)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _close(a, b, what, tol=2e-4):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs()
    assert torch.all(err <= tol * (b.abs() + float(b.abs().max()))), "%s: max err %g scale %g" % (what, float(err.max()), float(b.abs().max()))


class _Stage:
    """Caller-side buffers of one stage (what a C host would allocate)."""

    def __init__(self, rows_out, rows_in, C, dirac):
        z = lambda *s: torch.empty(*s, device=DEV)
        self.Z, self.stk, self.mean, self.var = z(rows_out, 2 * C), z(3, 2 * C), z(2 * C), z(2 * C)
        self.act = z(rows_in, C) if dirac else None
        self.Y = z(rows_out, C)
        self.dZ = z(rows_out, 2 * C)
        self.dg, self.dgamma, self.dbeta, self.dW, self.db = z(rows_in, C), z(2 * C), z(2 * C), z(C, 2 * C), z(C)


def _ws(nbytes):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=DEV)


@pytest.mark.parametrize("C", [128])
def test_dirac_block_through_stage_entry_points(C):
    from det import det_fill, det_tensor
    from surfacenetworks_b200 import _native as N, operators as OP, utils_pt as U, workloads as W
    meshes = W.make_mesh_ops(260, [0, 1]) + W.make_mesh_ops(240, [2])
    host = W.arap_batch(meshes, 0)
    B, nv, nf = 3, host["num_vertices"], host["num_faces"]
    D, DA = OP.Bsr4Operator.from_torch_coo(host["Di"].to(DEV)), OP.Bsr4Operator.from_torch_coo(host["DiA"].to(DEV))
    blk = det_fill(U.DirResNet2(C), 7, gain=0.5).to(DEV).train()
    ref = det_fill(U.DirResNet2(C), 7, gain=0.5).to(DEV).train()
    v = det_tensor((B * nv, C), 1).to(DEV)
    f = det_tensor((B * nf, C), 2).to(DEV)
    gv, gf = det_tensor((B * nv, C), 3).to(DEV), det_tensor((B * nf, C), 4).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    fw = _ws(N.lib.sn_stage_fwd_ws_bytes(C))
    s0, s1 = _Stage(B * nf, B * nv, C, True), _Stage(B * nv, B * nf, C, True)
    c0, c1 = blk.bn_fc0, blk.bn_fc1

    def fwd(op, s, x_self, x_gather, conv, residual):
        N.call("sn_dir_stage_fwd_f32", _ptr(op.browptr), _ptr(op.bcolind), _ptr(op.bval), op.n_brows, op.n_bcols, _ptr(x_self), C,
               _ptr(x_gather), C, C, _ptr(conv.bn.weight), _ptr(conv.bn.bias), _ptr(conv.fc.weight), _ptr(conv.fc.bias),
               _ptr(residual), C, _ptr(conv.bn.running_mean), _ptr(conv.bn.running_var), 0.1, conv.bn.eps, _ptr(s.Z), _ptr(s.act),
               _ptr(s.stk), _ptr(s.mean), _ptr(s.var), _ptr(s.Y), C, _ptr(fw), fw.numel(), st)

    fwd(D, s0, f, v, c0, None)                     # f_out = fc0(BN[elu(f) | D elu(v)])
    fwd(DA, s1, v, s0.Y, c1, v)                    # v_new = v + fc1(BN[elu(v) | D* elu(f_out)])

    def bwd(opT, s, rows_out, rows_in, dY, conv, g_extra):
        bw = _ws(N.lib.sn_stage_bwd_ws_bytes(rows_out, C))
        N.call("sn_dir_stage_bwd_f32", _ptr(opT.browptr), _ptr(opT.bcolind), _ptr(opT.bval), rows_out, rows_in, _ptr(dY), C,
               _ptr(s.Z), _ptr(s.act), _ptr(conv.fc.weight), _ptr(s.stk), _ptr(s.mean), C, _ptr(s.dZ), _ptr(s.dg), C,
               _ptr(g_extra), C, _ptr(s.dgamma), _ptr(s.dbeta), _ptr(s.dW), _ptr(s.db), _ptr(bw), bw.numel(), st)

    bwd(DA.T, s1, B * nv, B * nf, gv, c1, gf)      # d f_out (+ the downstream gradient of f_out); dZ_left = d v through elu(v)
    bwd(D.T, s0, B * nf, B * nv, s1.dg, c0, None)  # d v through the gather; dZ_left = d f
    g_v = s1.dZ[:, :C] + s0.dg + gv
    g_f = s0.dZ[:, :C]

    vr, fr = v.clone().requires_grad_(True), f.clone().requires_grad_(True)
    vo, fo = ref(D, DA, vr.view(B, nv, C), fr.view(B, nf, C))
    ((vo.reshape(-1, C) * gv).sum() + (fo.reshape(-1, C) * gf).sum()).backward()
    _close(s1.Y, vo.reshape(-1, C), "v_new")
    _close(s0.Y, fo.reshape(-1, C), "f_out")
    _close(g_v, vr.grad, "grad v")
    _close(g_f, fr.grad, "grad f")
    for s, conv in ((s0, ref.bn_fc0), (s1, ref.bn_fc1)):
        gs = float(conv.fc.weight.grad.abs().max())
        for got, want, name in ((s.dW, conv.fc.weight.grad, "dW"), (s.db, conv.fc.bias.grad, "db"),
                                (s.dgamma, conv.bn.weight.grad, "dgamma"), (s.dbeta, conv.bn.bias.grad, "dbeta")):
            err = float((got - want).abs().max())
            assert err <= 1e-3 * max(float(want.abs().max()), 1e-2 * gs), "%s: %g" % (name, err)
    for a, b in ((blk.bn_fc0.bn, ref.bn_fc0.bn), (blk.bn_fc1.bn, ref.bn_fc1.bn)):
        _close(a.running_mean, b.running_mean, "running_mean")
        _close(a.running_var, b.running_var, "running_var")
    # argument errors: short workspace, unsupported width
    with pytest.raises(N.SurfnetError):
        N.call("sn_dir_stage_fwd_f32", _ptr(D.browptr), _ptr(D.bcolind), _ptr(D.bval), D.n_brows, D.n_bcols, _ptr(f), C, _ptr(v), C,
               C, _ptr(c0.bn.weight), _ptr(c0.bn.bias), _ptr(c0.fc.weight), _ptr(c0.fc.bias), 0, C, 0, 0, 0.1, 1e-5, _ptr(s0.Z),
               _ptr(s0.act), _ptr(s0.stk), _ptr(s0.mean), _ptr(s0.var), _ptr(s0.Y), C, _ptr(fw), 16, st)
    assert N.lib.sn_dir_stage_fwd_f32(_ptr(D.browptr), _ptr(D.bcolind), _ptr(D.bval), D.n_brows, D.n_bcols, _ptr(f), 64, _ptr(v), 64,
                                      64, _ptr(c0.bn.weight), _ptr(c0.bn.bias), _ptr(c0.fc.weight), _ptr(c0.fc.bias), 0, 64, 0, 0, 0.1,
                                      1e-5, _ptr(s0.Z), _ptr(s0.act), _ptr(s0.stk), _ptr(s0.mean), _ptr(s0.var), _ptr(s0.Y), 64,
                                      _ptr(fw), fw.numel(), st) == N.SN_ERR_UNSUPPORTED


def test_laplacian_block_through_stage_entry_points():
    from det import det_fill, det_tensor
    from oracle import layers as O
    from surfacenetworks_b200 import _native as N, operators as OP, utils_pt as U, workloads as W
    C = 128
    meshes = W.make_mesh_ops(300, [0, 1]) + W.make_mesh_ops(280, [2])
    lb = W.lap_batch(meshes)
    B, nv = 3, lb["num_vertices"]
    L = OP.CsrOperator.from_torch_coo(lb["L"].to(DEV))
    blk = det_fill(U.LapResNet2(C), 9, gain=0.5).to(DEV).train()
    ref = det_fill(U.LapResNet2(C), 9, gain=0.5).to(DEV).train()
    x = det_tensor((B * nv, C), 1).to(DEV)
    gy = det_tensor((B * nv, C), 3).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    fw = _ws(N.lib.sn_stage_fwd_ws_bytes(C))
    s0, s1 = _Stage(B * nv, B * nv, C, False), _Stage(B * nv, B * nv, C, False)

    def fwd(s, xin, conv, residual):
        N.call("sn_lap_stage_fwd_f32", _ptr(L.rowptr), _ptr(L.colind), _ptr(L.val), L.n_rows, _ptr(xin), C, C, _ptr(conv.bn.weight),
               _ptr(conv.bn.bias), _ptr(conv.fc.weight), _ptr(conv.fc.bias), _ptr(residual), C, _ptr(conv.bn.running_mean),
               _ptr(conv.bn.running_var), 0.1, conv.bn.eps, _ptr(s.Z), _ptr(s.stk), _ptr(s.mean), _ptr(s.var), _ptr(s.Y), C,
               _ptr(fw), fw.numel(), st)

    fwd(s0, x, blk.bn_fc0, None)
    fwd(s1, s0.Y, blk.bn_fc1, x)
    LT = L.T
    bw = _ws(N.lib.sn_stage_bwd_ws_bytes(B * nv, C))

    def bwd(s, dY, conv, g_extra):
        N.call("sn_lap_stage_bwd_f32", _ptr(LT.rowptr), _ptr(LT.colind), _ptr(LT.val), LT.n_rows, _ptr(dY), C, _ptr(s.Z),
               _ptr(conv.fc.weight), _ptr(s.stk), _ptr(s.mean), C, _ptr(s.dZ), _ptr(s.dg), C, _ptr(g_extra), C, _ptr(s.dgamma),
               _ptr(s.dbeta), _ptr(s.dW), _ptr(s.db), _ptr(bw), bw.numel(), st)

    bwd(s1, gy, blk.bn_fc1, None)                  # gradient of the first stage's output
    bwd(s0, s1.dg, blk.bn_fc0, gy)                 # gradient of x: through stage 0, plus the residual path
    xr = x.clone().requires_grad_(True)
    out = ref(L, None, xr.view(B, nv, C))
    (out.reshape(-1, C) * gy).sum().backward()
    _close(s1.Y, out.reshape(-1, C), "out")
    _close(s0.dg, xr.grad, "grad x")
    for s, conv in ((s0, ref.bn_fc0), (s1, ref.bn_fc1)):
        gs = float(conv.fc.weight.grad.abs().max())
        for got, want, name in ((s.dW, conv.fc.weight.grad, "dW"), (s.db, conv.fc.bias.grad, "db"),
                                (s.dgamma, conv.bn.weight.grad, "dgamma"), (s.dbeta, conv.bn.bias.grad, "dbeta")):
            err = float((got - want).abs().max())
            assert err <= 1e-3 * max(float(want.abs().max()), 1e-2 * gs), "%s: %g" % (name, err)
    # forward against the oracle port (CPU, fp32)
    P = {k: v_.detach().cpu().clone() for k, v_ in det_fill(U.LapResNet2(C), 9, gain=0.5).state_dict().items()}
    o = O.lap_resnet2(P, lb["L"], x.cpu().view(B, nv, C))
    _close(s1.Y.cpu(), o.reshape(-1, C), "out vs oracle")
