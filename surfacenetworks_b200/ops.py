"""Autograd seam of the operator-application path.

Mirrors the reference's ``SparseBMMFunc`` (src/utils/cuda/sparse_bmm_func.py:23-72): forward applies the
sparse operator to a dense matrix, backward applies the TRANSPOSED operator to the incoming gradient and
returns no gradient for the operator itself (:60,72).  Differences: static new-style Functions, the CSR /
BSR structures (and their transposes) are built once per operator instead of on every call (:39,66-67),
and the ELU in front of the product and the ``torch.cat`` behind it (utils_pt.py:161-168 etc.) are folded
into the same Function so the concat buffer is written in place.
"""
from __future__ import annotations

import torch

from . import _native as N
from . import fused
from .operators import Bsr4Operator, CsrOperator, _check_dense, _ptr, _stream

__all__ = ["spmm", "stage_concat", "elu_into"]


def _as2d(x):
    """[B, n, C] contiguous -> [B*n, C] view (what ``x.view(-1, feat)`` does at utils_pt.py:167)."""
    if x.dim() == 3:
        return x.reshape(-1, x.shape[-1])
    return x


def elu_into(X, out):
    """out[:, :] = elu(X) through sn_elu_f32 (strided destination, e.g. the left half of a concat buffer)."""
    _check_dense(X, "X")
    _check_dense(out, "out")
    with torch.cuda.device(X.device):
        N.call("sn_elu_f32", _ptr(X), X.stride(0), _ptr(out), out.stride(0), X.shape[0], X.shape[1], _stream())
    return out


def _elu_bwd(A, a_is_raw, G, G2, out):
    with torch.cuda.device(G.device):
        N.call("sn_elu_bwd_f32", _ptr(A), A.stride(0), 1 if a_is_raw else 0, _ptr(G), G.stride(0),
               _ptr(G2), 0 if G2 is None else G2.stride(0), _ptr(out), out.stride(0), G.shape[0], G.shape[1], _stream())
    return out


class _Spmm(torch.autograd.Function):
    """Y = S @ X  (grad only w.r.t. X, as sparse_bmm_func.py:53-72)."""

    @staticmethod
    def forward(ctx, X, op):
        ctx.op = op
        return op.apply(X.contiguous())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gY):
        return ctx.op.T.apply(gY.contiguous()), None


def spmm(op, X):
    """Differentiable ``op @ X`` for a CsrOperator / Bsr4Operator and a 2-D dense X."""
    if not isinstance(op, (CsrOperator, Bsr4Operator)):
        raise TypeError("op must be a CsrOperator or Bsr4Operator")
    return _Spmm.apply(X, op)


class _StageConcat(torch.autograd.Function):
    """Z[rows_out, 2C] = [ elu(x_self) | S @ elu(x_gather) ]  -- one stage's operator application.

    Covers utils_pt.py:161-168 / 172-177 (Laplacian, x_self is x_gather) and :195-204 / :208-216 (Dirac D with
    x_self = f, x_gather = v; adjoint D* with x_self = v, x_gather = f_out).  The activation is materialised once
    per operand (sn_elu_f32) instead of being recomputed for every gathered copy of a row: each row is gathered
    ~3-7 times, and expm1 on the gather path costs more than the extra pass (measured: +60..100 us per SpMM at
    the ARAP size).  For the Laplacian the activated rows ARE the left half of Z and the SpMM gathers from there.
    """

    @staticmethod
    def forward(ctx, x_self, x_gather, op, same, want_stats, cell):
        xs = x_self.contiguous()
        rows_out, C = xs.shape
        Z = torch.empty(rows_out, 2 * C, dtype=torch.float32, device=xs.device)
        left, right = Z[:, :C], Z[:, C:]
        stats = None
        if want_stats and fused.elu_colstats_supported(xs, left):
            stats = fused.elu_colstats(xs, left)          # activation + BatchNorm statistics of the left half, one pass
        else:
            elu_into(xs, left)
        if same:
            op.apply(left, out=right)                     # gather from the activated left half (row stride 2C)
            ctx.save_for_backward(Z)
        else:
            xg = x_gather.contiguous()
            act = torch.empty_like(xg)
            elu_into(xg, act)
            op.apply(act, out=right)
            ctx.save_for_backward(Z, act)
        ctx.op, ctx.same, ctx.C, ctx.cell = op, same, C, cell
        if stats is None:
            empty = Z.new_empty(0)
            stats = (empty, empty.clone())
        ctx.mark_non_differentiable(stats[0], stats[1])
        return Z, stats[0], stats[1]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gZ, _gm=None, _gv=None):
        op, C = ctx.op, ctx.C
        gZ = gZ.contiguous()
        g_left, g_right = gZ[:, :C], gZ[:, C:]
        if ctx.same:
            (Z,) = ctx.saved_tensors
            g_res = ctx.cell.pop("residual_grad", None) if ctx.cell is not None else None
            # (g_left + S^T g) * elu'(x) + residual gradient, all in the store path
            t = op.T.apply_epilogue(g_right, G=g_left, A=Z[:, :C], G2=g_res)
            if t is None:
                t = op.T.apply(g_right)                   # S^T g
                _elu_bwd(Z[:, :C], False, g_left, t, t)   # (g_left + S^T g) * elu'(x), in place
                if g_res is not None:
                    t += g_res
            return t, None, None, None, None, None
        Z, act = ctx.saved_tensors
        if ctx.cell is not None and ctx.cell.get("left_premultiplied"):
            g_self = g_left                               # the dZ GEMM already applied elu'(x_self) (SN_GEMM_ELU_BWD_LEFT)
        else:
            g_self = torch.empty(gZ.shape[0], C, dtype=torch.float32, device=gZ.device)
            _elu_bwd(Z[:, :C], False, g_left, None, g_self)
        t = op.T.apply_epilogue(g_right, A=act)           # (S^T g) * elu'(x_gather) in the store path
        if t is None:
            t = op.T.apply(g_right)
            _elu_bwd(act, False, t, None, t)
        return g_self, t, None, None, None, None


def stage_concat(op, x_self, x_gather=None, want_stats=True, in_cell=None):
    """``[elu(x_self) | op @ elu(x_gather)]`` as one [rows, 2C] buffer; ``x_gather=None`` means x_self.

    With ``want_stats`` the activation pass also reduces the left half's BatchNorm statistics; they ride on the
    returned tensor (``Z._sn_left_stats``) and ``fused.bn_linear`` then only reduces the right half."""
    same = x_gather is None
    # not-same stages: the consumer (fused.bn_linear) may fold elu'(x_self) into its dZ GEMM epilogue; it says so
    # through this cell, which the backward above reads (default: not folded, run the elementwise pass)
    # in_cell (Laplacian blocks): the block's second stage leaves the residual's gradient there
    # (fused.bn_linear(res_cell=...)); it is added in this stage's backward SpMM epilogue instead of by autograd
    cell = in_cell if same else {"left_premultiplied": False}
    Z, mean_l, var_l = _StageConcat.apply(x_self, x_self if same else x_gather, op, same, want_stats, cell)
    if mean_l.numel():
        Z._sn_left_stats = (mean_l, var_l)
    if cell is not None and not same:
        Z._sn_stage_cell = cell
    return Z


class _DirBlock(torch.autograd.Function):
    """A whole DirResNet2 block (reference src/utils/utils_pt.py:191-220) as ONE autograd node:

        f_out = fc0(BN[elu(f) | D elu(v)]),   v_new = v + fc1(BN[elu(v) | D* elu(f_out)])

    Same kernels as the two-stage composition (stage_concat + fused.bn_linear); what the single node buys:
      * elu(v) is materialised once (left half of the vertex stage's buffer; D gathers from it in place);
      * backward: every gradient that autograd would accumulate with separate add kernels (v has three consumers,
        f_out two) rides in an epilogue instead --
            g_fout = (D*^T dZv_right) .* elu'(f_out) + g_f_downstream
            g_v    = (D^T dZf_right + dZv_left) .* elu'(v) + g_vnew           (sn_bsr4_spmm_epilogue_f32)
        and elu'(f) is applied by the dZ GEMM epilogue (SN_GEMM_ELU_BWD_LEFT).
    Used in training mode at the widths both epilogues cover; otherwise DirResNet2 composes the stage functions."""

    @staticmethod
    def forward(ctx, v2, f2, g0, b0, W0, c0, g1, b1, W1, c1, D, DA, bn0, bn1):
        C = v2.shape[1]
        v2, f2 = v2.contiguous(), f2.contiguous()
        Zv = torch.empty(v2.shape[0], 2 * C, dtype=torch.float32, device=v2.device)
        Zf = torch.empty(f2.shape[0], 2 * C, dtype=torch.float32, device=v2.device)
        st = torch.empty(4, 2 * C, dtype=torch.float32, device=v2.device)      # mean / var of Zv, mean / var of Zf
        fused.elu_colstats(v2, Zv[:, :C], st[0, :C], st[1, :C])
        fused.elu_colstats(f2, Zf[:, :C], st[2, :C], st[3, :C])
        stats_v, stats_f = (st[0], st[1], C), (st[2], st[3], C)
        D.apply(Zv[:, :C], out=Zf[:, C:])                       # faces <- vertices, gathers the activated rows in place
        cnt0, cnt1 = fused.bn_count_batch(bn0, True), fused.bn_count_batch(bn1, True)
        f_out, saved0 = fused.bn_linear_forward(Zf, g0, b0, W0, c0, None, bn0.running_mean, bn0.running_var, True,
                                                fused.bn_momentum(bn0), bn0.eps, stats_f, counter=cnt0)
        act_f = torch.empty_like(f_out)
        elu_into(f_out, act_f)
        DA.apply(act_f, out=Zv[:, C:])                          # vertices <- faces
        v_new, saved1 = fused.bn_linear_forward(Zv, g1, b1, W1, c1, v2, bn1.running_mean, bn1.running_var, True,
                                                fused.bn_momentum(bn1), bn1.eps, stats_v, counter=cnt1)
        ctx.save_for_backward(act_f, *saved0, *saved1)
        ctx.D, ctx.DA, ctx.C = D, DA, C
        ctx.set_materialize_grads(False)
        return v_new, f_out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_vnew, g_fdown):
        act_f, Zf, W0, stk0, mean0, Zv, W1, stk1, mean1 = ctx.saved_tensors
        C, D, DA = ctx.C, ctx.D, ctx.DA

        def dense(g):
            if g is not None and (g.stride(1) != 1 or g.stride(0) % 4 or g.data_ptr() % 16):
                g = g.contiguous()
            return g
        g_vnew, g_fdown = dense(g_vnew), dense(g_fdown)
        if g_vnew is None:
            g_vnew = torch.zeros(Zv.shape[0], C, dtype=torch.float32, device=Zv.device)
        dZv, dg1, db1, dW1, dc1 = fused.bn_linear_backward((Zv, W1, stk1, mean1), g_vnew, True, False)
        g_fout = _epilogue_or_passes(DA.T, dZv[:, C:], None, act_f, g_fdown)
        dZf, dg0, db0, dW0, dc0 = fused.bn_linear_backward((Zf, W0, stk0, mean0), g_fout, True, True)
        g_v = _epilogue_or_passes(D.T, dZf[:, C:], dZv[:, :C], Zv[:, :C], g_vnew)
        return g_v, dZf[:, :C], dg0, db0, dW0, dc0, dg1, db1, dW1, dc1, None, None, None, None


def _spmm_with_stats(op, X, out, mean, var):
    """out = op @ X and the column statistics of out: one launch when the row-group kernel covers the shape."""
    if op.apply_stats(X, out, mean, var) is None:
        op.apply(X, out=out)
        fused.colstats(out, mean, var)


class _FaceChainStart(torch.autograd.Function):
    """Entry of a chain of Dirac blocks: f [F, C] -> (Zf [F, 2C] whose LEFT half holds elu(f), st [2, 2C] whose first C
    columns hold that half's mean / biased variance).  The right halves are written by the first block of the chain."""

    @staticmethod
    def forward(ctx, f2):
        C = f2.shape[1]
        f2 = f2.contiguous()
        Zf = torch.empty(f2.shape[0], 2 * C, dtype=torch.float32, device=f2.device)
        st = torch.empty(2, 2 * C, dtype=torch.float32, device=f2.device)
        fused.elu_colstats(f2, Zf[:, :C], st[0, :C], st[1, :C])
        ctx.C = C
        ctx.mark_non_differentiable(st)
        return Zf, st

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gZf, _gst=None):
        # the consuming block's dZ GEMM has already applied elu'(f) to the left half (SN_GEMM_ELU_BWD_LEFT)
        return None if gZf is None else gZf[:, :ctx.C]


class _DirBlockChained(torch.autograd.Function):
    """DirResNet2 (reference src/utils/utils_pt.py:191-220) inside a stack whose face features only ever travel from one
    Dirac block to the next (as_rigid_as_possible/models.py:142-146, normal_predict/models.py:262-268,
    dense_correspondence/models.py:170-176).  Nobody reads the raw f_out: the next block takes elu(f_out) as the left half
    of its concat buffer and D* gathers elu(f_out).  So, compared with _DirBlock:

      * the face stage's GEMM writes ONLY elu(f_out), straight into the left half of the NEXT block's concat buffer, and
        reduces that half's BatchNorm statistics in its epilogue (sn_gemm_tf32_presplit_act_f32) -- the separate
        sn_elu_f32 pass and the next block's sn_elu_colstats_f32 pass over the faces are gone;
      * both SpMMs reduce the statistics of the right halves in their store path (sn_bsr4_spmm_stats_f32) -- the two
        sn_colstats_f32 passes are gone.
    Backward is _DirBlock's (the gradient w.r.t. raw f_out arrives as the left half of the next block's dZ, already
    multiplied by elu'(f_out) by that block's dZ GEMM)."""

    @staticmethod
    def forward(ctx, v2, Zf, stf, g0, b0, W0, c0, g1, b1, W1, c1, D, DA, bn0, bn1, last):
        C = v2.shape[1]
        v2 = v2.contiguous()
        dev = v2.device
        Zv = torch.empty(v2.shape[0], 2 * C, dtype=torch.float32, device=dev)
        stv = torch.empty(2, 2 * C, dtype=torch.float32, device=dev)
        fused.elu_colstats(v2, Zv[:, :C], stv[0, :C], stv[1, :C])
        _spmm_with_stats(D, Zv[:, :C], Zf[:, C:], stf[0, C:], stf[1, C:])        # faces <- vertices
        if last:
            Zf_next = st_next = None
            act_f = torch.empty(Zf.shape[0], C, dtype=torch.float32, device=dev)
            act = dict(act_out=act_f, want_raw=False)
        else:
            Zf_next = torch.empty_like(Zf)
            st_next = torch.empty_like(stf)
            act_f = Zf_next[:, :C]
            act = dict(act_out=act_f, mean=st_next[0, :C], var=st_next[1, :C], want_raw=False)
        cnt0, cnt1 = fused.bn_count_batch(bn0, True), fused.bn_count_batch(bn1, True)
        _, saved0 = fused.bn_linear_forward(Zf, g0, b0, W0, c0, None, bn0.running_mean, bn0.running_var, True,
                                            fused.bn_momentum(bn0), bn0.eps, (stf[0], stf[1], 2 * C),
                                            act=act, counter=cnt0)
        _spmm_with_stats(DA, act_f, Zv[:, C:], stv[0, C:], stv[1, C:])           # vertices <- faces
        v_new, saved1 = fused.bn_linear_forward(Zv, g1, b1, W1, c1, v2, bn1.running_mean, bn1.running_var, True,
                                                fused.bn_momentum(bn1), bn1.eps, (stv[0], stv[1], 2 * C), counter=cnt1)
        ctx.save_for_backward(act_f, *saved0, *saved1)
        ctx.D, ctx.DA, ctx.C = D, DA, C
        ctx.set_materialize_grads(False)
        if last:
            Zf_next = Zf.new_empty(0)
            st_next = Zf.new_empty(0)
        ctx.mark_non_differentiable(st_next)
        return v_new, Zf_next, st_next

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_vnew, g_Zfnext, _gst=None):
        act_f, Zf, W0, stk0, mean0, Zv, W1, stk1, mean1 = ctx.saved_tensors
        C, D, DA = ctx.C, ctx.D, ctx.DA
        g_fdown = None
        if g_Zfnext is not None and g_Zfnext.numel():
            g_fdown = g_Zfnext[:, :C]
            if g_fdown.stride(1) != 1 or g_fdown.stride(0) % 4 or g_fdown.data_ptr() % 16:
                g_fdown = g_fdown.contiguous()
        if g_vnew is not None and (g_vnew.stride(1) != 1 or g_vnew.stride(0) % 4 or g_vnew.data_ptr() % 16):
            g_vnew = g_vnew.contiguous()
        if g_vnew is None:
            g_vnew = torch.zeros(Zv.shape[0], C, dtype=torch.float32, device=Zv.device)
        dZv, dg1, db1, dW1, dc1 = fused.bn_linear_backward((Zv, W1, stk1, mean1), g_vnew, True, False)
        g_fout = _epilogue_or_passes(DA.T, dZv[:, C:], None, act_f, g_fdown)
        dZf, dg0, db0, dW0, dc0 = fused.bn_linear_backward((Zf, W0, stk0, mean0), g_fout, True, True)
        g_v = _epilogue_or_passes(D.T, dZf[:, C:], dZv[:, :C], Zv[:, :C], g_vnew)
        return g_v, dZf, None, dg0, db0, dW0, dc0, dg1, db1, dW1, dc1, None, None, None, None, None


def _epilogue_or_passes(opT, X, G, A, G2):
    """(opT @ X + G) .* elu'(A) + G2: one launch (row-group epilogue) or, where that kernel does not apply, the passes."""
    t = opT.apply_epilogue(X, G=G, A=A, G2=G2)
    if t is None:
        t = opT.apply(X)
        _elu_bwd(A, False, t, G, t)               # (opT @ X + G) .* elu'(A), in place
        if G2 is not None:
            t += G2
    return t


_DIR_BLOCK_WIDTHS = (64, 128)          # C and 2C both tensor-core GEMM widths, C covered by the row-group epilogue


def dir_block_supported(v2, f2, conv0, conv1):
    """True when the single-node Dirac block applies: CUDA fp32 rows, training-mode BatchNorm on both stages, widths
    covered by the epilogues, gradients enabled."""
    C = v2.shape[1]
    if not (torch.is_grad_enabled() and v2.is_cuda and v2.dtype == torch.float32 and f2.dtype == torch.float32):
        return False
    if C not in _DIR_BLOCK_WIDTHS or f2.shape[1] != C or v2.shape[0] == 0 or f2.shape[0] == 0:
        return False
    for conv in (conv0, conv1):
        bn, fc = conv.bn, conv.fc
        if not (bn.training or bn.running_mean is None) or fc.bias is None or bn.weight is None:
            return False
        if tuple(fc.weight.shape) != (C, 2 * C) or fc.weight.dtype != torch.float32:
            return False
    return True


def face_chain_start(f2):
    """(Zf, st) for the first block of a chain of Dirac blocks, see _DirBlockChained."""
    return _FaceChainStart.apply(f2)


def dir_block_chained(D, DA, v2, Zf, stf, conv0, conv1, last):
    """One DirResNet2 block inside a chain: returns (v_new, Zf_next, st_next); ``last`` = no Dirac block follows (the
    activated faces then go to a plain buffer, Zf_next / st_next come back empty)."""
    # num_batches_tracked: incremented by the fold kernels (fused.bn_count_batch), here for momentum=None
    return _DirBlockChained.apply(v2, Zf, stf, conv0.bn.weight, conv0.bn.bias, conv0.fc.weight, conv0.fc.bias, conv1.bn.weight,
                                  conv1.bn.bias, conv1.fc.weight, conv1.fc.bias, D, DA, conv0.bn, conv1.bn, bool(last))


def dir_block(D, DA, v2, f2, conv0, conv1):
    """(v_new, f_out) of one DirResNet2 block on rows; conv0 / conv1 are its two GraphConv1x1(2C -> C, "pre")."""
    return _DirBlock.apply(v2, f2, conv0.bn.weight, conv0.bn.bias, conv0.fc.weight, conv0.fc.bias, conv1.bn.weight,
                           conv1.bn.bias, conv1.fc.weight, conv1.fc.bias, D, DA, conv0.bn, conv1.bn)
