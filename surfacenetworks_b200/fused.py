"""Fused dense half of a ResNet stage:  Y = Linear(BatchNorm(Z)) (+ residual)  -- reference GraphConv1x1 with
batch_norm="pre" (src/utils/utils_pt.py:91-104) as used by every LapResNet2 / DirResNet2 / AvgResNet2 stage (:156-157).

The reference runs BatchNorm1d on ``x.transpose(1,2)`` and then nn.Linear: two transposes, a statistics pass, a
normalisation pass that writes a second [rows, 2C] tensor, and an fp32 SIMT GEMM.  Here:

  forward   sn_colstats_f32   one pass over Z -> per-column mean / biased variance (training mode)
            fold              W' = W diag(gamma rstd),  b' = b + W (beta - gamma mu rstd)        ([C x 2C], tiny)
            sn_gemm_tf32_f32  Y = Z W'^T + b' (+ residual)   tcgen05 3xTF32, fp32-grade accuracy
  backward  sn_gemm_tn_tf32_f32  G = dY^T Z (one [C x 2C] split-K product) and colsum(dY) give EVERYTHING
            BatchNorm's backward needs:
              dW = G diag(s) + colsum(dY) (x) t,   db = colsum(dY),   dbeta = colsum(dY) W,
              dgamma = rstd (sum_c W .* G - mu dbeta)
            sn_gemm_tf32_f32  dZ = dY (W diag(s)) + p .* Z + q      (p, q from dgamma, dbeta: the BN backward folded)

so Z is read once forward and twice backward, and no normalised copy of Z ever exists.
Shapes the tensor-core kernel does not cover (output width not in {64,128,256} or K % 32 != 0) take the plain
torch composite (F.batch_norm + F.linear on the GPU).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _native as N
from .operators import _ptr, _stream

__all__ = ["bn_linear", "gemm_tf32", "colstats", "fused_supported"]

_GEMM_N = (64, 128, 256)


def bn_momentum(bn):
    """The exponential-average factor nn.BatchNorm uses for this call (evaluate AFTER num_batches_tracked was incremented):
    ``momentum``, or -- for momentum=None, torch's cumulative moving average -- 1 / num_batches_tracked."""
    if bn.momentum is not None:
        return bn.momentum
    if bn.num_batches_tracked is None:
        return 0.0
    return 1.0 / float(bn.num_batches_tracked)          # (reads the counter back: momentum=None is not graph-capturable)


def bn_count_batch(bn, training):
    """nn.BatchNorm's ``num_batches_tracked += 1`` for one training-mode call.  Returns the counter tensor when the fold
    kernel should do the increment (sn_*_fold_fwd_f32: no separate launch per layer), None when it was done here (the
    cumulative-average mode momentum=None needs the new value on the host) or when there is nothing to count."""
    if not training or bn.num_batches_tracked is None:
        return None
    if bn.momentum is None or not bn.num_batches_tracked.is_cuda or bn.num_batches_tracked.dtype != torch.int64:
        bn.num_batches_tracked += 1
        return None
    return bn.num_batches_tracked


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def gemm_supported(n_out, k):
    return (n_out in _GEMM_N or (n_out % 256 == 0)) and k % 32 == 0


def fused_supported(z, weight):
    n_out, k = weight.shape
    return (z.is_cuda and z.dtype == torch.float32 and z.dim() == 2 and z.stride(1) == 1 and z.stride(0) % 4 == 0
            and gemm_supported(n_out, k) and gemm_supported(k, n_out) and k % 4 == 0 and 256 % (k // 4) == 0
            and z.data_ptr() % 16 == 0)


def colstats(Z, mean=None, var=None):
    """Per-column mean and biased variance over all rows of Z [rows, C] (sn_colstats_f32); ``mean`` / ``var`` may be
    given as contiguous [C] views (e.g. halves of a stage's full-width statistics vectors)."""
    rows, C = Z.shape
    if mean is None:
        mean = torch.empty(C, dtype=torch.float32, device=Z.device)
        var = torch.empty(C, dtype=torch.float32, device=Z.device)
    nb = N.lib.sn_colstats_ws_bytes(C)
    ws = _ws(nb, Z.device)
    if N.TIMER is not None:
        N.TIMER.annotate("colstats %dx%d" % (rows, C), 4 * rows * C, 3 * rows * C)
    with torch.cuda.device(Z.device):
        N.call("sn_colstats_f32", _ptr(Z), Z.stride(0), rows, C, _ptr(mean), _ptr(var), _ptr(ws), nb, _stream())
    return mean, var


def elu_colstats(X, out, mean=None, var=None):
    """out = elu(X) and the column statistics (mean, biased variance) of out in ONE pass (sn_elu_colstats_f32)."""
    rows, C = X.shape
    if mean is None:
        mean = torch.empty(C, dtype=torch.float32, device=X.device)
        var = torch.empty(C, dtype=torch.float32, device=X.device)
    nb = N.lib.sn_colstats_ws_bytes(C)
    ws = _ws(nb, X.device)
    if N.TIMER is not None:
        N.TIMER.annotate("elu_colstats %dx%d" % (rows, C), 8 * rows * C, 4 * rows * C)
    with torch.cuda.device(X.device):
        N.call("sn_elu_colstats_f32", _ptr(X), X.stride(0), _ptr(out), out.stride(0), rows, C, _ptr(mean), _ptr(var),
               _ptr(ws), nb, _stream())
    return mean, var


def elu_colstats_supported(X, out):
    C = X.shape[1]
    return (C % 4 == 0 and C <= 1024 and 256 % (C // 4) == 0 and X.stride(0) % 4 == 0 and out.stride(0) % 4 == 0 and
            X.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0 and X.shape[0] > 0)


def _legacy_flag():
    import os
    return N.SN_GEMM_LEGACY_SS if os.environ.get("SN_GEMM_LEGACY_SS") == "1" else 0


def gemm_tf32(A, B, bias=None, R=None, rscale=None, out=None, single_pass=False, group_bias=None, rows_per_group=0,
              elu_bwd_left=False, B_lo=None):
    """out[M, N] = A[M, K] @ B[N, K]^T + bias + group_bias[row // rows_per_group] + rscale * R
    (3xTF32 tensor-core GEMM; N > 256 is split in column blocks; B may be a row-strided view).
    ``B_lo``: B is already split (B = tf32(W), B_lo = W - B, same layout) -- no split kernel, no workspace."""
    M, K = A.shape
    Nn = B.shape[0]
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
    if B_lo is None and (B.stride(1) != 1 or B.stride(0) % 4 or B.data_ptr() % 16):
        B = B.contiguous()
    if group_bias is not None and Nn not in _GEMM_N:
        raise ValueError("group_bias needs N in %s" % (_GEMM_N,))
    step = Nn if Nn in _GEMM_N else 256
    flags = (N.SN_GEMM_SINGLE_PASS if single_pass else 0) | _legacy_flag()
    if elu_bwd_left:
        if R is None or step != Nn:
            raise ValueError("elu_bwd_left needs the residual operand and a single column block")
        flags |= N.SN_GEMM_ELU_BWD_LEFT
    with torch.cuda.device(A.device):
        for n0 in range(0, Nn, step):
            Bs = B[n0:n0 + step]
            if N.TIMER is not None:   # canonical HBM bytes: A read once, C written once, R read once, (pre-split) weights
                N.TIMER.annotate("gemm %dx%dx%d%s" % (M, step, K, "" if R is None else " +R"),
                                 4 * (M * K + M * step * (1 if R is None else 2) + 2 * step * K), 2 * M * step * K)
            common = (0 if bias is None else bias[n0:].data_ptr(), 0 if R is None else R[:, n0:].data_ptr(),
                      0 if R is None else R.stride(0), 0 if rscale is None else rscale[n0:].data_ptr(),
                      _ptr(group_bias), rows_per_group, out[:, n0:].data_ptr(), out.stride(0), M, step, K, flags)
            if B_lo is not None:
                N.call("sn_gemm_tf32_presplit_f32", _ptr(A), A.stride(0), _ptr(Bs), B_lo[n0:n0 + step].data_ptr(),
                       Bs.stride(0), *common, _stream())
            else:
                nb = N.lib.sn_gemm_tf32_ws_bytes(step, K)
                ws = _ws(nb, A.device)
                N.call("sn_gemm_tf32_f32", _ptr(A), A.stride(0), _ptr(Bs), Bs.stride(0), *common, _ptr(ws), nb, _stream())
    return out


def gemm_tf32_act(A, B_hi, B_lo, bias=None, R=None, rscale=None, out=None, act_out=None, mean=None, var=None,
                  group_bias=None, rows_per_group=0, want_raw=True):
    """The dense stage with the NEXT stage's activation fused into the epilogue (sn_gemm_tf32_presplit_act_f32):

        raw = A @ (B_hi + B_lo)^T + bias + group_bias[row // rows_per_group] + rscale * R
        act_out = elu(raw)                       (any row-strided [M, N] view, e.g. the left half of a concat buffer)
        mean / var = column statistics of act_out (when given: contiguous [N] views)

    ``want_raw=False`` skips the raw store entirely (the raw value has no other consumer).  Returns (raw | None, act_out)."""
    M, K = A.shape
    Nn = B_hi.shape[0]
    if Nn not in _GEMM_N:
        raise ValueError("gemm_tf32_act needs N in %s" % (_GEMM_N,))
    if act_out is None:
        act_out = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
    if want_raw and out is None:
        out = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
    if not want_raw:
        out = None
    nb = N.lib.sn_gemm_act_ws_bytes(Nn) if mean is not None else 0
    ws = _ws(nb, A.device) if nb else None
    if N.TIMER is not None:
        N.TIMER.annotate("gemm+act %dx%dx%d%s%s" % (M, Nn, K, "" if R is None else " +R", " raw+act" if want_raw else " act"),
                         4 * (M * K + M * Nn * ((1 if want_raw else 0) + 1 + (0 if R is None else 1)) + 2 * Nn * K), 2 * M * Nn * K)
    with torch.cuda.device(A.device):
        N.call("sn_gemm_tf32_presplit_act_f32", _ptr(A), A.stride(0), _ptr(B_hi), _ptr(B_lo), B_hi.stride(0), _ptr(bias),
               _ptr(R), 0 if R is None else R.stride(0), _ptr(rscale), _ptr(group_bias), rows_per_group, _ptr(out),
               0 if out is None else out.stride(0), _ptr(act_out), act_out.stride(0), _ptr(mean), _ptr(var), M, Nn, K, 0,
               _ptr(ws), nb, _stream())
    return out, act_out


def gemm_tn_supported(m, n):
    return m % 128 == 0 and n % 32 == 0


def gemm_tn_tf32(A, B, single_pass=False, colsum=False):
    """G[M, N] = A[R, M]^T @ B[R, N] (split-K tcgen05 3xTF32, deterministic); M in 128-blocks, N in <=256-blocks.
    ``colsum``: also return the column sums of A [M] from the same pass (sn_gemm_tn_colsum_tf32_f32)."""
    R, M = A.shape
    Nn = B.shape[1]
    G = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
    cs = torch.empty(M, dtype=torch.float32, device=A.device) if colsum else None
    flags = (N.SN_GEMM_SINGLE_PASS if single_pass else 0) | _legacy_flag()
    if colsum and _legacy_flag():
        cs = colstats(A)[0] * R
    with torch.cuda.device(A.device):
        for m0 in range(0, M, 128):
            for n0 in range(0, Nn, 256):
                n = min(256, Nn - n0)
                nb = N.lib.sn_gemm_tn_tf32_ws_bytes(R, n)
                ws = _ws(nb, A.device)
                if N.TIMER is not None:
                    N.TIMER.annotate("gemm_tn R=%d %dx%d" % (R, 128, n), 4 * R * (128 + n), 2 * R * 128 * n)
                if colsum and n0 == 0 and not _legacy_flag():
                    N.call("sn_gemm_tn_colsum_tf32_f32", A[:, m0:].data_ptr(), A.stride(0), B[:, n0:].data_ptr(), B.stride(0),
                           G[m0:, n0:].data_ptr(), G.stride(0), cs[m0:].data_ptr(), R, 128, n, flags, _ptr(ws), nb, _stream())
                else:
                    N.call("sn_gemm_tn_tf32_f32", A[:, m0:].data_ptr(), A.stride(0), B[:, n0:].data_ptr(), B.stride(0),
                           G[m0:, n0:].data_ptr(), G.stride(0), R, 128, n, flags, _ptr(ws), nb, _stream())
    return (G, cs) if colsum else G


def bn_linear_forward(Z, gamma, beta, W, b, residual, running_mean, running_var, training, momentum, eps, left_stats,
                      act=None, counter=None):
    """Forward of GraphConv1x1(batch_norm="pre") on rows (no autograd): statistics pass, BN folded into the weights,
    tcgen05 GEMM with the residual in its epilogue.  Returns (Y, saved) with ``saved`` = what bn_linear_backward needs.
    ``act``: dict(act_out=, mean=, var=, want_raw=) -- the GEMM also emits elu(Y) (and its column statistics) for the next
    stage (gemm_tf32_act); Y is None when ``want_raw`` is false."""
    rows, K = Z.shape
    Nn = W.shape[0]
    dev = Z.device
    if training:
        if left_stats is not None and len(left_stats) == 3:
            # (mean, var, Cl): full-width vectors whose first Cl entries the fused ELU pass has already written; the
            # statistics of the remaining columns land in place -- no concatenation kernels.  Cl == K: the producers
            # of both halves (GEMM activation epilogue, SpMM store path) have written everything.
            mean, var, Cl = left_stats
            if Cl < K:
                colstats(Z[:, Cl:], mean[Cl:], var[Cl:])
        elif left_stats is not None:              # left half already reduced by the fused ELU pass
            Cl = left_stats[0].numel()
            mr, vr = colstats(Z[:, Cl:])
            mean, var = torch.cat([left_stats[0], mr]), torch.cat([left_stats[1], vr])
        else:
            mean, var = colstats(Z)
    else:
        mean, var = running_mean, running_var
    W = W.contiguous()
    Wf = torch.empty(3, Nn, K, dtype=torch.float32, device=dev)       # [C, 2C] x (folded W', tf32(W'), W' - tf32(W'))
    bf = torch.empty(Nn, dtype=torch.float32, device=dev)
    stk = torch.empty(3, K, dtype=torch.float32, device=dev)          # s, t, rstd
    update = training and running_mean is not None
    with torch.cuda.device(dev):
        N.call("sn_bn_fold_fwd_f32", _ptr(mean), _ptr(var), _ptr(gamma), _ptr(beta), _ptr(W), _ptr(b), Nn, K, float(eps),
               _ptr(Wf[0]), _ptr(bf), _ptr(stk[0]), _ptr(stk[1]), _ptr(stk[2]), _ptr(running_mean) if update else 0,
               _ptr(running_var) if update else 0, float(momentum), rows, _ptr(Wf[1]), _ptr(Wf[2]),
               _ptr(counter) if training else 0, _stream())
    res = None if residual is None else residual.contiguous()
    if act is not None:
        Y, _ = gemm_tf32_act(Z, Wf[1], Wf[2], bias=bf, R=res, act_out=act["act_out"], mean=act.get("mean"), var=act.get("var"),
                             want_raw=act.get("want_raw", True))
    else:
        Y = gemm_tf32(Z, Wf[1], bias=bf, R=res, B_lo=Wf[2])
    return Y, (Z, W, stk, mean)


def bn_linear_backward(saved, dY, training, elu_bwd_left=False, elu_bwd_all=False):
    """Backward of bn_linear_forward: returns (dZ, dgamma, dbeta, dW, db).  ``elu_bwd_left``: the left half of Z holds
    activated values elu(x); its gradient leaves the dZ GEMM epilogue already multiplied by elu'(x).  ``elu_bwd_all``: ALL
    of Z holds activated values (the models' heads, conv2(elu(v))): the whole dZ is multiplied by elu'."""
    Z, W, stk, mean = saved
    if dY.stride(1) != 1 or dY.stride(0) % 4 or dY.data_ptr() % 16:     # row-strided views (halves of a dZ) are fine
        dY = dY.contiguous()
    rows, K = Z.shape
    Nn = W.shape[0]
    dev = Z.device
    if gemm_tn_supported(Nn, K):
        # [C, 2C] = dY^T Z: split-K tcgen05; colsum(dY) falls out of the same pass (dY crosses the registers of the warps
        # that feed it to tensor memory)
        G, sdY = gemm_tn_tf32(dY, Z, colsum=True)
    else:
        G = torch.mm(dY.t(), Z)
        sdY = dY.sum(0)
    dW = torch.empty_like(W)
    db = torch.empty(Nn, dtype=torch.float32, device=dev)
    vec = torch.empty(4, K, dtype=torch.float32, device=dev)          # dgamma, dbeta, p, q
    WsT = torch.empty(3, K, Nn, dtype=torch.float32, device=dev)      # (W diag(s))^T: full, tf32 hi, lo
    with torch.cuda.device(dev):
        N.call("sn_bn_fold_bwd_f32", _ptr(G), _ptr(sdY), _ptr(W), _ptr(stk[0]), _ptr(stk[1]), _ptr(stk[2]), _ptr(mean),
               Nn, K, rows, 1 if training else 0, _ptr(dW), _ptr(db), _ptr(vec[0]), _ptr(vec[1]), _ptr(vec[2]),
               _ptr(vec[3]), _ptr(WsT[0]), _ptr(WsT[1]), _ptr(WsT[2]), _stream())
    if elu_bwd_all:
        # (dY Ws + p Z + q) .* elu'(Z): the elu-backward epilogue of the AvgResNet2 stage without its per-mesh terms
        # (p = q = 0 in eval mode, written by the fold kernel)
        dZ = torch.empty(rows, K, dtype=torch.float32, device=dev)
        if N.TIMER is not None:
            N.TIMER.annotate("gemm+elu' %dx%dx%d" % (rows, K, Nn), 4 * (rows * Nn + 2 * rows * K + 2 * K * Nn), 2 * rows * K * Nn)
        with torch.cuda.device(dev):
            N.call("sn_gemm_tf32_presplit_elubwd_f32", _ptr(dY), dY.stride(0), _ptr(WsT[1]), _ptr(WsT[2]), Nn, _ptr(vec[3]),
                   _ptr(Z), Z.stride(0), _ptr(vec[2]), 0, 0, 0, 0, 0, _ptr(dZ), dZ.stride(0), rows, K, Nn, _stream())
    elif training:
        dZ = gemm_tf32(dY, WsT[1], bias=vec[3], R=Z, rscale=vec[2], elu_bwd_left=elu_bwd_left, B_lo=WsT[2])
    else:
        dZ = gemm_tf32(dY, WsT[1], B_lo=WsT[2])
    return dZ, vec[0], vec[1], dW, db


class _BnLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z, gamma, beta, W, b, residual, running_mean, running_var, training, momentum, eps, left_stats,
                elu_bwd_left=False, res_cell=None, counter=None):
        Y, saved = bn_linear_forward(Z, gamma, beta, W, b, residual, running_mean, running_var, training, momentum, eps,
                                     left_stats, counter=counter)
        ctx.save_for_backward(*saved)
        ctx.training, ctx.has_res, ctx.elu_bwd_left = training, residual is not None, elu_bwd_left
        ctx.res_cell = res_cell           # see ops.stage_concat(in_cell=...): where the residual's gradient goes instead
        return Y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dY):
        dZ, dgamma, dbeta, dW, db = bn_linear_backward(ctx.saved_tensors, dY, ctx.training,
                                                       ctx.elu_bwd_left and ctx.training)
        g_res = dY if ctx.has_res else None
        if g_res is not None and ctx.res_cell is not None:
            if g_res.stride(1) != 1 or g_res.stride(0) % 4 or g_res.data_ptr() % 16:
                g_res = g_res.contiguous()
            ctx.res_cell["residual_grad"] = g_res
            g_res = None
        return dZ, dgamma, dbeta, dW, db, g_res, None, None, None, None, None, None, None, None, None


def segment_sum(X, rows_per_seg, n_seg, weight=None, out=None, ws=None):
    """out[s, :] = sum over segment s's rows of weight[r] * X[r, :]  (sn_segment_sum_f32; weight None = 1)."""
    if out is None:
        out = torch.empty(n_seg, X.shape[1], dtype=torch.float32, device=X.device)
    nb = N.lib.sn_segment_sum_ws_bytes(n_seg, X.shape[1])
    if ws is None:
        ws = _ws(nb, X.device)
    with torch.cuda.device(X.device):
        N.call("sn_segment_sum_f32", _ptr(X), X.stride(0), _ptr(weight), rows_per_seg, n_seg, X.shape[1], _ptr(out),
               _ptr(ws), nb, _stream())
    return out


class _EluBnLinear(torch.autograd.Function):
    """GraphConv1x1("pre")(elu(x)) -- the heads of the model stacks (as_rigid_as_possible/models.py:121,151 and twins) -- as
    one node: activation + BatchNorm statistics in one pass (sn_elu_colstats_f32), and elu' applied by the dZ GEMM's
    epilogue in backward (no separate activation / derivative passes over the [rows, C] features)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, W, b, running_mean, running_var, training, momentum, eps, counter):
        rows, K = x.shape
        a = torch.empty(rows, K, dtype=torch.float32, device=x.device)
        if training:
            st = torch.empty(2, K, dtype=torch.float32, device=x.device)
            elu_colstats(x, a, st[0], st[1])
            left = (st[0], st[1], K)
        else:
            from .ops import elu_into
            elu_into(x, a)
            left = None
        Y, saved = bn_linear_forward(a, gamma, beta, W, b, None, running_mean, running_var, training, momentum, eps, left,
                                     counter=counter)
        ctx.save_for_backward(*saved)
        ctx.training = training
        return Y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dY):
        dx, dgamma, dbeta, dW, db = bn_linear_backward(ctx.saved_tensors, dY, ctx.training, elu_bwd_all=True)
        return dx, dgamma, dbeta, dW, db, None, None, None, None, None, None


def elu_bn_linear(x, bn, fc):
    """fc(bn(elu(x))) on rows [rows, K]: fused (see _EluBnLinear) when the widths allow it -- output widths between the
    tensor-core shapes are zero-padded and sliced like ``bn_linear`` -- else the composition."""
    n_out, k = fc.weight.shape
    n_pad = next((n for n in _GEMM_N if n >= n_out), None)
    x = x if (x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0) else x.contiguous()
    ok = (n_pad is not None and (n_pad == n_out or (n_out >= 64 and fc.bias is not None)) and x.is_cuda and
          x.dtype == torch.float32 and fc.bias is not None and elu_colstats_supported(x, x) and k in _GEMM_N and
          gemm_tn_supported(n_pad, k))
    if ok:
        W, b = fc.weight, fc.bias
        if n_pad != n_out:
            W = torch.cat([W, W.new_zeros(n_pad - n_out, k)], 0)
            b = torch.cat([b, b.new_zeros(n_pad - n_out)])
        ok = fused_supported(x, W)
    if not ok:
        return bn_linear(F.elu(x), bn, fc)
    training = bn.training or bn.running_mean is None
    counter = bn_count_batch(bn, training)
    y = _EluBnLinear.apply(x, bn.weight, bn.bias, W, b, bn.running_mean, bn.running_var, training, bn_momentum(bn), bn.eps,
                           counter)
    return y if n_pad == n_out else _SliceCols.apply(y, n_out)


class _AvgStage(torch.autograd.Function):
    """One AvgResNet2 stage (reference src/utils/utils_pt.py:230-243):

        a = elu(x);  avg_b = sum_r mask a / sum_r mask   (global_average, :120-122);
        Y = Linear(BatchNorm([a | avg_b broadcast over the mesh's rows])) (+ residual)

    The broadcast half of the concat buffer is never built: its columns are per-mesh constants, so
    Y[r] = W'_L a[r] + (W'_R avg_b + b') -- a K = C GEMM plus a per-mesh bias in the epilogue (group_bias) -- and its
    BatchNorm statistics are the (equal-weight) statistics of the B per-mesh averages.  Backward: G_L = dY^T a
    (split-K GEMM), G_R = (per-mesh sums of dY)^T avg; the usual folded BatchNorm backward on G = [G_L | G_R];
    dZ_L through the GEMM epilogue; the gradient of the averages is a [B, C] computation that comes back to the rows
    inside the ELU-backward kernel (sn_elu_bwd_group_f32).
    """

    @staticmethod
    def forward(ctx, x, maskw, inv_cnt, gamma, beta, W, b, residual, running_mean, running_var, training, momentum, eps,
                n_seg, rows_per_seg, in_cell=None, res_cell=None, counter=None):
        rows, C = x.shape
        Nn = W.shape[0]
        dev = x.device
        K = 2 * C
        a = torch.empty_like(x)
        avg = torch.empty(n_seg, C, dtype=torch.float32, device=dev)
        st = torch.empty(2, K, dtype=torch.float32, device=dev)                   # batch mean / biased variance of [a | avg]
        nb = N.lib.sn_avg_stage_ws_bytes(n_seg, C)
        ws = _ws(nb, dev)
        if N.TIMER is not None:
            N.TIMER.annotate("avg_pre %dx%d" % (rows, C), 8 * rows * C, 6 * rows * C)
        with torch.cuda.device(dev):
            # activation, left-half statistics and the masked per-mesh sums in ONE pass; then the tiny [B, C] reductions
            N.call("sn_avg_stage_pre_f32", _ptr(x), x.stride(0), _ptr(maskw), _ptr(inv_cnt), rows_per_seg, n_seg, C, _ptr(a),
                   a.stride(0), _ptr(st[0]), _ptr(st[1]), _ptr(avg), _ptr(ws), nb, _stream())
        mean, var = (st[0], st[1]) if training else (running_mean, running_var)
        W = W.contiguous()
        Wf = torch.empty(2, Nn, K, dtype=torch.float32, device=dev)               # tf32(W'), W' - tf32(W')
        stk = torch.empty(3, K, dtype=torch.float32, device=dev)
        u = torch.empty(n_seg, Nn, dtype=torch.float32, device=dev)               # per-mesh bias [B, Nn]
        update = training and running_mean is not None
        with torch.cuda.device(dev):
            N.call("sn_avg_fold_fwd_f32", _ptr(mean), _ptr(var), _ptr(gamma), _ptr(beta), _ptr(W), _ptr(b), Nn, C, float(eps),
                   _ptr(Wf[0]), _ptr(Wf[1]), _ptr(stk[0]), _ptr(stk[1]), _ptr(stk[2]), _ptr(running_mean) if update else 0,
                   _ptr(running_var) if update else 0, float(momentum), rows, _ptr(avg), n_seg, _ptr(u),
                   _ptr(counter) if training else 0, _stream())
        res = None if residual is None else residual.contiguous()
        Y = gemm_tf32(a, Wf[0][:, :C], R=res, group_bias=u, rows_per_group=rows_per_seg, B_lo=Wf[1][:, :C])
        ctx.save_for_backward(a, avg, W, stk, mean, maskw, inv_cnt)
        ctx.training, ctx.has_res, ctx.n_seg, ctx.rps = training, residual is not None, n_seg, rows_per_seg
        # in_cell / res_cell: the two stages of one AvgResNet2 block share a cell.  The stage that holds the block
        # residual (x + ...) leaves its gradient there instead of returning it, and the stage whose INPUT is that same x
        # adds it inside its ELU-backward kernel -- autograd's separate accumulation add disappears.
        ctx.in_cell, ctx.res_cell = in_cell, res_cell
        return Y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dY):
        a, avg, W, stk, mean, maskw, inv_cnt = ctx.saved_tensors
        dY = dY.contiguous()
        rows, C = a.shape
        Nn = W.shape[0]
        K = 2 * C
        dev = a.device
        # (Running the per-mesh sums on a side stream under the weight-gradient GEMM -- the only independent pair of
        # kernels in the step -- was measured in-process against this order: 11.17 vs 11.05 ms per step, slower.)
        SdY = segment_sum(dY, ctx.rps, ctx.n_seg)                                 # [B, Nn] per-mesh sums of dY
        GL = gemm_tn_tf32(dY, a)                                                  # [Nn, C] = dY^T a
        dW = torch.empty_like(W)
        db = torch.empty(Nn, dtype=torch.float32, device=dev)
        vec = torch.empty(4, K, dtype=torch.float32, device=dev)                  # dgamma, dbeta, p, q
        WsT = torch.empty(2, K, Nn, dtype=torch.float32, device=dev)              # (W diag(s))^T: tf32 hi, lo
        gb = torch.empty(ctx.n_seg, C, dtype=torch.float32, device=dev)           # gradient of the per-mesh averages / count
        with torch.cuda.device(dev):
            # G_R = SdY^T avg, the folded BatchNorm backward on [G_L | G_R] and gb, one launch
            N.call("sn_avg_fold_bwd_f32", _ptr(GL), GL.stride(0), _ptr(SdY), _ptr(avg), _ptr(W), _ptr(stk[0]), _ptr(stk[1]),
                   _ptr(stk[2]), _ptr(mean), _ptr(inv_cnt), Nn, C, ctx.n_seg, ctx.rps, 1 if ctx.training else 0, _ptr(dW),
                   _ptr(db), _ptr(vec[0]), _ptr(vec[1]), _ptr(vec[2]), _ptr(vec[3]), _ptr(WsT[0]), _ptr(WsT[1]), _ptr(gb),
                   _stream())
        p, q = vec[2], vec[3]
        dx = torch.empty_like(a)
        g3 = ctx.in_cell.pop("residual_grad", None) if ctx.in_cell is not None else None
        if g3 is not None and (g3.stride(1) != 1 or g3.stride(0) % 4 or g3.data_ptr() % 16):
            g3 = g3.contiguous()
        # dx = ((dY (W_L diag(s_L)) + p a + q + mask * gb[mesh]) .* elu'(a)) + g3: the dZ GEMM with the ELU backward, the
        # per-mesh gradient of the averages and the block-residual gradient in its epilogue (dZ_L is never written)
        if N.TIMER is not None:
            N.TIMER.annotate("gemm+elu' %dx%dx%d%s" % (rows, C, Nn, "" if g3 is None else " +R2"),
                             4 * (rows * Nn + rows * C * (2 if g3 is None else 3) + 2 * C * Nn), 2 * rows * C * Nn)
        with torch.cuda.device(dev):
            N.call("sn_gemm_tf32_presplit_elubwd_f32", _ptr(dY), dY.stride(0), _ptr(WsT[0]), _ptr(WsT[1]), Nn, _ptr(q), _ptr(a),
                   a.stride(0), _ptr(p), _ptr(gb), ctx.rps, _ptr(maskw), _ptr(g3), 0 if g3 is None else g3.stride(0), _ptr(dx),
                   dx.stride(0), rows, C, Nn, _stream())
        g_res = dY if ctx.has_res else None
        if g_res is not None and ctx.res_cell is not None:
            ctx.res_cell["residual_grad"] = dY           # picked up by the block's first stage (runs later in backward)
            g_res = None
        return (dx, None, None, vec[0], vec[1], dW, db, g_res, None, None, None, None, None, None, None, None, None, None)


def avg_stage_supported(x, weight):
    n_out, k = weight.shape
    C = x.shape[1]
    return (x.is_cuda and x.dtype == torch.float32 and k == 2 * C and n_out in _GEMM_N and n_out % 128 == 0 and
            C % 32 == 0 and C <= 256 and 256 % (C // 4) == 0 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and
            x.data_ptr() % 16 == 0)


_MASK_ATTR = "_sn_mask_info"


def mask_info(mask, B, V):
    """(row weights [B*V] fp32, 1 / per-mesh count [B, 1]) of a [B, V, 1] mask, computed once per mask tensor (every
    AvgResNet2 stage of a forward pass sees the same mask: 14 stages in the as_rigid_as_possible DirModel)."""
    key = (mask.data_ptr(), mask._version, tuple(mask.shape), mask.dtype)
    info = getattr(mask, _MASK_ATTR, None)
    if info is not None and info[0] == key:
        return info[1], info[2]
    maskw = mask.reshape(B * V).to(torch.float32).contiguous()
    inv_cnt = (1.0 / mask.reshape(B, V).sum(1, keepdim=True).to(torch.float32)).contiguous()      # [B, 1]
    try:
        setattr(mask, _MASK_ATTR, (key, maskw, inv_cnt))
    except Exception:  # pragma: no cover
        pass
    return maskw, inv_cnt


def avg_stage(x, mask, bn, fc, residual=None, in_cell=None, res_cell=None):
    """elu -> [x | global_average] -> BatchNorm -> Linear (+ residual) for x [B, V, C] (AvgResNet2 stage); returns
    [B*V, C_out] rows, or None when the shapes are outside the fused path."""
    B, V, C = x.shape
    x2 = x.reshape(B * V, C)
    if not avg_stage_supported(x2, fc.weight) or mask.shape[0] != B or mask.shape[1] != V:
        return None
    maskw, inv_cnt = mask_info(mask, B, V)
    training = bn.training or bn.running_mean is None
    counter = bn_count_batch(bn, training)
    momentum = bn_momentum(bn)
    res2 = None if residual is None else residual.reshape(B * V, -1)
    return _AvgStage.apply(x2, maskw, inv_cnt, bn.weight, bn.bias, fc.weight, fc.bias, res2, bn.running_mean,
                           bn.running_var, training, momentum, bn.eps, B, V, in_cell, res_cell, counter)


class _SmallKLinear(torch.autograd.Function):
    """nn.Linear with 3 or 6 input channels on rows (the models' conv1): one write-bound pass forward, one pass over dY
    backward (dW and db together) instead of a SIMT sgemm over 1e5 rows plus a column-sum."""

    @staticmethod
    def forward(ctx, x, W, b):
        rows, K = x.shape
        Nn = W.shape[0]
        x = x.contiguous()
        W = W.contiguous()
        y = torch.empty(rows, Nn, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.call("sn_linear_smallk_fwd_f32", _ptr(x), x.stride(0), _ptr(W), _ptr(b), _ptr(y), y.stride(0), rows, Nn, K, _stream())
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dY):
        x, W = ctx.saved_tensors
        rows, K = x.shape
        Nn = W.shape[0]
        if dY.stride(1) != 1 or dY.stride(0) % 4 or dY.data_ptr() % 16:
            dY = dY.contiguous()
        dW = torch.empty_like(W)
        db = torch.empty(Nn, dtype=torch.float32, device=x.device) if ctx.has_bias else None
        nb = N.lib.sn_linear_smallk_bwd_ws_bytes(Nn, K)
        ws = _ws(nb, x.device)
        with torch.cuda.device(x.device):
            N.call("sn_linear_smallk_bwd_f32", _ptr(dY), dY.stride(0), _ptr(x), x.stride(0), rows, Nn, K, _ptr(dW), _ptr(db),
                   _ptr(ws), nb, _stream())
        dx = dY @ W if ctx.needs_input_grad[0] else None
        return dx, dW, db


def smallk_linear_supported(x, fc):
    n_out, k = fc.weight.shape
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and k in (3, 6) and n_out % 4 == 0 and n_out <= 256
            and 256 % (n_out // 4) == 0 and x.shape[0] > 0 and fc.weight.dtype == torch.float32)


def smallk_linear(x, fc):
    """fc(x) for the 3 / 6-channel input layers (see _SmallKLinear); caller checks ``smallk_linear_supported``."""
    return _SmallKLinear.apply(x, fc.weight, fc.bias)


class _SliceCols(torch.autograd.Function):
    """y[:, :n] of a zero-padded [rows, n_pad] GEMM output; backward writes the padded gradient in one kernel
    (sn_head_pad_grad_f32) instead of autograd's zeros + strided copy."""

    @staticmethod
    def forward(ctx, yp, n):
        ctx.n, ctx.n_pad = n, yp.shape[1]
        return yp[:, :n]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        rows = g.shape[0]
        if g.stride(1) != 1 or g.stride(0) % 4 or g.data_ptr() % 16 or ctx.n % 4:
            out = g.new_zeros(rows, ctx.n_pad)
            out[:, :ctx.n] = g
            return out, None
        out = torch.empty(rows, ctx.n_pad, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            N.call("sn_head_pad_grad_f32", _ptr(g), g.stride(0), _ptr(out), out.stride(0), rows, ctx.n, ctx.n_pad, _stream())
        return out, None


class _HeadAddTiled(torch.autograd.Function):
    """``y + inputs[:, :, -3:].repeat(1, 1, n // 3)`` (as_rigid_as_possible/models.py:152 and twins) in one pass."""

    @staticmethod
    def forward(ctx, y, inputs):
        B, V, n = y.shape
        y2 = y.reshape(B * V, n)
        in2 = inputs.reshape(B * V, inputs.shape[2])
        out = torch.empty(B * V, n, dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            N.call("sn_head_add_tiled_f32", _ptr(y2), y2.stride(0), _ptr(in2), in2.stride(0), in2.shape[1], _ptr(out),
                   out.stride(0), B * V, n, _stream())
        return out.view(B, V, n)

    @staticmethod
    def backward(ctx, g):
        return g, None


def head_add_tiled(y, inputs, times):
    """``y + inputs[:, :, -3:].repeat(1, 1, times)``; fused when y is fp32 CUDA with a 16-byte-aligned row layout."""
    n = y.shape[2]
    y2 = y.reshape(-1, n) if y.dim() == 3 else None
    ok = (y.is_cuda and y.dtype == torch.float32 and y.dim() == 3 and n == 3 * times and n % 4 == 0 and not inputs.requires_grad
          and inputs.dtype == torch.float32 and inputs.is_contiguous() and y2 is not None and y2.stride(1) == 1
          and y2.stride(0) % 4 == 0 and y2.data_ptr() % 16 == 0 and y2.data_ptr() == y.data_ptr())
    if ok:
        return _HeadAddTiled.apply(y, inputs)
    return y + inputs[:, :, -3:].repeat(1, 1, times)


class _MaskedSmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, targets, maskw, scale):
        rows = maskw.numel()
        C = outputs.numel() // rows
        loss = torch.empty((), dtype=torch.float32, device=outputs.device)
        nb = N.lib.sn_masked_smooth_l1_ws_bytes()
        ws = _ws(nb, outputs.device)
        with torch.cuda.device(outputs.device):
            N.call("sn_masked_smooth_l1_fwd_f32", _ptr(outputs), _ptr(targets), _ptr(maskw), rows, C, float(scale), _ptr(loss),
                   _ptr(ws), nb, _stream())
        ctx.save_for_backward(outputs, targets, maskw)
        ctx.scale = float(scale)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        outputs, targets, maskw = ctx.saved_tensors
        rows = maskw.numel()
        C = outputs.numel() // rows
        g = g.contiguous().to(torch.float32)
        d = torch.empty_like(outputs)
        with torch.cuda.device(outputs.device):
            N.call("sn_masked_smooth_l1_bwd_f32", _ptr(outputs), _ptr(targets), _ptr(maskw), _ptr(g), rows, C, ctx.scale, _ptr(d),
                   _stream())
        return d, None, None, None


def masked_smooth_l1(outputs, targets, mask, scale):
    """``scale * F.smooth_l1_loss(outputs * mask, targets, reduction="sum")`` (as_rigid_as_possible/main.py:225-226) as one
    reduction pass forward and one elementwise pass backward; None when the layout is outside the fused path."""
    if not (outputs.is_cuda and outputs.dtype == torch.float32 and targets.dtype == torch.float32 and outputs.dim() == 3
            and outputs.shape == targets.shape and outputs.is_contiguous() and targets.is_contiguous()
            and mask.shape[:2] == outputs.shape[:2] and mask.numel() == outputs.shape[0] * outputs.shape[1]
            and outputs.shape[2] % 4 == 0 and not targets.requires_grad and not mask.requires_grad):
        return None
    maskw = mask_info(mask, outputs.shape[0], outputs.shape[1])[0]
    return _MaskedSmoothL1.apply(outputs, targets, maskw, scale)


class _Correlation(torch.autograd.Function):
    """out[b] = FA[b] @ FB[b]^T (dense_correspondence/models.py:199-203) on the tcgen05 3xTF32 kernel with the TMA-store
    epilogue (sn_gemm_nt_wide_tf32_f32): the [Na, Nb] result is written exactly once, nothing is padded or copied.
    Backward (dFA = G FB, dFB = G^T FA: contractions over thousands of vertices, plain library GEMM shapes) stays on
    torch.bmm."""

    @staticmethod
    def forward(ctx, FA, FB):
        B, Na, K = FA.shape
        Nb = FB.shape[1]
        FA, FB = FA.contiguous(), FB.contiguous()
        out = torch.empty(B, Na, Nb, dtype=torch.float32, device=FA.device)
        hi, lo = torch.empty_like(FB), torch.empty_like(FB)
        with torch.cuda.device(FA.device):
            for b in range(B):
                N.call("sn_split_tf32_f32", _ptr(FB[b]), K, Nb, K, _ptr(hi[b]), _ptr(lo[b]), _stream())
                if N.TIMER is not None:
                    N.TIMER.annotate("correlation %dx%dx%d" % (Na, Nb, K), 4 * (Na * K + 2 * Nb * K + Na * Nb), 2 * Na * Nb * K)
                N.call("sn_gemm_nt_wide_tf32_f32", _ptr(FA[b]), K, _ptr(hi[b]), _ptr(lo[b]), K, _ptr(out[b]), Nb, Na, Nb, K,
                       _stream())
        ctx.save_for_backward(FA, FB)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, G):
        FA, FB = ctx.saved_tensors
        dFA = torch.bmm(G, FB) if ctx.needs_input_grad[0] else None
        dFB = torch.bmm(G.transpose(1, 2), FA) if ctx.needs_input_grad[1] else None
        return dFA, dFB


def correlation(FA, FB):
    """``torch.bmm(FA, FB.transpose(1, 2))`` for [B, Na, K] x [B, Nb, K] feature matrices; the tensor-core kernel when the
    layout allows it (fp32 CUDA, 32 <= K <= 128, K % 4 == 0, Nb % 4 == 0, Nb >= 128), torch.bmm otherwise."""
    ok = (FA.is_cuda and FA.dtype == torch.float32 and FB.dtype == torch.float32 and FA.dim() == 3 and FB.dim() == 3
          and FA.shape[0] == FB.shape[0] and FA.shape[2] == FB.shape[2] and FA.shape[2] <= 128 and FA.shape[2] % 4 == 0
          and FB.shape[1] % 4 == 0 and FA.shape[1] > 0 and FB.shape[1] >= 128 and FA.shape[2] >= 32)
    if ok:
        return _Correlation.apply(FA, FB)
    return torch.bmm(FA, FB.transpose(1, 2))


def bn_linear_is_fused(rows_like, fc, residual_cols=None):
    """Will bn_linear take the fused path for a [rows, 2C] buffer made from ``rows_like`` [rows, C]?  (Same conditions as
    fused_supported, evaluated before the buffer exists.)"""
    n_out, k = fc.weight.shape
    return (rows_like.is_cuda and rows_like.dtype == torch.float32 and rows_like.dim() == 2 and k == 2 * rows_like.shape[1]
            and gemm_supported(n_out, k) and gemm_supported(k, n_out) and k % 4 == 0 and 256 % (k // 4) == 0
            and (residual_cols is None or residual_cols == n_out))


def bn_linear(z, bn, fc, residual=None, res_cell=None):
    """GraphConv1x1(batch_norm="pre") on rows: fused path when the shapes allow it, torch composite otherwise.
    ``res_cell``: honoured by the fused path only -- callers must check ``bn_linear_is_fused`` before handing one over."""
    if fused_supported(z, fc.weight) and (residual is None or (residual.shape == (z.shape[0], fc.weight.shape[0]))):
        training = bn.training or bn.running_mean is None
        counter = bn_count_batch(bn, training)
        momentum = bn_momentum(bn)
        left = getattr(z, "_sn_left_stats", None) if training else None
        if left is not None and not (z.shape[1] - left[0].numel()) % 4 == 0:
            left = None
        cell = getattr(z, "_sn_stage_cell", None)
        fold = bool(cell is not None and training and torch.is_grad_enabled() and fc.weight.shape[1] in _GEMM_N)
        if fold:
            cell["left_premultiplied"] = True
        return _BnLinear.apply(z, bn.weight, bn.bias, fc.weight, fc.bias, residual, bn.running_mean, bn.running_var,
                               training, momentum, bn.eps, left, fold, res_cell, counter)
    # Output widths between the tensor-core shapes (the 128 -> 120 head of the ARAP / dense_correspondence models,
    # conv2 at as_rigid_as_possible/models.py:121): zero-pad the Linear to the next supported width and slice -- the
    # cuBLAS fp32 SIMT GEMMs it replaces were 0.6 ms of the 19 ms step (profiles/r1b_launches_bench_summary.json)
    n_out, k = fc.weight.shape
    n_pad = next((n for n in _GEMM_N if n >= n_out), None)
    if residual is None and n_pad is not None and n_pad != n_out and n_out >= 64 and fc.bias is not None:
        w_pad = torch.cat([fc.weight, fc.weight.new_zeros(n_pad - n_out, k)], 0)
        if fused_supported(z, w_pad):
            b_pad = torch.cat([fc.bias, fc.bias.new_zeros(n_pad - n_out)])
            training = bn.training or bn.running_mean is None
            counter = bn_count_batch(bn, training)
            momentum = bn_momentum(bn)
            y = _BnLinear.apply(z, bn.weight, bn.bias, w_pad, b_pad, None, bn.running_mean, bn.running_var, training,
                                momentum, bn.eps, None, False, None, counter)
            return _SliceCols.apply(y, n_out)
    y = fc(bn(z))
    return y if residual is None else y + residual
