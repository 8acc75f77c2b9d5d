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


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def gemm_supported(n_out, k):
    return (n_out in _GEMM_N or (n_out % 256 == 0)) and k % 32 == 0


def fused_supported(z, weight):
    n_out, k = weight.shape
    return (z.is_cuda and z.dtype == torch.float32 and z.dim() == 2 and z.stride(1) == 1 and z.stride(0) % 4 == 0
            and gemm_supported(n_out, k) and gemm_supported(k, n_out) and k % 4 == 0 and 256 % (k // 4) == 0
            and z.data_ptr() % 16 == 0)


def colstats(Z):
    """Per-column mean and biased variance over all rows of Z [rows, C] (sn_colstats_f32)."""
    rows, C = Z.shape
    mean = torch.empty(C, dtype=torch.float32, device=Z.device)
    var = torch.empty(C, dtype=torch.float32, device=Z.device)
    nb = N.lib.sn_colstats_ws_bytes(C)
    ws = _ws(nb, Z.device)
    with torch.cuda.device(Z.device):
        N.call("sn_colstats_f32", _ptr(Z), Z.stride(0), rows, C, _ptr(mean), _ptr(var), _ptr(ws), nb, _stream())
    return mean, var


def gemm_tf32(A, B, bias=None, R=None, rscale=None, out=None, single_pass=False):
    """out[M, N] = A[M, K] @ B[N, K]^T + bias + rscale * R   (3xTF32 tensor-core GEMM; N > 256 is split in column blocks)."""
    M, K = A.shape
    Nn = B.shape[0]
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
    B = B.contiguous()
    step = Nn if Nn in _GEMM_N else 256
    nb = N.lib.sn_gemm_tf32_ws_bytes(step, K)
    flags = N.SN_GEMM_SINGLE_PASS if single_pass else 0
    with torch.cuda.device(A.device):
        for n0 in range(0, Nn, step):
            ws = _ws(nb, A.device)
            Bs = B[n0:n0 + step]
            N.call("sn_gemm_tf32_f32", _ptr(A), A.stride(0), _ptr(Bs), Bs.stride(0),
                   0 if bias is None else bias[n0:].data_ptr(), 0 if R is None else R[:, n0:].data_ptr(),
                   0 if R is None else R.stride(0), 0 if rscale is None else rscale[n0:].data_ptr(),
                   out[:, n0:].data_ptr(), out.stride(0), M, step, K, flags, _ptr(ws), nb, _stream())
    return out


def gemm_tn_supported(m, n):
    return m % 128 == 0 and n % 32 == 0


def gemm_tn_tf32(A, B, single_pass=False):
    """G[M, N] = A[R, M]^T @ B[R, N] (split-K tcgen05 3xTF32, deterministic); M in 128-blocks, N in <=256-blocks."""
    R, M = A.shape
    Nn = B.shape[1]
    G = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
    flags = N.SN_GEMM_SINGLE_PASS if single_pass else 0
    with torch.cuda.device(A.device):
        for m0 in range(0, M, 128):
            for n0 in range(0, Nn, 256):
                n = min(256, Nn - n0)
                nb = N.lib.sn_gemm_tn_tf32_ws_bytes(R, n)
                ws = _ws(nb, A.device)
                N.call("sn_gemm_tn_tf32_f32", A[:, m0:].data_ptr(), A.stride(0), B[:, n0:].data_ptr(), B.stride(0),
                       G[m0:, n0:].data_ptr(), G.stride(0), R, 128, n, flags, _ptr(ws), nb, _stream())
    return G


class _BnLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z, gamma, beta, W, b, residual, running_mean, running_var, training, momentum, eps):
        rows, K = Z.shape
        Nn = W.shape[0]
        dev = Z.device
        if training:
            mean, var = colstats(Z)
        else:
            mean, var = running_mean, running_var
        W = W.contiguous()
        Wf = torch.empty_like(W)                    # [C, 2C]: BatchNorm folded into the Linear
        bf = torch.empty(Nn, dtype=torch.float32, device=dev)
        stk = torch.empty(3, K, dtype=torch.float32, device=dev)          # s, t, rstd
        update = training and running_mean is not None
        with torch.cuda.device(dev):
            N.call("sn_bn_fold_fwd_f32", _ptr(mean), _ptr(var), _ptr(gamma), _ptr(beta), _ptr(W), _ptr(b), Nn, K, float(eps),
                   _ptr(Wf), _ptr(bf), _ptr(stk[0]), _ptr(stk[1]), _ptr(stk[2]), _ptr(running_mean) if update else 0,
                   _ptr(running_var) if update else 0, float(momentum), rows, _stream())
        res = None if residual is None else residual.contiguous()
        Y = gemm_tf32(Z, Wf, bias=bf, R=res)
        ctx.save_for_backward(Z, W, stk, mean)
        ctx.training, ctx.has_res = training, residual is not None
        return Y

    @staticmethod
    def backward(ctx, dY):
        Z, W, stk, mean = ctx.saved_tensors
        dY = dY.contiguous()
        rows, K = Z.shape
        Nn = W.shape[0]
        dev = Z.device
        if gemm_tn_supported(Nn, K):
            G = gemm_tn_tf32(dY, Z)                 # [C, 2C] = dY^T Z: split-K tcgen05 (MN-major operands)
            sdY = colstats(dY)[0] * rows            # colsum(dY) from the same deterministic statistics kernel
        else:
            G = torch.mm(dY.t(), Z)
            sdY = dY.sum(0)
        dW = torch.empty_like(W)
        db = torch.empty(Nn, dtype=torch.float32, device=dev)
        vec = torch.empty(4, K, dtype=torch.float32, device=dev)          # dgamma, dbeta, p, q
        WsT = torch.empty(K, Nn, dtype=torch.float32, device=dev)         # (W diag(s))^T
        with torch.cuda.device(dev):
            N.call("sn_bn_fold_bwd_f32", _ptr(G), _ptr(sdY), _ptr(W), _ptr(stk[0]), _ptr(stk[1]), _ptr(stk[2]), _ptr(mean),
                   Nn, K, rows, 1 if ctx.training else 0, _ptr(dW), _ptr(db), _ptr(vec[0]), _ptr(vec[1]), _ptr(vec[2]),
                   _ptr(vec[3]), _ptr(WsT), _stream())
        if ctx.training:
            dZ = gemm_tf32(dY, WsT, bias=vec[3], R=Z, rscale=vec[2])
        else:
            dZ = gemm_tf32(dY, WsT)
        return dZ, vec[0], vec[1], dW, db, (dY if ctx.has_res else None), None, None, None, None, None


def bn_linear(z, bn, fc, residual=None):
    """GraphConv1x1(batch_norm="pre") on rows: fused path when the shapes allow it, torch composite otherwise."""
    if fused_supported(z, fc.weight) and (residual is None or (residual.shape == (z.shape[0], fc.weight.shape[0]))):
        training = bn.training or bn.running_mean is None
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        momentum = 0.1 if bn.momentum is None else bn.momentum
        return _BnLinear.apply(z, bn.weight, bn.bias, fc.weight, fc.bias, residual, bn.running_mean, bn.running_var,
                               training, momentum, bn.eps)
    y = fc(bn(z))
    return y if residual is None else y + residual
