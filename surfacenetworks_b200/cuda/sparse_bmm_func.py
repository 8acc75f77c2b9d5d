"""``SparseBMMFunc`` -- reference src/utils/cuda/sparse_bmm_func.py:23-72: autograd wrapper of the batched sparse x
dense product, gradient only for the dense operand (``backward`` returns ``(None, grad_dense)``, :53-72).

    out = SparseBMMFunc.apply(matrix1, matrix2)     # also SparseBMMFunc()(matrix1, matrix2), the reference's style
    matrix1: 3-D sparse COO [B, R, C] (``sparse_cat``);  matrix2: dense [B, C, K];  out: [B, R, K]

The reference rebuilds the CSR of ``matrix1`` in every forward and of its transpose in every backward
(``batch_csr`` at :39 and :66-67; its caches are disabled with ``if False``).  Here both structures are built once per
sparse tensor object and cached on it (``operators.as_csr``).
"""
import torch

from ..operators import as_csr

__all__ = ["SparseBMMFunc"]


class _SparseBMM(torch.autograd.Function):

    @staticmethod
    def forward(ctx, matrix1, matrix2):
        if matrix1.dim() != 3 or matrix2.dim() != 3:
            raise ValueError("SparseBMMFunc expects a 3-D sparse [B, R, C] and a 3-D dense [B, C, K] operand")
        B, R, C = matrix1.shape
        if matrix2.size(0) != B or matrix2.size(1) != C:
            raise ValueError("dense operand must be [%d, %d, K], got %s" % (B, C, tuple(matrix2.shape)))
        op = as_csr(matrix1)
        ctx.op, ctx.dims = op, (B, R, C)
        dense = matrix2.contiguous()
        out = dense.new_empty(B, R, dense.size(2))
        op.apply(dense.view(B * C, -1), out=out.view(B * R, -1))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        B, R, C = ctx.dims
        g = grad_output.contiguous()
        grad = g.new_empty(B, C, g.size(2))
        ctx.op.T.apply(g.view(B * R, -1), out=grad.view(B * C, -1))
        return None, grad


class SparseBMMFunc(object):
    """Callable both ways: ``SparseBMMFunc.apply(m1, m2)`` and the reference's legacy ``SparseBMMFunc()(m1, m2)``
    (utils_pt.py:199,211)."""

    apply = staticmethod(_SparseBMM.apply)

    def __call__(self, matrix1, matrix2):
        return _SparseBMM.apply(matrix1, matrix2)
