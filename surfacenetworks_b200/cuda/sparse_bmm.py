"""``sparse_bmm(values, col_ind, col_ptr, size, dense) -> result`` -- reference src/utils/cuda/sparse_bmm.py:28-61
(+ sparse_bmm.cu:16-61): batched CSR x dense product, ``result[b] = S[b] @ dense[b]``.

    values [nnz] fp32 (storage order of the coalesced COO), col_ind / col_ptr from ``batch_csr``,
    size = (B, R, C) of the sparse operand, dense [B, C, K] fp32  ->  result [B, R, K] fp32 (allocated here by torch,
    like the reference, on dense's device; work is enqueued on torch's current stream).

One ``sn_csr_spmm_f32`` launch over the flattened block-diagonal operator replaces the reference's per-shape
NVRTC kernel (one thread per output element, batch on the fastest thread axis).
"""
import torch

from ..operators import CsrOperator, _require_cuda

__all__ = ["SparseBMM", "sparse_bmm"]


class SparseBMM(object):

    def __call__(self, values, col_ind, col_ptr, size, dense):
        _require_cuda(values, "values")
        _require_cuda(dense, "dense")
        B, R, C = int(size[0]), int(size[1]), int(size[2])
        if dense.dim() != 3 or dense.size(0) != B or dense.size(1) != C:
            raise ValueError("dense must be [%d, %d, K], got %s" % (B, C, tuple(dense.shape)))
        if values.dtype != torch.float32 or dense.dtype != torch.float32:
            raise TypeError("values and dense must be float32")
        values = values.contiguous()
        dense = dense.contiguous()
        nnz = values.numel()
        st = getattr(col_ptr, "_sn_structure", None)
        if st is not None and st[2:5] == (B, R, C) and st[5] == values.numel():
            rowptr, colind = st[0], st[1]
            val = values if values.numel() else values.new_zeros(1)
        else:
            # raw (col_ind, col_ptr) in the reference's layout: rebuild the flattened CSR32 with torch ops
            col_ptr = col_ptr.contiguous()
            rowptr = torch.cat([col_ptr[:, :R].reshape(-1), col_ptr[-1:, R]]).to(torch.int32)
            mesh_of_entry = torch.bucketize(torch.arange(nnz, device=values.device), col_ptr[:, R].contiguous(),
                                            right=True)
            colind = (col_ind.to(torch.int64) + mesh_of_entry * C).to(torch.int32)
            if nnz == 0:
                colind, values = colind.new_zeros(1), values.new_zeros(1)
            val = values
        op = CsrOperator(rowptr, colind, val, B * R, B * C, nnz=nnz)
        K = dense.size(2)
        result = dense.new_empty(B, R, K)
        op.apply(dense.view(B * C, K), out=result.view(B * R, K))
        return result


sparse_bmm = SparseBMM()
