"""``batch_csr(indices, size) -> (col_ind, col_ptr)`` -- reference src/utils/cuda/batch_csr.py:28-59 (+ batch_csr.cu).

The reference launches an NVRTC-compiled kernel (one per ``size``) over the coalesced indices of a 3-D sparse COO
tensor ``[B, R, C]`` and returns

    col_ind [nnz]        int64  column index (inside its mesh) of every entry, in storage order
    col_ptr [B, R + 1]   int64  GLOBAL offset of the first entry of row r of mesh b; col_ptr[b, R] = end of mesh b

Here ``sn_coo_to_csr32`` builds the flattened block-diagonal CSR32 of the same operator on the current stream; the two
returned tensors have the reference's layout and dtype.  Differences, both deliberate:
interior empty rows get correct pointers (the reference kernel leaves them at 0, batch_csr.cu:36-42).  Like the
reference, the indices must be in coalesced (batch, row, col) order.  The converted structure rides along on ``col_ptr`` (attribute ``_sn_structure``) so that
``sparse_bmm`` does not rebuild it.
"""
import torch

from ..operators import _coo_to_csr32, _require_cuda

__all__ = ["BatchCSR", "batch_csr"]


class BatchCSR(object):
    """Callable singleton like the reference's; there is no per-shape kernel cache because shapes are run-time
    arguments of the C ABI."""

    def __call__(self, indices, size):
        _require_cuda(indices, "indices")
        if indices.dim() != 2 or indices.size(0) != 3 or indices.dtype != torch.int64:
            raise ValueError("indices must be the [3, nnz] int64 index tensor of a 3-D sparse COO tensor")
        B, R, C = int(size[0]), int(size[1]), int(size[2])
        indices = indices.contiguous()
        nnz = indices.size(1)
        dummy = torch.zeros(max(nnz, 1), dtype=torch.float32, device=indices.device)
        # storage order is kept (SN_COO_SORTED: the indices of a coalesced tensor, as in sparse_bmm_func.py:39,66-67),
        # so the caller's value vector lines up with col_ind exactly as in the reference
        rowptr, colind, _, _ = _coo_to_csr32(indices[0], indices[1], indices[2], dummy, R, C, B * R, B * C, True)
        col_ptr = torch.as_strided(rowptr, (B, R + 1), (R, 1)).to(torch.int64)
        col_ind = indices[2].clone()
        col_ptr._sn_structure = (rowptr, colind, B, R, C, nnz)
        return col_ind, col_ptr


batch_csr = BatchCSR()
