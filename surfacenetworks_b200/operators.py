"""Device-resident sparse operators (CSR32 / BSR4) and their construction from torch COO tensors.

The reference hands its layers a torch sparse COO tensor built on the CPU every step
(``sparse_diag_cat`` / ``sparse_cat``, src/utils/utils_pt.py:21-53) and re-derives a CSR from it on
every forward AND backward (src/utils/cuda/sparse_bmm_func.py:39,66-67; caching is stubbed out with
``if False`` at :36,63).  Here an operator is converted ONCE on the GPU into

  * ``CsrOperator``  -- int32 rowptr / colind + fp32 values         (scalar cotangent Laplacian)
  * ``Bsr4Operator`` -- int32 block rowptr / colind + 16 fp32/block (quaternion Dirac D and adjoint D*)

together with (lazily) the transposed structure the backward pass needs, and cached on the torch tensor
it came from, so the per-step hot path never touches COO again.
"""
from __future__ import annotations

import torch

from . import _native as N

__all__ = ["CsrOperator", "Bsr4Operator", "as_csr", "as_bsr4", "clear_cache", "MeshOperatorCache", "Arena",
           "build_dirac_operators", "build_laplacian_operator", "pack_meshes"]


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            "surfacenetworks_b200 runs on CUDA (sm_100a) only; %s is on %s. There is no CPU fallback -- "
            "use the reference implementation for CPU runs." % (what, t.device))


def _check_dense(X, name):
    _require_cuda(X, name)
    if X.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, X.dtype))
    if X.dim() != 2 or (X.shape[1] > 1 and X.stride(1) != 1):
        raise ValueError("%s must be a 2-D row-major matrix (stride(1) == 1); got shape %s strides %s"
                         % (name, tuple(X.shape), X.stride()))


class Arena:
    """Replays a fixed sequence of device allocations: the first pass allocates, later passes (after ``begin()``) hand
    the same tensors back in the same order.  A per-step conversion pipeline run through an arena allocates nothing in
    steady state -- no caching-allocator traffic across streams, no cudaMalloc stalls (measured: the per-step COO
    conversion of bench.py's e2e path jittered between 20 and 230 ms per step without it)."""

    def __init__(self):
        self._bufs, self._i = [], 0

    def begin(self):
        self._i = 0

    def empty(self, numel, dtype, device):
        numel = int(numel)
        if self._i < len(self._bufs):
            t = self._bufs[self._i]
            if t.numel() >= numel and t.dtype == dtype and t.device == torch.device(device):
                self._i += 1
                return t[:numel]
            del self._bufs[self._i:]                 # the sequence changed: re-record from here
        t = torch.empty(numel, dtype=dtype, device=device)
        self._bufs.append(t)
        self._i += 1
        return t


def _empty(arena, numel, dtype, device):
    return torch.empty(numel, dtype=dtype, device=device) if arena is None else arena.empty(numel, dtype, device)


def _coo_to_csr32(batch, row, col, val, rows_per_batch, cols_per_batch, n_rows, n_cols, is_sorted, arena=None):
    """COO (int64, device) -> (rowptr, colind, val) via sn_coo_to_csr32 (replaces batch_csr.cu:13-47)."""
    dev = val.device
    nnz = int(val.numel())
    rowptr = _empty(arena, n_rows + 1, torch.int32, dev)
    colind = _empty(arena, max(nnz, 1), torch.int32, dev)
    out_val = _empty(arena, max(nnz, 1), torch.float32, dev)
    ws_bytes = 0 if is_sorted else N.lib.sn_coo_to_csr32_ws_bytes(nnz, n_rows)
    ws = _empty(arena, max(ws_bytes, 1), torch.uint8, dev)
    with torch.cuda.device(dev):
        N.call("sn_coo_to_csr32", _ptr(batch), _ptr(row), _ptr(col), _ptr(val), nnz, rows_per_batch, cols_per_batch,
               n_rows, n_cols, N.SN_COO_SORTED if is_sorted else 0, _ptr(rowptr), _ptr(colind), _ptr(out_val),
               _ptr(ws), ws_bytes, _stream())
    # nnz == 0 keeps the 1-element allocations so the C ABI never sees a null pointer
    return rowptr, (colind[:nnz] if nnz else colind), (out_val[:nnz] if nnz else out_val), nnz


class _CooSource:
    """The COO arrays an operator was built from -- kept only to build the transpose lazily."""

    def __init__(self, batch, row, col, val, rows_per_batch, cols_per_batch, n_rows, n_cols, is_sorted):
        self.batch, self.row, self.col, self.val = batch, row, col, val
        self.rows_per_batch, self.cols_per_batch = rows_per_batch, cols_per_batch
        self.n_rows, self.n_cols, self.is_sorted = n_rows, n_cols, is_sorted

    def transposed(self):
        return _CooSource(self.batch, self.col, self.row, self.val, self.cols_per_batch, self.rows_per_batch,
                          self.n_cols, self.n_rows, False)

    def to_csr(self, arena=None):
        return _coo_to_csr32(self.batch, self.row, self.col, self.val, self.rows_per_batch, self.cols_per_batch,
                             self.n_rows, self.n_cols, self.is_sorted, arena)

    @staticmethod
    def from_torch(S):
        """Accepts the reference's two sparse layouts: 2-D block-diagonal (sparse_diag_cat) or 3-D
        [B, R, C] (sparse_cat, the layout of the reference's dead SparseBMMFunc path)."""
        _require_cuda(S, "sparse operator")
        if S.layout != torch.sparse_coo:
            raise TypeError("expected a torch sparse COO tensor, got layout %s" % S.layout)
        idx, val = S._indices(), S._values()
        if val.dtype != torch.float32:
            raise TypeError("operator values must be float32 (the reference stores float32, "
                            "add_laplacian.py:61-65); got %s" % val.dtype)
        idx = idx.contiguous()
        val = val.contiguous()
        if S.dim() == 2:
            return _CooSource(None, idx[0], idx[1], val, 0, 0, S.shape[0], S.shape[1], S.is_coalesced())
        if S.dim() == 3:
            B, R, C = S.shape
            return _CooSource(idx[0], idx[1], idx[2], val, R, C, B * R, B * C, S.is_coalesced())
        raise ValueError("sparse operator must be 2-D or 3-D, got %d-D" % S.dim())


class CsrOperator:
    """Scalar CSR32 operator ``S`` [n_rows x n_cols]; ``apply`` computes ``Y = S @ X`` on the GPU."""

    kind = "csr"

    def __init__(self, rowptr, colind, val, n_rows, n_cols, source=None, nnz=None):
        self.rowptr, self.colind, self.val = rowptr, colind, val
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self._nnz = int(val.numel()) if nnz is None else int(nnz)
        self._source = source
        self._T = None

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    @property
    def nnz(self):
        return self._nnz

    @property
    def device(self):
        return self.rowptr.device

    @classmethod
    def from_source(cls, src):
        rowptr, colind, val, nnz = src.to_csr()
        return cls(rowptr, colind, val, src.n_rows, src.n_cols, src, nnz)

    @classmethod
    def from_torch_coo(cls, S):
        return cls.from_source(_CooSource.from_torch(S))

    @property
    def T(self):
        """Transposed operator (what backward applies, sparse_bmm_func.py:66-70) -- built once."""
        if self._T is None:
            if self._source is None:
                raise RuntimeError("transpose unavailable: operator was built without its COO source")
            self._T = CsrOperator.from_source(self._source.transposed())
            self._T._T = self
        return self._T

    def release_source(self):
        self._source = None

    def load_from(self, other, clamp=False):
        """Overwrite this operator's device arrays with ``other``'s (same shape, nnz within capacity), keeping addresses
        (see Bsr4Operator.load_from)."""
        if (other.n_rows, other.n_cols) != (self.n_rows, self.n_cols):
            raise ValueError("operator shapes differ")
        if other.nnz > self.colind.numel() and not clamp:
            raise ValueError("nnz %d exceeds the slot capacity %d" % (other.nnz, self.colind.numel()))
        nnz = min(other.nnz, self.colind.numel())
        self.rowptr.copy_(other.rowptr, non_blocking=True)
        self.colind[:nnz].copy_(other.colind[:nnz], non_blocking=True)
        self.val[:nnz].copy_(other.val[:nnz], non_blocking=True)
        self._nnz = nnz
        return self

    def algorithmic_bytes(self, C):
        """Canonical HBM bytes of one application at feature width C."""
        return 4 * (self.n_rows + 1) + 8 * self.nnz + 4 * self.n_cols * C + 4 * self.n_rows * C

    def flops(self, C):
        return 2 * self.nnz * C

    def apply(self, X, out=None, elu_input=False, direct_gather=False, variant=0):
        """``out[n_rows, C] = S @ (elu(X) if elu_input else X)``; X: [>= n_cols, C] (row stride free)."""
        _check_dense(X, "X")
        if X.shape[0] < self.n_cols:
            raise ValueError("X has %d rows, operator has %d columns" % (X.shape[0], self.n_cols))
        C = X.shape[1]
        if out is None:
            out = torch.empty(self.n_rows, C, dtype=torch.float32, device=X.device)
        _check_dense(out, "out")
        if out.shape[0] != self.n_rows or out.shape[1] != C:
            raise ValueError("out must be [%d, %d], got %s" % (self.n_rows, C, tuple(out.shape)))
        if N.TIMER is not None:
            N.TIMER.annotate("csr %dx%d C=%d" % (self.n_rows, self.n_cols, C), self.algorithmic_bytes(C), self.flops(C))
        with torch.cuda.device(X.device):
            N.call("sn_csr_spmm_f32", _ptr(self.rowptr), _ptr(self.colind), _ptr(self.val), _ptr(X), X.stride(0),
                   _ptr(out), out.stride(0), self.n_rows, C, N.spmm_flags(elu_input, direct_gather, False, variant),
                   _stream())
        return out

    def row_entries_hint(self):
        return 0          # (the CSR kernels take no hint: the shared-memory variant loses on the Laplacian, 47 -> 50 us)

    def apply_stats(self, X, out, mean, var):
        """``out = S @ X`` and the column statistics of ``out`` (mean, biased variance over all rows) in one launch, or
        None if unsupported for this shape (then run ``apply`` + ``fused.colstats``)."""
        return _apply_stats(self, "sn_csr_spmm_stats_f32", self.n_rows, self.n_cols,
                            (_ptr(self.rowptr), _ptr(self.colind), _ptr(self.val)), X, out, mean, var)

    def apply_epilogue(self, X, G=None, A=None, out=None, G2=None):
        """``(S @ X + G) * elu'(A) + G2`` (A = activated values) in one launch, or None if unsupported for this shape."""
        return _apply_epilogue(self, "sn_csr_spmm_epilogue_f32", self.n_rows, self.n_cols,
                               (_ptr(self.rowptr), _ptr(self.colind), _ptr(self.val)), X, G, A, out, G2)


def _apply_stats(op, entry, n_rows, n_cols, ptrs, X, out, mean, var):
    """``out = op @ X`` plus the per-column mean / biased variance of ``out`` from the same launch
    (sn_*_spmm_stats_f32); returns None when the row-group kernel does not cover the shape."""
    _check_dense(X, "X")
    _check_dense(out, "out")
    C = X.shape[1]
    if X.shape[0] < n_cols:
        raise ValueError("X has %d rows, operator has %d columns" % (X.shape[0], n_cols))
    if out.shape[0] != n_rows or out.shape[1] != C:
        raise ValueError("out must be [%d, %d], got %s" % (n_rows, C, tuple(out.shape)))
    if mean.numel() != C or var.numel() != C or not (mean.is_contiguous() and var.is_contiguous()):
        raise ValueError("mean / var must be contiguous [%d] vectors" % C)
    nb = N.lib.sn_spmm_stats_ws_bytes(C)
    ws = torch.empty(max(int(nb), 1), dtype=torch.uint8, device=X.device)
    if N.TIMER is not None:
        N.TIMER.annotate("%s %dx%d C=%d +stats" % (op.kind, n_rows, n_cols, C), op.algorithmic_bytes(C), op.flops(C))
    with torch.cuda.device(X.device):
        rc = N.call(entry, *ptrs, _ptr(X), X.stride(0), _ptr(out), out.stride(0), n_rows, C, _ptr(mean), _ptr(var),
                    N.spmm_flags(row_entries=op.row_entries_hint()), _ptr(ws), nb, _stream(), soft_unsupported=True)
    return None if rc == N.SN_ERR_UNSUPPORTED else out


# The store-path epilogue reads its operand rows through a shared-memory landing zone filled when the row starts; False
# selects the loads at the row's end (SN_SPMM_VARIANT(8); A/B timings, tools/ab_step.py --toggle operators.EPILOGUE_STAGED)
EPILOGUE_STAGED = True


def _apply_epilogue(op, entry, n_rows, n_cols, ptrs, X, G, A, out, G2=None):
    """``out = (op @ X + G) * elu'(A) + G2`` in one launch (sn_*_spmm_epilogue_f32); returns None when the row-group kernel
    does not cover the shape, so that the caller can run the separate passes."""
    _check_dense(X, "X")
    C = X.shape[1]
    if X.shape[0] < n_cols:
        raise ValueError("X has %d rows, operator has %d columns" % (X.shape[0], n_cols))
    if out is None:
        out = torch.empty(n_rows, C, dtype=torch.float32, device=X.device)
    for t, name in ((out, "out"), (G, "G"), (A, "A"), (G2, "G2")):
        if t is not None:
            _check_dense(t, name)
            if t.shape[0] != n_rows or t.shape[1] != C:
                raise ValueError("%s must be [%d, %d], got %s" % (name, n_rows, C, tuple(t.shape)))
    if N.TIMER is not None:      # canonical SpMM bytes + the epilogue operands, each read once
        extra = 4 * n_rows * C * ((G is not None) + (A is not None) + (G2 is not None))
        N.TIMER.annotate("%s %dx%d C=%d +epilogue" % (op.kind, n_rows, n_cols, C), op.algorithmic_bytes(C) + extra, op.flops(C))
    with torch.cuda.device(X.device):
        rc = N.call(entry, *ptrs, _ptr(X), X.stride(0), _ptr(out), out.stride(0), n_rows, C, _ptr(G),
                    0 if G is None else G.stride(0), _ptr(A), 0 if A is None else A.stride(0), _ptr(G2),
                    0 if G2 is None else G2.stride(0), 0 if EPILOGUE_STAGED else (8 << 8), _stream(), soft_unsupported=True)
    return None if rc == N.SN_ERR_UNSUPPORTED else out


class Bsr4Operator:
    """4x4-block CSR operator for the quaternion Dirac operators; scalar shape [4*n_brows x 4*n_bcols].

    ``apply`` implements the reference's ``view`` semantics (utils_pt.py:201-203): X is [n_bcols, C] node
    features, the q-th quarter of the channel vector is quaternion component q.
    """

    kind = "bsr4"

    def __init__(self, browptr, bcolind, bval, n_brows, n_bcols, source=None, n_blocks=None, max_row_blocks=0):
        self.browptr, self.bcolind, self.bval = browptr, bcolind, bval
        self.n_brows, self.n_bcols = int(n_brows), int(n_bcols)
        self._n_blocks = int(bcolind.numel()) if n_blocks is None else int(n_blocks)
        # largest number of blocks in a block-row (3 for D, the largest vertex valence for D*); informational
        self.max_row_blocks = int(max_row_blocks)
        self._source = source
        self._T = None

    @property
    def shape(self):
        return (4 * self.n_brows, 4 * self.n_bcols)

    @property
    def n_blocks(self):
        return self._n_blocks

    @property
    def device(self):
        return self.browptr.device

    @classmethod
    def from_source(cls, src, arena=None, block_capacity=None):
        """``arena``: take every device buffer from a replayable Arena (per-step conversions allocate nothing);
        ``block_capacity``: an upper bound of the block count known to the caller -- skips the read-back of the exact
        count, so the whole conversion is stream-ordered (the operator then reports the capacity as ``n_blocks``)."""
        if src.n_rows % 4 or src.n_cols % 4:
            raise ValueError("Dirac operator shape must be a multiple of 4 in both dims, got %dx%d"
                             % (src.n_rows, src.n_cols))
        rowptr, colind, val, _ = src.to_csr(arena)
        dev = val.device
        n_brows = src.n_rows // 4
        browptr = _empty(arena, n_brows + 1, torch.int32, dev)
        ws_bytes = N.lib.sn_csr32_to_bsr4_ws_bytes(src.n_rows)
        ws = _empty(arena, max(ws_bytes, 1), torch.uint8, dev)
        with torch.cuda.device(dev):
            N.call("sn_csr32_to_bsr4_count", _ptr(rowptr), _ptr(colind), src.n_rows, _ptr(browptr), _ptr(ws),
                   ws_bytes, _stream())
            if block_capacity is None:
                # one small read-back per operator conversion (not per step): block count + densest block-row
                stats = torch.stack([browptr[-1], (browptr[1:] - browptr[:-1]).max() if n_brows else browptr[-1]]).tolist()
                nb, max_row_blocks = int(stats[0]), int(stats[1])
            else:
                nb, max_row_blocks = int(block_capacity), 0
            bcolind = _empty(arena, max(nb, 1), torch.int32, dev)
            bval = _empty(arena, max(nb, 1) * 16, torch.float32, dev)
            N.call("sn_csr32_to_bsr4_fill", _ptr(rowptr), _ptr(colind), _ptr(val), src.n_rows, _ptr(browptr),
                   _ptr(bcolind), _ptr(bval), _stream())
        if nb:
            bcolind, bval = bcolind[:nb], bval[:nb * 16]
        return cls(browptr, bcolind, bval, n_brows, src.n_cols // 4, src, nb, max_row_blocks)

    @classmethod
    def from_torch_coo(cls, S, arena=None, block_capacity=None):
        return cls.from_source(_CooSource.from_torch(S), arena, block_capacity)

    @property
    def T(self):
        if self._T is None:
            self.build_transpose()
        return self._T

    def build_transpose(self, arena=None, block_capacity=None):
        """Builds (once) the transposed operator backward applies; ``arena`` / ``block_capacity`` as in from_source."""
        if self._T is None:
            if self._source is None:
                raise RuntimeError("transpose unavailable: operator was built without its COO source")
            self._T = Bsr4Operator.from_source(self._source.transposed(), arena, block_capacity)
            self._T._T = self
        return self._T

    def release_source(self):
        self._source = None

    def load_from(self, other, clamp=False):
        """Overwrite this operator's device arrays with ``other``'s (same shape, block count within capacity), keeping
        the buffers' addresses -- lets a captured CUDA graph be replayed on a new batch's operator.  ``clamp``: ``other``
        was built without a read-back and reports its buffer CAPACITY as block count (``from_source(block_capacity=)``,
        ``build_dirac_operators(sync=False)``); copy what fits the slot (the caller guarantees the real count does)."""
        if (other.n_brows, other.n_bcols) != (self.n_brows, self.n_bcols):
            raise ValueError("operator shapes differ")
        if other.n_blocks > self.bcolind.numel() and not clamp:
            raise ValueError("block count %d exceeds the slot capacity %d" % (other.n_blocks, self.bcolind.numel()))
        nb = min(other.n_blocks, self.bcolind.numel())
        self.browptr.copy_(other.browptr, non_blocking=True)
        self.bcolind[:nb].copy_(other.bcolind[:nb], non_blocking=True)
        self.bval[:16 * nb].copy_(other.bval[:16 * nb], non_blocking=True)
        self._n_blocks, self.max_row_blocks = nb, other.max_row_blocks
        return self

    def algorithmic_bytes(self, C):
        return 4 * (self.n_brows + 1) + 68 * self.n_blocks + 4 * self.n_bcols * C + 4 * self.n_brows * C

    def flops(self, C, stored_nnz_per_block=12):
        """Reference-stored nnz count (12 per Dirac block) by default; pass 16 for dense-block FLOPs."""
        return 2 * stored_nnz_per_block * self.n_blocks * (C // 4)

    def apply(self, X, out=None, elu_input=False, direct_gather=False, smem_stream=False, variant=0):
        _check_dense(X, "X")
        if X.shape[0] < self.n_bcols:
            raise ValueError("X has %d rows, operator has %d block columns" % (X.shape[0], self.n_bcols))
        C = X.shape[1]
        if C % 4:
            raise ValueError("feature width must be divisible by 4 for the quaternion view, got %d" % C)
        if out is None:
            out = torch.empty(self.n_brows, C, dtype=torch.float32, device=X.device)
        _check_dense(out, "out")
        if out.shape[0] != self.n_brows or out.shape[1] != C:
            raise ValueError("out must be [%d, %d], got %s" % (self.n_brows, C, tuple(out.shape)))
        if N.TIMER is not None:
            N.TIMER.annotate("bsr4 %dx%d C=%d" % (self.n_brows, self.n_bcols, C), self.algorithmic_bytes(C), self.flops(C))
        with torch.cuda.device(X.device):
            flags = N.spmm_flags(elu_input, direct_gather, smem_stream, variant, row_entries=self.row_entries_hint())
            N.call("sn_bsr4_spmm_f32", _ptr(self.browptr), _ptr(self.bcolind), _ptr(self.bval),
                   _ptr(X), X.stride(0), _ptr(out), out.stride(0), self.n_brows, C, flags, _stream())
        return out

    def row_entries_hint(self):
        """Mean blocks per block-row, rounded up (D: 3, D*: the mean vertex valence, ~6): the SN_SPMM_ROW_ENTRIES hint."""
        return min(15, -(-self.n_blocks // max(self.n_brows, 1))) if self.n_blocks > 0 else 0

    def apply_stats(self, X, out, mean, var):
        """``out = S @ X`` and the column statistics of ``out`` in one launch (see CsrOperator.apply_stats)."""
        return _apply_stats(self, "sn_bsr4_spmm_stats_f32", self.n_brows, self.n_bcols,
                            (_ptr(self.browptr), _ptr(self.bcolind), _ptr(self.bval)), X, out, mean, var)

    def apply_epilogue(self, X, G=None, A=None, out=None, G2=None):
        """``(S @ X + G) * elu'(A) + G2`` (A = activated values) in one launch, or None if unsupported for this shape."""
        if X.shape[1] % 4:
            raise ValueError("feature width must be divisible by 4 for the quaternion view, got %d" % X.shape[1])
        return _apply_epilogue(self, "sn_bsr4_spmm_epilogue_f32", self.n_brows, self.n_bcols,
                               (_ptr(self.browptr), _ptr(self.bcolind), _ptr(self.bval)), X, G, A, out, G2)


# ---------------------------------------------------------------------------------------------------
# GPU operator construction: padded mesh batch -> batch operators, no scipy / COO round trip.
class _StructureSource:
    """Stand-in for the COO source of an operator that was built directly in CSR32 / BSR4 form: expands the stored
    structure to scalar COO on the device the first time the transpose is requested."""

    def __init__(self, kind, ptr, ind, val, n_rows, n_cols, n_entries):
        self.kind, self.ptr, self.ind, self.val = kind, ptr, ind, val
        self.n_rows, self.n_cols, self.n_entries = n_rows, n_cols, n_entries

    def _coo(self):
        n = self.n_entries
        counts = (self.ptr[1:] - self.ptr[:-1]).to(torch.int64)
        major = torch.repeat_interleave(torch.arange(counts.numel(), device=self.ptr.device), counts, output_size=n)
        minor = self.ind[:n].to(torch.int64)
        if self.kind == "csr":
            return major, minor, self.val[:n].contiguous()
        # bval[16 k + 4 q + s] = B_k[(q + s) % 4][q]
        q = torch.arange(4, device=self.ptr.device).view(1, 4, 1)
        sft = torch.arange(4, device=self.ptr.device).view(1, 1, 4)
        row = (4 * major.view(-1, 1, 1) + (q + sft) % 4).reshape(-1)
        col = (4 * minor.view(-1, 1, 1) + q + 0 * sft).reshape(-1)
        return row, col, self.val[:16 * n].contiguous()

    def transposed(self):
        row, col, val = self._coo()
        return _CooSource(None, col, row, val, 0, 0, self.n_cols, self.n_rows, False)


def _check_mesh_batch(V, F):
    _require_cuda(V, "V")
    _require_cuda(F, "F")
    if V.dim() != 3 or V.size(2) != 3 or F.dim() != 3 or F.size(2) != 3 or V.size(0) != F.size(0):
        raise ValueError("expected V [B, v_pad, 3] and F [B, f_pad, 3], got %s and %s" % (tuple(V.shape), tuple(F.shape)))
    return V.to(torch.float64).contiguous(), F.to(torch.int32).contiguous()


def _mesh_ws(n, v_pad, f_pad, dev):
    nbytes = N.lib.sn_mesh_ws_bytes(n, v_pad, f_pad)
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev), nbytes


def build_dirac_operators(V, F, with_transposes=True, sync=True, buffers=None):
    """Batch Dirac operator ``D`` [B*f_pad x B*v_pad] and adjoint ``D*`` [B*v_pad x B*f_pad] (block rows x block
    columns) built on the GPU from padded positions ``V`` [B, v_pad, 3] and faces ``F`` [B, f_pad, 3] (local vertex
    indices, padding faces = -1).  Same values as the reference's mesh.dirac (src/utils/mesh.py:35-64) after its
    ``.astype('float32')``; replaces the offline numpy construction + sparse_diag_cat + upload.

    ``with_transposes``: D^T and (D*)^T (what backward applies) come out of the same kernels and are attached as
    ``D.T`` / ``DA.T``.  ``sync=False`` skips the two small read-backs (status check, block count): the operators then
    carry the capacity 3*B*f_pad as their block count (only the accounting of ``algorithmic_bytes`` is affected), so
    the whole construction is stream-ordered -- for per-step geometry inside a pipelined training loop.  ``buffers``:
    a dict that keeps the output / workspace tensors between calls (same batch shape), so a per-step rebuild allocates
    nothing."""
    V, F = _check_mesh_batch(V, F)
    n, v_pad, f_pad = V.size(0), V.size(1), F.size(1)
    dev = V.device
    cap = max(3 * n * f_pad, 1)
    i32 = dict(dtype=torch.int32, device=dev)
    store = buffers if buffers is not None else {}
    if store.get("shape") != (n, v_pad, f_pad, bool(with_transposes), str(dev)):
        store.clear()
        store["shape"] = (n, v_pad, f_pad, bool(with_transposes), str(dev))

    def buf(name, numel, dtype):
        if name not in store:
            store[name] = torch.empty(numel, dtype=dtype, device=dev)
        return store[name]
    d_ptr, d_ind = buf("d_ptr", n * f_pad + 1, torch.int32), buf("d_ind", cap, torch.int32)
    a_ptr, a_ind = buf("a_ptr", n * v_pad + 1, torch.int32), buf("a_ind", cap, torch.int32)
    d_val, a_val = buf("d_val", 16 * cap, torch.float32), buf("a_val", 16 * cap, torch.float32)
    status = buf("status", 1, torch.int32)
    nbytes = N.lib.sn_mesh_ws_bytes(n, v_pad, f_pad)
    ws = buf("ws", max(nbytes, 1), torch.uint8)
    dt_ind = dt_val = at_ind = at_val = None
    if with_transposes:
        dt_ind, at_ind = buf("dt_ind", cap, torch.int32), buf("at_ind", cap, torch.int32)
        dt_val, at_val = buf("dt_val", 16 * cap, torch.float32), buf("at_val", 16 * cap, torch.float32)
    with torch.cuda.device(dev):
        N.call("sn_mesh_dirac_bsr4", _ptr(V), _ptr(F), n, v_pad, f_pad, _ptr(d_ptr), _ptr(d_ind), _ptr(d_val),
               _ptr(a_ptr), _ptr(a_ind), _ptr(a_val), _ptr(dt_ind), _ptr(dt_val), _ptr(at_ind), _ptr(at_val),
               _ptr(status), _ptr(ws), nbytes, _stream())
    if sync:
        nb = int(d_ptr[-1].item())                   # one read-back per construction, like from_source
        max_a = int((a_ptr[1:] - a_ptr[:-1]).max().item()) if n * v_pad else 0
    else:
        nb, max_a = cap, 0
    keep = max(nb, 1)
    D = Bsr4Operator(d_ptr, d_ind[:keep], d_val[:16 * keep], n * f_pad, n * v_pad,
                     _StructureSource("bsr4", d_ptr, d_ind, d_val, 4 * n * f_pad, 4 * n * v_pad, nb) if sync else None,
                     nb, 3 if nb else 0)
    DA = Bsr4Operator(a_ptr, a_ind[:keep], a_val[:16 * keep], n * v_pad, n * f_pad,
                      _StructureSource("bsr4", a_ptr, a_ind, a_val, 4 * n * v_pad, 4 * n * f_pad, nb) if sync else None,
                      nb, max_a)
    if with_transposes:
        D._T = Bsr4Operator(a_ptr, dt_ind[:keep], dt_val[:16 * keep], n * v_pad, n * f_pad, None, nb, max_a)
        DA._T = Bsr4Operator(d_ptr, at_ind[:keep], at_val[:16 * keep], n * f_pad, n * v_pad, None, nb, 3 if nb else 0)
        D._T._T, DA._T._T = D, DA
    D.status = DA.status = status                    # device int32: 0, or the largest per-vertex face count when some
    #                                                  vertex exceeds 64 faces (informational: any valence is supported)
    return D, DA


def build_laplacian_operator(V, F):
    """Batch cotangent Laplacian ``A^-1 (D - W)`` [B*v_pad x B*v_pad] built on the GPU (reference recipe
    src/as_rigid_as_possible/add_laplacian.py:50-56 over mesh.py:17-26,67-80,102-112 and graph.py:40-49)."""
    V, F = _check_mesh_batch(V, F)
    n, v_pad, f_pad = V.size(0), V.size(1), F.size(1)
    dev = V.device
    cap = max(n * (v_pad + 6 * f_pad), 1)
    rowptr = torch.empty(n * v_pad + 1, dtype=torch.int32, device=dev)
    colind = torch.empty(cap, dtype=torch.int32, device=dev)
    val = torch.empty(cap, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws, nbytes = _mesh_ws(n, v_pad, f_pad, dev)
    with torch.cuda.device(dev):
        N.call("sn_mesh_laplacian_csr", _ptr(V), _ptr(F), n, v_pad, f_pad, _ptr(rowptr), _ptr(colind), _ptr(val),
               _ptr(status), _ptr(ws), nbytes, _stream())
    nnz = int(rowptr[-1].item())
    return CsrOperator(rowptr, colind[:max(nnz, 1)], val[:max(nnz, 1)], n * v_pad, n * v_pad,
                       _StructureSource("csr", rowptr, colind, val, n * v_pad, n * v_pad, nnz), nnz)


def pack_meshes(meshes, device, v_pad=None, f_pad=None):
    """List of (V [nv, 3], F [nf, 3]) numpy meshes -> padded device tensors (V [B, v_pad, 3] fp64 zero-padded,
    F [B, f_pad, 3] int32 padded with -1) for build_dirac_operators / build_laplacian_operator."""
    import numpy as np
    v_pad = v_pad or max(v.shape[0] for v, _ in meshes)
    f_pad = f_pad or max(f.shape[0] for _, f in meshes)
    Vp = np.zeros((len(meshes), v_pad, 3), dtype=np.float64)
    Fp = np.full((len(meshes), f_pad, 3), -1, dtype=np.int32)
    for i, (v, f) in enumerate(meshes):
        Vp[i, :v.shape[0]] = v
        Fp[i, :f.shape[0]] = f
    return torch.from_numpy(Vp).to(device), torch.from_numpy(Fp).to(device)


# ---------------------------------------------------------------------------------------------------
# Cache: converted operators live on the torch tensor object they came from (``S._sn_ops``).
_ATTR = "_sn_ops"


def _cached(S, kind, builder):
    cache = getattr(S, _ATTR, None)
    if cache is None:
        cache = {}
        try:
            setattr(S, _ATTR, cache)
        except Exception:  # pragma: no cover - exotic tensor subclasses
            return builder(S)
    # values AND indices identify the operator: an in-place edit of either (or a reshape) must rebuild the structures
    idx = S._indices()
    key = (kind, S._values().data_ptr(), S._values()._version, idx.data_ptr(), idx._version, S._nnz(), tuple(S.shape))
    op = cache.get(key)
    if op is None:
        cache.clear()
        op = cache[key] = builder(S)
    return op


def clear_cache(S):
    if hasattr(S, _ATTR):
        getattr(S, _ATTR).clear()


def as_csr(S):
    """torch sparse COO (2-D block-diagonal or 3-D batched) or CsrOperator -> CsrOperator (cached)."""
    if isinstance(S, CsrOperator):
        return S
    if isinstance(S, Bsr4Operator):
        raise TypeError("expected a scalar (Laplacian) operator, got a Bsr4Operator")
    return _cached(S, "csr", CsrOperator.from_torch_coo)


def as_bsr4(S):
    """torch sparse COO (2-D block-diagonal or 3-D batched) or Bsr4Operator -> Bsr4Operator (cached)."""
    if isinstance(S, Bsr4Operator):
        return S
    if isinstance(S, CsrOperator):
        raise TypeError("expected a Dirac (4x4-block) operator, got a CsrOperator")
    return _cached(S, "bsr4", Bsr4Operator.from_torch_coo)


# ---------------------------------------------------------------------------------------------------
class MeshOperatorCache:
    """Per-mesh operators resident on the GPU + batch assembly on the GPU.

    The reference converts every sampled mesh's scipy operators to torch COO, concatenates, sorts and uploads them on
    every training step (src/as_rigid_as_possible/main.py:142-185 -> utils_pt.sparse_diag_cat).  With this cache each
    mesh is converted once (``add``); ``assemble`` builds the block-diagonal batch operator -- and its transpose, for
    backward -- from the cached parts with sn_assemble_block_diag, optionally straight into existing operator
    buffers (``out=``) so a captured CUDA graph can be replayed on the new batch.
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self._ops = {}

    def __contains__(self, key):
        return key in self._ops

    def add(self, key, S, kind):
        """Register mesh ``key``'s operator: ``S`` is a torch sparse COO tensor (any device) or a scipy matrix;
        ``kind`` is "csr" (Laplacian) or "bsr4" (Dirac / adjoint).  Converts it and its transpose on the GPU."""
        if not isinstance(S, torch.Tensor):
            import numpy as np
            S = S.tocoo()
            S = torch.sparse_coo_tensor(torch.from_numpy(np.stack([S.row, S.col]).astype(np.int64)),
                                        torch.from_numpy(S.data.astype(np.float32)), S.shape)
        S = S.to(self.device)
        op = (CsrOperator if kind == "csr" else Bsr4Operator).from_torch_coo(S if S.is_coalesced() else S.coalesce())
        op.T                      # build the transpose once
        op.release_source()
        op.T.release_source()
        self._ops[(key, kind)] = op
        return op

    def get(self, key, kind):
        return self._ops[(key, kind)]

    @staticmethod
    def _arrays(op):
        if op.kind == "csr":
            return op.rowptr, op.colind, op.val, op.n_rows, op.nnz
        return op.browptr, op.bcolind, op.bval, op.n_brows, op.n_blocks

    @staticmethod
    def _n_cols(op):
        return op.n_cols if op.kind == "csr" else op.n_bcols

    def _table(self, parts, kind, rows_pad, cols_pad):
        """(device table [n, 6] int64 of (rowptr, colind, val, n_rows, n_entries, entry offset), total entries, largest
        row length) of one block-diagonal assembly: the host-side half of ``_assemble_one``."""
        table, off = [], 0
        for op in parts:
            rp, ci, va, n_rows, n_ent = self._arrays(op)
            if n_rows > rows_pad:
                raise ValueError("mesh operator has %d rows, more than the padded size %d" % (n_rows, rows_pad))
            if self._n_cols(op) > cols_pad:       # an oversized mesh would bleed into the next block's columns
                raise ValueError("mesh operator has %d columns, more than the padded size %d" % (self._n_cols(op), cols_pad))
            table += [rp.data_ptr(), ci.data_ptr(), va.data_ptr(), n_rows, n_ent, off]
            off += n_ent
        table = torch.tensor(table, dtype=torch.int64).to(self.device, non_blocking=True)
        mrb = 0 if kind == "csr" else max(op.max_row_blocks for op in parts)
        return table, off, mrb

    def plan(self, keys, kind, rows_pad, cols_pad):
        """The host-side half of ``assemble`` for the meshes ``keys`` -- validation, the two pointer tables (operator and
        transpose) and their upload on the CURRENT stream -- so that a data loader can prepare the next batch while the
        GPU is busy; ``assemble(..., plan=p)`` then only launches the assembly kernels.  The plan stays valid as long as
        the cached operators it points to."""
        parts = [self._ops[(k, kind)] for k in keys]
        return {"kind": kind, "n": len(parts), "rows_pad": rows_pad, "cols_pad": cols_pad,
                "fwd": self._table(parts, kind, rows_pad, cols_pad),
                "bwd": self._table([p.T for p in parts], kind, cols_pad, rows_pad)}

    def _assemble_one(self, tab, n, kind, rows_pad, cols_pad, out):
        dev = self.device
        table, off, mrb = tab
        table.record_stream(torch.cuda.current_stream(dev))      # (a plan may have been uploaded on another stream)
        vpe = 1 if kind == "csr" else 16
        if out is None:
            rowptr = torch.empty(n * rows_pad + 1, dtype=torch.int32, device=dev)
            colind = torch.empty(max(off, 1), dtype=torch.int32, device=dev)
            val = torch.empty(max(off, 1) * vpe, dtype=torch.float32, device=dev)
        else:
            rowptr, colind, val = self._arrays(out)[:3]
            if rowptr.numel() != n * rows_pad + 1 or colind.numel() < off:
                raise ValueError("output operator buffers do not fit this batch")
        with torch.cuda.device(dev):
            N.call("sn_assemble_block_diag", _ptr(table), n, rows_pad, cols_pad, off, vpe, _ptr(rowptr), _ptr(colind),
                   _ptr(val), _stream())
        if out is not None:
            if kind == "csr":
                out._nnz = off
            else:
                out._n_blocks = off
                out.max_row_blocks = mrb
            return out
        if kind == "csr":
            return CsrOperator(rowptr, colind[:off] if off else colind, val[:off] if off else val, n * rows_pad,
                               n * cols_pad, None, off)
        return Bsr4Operator(rowptr, colind[:off] if off else colind, val[:16 * off] if off else val, n * rows_pad,
                            n * cols_pad, None, off, mrb)

    def assemble(self, keys, kind, rows_pad, cols_pad, out=None, plan=None):
        """Block-diagonal batch operator of the meshes ``keys`` (in order), each padded to ``rows_pad x cols_pad``
        (block rows / columns for "bsr4", scalar for "csr"), with its transpose attached.  ``out``: an operator of
        the same batch shape whose buffers (and whose transpose's) are overwritten in place.  ``plan``: the result of
        ``plan(keys, kind, rows_pad, cols_pad)`` prepared earlier (``keys`` is then ignored)."""
        if plan is None:
            plan = self.plan(keys, kind, rows_pad, cols_pad)
        elif (plan["kind"], plan["rows_pad"], plan["cols_pad"]) != (kind, rows_pad, cols_pad):
            raise ValueError("assembly plan was made for another operator kind / padded size")
        n = plan["n"]
        fwd = self._assemble_one(plan["fwd"], n, kind, rows_pad, cols_pad, out)
        bwd = self._assemble_one(plan["bwd"], n, kind, cols_pad, rows_pad, None if out is None else out.T)
        fwd._T, bwd._T = bwd, fwd
        return fwd
