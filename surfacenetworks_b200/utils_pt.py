"""Layer library -- the torch.nn.Module surface of reference ``src/utils/utils_pt.py``, backed by
libsurfnet_b200's sm_100a kernels.

Same class names, constructor arguments, ``forward`` signatures, return values and ``state_dict`` keys
(``bn_fc{0,1}.bn.*``, ``bn_fc{0,1}.fc.*``) as the reference, so its model files
(``as_rigid_as_possible/models.py`` etc.) and checkpoints work unchanged:

    reference (utils_pt.py)                 here
    --------------------------------------  -----------------------------------------------------------
    LapResNet2.forward       :159-180       ops.stage_concat(CsrOperator)  -> sn_elu_f32 + sn_csr_spmm_f32
    DirResNet2.forward       :191-220       ops.stage_concat(Bsr4Operator) -> sn_elu_f32 + sn_bsr4_spmm_f32
    GraphConv1x1.forward     :91-104        BatchNorm over the flattened [B*N, C] rows (no transposes) + Linear
    sparse_cat / sparse_diag_cat :21-53     vectorised host assembly (same coalesced COO result)
    DenseLapResNet2 / dense L  :132-148     torch.bmm (plain library GEMM; small meshes only)

Operators may be given as torch sparse COO tensors (2-D block-diagonal from ``sparse_diag_cat`` or 3-D from
``sparse_cat`` -- the reference's 3-D branch is dead code, utils_pt.py:197-199 raises NameError), as dense
``[B, V, V]`` tensors (Laplacian only), or pre-converted ``CsrOperator`` / ``Bsr4Operator`` objects.

CUDA only: CPU tensors raise (no fallback path; use the reference for CPU runs).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused, ops
from .operators import Bsr4Operator, CsrOperator, as_bsr4, as_csr

__all__ = [
    "sparse_cat", "sparse_diag_cat", "sp_sparse_to_pt_sparse", "to_dense_batched", "GraphConv1x1",
    "GraphBatchNorm", "global_average", "DenseLapResNet2", "LapResNet2", "DirResNet2", "AvgResNet2", "MlpResNet2",
]


# ------------------------------------------------------------------------------------------------ batching helpers
def _gather_coo(tensors):
    idx = [t._indices() for t in tensors]
    val = torch.cat([t._values() for t in tensors], 0)
    counts = torch.tensor([i.shape[1] for i in idx], dtype=torch.long)
    which = torch.repeat_interleave(torch.arange(len(tensors), dtype=torch.long), counts).to(val.device)
    return torch.cat(idx, 1), val, which


def _already_coalesced(which, idx, size0, size1):
    """True when the concatenated entries are already in coalesced order (strictly increasing (mesh, row, col) keys,
    every row / column inside its block): the usual case -- scipy CSR operators converted by sp_sparse_to_pt_sparse come
    out row-major with sorted columns -- and then ``.coalesce()`` (a 9M-entry sort per ARAP batch, most of the 1.6 s the
    reference spends in sparse_diag_cat per step) would return exactly the same tensor."""
    if idx.shape[1] == 0:
        return True
    if int(idx.min()) < 0 or int(idx[0].max()) >= size0 or int(idx[1].max()) >= size1:
        return False
    key = (which * size0 + idx[0]) * size1 + idx[1]
    return bool((key[1:] > key[:-1]).all())


def sparse_cat(tensors, size0, size1):
    """List of 2-D COO operators -> one coalesced 3-D COO ``[B, size0, size1]`` (utils_pt.py:21-39)."""
    idx, val, which = _gather_coo(tensors)
    idx3 = torch.cat([which.unsqueeze(0), idx], 0)
    shape = (len(tensors), size0, size1)
    if _already_coalesced(which, idx, size0, size1):
        return torch.sparse_coo_tensor(idx3, val, shape, is_coalesced=True)
    return torch.sparse_coo_tensor(idx3, val, shape).coalesce()


def sparse_diag_cat(tensors, size0, size1):
    """List of 2-D COO operators -> coalesced block-diagonal ``[B*size0, B*size1]`` (utils_pt.py:41-53)."""
    idx, val, which = _gather_coo(tensors)
    shift = torch.stack([which * size0, which * size1], 0)
    shape = (len(tensors) * size0, len(tensors) * size1)
    if _already_coalesced(which, idx, size0, size1):
        return torch.sparse_coo_tensor(idx + shift, val, shape, is_coalesced=True)
    return torch.sparse_coo_tensor(idx + shift, val, shape).coalesce()


def sp_sparse_to_pt_sparse(L):
    """scipy sparse matrix -> (uncoalesced) torch COO with the matrix's dtype (utils_pt.py:56-69)."""
    L = L.tocoo()
    idx = torch.from_numpy(np.stack([L.row, L.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(L.data), L.shape)


def to_dense_batched(x, batch_size):
    """utils_pt.py:71-74."""
    return x.to_dense().unsqueeze(0).repeat(batch_size, 1, 1)


# ------------------------------------------------------------------------------------------------ dense stage
class GraphConv1x1(nn.Module):
    """Per-node Linear with optional BatchNorm before ("pre") or after ("post") -- utils_pt.py:76-104.

    The reference normalises ``x.transpose(1, 2)`` ([B, C, N]) with BatchNorm1d, i.e. per-channel statistics
    over all B*N rows (padded rows included); the same statistics come from BatchNorm1d on the flattened
    [B*N, C] matrix, which skips both transposes.
    """

    def __init__(self, num_inputs, num_outputs, batch_norm=None):
        super().__init__()
        self.num_inputs, self.num_outputs, self.batch_norm = num_inputs, num_outputs, batch_norm
        if batch_norm == "pre":
            self.bn = nn.BatchNorm1d(num_inputs)
        if batch_norm == "post":
            self.bn = nn.BatchNorm1d(num_outputs)
        self.fc = nn.Linear(num_inputs, num_outputs)

    def forward_rows(self, z, residual=None, res_cell=None):
        """z: [rows, num_inputs] -> [rows, num_outputs] (+ residual).  "pre" BatchNorm + Linear runs as the fused
        stage of ``fused.py`` (statistics pass, BN folded into the weights, tcgen05 GEMM with the residual in its
        epilogue) whenever the widths allow it."""
        if self.batch_norm == "pre":
            return fused.bn_linear(z, self.bn, self.fc, residual, res_cell)
        if fused.smallk_linear_supported(z, self.fc):          # the 3 / 6-channel input layers (conv1)
            z = fused.smallk_linear(z, self.fc)
        else:
            z = self.fc(z)
        if self.batch_norm == "post":
            z = self.bn(z)
        return z if residual is None else z + residual

    def forward(self, x):
        batch_size, num_nodes, num_inputs = x.size()
        assert num_inputs == self.num_inputs
        return self.forward_rows(x.reshape(-1, num_inputs)).view(batch_size, num_nodes, self.num_outputs)

    def forward_elu(self, x):
        """``self(F.elu(x))`` -- how every model stack of the reference ends (as_rigid_as_possible/models.py:121,151) -- with
        the activation inside the fused stage: one pass for elu + BatchNorm statistics, elu' in the dZ GEMM epilogue."""
        if self.batch_norm != "pre" or not x.is_cuda:
            return self(F.elu(x))
        batch_size, num_nodes, num_inputs = x.size()
        assert num_inputs == self.num_inputs
        return fused.elu_bn_linear(x.reshape(-1, num_inputs), self.bn, self.fc).view(batch_size, num_nodes, self.num_outputs)


class GraphBatchNorm(nn.Module):
    """BatchNorm over [B*N, C] that always uses batch statistics (utils_pt.py:107-118)."""

    def __init__(self, num_inputs):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_inputs)

    def forward(self, x):
        self.bn.train()
        b, n, c = x.size()
        return self.bn(x.reshape(b * n, c)).view(b, n, c)


def global_average(x, mask):
    """Masked mean over the node axis, keepdim (utils_pt.py:120-122)."""
    m = mask.expand_as(x)
    return (x * m).sum(1, keepdim=True) / m.sum(1, keepdim=True)


# ------------------------------------------------------------------------------------------------ ResNet blocks
class _TwoStageBlock(nn.Module):
    """Shared shape of the reference's ResNet blocks: two GraphConv1x1(2C -> C, "pre") stages."""

    def __init__(self, num_outputs):
        super().__init__()
        self.num_outputs = num_outputs
        self.bn_fc0 = GraphConv1x1(2 * num_outputs, num_outputs, batch_norm="pre")
        self.bn_fc1 = GraphConv1x1(2 * num_outputs, num_outputs, batch_norm="pre")


class LapResNet2(_TwoStageBlock):
    """x -> x + fc1(BN[e1 | L e1]),  e1 = elu(fc0(BN[e0 | L e0])),  e0 = elu(x)   (utils_pt.py:151-180).

    ``L``: torch sparse COO (block-diagonal 2-D or batched 3-D), ``CsrOperator``, or dense ``[B, V, V]``.
    ``mask`` is unused, as in the reference.
    """

    def forward(self, L, mask, inputs):
        batch, node, feat = inputs.size()
        if isinstance(L, torch.Tensor) and L.layout is torch.strided:
            return _dense_lap_block(self, L, inputs)
        op = as_csr(L)
        x = inputs.reshape(batch * node, feat)
        # training: the residual's gradient skips autograd's accumulation add -- the second stage leaves it in `cell`,
        # the first stage's backward SpMM adds it in its store path (ops.stage_concat / fused.bn_linear)
        cell = None
        if torch.is_grad_enabled() and x.requires_grad and fused.bn_linear_is_fused(x, self.bn_fc1.fc, feat):
            cell = {}
        y = self.bn_fc0.forward_rows(ops.stage_concat(op, x, in_cell=cell))
        y = self.bn_fc1.forward_rows(ops.stage_concat(op, y), residual=x, res_cell=cell)   # "+ inputs" rides in the epilogue
        return y.view(batch, node, feat)


def _dense_lap_block(block, L, inputs):
    x = F.elu(inputs)
    x = block.bn_fc0(torch.cat([x, torch.bmm(L, x)], 2))
    x = F.elu(x)
    x = block.bn_fc1(torch.cat([x, torch.bmm(L, x)], 2))
    return x + inputs


class DenseLapResNet2(_TwoStageBlock):
    """LapResNet2 with a dense ``[B, V, V]`` Laplacian through torch.bmm (utils_pt.py:124-148)."""

    def forward(self, L, mask, inputs):
        return _dense_lap_block(self, L, inputs)


class DirResNet2(_TwoStageBlock):
    """Dirac block (utils_pt.py:182-220):

        f_out = fc0(BN[elu(f) | D  elu(v)])            faces    <- vertices
        v_out = fc1(BN[elu(v) | D* elu(f_out)])        vertices <- faces
        return v + v_out, f_out

    ``Di`` [4F x 4V] / ``DiA`` [4V x 4F]: torch sparse COO (2-D block-diagonal or 3-D) or ``Bsr4Operator``.
    """

    def __init__(self, num_outputs, res_f=False):
        super().__init__(num_outputs)
        self.res_f = res_f

    def forward(self, Di, DiA, v, f):
        batch_size, num_nodes, num_inputs = v.size()
        _, num_faces, _ = f.size()
        D, DA = as_bsr4(Di), as_bsr4(DiA)
        v2 = v.reshape(batch_size * num_nodes, num_inputs)
        f2 = f.reshape(batch_size * num_faces, num_inputs)
        if ops.dir_block_supported(v2, f2, self.bn_fc0, self.bn_fc1):       # training: the whole block is one node
            v_new, f_out = ops.dir_block(D, DA, v2, f2, self.bn_fc0, self.bn_fc1)
            return v_new.view(batch_size, num_nodes, num_inputs), f_out.view(batch_size, num_faces, num_inputs)
        f_out = self.bn_fc0.forward_rows(ops.stage_concat(D, f2, v2))
        v_new = self.bn_fc1.forward_rows(ops.stage_concat(DA, v2, f_out), residual=v2)   # v + v_out in the epilogue
        return v_new.view(batch_size, num_nodes, num_inputs), f_out.view(batch_size, num_faces, num_inputs)


    # ---- chained form (stacks where the face features only travel from one Dirac block to the next, see ops._DirBlockChained)
    def chain_supported(self, v, f):
        """True when ``forward_chained`` applies: training-mode fused block at a width both epilogues cover."""
        if v.dim() != 3 or f.dim() != 3:
            return False
        v2 = v.reshape(-1, v.size(2))
        f2 = f.reshape(-1, f.size(2))
        return ops.dir_block_supported(v2, f2, self.bn_fc0, self.bn_fc1)

    def forward_chained(self, Di, DiA, v, face_state, last=False):
        """``forward`` with the face features handed over in activated form: ``face_state`` = ``ops.face_chain_start(f)``
        or the state returned by the previous Dirac block.  Returns (v_out, next_face_state)."""
        batch_size, num_nodes, num_inputs = v.size()
        D, DA = as_bsr4(Di), as_bsr4(DiA)
        v2 = v.reshape(batch_size * num_nodes, num_inputs)
        Zf, stf = face_state
        v_new, Zf_next, st_next = ops.dir_block_chained(D, DA, v2, Zf, stf, self.bn_fc0, self.bn_fc1, last)
        return v_new.view(batch_size, num_nodes, num_inputs), (Zf_next, st_next)


class AvgResNet2(_TwoStageBlock):
    """Global-average block, no sparse operator (utils_pt.py:222-243):
    x -> x + fc1(BN[e1 | avg(e1)]),  e1 = elu(fc0(BN[e0 | avg(e0)])),  e0 = elu(x), avg = masked mean over the mesh.

    At the widths the tensor-core stage covers, each stage runs as ``fused.avg_stage``: the broadcast half of the
    concat is never materialised (K = C GEMM + per-mesh bias).  Other widths take the torch composite below.
    """

    def forward(self, L, mask, inputs):
        B, V, C = inputs.size()
        # the residual's gradient travels through this cell into the first stage's ELU-backward kernel (fused.py);
        # only when autograd will really run both stages' backward in one pass over the same `inputs`
        cell = {} if (torch.is_grad_enabled() and inputs.requires_grad) else None
        y = fused.avg_stage(inputs, mask, self.bn_fc0.bn, self.bn_fc0.fc, in_cell=cell) if inputs.is_cuda else None
        if y is not None:
            y = fused.avg_stage(y.view(B, V, -1), mask, self.bn_fc1.bn, self.bn_fc1.fc, residual=inputs, res_cell=cell)
        if y is not None:
            return y.view(B, V, -1)
        x = F.elu(inputs)
        x = self.bn_fc0(torch.cat([x, global_average(x, mask).expand_as(x)], 2))
        x = F.elu(x)
        x = self.bn_fc1(torch.cat([x, global_average(x, mask).expand_as(x)], 2))
        return x + inputs


class MlpResNet2(nn.Module):
    """BN -> ELU -> Linear, twice, plus residual (utils_pt.py:245-263)."""

    def __init__(self, num_outputs):
        super().__init__()
        self.num_outputs = num_outputs
        self.bn0 = GraphBatchNorm(num_outputs)
        self.fc0 = GraphConv1x1(num_outputs, num_outputs, batch_norm=None)
        self.bn1 = GraphBatchNorm(num_outputs)
        self.fc1 = GraphConv1x1(num_outputs, num_outputs, batch_norm=None)

    def forward(self, L, mask, inputs):
        x = self.fc0(F.elu(self.bn0(inputs)))
        x = self.fc1(F.elu(self.bn1(x)))
        return x + inputs
