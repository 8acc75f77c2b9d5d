"""Callers of the hot path: the model stacks that define the shapes the
kernels see.  Same constructor arguments, ``forward`` signatures and ``state_dict`` keys
(``conv1.fc.*``, ``rn{i}.bn_fc{0,1}.{bn,fc}.*``, ``conv2.{bn,fc}.*``) as the reference files:

    ArapLapModel / ArapDirModel / ArapAvgModel / ArapMlpModel
        src/as_rigid_as_possible/models.py:21-52 (Model), :108-152 (DirModel), :54-78, :80-105
    LapEncoder                      src/mesh_mnist/models_vae.py:22-51   (5 x LapResNet2(128), cfg2)
    DirDeepModel                    src/normal_predict/models.py:234-280 (30-block Dirac stack)
    DcLapModel / DcDirModel / SiameseModel
        src/dense_correspondence/models.py:21-48 (Model), :140-182 (DirModel), :184-203 (SiameseModel: the
        FA . FB^T correlation of BASELINE cfg5)
    LapResNet2General               src/normal_predict/models.py:447-477 (_LapResNet2: inner_layers, in != out)

Everything here is a thin loop over ``utils_pt`` blocks; the arithmetic lives in libsurfnet_b200.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused
from . import ops
from . import utils_pt as utils
from .operators import as_bsr4, as_csr

__all__ = ["ArapLapModel", "ArapDirModel", "ArapAvgModel", "ArapMlpModel", "LapEncoder", "DirDeepModel", "LapResNet2General", "DcLapModel",
           "DcDirModel", "SiameseModel", "arap_loss"]


def _add_blocks(model, kinds, width):
    for i, kind in enumerate(kinds):
        model.add_module("rn{}".format(i), kind(width))


def _last3_tiled(inputs, times):
    return inputs[:, :, -3:].repeat(1, 1, times)


def _add_last3_tiled(y, inputs, times):
    """``y + inputs[:, :, -3:].repeat(1, 1, times)`` (reference as_rigid_as_possible/models.py:152): one fused pass."""
    return fused.head_add_tiled(y, inputs, times)


def _num_faces(Di, DiA, batch_size):
    if isinstance(DiA, torch.Tensor):
        return DiA.size(2) // 4 if DiA.dim() == 3 else DiA.size(1) // 4 // batch_size
    return as_bsr4(DiA).n_bcols // batch_size


def _dirac_stack(model, n_layers, D, DA, mask, v, f):
    """The reference's block loop ``v, f = rn_i(Di, DiA, v, f)`` / ``v = rn_i(L, mask, v)`` (as_rigid_as_possible/models.py:
    142-146 and its twins).  Returns v; the final f is discarded by every caller, as in the reference.  When the Dirac
    blocks run their fused training path the face features are handed from block to block in ACTIVATED form (the only
    form anybody reads): utils_pt.DirResNet2.forward_chained."""
    blocks = [model._modules["rn{}".format(i)] for i in range(n_layers)]
    dirac = [i for i, b in enumerate(blocks) if isinstance(b, utils.DirResNet2)]
    chained = bool(dirac) and all(blocks[i].chain_supported(v, f) for i in dirac)
    state = ops.face_chain_start(f.reshape(-1, f.size(2))) if chained else None
    for i, blk in enumerate(blocks):
        if isinstance(blk, utils.DirResNet2):
            if chained:
                v, state = blk.forward_chained(D, DA, v, state, last=(i == dirac[-1]))
            else:
                v, f = blk(D, DA, v, f)
        else:
            v = blk(None, mask, v)
    return v


class ArapLapModel(nn.Module):
    """as_rigid_as_possible ``Model(layer, dense)``: conv1(6->128), alternating Lap / Avg blocks, conv2(128->120)."""

    def __init__(self, layer, dense=False):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(6, 128, batch_norm=None)
        self.layer = layer
        lap = utils.DenseLapResNet2 if dense else utils.LapResNet2
        _add_blocks(self, [lap if i % 2 == 0 else utils.AvgResNet2 for i in range(layer)], 128)
        self.conv2 = utils.GraphConv1x1(128, 120, batch_norm="pre")

    def forward(self, L, mask, inputs):
        x = self.conv1(inputs)
        for i in range(self.layer):
            x = self._modules["rn{}".format(i)](L, mask, x)
        return _add_last3_tiled(self.conv2.forward_elu(x), inputs, 40)


class ArapAvgModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(6, 128, batch_norm=None)
        _add_blocks(self, [utils.AvgResNet2] * 15, 128)
        self.conv2 = utils.GraphConv1x1(128, 120, batch_norm="pre")

    def forward(self, L, mask, inputs):
        x = self.conv1(inputs)
        for i in range(15):
            x = self._modules["rn{}".format(i)](L, mask, x)
        return _add_last3_tiled(self.conv2.forward_elu(x), inputs, 40)


class ArapMlpModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(6, 128, batch_norm=None)
        _add_blocks(self, [utils.MlpResNet2] * 15, 128)
        self.bn = utils.GraphBatchNorm(128)
        self.conv2 = utils.GraphConv1x1(128, 120, batch_norm=None)

    def forward(self, L, mask, inputs):
        x = self.conv1(inputs)
        for i in range(15):
            x = self._modules["rn{}".format(i)](L, mask, x)
        return self.conv2(F.elu(self.bn(x))) + _last3_tiled(inputs, 40)


class ArapDirModel(nn.Module):
    """as_rigid_as_possible ``DirModel``: 8 DirResNet2 + 7 AvgResNet2 blocks at width 128 (BASELINE cfg3/cfg4)."""

    def __init__(self):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(6, 128, batch_norm=None)
        _add_blocks(self, [utils.DirResNet2 if i % 2 == 0 else utils.AvgResNet2 for i in range(15)], 128)
        self.do = nn.Dropout2d()  # declared (and unused) by the reference; kept for module-tree parity
        self.conv2 = utils.GraphConv1x1(128, 120, batch_norm="pre")

    def forward(self, Di, DiA, mask, inputs):
        batch_size = inputs.size(0)
        D, DA = as_bsr4(Di), as_bsr4(DiA)
        v = self.conv1(inputs)
        f = v.new_zeros(batch_size, _num_faces(Di, DiA, batch_size), 128)
        v = _dirac_stack(self, 15, D, DA, mask, v, f)
        return _add_last3_tiled(self.conv2.forward_elu(v), inputs, 40)


def arap_loss(outputs, targets, mask, batch_size):
    """Masked smooth-L1, summed, per mesh (src/as_rigid_as_possible/main.py:225-226)."""
    loss = fused.masked_smooth_l1(outputs, targets, mask, 1.0 / batch_size) if outputs.is_cuda else None
    if loss is not None:
        return loss
    return F.smooth_l1_loss(outputs * mask.expand_as(outputs), targets, reduction="sum") / batch_size


class LapEncoder(nn.Module):
    """mesh_mnist VAE encoder: conv1(3->128), 5 x LapResNet2(128), BN-conv, masked mean, two heads."""

    def __init__(self):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(3, 128, batch_norm=None)
        self.num_layers = 5
        _add_blocks(self, [utils.LapResNet2] * 5, 128)
        self.bn_conv2 = utils.GraphConv1x1(128, 128, batch_norm="pre")
        self.fc_mu = nn.Linear(128, 100)
        self.fc_logvar = nn.Linear(128, 100)

    def forward(self, inputs, L, mask):
        L = as_csr(L) if not (isinstance(L, torch.Tensor) and L.layout is torch.strided) else L
        x = self.conv1(inputs)
        for i in range(self.num_layers):
            x = self._modules["rn{}".format(i)](L, mask, x)
        x = F.elu(self.bn_conv2.forward_elu(x))
        x = utils.global_average(x, mask).squeeze()
        return self.fc_mu(x), self.fc_logvar(x)


class LapResNet2General(nn.Module):
    """normal_predict ``_LapResNet2``: ``inner_layers`` stages, num_inputs may differ from num_outputs."""

    def __init__(self, num_inputs, num_outputs=None, bnmode="", inner_layers=2):
        super().__init__()
        num_outputs = num_inputs if num_outputs is None else num_outputs
        self.num_outputs = num_outputs
        if bnmode is not None:
            bnmode = bnmode + "pre"
        self.layer = inner_layers
        widths = [num_inputs] + [num_outputs] * (inner_layers - 1)
        for i, w in enumerate(widths):
            self.add_module("bn_fc{}".format(i), utils.GraphConv1x1(2 * w, num_outputs, batch_norm=bnmode))

    def forward(self, L, mask, inputs):
        batch, node, _ = inputs.size()
        dense = isinstance(L, torch.Tensor) and L.layout is torch.strided
        x = inputs
        for i in range(self.layer):
            conv = self._modules["bn_fc{}".format(i)]
            if dense:
                x = F.elu(x)
                x = conv(torch.cat([x, torch.bmm(L, x)], 2))
            else:
                z = ops.stage_concat(as_csr(L), x.reshape(batch * node, x.size(2)))
                x = conv.forward_rows(z).view(batch, node, self.num_outputs)
        if self.num_outputs <= inputs.size(2):
            return x + inputs[:, :, :self.num_outputs]
        return x + torch.cat([inputs] * 2, dim=2)


class DirDeepModel(nn.Module):
    """normal_predict ``DirDeepModel`` (models.py:234-280): ``forward((Di, DiA), mask, inputs)``."""

    def __init__(self, in_features=3, out_features=1, layers=30):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(in_features, 128, batch_norm=None)
        self.feature_width = 128
        self.layer_num = layers
        _add_blocks(self, [utils.DirResNet2 if i % 2 == 0 else utils.AvgResNet2 for i in range(layers)], 128)
        self.do = nn.Dropout2d()
        self.conv2 = utils.GraphConv1x1(128, out_features, batch_norm="pre")

    def forward(self, DiDA, mask, inputs):
        Di, DiA = DiDA
        batch_size = inputs.size(0)
        D, DA = as_bsr4(Di), as_bsr4(DiA)
        v = self.conv1(inputs)
        f = v.new_zeros(batch_size, DA.n_bcols // batch_size, self.feature_width)
        v = _dirac_stack(self, self.layer_num, D, DA, mask, v, f)
        return F.elu(self.conv2(v))

    def fuzzy_load(self, pre_dict):
        own = self.state_dict()
        own.update({k: v for k, v in pre_dict.items() if k in own})
        self.load_state_dict(own)


class DcLapModel(nn.Module):
    """dense_correspondence ``Model(layer)`` (models.py:21-48): conv1(3->128), Lap / Avg blocks, conv2(128->120)."""

    def __init__(self, layer):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(3, 128, batch_norm=None)
        self.layer = layer
        _add_blocks(self, [utils.LapResNet2 if i % 2 == 0 else utils.AvgResNet2 for i in range(layer)], 128)
        self.conv2 = utils.GraphConv1x1(128, 120, batch_norm="pre")

    def forward(self, L, mask, inputs):
        Lop = L if isinstance(L, torch.Tensor) and L.layout == torch.strided else as_csr(L)
        x = self.conv1(inputs)
        for i in range(self.layer):
            x = self._modules["rn{}".format(i)](Lop, mask, x)
        return _add_last3_tiled(self.conv2.forward_elu(x), inputs, 40)


class DcDirModel(nn.Module):
    """dense_correspondence ``DirModel(layer)`` (models.py:140-182).  The reference reads the face count from
    ``DiA.size(2)`` (:166), which only exists for 3-D operators whose DirResNet2 branch is dead;
    here 2-D block-diagonal and 3-D operators both work."""

    def __init__(self, layer):
        super().__init__()
        self.conv1 = utils.GraphConv1x1(3, 128, batch_norm=None)
        self.layer = layer
        _add_blocks(self, [utils.DirResNet2 if i % 2 == 0 else utils.AvgResNet2 for i in range(layer)], 128)
        self.do = nn.Dropout2d()  # declared and unused by the reference
        self.conv2 = utils.GraphConv1x1(128, 120, batch_norm="pre")

    def forward(self, Di, DiA, mask, inputs):
        batch_size = inputs.size(0)
        D, DA = as_bsr4(Di), as_bsr4(DiA)
        v = self.conv1(inputs)
        f = v.new_zeros(batch_size, DA.n_bcols // batch_size, 128)
        v = _dirac_stack(self, self.layer, D, DA, mask, v, f)
        return _add_last3_tiled(self.conv2.forward_elu(v), inputs, 40)


class SiameseModel(nn.Module):
    """dense_correspondence ``SiameseModel(model, layer)`` (models.py:184-203): one shared tower applied to both shapes,
    then the all-pairs feature correlation ``FA @ FB^T`` ([B, Na, 120] x [B, 120, Nb] -> [B, Na, Nb]).

    The towers run on the libsurfnet_b200 kernels.  The correlation's cost is writing the Na x Nb result (196 MB per
    7000-vertex pair): it runs on the tcgen05 3xTF32 kernel with the TMA-store epilogue (``fused.correlation``,
    sn_gemm_nt_wide_tf32_f32); its backward -- two contractions over the vertices -- on torch.bmm."""

    def __init__(self, model="dirac", layer=15):
        super().__init__()
        if "dir" in model:
            self.model = DcDirModel(layer)
        elif "lap" in model:
            self.model = DcLapModel(layer)
        else:
            raise ValueError("SiameseModel: supported towers are 'dirac' and 'lap', got %r" % (model,))

    def forward(self, OperationA, OperationB, inputA, inputB):
        FA = self.model(*OperationA, inputA)
        FB = self.model(*OperationB, inputB)
        return fused.correlation(FA, FB)
