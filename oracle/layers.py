"""Functional torch-CPU restatement of the reference layer library.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Each function follows the reference lines it cites; parameters come in a flat dict ``P`` keyed exactly like the
reference modules' ``state_dict()`` (``bn_fc0.bn.weight``, ``bn_fc0.fc.bias``, ...), so a state_dict saved by
either implementation drives it.  The sparse products go through ``torch.mm(sparse_coo, dense)`` on the CPU --
the very call the reference makes (src/utils/utils_pt.py:167,176,202,214); that arithmetic is PyTorch ATen's
(third-party, reference pin torch==1.0.0, README.md:45; this image: see torch.__version__), and
``oracle/sn_oracle.c`` restates it in plain C for an independent check.

Pinned by tests/test_oracle_golden.py against tests/golden/*.npz, which were produced by the reference's own
modules (tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def _sub(P, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in P.items() if k.startswith(prefix)}


def batch_norm_rows(x, P, training=True, momentum=0.1, eps=1e-5):
    """nn.BatchNorm1d applied the reference's way: on x.transpose(1, 2) = [B, C, N]  (utils_pt.py:98,101)."""
    y = F.batch_norm(x.transpose(1, 2), P.get("running_mean"), P.get("running_var"), P["weight"], P["bias"],
                     training, momentum, eps)
    if training and "num_batches_tracked" in P:
        P["num_batches_tracked"] += 1
    return y.transpose(1, 2)


def graph_conv1x1(x, P, batch_norm=None, training=True):
    """GraphConv1x1.forward, utils_pt.py:91-104.  x: [B, N, Cin]."""
    if batch_norm == "pre":
        x = batch_norm_rows(x, _sub(P, "bn."), training)
    x = F.linear(x, P["fc.weight"], P["fc.bias"])
    if batch_norm == "post":
        x = batch_norm_rows(x, _sub(P, "bn."), training)
    return x


def graph_batch_norm(x, P):
    """GraphBatchNorm.forward, utils_pt.py:112-118: always batch statistics, on the flattened rows."""
    b, n, c = x.shape
    Pb = _sub(P, "bn.")
    y = F.batch_norm(x.reshape(b * n, c), Pb.get("running_mean"), Pb.get("running_var"), Pb["weight"], Pb["bias"],
                     True, 0.1, 1e-5)
    if "num_batches_tracked" in Pb:
        Pb["num_batches_tracked"] += 1
    return y.view(b, n, c)


def global_average(x, mask):
    """utils_pt.py:120-122."""
    m = mask.expand_as(x)
    return (x * m).sum(1, keepdim=True) / m.sum(1, keepdim=True)


def apply_laplacian(L, x):
    """utils_pt.py:164-167: dense -> bmm, sparse block-diagonal -> mm on the flattened batch."""
    b, n, c = x.shape
    if L.layout is torch.strided:
        return torch.bmm(L, x)
    return torch.mm(L, x.reshape(-1, c)).view(b, n, c)


def lap_resnet2(P, L, x, training=True):
    """LapResNet2.forward / DenseLapResNet2.forward, utils_pt.py:159-180 / 132-148."""
    h = x
    for stage in ("bn_fc0.", "bn_fc1."):
        h = F.elu(h)
        h = torch.cat([h, apply_laplacian(L, h)], 2)
        h = graph_conv1x1(h, _sub(P, stage), "pre", training)
    return h + x


def apply_dirac(D, x, rows_out):
    """The quaternion view, utils_pt.py:201-203 / 213-215: [B, n, C] -> view [B*n*4, C/4] -> mm -> [B, rows_out, C]."""
    b, n, c = x.shape
    return torch.mm(D, x.reshape(b * n * 4, c // 4)).view(b, rows_out, c)


def dir_resnet2(P, Di, DiA, v, f, training=True):
    """DirResNet2.forward, utils_pt.py:191-220 (2-D block-diagonal operators)."""
    nv, nf = v.shape[1], f.shape[1]
    x_in, f_in = F.elu(v), F.elu(f)
    f_out = graph_conv1x1(torch.cat([f_in, apply_dirac(Di, x_in, nf)], 2), _sub(P, "bn_fc0."), "pre", training)
    v_out = graph_conv1x1(torch.cat([x_in, apply_dirac(DiA, F.elu(f_out), nv)], 2), _sub(P, "bn_fc1."), "pre",
                          training)
    return v + v_out, f_out


def avg_resnet2(P, mask, x, training=True):
    """AvgResNet2.forward, utils_pt.py:230-243."""
    h = x
    for stage in ("bn_fc0.", "bn_fc1."):
        h = F.elu(h)
        h = torch.cat([h, global_average(h, mask).expand_as(h).contiguous()], 2)
        h = graph_conv1x1(h, _sub(P, stage), "pre", training)
    return h + x


def mlp_resnet2(P, x):
    """MlpResNet2.forward, utils_pt.py:255-263."""
    h = graph_conv1x1(F.elu(graph_batch_norm(x, _sub(P, "bn0."))), _sub(P, "fc0."))
    h = graph_conv1x1(F.elu(graph_batch_norm(h, _sub(P, "bn1."))), _sub(P, "fc1."))
    return h + x


def sparse_diag_cat(tensors, size0, size1):
    """utils_pt.py:41-53, restated: shift each operator's indices by i*(size0, size1), concatenate, coalesce."""
    idx, val = [], []
    for i, t in enumerate(tensors):
        idx.append(t._indices() + torch.tensor([[i * size0], [i * size1]]))
        val.append(t._values())
    return torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val), (len(tensors) * size0, len(tensors) * size1)).coalesce()


def sparse_cat(tensors, size0, size1):
    """utils_pt.py:21-39, restated: prepend the batch index, concatenate, coalesce -> [B, size0, size1]."""
    idx, val = [], []
    for i, t in enumerate(tensors):
        ii = t._indices()
        idx.append(torch.cat([torch.full((1, ii.shape[1]), i, dtype=torch.long), ii], 0))
        val.append(t._values())
    return torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val), (len(tensors), size0, size1)).coalesce()


def arap_dir_model(P, Di, DiA, mask, inputs, training=True):
    """as_rigid_as_possible DirModel.forward, src/as_rigid_as_possible/models.py:128-152."""
    b = inputs.shape[0]
    v = graph_conv1x1(inputs, _sub(P, "conv1."))
    f = torch.zeros(b, DiA.shape[1] // 4 // b, 128)
    for i in range(15):
        Pi = _sub(P, "rn%d." % i)
        if i % 2 == 0:
            v, f = dir_resnet2(Pi, Di, DiA, v, f, training)
        else:
            v = avg_resnet2(Pi, mask, v, training)
    x = graph_conv1x1(F.elu(v), _sub(P, "conv2."), "pre", training)
    return x + inputs[:, :, -3:].repeat(1, 1, 40)


def arap_lap_model(P, L, mask, inputs, layers=15, training=True):
    """as_rigid_as_possible Model.forward, src/as_rigid_as_possible/models.py:41-52."""
    x = graph_conv1x1(inputs, _sub(P, "conv1."))
    for i in range(layers):
        Pi = _sub(P, "rn%d." % i)
        x = lap_resnet2(Pi, L, x, training) if i % 2 == 0 else avg_resnet2(Pi, mask, x, training)
    x = graph_conv1x1(F.elu(x), _sub(P, "conv2."), "pre", training)
    return x + inputs[:, :, -3:].repeat(1, 1, 40)


def arap_loss(outputs, targets, mask, batch_size):
    """src/as_rigid_as_possible/main.py:225-226."""
    return F.smooth_l1_loss(outputs * mask.expand_as(outputs), targets, reduction="sum") / batch_size


def lap_resnet2_general(P, L, x, inner_layers=2, training=True):
    """normal_predict _LapResNet2.forward, src/normal_predict/models.py:462-477 (bnmode='' -> "pre")."""
    h = x
    for i in range(inner_layers):
        h = F.elu(h)
        h = torch.cat([h, apply_laplacian(L, h)], 2)
        h = graph_conv1x1(h, _sub(P, "bn_fc%d." % i), "pre", training)
    n_out = h.shape[2]
    if n_out <= x.shape[2]:
        return h + x[:, :, :n_out]
    return h + torch.cat([x] * 2, dim=2)


def dir_deep_model(P, Di, DiA, mask, inputs, layers, training=True):
    """normal_predict DirDeepModel.forward, src/normal_predict/models.py:255-274."""
    b = inputs.shape[0]
    v = graph_conv1x1(inputs, _sub(P, "conv1."))
    f = v.new_zeros(b, DiA.shape[-1] // 4 // b, 128)
    for i in range(layers):
        Pi = _sub(P, "rn%d." % i)
        if i % 2 == 0:
            v, f = dir_resnet2(Pi, Di, DiA, v, f, training)
        else:
            v = avg_resnet2(Pi, mask, v, training)
    return F.elu(graph_conv1x1(v, _sub(P, "conv2."), "pre", training))


def lap_encoder(P, inputs, L, mask, training=True):
    """mesh_mnist LapEncoder.forward, src/mesh_mnist/models_vae.py:38-51."""
    x = graph_conv1x1(inputs, _sub(P, "conv1."))
    for i in range(5):
        x = lap_resnet2(_sub(P, "rn%d." % i), L, x, training)
    x = F.elu(graph_conv1x1(F.elu(x), _sub(P, "bn_conv2."), "pre", training))
    x = global_average(x, mask).squeeze()
    return F.linear(x, P["fc_mu.weight"], P["fc_mu.bias"]), F.linear(x, P["fc_logvar.weight"], P["fc_logvar.bias"])


def dc_dir_model(P, Di, DiA, mask, inputs, layers, training=True):
    """dense_correspondence DirModel.forward, src/dense_correspondence/models.py:161-182 (2-D block-diagonal operators)."""
    b = inputs.shape[0]
    v = graph_conv1x1(inputs, _sub(P, "conv1."))
    f = v.new_zeros(b, DiA.shape[-1] // 4 // b, 128)
    for i in range(layers):
        Pi = _sub(P, "rn%d." % i)
        if i % 2 == 0:
            v, f = dir_resnet2(Pi, Di, DiA, v, f, training)
        else:
            v = avg_resnet2(Pi, mask, v, training)
    x = graph_conv1x1(F.elu(v), _sub(P, "conv2."), "pre", training)
    return x + inputs[:, :, -3:].repeat(1, 1, 40)


def siamese(P, op_a, op_b, input_a, input_b, layers, tower="dirac", training=True):
    """dense_correspondence SiameseModel.forward, src/dense_correspondence/models.py:199-203: shared tower on both
    shapes, then torch.bmm(FA, FB^T).  BatchNorm running statistics are not tracked by this functional restatement."""
    Pm = _sub(P, "model.")
    if tower == "dirac":
        FA = dc_dir_model(Pm, *op_a, input_a, layers, training)
        FB = dc_dir_model(Pm, *op_b, input_b, layers, training)
    else:
        FA = arap_lap_model(Pm, *op_a, input_a, layers, training)
        FB = arap_lap_model(Pm, *op_b, input_b, layers, training)
    return torch.bmm(FA, FB.transpose(1, 2))
