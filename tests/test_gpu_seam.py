"""The reference's native seam (src/utils/cuda/: batch_csr, sparse_bmm, SparseBMMFunc) re-exported with the same names,
signatures and layouts by surfacenetworks_b200.cuda -- checked against the oracle's restatement of the two reference
kernels (oracle/sn_oracle.c: batch_csr.cu:13-47, sparse_bmm.cu:16-61) and against dense torch.bmm, the check the
reference's own __main__ blocks print (sparse_bmm.py:65-94, sparse_bmm_func.py:74-111)."""
import numpy as np
import pytest
import torch

from det import det_array
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
EPS32 = float(np.finfo(np.float32).eps)


def _cat_L(golden):
    b = golden("batching")
    S = golden.pt_coo("batching", "cat_L", DEV)
    return b, S


def test_batch_csr_layout_matches_reference_kernel(golden):
    from surfacenetworks_b200.cuda import batch_csr
    b, S = _cat_L(golden)
    B, R, C = S.shape
    col_ind, col_ptr = batch_csr(S._indices(), S.size())
    ref_ind, ref_ptr = c_oracle.batch_csr(b["cat_L_idx"], B, R)
    assert col_ind.dtype == torch.int64 and col_ptr.dtype == torch.int64 and tuple(col_ptr.shape) == (B, R + 1)
    assert np.array_equal(col_ind.cpu().numpy(), ref_ind)
    assert np.array_equal(col_ptr.cpu().numpy(), ref_ptr)


@pytest.mark.parametrize("K", [1, 5, 16, 128])
def test_sparse_bmm_matches_reference_kernel(golden, K):
    from surfacenetworks_b200.cuda import batch_csr, sparse_bmm
    b, S = _cat_L(golden)
    B, R, C = S.shape
    dense = det_array((B, C, K), 50 + K)
    dg = torch.from_numpy(dense).to(DEV)
    col_ind, col_ptr = batch_csr(S._indices(), S.size())
    out = sparse_bmm(S._values(), col_ind, col_ptr, S.size(), dg)
    assert tuple(out.shape) == (B, R, K)
    ref = c_oracle.sparse_bmm(b["cat_L_val"], *c_oracle.batch_csr(b["cat_L_idx"], B, R), B, R, dense)
    dense_bmm = torch.bmm(S.to_dense(), dg)
    mag = torch.bmm(S.to_dense().abs(), dg.abs()).cpu().numpy()
    assert np.all(np.abs(out.cpu().numpy() - ref) <= 32 * EPS32 * mag + 1e-30)
    assert np.all(np.abs(out.cpu().numpy() - dense_bmm.cpu().numpy()) <= 32 * EPS32 * mag + 1e-30)
    # raw (col_ind, col_ptr) without the cached structure (e.g. tensors that went through .clone())
    out2 = sparse_bmm(S._values(), col_ind.clone(), col_ptr.clone(), S.size(), dg)
    assert torch.equal(out, out2)


def test_sparse_bmm_func_forward_backward(golden):
    from surfacenetworks_b200.cuda import SparseBMMFunc
    _, S = _cat_L(golden)
    B, R, C = S.shape
    K = 24
    x = torch.from_numpy(det_array((B, C, K), 3)).to(DEV).requires_grad_(True)
    g = torch.from_numpy(det_array((B, R, K), 4)).to(DEV)
    y = SparseBMMFunc.apply(S, x)
    y.backward(g)
    xd = x.detach().clone().requires_grad_(True)
    yd = torch.bmm(S.to_dense(), xd)
    yd.backward(g)
    Sd = S.to_dense().abs()
    assert torch.all((y - yd).abs() <= 32 * EPS32 * torch.bmm(Sd, xd.abs()) + 1e-30)
    assert torch.all((x.grad - xd.grad).abs() <= 32 * EPS32 * torch.bmm(Sd.transpose(1, 2), g.abs()) + 1e-30)
    y2 = SparseBMMFunc()(S, x.detach())                      # the reference's legacy call style, utils_pt.py:199
    assert torch.equal(y2, y.detach())
    with pytest.raises(ValueError):
        SparseBMMFunc.apply(S, x.detach()[0])


def _to_3d(S2, B):
    """Block-diagonal 2-D COO [B*R x B*C] -> the sparse_cat layout [B, R, C] (utils_pt.py:21-39)."""
    R, C = S2.shape[0] // B, S2.shape[1] // B
    idx, val = S2._indices(), S2._values()
    b = idx[0] // R
    return torch.sparse_coo_tensor(torch.stack([b, idx[0] - b * R, idx[1] - b * C]), val, (B, R, C)).coalesce()


def test_dirac_block_accepts_the_3d_layout(golden):
    """utils_pt.py:197-199,209-211: a 3-D Di / DiA routes through SparseBMMFunc in the reference (dead there: NameError);
    here the 3-D layout gives the same block output as the 2-D block-diagonal one."""
    from surfacenetworks_b200 import utils_pt as U
    from det import det_fill
    b = golden("batching")
    nv, nf = int(b["nv"]), int(b["nf"])
    C = 16
    v = torch.from_numpy(det_array((2, nv, C), 1)).to(DEV)
    f = torch.from_numpy(det_array((2, nf, C), 2)).to(DEV)
    Di, DiA = golden.pt_coo("batching", "diag_Di", DEV), golden.pt_coo("batching", "diag_DiA", DEV)
    out2 = det_fill(U.DirResNet2(C), 5).to(DEV)(Di, DiA, v, f)
    out3 = det_fill(U.DirResNet2(C), 5).to(DEV)(_to_3d(Di, 2), _to_3d(DiA, 2), v, f)
    assert torch.equal(out2[0], out3[0]) and torch.equal(out2[1], out3[1])
