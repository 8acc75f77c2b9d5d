"""The model stacks' two ends on their own kernels (csrc/head_tail.cu) against the torch composites the reference runs:
conv1 = nn.Linear(3 | 6 -> N) (as_rigid_as_possible/models.py:112, utils_pt.py:89), the `+ inputs[:, :, -3:].repeat(1, 1, 40)`
output head (models.py:152) and the masked smooth-L1 loss (main.py:225-226).  fp32 elementwise work: tolerances are a few
ulps of the summed magnitudes (stated per check)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,N,rows", [(6, 128, 128000), (3, 128, 16000), (6, 64, 777), (3, 256, 5)])
def test_small_k_linear_forward_backward(K, N, rows):
    from surfacenetworks_b200 import fused
    g = torch.Generator(device=DEV).manual_seed(K * N + rows)
    fc = torch.nn.Linear(K, N).to(DEV)
    x = torch.randn(rows, K, device=DEV, generator=g)
    w = torch.randn(rows, N, device=DEV, generator=g)
    assert fused.smallk_linear_supported(x, fc)
    y = fused.smallk_linear(x, fc)
    (y * w).sum().backward()
    gw, gb = fc.weight.grad.clone(), fc.bias.grad.clone()
    fc.zero_grad()
    y64 = x.double() @ fc.weight.double().t() + fc.bias.double()
    mag = x.double().abs() @ fc.weight.double().abs().t() + fc.bias.double().abs()
    assert torch.all((y.double() - y64).abs() <= 4e-7 * mag + 1e-30)
    gw64 = w.double().t() @ x.double()
    gwm = w.double().abs().t() @ x.double().abs()
    assert torch.all((gw.double() - gw64).abs() <= 1e-6 * gwm + 1e-30), float(((gw.double() - gw64).abs() / gwm).max())
    gb64 = w.double().sum(0)
    assert torch.all((gb.double() - gb64).abs() <= 1e-6 * w.double().abs().sum(0) + 1e-30)
    # deterministic
    y2 = fused.smallk_linear(x, fc)
    (y2 * w).sum().backward()
    assert torch.equal(fc.weight.grad, gw) and torch.equal(fc.bias.grad, gb)
    # input gradient (torch path) when requested
    xr = x.clone().requires_grad_(True)
    fused.smallk_linear(xr, fc).sum().backward()
    assert torch.allclose(xr.grad, fc.weight.detach().sum(0).expand_as(xr), rtol=1e-5, atol=1e-6)


def test_head_add_tiled_and_padded_slice_gradient():
    from surfacenetworks_b200 import fused
    B, V, n, n_pad = 3, 501, 120, 128
    g = torch.Generator(device=DEV).manual_seed(1)
    yp = torch.randn(B * V, n_pad, device=DEV, generator=g).requires_grad_(True)
    inputs = torch.randn(B, V, 6, device=DEV, generator=g)
    w = torch.randn(B, V, n, device=DEV, generator=g)
    y = fused._SliceCols.apply(yp, n).view(B, V, n)
    out = fused.head_add_tiled(y, inputs, 40)
    ref = yp.detach()[:, :n].reshape(B, V, n) + inputs[:, :, -3:].repeat(1, 1, 40)
    assert torch.equal(out, ref)
    (out * w).sum().backward()
    gref = torch.zeros(B * V, n_pad, device=DEV)
    gref[:, :n] = w.reshape(B * V, n)
    assert torch.equal(yp.grad, gref)
    # layouts outside the fused path take the torch composite (same values)
    y3 = torch.randn(B, V, 9, device=DEV, generator=g)
    assert torch.equal(fused.head_add_tiled(y3, inputs, 3), y3 + inputs[:, :, -3:].repeat(1, 1, 3))


@pytest.mark.parametrize("B,V,C", [(64, 2000, 120), (3, 77, 120), (2, 5, 4)])
def test_masked_smooth_l1_matches_torch(B, V, C):
    from surfacenetworks_b200 import models as M
    g = torch.Generator(device=DEV).manual_seed(B + V)
    out = (3 * torch.randn(B, V, C, device=DEV, generator=g)).requires_grad_(True)
    tgt = torch.randn(B, V, C, device=DEV, generator=g)
    mask = (torch.rand(B, V, 1, device=DEV, generator=g) > 0.2).float()
    loss = M.arap_loss(out, tgt, mask, B)
    (loss * 1.7).backward()
    got = out.grad.clone()
    out.grad = None
    ref = F.smooth_l1_loss(out * mask.expand_as(out), tgt, reduction="sum") / B
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * abs(float(ref)) + 1e-12          # fp64-accumulated vs torch's fp32 tree
    assert torch.allclose(got, out.grad, rtol=2e-6, atol=1e-9)
    l2 = M.arap_loss(out.detach(), tgt, mask, B)
    assert float(l2) == float(loss)
