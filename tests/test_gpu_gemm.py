"""tcgen05 TF32 GEMM (sn_gemm_tf32_f32) vs an fp64 reference.

Tolerance: the reference Linear is fp32; 3xTF32 (default) must stay within 8 * eps_fp32 * K-independent bound
|err| <= 2e-6 * (|A| |B|^T + |bias| + |rscale R|); the single-pass mode is checked at 2e-3 of the same magnitude.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_gemm(A, B, bias=None, R=None, rscale=None, single_pass=False, out=None):
    from surfacenetworks_b200 import _native as N
    M, K = A.shape
    Nn = B.shape[0]
    C = torch.empty(M, Nn, device=DEV) if out is None else out
    p = lambda t: 0 if t is None else t.data_ptr()
    N.call("sn_gemm_tf32_f32", p(A), A.stride(0), p(B), B.stride(0), p(bias), p(R), 0 if R is None else R.stride(0),
           p(rscale), p(C), C.stride(0), M, Nn, K, N.SN_GEMM_SINGLE_PASS if single_pass else 0,
           torch.cuda.current_stream().cuda_stream)
    return C


def reference(A, B, bias, R, rscale):
    A64, B64 = A.double(), B.double()
    y = A64 @ B64.t()
    mag = A64.abs() @ B64.abs().t()
    if bias is not None:
        y = y + bias.double()
        mag = mag + bias.double().abs()
    if R is not None:
        s = rscale.double() if rscale is not None else torch.ones(B.shape[0], device=DEV, dtype=torch.float64)
        y = y + s * R.double()
        mag = mag + (s * R.double()).abs()
    return y, mag


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1000, 128, 256), (4096 + 77, 256, 128), (300, 64, 64), (20000, 128, 128)])
@pytest.mark.parametrize("mode", ["plain", "bias", "full"])
def test_gemm_matches_fp64(M, N, K, mode):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g) * torch.exp(torch.randn(M, 1, device=DEV, generator=g))
    B = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g) if mode != "plain" else None
    R = torch.randn(M, N, device=DEV, generator=g) if mode == "full" else None
    rscale = torch.randn(N, device=DEV, generator=g) if mode == "full" else None
    y64, mag = reference(A, B, bias, R, rscale)
    C = run_gemm(A, B, bias, R, rscale)
    torch.cuda.synchronize()
    err = (C.double() - y64).abs()
    assert torch.all(err <= 2e-6 * mag + 1e-30), "3xTF32: max err/mag %g" % float((err / (mag + 1e-300)).max())
    C1 = run_gemm(A, B, bias, R, rscale, single_pass=True)
    err1 = (C1.double() - y64).abs()
    assert torch.all(err1 <= 2e-3 * mag + 1e-30), "TF32: max err/mag %g" % float((err1 / (mag + 1e-300)).max())
    # the fp32 cuBLAS result (what the reference's nn.Linear computes on a GPU) is at most 8x closer to fp64 than we are
    ref32 = torch.nn.functional.linear(A, B, bias)
    if R is not None:
        ref32 = ref32 + rscale * R
    e32 = float(((ref32.double() - y64).abs() / (mag + 1e-300)).max())
    assert float((err / (mag + 1e-300)).max()) <= max(8 * e32, 2e-6)


def test_gemm_strided_operands_and_repeatability():
    """A = the [rows, 2C] concat buffer view, C written into a wider buffer; two launches give identical bits."""
    g = torch.Generator(device=DEV).manual_seed(5)
    M, N, K = 3000, 128, 256
    Zbuf = torch.randn(M, K + 64, device=DEV, generator=g)
    A = Zbuf[:, :K]
    B = torch.randn(N, K, device=DEV, generator=g) / 16
    out = torch.zeros(M, 2 * N, device=DEV)
    run_gemm(A, B, out=out[:, N:])
    y64, mag = reference(A, B, None, None, None)
    assert torch.all((out[:, N:].double() - y64).abs() <= 2e-6 * mag)
    assert torch.all(out[:, :N] == 0)
    again = torch.zeros(M, 2 * N, device=DEV)
    run_gemm(A, B, out=again[:, N:])
    assert torch.equal(out, again)


def test_gemm_argument_errors():
    from surfacenetworks_b200 import _native as N
    A = torch.zeros(128, 64, device=DEV)
    B = torch.zeros(100, 64, device=DEV)
    with pytest.raises(N.SurfnetError):
        run_gemm(A, B)                       # N = 100 unsupported
    with pytest.raises(N.SurfnetError):
        run_gemm(torch.zeros(128, 48, device=DEV), torch.zeros(128, 48, device=DEV))   # K % 32 != 0
