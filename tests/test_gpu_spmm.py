"""Parity of the sm_100a SpMM / format kernels (through the C ABI) against the oracle and the reference goldens.

Tolerance (stated once, used everywhere): componentwise
    |y - y64| <= 32 * eps_fp32 * (|S| |x|)
where y64 is the double-precision product and |S||x| the magnitude bound, both from oracle/sn_oracle.c.  The
reference's own fp32 torch.mm result sits at <= 2.3e-7 (|S||x|) (~2 eps) from y64; absolute
tolerances are meaningless here because |L| reaches 1e5.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from det import det_array
from oracle import c_oracle

pytestmark = pytest.mark.gpu
EPS32 = float(np.finfo(np.float32).eps)
DEV = "cuda"


def ops_mod():
    from surfacenetworks_b200 import operators
    return operators


def within_bound(y, y64, bound, what, k=32):
    y = np.asarray(y, dtype=np.float64)
    err = np.abs(y - y64)
    tol = k * EPS32 * bound + 1e-30
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError("%s: %d elements exceed %d eps |S||x|; worst at %s err %g tol %g"
                             % (what, int((err > tol).sum()), k, i, err[i], tol[i]))
    return float((err / (EPS32 * bound + 1e-300)).max())


def coo_cuda(idx, val, shape, coalesced=True):
    return torch.sparse_coo_tensor(torch.from_numpy(np.asarray(idx)), torch.from_numpy(np.asarray(val)), shape,
                                   is_coalesced=coalesced).to(DEV)


# ------------------------------------------------------------------------------------------- golden fixtures
def test_cube_golden(golden):
    """BASELINE cfg1: cube.ply, 8-vertex Laplacian x 16-dim features (+ the cube's Dirac pair)."""
    O, d = ops_mod(), golden("spmm")
    for name, xkey, ykey, kind in (("cube_L", "cube_x", "cube_Lx", "csr"), ("cube_Di", "cube_x", "cube_Dix", "bsr4"),
                                   ("cube_DiA", "cube_f", "cube_DiAf", "bsr4")):
        row, col, val, shape = golden.coo("operators", name)
        S = coo_cuda(np.stack([row, col]), val, shape, coalesced=False).coalesce()
        x = torch.from_numpy(d[xkey]).to(DEV)
        if kind == "csr":
            y = O.as_csr(S).apply(x)
            y64, bound = c_oracle.coo_mm_f64(row, col, val, shape[0], d[xkey])
        else:
            y = O.as_bsr4(S).apply(x)
            y64, bound = c_oracle.dirac_view_mm_f64(row, col, val, shape[0] // 4, d[xkey])
        within_bound(y.cpu().numpy(), y64, bound, name)
        assert_close(y.cpu().numpy(), d[ykey], 1e-5, 1e-5, name + " vs reference torch.mm")


@pytest.mark.parametrize("C", [4, 32, 40, 128])
def test_batch_golden(golden, C):
    """Ragged batch of two meshes (block-diagonal, padded rows empty): forward and transposed products."""
    O, d, b = ops_mod(), golden("spmm"), golden("batching")
    nv, nf = int(b["nv"]), int(b["nf"])
    x = torch.from_numpy(d["b_x%d" % C]).to(DEV)
    f = torch.from_numpy(d["b_f%d" % C]).to(DEV)

    L = O.as_csr(golden.pt_coo("batching", "diag_L", DEV))
    idx, val = b["diag_L_idx"], b["diag_L_val"]
    y64, bound = c_oracle.coo_mm_f64(idx[0], idx[1], val, 2 * nv, d["b_x%d" % C])
    within_bound(L.apply(x).cpu().numpy(), y64, bound, "L x")
    within_bound(d["b_Lx%d" % C], y64, bound, "reference L x")
    y64, bound = c_oracle.coo_mm_f64(idx[1], idx[0], val, 2 * nv, d["b_x%d" % C])
    within_bound(L.T.apply(x).cpu().numpy(), y64, bound, "L^T x")
    within_bound(d["b_LTx%d" % C], y64, bound, "reference L^T x")

    Di = O.as_bsr4(golden.pt_coo("batching", "diag_Di", DEV))
    idx, val = b["diag_Di_idx"], b["diag_Di_val"]
    assert Di.n_brows == 2 * nf and Di.n_bcols == 2 * nv
    y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, 2 * nf, d["b_x%d" % C])
    within_bound(Di.apply(x).cpu().numpy(), y64, bound, "Di x")
    within_bound(d["b_Dix%d" % C], y64, bound, "reference Di x")
    y64, bound = c_oracle.dirac_view_mm_f64(idx[1], idx[0], val, 2 * nv, d["b_f%d" % C])
    within_bound(Di.T.apply(f).cpu().numpy(), y64, bound, "Di^T f")
    within_bound(d["b_DiTf%d" % C], y64, bound, "reference Di^T f")

    DiA = O.as_bsr4(golden.pt_coo("batching", "diag_DiA", DEV))
    idx, val = b["diag_DiA_idx"], b["diag_DiA_val"]
    y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, 2 * nv, d["b_f%d" % C])
    within_bound(DiA.apply(f).cpu().numpy(), y64, bound, "DiA f")
    within_bound(d["b_DiAf%d" % C], y64, bound, "reference DiA f")


def test_3d_layout_equals_block_diagonal(golden):
    """sparse_cat's [B, R, C] layout (input of the reference's batch_csr kernel) -> same CSR as sparse_diag_cat."""
    O, b = ops_mod(), golden("batching")
    L2 = O.as_csr(golden.pt_coo("batching", "diag_L", DEV))
    L3 = O.as_csr(golden.pt_coo("batching", "cat_L", DEV))
    assert L3.shape == L2.shape
    assert torch.equal(L2.rowptr, L3.rowptr) and torch.equal(L2.colind, L3.colind) and torch.equal(L2.val, L3.val)
    # and it reproduces the reference kernel's col_ptr (global offsets per batch) modulo layout
    nv = int(b["nv"])
    col_ind, col_ptr = c_oracle.batch_csr(b["cat_L_idx"], 2, nv)
    rp = L3.rowptr.cpu().numpy()
    assert np.array_equal(rp[:nv + 1], col_ptr[0]) and np.array_equal(rp[nv:], col_ptr[1])
    assert np.array_equal(L3.colind.cpu().numpy() - np.repeat([0, nv], [col_ptr[0, -1], col_ptr[1, -1] - col_ptr[0, -1]]), col_ind)
    D2 = O.as_bsr4(golden.pt_coo("batching", "diag_Di", DEV))
    D3 = O.as_bsr4(golden.pt_coo("batching", "cat_Di", DEV))
    assert torch.equal(D2.browptr, D3.browptr) and torch.equal(D2.bcolind, D3.bcolind) and torch.equal(D2.bval, D3.bval)


# ------------------------------------------------------------------------------------------- format edge cases
def random_coo(rng, n_rows, n_cols, nnz, empty_rows=()):
    allowed = np.setdiff1d(np.arange(n_rows), np.asarray(empty_rows, dtype=np.int64))
    row = allowed[rng.integers(0, len(allowed), nnz)]
    col = rng.integers(0, n_cols, nnz)
    val = rng.standard_normal(nnz).astype(np.float32)
    return row.astype(np.int64), col.astype(np.int64), val


@pytest.mark.parametrize("n_rows,n_cols,nnz", [(1, 1, 1), (37, 53, 200), (1000, 40, 5000), (64, 64, 0), (5000, 5000, 3)])
def test_unsorted_duplicates_empty_rows(n_rows, n_cols, nnz):
    """Unsorted COO with duplicate entries and empty rows anywhere (interior ones break the reference kernel)."""
    O = ops_mod()
    rng = np.random.default_rng(n_rows * 7 + nnz)
    row, col, val = random_coo(rng, n_rows, n_cols, nnz, empty_rows=(0, n_rows // 2, n_rows - 1) if n_rows > 3 else ())
    S = coo_cuda(np.stack([row, col]), val, (n_rows, n_cols), coalesced=False)
    op = O.CsrOperator.from_torch_coo(S)
    rp = op.rowptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == nnz and np.all(np.diff(rp) >= 0)
    assert np.array_equal(np.diff(rp), np.bincount(row, minlength=n_rows))
    x = det_array((n_cols, 8), 3)
    y64, bound = c_oracle.coo_mm_f64(row, col, val, n_rows, x)
    within_bound(op.apply(torch.from_numpy(x).to(DEV)).cpu().numpy(), y64, bound, "unsorted csr")
    xt = det_array((n_rows, 8), 4)
    y64, bound = c_oracle.coo_mm_f64(col, row, val, n_cols, xt)
    within_bound(op.T.apply(torch.from_numpy(xt).to(DEV)).cpu().numpy(), y64, bound, "unsorted csr^T")
    # the sorted fast path must agree with the general path on coalesced input
    Sc = S.coalesce()
    a, b = O.CsrOperator.from_torch_coo(Sc), None
    idx = Sc._indices().cpu().numpy()
    perm = rng.permutation(idx.shape[1])
    b = O.CsrOperator.from_torch_coo(coo_cuda(idx[:, perm], Sc._values().cpu().numpy()[perm], (n_rows, n_cols), False))
    n = a.nnz
    assert n == b.nnz and torch.equal(a.rowptr, b.rowptr)
    assert torch.equal(a.colind[:n], b.colind[:n]) and torch.equal(a.val[:n], b.val[:n])


def test_bsr4_general_blocks():
    """General 4x4 blocks (not just quaternion structure): dense random blocks, ragged block rows, empty block rows."""
    O = ops_mod()
    rng = np.random.default_rng(5)
    nbr, nbc = 50, 30
    rows, cols, vals = [], [], []
    for br in range(nbr):
        if br % 7 == 3:
            continue  # empty block row
        for bc in rng.choice(nbc, size=rng.integers(1, 9), replace=False):
            blk = rng.standard_normal((4, 4)).astype(np.float32)
            blk[rng.random((4, 4)) < 0.3] = 0  # partially filled blocks
            p, q = np.nonzero(blk)
            rows += list(4 * br + p)
            cols += list(4 * bc + q)
            vals += list(blk[p, q])
    row, col, val = np.array(rows, np.int64), np.array(cols, np.int64), np.array(vals, np.float32)
    perm = rng.permutation(len(val))
    S = coo_cuda(np.stack([row[perm], col[perm]]), val[perm], (4 * nbr, 4 * nbc), coalesced=False)
    op = O.Bsr4Operator.from_torch_coo(S)
    for C in (4, 16, 24, 64, 128, 256):
        x = det_array((nbc, C), C)
        y64, bound = c_oracle.dirac_view_mm_f64(row, col, val, nbr, x)
        within_bound(op.apply(torch.from_numpy(x).to(DEV)).cpu().numpy(), y64, bound, "bsr4 C=%d" % C)
        g = det_array((nbr, C), C + 1)
        y64, bound = c_oracle.dirac_view_mm_f64(col, row, val, nbc, g)
        within_bound(op.T.apply(torch.from_numpy(g).to(DEV)).cpu().numpy(), y64, bound, "bsr4^T C=%d" % C)


@pytest.mark.parametrize("C", [1, 3, 4, 8, 16, 20, 64, 100, 128, 192, 256, 512])
def test_csr_feature_widths_and_strides(C):
    """Every dispatch bucket of sn_csr_spmm_f32, with strided operands (the concat-buffer case) and ELU-on-load."""
    O = ops_mod()
    rng = np.random.default_rng(C)
    n = 301
    row, col, val = random_coo(rng, n, n, 2000)
    op = O.CsrOperator.from_torch_coo(coo_cuda(np.stack([row, col]), val, (n, n), False))
    x = det_array((n, C), 100 + C)
    xg = torch.from_numpy(x).to(DEV)
    y64, bound = c_oracle.coo_mm_f64(row, col, val, n, x)
    y = op.apply(xg)
    within_bound(y.cpu().numpy(), y64, bound, "contiguous")
    within_bound(op.apply(xg, direct_gather=True).cpu().numpy(), y64, bound, "direct-gather kernel")
    for variant in (1, 2, 3, 5, 9):        # tuning variants of the row-group kernel share its summation order
        assert torch.equal(y, op.apply(xg, variant=variant)), "row-group variant %d" % variant
    Z = torch.zeros(n, 2 * C + 4, device=DEV)
    Z[:, :C] = xg
    op.apply(Z[:, :C], out=Z[:, C:2 * C])
    within_bound(Z[:, C:2 * C].cpu().numpy(), y64, bound, "strided in/out")
    assert torch.all(Z[:, 2 * C:] == 0) and torch.equal(Z[:, :C], xg)        # nothing written outside the target
    xe = c_oracle.elu_f32(x)
    y64, bound = c_oracle.coo_mm_f64(row, col, val, n, xe)
    within_bound(op.apply(xg, elu_input=True).cpu().numpy(), y64, bound, "elu on load", k=64)


@pytest.mark.parametrize("C", [4, 8, 16, 32, 48, 64, 128, 256, 512])
def test_bsr4_feature_widths_and_strides(golden, C):
    O, b = ops_mod(), golden("batching")
    nv, nf = int(b["nv"]), int(b["nf"])
    Di = O.as_bsr4(golden.pt_coo("batching", "diag_Di", DEV))
    idx, val = b["diag_Di_idx"], b["diag_Di_val"]
    x = det_array((2 * nv, C), 200 + C)
    xg = torch.from_numpy(x).to(DEV)
    y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, 2 * nf, x)
    y = Di.apply(xg)
    within_bound(y.cpu().numpy(), y64, bound, "contiguous")
    # tuning variants of the row-group kernel and the small-operator kernel (6; 7 = persistent kernel forced) share one
    # summation order; below C = 32 the persistent kernel does not exist and 1-5 / 7 fall back to the direct-gather kernel
    for variant in (1, 2, 3, 5, 6, 7, 9) if C >= 32 else (6,):
        assert torch.equal(y, Di.apply(xg, variant=variant)), "row-group variant %d" % variant
    if C == 16:
        within_bound(Di.apply(xg, variant=7).cpu().numpy(), y64, bound, "C = 16 without the small-operator kernel")
    # cp.async streaming kernel (C = 128/256/512) and direct-gather kernel use the same summation order: bit-identical
    yd = Di.apply(xg, direct_gather=True)
    within_bound(yd.cpu().numpy(), y64, bound, "direct-gather kernel")
    assert torch.equal(yd, Di.apply(xg, smem_stream=True))
    Xs = torch.zeros(2 * nv, 2 * C, device=DEV)
    Xs[:, :C] = xg
    Z = torch.zeros(2 * nf, 2 * C, device=DEV)
    Di.apply(Xs[:, :C], out=Z[:, C:])
    within_bound(Z[:, C:].cpu().numpy(), y64, bound, "strided")
    assert torch.all(Z[:, :C] == 0)
    y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, 2 * nf, c_oracle.elu_f32(x))
    within_bound(Di.apply(xg, elu_input=True).cpu().numpy(), y64, bound, "elu on load", k=64)


# ------------------------------------------------------------------------------------------- row-group kernel paths
@pytest.mark.parametrize("C", [16, 32, 64, 128, 256, 512])
def test_rowgroup_long_rows_and_empty_runs(C):
    """Row-group kernel (spmm_rowgroup.cu) corner paths: rows far longer than the staged index ring (global-load
    fallback), runs of empty rows at the start / middle / end of a warp-tile, one dense row among empty ones,
    row counts that are not a multiple of the warp-tile, every tile length (variants 1-3)."""
    O = ops_mod()
    rng = np.random.default_rng(1000 + C)
    n_rows, n_cols = 333, 190
    lens = rng.integers(0, 4, n_rows)
    lens[5] = 150                       # longer than any staged ring
    lens[40:90] = 0                     # empty warp-tiles
    lens[90] = 97
    lens[300:] = rng.integers(20, 60, n_rows - 300)
    lens[-1] = 0
    row = np.repeat(np.arange(n_rows), lens).astype(np.int64)
    col = np.concatenate([np.sort(rng.choice(n_cols, n, replace=False)) for n in lens]).astype(np.int64) \
        if row.size else np.zeros(0, np.int64)
    val = rng.standard_normal(row.size).astype(np.float32)
    op = O.CsrOperator.from_torch_coo(coo_cuda(np.stack([row, col]), val, (n_rows, n_cols), True))
    x = det_array((n_cols, C), 300 + C)
    xg = torch.from_numpy(x).to(DEV)
    y64, bound = c_oracle.coo_mm_f64(row, col, val, n_rows, x)
    y = op.apply(xg)
    within_bound(y.cpu().numpy(), y64, bound, "csr rowgroup")
    # 5: two gathers in flight in registers; 6 / 7: small-operator kernel (spmm_rowdirect.cu) / persistent kernel forced;
    # 9: gathers through shared memory (block operators; the scalar family falls back to its default)
    for variant in (1, 2, 3, 5, 6, 7, 9):
        assert torch.equal(y, op.apply(xg, variant=variant)), "csr variant %d" % variant
    # the same pattern as 4x4 blocks (dense random blocks): block row r has lens[r] blocks
    blk = rng.standard_normal((row.size, 4, 4)).astype(np.float32)
    p, q = np.meshgrid(np.arange(4), np.arange(4), indexing="ij")
    brow = (4 * row[:, None, None] + p).ravel()
    bcol = (4 * col[:, None, None] + q).ravel()
    opb = O.Bsr4Operator.from_torch_coo(coo_cuda(np.stack([brow, bcol]), blk.ravel(), (4 * n_rows, 4 * n_cols), False))
    assert opb.n_blocks == row.size
    y64, bound = c_oracle.dirac_view_mm_f64(brow, bcol, blk.ravel(), n_rows, x)
    y = opb.apply(xg)
    within_bound(y.cpu().numpy(), y64, bound, "bsr4 rowgroup")
    within_bound(opb.apply(xg, direct_gather=True).cpu().numpy(), y64, bound, "bsr4 direct")
    for variant in (1, 2, 3, 5, 6, 7, 9) if C >= 32 else (6,):      # C = 16: see test_bsr4_feature_widths_and_strides
        assert torch.equal(y, opb.apply(xg, variant=variant)), "bsr4 variant %d" % variant
    # the row-length hint only changes how many gathers are in flight (here it is wrong on purpose: rows hold up to 150)
    opb.max_row_blocks = 3
    assert torch.equal(y, opb.apply(xg, variant=6)), "bsr4 small-operator kernel, three entries in flight"


@pytest.mark.parametrize("C,n_rows", [(16, 480_000), (64, 200_000), (128, 150_001), (512, 40_000)])
def test_rowgroup_persistent_warps(C, n_rows):
    """Enough warp-tiles that every persistent warp walks several of them (index ring rotation, both tile lengths)."""
    O = ops_mod()
    rng = np.random.default_rng(C)
    n_cols = 50_000
    nnz = 6 * n_rows
    row = np.sort(rng.integers(0, n_rows, nnz)).astype(np.int64)
    col = rng.integers(0, n_cols, nnz).astype(np.int64)
    val = rng.standard_normal(nnz).astype(np.float32)
    op = O.CsrOperator.from_torch_coo(coo_cuda(np.stack([row, col]), val, (n_rows, n_cols), False))
    x = det_array((n_cols, C), 7)
    xg = torch.from_numpy(x).to(DEV)
    y = op.apply(xg)
    yd = op.apply(xg, direct_gather=True)
    mag = O.CsrOperator(op.rowptr, op.colind, op.val.abs(), op.n_rows, op.n_cols).apply(xg.abs(), direct_gather=True)
    assert torch.all((y - yd).abs() <= 64 * EPS32 * mag + 1e-30)
    for variant in (1, 3, 6, 9):               # 6: the small-operator kernel forced onto a large operator
        assert torch.equal(y, op.apply(xg, variant=variant)), "variant %d" % variant
    # oracle check on a row sample (full fp64 product of 480k x 16 ... 150k x 128 stays cheap on the CPU)
    y64, bound = c_oracle.coo_mm_f64(row, col, val, n_rows, x)
    within_bound(y.cpu().numpy(), y64, bound, "csr persistent")


@pytest.mark.parametrize("C", [32, 64, 128, 256, 512, 48])
def test_spmm_epilogue_matches_separate_passes(C):
    """sn_*_spmm_epilogue_f32: Y = (S X + G) .* elu'(A) in the store path == SpMM, add, elementwise derivative as three
    steps (same fp32 operations in the same order: bit-identical); unsupported widths report None."""
    O = ops_mod()
    rng = np.random.default_rng(C)
    n_rows, n_cols = 500, 300
    row, col, val = random_coo(rng, n_rows, n_cols, 3000, empty_rows=(0, 7, 8, 9, 499))
    opc = O.CsrOperator.from_torch_coo(coo_cuda(np.stack([row, col]), val, (n_rows, n_cols), False))
    brow, bcol, bvals = random_coo(rng, 4 * n_rows, 4 * n_cols, 9000, empty_rows=range(16, 40))
    opb = O.Bsr4Operator.from_torch_coo(coo_cuda(np.stack([brow, bcol]), bvals, (4 * n_rows, 4 * n_cols), False))
    X = torch.from_numpy(det_array((n_cols, C), 1)).to(DEV)
    Gbuf = torch.from_numpy(det_array((n_rows, 2 * C), 2)).to(DEV)
    G = Gbuf[:, :C]                                                          # row-strided, like the halves of a dZ
    A = torch.nn.functional.elu(torch.from_numpy(det_array((n_rows, C), 3)).to(DEV))
    dA = torch.where(A > 0, torch.ones_like(A), A + 1)
    for op in (opc, opb):
        plain = op.apply(X)
        for g_, a_ in ((G, A), (None, A), (G, None)):
            y = op.apply_epilogue(X, G=g_, A=a_)
            if C == 48:
                assert y is None
                continue
            expect = plain if g_ is None else plain + g_
            expect = expect if a_ is None else expect * dA
            assert torch.equal(y, expect), (op.kind, g_ is None, a_ is None)
        if C != 48:                                                          # Y may alias G
            Gc = G.clone()
            y = op.apply_epilogue(X, G=Gc, A=A, out=Gc)
            assert torch.equal(y, (plain + G) * dA)
            G2 = Gbuf[:, C:]                                                 # addend behind the derivative
            assert torch.equal(op.apply_epilogue(X, G=G, A=A, G2=G2), (plain + G) * dA + G2)
            assert torch.equal(op.apply_epilogue(X, A=A, G2=G2), plain * dA + G2)
            assert torch.equal(op.apply_epilogue(X, G2=G2), plain + G2)


def test_elu_kernels():
    from surfacenetworks_b200 import ops
    x = np.concatenate([-np.logspace(-6, 1.2, 300), np.logspace(-6, 1.2, 300), [0.0, -0.0]]).astype(np.float32)
    x = np.tile(x, 5)[: 5 * 600].reshape(-1, 24)
    xg = torch.from_numpy(x).to(DEV)
    Z = torch.zeros(x.shape[0], 48, device=DEV)
    ops.elu_into(xg, Z[:, :24])
    assert_close(Z[:, :24].cpu().numpy(), c_oracle.elu_f32(x), 4 * EPS32, 1e-37, "elu")
    g = torch.from_numpy(det_array(x.shape, 9)).to(DEV)
    g2 = torch.from_numpy(det_array(x.shape, 10)).to(DEV)
    ref = (g + g2) * torch.where(xg > 0, torch.ones_like(xg), torch.exp(xg))
    out = torch.empty_like(g)
    ops._elu_bwd(xg, True, g, g2, out)
    assert_close(out.cpu().numpy(), ref.cpu().numpy(), 1e-6, 1e-7, "elu' from raw")
    ops._elu_bwd(Z[:, :24], False, g, None, out)
    ref = g * torch.where(xg > 0, torch.ones_like(xg), torch.exp(xg))
    assert_close(out.cpu().numpy(), ref.cpu().numpy(), 1e-6, 1e-7, "elu' from activated")


# ------------------------------------------------------------------------------------------- mesh-sized parity
@pytest.mark.parametrize("V,B,C", [(500, 8, 128), (2000, 4, 128), (2000, 2, 256), (700, 3, 64)])
def test_mesh_operators_vs_oracle(V, B, C):
    """Synthetic meshes at the BASELINE per-mesh sizes (mesh_mnist ~500 V, ARAP ~2k V), small batch."""
    from surfacenetworks_b200 import geometry, utils_pt as U
    O = ops_mod()
    Ls, Ds, DAs, nv, nf = [], [], [], 0, 0
    for s in range(B):
        Vv, F = geometry.synth_mesh(V - 3 * s, 100 + s)   # ragged: different sizes, padded to the max
        nv, nf = max(nv, Vv.shape[0]), max(nf, F.shape[0])
        D, DA = geometry.build_dirac(Vv, F)
        Ls.append(U.sp_sparse_to_pt_sparse(geometry.build_laplacian(Vv, F)))
        Ds.append(U.sp_sparse_to_pt_sparse(D))
        DAs.append(U.sp_sparse_to_pt_sparse(DA))
    L = U.sparse_diag_cat(Ls, nv, nv)
    Di = U.sparse_diag_cat(Ds, 4 * nf, 4 * nv)
    DiA = U.sparse_diag_cat(DAs, 4 * nv, 4 * nf)
    x = det_array((B * nv, C), 1)
    f = det_array((B * nf, C), 2)
    xg, fg = torch.from_numpy(x).to(DEV), torch.from_numpy(f).to(DEV)
    for S, kind, inp, ing, rows in ((L, "csr", x, xg, B * nv), (Di, "bsr4", x, xg, B * nf), (DiA, "bsr4", f, fg, B * nv)):
        idx, val = S._indices().numpy(), S._values().numpy()
        op = (O.as_csr if kind == "csr" else O.as_bsr4)(S.to(DEV))
        if kind == "csr":
            y64, bound = c_oracle.coo_mm_f64(idx[0], idx[1], val, rows, inp)
        else:
            y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, rows, inp)
        y = op.apply(ing)
        worst = within_bound(y.cpu().numpy(), y64, bound, kind)
        assert worst < 32
        assert torch.equal(y, op.apply(ing)), "run-to-run bit reproducibility"
        if kind == "bsr4":
            assert op.max_row_blocks >= 3
            yd = op.apply(ing, direct_gather=True)
            within_bound(yd.cpu().numpy(), y64, bound, "direct-gather kernel")
            assert torch.equal(yd, op.apply(ing, smem_stream=True)), "streaming vs direct-gather kernel"
            for variant in (1, 2, 3, 5, 9):
                assert torch.equal(y, op.apply(ing, variant=variant)), "row-group variant %d" % variant
            within_bound(op.T.apply(y).cpu().numpy(), *c_oracle.dirac_view_mm_f64(idx[1], idx[0], val, S.shape[1] // 4, y.cpu().numpy()), "bsr4^T")
        # the reference's own path on the same inputs (CPU torch.mm) obeys the same bound
        ref = torch.mm(S, torch.from_numpy(inp).view(S.shape[1], -1)).view(rows, C).numpy()
        within_bound(ref, y64, bound, "reference torch.mm")


def op_abs_of(O, op, kind):
    if kind == "csr":
        return O.CsrOperator(op.rowptr, op.colind, op.val.abs(), op.n_rows, op.n_cols)
    return O.Bsr4Operator(op.browptr, op.bcolind, op.bval.abs(), op.n_brows, op.n_bcols)


def test_full_size_properties():
    """BASELINE cfg3 size (B=64 meshes x 2000 V, C=128): size-independent properties instead of an O(nnz) CPU check.

    adjoint identity  <S x, y> = <x, S^T y>   (validates the transposed structures used by backward)
    linearity         S(a x + b z) = a S x + b S z
    null space        L 1 = 0 up to the |L||1| bound
    """
    from surfacenetworks_b200 import geometry, utils_pt as U
    O = ops_mod()
    B, V, C = 64, 2000, 128
    meshes = [geometry.synth_mesh(V, s) for s in range(4)]          # 4 distinct meshes, repeated 16x
    nf = max(F.shape[0] for _, F in meshes)
    Ls = [U.sp_sparse_to_pt_sparse(geometry.build_laplacian(Vv, F)) for Vv, F in meshes]
    DD = [geometry.build_dirac(Vv, F) for Vv, F in meshes]
    L = U.sparse_diag_cat([Ls[i % 4] for i in range(B)], V, V).to(DEV)
    Di = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(DD[i % 4][0]) for i in range(B)], 4 * nf, 4 * V).to(DEV)
    DiA = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(DD[i % 4][1]) for i in range(B)], 4 * V, 4 * nf).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    for S, kind, n_in, n_out in ((L, "csr", B * V, B * V), (Di, "bsr4", B * V, B * nf), (DiA, "bsr4", B * nf, B * V)):
        op = (O.as_csr if kind == "csr" else O.as_bsr4)(S)
        x = torch.randn(n_in, C, device=DEV, generator=g)
        z = torch.randn(n_in, C, device=DEV, generator=g)
        y = torch.randn(n_out, C, device=DEV, generator=g)
        Sx = op.apply(x)
        # row-group kernel vs the direct-gather kernel (validated against the oracle at the smaller sizes above)
        assert torch.all((Sx - op.apply(x, direct_gather=True)).abs() <= 64 * EPS32 * op_abs_of(O, op, kind).apply(x.abs()) + 1e-30)
        for variant in (1, 2, 3, 5, 9):
            assert torch.equal(Sx, op.apply(x, variant=variant)), (kind, variant)
        lhs = (Sx.double() * y.double()).sum().item()
        rhs = (x.double() * op.T.apply(y).double()).sum().item()
        scale = (Sx.double().abs() * y.double().abs()).sum().item()
        assert abs(lhs - rhs) <= 1e-6 * scale, (kind, lhs, rhs, scale)
        lin = op.apply(0.5 * x - 2.0 * z)
        ref = 0.5 * Sx - 2.0 * op.apply(z)
        absS = op.apply(x.abs() + z.abs()).abs() + 1.0   # loose magnitude proxy
        op_abs = op_abs_of(O, op, kind)
        mag = op_abs.apply(0.5 * x.abs() + 2.0 * z.abs())
        assert torch.all((lin - ref).abs() <= 64 * EPS32 * mag + 1e-30), kind
        del absS
        if kind == "csr":
            ones = torch.ones(n_in, C, device=DEV)
            assert torch.all(op.apply(ones).abs() <= 32 * EPS32 * op_abs.apply(ones) + 1e-30)


def test_error_paths():
    O = ops_mod()
    from surfacenetworks_b200 import _native as N
    S = coo_cuda(np.array([[0, 1], [1, 0]]), np.array([1.0, 2.0], np.float32), (2, 2), False)
    op = O.as_csr(S)
    with pytest.raises(ValueError):
        op.apply(torch.zeros(1, 4, device=DEV))                   # too few rows
    with pytest.raises(TypeError):
        op.apply(torch.zeros(2, 4, device=DEV, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        op.apply(torch.zeros(2, 4))                               # CPU tensor: no fallback
    with pytest.raises(ValueError):
        O.Bsr4Operator.from_torch_coo(coo_cuda(np.array([[0], [0]]), np.array([1.0], np.float32), (6, 8), False))
    with pytest.raises(TypeError):
        O.as_csr(S.double())
    assert N.lib.sn_csr_spmm_f32(0, 0, 0, 0, 4, 0, 4, 2, 4, 0, 0) == N.SN_ERR_ARG
    assert N.lib.sn_bsr4_spmm_f32(1, 1, 1, 1, 6, 1, 6, 2, 6, 0, 0) == N.SN_ERR_UNSUPPORTED   # C % 4 != 0
    with pytest.raises(N.SurfnetError):
        N.call("sn_csr_spmm_f32", 0, 0, 0, 0, 4, 0, 4, 2, 4, 0, 0)


def test_gpu_batch_assembly_matches_host_assembly():
    """MeshOperatorCache.assemble (sn_assemble_block_diag) == converting the host-assembled block-diagonal COO:
    identical row pointers / column indices / values for D, D*, L and their transposes, ragged meshes padded."""
    from surfacenetworks_b200 import utils_pt as U, workloads as W
    O = ops_mod()
    meshes = W.make_mesh_ops(120, [0, 1]) + W.make_mesh_ops(97, [2]) + W.make_mesh_ops(120, [3])
    nv = max(m.num_vertices for m in meshes)
    nf = max(m.num_faces for m in meshes)
    cache = O.MeshOperatorCache(DEV)
    for i, m in enumerate(meshes):
        cache.add(("Di", i), m.Di, "bsr4")
        cache.add(("DiA", i), U.sp_sparse_to_pt_sparse(m.DiA), "bsr4")     # scipy and torch COO inputs both work
        cache.add(("L", i), m.L, "csr")
    order = [2, 0, 3, 1]
    sel = [meshes[i] for i in order]
    host = W.arap_batch(sel, 0)
    for name, rows_pad, cols_pad in (("Di", nf, nv), ("DiA", nv, nf)):
        ref = O.Bsr4Operator.from_torch_coo(host[name].to(DEV))
        got = cache.assemble([(name, i) for i in order], "bsr4", rows_pad, cols_pad)
        for a, b in ((got, ref), (got.T, ref.T)):
            assert (a.n_brows, a.n_bcols, a.n_blocks) == (b.n_brows, b.n_bcols, b.n_blocks)
            assert torch.equal(a.browptr, b.browptr) and torch.equal(a.bcolind, b.bcolind) and torch.equal(a.bval, b.bval)
        # in-place re-assembly of another batch into the same buffers (CUDA-graph replay case)
        slot = cache.assemble([(name, i) for i in order], "bsr4", rows_pad, cols_pad)
        ptrs = (slot.browptr.data_ptr(), slot.bval.data_ptr(), slot.T.bval.data_ptr())
        order2 = [1, 3, 0, 2]
        cache.assemble([(name, i) for i in order2], "bsr4", rows_pad, cols_pad, out=slot)
        ref2 = O.Bsr4Operator.from_torch_coo(W.arap_batch([meshes[i] for i in order2], 0)[name].to(DEV))
        assert ptrs == (slot.browptr.data_ptr(), slot.bval.data_ptr(), slot.T.bval.data_ptr())
        assert torch.equal(slot.browptr, ref2.browptr) and torch.equal(slot.bcolind, ref2.bcolind)
        assert torch.equal(slot.bval, ref2.bval) and torch.equal(slot.T.bval, ref2.T.bval)
        x = torch.randn(slot.n_bcols, 128, device=DEV)
        assert torch.equal(slot.apply(x), ref2.apply(x))
        # the same through a plan prepared ahead of time on another stream (what a data loader does while the GPU is busy)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            plan = cache.plan([(name, i) for i in order], "bsr4", rows_pad, cols_pad)
        torch.cuda.current_stream().wait_stream(side)
        cache.assemble(None, "bsr4", rows_pad, cols_pad, out=slot, plan=plan)
        assert torch.equal(slot.browptr, ref.browptr) and torch.equal(slot.bcolind, ref.bcolind)
        assert torch.equal(slot.bval, ref.bval) and torch.equal(slot.T.bval, ref.T.bval) and slot.n_blocks == ref.n_blocks
        with pytest.raises(ValueError):
            cache.assemble(None, "bsr4", rows_pad + 1, cols_pad, out=slot, plan=plan)
    refL = O.CsrOperator.from_torch_coo(W.lap_batch(sel)["L"].to(DEV))
    gotL = cache.assemble([("L", i) for i in order], "csr", nv, nv)
    for a, b in ((gotL, refL), (gotL.T, refL.T)):
        assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.colind, b.colind) and torch.equal(a.val, b.val)


# ------------------------------------------------------------------------------------------- oracle at the TIMED sizes
def _check_against_oracle(op, S, X, what, transposed=False):
    """op.apply(X) (or op.T.apply) against the double-precision oracle product of the COO operator S on the same input."""
    idx, val = S._indices().numpy(), S._values().numpy()
    r, c = (idx[1], idx[0]) if transposed else (idx[0], idx[1])
    n_out = S.shape[1] if transposed else S.shape[0]
    y = (op.T if transposed else op).apply(X)
    xh = X.cpu().numpy()
    if op.kind == "csr":
        y64, bound = c_oracle.coo_mm_f64(r, c, val, n_out, xh)
    else:
        y64, bound = c_oracle.dirac_view_mm_f64(r, c, val, n_out // 4, xh)
    return within_bound(y.cpu().numpy(), y64, bound, what)


def test_full_cfg3_size_vs_oracle():
    """BASELINE configs[2] at FULL size -- 64 distinct meshes x 2000 V, C = 128, the tensors bench.py times -- against the
    double-precision oracle: L, D, D*, and the transposed operators the backward pass applies (D^T, D*^T, L^T)."""
    from surfacenetworks_b200 import workloads as W
    O = ops_mod()
    meshes = W.make_mesh_ops(2000, range(64))
    b = W.arap_batch(meshes, 0)
    L = W.lap_batch(meshes)["L"]
    C = 128
    g = torch.Generator(device=DEV).manual_seed(3)
    for name, S, kind in (("L", L, "csr"), ("D", b["Di"], "bsr4"), ("D*", b["DiA"], "bsr4")):
        op = (O.as_csr if kind == "csr" else O.as_bsr4)(S.to(DEV))
        div = 1 if kind == "csr" else 4
        x_in = torch.randn(S.shape[1] // div, C, device=DEV, generator=g)
        x_out = torch.randn(S.shape[0] // div, C, device=DEV, generator=g)
        assert _check_against_oracle(op, S, x_in, name) < 32
        assert _check_against_oracle(op, S, x_out, name + "^T", transposed=True) < 32


@pytest.mark.parametrize("C", [16, 32, 64, 128, 256, 512])
def test_cfg5_faust_size_vs_oracle(C):
    """BASELINE configs[4]: one ~7000-vertex mesh, Dirac D / D* at every width of the 16-512 sweep bench.py times."""
    from surfacenetworks_b200 import workloads as W
    O = ops_mod()
    b = W.arap_batch(W.make_mesh_ops(7000, [0]), 0)
    g = torch.Generator(device=DEV).manual_seed(C)
    for name, S in (("D", b["Di"]), ("D*", b["DiA"])):
        op = O.as_bsr4(S.to(DEV))
        x = torch.randn(S.shape[1] // 4, C, device=DEV, generator=g)
        assert _check_against_oracle(op, S, x, "%s C=%d" % (name, C)) < 32


def test_cfg2_mesh_mnist_size_vs_oracle():
    """BASELINE configs[1]: 32 meshes x ~500 V, Laplacian at 128 features (and its transpose)."""
    from surfacenetworks_b200 import workloads as W
    O = ops_mod()
    L = W.lap_batch(W.make_mesh_ops(500, range(32)))["L"]
    op = O.as_csr(L.to(DEV))
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(L.shape[1], 128, device=DEV, generator=g)
    assert _check_against_oracle(op, L, x, "L") < 32
    assert _check_against_oracle(op, L, x, "L^T", transposed=True) < 32
