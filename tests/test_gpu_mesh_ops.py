"""GPU operator construction (sn_mesh_dirac_bsr4 / sn_mesh_laplacian_csr) against operators built by
the reference's own code (tests/golden/operators.npz: mesh.dirac, mesh.cotangent_weights, graph.laplacian on cube.ply)
and against the host builder (geometry.py, itself pinned to the reference in tests/test_geometry.py) on ragged batches.

Bar: identical sparsity structure; values within 1 fp32 ulp (the kernels evaluate the geometry in fp64 in the
reference's operation order and round once, so they are expected -- and reported -- to be bit-identical)."""
import numpy as np
import pytest
import torch

from det import det_array

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops_mod():
    from surfacenetworks_b200 import operators
    return operators


def same_values(a, b, what):
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    tol = np.spacing(np.abs(b).astype(np.float32)).astype(np.float64)        # 1 ulp of the expected value
    err = np.abs(a.astype(np.float64) - b.astype(np.float64))
    assert np.all(err <= tol), "%s: %d values differ by more than 1 ulp (max %g)" % (what, int((err > tol).sum()), err.max())
    return float((a == b).mean())


def check_bsr4(got, ref, what):
    assert (got.n_brows, got.n_bcols, got.n_blocks) == (ref.n_brows, ref.n_bcols, ref.n_blocks), what
    assert torch.equal(got.browptr, ref.browptr), what + " row pointers"
    assert torch.equal(got.bcolind[:got.n_blocks], ref.bcolind[:ref.n_blocks]), what + " block columns"
    return same_values(got.bval[:16 * got.n_blocks], ref.bval[:16 * ref.n_blocks], what + " values")


def check_csr(got, ref, what):
    assert (got.n_rows, got.n_cols, got.nnz) == (ref.n_rows, ref.n_cols, ref.nnz), (what, got.nnz, ref.nnz)
    assert torch.equal(got.rowptr, ref.rowptr), what + " row pointers"
    assert torch.equal(got.colind[:got.nnz], ref.colind[:ref.nnz]), what + " columns"
    return same_values(got.val[:got.nnz], ref.val[:ref.nnz], what + " values")


def test_cube_matches_reference_built_operators(golden):
    """BASELINE cfg1 mesh: the operators the reference's mesh.py / graph.py produce for cube.ply."""
    from surfacenetworks_b200 import geometry
    O = ops_mod()
    V, F = geometry.cube_mesh()
    Vg, Fg = O.pack_meshes([(V, F)], DEV)
    D, DA = O.build_dirac_operators(Vg, Fg)
    L = O.build_laplacian_operator(Vg, Fg)
    ref = {}
    for name in ("cube_L", "cube_Di", "cube_DiA"):
        r, c, v, shape = golden.coo("operators", name)
        ref[name] = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), shape).coalesce().to(DEV)
    exact = [check_bsr4(D, O.Bsr4Operator.from_torch_coo(ref["cube_Di"]), "cube D"),
             check_bsr4(DA, O.Bsr4Operator.from_torch_coo(ref["cube_DiA"]), "cube D*"),
             check_csr(L, O.CsrOperator.from_torch_coo(ref["cube_L"]), "cube L")]
    print("bit-identical fractions (D, D*, L):", exact)


@pytest.mark.parametrize("sizes", [(60, 45, 60), (500, 480, 512, 300)])
def test_ragged_batch_matches_host_builder(sizes):
    """Ragged batches (padding vertices and faces): GPU-built batch operators == host-built + sparse_diag_cat +
    conversion, including the transposed structures used by backward."""
    from surfacenetworks_b200 import geometry, utils_pt as U
    O = ops_mod()
    meshes = [geometry.synth_mesh(n, 10 + i) for i, n in enumerate(sizes)]
    nv = max(v.shape[0] for v, _ in meshes)
    nf = max(f.shape[0] for _, f in meshes)
    Vg, Fg = O.pack_meshes(meshes, DEV)
    D, DA = O.build_dirac_operators(Vg, Fg)
    L = O.build_laplacian_operator(Vg, Fg)
    DD = [geometry.build_dirac(v, f) for v, f in meshes]
    Dh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(d[0]) for d in DD], 4 * nf, 4 * nv).to(DEV)
    DAh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(d[1]) for d in DD], 4 * nv, 4 * nf).to(DEV)
    Lh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(geometry.build_laplacian(v, f)) for v, f in meshes], nv, nv).to(DEV)
    Dr, DAr, Lr = O.Bsr4Operator.from_torch_coo(Dh), O.Bsr4Operator.from_torch_coo(DAh), O.CsrOperator.from_torch_coo(Lh)
    exact = [check_bsr4(D, Dr, "D"), check_bsr4(DA, DAr, "D*"), check_csr(L, Lr, "L")]
    print("bit-identical fractions (D, D*, L):", exact)
    assert min(exact) > 0.999
    # transposes built from the stored structure == transposes built from the COO source
    check_bsr4(D.T, Dr.T, "D^T")
    check_bsr4(DA.T, DAr.T, "D*^T")
    check_csr(L.T, Lr.T, "L^T")
    # stream-ordered construction (no read-backs): same arrays, capacity-sized block count
    D2, DA2 = O.build_dirac_operators(Vg, Fg, sync=False)
    nb = D.n_blocks
    assert int(D2.status.item()) == 0 and D2.n_blocks == 3 * len(sizes) * nf
    for got, ref in ((D2, D), (DA2, DA), (D2.T, D.T), (DA2.T, DA.T)):
        assert torch.equal(got.browptr, ref.browptr) and torch.equal(got.bcolind[:nb], ref.bcolind[:nb])
        assert torch.equal(got.bval[:16 * nb], ref.bval[:16 * nb])
    # and the operators drive the layers: one DirResNet2 forward on GPU-built vs host-built operators
    from det import det_fill
    C = 32
    v = torch.from_numpy(det_array((len(sizes), nv, C), 1)).to(DEV)
    f = torch.from_numpy(det_array((len(sizes), nf, C), 2)).to(DEV)
    out_g = det_fill(U.DirResNet2(C), 5).to(DEV)(D, DA, v, f)
    out_h = det_fill(U.DirResNet2(C), 5).to(DEV)(Dh, DAh, v, f)
    assert torch.allclose(out_g[0], out_h[0], rtol=1e-5, atol=1e-5) and torch.allclose(out_g[1], out_h[1], rtol=1e-5, atol=1e-5)


def test_full_size_batch_and_errors():
    """BASELINE cfg3 size (64 x 2000 V): structure invariants instead of a host rebuild -- every real face row of D has
    3 blocks, D* has as many blocks as D, L rows sum to ~0 (L 1 = 0), and D* == D^T scaled (adjoint identity)."""
    from surfacenetworks_b200 import geometry
    O = ops_mod()
    base = [geometry.synth_mesh(2000, s) for s in range(4)]
    meshes = [base[i % 4] for i in range(64)]
    Vg, Fg = O.pack_meshes(meshes, DEV)
    D, DA = O.build_dirac_operators(Vg, Fg)
    L = O.build_laplacian_operator(Vg, Fg)
    n_faces = sum(f.shape[0] for _, f in meshes)
    assert D.n_blocks == DA.n_blocks == 3 * n_faces
    cnt = (D.browptr[1:] - D.browptr[:-1])
    assert set(cnt.unique().tolist()) <= {0, 3}
    ones = torch.ones(L.n_cols, 16, device=DEV)
    Labs = O.CsrOperator(L.rowptr, L.colind, L.val.abs(), L.n_rows, L.n_cols)
    assert torch.all(L.apply(ones).abs() <= 64 * 1.2e-7 * Labs.apply(ones) + 1e-30)
    # padded batches repeat 4 distinct meshes: mesh i and mesh i + 4 get identical blocks
    nb0 = int(D.browptr[4 * Fg.shape[1]].item())
    assert torch.equal(D.bval[:16 * nb0], D.bval[16 * nb0:32 * nb0])
    # error paths: CPU tensors refuse, malformed shapes raise
    with pytest.raises(RuntimeError):
        O.build_dirac_operators(Vg.cpu(), Fg.cpu())
    with pytest.raises(ValueError):
        O.build_laplacian_operator(Vg[:, :, :2], Fg)


@pytest.mark.parametrize("n_fan", [70, 300])
def test_high_valence_vertex_matches_host_builder(n_fan):
    """A vertex in more than 64 faces (pole of a UV sphere, cone apex): the reference handles any valence (mesh.py:35-64,
    102-112 are dense O(V^2) numpy), so do the kernels -- such vertices take the same code over global scratch.  Operators
    must equal the host builder's (itself pinned to the reference in tests/test_geometry.py), status reports the valence."""
    from surfacenetworks_b200 import geometry, utils_pt as U
    O = ops_mod()
    rng = np.random.default_rng(n_fan)
    ang = np.sort(rng.random(n_fan + 1)) * 1.9 * np.pi
    ring = np.stack([np.cos(ang), np.sin(ang), 0.2 * rng.random(n_fan + 1)], 1) * (1.0 + 0.3 * rng.random((n_fan + 1, 1)))
    Vf = np.concatenate([np.array([[0.0, 0.0, 0.5]]), ring])                        # apex 0 + an open fan around it
    fan = np.array([[0, i + 1, i + 2] for i in range(n_fan)], dtype=np.int64)       # vertex 0 in n_fan faces
    small = geometry.synth_mesh(50, 3)
    meshes = [(Vf, fan), small]
    nv = max(v.shape[0] for v, _ in meshes)
    nf = max(f.shape[0] for _, f in meshes)
    Vg, Fg = O.pack_meshes(meshes, DEV)
    D, DA = O.build_dirac_operators(Vg, Fg)
    L = O.build_laplacian_operator(Vg, Fg)
    assert int(D.status.item()) == n_fan                                             # informational: largest valence > 64
    DD = [geometry.build_dirac(v, f) for v, f in meshes]
    Dh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(d[0]) for d in DD], 4 * nf, 4 * nv).to(DEV)
    DAh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(d[1]) for d in DD], 4 * nv, 4 * nf).to(DEV)
    Lh = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(geometry.build_laplacian(v, f)) for v, f in meshes], nv, nv).to(DEV)
    Dr, DAr, Lr = O.Bsr4Operator.from_torch_coo(Dh), O.Bsr4Operator.from_torch_coo(DAh), O.CsrOperator.from_torch_coo(Lh)
    check_bsr4(D, Dr, "D")
    check_bsr4(DA, DAr, "D*")
    check_bsr4(D.T, Dr.T, "D^T")
    check_bsr4(DA.T, DAr.T, "D*^T")
    check_csr(L, Lr, "L")
    # stream-ordered construction (no read-back) is memory-safe and identical for the high-valence rows as well
    D2, DA2 = O.build_dirac_operators(Vg, Fg, sync=False)
    nb = D.n_blocks
    assert torch.equal(DA2.browptr, DA.browptr) and torch.equal(DA2.bcolind[:nb], DA.bcolind[:nb])
    assert torch.equal(DA2.bval[:16 * nb], DA.bval[:16 * nb])
