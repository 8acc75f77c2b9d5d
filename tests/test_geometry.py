"""Host-side operator construction vs matrices produced by the reference's own mesh.py / graph.py."""
import numpy as np
import pytest
import scipy.sparse as sp

from surfacenetworks_b200 import geometry


def _csr(row, col, val, shape):
    m = sp.csr_matrix((val, (row, col)), shape=shape)
    m.sort_indices()
    return m


@pytest.mark.parametrize("mesh", ["cube", "s60", "s45"])
def test_operators_match_reference(golden, mesh):
    d = golden("operators")
    V, F = d[mesh + "_V"], d[mesh + "_F"]
    L = geometry.build_laplacian(V, F)
    D, DA = geometry.build_dirac(V, F)
    for name, ours in (("L", L), ("Di", D), ("DiA", DA)):
        ref = _csr(*golden.coo("operators", "%s_%s" % (mesh, name)))
        ours = ours.tocsr()
        ours.sort_indices()
        assert ours.shape == ref.shape
        assert np.array_equal(ours.indptr, ref.indptr) and np.array_equal(ours.indices, ref.indices), name
        # same float64 arithmetic order as mesh.py -> identical float32 values
        assert np.array_equal(ours.data, ref.data), name


def test_cube_facts(golden):
    """cube L 8x8 / 44 nnz, Di 48x32 / 192 nnz, DiA 32x48 / 192 nnz."""
    V, F = geometry.cube_mesh()
    L = geometry.build_laplacian(V, F)
    D, DA = geometry.build_dirac(V, F)
    assert (L.shape, L.nnz) == ((8, 8), 44)
    assert (D.shape, D.nnz) == ((48, 32), 192)
    assert (DA.shape, DA.nnz) == ((32, 48), 192)
    assert abs(np.asarray(L.astype(np.float64).sum(1))).max() < 1e-4  # L 1 = 0


def test_dirac_structure():
    """Block = -Q(e)/(2A): zero diagonal, 12 stored nnz per block; DA(j,f) = D(f,j)^T A_f / A_v."""
    V, F = geometry.synth_mesh(80, 5)
    D, DA = geometry.build_dirac(V, F, dtype=np.float64)
    assert D.nnz == 36 * F.shape[0]
    Dd, DAd = D.toarray(), DA.toarray()
    Af = geometry.face_areas(V, F)
    Av = np.zeros(V.shape[0])
    np.add.at(Av, F.reshape(-1), np.repeat(Af / 3, 3))
    for f in (0, 7, F.shape[0] - 1):
        for j in F[f]:
            blk = Dd[4 * f:4 * f + 4, 4 * j:4 * j + 4]
            assert np.all(np.diag(blk) == 0)
            np.testing.assert_allclose(DAd[4 * j:4 * j + 4, 4 * f:4 * f + 4], blk.T * Af[f] / Av[j], rtol=1e-12)


def test_synth_mesh_deterministic_and_sized():
    V, F = geometry.synth_mesh(500, 3)
    V2, F2 = geometry.synth_mesh(500, 3)
    assert np.array_equal(V, V2) and np.array_equal(F, F2)
    assert V.shape == (500, 3) and 900 < F.shape[0] < 1000      # F ~ 2V - 2 - hull
    assert geometry.face_areas(V, F).min() > 1e-6
    L = geometry.build_laplacian(V, F)
    assert 6.0 < L.nnz / 500 < 7.5                               # ~6.9 nnz per row


def test_ply_reader(tmp_path):
    V, F = geometry.cube_mesh()
    p = tmp_path / "c.ply"
    lines = ["ply", "format ascii 1.0", "element vertex 8", "property float32 x", "property float32 y",
             "property float32 z", "element face 12", "property list uint8 int32 vertex_indices", "end_header"]
    lines += ["%g %g %g" % tuple(v) for v in V] + ["3 %d %d %d" % tuple(f) for f in F]
    p.write_text("\n".join(lines) + "\n")
    V2, F2 = geometry.read_ply_ascii(str(p))
    assert np.array_equal(V, V2) and np.array_equal(F, F2)


@pytest.mark.parametrize("method", ["bisect", "morton", "rcm"])
def test_locality_order_renumbers_the_same_mesh(method):
    """geometry.locality_order / reorder_mesh: both orders are permutations, the renumbered mesh is the same mesh, and
    its operators are the original ones with rows / columns permuted (D' = P_f D P_v^T blockwise, L' = P_v L P_v^T up to
    the summation order of duplicate contributions)."""
    V, F = geometry.synth_mesh(120, 3)
    vorder, forder = geometry.locality_order(V, F, method)
    assert sorted(vorder.tolist()) == list(range(V.shape[0])) and sorted(forder.tolist()) == list(range(F.shape[0]))
    V2, F2 = geometry.reorder_mesh(V, F, vorder, forder)
    assert np.array_equal(V2[F2], V[F[forder]])                     # every face keeps its corner positions, in order
    D, DA = geometry.build_dirac(V, F)
    D2, DA2 = geometry.build_dirac(V2, F2)
    rows = (4 * forder[:, None] + np.arange(4)).ravel()
    cols = (4 * vorder[:, None] + np.arange(4)).ravel()
    assert np.array_equal(D2.toarray(), D.toarray()[np.ix_(rows, cols)])
    # D* divides by vertex areas accumulated over the incident faces: their order changes, the value by a rounding
    np.testing.assert_allclose(DA2.toarray(), DA.toarray()[np.ix_(cols, rows)], rtol=1e-6, atol=0)
    L, L2 = geometry.build_laplacian(V, F), geometry.build_laplacian(V2, F2)
    ref = L.toarray()[np.ix_(vorder, vorder)]
    np.testing.assert_allclose(L2.toarray(), ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max())
    if method == "bisect":                                           # consecutive faces form patches: fewer distinct corners
        def span(Fx):
            return np.mean([len(np.unique(Fx[i:i + 32])) for i in range(0, len(Fx) - 32, 32)])
        assert span(F2) < span(F)
