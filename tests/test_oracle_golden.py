"""Pin the oracle (oracle/) against fixtures produced by the reference's own code (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from det import det_fill, det_tensor
from oracle import c_oracle, layers as O
from surfacenetworks_b200 import models as M
from surfacenetworks_b200 import utils_pt as U

EPS32 = float(np.finfo(np.float32).eps)


def params_of(module, seed, gain=1.0):
    """Deterministic parameters in the reference's state_dict layout (keys shared by both implementations)."""
    det_fill(module, seed, gain)
    P = {}
    for k, v in module.state_dict().items():
        v = v.clone()
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
        P[k] = v
    return P


def check_record(golden, tag, outs, inputs, P, weights, rtol=2e-5, atol=2e-5):
    d = golden("layers")
    loss = sum((o * w).sum() for o, w in zip(outs, weights))
    loss.backward()
    for i, o in enumerate(outs):
        assert_close(o.detach().numpy(), d["%s/out%d" % (tag, i)], rtol, atol, "%s out%d" % (tag, i))
    for i, t in enumerate(inputs):
        g = d["%s/gin%d" % (tag, i)]
        assert_close(t.grad.numpy(), g, rtol, atol * max(1.0, np.abs(g).max()), "%s gin%d" % (tag, i))
    for k, v in P.items():
        key = "%s/gparam.%s" % (tag, k)
        if key in d.files:
            g = d[key]
            assert_close(v.grad.numpy(), g, rtol, atol * max(1.0, np.abs(g).max()), key)
        key = "%s/buf.%s" % (tag, k)
        if key in d.files:
            assert_close(v.detach().numpy(), d[key], rtol, atol, key)


# ---------------------------------------------------------------------------------------------- plain-C oracle
def test_c_coo_mm_cube(golden):
    d = golden("spmm")
    row, col, val, shape = golden.coo("operators", "cube_L")
    y = c_oracle.coo_mm_f32(row, col, val, shape[0], d["cube_x"])
    y64, bound = c_oracle.coo_mm_f64(row, col, val, shape[0], d["cube_x"])
    assert np.all(np.abs(d["cube_Lx"] - y64) <= 32 * EPS32 * bound + 1e-30)   # the reference's own result obeys the bound
    assert np.all(np.abs(y - y64) <= 32 * EPS32 * bound + 1e-30)
    assert_close(y, d["cube_Lx"], 1e-6, 1e-6, "cube Lx")


@pytest.mark.parametrize("C", [4, 32, 40, 128])
def test_c_coo_and_dirac_view_batch(golden, C):
    d, b = golden("spmm"), golden("batching")
    nv, nf = int(b["nv"]), int(b["nf"])
    idx, val = b["diag_L_idx"], b["diag_L_val"]
    y64, bound = c_oracle.coo_mm_f64(idx[0], idx[1], val, 2 * nv, d["b_x%d" % C])
    assert np.all(np.abs(d["b_Lx%d" % C] - y64) <= 32 * EPS32 * bound + 1e-30)
    assert np.all(np.abs(c_oracle.coo_mm_f32(idx[0], idx[1], val, 2 * nv, d["b_x%d" % C]) - y64) <= 32 * EPS32 * bound + 1e-30)
    # transpose product (what backward applies)
    yT64, boundT = c_oracle.coo_mm_f64(idx[1], idx[0], val, 2 * nv, d["b_x%d" % C])
    assert np.all(np.abs(d["b_LTx%d" % C] - yT64) <= 32 * EPS32 * boundT + 1e-30)
    # Dirac view
    idx, val = b["diag_Di_idx"], b["diag_Di_val"]
    y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, 2 * nf, d["b_x%d" % C])
    assert np.all(np.abs(d["b_Dix%d" % C] - y64) <= 32 * EPS32 * bound + 1e-30)
    y32 = c_oracle.dirac_view_mm_f32(idx[0], idx[1], val, 2 * nf, d["b_x%d" % C])
    assert np.all(np.abs(y32 - y64) <= 32 * EPS32 * bound + 1e-30)
    idx, val = b["diag_DiA_idx"], b["diag_DiA_val"]
    y64, bound = c_oracle.dirac_view_mm_f64(idx[0], idx[1], val, 2 * nv, d["b_f%d" % C])
    assert np.all(np.abs(d["b_DiAf%d" % C] - y64) <= 32 * EPS32 * bound + 1e-30)
    # D^T through the view == transpose of the scalar matrix applied through the view
    idx, val = b["diag_Di_idx"], b["diag_Di_val"]
    y64, bound = c_oracle.dirac_view_mm_f64(idx[1], idx[0], val, 2 * nv, d["b_f%d" % C])
    assert np.all(np.abs(d["b_DiTf%d" % C] - y64) <= 32 * EPS32 * bound + 1e-30)


def test_c_batch_csr_and_sparse_bmm(golden):
    """Reference's native pair (batch_csr.cu + sparse_bmm.cu) restated; checked against torch.mm results."""
    d, b = golden("spmm"), golden("batching")
    nv = int(b["nv"])
    idx3, val = b["cat_L_idx"], b["cat_L_val"]
    col_ind, col_ptr = c_oracle.batch_csr(idx3, 2, nv)
    assert col_ptr[0, 0] == 0 and col_ptr[-1, -1] == val.size
    assert np.all(np.diff(col_ptr.reshape(-1)) >= 0)                      # monotone, incl. empty padded rows
    x = d["b_x32"].reshape(2, nv, 32)
    y = c_oracle.sparse_bmm(val, col_ind, col_ptr, 2, nv, x)
    assert_close(y.reshape(2 * nv, 32), d["b_Lx32"], 1e-5, 1e-3, "sparse_bmm vs torch.mm")
    # interior empty rows (the reference kernel's defect, batch_csr.cu:36-42): ranges must be empty, not 0
    idx = np.array([[0, 0, 1], [0, 3, 1], [1, 2, 0]], dtype=np.int64)
    ci, cp = c_oracle.batch_csr(idx, 2, 4)
    assert cp.tolist() == [[0, 1, 1, 1, 2], [2, 2, 3, 3, 3]]


def test_c_elu():
    x = np.concatenate([-np.logspace(-8, 1.5, 500), np.logspace(-8, 1.5, 500), [0.0]]).astype(np.float32)
    # libm expm1f vs torch's vectorised expm1: both within 1 ulp of the exact value
    assert_close(c_oracle.elu_f32(x), torch.nn.functional.elu(torch.from_numpy(x)).numpy(), 3 * EPS32, 0, "elu")


# ---------------------------------------------------------------------------------------------- batching helpers
def test_batching_restatement(golden):
    b = golden("batching")
    ops = golden("operators")
    nv, nf = int(b["nv"]), int(b["nf"])

    def pt(mesh, name):
        r, c, v, shape = golden.coo("operators", "%s_%s" % (mesh, name))
        return torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), shape)

    for name, s0, s1 in (("L", nv, nv), ("Di", 4 * nf, 4 * nv), ("DiA", 4 * nv, 4 * nf)):
        t = O.sparse_diag_cat([pt("s60", name), pt("s45", name)], s0, s1)
        assert np.array_equal(t._indices().numpy(), b["diag_%s_idx" % name])
        assert np.array_equal(t._values().numpy(), b["diag_%s_val" % name])
    t = O.sparse_cat([pt("s60", "L"), pt("s45", "L")], nv, nv)
    assert np.array_equal(t._indices().numpy(), b["cat_L_idx"]) and np.array_equal(t._values().numpy(), b["cat_L_val"])
    del ops


# ---------------------------------------------------------------------------------------------- layers
def test_layers_cube(golden):
    Lc = golden.pt_coo("batching", "pt_L1")  # just to exercise the loader
    del Lc
    r, c, v, shape = golden.coo("operators", "cube_L")
    L = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), shape).coalesce()
    x = det_tensor((1, 8, 16), 31).requires_grad_(True)
    P = params_of(U.LapResNet2(16), 1)
    check_record(golden, "cube_lap", [O.lap_resnet2(P, L, x)], [x], P, [det_tensor((1, 8, 16), 33)])

    r, c, v, shape = golden.coo("operators", "cube_Di")
    Di = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), shape).coalesce()
    r, c, v, shape = golden.coo("operators", "cube_DiA")
    DiA = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), shape).coalesce()
    x = det_tensor((1, 8, 16), 31).requires_grad_(True)
    f = det_tensor((1, 12, 16), 32).requires_grad_(True)
    P = params_of(U.DirResNet2(16), 2)
    check_record(golden, "cube_dir", list(O.dir_resnet2(P, Di, DiA, x, f)), [x, f], P,
                 [det_tensor((1, 8, 16), 34), det_tensor((1, 12, 16), 35)])


@pytest.fixture()
def batch(golden):
    b = golden("batching")
    nv, nf = int(b["nv"]), int(b["nf"])
    return dict(nv=nv, nf=nf, L=golden.pt_coo("batching", "diag_L"), Di=golden.pt_coo("batching", "diag_Di"),
                DiA=golden.pt_coo("batching", "diag_DiA"), mask=torch.from_numpy(golden("layers")["mask"]),
                x=lambda: det_tensor((2, nv, 32), 41).requires_grad_(True),
                f=lambda: det_tensor((2, nf, 32), 42).requires_grad_(True),
                wv=det_tensor((2, nv, 32), 43), wf=det_tensor((2, nf, 32), 44))


def test_layers_batch(golden, batch):
    B = batch
    x = B["x"]()
    P = params_of(U.LapResNet2(32), 3)
    check_record(golden, "b_lap", [O.lap_resnet2(P, B["L"], x)], [x], P, [B["wv"]])

    x, f = B["x"](), B["f"]()
    P = params_of(U.DirResNet2(32), 4)
    check_record(golden, "b_dir", list(O.dir_resnet2(P, B["Di"], B["DiA"], x, f)), [x, f], P, [B["wv"], B["wf"]])

    x = B["x"]()
    P = params_of(U.AvgResNet2(32), 5)
    check_record(golden, "b_avg", [O.avg_resnet2(P, B["mask"], x)], [x], P, [B["wv"]])

    x = B["x"]()
    P = params_of(U.MlpResNet2(32), 6)
    check_record(golden, "b_mlp", [O.mlp_resnet2(P, x)], [x], P, [B["wv"]])

    x = B["x"]()
    Ld = B["L"].to_dense()
    nv = B["nv"]
    Ld = torch.stack([Ld[i * nv:(i + 1) * nv, i * nv:(i + 1) * nv] for i in range(2)])
    P = params_of(U.DenseLapResNet2(32), 3)
    check_record(golden, "b_denselap", [O.lap_resnet2(P, Ld, x)], [x], P, [B["wv"]])

    for bn in (None, "pre", "post"):
        x = B["x"]()
        P = params_of(U.GraphConv1x1(32, 24, batch_norm=bn), 7)
        check_record(golden, "b_conv_%s" % bn, [O.graph_conv1x1(x, P, bn)], [x], P, [det_tensor((2, nv, 24), 45)])

    x = B["x"]()
    P = params_of(U.GraphBatchNorm(32), 8)
    check_record(golden, "b_gbn", [O.graph_batch_norm(x, P)], [x], P, [B["wv"]])
    assert_close(O.global_average(det_tensor((2, nv, 32), 41), B["mask"]).numpy(), golden("layers")["b_global_average"],
                 1e-6, 1e-6, "global_average")


def test_layers_eval_mode(golden, batch):
    B, d = batch, golden("layers")
    with torch.no_grad():
        P = params_of(U.LapResNet2(32), 3)
        assert_close(O.lap_resnet2(P, B["L"], B["x"](), training=False).numpy(), d["b_lap_eval/out0"], 2e-5, 2e-5, "lap eval")
        P = params_of(U.DirResNet2(32), 4)
        v, f = O.dir_resnet2(P, B["Di"], B["DiA"], B["x"](), B["f"](), training=False)
        assert_close(v.numpy(), d["b_dir_eval/out0"], 2e-5, 2e-5, "dir eval v")
        assert_close(f.numpy(), d["b_dir_eval/out1"], 2e-5, 2e-5, "dir eval f")


@pytest.mark.parametrize("tag", ["dir", "lap"])
def test_arap_models(golden, batch, tag):
    B, d = batch, golden("arap_models")
    inputs, targets = torch.from_numpy(d["inputs"]), torch.from_numpy(d["targets"])
    if tag == "dir":
        P = params_of(M.ArapDirModel(), 9, 0.25)
        out = O.arap_dir_model(P, B["Di"], B["DiA"], B["mask"], inputs)
    else:
        P = params_of(M.ArapLapModel(15), 10, 0.25)
        out = O.arap_lap_model(P, B["L"], B["mask"], inputs)
    loss = O.arap_loss(out, targets, B["mask"], 2)
    loss.backward()
    assert_close(out.detach().numpy(), d[tag + "/out"], 1e-4, 1e-4, tag + " out")
    assert_close(loss.item(), d[tag + "/loss"], 1e-5, 0, tag + " loss")
    for k in d.files:
        if k.startswith(tag + "/g."):
            g = d[k]
            assert_close(P[k[len(tag) + 3:]].grad.numpy(), g, 1e-3, 1e-4 * np.abs(g).max(), k)


def test_other_callers(golden, batch):
    """normal_predict _LapResNet2 / DirDeepModel and the mesh_mnist LapEncoder."""
    B, d = batch, golden("callers")
    x32 = det_tensor((2, B["nv"], 32), 62)
    x3 = torch.from_numpy(d["x3"])
    with torch.no_grad():
        for tag, mk, seed, inner in (("lapgen_32_64_3", lambda: M.LapResNet2General(32, 64, inner_layers=3), 11, 3),
                                     ("lapgen_32", lambda: M.LapResNet2General(32), 12, 2),
                                     ("lapgen_32_16_1", lambda: M.LapResNet2General(32, 16, inner_layers=1), 15, 1)):
            P = params_of(mk(), seed, 0.5)
            assert_close(O.lap_resnet2_general(P, B["L"], x32, inner).numpy(), d[tag + "/out0"], 2e-5, 2e-5, tag)
        P = params_of(M.DirDeepModel(3, 1, layers=4), 13, 0.25)
        assert_close(O.dir_deep_model(P, B["Di"], B["DiA"], B["mask"], x3, 4).numpy(), d["dirdeep4/out0"], 2e-5, 2e-5, "dirdeep4")
        P = params_of(M.LapEncoder(), 14, 0.25)
        mu, lv = O.lap_encoder(P, x3, B["L"], B["mask"])
        assert_close(mu.numpy(), d["lapencoder/out0"], 2e-5, 2e-5, "lapencoder mu")
        assert_close(lv.numpy(), d["lapencoder/out1"], 2e-5, 2e-5, "lapencoder logvar")


def test_dense_correspondence_siamese(golden, batch):
    """dense_correspondence Model(5) and SiameseModel('lap', 3) (models.py:21-48,184-203): oracle vs the reference."""
    B, d = batch, golden("siamese")
    xa, xb = torch.from_numpy(d["xa"]), torch.from_numpy(d["xb"])
    with torch.no_grad():
        P = params_of(M.DcLapModel(5), 17, 0.25)
        assert_close(O.arap_lap_model(P, B["L"], B["mask"], xa, 5).numpy(), d["dclap5/out0"], 2e-5, 2e-5, "dc Model(5)")
        P = params_of(M.SiameseModel("lap", 3), 16, 0.25)
        out = O.siamese(P, (B["L"], B["mask"]), (B["L"], B["mask"]), xa, xb, 3, "lap")
        scale = float(np.abs(d["siamese_lap3/out0"]).max())
        assert_close(out.numpy(), d["siamese_lap3/out0"], 1e-4, 1e-5 * scale, "siamese lap3")


def test_mesh_operator_restatement_matches_reference_built_operators(golden):
    """oracle/mesh_ops.py (the per-entry arithmetic the GPU construction kernels replay) reproduces, bit for bit, the
    operators the reference's mesh.py / graph.py built for cube.ply (tests/golden/operators.npz), and the host builder on a
    generic synthetic mesh."""
    from oracle import mesh_ops
    from surfacenetworks_b200 import geometry
    V, F = geometry.cube_mesh()
    r, c, v, shape = golden.coo("operators", "cube_L")
    rr, cc, vv = mesh_ops.laplacian_coo(V, F)
    assert sorted(zip(r.tolist(), c.tolist(), v.tolist())) == sorted(zip(rr.tolist(), cc.tolist(), vv.tolist()))
    D, DA = mesh_ops.dirac_entries(V, F)
    for name, got in (("cube_Di", D), ("cube_DiA", DA)):
        r, c, v, shape = golden.coo("operators", name)
        ref = {(int(a), int(b)): np.float32(x) for a, b, x in zip(r, c, v)}
        assert ref == got, name
    V, F = geometry.synth_mesh(120, 5)
    L = geometry.build_laplacian(V, F).tocoo()
    rr, cc, vv = mesh_ops.laplacian_coo(V, F)
    assert sorted(zip(L.row.tolist(), L.col.tolist(), L.data.tolist())) == sorted(zip(rr.tolist(), cc.tolist(), vv.tolist()))
    Dh, DAh = [m.tocoo() for m in geometry.build_dirac(V, F)]
    D, DA = mesh_ops.dirac_entries(V, F)
    for S, got in ((Dh, D), (DAh, DA)):
        assert {(int(a), int(b)): np.float32(x) for a, b, x in zip(S.row, S.col, S.data)} == got


def test_cpu_arm_workload_is_pinned(golden):
    """oracle/workload.py (what bench.py's CPU arms run on, independent of the product package): its Dirac operators
    equal the reference-built cube operators and the per-entry restatement; meshes, batch operators and the parameter
    layout equal the product's host-side builders, so both bench arms see the same workload."""
    from oracle import mesh_ops, workload as W
    from surfacenetworks_b200 import geometry, workloads as PW

    def as_dict(m):
        m = m.tocoo()
        return {(int(r), int(c)): np.float32(v) for r, c, v in zip(m.row, m.col, m.data)}

    V, F = geometry.cube_mesh()
    D, DA = W.dirac_operators(V, F)
    for name, m in (("cube_Di", D), ("cube_DiA", DA)):
        r, c, v, shape = golden.coo("operators", name)
        assert m.shape == shape
        assert as_dict(m) == {(int(a), int(b)): np.float32(x) for a, b, x in zip(r, c, v)}, name
    V, F = W.synth_mesh(120, 5)
    Vp, Fp = geometry.synth_mesh(120, 5)
    assert np.array_equal(V, Vp) and np.array_equal(F, Fp)
    D, DA = W.dirac_operators(V, F)
    De, DAe = mesh_ops.dirac_entries(V, F)
    assert as_dict(D) == De and as_dict(DA) == DAe
    # batch: same coalesced block-diagonal operators, inputs, targets and mask as the product's workload builder
    meshes = [W.synth_mesh(n, s) for n, s in ((90, 1), (70, 2), (90, 3))]
    b = W.arap_batch(meshes, seed=4)
    pb = PW.arap_batch([PW.MeshOps(v, f) for v, f in meshes], seed=4)
    for k in ("inputs", "targets", "mask"):
        assert torch.equal(b[k], pb[k]), k
    for k in ("Di", "DiA"):
        assert b[k].shape == pb[k].shape
        assert torch.equal(b[k]._indices(), pb[k]._indices()) and torch.equal(b[k]._values(), pb[k]._values()), k
    # parameter dictionary: the reference DirModel's state_dict layout (= the product model's)
    from surfacenetworks_b200 import models as M
    sd = M.ArapDirModel().state_dict()
    P = W.arap_dir_params(0)
    assert list(P.keys()) == list(sd.keys())
    assert all(P[k].shape == sd[k].shape and P[k].dtype == sd[k].dtype for k in sd)
