"""Host-side checks that need no GPU: the C ABI library loads and exports every symbol the header declares,
argument validation, batching helpers vs the reference's outputs, module surface / state_dict layout."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO

HEADER = os.path.join(REPO, "include", "surfnet_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from surfacenetworks_b200 import _native as N
    syms = declared_symbols()
    assert len(syms) >= 10
    lib = ctypes.CDLL(N.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libsurfnet_b200.so does not export %s" % s
    assert set(N.SIGNATURES) == set(syms), "ctypes signature table and header disagree"
    assert N.version() == 100
    assert b"workspace" in N.lib.sn_status_string(N.SN_ERR_WORKSPACE)


def test_argument_validation_without_a_gpu():
    """Argument errors are detected on the host before any CUDA call."""
    from surfacenetworks_b200 import _native as N
    L = N.lib
    assert L.sn_csr_spmm_f32(0, 0, 0, 0, 4, 0, 4, 8, 4, 0, 0) == N.SN_ERR_ARG          # null pointers
    assert L.sn_csr_spmm_f32(0, 0, 0, 0, 4, 0, 4, -1, 4, 0, 0) == N.SN_ERR_ARG         # negative size
    assert L.sn_csr_spmm_f32(0, 0, 0, 0, 4, 0, 4, 0, 4, 0, 0) == N.SN_OK               # empty problem
    assert L.sn_bsr4_spmm_f32(1, 1, 1, 1, 6, 1, 6, 2, 6, 0, 0) == N.SN_ERR_UNSUPPORTED  # C % 4 != 0
    assert L.sn_bsr4_spmm_f32(1, 1, 1, 1, 4, 1, 8, 2, 8, 0, 0) == N.SN_ERR_ARG          # ldx < C
    assert L.sn_csr32_to_bsr4_count(1, 1, 6, 1, 1, 1 << 20, 0) == N.SN_ERR_UNSUPPORTED  # rows % 4 != 0
    assert L.sn_coo_to_csr32(0, 1, 1, 1, 5, 0, 0, 4, 4, 0, 1, 1, 1, 0, 0, 0) == N.SN_ERR_WORKSPACE
    assert L.sn_coo_to_csr32(0, 1, 1, 1, 1 << 33, 0, 0, 4, 4, 0, 1, 1, 1, 0, 0, 0) == N.SN_ERR_OVERFLOW
    assert L.sn_elu_f32(0, 4, 0, 4, 3, 4, 0) == N.SN_ERR_ARG
    assert L.sn_coo_to_csr32_ws_bytes(1000, 100) >= 4 * 1000 + 4 * 101
    with pytest.raises(N.SurfnetError) as e:
        N.call("sn_csr_spmm_f32", 0, 0, 0, 0, 4, 0, 4, 8, 4, 0, 0)
    assert e.value.status == N.SN_ERR_ARG and "invalid argument" in str(e.value)


def test_batching_helpers_match_reference(golden):
    """sparse_diag_cat / sparse_cat / sp_sparse_to_pt_sparse: identical indices and values to utils_pt.py:21-69."""
    import scipy.sparse as sp
    from surfacenetworks_b200 import utils_pt as U
    b = golden("batching")
    nv, nf = int(b["nv"]), int(b["nf"])

    def scipy_op(mesh, name):
        r, c, v, shape = golden.coo("operators", "%s_%s" % (mesh, name))
        return sp.coo_matrix((v, (r, c)), shape=shape)

    t = U.sp_sparse_to_pt_sparse(scipy_op("s45", "L"))
    assert np.array_equal(t._indices().numpy(), b["pt_L1_idx"]) and np.array_equal(t._values().numpy(), b["pt_L1_val"])
    assert t._values().dtype == torch.float32 and t._indices().dtype == torch.int64
    for name, s0, s1 in (("L", nv, nv), ("Di", 4 * nf, 4 * nv), ("DiA", 4 * nv, 4 * nf)):
        parts = [U.sp_sparse_to_pt_sparse(scipy_op(m, name)) for m in ("s60", "s45")]
        d = U.sparse_diag_cat(parts, s0, s1)
        assert d.is_coalesced() and tuple(d.shape) == (2 * s0, 2 * s1)
        assert np.array_equal(d._indices().numpy(), b["diag_%s_idx" % name])
        assert np.array_equal(d._values().numpy(), b["diag_%s_val" % name])
    for name, s0, s1 in (("L", nv, nv), ("Di", 4 * nf, 4 * nv)):
        parts = [U.sp_sparse_to_pt_sparse(scipy_op(m, name)) for m in ("s60", "s45")]
        c3 = U.sparse_cat(parts, s0, s1)
        assert tuple(c3.shape) == (2, s0, s1)
        assert np.array_equal(c3._indices().numpy(), b["cat_%s_idx" % name])
        assert np.array_equal(c3._values().numpy(), b["cat_%s_val" % name])
    dense = U.to_dense_batched(U.sp_sparse_to_pt_sparse(scipy_op("s45", "L")), 3)
    assert dense.shape == (3, 45, 45)


def test_workload_block_diag_equals_sparse_diag_cat():
    """bench.py's sort-free batch assembly gives exactly what sparse_diag_cat(...).coalesce() gives."""
    from surfacenetworks_b200 import utils_pt as U, workloads as W
    meshes = W.make_mesh_ops(40, [0, 1, 2])
    nv = max(m.num_vertices for m in meshes)
    nf = max(m.num_faces for m in meshes)
    batch = W.arap_batch(meshes, 0)
    ref = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(m.Di) for m in meshes], 4 * nf, 4 * nv)
    assert torch.equal(batch["Di"]._indices(), ref._indices()) and torch.equal(batch["Di"]._values(), ref._values())
    ref = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(m.DiA) for m in meshes], 4 * nv, 4 * nf)
    assert torch.equal(batch["DiA"]._indices(), ref._indices()) and torch.equal(batch["DiA"]._values(), ref._values())
    lb = W.lap_batch(meshes)
    ref = U.sparse_diag_cat([U.sp_sparse_to_pt_sparse(m.L) for m in meshes], nv, nv)
    assert torch.equal(lb["L"]._indices(), ref._indices()) and torch.equal(lb["L"]._values(), ref._values())
    assert batch["inputs"].shape == (3, nv, 6) and batch["targets"].shape == (3, nv, 120)


def test_module_surface_and_state_dict_layout():
    """Same constructor arguments and state_dict keys as the reference modules."""
    from surfacenetworks_b200 import models as M, utils_pt as U
    blk = U.DirResNet2(32, res_f=True)
    assert blk.res_f is True and blk.num_outputs == 32
    expect = ["bn_fc%d.bn.%s" % (i, k) for i in (0, 1)
              for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    expect += ["bn_fc%d.fc.%s" % (i, k) for i in (0, 1) for k in ("weight", "bias")]
    for cls in (U.LapResNet2, U.DirResNet2, U.AvgResNet2, U.DenseLapResNet2):
        sd = cls(32).state_dict()
        assert sorted(sd) == sorted(expect)
        assert sd["bn_fc0.fc.weight"].shape == (32, 64) and sd["bn_fc1.bn.weight"].shape == (64,)
    assert sorted(U.MlpResNet2(8).state_dict())[:2] == ["bn0.bn.bias", "bn0.bn.num_batches_tracked"]
    assert U.GraphConv1x1(4, 5, batch_norm=None).state_dict().keys() == {"fc.weight", "fc.bias"}
    for model, n in ((M.ArapDirModel(), 1018872), (M.ArapLapModel(15), 1018872)):
        assert sum(p.numel() for p in model.parameters()) == n
        assert "rn14.bn_fc1.fc.weight" in model.state_dict() and "conv2.bn.running_var" in model.state_dict()
    g = M.LapResNet2General(16, 32, inner_layers=3)
    assert g.state_dict()["bn_fc0.fc.weight"].shape == (32, 32) and g.state_dict()["bn_fc2.fc.weight"].shape == (32, 64)


def test_cpu_tensors_are_refused_loudly():
    """There is no CPU fallback in the product path."""
    from surfacenetworks_b200 import operators, ops, utils_pt as U
    S = torch.sparse_coo_tensor(torch.tensor([[0, 1], [1, 0]]), torch.tensor([1.0, 2.0]), (2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        operators.as_csr(S)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.elu_into(torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        U.LapResNet2(4)(S, None, torch.zeros(1, 2, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "surfacenetworks_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "sn_oracle" not in text, f


def test_new_entry_points_validate_arguments_without_a_gpu():
    """The round-1b entry points (row-group epilogue SpMM, mesh construction, GEMM flags) reject bad arguments on the
    host before any CUDA call."""
    from surfacenetworks_b200 import _native as N
    L = N.lib
    # epilogue SpMM: null operands / short leading dimensions / flags that select another kernel
    assert L.sn_bsr4_spmm_epilogue_f32(0, 0, 0, 0, 128, 0, 128, 8, 128, 0, 0, 0, 0, 0, 0, 0, 0) == N.SN_ERR_ARG
    assert L.sn_bsr4_spmm_epilogue_f32(16, 16, 16, 16, 128, 16, 128, 8, 128, 16, 64, 0, 0, 0, 0, 0, 0) == N.SN_ERR_ARG
    assert L.sn_bsr4_spmm_epilogue_f32(16, 16, 16, 16, 128, 16, 128, 8, 128, 0, 0, 0, 0, 0, 0,
                                       N.SN_SPMM_DIRECT_GATHER, 0) == N.SN_ERR_UNSUPPORTED
    assert L.sn_csr_spmm_epilogue_f32(16, 16, 16, 16, 48, 16, 48, 8, 48, 0, 0, 16, 48, 0, 0, 0, 0) == N.SN_ERR_UNSUPPORTED
    assert L.sn_csr_spmm_epilogue_f32(16, 16, 16, 16, 64, 16, 64, 0, 64, 0, 0, 0, 0, 0, 0, 0, 0) == N.SN_OK   # no rows
    # mesh construction: sizes, workspace
    assert L.sn_mesh_ws_bytes(0, 10, 10) == 0 and L.sn_mesh_ws_bytes(2, 100, 200) > 0
    assert L.sn_mesh_dirac_bsr4(0, 0, 2, 100, 200, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) == N.SN_ERR_ARG
    assert L.sn_mesh_dirac_bsr4(16, 16, 2, 100, 200, 16, 16, 16, 16, 16, 16, 0, 0, 0, 0, 16, 16, 8, 0) == N.SN_ERR_WORKSPACE
    assert L.sn_mesh_laplacian_csr(16, 16, 1 << 40, 100, 200, 16, 16, 16, 16, 16, 1 << 40, 0) == N.SN_ERR_OVERFLOW
    assert L.sn_mesh_laplacian_csr(16, 16, 0, 100, 200, 16, 16, 16, 16, 16, 0, 0) == N.SN_OK                 # no meshes
    # GEMM: the elu' epilogue needs the residual operand
    assert L.sn_gemm_tf32_f32(16, 128, 16, 128, 0, 0, 0, 0, 0, 0, 16, 256, 128, 256, 128, N.SN_GEMM_ELU_BWD_LEFT, 0, 0, 0) \
        == N.SN_ERR_ARG
    assert N.spmm_flags(True, False, True, 3) == (N.SN_SPMM_ELU_INPUT | N.SN_SPMM_SMEM_STREAM | (3 << 8))


def test_mesh_packing_and_reference_seam_surface():
    """pack_meshes pads like the batch operators expect; the src/utils/cuda mirror exposes the reference's names and
    refuses CPU tensors (no fallback)."""
    from surfacenetworks_b200 import cuda as C, geometry, operators as O
    V1, F1 = geometry.cube_mesh()
    V2, F2 = geometry.synth_mesh(30, 3)
    Vp, Fp = O.pack_meshes([(V1, F1), (V2, F2)], "cpu")
    assert Vp.shape == (2, 30, 3) and Vp.dtype == torch.float64 and Fp.dtype == torch.int32
    assert Fp.shape == (2, max(F1.shape[0], F2.shape[0]), 3)
    assert torch.all(Fp[0, F1.shape[0]:] == -1) and torch.all(Vp[0, 8:] == 0)
    assert np.array_equal(Fp[1, :F2.shape[0]].numpy(), F2) and np.array_equal(Vp[0, :8].numpy(), V1)
    with pytest.raises(RuntimeError):
        O.build_dirac_operators(Vp, Fp)                                  # CPU tensors: loud failure
    assert callable(C.batch_csr) and callable(C.sparse_bmm) and hasattr(C.SparseBMMFunc, "apply")
    idx = torch.zeros(3, 4, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        C.batch_csr(idx, (1, 2, 2))
    S = torch.sparse_coo_tensor(idx, torch.ones(4), (1, 2, 2))
    with pytest.raises(RuntimeError):
        C.SparseBMMFunc.apply(S, torch.zeros(1, 2, 3))


def test_arena_replays_allocations_and_structure_source_expands_blocks():
    """Host logic behind the allocation-free e2e path and the transposes of GPU-built operators (device-agnostic, so
    checked here on CPU tensors): Arena hands back the same storage in the same order after begin(); the BSR4 -> COO
    expansion follows the rotated block layout bval[16k + 4q + s] = B[(q + s) % 4][q]."""
    from surfacenetworks_b200 import operators as O
    ar = O.Arena()
    a1, b1 = ar.empty(10, torch.float32, "cpu"), ar.empty(7, torch.int32, "cpu")
    ar.begin()
    a2, b2 = ar.empty(10, torch.float32, "cpu"), ar.empty(5, torch.int32, "cpu")          # smaller request: a prefix view
    assert a2.data_ptr() == a1.data_ptr() and b2.data_ptr() == b1.data_ptr() and b2.numel() == 5
    ar.begin()
    c = ar.empty(10, torch.float64, "cpu")                                                 # sequence changed: re-recorded
    assert c.dtype == torch.float64 and c.data_ptr() != a1.data_ptr()
    ar.begin()
    assert ar.empty(10, torch.float64, "cpu").data_ptr() == c.data_ptr()
    # two block rows, three blocks: (0,1), (1,0), (1,2); dense random 4x4 blocks
    rng = np.random.default_rng(0)
    blocks = rng.standard_normal((3, 4, 4)).astype(np.float32)
    ptr = torch.tensor([0, 1, 3], dtype=torch.int32)
    ind = torch.tensor([1, 0, 2], dtype=torch.int32)
    bval = np.empty((3, 16), np.float32)
    for k in range(3):
        for q in range(4):
            for s_ in range(4):
                bval[k, 4 * q + s_] = blocks[k, (q + s_) % 4, q]
    src = O._StructureSource("bsr4", ptr, ind, torch.from_numpy(bval.ravel()), 8, 12, 3)
    row, col, val = src._coo()
    dense = np.zeros((8, 12), np.float32)
    dense[row.numpy(), col.numpy()] = val.numpy()
    expect = np.zeros((8, 12), np.float32)
    for k, (br, bc) in enumerate([(0, 1), (1, 0), (1, 2)]):
        expect[4 * br:4 * br + 4, 4 * bc:4 * bc + 4] = blocks[k]
    assert np.array_equal(dense, expect)
    t = src.transposed()
    assert (t.n_rows, t.n_cols) == (12, 8) and torch.equal(t.row, col) and torch.equal(t.col, row)
    csr = O._StructureSource("csr", torch.tensor([0, 2, 2, 3], dtype=torch.int32), torch.tensor([1, 3, 0], dtype=torch.int32),
                             torch.tensor([1.0, 2.0, 3.0]), 3, 4, 3)
    r, c_, v = csr._coo()
    assert r.tolist() == [0, 0, 2] and c_.tolist() == [1, 3, 0] and v.tolist() == [1.0, 2.0, 3.0]


def test_spmm_flag_word_and_row_length_hint():
    """The flags word of the SpMM entry points (include/surfnet_b200.h): variant in bits 8-11, row-length hint in 12-15;
    the hint of a block operator is its mean block count per row, rounded up."""
    from surfacenetworks_b200 import _native as N, operators as OP
    assert N.spmm_flags() == 0
    assert N.spmm_flags(elu_input=True, variant=9, row_entries=6) == (1 | (9 << 8) | (6 << 12))
    assert N.spmm_flags(row_entries=40) == 0 and N.spmm_flags(row_entries=-1) == 0          # out of range: no hint
    t = torch.zeros(0)
    assert OP.Bsr4Operator(t, t, t, 100, 50, n_blocks=300).row_entries_hint() == 3           # D: three blocks per face
    assert OP.Bsr4Operator(t, t, t, 100, 200, n_blocks=590).row_entries_hint() == 6          # D*: mean valence ~5.9
    assert OP.Bsr4Operator(t, t, t, 100, 200, n_blocks=0).row_entries_hint() == 0
    assert OP.Bsr4Operator(t, t, t, 2, 200, n_blocks=400).row_entries_hint() == 15           # capped to the 4-bit field


def test_bn_count_batch_host_side():
    """fused.bn_count_batch: the counter is handed to the fold kernel only when it lives on the GPU and a fixed momentum is
    set; otherwise (cumulative average, CPU buffers) it is incremented on the host like nn.BatchNorm does."""
    import torch.nn as nn
    from surfacenetworks_b200 import fused
    bn = nn.BatchNorm1d(4)
    assert fused.bn_count_batch(bn, training=False) is None and int(bn.num_batches_tracked) == 0
    assert fused.bn_count_batch(bn, training=True) is None and int(bn.num_batches_tracked) == 1     # CPU buffer: host add
    bn = nn.BatchNorm1d(4, momentum=None)
    assert fused.bn_count_batch(bn, training=True) is None and int(bn.num_batches_tracked) == 1
    assert fused.bn_momentum(bn) == 1.0
    bn = nn.BatchNorm1d(4, track_running_stats=False)
    assert fused.bn_count_batch(bn, training=True) is None
