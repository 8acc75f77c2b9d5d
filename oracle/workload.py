"""Self-contained synthetic as_rigid_as_possible workload for the CPU arms of bench.py.  TEST INFRASTRUCTURE.

``bench.py --impl reference`` and the ``cpu_baseline`` leg must not depend on the product package (a reference arm that
imports ``surfacenetworks_b200`` ends up with the product's shared object in its process).  Everything those legs need
besides ``oracle/layers.py`` lives here, in numpy / scipy / torch-CPU only:

  synth_mesh          the synthetic height-field mesh of SURVEY.md 8(d) (same recipe and seeds as the product's
                      geometry.synth_mesh: uniform 2-D points -> Delaunay -> z = 0.3 U[0,1], minimum-area rejection,
                      cf. reference src/mesh_mnist/create_data.py:62-99)
  dirac_operators     D / D* of one mesh, vectorised: block(f, j) = -Q(0, V[j+1] - V[j+2]) / (2 A_f),
                      block*(j, f) = block(f, j)^T A_f / A_v[j]      (reference src/utils/mesh.py:28-64)
  arap_batch          what the reference's sample_batch hands to the model (src/as_rigid_as_possible/main.py:98-185):
                      zero-padded inputs / targets / mask and the block-diagonal COO operators
  reference_assembly  the reference's per-step batch assembly, utils_pt.py:41-53: offset, concatenate, ``.coalesce()``
  arap_dir_params     an initial parameter dictionary with the reference DirModel's state_dict layout
                      (src/as_rigid_as_possible/models.py:108-126) and torch's default Linear / BatchNorm initialisation

Pinned by tests/test_oracle_golden.py: operators equal to the per-entry restatement (oracle/mesh_ops.py, itself pinned
to the reference-built cube operators) and to the product's host builder on the same meshes.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy import sparse

__all__ = ["synth_mesh", "face_areas", "dirac_operators", "arap_batch", "reference_assembly", "arap_dir_params"]


def _sqdist(d):
    return d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]      # mesh.py:24, left to right


def face_areas(V, F):
    """Heron areas with the 1e-6 floor of mesh.py:67-80."""
    v0, v1, v2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    a, b, c = np.sqrt(_sqdist(v0 - v1)), np.sqrt(_sqdist(v1 - v2)), np.sqrt(_sqdist(v2 - v0))
    s = (a + b + c) / 2
    prod = s * (s - a) * (s - b) * (s - c)
    out = np.full(F.shape[0], 1e-6)
    good = prod > 0
    out[good] = np.sqrt(prod[good])
    return out


def synth_mesh(num_vertices, seed, min_area=1e-6):
    from scipy.spatial import Delaunay
    attempt = 0
    while True:
        rng = np.random.default_rng([int(seed), attempt])
        P = rng.random((num_vertices, 2))
        z = 0.3 * rng.random(num_vertices)
        tri = Delaunay(P)
        V = np.concatenate([tri.points, z[:, None]], axis=1).astype(np.float64)
        F = np.asarray(tri.simplices, dtype=np.int64)
        if face_areas(V, F).min() > min_area:
            return V, F
        attempt += 1


def dirac_operators(V, F):
    """(D [4F x 4V], D* [4V x 4F]) as scipy CSR float32 with sorted rows (mesh.py:35-64, then .astype('float32'))."""
    nf, nv = F.shape[0], V.shape[0]
    Af = face_areas(V, F)
    Av = np.zeros(nv)
    np.add.at(Av, F.reshape(-1), np.repeat(Af / 3, 3))                                # mesh.py:44-45
    f = np.repeat(np.arange(nf), 3)
    c = np.tile(np.arange(3), nf)
    j = F[f, c]
    e = V[F[f, (c + 1) % 3]] - V[F[f, (c + 2) % 3]]                                   # mesh.py:49-51
    z = np.zeros(e.shape[0])
    b, cc, d = e[:, 0], e[:, 1], e[:, 2]
    Q = np.stack([np.stack([z, -b, -cc, -d], 1), np.stack([b, z, -d, cc], 1),
                  np.stack([cc, d, z, -b], 1), np.stack([d, -cc, b, z], 1)], 1)       # mesh.py:28-33
    mat = -Q / (2 * Af[f])[:, None, None]                                            # mesh.py:57
    matA = np.transpose(mat, (0, 2, 1)) * Af[f][:, None, None] / Av[j][:, None, None]   # mesh.py:59
    p = np.arange(4)
    rows = 4 * f[:, None, None] + p[None, :, None] + 0 * p[None, None, :]
    cols = 4 * j[:, None, None] + p[None, None, :] + 0 * p[None, :, None]
    nz, nzA = mat != 0, matA != 0
    D = sparse.csr_matrix((mat[nz], (rows[nz], cols[nz])), shape=(4 * nf, 4 * nv)).astype(np.float32)
    rowsA = 4 * j[:, None, None] + p[None, :, None] + 0 * p[None, None, :]
    colsA = 4 * f[:, None, None] + p[None, None, :] + 0 * p[None, :, None]
    DA = sparse.csr_matrix((matA[nzA], (rowsA[nzA], colsA[nzA])), shape=(4 * nv, 4 * nf)).astype(np.float32)
    D.sort_indices()
    DA.sort_indices()
    return D, DA


def _torch_coo(m):
    """scipy -> torch sparse COO the way sp_sparse_to_pt_sparse does (utils_pt.py:56-69): int64 indices, uncoalesced."""
    m = m.tocoo()
    idx = torch.from_numpy(np.stack([m.row, m.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data.astype(np.float32)), m.shape)


def reference_assembly(mats, size0, size1):
    """sparse_diag_cat, utils_pt.py:41-53: per-mesh COO tensors offset to their diagonal block, concatenated, coalesced
    (a sort of all nnz -- the reference does this on the host every training step, main.py:172-177)."""
    idx, val = [], []
    for i, t in enumerate(mats):
        ii = t._indices().clone()
        ii[0] += i * size0
        ii[1] += i * size1
        idx.append(ii)
        val.append(t._values())
    return torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val), (len(mats) * size0, len(mats) * size1)).coalesce()


def arap_batch(meshes, seed=0):
    """meshes: list of (V, F).  Returns the batch dictionary plus the per-mesh torch COO operators (for timing the
    reference-style assembly)."""
    B = len(meshes)
    nv = max(v.shape[0] for v, _ in meshes)
    nf = max(f.shape[0] for _, f in meshes)
    rng = np.random.default_rng(seed)
    inputs = np.zeros((B, nv, 6), dtype=np.float32)
    targets = np.zeros((B, nv, 120), dtype=np.float32)
    mask = np.zeros((B, nv, 1), dtype=np.float32)
    per_mesh = []
    for b, (V, F) in enumerate(meshes):
        n = V.shape[0]
        frame0 = V.astype(np.float32)
        vel = 0.01 * rng.standard_normal((n, 3)).astype(np.float32)
        inputs[b, :n, :3] = frame0
        inputs[b, :n, 3:] = frame0 + vel
        steps = np.arange(2, 42, dtype=np.float32)[None, :, None]
        targets[b, :n] = (frame0[:, None, :] + steps * vel[:, None, :]).reshape(n, 120)
        mask[b, :n] = 1
        D, DA = dirac_operators(V, F)
        per_mesh.append((_torch_coo(D), _torch_coo(DA)))
    Di = reference_assembly([d for d, _ in per_mesh], 4 * nf, 4 * nv)
    DiA = reference_assembly([a for _, a in per_mesh], 4 * nv, 4 * nf)
    return {"inputs": torch.from_numpy(inputs), "targets": torch.from_numpy(targets), "mask": torch.from_numpy(mask),
            "Di": Di, "DiA": DiA, "per_mesh": per_mesh, "num_vertices": nv, "num_faces": nf, "batch_size": B}


def _linear(P, prefix, n_in, n_out, gen):
    bound = 1.0 / math.sqrt(n_in)                      # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    P[prefix + "weight"] = (torch.rand(n_out, n_in, generator=gen) * 2 - 1) * bound
    P[prefix + "bias"] = (torch.rand(n_out, generator=gen) * 2 - 1) * bound


def _batch_norm(P, prefix, n):
    P[prefix + "weight"] = torch.ones(n)
    P[prefix + "bias"] = torch.zeros(n)
    P[prefix + "running_mean"] = torch.zeros(n)
    P[prefix + "running_var"] = torch.ones(n)
    P[prefix + "num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def arap_dir_params(seed=0, width=128, layers=15):
    """state_dict layout of the reference DirModel (models.py:108-126): conv1 6 -> C, rn0..rn14 with two
    GraphConv1x1(2C, C, "pre") each, conv2 C -> 120 with BatchNorm "pre"."""
    gen = torch.Generator().manual_seed(seed)
    P = {}
    _linear(P, "conv1.fc.", 6, width, gen)
    for i in range(layers):
        for s in (0, 1):
            _batch_norm(P, "rn%d.bn_fc%d.bn." % (i, s), 2 * width)
            _linear(P, "rn%d.bn_fc%d.fc." % (i, s), 2 * width, width, gen)
    _batch_norm(P, "conv2.bn.", width)
    _linear(P, "conv2.fc.", width, 120, gen)
    return P
