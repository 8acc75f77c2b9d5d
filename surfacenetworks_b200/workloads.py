"""Synthetic batches of the BASELINE.json configurations, built on the host.

What the reference's ``sample_batch`` produces each step (src/as_rigid_as_possible/main.py:98-185): zero-padded
``inputs`` / ``targets`` / ``mask`` and the block-diagonal batch operators as torch sparse COO (int64 indices, fp32
values).  Per-mesh operators are CSR with sorted rows, so stacking them mesh by mesh already yields the coalesced
order ``sparse_diag_cat(...).coalesce()`` (utils_pt.py:41-53) would produce -- no 1.6 s sort.
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry

__all__ = ["MeshOps", "make_mesh_ops", "block_diag_coo", "arap_batch", "lap_batch"]


class MeshOps:
    """One mesh and its three operators (scipy CSR, float32)."""

    def __init__(self, V, F):
        self.V, self.F = V, F
        self.L = geometry.build_laplacian(V, F)
        self.Di, self.DiA = geometry.build_dirac(V, F)
        for m in (self.L, self.Di, self.DiA):
            m.sort_indices()

    @property
    def num_vertices(self):
        return self.V.shape[0]

    @property
    def num_faces(self):
        return self.F.shape[0]


def make_mesh_ops(num_vertices, seeds, order=None):
    """Synthetic meshes with their operators; ``order`` ("bisect", "morton", "rcm", "morton_xy") renumbers each mesh with
    ``geometry.locality_order`` first (a preprocessing choice: same meshes, same per-vertex results, renumbered)."""
    out = []
    for s in seeds:
        V, F = geometry.synth_mesh(num_vertices, s)
        if order and order != "none":
            kw = {"method": "morton", "axes": (0, 1)} if order == "morton_xy" else {"method": order}
            V, F = geometry.reorder_mesh(V, F, *geometry.locality_order(V, F, **kw))
        out.append(MeshOps(V, F))
    return out


def block_diag_coo(mats, size0, size1):
    """Per-mesh CSR matrices -> coalesced block-diagonal torch COO [B*size0, B*size1] (CPU, int64 / fp32)."""
    rows, cols, vals = [], [], []
    for i, m in enumerate(mats):
        m = m.tocsr()
        if not m.has_sorted_indices:
            m = m.sorted_indices()
        if m.shape[0] > size0 or m.shape[1] > size1:
            raise ValueError("operator %s larger than the padded size (%d, %d)" % (m.shape, size0, size1))
        counts = np.diff(m.indptr)
        rows.append(np.repeat(np.arange(m.shape[0], dtype=np.int64), counts) + i * size0)
        cols.append(m.indices.astype(np.int64) + i * size1)
        vals.append(m.data.astype(np.float32, copy=False))
    idx = torch.from_numpy(np.stack([np.concatenate(rows), np.concatenate(cols)]))
    val = torch.from_numpy(np.concatenate(vals))
    return torch.sparse_coo_tensor(idx, val, (len(mats) * size0, len(mats) * size1), is_coalesced=True)


def arap_batch(meshes, seed=0, dirac=True):
    """as_rigid_as_possible-shaped batch: inputs [B,V,6] (two frames), targets [B,V,120] (40 frames), mask [B,V,1]."""
    B = len(meshes)
    nv = max(m.num_vertices for m in meshes)
    nf = max(m.num_faces for m in meshes)
    rng = np.random.default_rng(seed)
    inputs = np.zeros((B, nv, 6), dtype=np.float32)
    targets = np.zeros((B, nv, 120), dtype=np.float32)
    mask = np.zeros((B, nv, 1), dtype=np.float32)
    for b, m in enumerate(meshes):
        n = m.num_vertices
        frame0 = m.V.astype(np.float32)
        vel = 0.01 * rng.standard_normal((n, 3)).astype(np.float32)
        inputs[b, :n, :3] = frame0
        inputs[b, :n, 3:] = frame0 + vel
        steps = np.arange(2, 42, dtype=np.float32)[None, :, None]
        targets[b, :n] = (frame0[:, None, :] + steps * vel[:, None, :]).reshape(n, 120)
        mask[b, :n] = 1
    out = {"inputs": torch.from_numpy(inputs), "targets": torch.from_numpy(targets), "mask": torch.from_numpy(mask),
           "num_vertices": nv, "num_faces": nf, "batch_size": B}
    if dirac:
        out["Di"] = block_diag_coo([m.Di for m in meshes], 4 * nf, 4 * nv)
        out["DiA"] = block_diag_coo([m.DiA for m in meshes], 4 * nv, 4 * nf)
    else:
        out["L"] = block_diag_coo([m.L for m in meshes], nv, nv)
    return out


def lap_batch(meshes):
    """mesh_mnist-shaped Laplacian batch: block-diagonal L and vertex positions as 3-channel inputs."""
    nv = max(m.num_vertices for m in meshes)
    inputs = np.zeros((len(meshes), nv, 3), dtype=np.float32)
    mask = np.zeros((len(meshes), nv, 1), dtype=np.float32)
    for b, m in enumerate(meshes):
        inputs[b, :m.num_vertices] = m.V
        mask[b, :m.num_vertices] = 1
    return {"inputs": torch.from_numpy(inputs), "mask": torch.from_numpy(mask), "num_vertices": nv,
            "L": block_diag_coo([m.L for m in meshes], nv, nv), "batch_size": len(meshes)}
