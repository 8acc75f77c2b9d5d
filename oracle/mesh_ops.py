"""Per-face / per-vertex restatement of the reference's mesh operators.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference builds its operators with dense numpy temporaries:

    mesh.dist / mesh.area / mesh.cotangent_weights   src/utils/mesh.py:17-26, 67-80, 102-112
    graph.laplacian (normalized=False)               src/utils/graph.py:40-49
    recipe A^-1 (D - W), .astype('float32')          src/as_rigid_as_possible/add_laplacian.py:50-61
    mesh.dirac (D, D*)                               src/utils/mesh.py:35-64

This file restates the same arithmetic entry by entry, in plain Python loops over faces and vertices and in the
reference's fp64 operation order (left-to-right sums, no fused multiply-add), i.e. the order the GPU construction
kernels (surfacenetworks_b200/csrc/mesh_ops.cu) replay.  Small meshes only (pure-Python loops).  Pinned by
tests/test_oracle_golden.py against the operators the reference's own code produced for cube.ply
(tests/golden/operators.npz).
"""
import numpy as np

f64 = np.float64


def face_geometry(V, i0, i1, i2):
    """Edge lengths (l01, l12, l20) and Heron area with the 1e-6 floor -- mesh.py:17-26, 67-80."""
    def sqdist(a, b):
        d = V[a] - V[b]
        return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
    l01, l12, l20 = np.sqrt(sqdist(i0, i1)), np.sqrt(sqdist(i1, i2)), np.sqrt(sqdist(i2, i0))
    s = ((l01 + l12) + l20) / f64(2.0)
    prod = ((s * (s - l01)) * (s - l12)) * (s - l20)
    return l01, l12, l20, (np.sqrt(prod) if prod > 0 else f64(1e-6))


def _incidence(V, F):
    inc = [[] for _ in range(V.shape[0])]
    for f in range(F.shape[0]):
        for c in range(3):
            inc[F[f, c]].append((f, c))               # ascending (face, corner): the order of np.add.at in the reference
    return inc


def laplacian_coo(V, F):
    """(row, col, float32 value) of A^-1 (diag(colsum W) - W) in row-major order with ascending columns."""
    V = np.asarray(V, dtype=f64)
    rows, cols, vals = [], [], []
    for vi, keys in enumerate(_incidence(V, F)):
        if not keys:
            continue
        nb, wij, wji, A = [], [], [], f64(0.0)
        for f, c in keys:
            ff = F[f]
            l01, l12, l20, a = face_geometry(V, ff[0], ff[1], ff[2])
            sq = np.zeros((3, 3))
            sq[0, 1] = sq[1, 0] = l01 * l01
            sq[1, 2] = sq[2, 1] = l12 * l12
            sq[2, 0] = sq[0, 2] = l20 * l20
            den = f64(8.0) * a + f64(1e-6)                                   # mesh.py:109
            x = (a / f64(3.0)) / f64(4.0)                                    # mesh.py:110, once per permutation
            A = (A + x) + x
            o1, o2 = (1 if c == 0 else 0), (1 if c == 2 else 2)             # itertools.permutations order
            for pj, pk in ((o1, o2), (o2, o1)):
                cij = ((-sq[c, pj] + sq[pj, pk]) + sq[pk, c]) / den          # contribution to W[i, j]
                cji = ((-sq[pj, c] + sq[c, pk]) + sq[pk, pj]) / den          # contribution to W[j, i] (column i)
                j, pos = ff[pj], len(nb)
                while pos > 0 and nb[pos - 1] > j:                           # stable: equal neighbours keep face order
                    pos -= 1
                nb.insert(pos, j)
                wij.insert(pos, cij)
                wji.insert(pos, cji)
        ainv = f64(1.0) / (A + f64(1e-9))                                    # add_laplacian.py:53

        def merged(w):
            s = 0
            while s < len(nb):
                acc, e = w[s], s + 1
                while e < len(nb) and nb[e] == nb[s]:
                    acc = acc + w[e]
                    e += 1
                yield nb[s], acc
                s = e
        d = f64(0.0)
        for _, w in merged(wji):                                             # graph.py:44: degrees = column sums of W
            d = d + w
        diag_done = False
        for j, w in merged(wij):
            if not diag_done and j > vi:
                if d != 0:
                    rows.append(vi), cols.append(vi), vals.append(np.float32(ainv * d))
                diag_done = True
            if w != 0:                                                        # csr_matrix(dense) drops exact zeros
                rows.append(vi), cols.append(j), vals.append(np.float32(ainv * (-w)))
        if not diag_done and d != 0:
            rows.append(vi), cols.append(vi), vals.append(np.float32(ainv * d))
    return np.array(rows, np.int64), np.array(cols, np.int64), np.array(vals, np.float32)


def dirac_entries(V, F):
    """Dicts {(row, col): float32} of the non-zero entries of D [4F x 4V] and D* [4V x 4F] -- mesh.py:35-64."""
    V = np.asarray(V, dtype=f64)
    nf, nv = F.shape[0], V.shape[0]
    Af = np.array([face_geometry(V, *F[f])[3] for f in range(nf)])
    Av = np.zeros(nv)
    for f in range(nf):
        for c in range(3):
            Av[F[f, c]] = Av[F[f, c]] + Af[f] / f64(3.0)                     # mesh.py:44-45
    D, DA = {}, {}
    for f in range(nf):
        for c in range(3):
            b, cc, d = V[F[f, (c + 1) % 3]] - V[F[f, (c + 2) % 3]]
            Q = np.array([[0, -b, -cc, -d], [b, 0, -d, cc], [cc, d, 0, -b], [d, -cc, b, 0]], dtype=f64)   # mesh.py:28-33
            M = (-Q) / (f64(2.0) * Af[f])                                    # mesh.py:47-58
            j = F[f, c]
            T = (M.T * Af[f]) / Av[j]                                        # mesh.py:59
            for p in range(4):
                for q in range(4):
                    if M[p, q] != 0:
                        D[(4 * f + p, 4 * j + q)] = np.float32(M[p, q])
                    if T[p, q] != 0:
                        DA[(4 * j + p, 4 * f + q)] = np.float32(T[p, q])
    return D, DA
