"""oracle/ -- CPU restatement of the reference's operator-application path.  TEST INFRASTRUCTURE ONLY.

Importable from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs, and from nowhere else: the product package ``surfacenetworks_b200`` must never import this.

Pinning: the reference has no known-answer tests for this path, so the oracle is pinned
against fixtures produced by RUNNING THE REFERENCE in the build container (``tests/golden/make_golden.py``
-> ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks every function here against them.

  c_oracle.py  ctypes wrappers of sn_oracle.c (plain C: COO x dense, batch_csr, sparse_bmm, Dirac view, ELU)
  layers.py    functional torch-CPU restatement of utils_pt.py's layers and the ARAP / dense_correspondence model stacks
  mesh_ops.py  per-entry restatement of the mesh operators (mesh.py / graph.py) in the reference's fp64 operation order
"""
