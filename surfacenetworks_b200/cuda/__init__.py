"""Drop-in mirror of the reference's native seam ``src/utils/cuda/`` (batch_csr, sparse_bmm, SparseBMMFunc).

Same module names, call signatures and tensor layouts as the reference's cupy / NVRTC wrappers; the work is done by
libsurfnet_b200.so (``sn_coo_to_csr32`` and ``sn_csr_spmm_f32``) instead of kernels JIT-compiled per shape.
"""
from .batch_csr import BatchCSR, batch_csr  # noqa: F401
from .sparse_bmm import SparseBMM, sparse_bmm  # noqa: F401
from .sparse_bmm_func import SparseBMMFunc  # noqa: F401
