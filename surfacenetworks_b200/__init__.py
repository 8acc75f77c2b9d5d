"""surfacenetworks_b200 -- B200 (sm_100a) implementation of SurfaceNetworks' operator-application hot path.

    utils_pt   : the reference's layer library surface (LapResNet2, DirResNet2, GraphConv1x1, ...)
    models     : the model stacks that call it (as_rigid_as_possible DirModel / Model, ...)
    operators  : device-resident CSR32 / BSR4 operators built once from torch COO, assembled from a per-mesh cache, or
                 constructed on the GPU from vertex positions + faces
    ops        : autograd seam (forward S @ x, backward S^T @ g; the single-node Dirac block)
    fused      : the dense half of a stage (BatchNorm folded into a tcgen05 3xTF32 GEMM) and the AvgResNet2 stage
    cuda       : the reference's native seam under its own names (batch_csr, sparse_bmm, SparseBMMFunc)
    geometry   : host-side mesh operators (cotangent Laplacian, Dirac, adjoint) + synthetic meshes
    _native    : ctypes binding of libsurfnet_b200.so (include/surfnet_b200.h)

The CUDA library is mandatory: importing ``utils_pt`` / ``ops`` / ``operators`` raises if it is not built.
"""
__version__ = "0.1.0"
