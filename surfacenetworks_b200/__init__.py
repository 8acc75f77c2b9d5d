"""surfacenetworks_b200 -- B200 (sm_100a) implementation of SurfaceNetworks' operator-application hot path.

    utils_pt   : the reference's layer library surface (LapResNet2, DirResNet2, GraphConv1x1, ...)
    models     : the model stacks that call it (as_rigid_as_possible DirModel / Model, ...)
    operators  : device-resident CSR32 / BSR4 operators built once from torch COO
    ops        : autograd seam (forward S @ x, backward S^T @ g)
    geometry   : host-side mesh operators (cotangent Laplacian, Dirac, adjoint) + synthetic meshes
    _native    : ctypes binding of libsurfnet_b200.so (include/surfnet_b200.h)

The CUDA library is mandatory: importing ``utils_pt`` / ``ops`` / ``operators`` raises if it is not built.
"""
__version__ = "0.1.0"
