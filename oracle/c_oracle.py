"""ctypes wrappers around oracle/sn_oracle.c (built by oracle/Makefile into oracle/_build/).  Test infrastructure."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsn_oracle.so")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "sn_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def coo_mm_f32(row, col, val, n_rows, X):
    """torch.mm(sparse_coo, dense) restated: fp32 accumulation in storage order."""
    row, col, val, X = _i64(row), _i64(col), _f32(val), _f32(X)
    C = X.shape[1]
    out = np.empty((n_rows, C), dtype=np.float32)
    lib().oracle_coo_mm_f32(_p(row), _p(col), _p(val), ctypes.c_int64(val.size), ctypes.c_int64(n_rows), _p(X),
                            ctypes.c_int64(C), _p(out), ctypes.c_int64(C), ctypes.c_int64(C))
    return out


def coo_mm_f64(row, col, val, n_rows, X):
    """Double-precision product and the componentwise magnitude bound |S||x|."""
    row, col, val, X = _i64(row), _i64(col), _f32(val), _f32(X)
    C = X.shape[1]
    out = np.empty((n_rows, C), dtype=np.float64)
    bound = np.empty((n_rows, C), dtype=np.float64)
    lib().oracle_coo_mm_f64(_p(row), _p(col), _p(val), ctypes.c_int64(val.size), ctypes.c_int64(n_rows), _p(X),
                            ctypes.c_int64(C), _p(out), _p(bound), ctypes.c_int64(C))
    return out, bound


def batch_csr(indices, B, R):
    """Reference batch_csr.cu semantics (with empty rows handled): returns col_ind[nnz], col_ptr[B, R+1]."""
    indices = _i64(indices)
    nnz = indices.shape[1]
    col_ind = np.empty(nnz, dtype=np.int64)
    col_ptr = np.empty((B, R + 1), dtype=np.int64)
    lib().oracle_batch_csr(_p(indices), ctypes.c_int64(nnz), ctypes.c_int64(B), ctypes.c_int64(R), _p(col_ind),
                           _p(col_ptr))
    return col_ind, col_ptr


def sparse_bmm(values, col_ind, col_ptr, B, R, dense):
    """Reference sparse_bmm.cu semantics: batched CSR x dense [B, Rd, Cd] -> [B, R, Cd]."""
    values, col_ind, col_ptr, dense = _f32(values), _i64(col_ind), _i64(col_ptr), _f32(dense)
    _, Rd, Cd = dense.shape
    out = np.empty((B, R, Cd), dtype=np.float32)
    lib().oracle_sparse_bmm(_p(values), _p(col_ind), _p(col_ptr), ctypes.c_int64(B), ctypes.c_int64(R), _p(dense),
                            ctypes.c_int64(Rd), ctypes.c_int64(Cd), _p(out))
    return out


def dirac_view_mm_f32(row, col, val, R, X):
    """[4R x 4Cn] operator applied through the quaternion view of utils_pt.py:201-203; X [Cn, C] -> [R, C]."""
    row, col, val, X = _i64(row), _i64(col), _f32(val), _f32(X)
    C = X.shape[1]
    out = np.empty((R, C), dtype=np.float32)
    lib().oracle_dirac_view_mm_f32(_p(row), _p(col), _p(val), ctypes.c_int64(val.size), ctypes.c_int64(R), _p(X),
                                   _p(out), ctypes.c_int64(C))
    return out


def dirac_view_mm_f64(row, col, val, R, X):
    row, col, val, X = _i64(row), _i64(col), _f32(val), _f32(X)
    C = X.shape[1]
    out = np.empty((R, C), dtype=np.float64)
    bound = np.empty((R, C), dtype=np.float64)
    lib().oracle_dirac_view_mm_f64(_p(row), _p(col), _p(val), ctypes.c_int64(val.size), ctypes.c_int64(R), _p(X),
                                   _p(out), _p(bound), ctypes.c_int64(C))
    return out, bound


def elu_f32(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().oracle_elu_f32(_p(x), _p(y), ctypes.c_int64(x.size))
    return y
