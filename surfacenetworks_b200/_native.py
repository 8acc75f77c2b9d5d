"""ctypes binding of libsurfnet_b200.so (C ABI in include/surfnet_b200.h).

The library is the product: there is no Python / CPU fallback.  If the shared object is missing or a
symbol cannot be resolved, importing this module raises -- the layers must fail loudly rather than run
on something else (the reference's seam had no error handling at all, src/utils/cuda/sparse_bmm.py:57-59;
here every call's status code is checked).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SURFNET_B200_LIB", os.path.join(_HERE, "libsurfnet_b200.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libsurfnet_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C surfacenetworks_b200/csrc` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_i64, _i32, _f32, _ptr, _sz, _int = c_int64, c_int32, c_float, c_void_p, c_size_t, c_int

# name -> (restype, argtypes).  Pointers are passed as raw addresses (tensor.data_ptr()).
SIGNATURES = {
    "sn_version": (_int, []),
    "sn_status_string": (c_char_p, [_int]),
    "sn_coo_to_csr32_ws_bytes": (_sz, [_i64, _i64]),
    "sn_coo_to_csr32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _int, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_csr32_to_bsr4_ws_bytes": (_sz, [_i64]),
    "sn_csr32_to_bsr4_count": (_int, [_ptr, _ptr, _i64, _ptr, _ptr, _sz, _ptr]),
    "sn_csr32_to_bsr4_fill": (_int, [_ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr]),
    "sn_assemble_block_diag": (_int, [_ptr, _i64, _i64, _i64, _i64, _int, _ptr, _ptr, _ptr, _ptr]),
    "sn_csr_spmm_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _int, _ptr]),
    "sn_bsr4_spmm_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _int, _ptr]),
    "sn_elu_f32": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr]),
    "sn_gemm_tn_tf32_ws_bytes": (_sz, [_i64, _i64]),
    "sn_gemm_tn_tf32_f32": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _int, _ptr, _sz, _ptr]),
    "sn_gemm_tn_colsum_tf32_f32": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _int, _ptr, _sz, _ptr]),
    "sn_bn_fold_fwd_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                  _f32, _i64, _ptr, _ptr, _ptr, _ptr]),
    "sn_bn_fold_bwd_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _int, _ptr, _ptr, _ptr, _ptr,
                                  _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "sn_colstats_ws_bytes": (_sz, [_i64]),
    "sn_colstats_f32": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_elu_colstats_f32": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_gemm_tf32_ws_bytes": (_sz, [_i64, _i64]),
    "sn_gemm_tf32_f32": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _int,
                                _ptr, _sz, _ptr]),
    "sn_gemm_tf32_presplit_f32": (_int, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64,
                                         _i64, _int, _ptr]),
    "sn_gemm_act_ws_bytes": (_sz, [_i64]),
    "sn_gemm_tf32_presplit_act_f32": (_int, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _ptr,
                                             _i64, _ptr, _ptr, _i64, _i64, _i64, _int, _ptr, _sz, _ptr]),
    "sn_spmm_stats_ws_bytes": (_sz, [_i64]),
    "sn_bsr4_spmm_stats_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _ptr, _int, _ptr, _sz, _ptr]),
    "sn_csr_spmm_stats_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _ptr, _int, _ptr, _sz, _ptr]),
    "sn_avg_stage_ws_bytes": (_sz, [_i64, _i64]),
    "sn_avg_stage_pre_f32": (_int, [_ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_avg_fold_fwd_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                   _f32, _i64, _ptr, _i64, _ptr, _ptr, _ptr]),
    "sn_avg_fold_bwd_f32": (_int, [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _int,
                                   _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "sn_linear_smallk_fwd_f32": (_int, [_ptr, _i64, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr]),
    "sn_linear_smallk_bwd_ws_bytes": (_sz, [_i64, _i64]),
    "sn_linear_smallk_bwd_f32": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_head_add_tiled_f32": (_int, [_ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _i64, _i64, _ptr]),
    "sn_head_pad_grad_f32": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr]),
    "sn_masked_smooth_l1_ws_bytes": (_sz, []),
    "sn_masked_smooth_l1_fwd_f32": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _f32, _ptr, _ptr, _sz, _ptr]),
    "sn_masked_smooth_l1_bwd_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _f32, _ptr, _ptr]),
    "sn_gemm_nt_wide_tf32_f32": (_int, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr]),
    "sn_stage_fwd_ws_bytes": (_sz, [_i64]),
    "sn_stage_bwd_ws_bytes": (_sz, [_i64, _i64]),
    "sn_dir_stage_fwd_f32": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64,
                                    _ptr, _ptr, _f32, _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _sz, _ptr]),
    "sn_lap_stage_fwd_f32": (_int, [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _f32,
                                    _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _sz, _ptr]),
    "sn_dir_stage_bwd_f32": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64,
                                    _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_lap_stage_bwd_f32": (_int, [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64,
                                    _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_gemm_tf32_presplit_elubwd_f32": (_int, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64,
                                                _ptr, _i64, _i64, _i64, _i64, _ptr]),
    "sn_split_tf32_f32": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr]),
    "sn_csr_spmm_epilogue_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                                        _int, _ptr]),
    "sn_bsr4_spmm_epilogue_f32": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                                         _int, _ptr]),
    "sn_mesh_ws_bytes": (_sz, [_i64, _i64, _i64]),
    "sn_mesh_dirac_bsr4": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                  _ptr, _sz, _ptr]),
    "sn_mesh_laplacian_csr": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "sn_segment_sum_ws_bytes": (_sz, [_i64, _i64]),
    "sn_segment_sum_f32": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _ptr, _sz, _ptr]),
    "sn_elu_bwd_group_f32": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr]),
    "sn_elu_bwd_f32": (_int, [_ptr, _i64, _int, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == broken build; do not swallow it
    _fn.restype = _res
    _fn.argtypes = _args

SN_OK = 0
SN_ERR_ARG, SN_ERR_UNSUPPORTED, SN_ERR_WORKSPACE, SN_ERR_OVERFLOW = -1, -2, -3, -4
SN_COO_SORTED = 1
SN_SPMM_ELU_INPUT = 1
SN_SPMM_DIRECT_GATHER = 2
SN_SPMM_SMEM_STREAM = 4


def spmm_flags(elu_input=False, direct_gather=False, smem_stream=False, variant=0, row_entries=0):
    """flags word of sn_csr_spmm_f32 / sn_bsr4_spmm_f32 (include/surfnet_b200.h).  ``row_entries``: SN_SPMM_ROW_ENTRIES
    hint (typical entries per row; 0 = unknown)."""
    return ((SN_SPMM_ELU_INPUT if elu_input else 0) | (SN_SPMM_DIRECT_GATHER if direct_gather else 0)
            | (SN_SPMM_SMEM_STREAM if smem_stream else 0) | ((int(variant) & 15) << 8)
            | ((int(row_entries) if 0 < int(row_entries) < 16 else 0) << 12))


SN_GEMM_SINGLE_PASS = 1
SN_GEMM_NO_L2_PREFETCH = 2
SN_GEMM_ELU_BWD_LEFT = 4
SN_GEMM_LEGACY_SS = 8


class SurfnetError(RuntimeError):
    def __init__(self, fn, status):
        self.status = status
        msg = lib.sn_status_string(status)
        super().__init__("%s failed: status %d (%s)" % (fn, status, msg.decode() if msg else "?"))


def check(fn, status):
    if status != SN_OK:
        raise SurfnetError(fn, status)


# Launch accounting: every successful call of a kernel-launching entry point is counted here (bench.py reports
# the total inside its timed region as "gpu_launches"); TIMER, when set, brackets selected calls with CUDA events
# on the launching stream (bench.py's live roofline measurement).
CALL_COUNTS = {}
TIMER = None


class KernelTimer:
    """Collects (name, tag, bytes, flops, start_event, end_event) for calls whose name is in ``names`` (None: every
    kernel-launching entry point)."""

    def __init__(self, names=None):
        self.names = set(n for n in SIGNATURES if not n.endswith("_ws_bytes") and n not in ("sn_version", "sn_status_string")) \
            if names is None else set(names)
        self.records = []
        self._meta = None

    def annotate(self, tag, nbytes, flops):
        self._meta = (tag, nbytes, flops)

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, tag, nbytes, flops, e0, e1 in self.records:
            s = out.setdefault((name, tag), {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            s["launches"] += 1
            s["ms"] += e0.elapsed_time(e1)
            s["bytes"] += nbytes
            s["flops"] += flops
        return out


def call(name, *args, soft_unsupported=False):
    """Call an int-returning entry point and raise SurfnetError on a non-zero status.  With ``soft_unsupported`` a
    SN_ERR_UNSUPPORTED status (nothing was launched) is returned to the caller instead, who then takes another path."""
    t = TIMER
    if t is not None and name in t.names:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        meta, t._meta = t._meta, None
        if soft_unsupported and rc == SN_ERR_UNSUPPORTED:
            return rc
        check(name, rc)
        tag, nbytes, flops = meta or ("", 0, 0)
        t.records.append((name, tag, nbytes, flops, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
        if soft_unsupported and rc == SN_ERR_UNSUPPORTED:
            return rc
        check(name, rc)
    CALL_COUNTS[name] = CALL_COUNTS.get(name, 0) + 1
    return SN_OK


def version():
    return lib.sn_version()
