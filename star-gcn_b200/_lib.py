"""ctypes binding of libstargcn_b200.so (the C ABI in include/stargcn_b200.h).

There is NO CPU or eager fallback: if the shared library cannot be loaded (and cannot be
built because nvcc is absent) importing this module raises, and every operator in this
package therefore fails loudly instead of silently computing somewhere else.
"""
import ctypes
import os

from . import _build

SG_OK = 0
REQ = {"null": 0, "write": 1, "add": 3}
POOL = {"sum": 0, "avg": 1, "mean": 1, "max": 2}
REDUCE = {"sum": 0, "max": 2, "min": 3}
BCAST = {"add": 0, "mul": 1, "to": 2, "sub": 3, "div": 4}

_c_int, _c_sz, _c_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/stargcn_b200.h one to one
SIGNATURES = {
    "sg_last_error": (ctypes.c_char_p, []),
    "sg_abi_version": (_c_int, []),
    "sg_launch_count": (ctypes.c_longlong, []),
    "sg_launch_count_reset": (None, []),
    "sg_dev_option": (_c_int, [_c_int, _c_int]),
    "sg_seg_ids": (_c_int, [_c_p, _c_p, _c_int, _c_int, _c_p]),
    "sg_csr_transpose_ws_bytes": (_c_sz, [_c_int, _c_int, _c_int]),
    "sg_csr_transpose": (_c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_p, _c_sz, _c_p]),
    "sg_plan_bytes": (_c_sz, [_c_int, _c_int, _c_int]),
    "sg_plan_partial_rows": (_c_sz, [_c_int, _c_int, _c_int]),
    "sg_plan_build": (_c_int, [_c_p, _c_sz, _c_p, _c_int, _c_int, _c_int, _c_p]),
    "sg_weighted_pool_fwd": (_c_int, [_c_p] * 5 + [_c_int] * 6 + [_c_p, _c_int, _c_p, _c_p]),
    "sg_weighted_pool_bwd_data": (_c_int, [_c_p] * 6 + [_c_int] * 6 + [_c_p, _c_int, _c_p, _c_p]),
    "sg_take_k_corr": (_c_int, [_c_p] * 5 + [_c_int] * 6 + [_c_p]),
    "sg_seg_pool_fwd": (_c_int, [_c_p] * 5 + [_c_int] * 6 + [_c_p, _c_int, _c_p, _c_p]),
    "sg_seg_pool_bwd": (_c_int, [_c_p] * 7 + [_c_int] * 7 + [_c_p, _c_int, _c_p, _c_p]),
    "sg_seg_reduce": (_c_int, [_c_p] * 3 + [_c_int] * 5 + [_c_p]),
    "sg_seg_broadcast_binary": (_c_int, [_c_p] * 4 + [_c_int] * 5 + [_c_p]),
    "sg_seg_softmax_fwd": (_c_int, [_c_p] * 3 + [_c_int] * 3 + [_c_p]),
    "sg_seg_softmax_bwd": (_c_int, [_c_p] * 4 + [_c_int] * 4 + [_c_p]),
    "sg_multilink_agg_fwd": (_c_int, [_c_p] * 6 + [_c_int] * 5 + [_c_p, _c_int, _c_p, _c_p]),
    "sg_multilink_agg_fwd_split": (_c_int, [_c_p, _c_p, _c_int, _c_p, _c_p, _c_p, _c_p] + [_c_int] * 5 +
                                   [_c_p, _c_int, _c_p, _c_p]),
    "sg_multilink_transpose_finish": (_c_int, [_c_p] * 5 + [_c_int] * 3 + [_c_p]),
    "sg_gemm_split_ws_bytes": (_c_sz, [_c_int, _c_int, _c_int]),
    "sg_gemm_tf32x3": (_c_int, [_c_p, _c_int, _c_p, _c_p, _c_int, _c_p, _c_p, _c_int] + [_c_int] * 5 +
                       [ctypes.c_float, _c_p, _c_int, _c_p, _c_p]),
    "sg_csr_support": (_c_int, [_c_p] * 5 + [_c_int] * 3 + [_c_p]),
    "sg_sampler_ws_bytes": (_c_sz, [_c_int, _c_int]),
    "sg_sample_neighbors_count": (_c_int, [_c_p] * 3 + [_c_int] * 2 + [_c_p, _c_p]),
    "sg_sample_neighbors_fill": (_c_int, [_c_p] * 4 + [_c_int, ctypes.c_ulonglong, _c_p]),
    "sg_multilink_split": (_c_int, [_c_p] * 12 + [_c_int] * 2 + [_c_p, _c_p]),
    "sg_remove_edges_ws_bytes": (_c_sz, [_c_int, _c_int]),
    "sg_remove_edges_count": (_c_int, [_c_p] * 5 + [_c_int] * 3 + [_c_p, _c_p]),
    "sg_remove_edges_fill": (_c_int, [_c_p] * 6 + [_c_int] * 2 + [_c_p, _c_p]),
    "sg_masked_support": (_c_int, [_c_p] * 7 + [_c_int] * 2 + [_c_p]),
    "sg_bincount": (_c_int, [_c_p, _c_p, _c_int, _c_int, _c_p]),
    "sg_unique_inverse_ws_bytes": (_c_sz, [_c_int]),
    "sg_unique_inverse": (_c_int, [_c_p] * 4 + [_c_int, _c_p, _c_sz, _c_p]),
    "sg_optim_chunk": (_c_int, []),
    "sg_global_norm": (_c_int, [_c_p] * 4 + [_c_int, ctypes.c_float, _c_p, _c_p]),
    "sg_multi_adam": (_c_int, [_c_p] * 6 + [_c_int] + [ctypes.c_float] * 6 + [_c_p, _c_int, _c_p]),
    "sg_masked_embed_fwd": (_c_int, [_c_p] * 5 + [_c_int] * 3 + [_c_p]),
    "sg_reduce_ws_bytes": (_c_sz, []),
    "sg_sq_err_fwd": (_c_int, [_c_p, _c_p, _c_p, ctypes.c_longlong, ctypes.c_float, _c_p, _c_p]),
    "sg_sq_err_bwd": (_c_int, [_c_p] * 5 + [ctypes.c_longlong, ctypes.c_float, _c_p]),
    "sg_rowdot_fwd": (_c_int, [_c_p] * 3 + [_c_int] * 2 + [_c_p]),
    "sg_rowdot_bwd": (_c_int, [_c_p] * 5 + [_c_int] * 2 + [_c_p]),
    "sg_colsum_ws_bytes": (_c_sz, [_c_int]),
    "sg_colsum": (_c_int, [_c_p] * 3 + [_c_int] * 3 + [_c_p, _c_p]),
    "sg_split_tf32": (_c_int, [_c_p, _c_p, _c_int, _c_p, _c_int, _c_int, _c_int, _c_int, _c_p]),
    "sg_act_bwd_split": (_c_int, [_c_p, _c_p, _c_int, _c_p, _c_p, _c_int, _c_int, ctypes.c_float, _c_p]),
    "sg_gemm_trace_read": (_c_int, [_c_p]),
    "sg_tma_probe": (_c_int, [_c_p] + [_c_int] * 7 + [_c_p]),
    "sg_upload_segments": (_c_int, [_c_p, _c_p, _c_p, _c_int, _c_p]),
    "sg_row_gather_probe": (_c_int, [_c_p, _c_p, _c_int, _c_int, _c_int, ctypes.c_uint, _c_p]),
    "sg_multilink_agg_bwd": (_c_int, [_c_p] * 5 + [_c_int] * 6 + [_c_p, _c_int, _c_p, _c_p]),
    "sg_peer_push_rows": (_c_int, [_c_p, _c_p, ctypes.c_longlong, _c_int, _c_p]),
    "sg_peer_barrier": (_c_int, [_c_p, _c_p, _c_int, _c_int, ctypes.c_double, _c_p]),
    "sg_peer_reduce": (_c_int, [_c_p, _c_p, ctypes.c_longlong, ctypes.c_longlong, _c_int, _c_int, _c_p]),
    "sg_multilink_agg_bwd_peer": (_c_int, [_c_p, _c_p, _c_int] + [_c_p] * 4 + [_c_int] * 5 + [_c_p, _c_int, _c_p, _c_p]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is missing and nvcc exists).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class StarGCNError(RuntimeError):
    """Raised when a libstargcn_b200 entry point reports an error (cf. MXNetError)."""


def check(rc, what):
    if rc != SG_OK:
        msg = load().sg_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        raise StarGCNError(f"{what} failed (code {rc}): {msg}")


DEV_OPTIONS = {"gather_variant": 0, "gather_grid": 1, "gemm_arrive": 2, "gemm_chain": 3, "gather_threads": 4, "peer_push_blocks": 5, "gemm_trace": 8}


def dev_option(name, value):
    """Development only: set a tuning option of the library (see sg_dev_option); 0 = shipped behaviour."""
    check(load().sg_dev_option(DEV_OPTIONS[name], int(value)), "sg_dev_option")


def launch_count():
    return int(load().sg_launch_count())


def reset_launch_count():
    load().sg_launch_count_reset()
