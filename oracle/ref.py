"""TEST INFRASTRUCTURE ONLY — ctypes front-end of ``oracle/_ref/*.so``.

Those two libraries are the reference's OWN CPU code, compiled unmodified from
/root/reference by oracle/Makefile (see oracle/ref_wrap/*):

  libsegops_ref.so         seg_ops_cuda/seg_ops.cu:788-844,1048-1129 (authors' CPU loops)
  libgraph_sampler_ref.so  GraphSampler/graph_sampler.cpp (host CSR bookkeeping)

They exist in the authoring container and travel to the GPU box prebuilt; ``available()``
says whether they can be loaded.  They are used to pin ``seg_ops_oracle.c`` and as the
``--impl reference`` CPU arm of bench.py.
"""
import ctypes
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_libs = {}


def _load(name):
    if name not in _libs:
        path = os.path.join(_DIR, name)
        _libs[name] = ctypes.CDLL(path) if os.path.exists(path) else None
    return _libs[name]


def segops_lib():
    return _load("libsegops_ref.so")


def sampler_lib():
    return _load("libgraph_sampler_ref.so")


def available():
    return segops_lib() is not None and sampler_lib() is not None


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- seg_ops.cu CPU loops (write semantics only: the prototype always overwrites) ----
def weighted_pool_fwd(data, weights, indices, indptr):
    data, w, ids, ptr = _f(data), _f(weights), _i(indices), _i(indptr)
    K, n_nb, F = data.shape
    n_seg = len(ptr) - 1
    out = np.full((K, n_seg, F), 7.25, np.float32)
    segops_lib().ref_weighted_pool_fwd(_p(out), _p(w), _p(data), _p(ids), _p(ptr), K, n_seg, n_nb, len(ids), F)
    return out


def weighted_pool_bwd_data(gout, weights, indices, indptr, n_nb):
    g, w, ids, ptr = _f(gout), _f(weights), _i(indices), _i(indptr)
    K, n_seg, F = g.shape
    out = np.full((K, n_nb, F), 7.25, np.float32)
    segops_lib().ref_weighted_pool_bwd_data(_p(out), _p(w), _p(g), _p(ids), _p(ptr), K, n_seg, n_nb, len(ids), F)
    return out


def take_k_corr(embed1, embed2, ids, indptr):
    e1, e2, ids, ptr = _f(embed1), _f(embed2), _i(ids), _i(indptr)
    K, n_node, F = e1.shape
    out = np.full((K, len(ids)), 7.25, np.float32)
    segops_lib().ref_take_k_corr(_p(out), _p(e1), _p(e2), _p(ids), _p(ptr), K, n_node, e2.shape[1], len(ids), F)
    return out


_POOL = {"sum": 0, "avg": 1, "mean": 1, "max": 2}


def seg_pool_fwd(data, indices, indptr, pool_type):
    data, ids, ptr = _f(data), _i(indices), _i(indptr)
    B, total, F = data.shape
    n_seg = len(ptr) - 1
    out = np.full((B, n_seg, F), 7.25, np.float32)
    am = np.full((B, n_seg, F), 12345, np.int32)
    segops_lib().ref_seg_pool_fwd(_p(out), _p(am), _p(data), _p(ids), _p(ptr), B, n_seg, F, total, len(ids),
                                  _POOL[pool_type])
    return out, am


def seg_pool_bwd(gout, argmax, indices, indptr, total, pool_type):
    g, am, ids, ptr = _f(gout), _i(argmax), _i(indices), _i(indptr)
    B, n_seg, F = g.shape
    out = np.full((B, total, F), 7.25, np.float32)
    segops_lib().ref_seg_pool_bwd(_p(out), _p(g), _p(am), _p(ids), _p(ptr), B, n_seg, F, total, len(ids),
                                  _POOL[pool_type])
    return out


# ---- GraphSampler bookkeeping ----
def gen_row_indices_by_indptr(indptr, nnz):
    ptr = _i(indptr)
    out = np.empty(nnz, np.int32)
    n = sampler_lib().ref_gen_row_indices_by_indptr(_p(ptr), len(ptr) - 1, nnz, _p(out))
    assert n == nnz
    return out


def get_support(row_degrees, col_degrees, indptr, end_points, symm=True):
    rd, cd, ptr, ep = _i(row_degrees), _i(col_degrees), _i(indptr), _i(end_points)
    out = np.empty(len(ep), np.float32)
    sampler_lib().ref_get_support(_p(rd), _p(cd), _p(ptr), _p(ep), len(ptr) - 1, len(ep), int(symm), _p(out))
    return out


def multi_link_split_by_value(edge_values, indptr, possible_values):
    ev, ptr, pv = _f(edge_values), _i(indptr), _f(possible_values)
    n, nnz, R = len(ptr) - 1, len(ev), len(pv)
    idx = np.empty(max(nnz, 1), np.int32)
    cnt = np.empty(R, np.int32)
    ptrs = np.empty((R, n + 1), np.int32)
    w = sampler_lib().ref_multi_link_split_by_value(_p(ev), _p(ptr), _p(pv), n, nnz, R, _p(idx), _p(cnt), _p(ptrs))
    assert w == nnz
    offs = np.concatenate([[0], np.cumsum(cnt)])
    return [idx[offs[r]:offs[r + 1]].copy() for r in range(R)], [ptrs[r].copy() for r in range(R)]


def random_sample_fix_neighbor(seed, src_indptr, sel_indices, neighbor_num):
    ptr, sel = _i(src_indptr), _i(sel_indices)
    cap = int(ptr[-1]) + 1
    out = np.empty(cap, np.int32)
    optr = np.empty(len(sel) + 1, np.int32)
    n = sampler_lib().ref_random_sample_fix_neighbor(seed, _p(ptr), _p(sel), len(sel), neighbor_num, cap, _p(out),
                                                     _p(optr))
    assert n >= 0
    return out[:n].copy(), optr


def remove_edges(end_points, values, indptr, rm_rows, rm_cols):
    ep, val, ptr, rr, rc = _i(end_points), _f(values), _i(indptr), _i(rm_rows), _i(rm_cols)
    oep = np.empty(max(len(ep), 1), np.int32)
    oval = np.empty(max(len(ep), 1), np.float32)
    optr = np.empty(len(ptr), np.int32)
    n = sampler_lib().ref_remove_edges(_p(ep), _p(val), _p(ptr), _p(rr), _p(rc), len(ptr) - 1, len(ep), len(rr),
                                       _p(oep), _p(oval), _p(optr))
    return oep[:n].copy(), oval[:n].copy(), optr
