"""TEST INFRASTRUCTURE ONLY — numpy/ctypes front-end of ``seg_ops_oracle.c``.

Every function mirrors one reference CPU operator (file:line in the C source header) and
takes/returns numpy arrays (float32 / int32, C-contiguous).  ``req`` follows MXNet's
OpReqType: 'write' | 'add' | 'null' (test_seg_ops.py:130 exercises add and write).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
REQ = {"null": 0, "write": 1, "add": 3}
POOL = {"sum": 0, "avg": 1, "mean": 1, "max": 2}
REDUCE = {"sum": 0, "max": 2, "min": 3}
BCAST = {"add": 0, "mul": 1, "to": 2, "minus": 3, "div": 4}


def build():
    """(Re)build liborc_segops.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc_segops.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _out(shape, req, init, dtype=np.float32):
    if req == "add":
        assert init is not None, "req='add' needs the buffer to accumulate into"
        out = np.array(init, dtype=dtype, order="C", copy=True)
        assert out.shape == tuple(shape)
    else:
        out = np.full(shape, 7.25 if dtype == np.float32 else 12345, dtype=dtype)  # poison
    return out, out.ctypes.data_as(ctypes.c_void_p)


def seg_ids(indptr, nnz=None):
    indptr, pi = _i(indptr)
    nnz = int(indptr[-1]) if nnz is None else nnz
    out = np.full((max(nnz, 0),), -1, np.int32)
    lib().orc_seg_ids(out.ctypes.data_as(ctypes.c_void_p), pi, ctypes.c_int(len(indptr) - 1))
    return out


def seg_reduce(data, indptr, kind="sum", req="write", init=None):
    data, pd = _f(data)
    indptr, pi = _i(indptr)
    B, nnz = data.shape
    n_seg = len(indptr) - 1
    out, po = _out((B, n_seg), req, init)
    lib().orc_seg_reduce(po, pd, pi, B, nnz, n_seg, REDUCE[kind], REQ[req])
    return out


def seg_sum(data, indptr, **kw):
    return seg_reduce(data, indptr, "sum", **kw)


def seg_broadcast_binary(lhs, rhs, indptr, op, nnz=None, req="write", init=None):
    rhs, pr = _f(rhs)
    indptr, pi = _i(indptr)
    B, n_seg = rhs.shape
    if lhs is None:
        pl = ctypes.c_void_p(0)
    else:
        lhs, pl = _f(lhs)
        nnz = lhs.shape[1]
    out, po = _out((B, nnz), req, init)
    lib().orc_seg_broadcast_binary(po, pl, pr, pi, B, nnz, n_seg, BCAST[op], REQ[req])
    return out


def seg_broadcast_add(lhs, rhs, indptr, **kw):
    return seg_broadcast_binary(lhs, rhs, indptr, "add", **kw)


def seg_broadcast_mul(lhs, rhs, indptr, **kw):
    return seg_broadcast_binary(lhs, rhs, indptr, "mul", **kw)


def seg_broadcast_to(data, indptr, nnz, **kw):
    return seg_broadcast_binary(None, data, indptr, "to", nnz=nnz, **kw)


def seg_softmax(data, indptr):
    data, pd = _f(data)
    indptr, pi = _i(indptr)
    out = np.empty_like(data)
    lib().orc_seg_softmax(out.ctypes.data_as(ctypes.c_void_p), pd, pi, data.shape[0], data.shape[1],
                          len(indptr) - 1)
    return out


def seg_softmax_bwd(ograd, val, indptr, req="write", init=None):
    ograd, pg = _f(ograd)
    val, pv = _f(val)
    indptr, pi = _i(indptr)
    out, po = _out(ograd.shape, req, init)
    lib().orc_seg_softmax_bwd(po, pg, pv, pi, ograd.shape[0], ograd.shape[1], len(indptr) - 1, REQ[req])
    return out


def seg_take_k_corr(embed1, embed2, neighbor_ids, neighbor_indptr, req="write", init=None):
    e1, p1 = _f(embed1)
    e2, p2 = _f(embed2)
    ids, pi = _i(neighbor_ids)
    ptr, pp = _i(neighbor_indptr)
    K, n_node, F = e1.shape
    out, po = _out((K, len(ids)), req, init)
    lib().orc_take_k_corr(po, p1, p2, pi, pp, K, n_node, e2.shape[1], len(ids), F, REQ[req])
    return out


def seg_weighted_pool(data, weights, indices, indptr, req="write", init=None):
    data, pd = _f(data)
    w, pw = _f(weights)
    ids, pi = _i(indices)
    ptr, pp = _i(indptr)
    K, n_nb, F = data.shape
    n_seg = len(ptr) - 1
    out, po = _out((K, n_seg, F), req, init)
    lib().orc_weighted_pool_fwd(po, pd, pw, pi, pp, K, n_seg, n_nb, len(ids), F, REQ[req])
    return out


def seg_weighted_pool_bwd_data(gout, weights, indices, indptr, n_nb, req="write", init=None):
    g, pg = _f(gout)
    w, pw = _f(weights)
    ids, pi = _i(indices)
    ptr, pp = _i(indptr)
    K, n_seg, F = g.shape
    out, po = _out((K, n_nb, F), req, init)
    lib().orc_weighted_pool_bwd_data(po, pg, pw, pi, pp, K, n_seg, n_nb, len(ids), F, REQ[req])
    return out


def seg_pool(data, indices, indptr, pool_type="sum", return_argmax=False):
    data, pd = _f(data)
    ids, pi = _i(indices)
    ptr, pp = _i(indptr)
    B, n_nb, F = data.shape
    n_seg = len(ptr) - 1
    out = np.full((B, n_seg, F), 7.25, np.float32)
    am = np.full((B, n_seg, F), 12345, np.int32)
    lib().orc_seg_pool_fwd(out.ctypes.data_as(ctypes.c_void_p), am.ctypes.data_as(ctypes.c_void_p), pd, pi,
                           pp, B, n_seg, n_nb, len(ids), F, POOL[pool_type])
    return (out, am) if return_argmax else out


def seg_pool_bwd(gout, argmax, indices, indptr, n_nb, pool_type="sum", req="write", init=None):
    g, pg = _f(gout)
    ids, pi = _i(indices)
    ptr, pp = _i(indptr)
    B, n_seg, F = g.shape
    if argmax is None:
        pa = ctypes.c_void_p(0)
    else:
        argmax, pa = _i(argmax)
    out, po = _out((B, n_nb, F), req, init)
    lib().orc_seg_pool_bwd(po, pg, pa, pi, pp, B, n_seg, n_nb, len(ids), F, POOL[pool_type], REQ[req])
    return out


def csr_transpose(indices, indptr, n_nb):
    ids, pi = _i(indices)
    ptr, pp = _i(indptr)
    nnz = len(ids)
    t_indptr = np.empty(n_nb + 1, np.int32)
    t_perm = np.empty(nnz, np.int32)
    t_seg = np.empty(nnz, np.int32)
    lib().orc_csr_transpose(t_indptr.ctypes.data_as(ctypes.c_void_p), t_perm.ctypes.data_as(ctypes.c_void_p),
                            t_seg.ctypes.data_as(ctypes.c_void_p), pi, pp, len(ptr) - 1, n_nb, nnz)
    return t_indptr, t_perm, t_seg
