"""Device-resident graph + neighbour sampling.

``DeviceCSR`` is the device form of one direction of the reference's ``CSRMat``
(mxgraph/graph.py:261-316): ``ind_ptr``, ``end_points`` (column indices), ``values`` (ratings),
``multi_link`` (the possible rating values) and the cached ``support`` (graph.py:414-429).
``sample_neighbors`` mirrors ``CSRMat.sample_neighbors`` (graph.py:677-748) with
``use_multi_link=True`` but never leaves the device: its result is the relation-major
:class:`MultiLinkCSR` the fused aggregation consumes, instead of 4·R numpy arrays that
``heter_sage`` re-uploads on every call (mxgraph/layers/layers.py:366-377).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check
from .graph import MultiLinkCSR
from .seg_op import _bytes, _p, _stream


def _dev_i32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device, torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)


def _dev_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device, torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


class DeviceCSR:
    def __init__(self, ind_ptr, end_points, values, multi_link, n_cols, row_degrees=None, col_degrees=None,
                 support=None, symm=True, device="cuda"):
        self.device = torch.device(device)
        self.ind_ptr = _dev_i32(ind_ptr, self.device)
        self.end_points = _dev_i32(end_points, self.device)
        self.values = _dev_f32(values, self.device)
        self.multi_link = _dev_f32(multi_link, self.device)
        self.n_rows = self.ind_ptr.numel() - 1
        self.n_cols = int(n_cols)
        self.nnz = self.end_points.numel()
        self.R = self.multi_link.numel()
        self._calls = 0              # advances with every sampling call that does not pass its own seed
        if support is not None:
            self.support = _dev_f32(support, self.device)
        else:
            if row_degrees is None or (symm and col_degrees is None):
                raise ValueError("give either `support` or the degree arrays it is computed from")
            self.support = self.compute_support(_dev_i32(row_degrees, self.device),
                                                _dev_i32(col_degrees, self.device) if col_degrees is not None else None, symm)

    def compute_support(self, row_degrees, col_degrees, symm=True):
        """get_support (GraphSampler/graph_sampler.cpp:393-420), on the device, bit-exact."""
        out = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=self.device)[:self.nnz]
        check(_lib.load().sg_csr_support(_p(out), _p(row_degrees), _p(col_degrees), _p(self.ind_ptr), _p(self.end_points),
                                         self.n_rows, self.nnz, int(bool(symm)), _stream()), "sg_csr_support")
        return out

    def remove_edges(self, rm_rows, rm_cols, symm=True, col_degrees=None):
        """CSRMat.remove_edges_by_ind (mxgraph/graph.py:631-658 -> graph_sampler.cpp:154-201): a new DeviceCSR
        without the listed (row index, column index) pairs; order inside rows is preserved.  The support is
        recomputed from the new degrees (``col_degrees``: degrees of the column side after the removal —
        the row degrees of the reverse matrix; computed here with a histogram when not given)."""
        lib = _lib.load()
        dev = self.device
        rr, rc = _dev_i32(rm_rows, dev), _dev_i32(rm_cols, dev)
        if rr.numel() != rc.numel():
            raise ValueError("rm_rows and rm_cols must have the same length")
        ws = _bytes(lib.sg_remove_edges_ws_bytes(self.n_rows, self.nnz), dev)
        new_ptr = torch.empty(self.n_rows + 1, dtype=torch.int32, device=dev)
        check(lib.sg_remove_edges_count(_p(new_ptr), _p(self.ind_ptr), _p(self.end_points), _p(rr), _p(rc), self.n_rows,
                                        self.nnz, rr.numel(), _p(ws), _stream()), "sg_remove_edges_count")
        new_nnz = int(new_ptr[-1].item())
        new_ep = torch.empty(max(new_nnz, 1), dtype=torch.int32, device=dev)[:new_nnz]
        new_val = torch.empty(max(new_nnz, 1), dtype=torch.float32, device=dev)[:new_nnz]
        check(lib.sg_remove_edges_fill(_p(new_ep), _p(new_val), _p(new_ptr), _p(self.ind_ptr), _p(self.end_points),
                                       _p(self.values), self.n_rows, self.nnz, _p(ws), _stream()), "sg_remove_edges_fill")
        row_deg = new_ptr[1:] - new_ptr[:-1]
        if symm and col_degrees is None:
            col_degrees = torch.empty(self.n_cols, dtype=torch.int32, device=dev)
            check(lib.sg_bincount(_p(col_degrees), _p(new_ep), new_nnz, self.n_cols, _stream()), "sg_bincount")
        return DeviceCSR(new_ptr, new_ep, new_val, self.multi_link, self.n_cols, row_degrees=row_deg.contiguous(),
                         col_degrees=col_degrees, symm=symm, device=dev)

    def _next_seed(self, seed):
        """``seed=None`` draws a fresh stream per call (the reference's mt19937 engines advance across calls,
        graph_sampler.cpp:742-779); an explicit seed makes the sample a pure function of (seed, row, draw)."""
        if seed is not None:
            return int(seed)
        self._calls += 1
        return (self._calls * 0x9E3779B97F4A7C15 + 0x1234567) & (2 ** 64 - 1)

    def sample_positions(self, src_inds=None, num_neighbors=-1, seed=None):
        """random_sample_fix_neighbor: (sampled positions on the nnz axis, dst_ind_ptr) as device tensors.
        Fan-outs above 256 per row are rejected by the kernel (its per-row pool lives in shared memory); the
        shipped configurations use -1 (all neighbours)."""
        lib = _lib.load()
        seed = self._next_seed(seed)
        sel = None if src_inds is None else _dev_i32(src_inds, self.device)
        n_sel = self.n_rows if sel is None else sel.numel()
        k = -1 if num_neighbors is None else int(num_neighbors)
        dst_indptr = torch.empty(n_sel + 1, dtype=torch.int32, device=self.device)
        ws = _bytes(lib.sg_sampler_ws_bytes(n_sel, 1), self.device)
        check(lib.sg_sample_neighbors_count(_p(dst_indptr), _p(self.ind_ptr), _p(sel), n_sel, k, _p(ws), _stream()),
              "sg_sample_neighbors_count")
        # upper bound known on the host without a sync: every edge (k < 0) or k per row
        cap = self.nnz if (k < 0 and sel is None) else None
        nnz_s = cap if cap is not None else int(dst_indptr[-1].item())
        sampled = torch.empty(max(nnz_s, 1), dtype=torch.int32, device=self.device)[:nnz_s]
        check(lib.sg_sample_neighbors_fill(_p(sampled), _p(dst_indptr), _p(self.ind_ptr), _p(sel), n_sel,
                                           ctypes.c_ulonglong(int(seed) & (2 ** 64 - 1)), _stream()),
              "sg_sample_neighbors_fill")
        return sampled, dst_indptr, n_sel

    def split(self, sampled, dst_indptr, n_sel, want_index=False, want_values=False, check_values=False):
        """multi_link_split + per-level takes -> (cat_indptr, end_points_cat, support_cat[, split_index, values_cat])."""
        lib = _lib.load()
        nnz_s = sampled.numel()
        R = self.R
        dev = self.device
        cat_indptr = torch.empty(R * n_sel + 1, dtype=torch.int32, device=dev)
        ep_cat = torch.empty(max(nnz_s, 1), dtype=torch.int32, device=dev)[:nnz_s]
        sup_cat = torch.empty(max(nnz_s, 1), dtype=torch.float32, device=dev)[:nnz_s]
        split_index = torch.empty(max(nnz_s, 1), dtype=torch.int32, device=dev)[:nnz_s] if want_index else None
        val_cat = torch.empty(max(nnz_s, 1), dtype=torch.float32, device=dev)[:nnz_s] if want_values else None
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = _bytes(lib.sg_sampler_ws_bytes(n_sel, R), dev)
        check(lib.sg_multilink_split(_p(cat_indptr), _p(split_index), _p(ep_cat), _p(sup_cat), _p(val_cat), _p(bad),
                                     _p(self.values), _p(self.end_points), _p(self.support), _p(sampled), _p(dst_indptr),
                                     _p(self.multi_link), R, n_sel, _p(ws), _stream()), "sg_multilink_split")
        if check_values and int(bad.item()):
            raise ValueError("an edge value is not in multi_link (graph_sampler.cpp:303 ASSERT)")
        return cat_indptr, ep_cat, sup_cat, split_index, val_cat

    def sample_neighbors(self, src_inds=None, num_neighbors=-1, seed=None):
        """CSRMat.sample_neighbors(use_multi_link=True) -> MultiLinkCSR over the selected rows.
        End points are column INDICES of this matrix (the reference maps them to node ids and
        gen_plan maps those back to local indices; with every column present the two coincide)."""
        sampled, dst_indptr, n_sel = self.sample_positions(src_inds, num_neighbors, seed)
        cat_indptr, ep_cat, sup_cat, _, _ = self.split(sampled, dst_indptr, n_sel)
        return MultiLinkCSR.from_device(ep_cat, sup_cat, cat_indptr, self.R, n_sel, self.n_cols)


def unique_inverse(data):
    """(unique values in first-occurrence order, inverse indices) of an int32 device tensor — the serial
    ``unique_inverse`` (GraphSampler/graph_sampler.h:510-534) that ``merge_nodes`` relies on."""
    lib = _lib.load()
    if not isinstance(data, torch.Tensor) or not data.is_cuda or data.dtype != torch.int32 or data.dim() != 1:
        raise TypeError("data must be a 1-D int32 CUDA tensor")
    data = data.contiguous()
    n = data.numel()
    dev = data.device
    uniq = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    inverse = torch.empty(max(n, 1), dtype=torch.int32, device=dev)[:n]
    n_unique = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.sg_unique_inverse_ws_bytes(n)
    if ws_bytes == 0:
        check(2, "sg_unique_inverse_ws_bytes")
    ws = _bytes(ws_bytes, dev)
    check(lib.sg_unique_inverse(_p(uniq), _p(inverse), _p(n_unique), _p(data), n, _p(ws), ws_bytes, _stream()),
          "sg_unique_inverse")
    return uniq[:int(n_unique.item())], inverse


def merge_nodes(node_ids_l):
    """mxgraph.graph.merge_nodes for a list of int32 device tensors: (uniq_node_ids, [indices per input])."""
    if isinstance(node_ids_l, torch.Tensor):
        return unique_inverse(node_ids_l)
    uniq, inv = unique_inverse(torch.cat([t.reshape(-1) for t in node_ids_l]))
    out, begin = [], 0
    for t in node_ids_l:
        out.append(inv[begin:begin + t.numel()])
        begin += t.numel()
    return uniq, out


__all__ = ["DeviceCSR", "unique_inverse", "merge_nodes"]
