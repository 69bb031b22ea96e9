"""Device-resident heterogeneous graph and the computing plan built on it.

``DeviceHeterGraph`` is the device form of the slice of ``mxgraph.graph.HeterGraph`` that
``StackedHeterGCNLayers.gen_plan`` touches (mxgraph/layers/layers.py:260-337): ``meta_graph`` and, per
``(src_key, dst_key)``, a CSR matrix with the node ids of its rows and columns (mxgraph/graph.py:261-316).
``gen_plan`` / ``merge_node_ids_dict`` below produce the SAME plan as the host versions in
``layers/layers.py`` / ``hetergraph.py`` — identical node lists (ids in order of first appearance, the
serial ``unique_inverse`` of GraphSampler/graph_sampler.h:510-534), identical local indices, identical
per-level CSRs — but every array stays on the device: neighbour lists are sampled, split by rating level
and mapped to node ids by the kernels of csrc/sampler.cu, ids are merged by the device sort-unique-inverse
(csrc/plan.cu), and each ``(src, dst)`` entry comes out as the relation-major :class:`MultiLinkCSR` the
fused aggregation consumes.  What crosses to the host is a handful of COUNTS per depth (number of distinct
nodes, number of sampled edges) — the sizes of the next allocations; no index or feature data does.
"""
import numpy as np
import torch

from .graph import MultiLinkCSR
from .sampler import DeviceCSR, unique_inverse


def _i32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1), dtype=np.int32)).to(device)


class DeviceCSRMat:
    """One ``(src_key, dst_key)`` matrix: a :class:`DeviceCSR` plus the node ids of its rows / columns and the
    id -> row lookup (the reference keeps a ``reverse_row_map`` dict, graph.py:300-316; here a dense int32 table)."""

    def __init__(self, csr, row_ids, col_ids):
        dev = csr.device
        self.csr = csr
        self.row_ids, self.col_ids = _i32(row_ids, dev), _i32(col_ids, dev)
        if self.row_ids.numel() != csr.n_rows or self.col_ids.numel() != csr.n_cols:
            raise ValueError("row_ids / col_ids must name every row / column of the matrix")
        n_ids = int(self.row_ids.max().item()) + 1 if csr.n_rows else 0
        self._row_of = torch.full((max(n_ids, 1),), -1, dtype=torch.int32, device=dev)
        self._row_of[self.row_ids.long()] = torch.arange(csr.n_rows, dtype=torch.int32, device=dev)
        n_cids = int(self.col_ids.max().item()) + 1 if csr.n_cols else 0
        self._col_of = torch.full((max(n_cids, 1),), -1, dtype=torch.int32, device=dev)
        self._col_of[self.col_ids.long()] = torch.arange(csr.n_cols, dtype=torch.int32, device=dev)

    def rows_of(self, node_ids):
        """Row index of every node id (ids come from this graph's own lists, so they are always present)."""
        return self._row_of[node_ids.long()]

    def cols_of(self, node_ids):
        """Column index of every (destination-type) node id."""
        return self._col_of[node_ids.long()]

    def sample_neighbors(self, src_ids, num_neighbors=-1, seed=None):
        """CSRMat.sample_neighbors(src_ids, use_multi_link=True) (graph.py:677-748) on the device.
        Returns (end-point NODE IDS concatenated level-major, concatenated indptr, concatenated support, n_sel):
        the concatenation of the reference's per-level ``end_points_l`` / ``ind_ptr_l`` / ``support_l`` lists."""
        rows = self.rows_of(src_ids)
        sampled, dst_indptr, n_sel = self.csr.sample_positions(rows, num_neighbors, seed)
        cat_indptr, ep_cat, sup_cat, _, _ = self.csr.split(sampled, dst_indptr, n_sel)
        return self.col_ids[ep_cat.long()], cat_indptr, sup_cat, n_sel


class DeviceHeterGraph:
    def __init__(self, meta_graph, mats):
        self.meta_graph = meta_graph
        self._mats = mats

    def __getitem__(self, key):
        return self._mats[key]

    @property
    def device(self):
        return next(iter(self._mats.values())).csr.device

    @classmethod
    def from_synth(cls, g, user="user", item="item", device="cuda"):
        """Graph over a ``synth.make_bipartite`` dict (node ids = 0..N-1 on each side)."""
        uid, iid = np.arange(g["n_user"], dtype=np.int32), np.arange(g["n_item"], dtype=np.int32)
        mats = {}
        for key, c, rid, cid in (((user, item), g["u2i"], uid, iid), ((item, user), g["i2u"], iid, uid)):
            csr = DeviceCSR(c["indptr"], c["cols"], c["vals"], g["levels"], len(cid), support=c["support"], device=device)
            mats[key] = DeviceCSRMat(csr, rid, cid)
        return cls({user: {item: "rating"}, item: {user: "rev_rating"}}, mats)


def merge_nodes(arrays):
    """mxgraph.graph.merge_nodes on the device: (distinct ids in first-appearance order over the concatenation,
    [index of every element of every input into that list])."""
    sizes = [int(a.numel()) for a in arrays]
    if sum(sizes) == 0:
        dev = arrays[0].device
        return torch.zeros(0, dtype=torch.int32, device=dev), [torch.zeros(0, dtype=torch.int32, device=dev) for _ in arrays]
    flat = torch.cat([a.reshape(-1) for a in arrays]) if len(arrays) > 1 else arrays[0].reshape(-1)
    uniq, inv = unique_inverse(flat.contiguous())
    out, begin = [], 0
    for n in sizes:
        out.append(inv[begin:begin + n])
        begin += n
    return uniq, out


def merge_node_ids_dict(dicts, device):
    """Device version of ``hetergraph.merge_node_ids_dict`` (mxgraph/graph.py:166-219 for plain requests)."""
    per_type = {}
    for d in dicts:
        for key, ids in d.items():
            per_type.setdefault(key, []).append(_i32(ids, device))
    merged, inverse = {}, {}
    for key, arrays in per_type.items():
        merged[key], inverse[key] = merge_nodes(arrays)
    cursor = {key: 0 for key in per_type}
    out = []
    for d in dicts:
        idx = {}
        for key in d:
            idx[key] = inverse[key][cursor[key]]
            cursor[key] += 1
        out.append(idx)
    return merged, out


def gen_plan(stack, graph, sel_node_ids_dict, graph_sampler_args=None, symm=True, seed=None):
    """``StackedHeterGCNLayers.gen_plan`` (layers.py:260-337) on a :class:`DeviceHeterGraph`.  Same output format:
    ``(required_ids_of_depth_0, [[ids_dict, {src: [row_inds, restore_idx, {dst: entry}]}] per depth])`` where
    every id / index array is an int32 device tensor and ``entry = [MultiLinkCSR, None, None, None, MultiLinkCSR]``
    (the device plan entry ``heter_sage`` consumes directly; ``MultiLinkCSR.to_lists()`` gives the reference's
    per-level ``[end_points_l, ind_ptr_l, support_l]`` back for inspection)."""
    dev = graph.device
    n_depth = len(stack)
    plan = [None] * n_depth
    selected = {k: _i32(v, dev) for k, v in sel_node_ids_dict.items()}
    for depth in reversed(range(n_depth)):
        restore = {}
        if depth == n_depth - 1:          # only the outermost request may contain duplicates
            for key, ids in list(selected.items()):
                selected[key], restore[key] = unique_inverse(ids)
        entries, pending = {}, {}
        for src, ids in selected.items():
            neigh = {}
            for dst in graph.meta_graph[src]:
                if not stack[depth].aggregators[(src, dst)].use_multi_link:
                    raise NotImplementedError("device gen_plan covers the multi-link (rating level) aggregators")
                fan = -1 if graph_sampler_args is None else graph_sampler_args[(src, dst)]
                ep_ids, cat_indptr, sup_cat, n_sel = graph[src, dst].sample_neighbors(ids, fan, seed)
                neigh[dst] = [cat_indptr, sup_cat, n_sel, graph[src, dst].csr.R]
                pending.setdefault(dst, []).append((src, ep_ids))
            entries[src] = neigh
        merged_ids, args = {}, {src: [None, restore.get(src), {}] for src in selected}
        for key in list(dict.fromkeys(list(pending) + list(selected))):
            arrays, owners = [], []
            for src, ep_ids in pending.get(key, []):
                arrays.append(ep_ids)
                owners.append(src)
            if key in selected:
                arrays.append(selected[key])
                owners.append(None)
            merged_ids[key], inverse = merge_nodes(arrays)
            for src, inv in zip(owners, inverse):
                if src is None:
                    args[key][0] = inv
                else:
                    cat_indptr, sup_cat, n_sel, R = entries[src][key]
                    # the inverse is a slice of the merged array: give the plan entry its own aligned storage
                    csr = MultiLinkCSR.from_device(inv.clone(), sup_cat, cat_indptr, R, n_sel, int(merged_ids[key].numel()))
                    args[src][2][key] = [csr, None, None, None, csr]
        plan[depth] = [merged_ids, args]
        selected = merged_ids
    return plan[0][0], plan


__all__ = ["DeviceCSRMat", "DeviceHeterGraph", "gen_plan", "merge_nodes", "merge_node_ids_dict"]
