"""Host-side id bookkeeping used when a computing plan is built on the CPU.

``unordered_unique`` / ``merge_nodes`` give the same results as the serial (order-defining) variant of
the reference's ``unique_inverse`` (GraphSampler/graph_sampler.h:510-534, reached through
mxgraph/graph.py:63-163): distinct ids in order of FIRST appearance plus the index of every input
element into that list.  The device versions live in ``stargcn_b200.sampler`` (``unique_inverse``,
``merge_nodes``) and are bit-exact with these.
"""
import numpy as np


def unordered_unique(ids, return_inverse=False):
    ids = np.asarray(ids)
    if ids.size == 0:
        empty = ids.astype(np.int32)
        return (empty, np.zeros(0, np.int32)) if return_inverse else empty
    sorted_vals, first_pos, inv_sorted = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first_pos, kind="stable")          # sorted-unique slot -> rank by first appearance
    uniq = sorted_vals[order].astype(ids.dtype)
    if not return_inverse:
        return uniq
    rank = np.empty(order.size, np.int32)
    rank[order] = np.arange(order.size, dtype=np.int32)
    return uniq, rank[inv_sorted.reshape(-1)].astype(np.int32)


def merge_nodes(node_ids):
    """One array -> (uniq, inverse).  A list of arrays -> (uniq over the concatenation, [inverse per array])."""
    if isinstance(node_ids, np.ndarray):
        return unordered_unique(node_ids, return_inverse=True)
    sizes = [int(np.asarray(a).size) for a in node_ids]
    flat = np.concatenate([np.asarray(a).reshape(-1) for a in node_ids]) if sizes else np.zeros(0, np.int32)
    uniq, inv = unordered_unique(flat, return_inverse=True)
    cuts = np.cumsum([0] + sizes)
    return uniq, [inv[cuts[k]:cuts[k + 1]] for k in range(len(sizes))]


__all__ = ["unordered_unique", "merge_nodes"]


def merge_node_ids_dict(dicts):
    """Several ``{node_type: ids}`` requests -> (``{node_type: distinct ids in first-appearance order}``,
    one ``{node_type: index into those}`` per request) — what mxgraph.graph.merge_node_ids_dict
    (graph.py:166-219) returns for plain (non-edge) requests.  Requests are scanned in order, so the ids of
    an earlier request come first in the merged list."""
    per_type = {}
    for d in dicts:
        for key, ids in d.items():
            per_type.setdefault(key, []).append(np.asarray(ids).reshape(-1))
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


__all__ += ["merge_node_ids_dict"]
