"""``HeterGCNLayer`` / ``StackedHeterGCNLayers`` / ``LayerDictionary`` / ``InnerProductLayer``
with the reference's constructor keywords and call signatures
(mxgraph/layers/layers.py:8-39,42-208,210-222,224-385) on torch.nn.Module.

``heter_sage`` consumes the computing plan in the reference's own format
(``[prev_level_ids_dict, {src_key: [uniq_sel_inds, sel_node_idx, {dst_key: [end_points,
edge_values, ind_ptr, support]}]}]`` per depth, layers.py:303-336) but uploads every
``(src_key, dst_key)`` entry ONCE into a device-resident :class:`MultiLinkCSR` that is cached
on the plan object, instead of four ``nd.array`` copies per level per call (layers.py:366-377).
"""
import warnings

import numpy as np
import torch
import torch.nn as nn

from ..graph import MultiLinkCSR
from .aggregators import GCNAggregator, MultiLinkGCNAggregator
from .common import Dense, activation_code, get_activation


class LayerDictionary(nn.Module):
    """Mapping from hashable keys — typically ``(src_key, dst_key)`` tuples, which ``nn.ModuleDict`` cannot
    hold — to sub-modules.  Same surface as the reference container (mxgraph/layers/layers.py:8-39):
    ``len``, ``[]`` get/set, ``in``, ``keys()``."""

    def __init__(self, **kwargs):
        super().__init__()
        self._slot = {}                   # key -> position in the module list
        self._mods = nn.ModuleList()

    def __len__(self):
        return len(self._mods)

    def __contains__(self, key):
        return key in self._slot

    def keys(self):
        return self._slot.keys()

    def __getitem__(self, key):
        return self._mods[self._slot[key]]

    def __setitem__(self, key, layer):
        pos = self._slot.get(key)
        if pos is None:
            self._slot[key] = len(self._mods)
            self._mods.append(layer)
        else:
            warnings.warn(f"LayerDictionary: replacing the layer stored under {key!r}")
            self._mods[pos] = layer


def _per_key(value, keys):
    return value if isinstance(value, dict) else {k: value for k in keys}


class HeterGCNLayer(nn.Module):
    """One heterogeneous graph-convolution layer: for every source node type, aggregate each neighbour type
    with its own (multi-link) aggregator, combine the neighbour types (+ optionally a Dense of the node's own
    features), then Dense(out_units) + activation.  Constructor keywords and ``forward_single`` follow
    mxgraph/layers/layers.py:42-187; the Dense and its activation run as one tcgen05 GEMM."""

    def __init__(self, meta_graph, multi_link_structure, agg_units, out_units, source_keys=None, dropout_rate=0.0,
                 agg_ordinal_sharing=False, agg_accum="stack", agg_act="relu", layer_accum="stack",
                 accum_self=False, out_act=None, prefix=None, params=None):
        super().__init__()
        if layer_accum not in ("stack", "sum"):
            raise NotImplementedError(layer_accum)
        self._meta_graph = meta_graph
        self._source_keys = list(meta_graph.keys() if source_keys is None else source_keys)
        self._layer_accum, self._accum_self = layer_accum, accum_self
        self._out_act = get_activation(out_act)
        self.dropout = nn.Dropout(dropout_rate)
        units_of = _per_key(agg_units, meta_graph)
        out_of = _per_key(out_units, self._source_keys)

        self._aggregators = LayerDictionary()
        for src in self._source_keys:
            for dst in meta_graph[src]:
                n_links = multi_link_structure[(src, dst)]
                if n_links is None:
                    agg = GCNAggregator(units=units_of[src], act=agg_act, dropout_rate=dropout_rate)
                else:
                    agg = MultiLinkGCNAggregator(units=units_of[src], num_links=n_links, act=agg_act,
                                                 dropout_rate=dropout_rate, ordinal_sharing=agg_ordinal_sharing,
                                                 accum=agg_accum)
                self._aggregators[(src, dst)] = agg
        self._out_fcs = LayerDictionary()
        self._self_fcs = LayerDictionary() if accum_self else None
        for key, units in out_of.items():
            if units is None:
                continue
            self._out_fcs[key] = Dense(units)
            if accum_self:
                self._self_fcs[key] = nn.Sequential(nn.Dropout(dropout_rate), Dense(units), nn.Dropout(dropout_rate))

    @property
    def aggregators(self):
        return self._aggregators

    def _aggregate(self, key, dst_key, entry):
        feats, end_points, _edge_values, indptr, support = entry
        agg = self._aggregators[(key, dst_key)]
        if isinstance(end_points, MultiLinkCSR):       # device-resident plan entry: indptr / support live inside it
            inner = agg._agg if isinstance(agg, GCNAggregator) else agg
            return inner(feats, end_points)
        return agg(feats, end_points, indptr, support) if agg.use_support else agg(feats, end_points, indptr)

    def forward_single(self, key, base_feas, neighbor_data):
        """``neighbor_data[dst_key] = (features, end_points, edge_values, indptr, support)`` as in
        layers.py:147-187; ``end_points`` may be a prebuilt :class:`MultiLinkCSR`."""
        parts = [self.dropout(self._aggregate(key, dst, neighbor_data[dst])) for dst in self._meta_graph[key]]
        if self._accum_self:
            parts.append(self._self_fcs[key](base_feas))
        if len(parts) == 1:
            merged = parts[0]
        elif self._layer_accum == "stack":
            merged = torch.cat(parts, dim=1)
        else:
            merged = torch.stack(parts, dim=0).sum(dim=0)
        code = activation_code(self._out_act)
        if code is None:                               # activation the GEMM epilogue cannot carry
            return self._out_act(self._out_fcs[key](merged))
        return self._out_fcs[key](merged, act={0: None, 1: "leaky", 2: "relu"}[code])

    def forward(self, base_feas, neighbor_data):
        missing = [k for k in base_feas if k not in neighbor_data]
        if missing:
            raise KeyError(f"no neighbour data for node types {missing}")
        return {k: self.forward_single(k, feas, neighbor_data[k]) for k, feas in base_feas.items()}


class InnerProductLayer(nn.Module):
    """sum_d data1 * data2 (optionally after a shared Dense(mid_units)) — layers.py:210-222."""

    def __init__(self, mid_units=None, **kwargs):
        super().__init__()
        self._mid_map = Dense(mid_units) if mid_units is not None else None

    def forward(self, data1, data2):
        from ..decoder import inner_product
        if self._mid_map is not None:
            data1, data2 = self._mid_map(data1), self._mid_map(data2)
        return inner_product(data1, data2)


def _take_rows(x, idx):
    """mx.nd.take(x, idx) on axis 0 (row gather kernel; differentiable)."""
    from ..decoder import take_rows
    if not isinstance(idx, torch.Tensor):
        idx = torch.from_numpy(np.ascontiguousarray(np.asarray(idx), dtype=np.int32))
    return take_rows(x, idx.to(x.device, torch.int32))


class StackedHeterGCNLayers(nn.Module):
    """A stack of HeterGCNLayers (or one layer applied ``recurrent_layer_num`` times), the multi-hop plan
    that feeds it and its execution — surface of mxgraph/layers/layers.py:224-385."""

    def __init__(self, recurrent_layer_num=None, **kwargs):
        super().__init__()
        self._recurrent_layer_num = recurrent_layer_num
        self._blocks = nn.ModuleList()

    def __len__(self):
        if self._recurrent_layer_num is None:
            return len(self._blocks)
        return self._recurrent_layer_num if len(self._blocks) else 0

    def __getitem__(self, depth):
        if self._recurrent_layer_num is None:
            return self._blocks[depth]
        if not 0 <= depth < self._recurrent_layer_num:
            raise KeyError(f"{depth} is out of range. Layer number={len(self)}")
        return self._blocks[0]

    def add(self, *blocks):
        if self._recurrent_layer_num is not None and len(self._blocks) + len(blocks) > 1:
            raise ValueError("a recurrent stack holds exactly one block")
        for block in blocks:
            if not isinstance(block, HeterGCNLayer):
                raise TypeError("StackedHeterGCNLayers.add expects HeterGCNLayer instances")
            self._blocks.append(block)

    # ---- plan construction (host side; the device pieces are stargcn_b200.sampler) ----
    def _sample_depth(self, graph, depth, selected, fanout, symm):
        """Sample the neighbourhoods of one depth.  Returns the per-source entries (end points still unset)
        and, per node type, every id array that has to be mapped to a local row index."""
        entries, pending = {}, {}
        for src, ids in selected.items():
            neigh = {}
            for dst in graph.meta_graph[src]:
                multi = self[depth].aggregators[(src, dst)].use_multi_link
                ep_ids, values, indptr, support = graph[src, dst].sample_neighbors(
                    src_ids=ids, symm=symm, use_multi_link=multi, num_neighbors=fanout[(src, dst)])
                neigh[dst] = [None, values, indptr, support]
                pending.setdefault(dst, []).append((src, ep_ids))
            entries[src] = neigh
        return entries, pending

    def gen_plan(self, graph, sel_node_ids_dict, graph_sampler_args=None, symm=True):
        """Multi-hop computing plan, top depth first.  Output format of layers.py:260-337:
        ``(required_ids_of_depth_0, [[ids_dict, {src: [row_inds, restore_idx, {dst: [end_points, edge_values,
        ind_ptr, support]}]}] per depth])`` with end points as LOCAL row indices into the previous depth's
        merged node list.  (The reference's loop body still holds a debugging ``print`` / ``input()`` pair,
        :319-320, that blocks on stdin; nothing of the sort here.)"""
        from ..hetergraph import merge_nodes, unordered_unique

        n_depth = len(self)
        plan = [None] * n_depth
        selected = dict(sel_node_ids_dict)
        for depth in reversed(range(n_depth)):
            restore = {}
            if depth == n_depth - 1:          # only the outermost request may contain duplicates
                for key, ids in list(selected.items()):
                    selected[key], restore[key] = unordered_unique(ids, return_inverse=True)
            entries, pending = self._sample_depth(graph, depth, selected, graph_sampler_args, symm)

            merged_ids, args = {}, {src: [None, restore.get(src), entries[src]] for src in selected}
            for key in set(pending) | set(selected):
                arrays, owners = [], []       # owners[k] says where the k-th inverse array goes
                for src, ep_ids in pending.get(key, []):
                    if isinstance(ep_ids, np.ndarray):
                        arrays.append(ep_ids)
                        owners.append((src, None))
                    else:                     # multi-link: one id array per rating level
                        arrays.extend(ep_ids)
                        owners.extend((src, lvl) for lvl in range(len(ep_ids)))
                if key in selected:
                    arrays.append(selected[key])
                    owners.append((None, None))
                merged_ids[key], inverse = merge_nodes(arrays)
                for (src, lvl), inv in zip(owners, inverse):
                    if src is None:
                        args[key][0] = inv
                    elif lvl is None:
                        entries[src][key][0] = inv
                    else:
                        if entries[src][key][0] is None:
                            entries[src][key][0] = []
                        entries[src][key][0].append(inv)
            plan[depth] = [merged_ids, args]
            selected = merged_ids
        return plan[0][0], plan

    @staticmethod
    def _device_entry(agg_info, dst_key, n_nb, device):
        """Upload one (src, dst) plan entry once and keep the device form in the entry's fifth slot."""
        entry = agg_info[dst_key]
        cached = entry[4] if len(entry) > 4 else None
        if isinstance(cached, MultiLinkCSR) and cached.device == device:
            return cached
        end_points, _values, ind_ptr, support = entry[:4]
        if not isinstance(end_points, (list, tuple)):
            end_points, ind_ptr, support = [end_points], [ind_ptr], [support]
        csr = MultiLinkCSR(end_points, ind_ptr, support, n_nb, device=device)
        if isinstance(entry, list):
            if len(entry) > 4:
                entry[4] = csr
            else:
                entry.append(csr)
        return csr

    def heter_sage(self, input_dict, computing_plan):
        """Execute the plan (layers.py:339-385): depth by depth, every node type aggregates from the previous
        depth's features; the last depth is restored to the caller's (possibly repeated) order."""
        device = next(iter(input_dict.values())).device
        feats, last = input_dict, len(self) - 1
        for depth in range(len(self)):
            _ids, args = computing_plan[depth]
            out = {}
            for src, (row_inds, restore_idx, agg_info) in args.items():
                own = _take_rows(feats[src], row_inds)
                neigh = {dst: (feats[dst], self._device_entry(agg_info, dst, feats[dst].shape[0], device), None, None, None)
                         for dst in agg_info}
                h = self[depth].forward_single(key=src, base_feas=own, neighbor_data=neigh)
                out[src] = _take_rows(h, restore_idx) if depth == last and restore_idx is not None else h
            feats = out
        return feats
