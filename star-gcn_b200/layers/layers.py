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
    """Key -> layer mapping; keys may be tuples such as (src_key, dst_key) (layers.py:8-39)."""

    def __init__(self, **kwargs):
        super().__init__()
        self._key2idx = dict()
        self._layers = nn.ModuleList()
        self._nlayers = 0

    def __len__(self):
        return len(self._layers)

    def __setitem__(self, key, layer):
        if key in self._key2idx:
            warnings.warn("Duplicate Key. Need to test the code!")
            self._layers[self._key2idx[key]] = layer
        else:
            self._layers.append(layer)
            self._key2idx[key] = self._nlayers
            self._nlayers += 1

    def __getitem__(self, key):
        return self._layers[self._key2idx[key]]

    def __contains__(self, key):
        return key in self._key2idx

    def keys(self):
        return self._key2idx.keys()


class HeterGCNLayer(nn.Module):
    def __init__(self, meta_graph, multi_link_structure, agg_units, out_units, source_keys=None, dropout_rate=0.0,
                 agg_ordinal_sharing=False, agg_accum="stack", agg_act="relu", layer_accum="stack",
                 accum_self=False, out_act=None, prefix=None, params=None):
        super().__init__()
        self._meta_graph = meta_graph
        if source_keys is None:
            source_keys = meta_graph.keys()
        self._source_keys = list(source_keys)
        if not isinstance(out_units, dict):
            out_units = {k: out_units for k in self._source_keys}
        if not isinstance(agg_units, dict):
            agg_units = {k: agg_units for k in meta_graph}
        self._layer_accum = layer_accum
        self._accum_self = accum_self
        self._out_act = get_activation(out_act)
        self.dropout = nn.Dropout(dropout_rate)  # dropout before feeding the out layer (layers.py:91)
        self._aggregators = LayerDictionary()
        for src_key in self._source_keys:
            for dst_key in meta_graph[src_key]:
                if multi_link_structure[(src_key, dst_key)] is None:
                    self._aggregators[(src_key, dst_key)] = GCNAggregator(
                        units=agg_units[src_key], act=agg_act, dropout_rate=dropout_rate)
                else:
                    self._aggregators[(src_key, dst_key)] = MultiLinkGCNAggregator(
                        units=agg_units[src_key], num_links=multi_link_structure[(src_key, dst_key)], act=agg_act,
                        dropout_rate=dropout_rate, ordinal_sharing=agg_ordinal_sharing, accum=agg_accum)
        self._out_fcs = LayerDictionary()
        for key, ele_units in out_units.items():
            if ele_units is not None:
                self._out_fcs[key] = Dense(ele_units)
        if self._accum_self:
            self._self_fcs = LayerDictionary()
            for key, ele_units in out_units.items():
                if ele_units is not None:
                    self._self_fcs[key] = nn.Sequential(nn.Dropout(dropout_rate), Dense(ele_units),
                                                        nn.Dropout(dropout_rate))

    @property
    def aggregators(self):
        return self._aggregators

    def forward_single(self, key, base_feas, neighbor_data):
        """neighbor_data: {dst_key: (feas, end_points, edge_values, indptr, support)} (layers.py:147-187).
        ``end_points`` may be a prebuilt MultiLinkCSR (then indptr/support are ignored)."""
        out_l = []
        for dst_key in self._meta_graph[key]:
            neighbor_feas, end_points, edge_values, indptr, support = neighbor_data[dst_key]
            agg = self._aggregators[(key, dst_key)]
            if isinstance(agg, GCNAggregator) and isinstance(end_points, MultiLinkCSR):
                out = agg._agg(neighbor_feas, end_points)
            elif agg.use_support:
                out = agg(neighbor_feas, end_points, indptr, support)
            else:
                out = agg(neighbor_feas, end_points, indptr)
            out_l.append(self.dropout(out))
        if self._accum_self:
            out_l.append(self._self_fcs[key](base_feas))
        if len(out_l) == 1:
            out = out_l[0]
        elif self._layer_accum == "stack":
            out = torch.cat(out_l, dim=1)
        elif self._layer_accum == "sum":
            out = torch.stack(out_l, dim=0).sum(dim=0)
        else:
            raise NotImplementedError
        code = activation_code(self._out_act)
        if code is not None:  # activation rides in the GEMM epilogue
            return self._out_fcs[key](out, act={0: None, 1: "leaky", 2: "relu"}[code])
        return self._out_act(self._out_fcs[key](out))

    def forward(self, base_feas, neighbor_data):
        out = {}
        for key, ele_feas in base_feas.items():
            assert key in neighbor_data
            out[key] = self.forward_single(key, ele_feas, neighbor_data[key])
        return out


class InnerProductLayer(nn.Module):
    def __init__(self, mid_units=None, **kwargs):
        super().__init__()
        self._mid_units = mid_units
        if self._mid_units is not None:
            self._mid_map = Dense(mid_units)

    def forward(self, data1, data2):
        if self._mid_units is not None:
            data1 = self._mid_map(data1)
            data2 = self._mid_map(data2)
        from ..decoder import inner_product
        return inner_product(data1, data2)


def _take_rows(x, idx):
    """mx.nd.take(x, idx) on axis 0 with int32 indices."""
    if not isinstance(idx, torch.Tensor):
        idx = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
    return x.index_select(0, idx.to(x.device, torch.int64))


class StackedHeterGCNLayers(nn.Module):
    """Stack multiple HeterGCNLayers (layers.py:224-258)."""

    def __init__(self, recurrent_layer_num=None, **kwargs):
        super().__init__()
        self._recurrent_layer_num = recurrent_layer_num
        self._blocks = nn.ModuleList()

    def __len__(self):
        if self._recurrent_layer_num is None:
            return len(self._blocks)
        return 0 if len(self._blocks) == 0 else self._recurrent_layer_num

    def __getitem__(self, key):
        if self._recurrent_layer_num is not None:
            if key < self._recurrent_layer_num:
                return self._blocks[0]
            raise KeyError("{} is out of range. Layer number={}".format(key, len(self)))
        return self._blocks[key]

    def add(self, *blocks):
        if self._recurrent_layer_num is not None:
            if len(self._blocks) == 1:
                raise ValueError("Cannot add more blocks if `use_recurrent` flag is turned on!")
            if len(blocks) > 1:
                raise ValueError("Can only add a single block if `use_recurrent` flag is turned on!")
        for block in blocks:
            assert isinstance(block, HeterGCNLayer)
            self._blocks.append(block)

    def gen_plan(self, graph, sel_node_ids_dict, graph_sampler_args=None, symm=True):
        """Host-side multi-hop plan, same output structure as layers.py:260-337 (minus the stray
        ``print``/``input()`` debugging lines at :319-320 that block the reference on stdin)."""
        from ..hetergraph import merge_nodes, unordered_unique

        computing_plan = [None for _ in range(len(self))]
        for depth in range(len(self) - 1, -1, -1):
            prev_level_ids_dict, agg_args_dict = dict(), dict()
            all_neighbor_ids_dict, all_src_ids_dict = dict(), dict()
            for src_key, sel_node_ids in sel_node_ids_dict.items():
                if depth == len(self) - 1:
                    uniq_sel_node_ids, sel_node_idx = unordered_unique(sel_node_ids, return_inverse=True)
                else:
                    uniq_sel_node_ids, sel_node_idx = sel_node_ids, None
                agg_args_dict[src_key] = [uniq_sel_node_ids, sel_node_idx, dict()]
                all_src_ids_dict[src_key] = uniq_sel_node_ids
                for dst_key in graph.meta_graph[src_key]:
                    use_multi_link = self[depth].aggregators[(src_key, dst_key)].use_multi_link
                    end_points_ids, edge_values, ind_ptr, support = graph[src_key, dst_key].sample_neighbors(
                        src_ids=uniq_sel_node_ids, symm=symm, use_multi_link=use_multi_link,
                        num_neighbors=graph_sampler_args[(src_key, dst_key)])
                    agg_args_dict[src_key][2][dst_key] = [None, edge_values, ind_ptr, support]
                    all_neighbor_ids_dict.setdefault(dst_key, dict())[src_key] = end_points_ids
            for key in set(all_neighbor_ids_dict.keys()) | set(all_src_ids_dict.keys()):
                node_ids_l = []
                if key in all_neighbor_ids_dict:
                    for _, end_points in all_neighbor_ids_dict[key].items():
                        if isinstance(end_points, np.ndarray):
                            node_ids_l.append(end_points)
                        else:
                            node_ids_l.extend(end_points)
                if key in all_src_ids_dict:
                    node_ids_l.append(all_src_ids_dict[key])
                uniq_node_ids, node_inds_l = merge_nodes(node_ids_l)
                prev_level_ids_dict[key] = uniq_node_ids
                curr = 0
                if key in all_neighbor_ids_dict:
                    for src_key, end_points in all_neighbor_ids_dict[key].items():
                        if isinstance(end_points, np.ndarray):
                            agg_args_dict[src_key][2][key][0] = node_inds_l[curr]
                            curr += 1
                        else:
                            agg_args_dict[src_key][2][key][0] = node_inds_l[curr:(curr + len(end_points))]
                            curr += len(end_points)
                if key in all_src_ids_dict:
                    agg_args_dict[key][0] = node_inds_l[curr]
            computing_plan[depth] = [prev_level_ids_dict, agg_args_dict]
            sel_node_ids_dict = prev_level_ids_dict
        return computing_plan[0][0], computing_plan

    @staticmethod
    def _device_entry(agg_info, dst_key, n_nb, device):
        """Upload one (src,dst) plan entry once; cached in the plan's own list (5th slot)."""
        entry = agg_info[dst_key]
        if len(entry) > 4 and isinstance(entry[4], MultiLinkCSR) and entry[4].device == device:
            return entry[4]
        end_points, edge_values, ind_ptr, support = entry[:4]
        if isinstance(end_points, (list, tuple)):
            csr = MultiLinkCSR(end_points, ind_ptr, support, n_nb, device=device)
        else:
            csr = MultiLinkCSR([end_points], [ind_ptr], [support], n_nb, device=device)
        if isinstance(entry, list):
            if len(entry) > 4:
                entry[4] = csr
            else:
                entry.append(csr)
        return csr

    def heter_sage(self, input_dict, computing_plan):
        """Run the stacked layers over the plan (layers.py:339-385)."""
        device = next(iter(input_dict.values())).device
        ret = dict()
        for depth in range(len(self)):
            ret = dict()
            prev_level_ids_dict, agg_args_dict = computing_plan[depth]
            for src_key in agg_args_dict:
                uniq_sel_node_inds, sel_node_idx, agg_info_dict = agg_args_dict[src_key]
                nd_src_feas = _take_rows(input_dict[src_key], uniq_sel_node_inds)
                neighbor_data = {}
                for dst_key in agg_info_dict:
                    nd_neighbor_feas = input_dict[dst_key]
                    csr = self._device_entry(agg_info_dict, dst_key, nd_neighbor_feas.shape[0], device)
                    neighbor_data[dst_key] = (nd_neighbor_feas, csr, None, None, None)
                ret[src_key] = self[depth].forward_single(key=src_key, base_feas=nd_src_feas,
                                                          neighbor_data=neighbor_data)
                if depth == len(self) - 1:
                    ret[src_key] = _take_rows(ret[src_key], sel_node_idx)
            input_dict = ret
        return ret
