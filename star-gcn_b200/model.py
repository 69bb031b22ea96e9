"""STAR-GCN encoder–decoder stack assembled from the device kernels: the model-level call surface of
``Net`` in the reference's experiment script (experiments/STAR-GCN.py:167-461), for the shipped
configuration family (USE_EMBED, no side features, USE_DAE, non-recurrent blocks):

    embed_layers[key]            nn.Embedding(N_key, EMBED.UNITS), U(-0.1, 0.1)                 :173-181
    encoders[b]                  StackedHeterGCNLayers with one HeterGCNLayer(AGG -> OUT)       :194-222
    embed_maps[b][key]           Dense -> act -> Dense back to EMBED.UNITS                      :226-246
    rating_{user,item}_projs[b]  Dense(GEN_RATING.MID_MAP)                                      :249-259
    gen_ratings                  InnerProductLayer                                              :261

``forward`` builds the per-block plans top-down (every block predicts the ratings of the batch pairs and
reconstructs the embeddings of the recon nodes; block b also produces the inputs block b+1 needs) and runs
them bottom-up, as :340-461 does.  ``loss`` is the training objective of :611-628.
"""
import numpy as np
import torch
import torch.nn as nn

from . import decoder, devgraph
from .hetergraph import merge_node_ids_dict
from .layers import HeterGCNLayer, InnerProductLayer, LayerDictionary, StackedHeterGCNLayers
from .layers.common import Dense


class EmbedTable(nn.Module):
    def __init__(self, num_nodes, units):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(num_nodes, units).uniform_(-0.1, 0.1))


def _ids(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.int32)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.int32)).to(device)


class StarGCN(nn.Module):
    def __init__(self, meta_graph, multi_link_structure, num_nodes, name_user, name_item, embed_units=64,
                 agg_units=250, out_units=75, n_blocks=2, mid_map=64, agg_accum="sum", act="leaky", dropout=0.0):
        super().__init__()
        self._name_user, self._name_item, self._n_blocks = name_user, name_item, n_blocks
        self.embed_layers = LayerDictionary()
        for key, n in num_nodes.items():
            self.embed_layers[key] = EmbedTable(n, embed_units)
        self.encoders, self.embed_maps = nn.ModuleList(), nn.ModuleList()
        self.rating_user_projs, self.rating_item_projs = nn.ModuleList(), nn.ModuleList()
        for _ in range(n_blocks):
            enc = StackedHeterGCNLayers()
            enc.add(HeterGCNLayer(meta_graph=meta_graph, multi_link_structure=multi_link_structure, dropout_rate=dropout,
                                  agg_units=agg_units, out_units=out_units, agg_accum=agg_accum, agg_act=act, out_act=act))
            self.encoders.append(enc)
            maps = LayerDictionary()
            for key in meta_graph:
                maps[key] = decoder.EmbedMap(embed_units, act=act)
            self.embed_maps.append(maps)
            self.rating_user_projs.append(Dense(mid_map))
            self.rating_item_projs.append(Dense(mid_map))
        self.gen_ratings = InnerProductLayer()
        self.last_plans = None

    @property
    def device(self):
        return next(self.parameters()).device

    def get_embed(self, node_ids_dict, embed_noise_dict=None, use_mask=True):
        """{key: ids} -> {key: embeddings}; with ``use_mask`` ids go through the noise table first
        (-1 = zero vector, i = row i) — Net.get_embed :264-300."""
        dev = self.device
        out = {}
        for key, ids in node_ids_dict.items():
            noise = _ids(embed_noise_dict[key], dev) if use_mask else None
            out[key] = decoder.get_embed(self.embed_layers[key].weight, _ids(ids, dev), noise, use_mask=use_mask)
        return out

    def forward(self, graph, rating_node_pairs=None, embed_noise_dict=None, recon_node_ids_dict=None,
                graph_sampler_args=None, symm=True):
        """Returns (pred_ratings per block, pred_embeddings per block, gt_embeddings)."""
        if rating_node_pairs is None and recon_node_ids_dict is None:
            raise NotImplementedError("need rating pairs, recon nodes, or both")
        dev = self.device
        gt = self.get_embed(recon_node_ids_dict, use_mask=False) if recon_node_ids_dict is not None else {}
        # a DeviceHeterGraph keeps the whole plan construction on the device (devgraph.gen_plan): no index array
        # crosses PCIe in either direction; a host graph object goes through the numpy mirror of the reference
        on_device = isinstance(graph, devgraph.DeviceHeterGraph)

        # ---- plans, last block first: what a block must output = rating nodes + recon nodes + next block's inputs ----
        plans, lookups, needed = [None] * self._n_blocks, [None] * self._n_blocks, {}
        for b in reversed(range(self._n_blocks)):
            requests, names = [], []
            if rating_node_pairs is not None:
                requests.append({self._name_user: rating_node_pairs[0], self._name_item: rating_node_pairs[1]})
                names.append("rating")
            if recon_node_ids_dict is not None:
                requests.append(recon_node_ids_dict)
                names.append("recon")
            requests.append(needed)
            names.append("next")
            if on_device:
                selected, idx = devgraph.merge_node_ids_dict(requests, dev)
                lookups[b] = dict(zip(names, idx))
                needed, plans[b] = devgraph.gen_plan(self.encoders[b], graph, selected, graph_sampler_args, symm)
            else:
                selected, idx = merge_node_ids_dict(requests)
                lookups[b] = dict(zip(names, idx))
                needed, plans[b] = self.encoders[b].gen_plan(graph=graph, sel_node_ids_dict=selected,
                                                             graph_sampler_args=graph_sampler_args, symm=symm)
        self.last_plans = (plans, lookups, needed)

        # ---- execution, first block first ----
        feats = self.get_embed(needed, embed_noise_dict, use_mask=embed_noise_dict is not None)
        pred_ratings, pred_embeddings = [], []
        for b in range(self._n_blocks):
            h = self.encoders[b].heter_sage(feats, plans[b])
            look = lookups[b]
            if "rating" in look:
                u = self.rating_user_projs[b](decoder.take_rows(h[self._name_user], _ids(look["rating"][self._name_user], dev)))
                v = self.rating_item_projs[b](decoder.take_rows(h[self._name_item], _ids(look["rating"][self._name_item], dev)))
                pred_ratings.append(self.gen_ratings(u, v))
            if "recon" in look:
                pred_embeddings.append({key: self.embed_maps[b][key](h[key], _ids(idx, dev))
                                        for key, idx in look["recon"].items()})
            if b < self._n_blocks - 1:
                feats = {key: self.embed_maps[b][key](h[key], _ids(idx, dev)) for key, idx in look["next"].items()}
        return pred_ratings, pred_embeddings, gt

    @staticmethod
    def loss(pred_ratings, pred_embeddings, gt_embeddings, gt_ratings, rating_mean=0.0, rating_std=1.0, recon_lambda=0.1):
        """sum_b L2Loss(pred_b, (y - mean)/std).mean() + lambda * sum_b sum_key mean_n sum_d (gt - pred)^2   (:611-628)"""
        target = (gt_ratings - rating_mean) / rating_std
        total = None
        for p in pred_ratings:
            term = decoder.l2_loss(p, target)
            total = term if total is None else total + term
        for block in pred_embeddings:
            for key, pred in block.items():
                term = recon_lambda * decoder.recon_loss(gt_embeddings[key], pred)
                total = term if total is None else total + term
        return total


__all__ = ["StarGCN", "EmbedTable"]
