"""``GCNAggregator`` / ``MultiLinkGCNAggregator`` with the reference's call surface
(mxgraph/layers/aggregators.py:21-163) on torch.nn.Module, computing through the fused
sm_100a aggregation kernels.

Reference data flow per rating level r (aggregators.py:133-150):
    H_r = FullyConnected(dropout(X), W_r, b_r)            (N_nb, U_r)
    out_r = seg_weighted_pool(H_r, support_r, end_points_r, indptr_r)
    out = act(concat_r out_r | sum_r out_r)
which gathers U_r = 250 floats per edge and launches 2R+2 operators.  This module computes
the same function aggregate-first:
    A[i, r, :] = sum_p support_r[p] * X[end_points_r[p], :]       one fused launch, D floats/edge
    s[i, r]    = sum_p support_r[p]
    out_r[i]   = A[i, r, :] W_r^T + s[i, r] * b_r                 one GEMM over K = R*D
Set ``reference_order=True`` to run the reference's operator order through the per-level
``seg_op.seg_weighted_pool`` kernels instead (used by the parity tests).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.lazy import LazyModuleMixin

from .. import seg_op
from ..graph import FUSED_DIMS, MultiLinkCSR, fused_agg_transform, multilink_aggregate, pack_w_ext
from .common import activation_code, get_activation, xavier_in_uniform_


class BaseAggregator(LazyModuleMixin, nn.Module):
    cls_to_become = None

    def initialize_parameters(self, *args, **kwargs):   # aggregators without deferred shapes
        pass

    @property
    def use_mulit_link(self):  # (sic) spelled as in the reference, aggregators.py:9
        raise NotImplementedError

    @property
    def use_support(self):
        raise NotImplementedError

    @property
    def use_edge_type(self):
        raise NotImplementedError


class MultiLinkGCNAggregator(BaseAggregator):
    def __init__(self, units, num_links, act=None, dropout_rate=0.0, ordinal_sharing=True, accum="stack",
                 in_units=None, reference_order=False, **kwargs):
        super().__init__()
        self._units = units
        self._num_links = num_links
        self._act = get_activation(act)
        self._ordinal_sharing = ordinal_sharing
        self._accum = accum
        if accum not in ("stack", "sum"):
            raise NotImplementedError(accum)
        if self._accum == "stack":
            assert units % num_links == 0, "units should be divisible by the num_links "
            self._units = self._units // num_links
        self.reference_order = reference_order
        self.grad_group = None     # torch.distributed group: all-reduce the weight gradient inside backward
        self.halo_plan = None      # dist.HaloPlan of mode 'peer': neighbor_data holds the rank's OWN rows and the
                                   # exchange runs inside the fused op over NVLink peer memory
        self.tensor_cores = True   # False: unfused path (generic gather kernel + Dense GEMM), kept for A/B runs
        self.dropout = nn.Dropout(dropout_rate)
        # parameters are named weight{i} / bias{i} as in the reference (aggregators.py:86-97)
        for i in range(num_links):
            self.register_parameter(f"weight{i}", nn.UninitializedParameter())
            self.register_parameter(f"bias{i}", nn.UninitializedParameter())
        self._plan_cache = {}
        if in_units is not None:
            self._materialize(in_units, None)

    @property
    def use_multi_link(self):
        return True

    @property
    def use_support(self):
        return True

    @property
    def use_edge_type(self):
        return False

    def initialize_parameters(self, neighbor_data, *args, **kwargs):
        """torch lazy-module protocol: shapes come from the first input (MXNet deferred init); state_dict() /
        load_state_dict() work before that (load materialises from the checkpoint's shapes)."""
        if self.has_uninitialized_params():
            self._materialize(neighbor_data.shape[-1], neighbor_data.device)

    def _materialize(self, in_units, device):
        for i in range(self._num_links):
            w, b = getattr(self, f"weight{i}"), getattr(self, f"bias{i}")
            if isinstance(w, nn.UninitializedParameter):
                w.materialize((self._units, in_units), device=device, dtype=torch.float32)
                xavier_in_uniform_(w, in_units)
            if isinstance(b, nn.UninitializedParameter):
                b.materialize((self._units,), device=device, dtype=torch.float32)
                with torch.no_grad():
                    b.zero_()

    def _effective_params(self):
        """Per-level (W_r, b_r); ordinal sharing accumulates levels (aggregators.py:134-140)."""
        ws, bs = [], []
        w = b = None
        for i in range(self._num_links):
            wi, bi = getattr(self, f"weight{i}"), getattr(self, f"bias{i}")
            if i > 0 and self._ordinal_sharing:
                w, b = w + wi, b + bi
            else:
                w, b = wi, bi
            ws.append(w)
            bs.append(b)
        return ws, bs

    def _plan(self, neighbor_rows, end_points_l, indptr_l, support_l):
        """Device plan of the three per-level lists.  Torch tensors are cached by identity AND version of every
        tensor (an in-place update of any of them rebuilds the plan); numpy inputs cannot be watched for in-place
        changes, so they are uploaded on every call — as the reference does (layers.py:366-377)."""
        if isinstance(end_points_l, MultiLinkCSR):
            return end_points_l
        tensors = list(end_points_l) + list(indptr_l) + list(support_l)
        if not all(isinstance(t, torch.Tensor) for t in tensors):
            return MultiLinkCSR(end_points_l, indptr_l, support_l, neighbor_rows)
        key = tuple(id(t) for t in tensors) + (int(neighbor_rows),)
        versions = [t._version for t in tensors]
        hit = self._plan_cache.get(key)
        if hit is not None and hit[0] == versions:
            return hit[1]
        csr = MultiLinkCSR(end_points_l, indptr_l, support_l, neighbor_rows)
        if len(self._plan_cache) >= 4:           # a plan holds ~40 B per edge of derived arrays: keep few
            self._plan_cache.pop(next(iter(self._plan_cache)))
        # keep the keyed tensors alive so ids cannot be recycled
        self._plan_cache[key] = (versions, csr, tensors)
        return csr

    def forward(self, neighbor_data, end_points_l, indptr_l=None, support_l=None):
        """neighbor_data (N_nb, D); the three lists as in aggregators.py:111-128 — or a prebuilt
        :class:`MultiLinkCSR` in place of ``end_points_l``."""
        if self.has_uninitialized_params():
            self._materialize(neighbor_data.shape[-1], neighbor_data.device)
        neighbor_data = self.dropout(neighbor_data)
        ws, bs = self._effective_params()
        if self.reference_order:
            return self._act(self._forward_reference_order(neighbor_data, end_points_l, indptr_l, support_l, ws, bs))
        n_rows = neighbor_data.shape[0] if self.halo_plan is None else self.halo_plan.n_ext
        csr = self._plan(n_rows, end_points_l, indptr_l, support_l)
        if csr.R != self._num_links:
            raise ValueError(f"plan has {csr.R} links, aggregator was built for {self._num_links}")
        D = neighbor_data.shape[1]
        code = activation_code(self._act)
        # One packed operand for every accumulation mode: 'sum' lays the level weights side by side,
        # w_ext = [W_0 | ... | W_{R-1} | b_0 ... b_{R-1}]  (U, R*D + R);  'stack' (aggregators.py:79-81,151-153 —
        # shipped by cfg/inductive_ml_100k_item_*.yml) is the same GEMM with a BLOCK-DIAGONAL operand: row block r
        # holds W_r in the columns of level r and b_r in bias column r, so out[:, r*U_r:(r+1)*U_r] = agg_r W_r^T +
        # wsum_r b_r without a batched library GEMM (the zero blocks cost R x the flops of a layer that is tiny).
        if self._accum == "sum" or self._num_links == 1:
            w_ext = pack_w_ext(ws, bs)
        else:
            w_ext = torch.cat([torch.block_diag(*ws), torch.block_diag(*[b.unsqueeze(1) for b in bs])], dim=1)
        slope = {0: 1.0, 1: 0.1, 2: 0.0}.get(code, 1.0)     # activations the epilogue cannot carry run afterwards
        if D in FUSED_DIMS and self.tensor_cores:
            # gather (1 launch) + tcgen05 3xTF32 GEMM with the activation in its epilogue
            out = fused_agg_transform(neighbor_data, w_ext, csr, slope, self.grad_group, self.halo_plan)
        else:
            # any other width: the generic gather kernel, then the same tensor-core GEMM through the Dense path
            from ..decoder import fused_dense
            agg, wsum = multilink_aggregate(neighbor_data, csr)       # (n_dst, R*D), (n_dst, R)
            out = fused_dense(torch.cat([agg, wsum], dim=1), w_ext, None,
                              {1.0: None, 0.1: "leaky", 0.0: "relu"}[slope])
        return out if code is not None else self._act(out)

    def _forward_reference_order(self, neighbor_data, end_points_l, indptr_l, support_l, ws, bs):
        if isinstance(end_points_l, MultiLinkCSR):
            raise ValueError("reference_order needs the per-level lists, not a MultiLinkCSR")
        out_l = []
        for i in range(self._num_links):
            feat = F.linear(neighbor_data, ws[i], bs[i])
            nnz = int(indptr_l[i][-1].item())
            out = seg_op.seg_weighted_pool(data=feat.unsqueeze(0), weights=support_l[i][:nnz].unsqueeze(0),
                                           indices=end_points_l[i][:nnz], indptr=indptr_l[i])
            out_l.append(out.reshape(-1, out.shape[-1]))
        if len(out_l) == 1:
            return out_l[0]
        if self._accum == "stack":
            return torch.cat(out_l, dim=1)
        return torch.stack(out_l, dim=0).sum(dim=0)


class GCNAggregator(BaseAggregator):
    def __init__(self, units, act=None, dropout_rate=0.0, **kwargs):
        super().__init__()
        self._agg = MultiLinkGCNAggregator(units=units, num_links=1, act=act, dropout_rate=dropout_rate, **kwargs)

    @property
    def use_multi_link(self):
        return False

    @property
    def use_support(self):
        return True

    @property
    def use_edge_type(self):
        return False

    def forward(self, neighbor_data, end_points, indptr, support):
        return self._agg(neighbor_data, [end_points], [indptr], [support])
